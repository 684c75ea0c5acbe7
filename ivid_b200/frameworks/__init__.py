from .gaussian_diffusion import GaussianDiffusion, ClassifierFreeGuidance, InpaintCFG, SuperResCFG
from . import utils
