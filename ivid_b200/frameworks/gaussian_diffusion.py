"""Diffusion frameworks — host-side mirrors of the reference classes (same names, ctor kwargs, attributes):

    GaussianDiffusion        diffusion/frameworks/gaussian_diffusion.py:12-116
    ClassifierFreeGuidance   diffusion/frameworks/classifier_free_guidance.py:12-75
    InpaintCFG               diffusion/frameworks/inpaint_cfg.py:11-128
    SuperResCFG              diffusion/frameworks/sr_cfg.py:11-96

`model_inference` keeps the reference semantics but evaluates both classifier-free-guidance halves in ONE batch-2N
native forward.  Training losses are out of scope of the sampling hot path.
"""
from __future__ import annotations

import ctypes
import inspect

import numpy as np
import torch
import torch.nn.functional as F

from .. import _lib
from ..utils import edict
from .utils import get_betas_by_name

__all__ = ["GaussianDiffusion", "ClassifierFreeGuidance", "InpaintCFG", "SuperResCFG"]


def _extract(arr, timesteps, broadcast_shape):
    # reference frameworks/utils.py:63-80
    res = torch.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


def _unwrap(backbone):
    return backbone.module if hasattr(backbone, "module") else backbone


class GaussianDiffusion:
    """Utilities for sampling diffusion models (reference gaussian_diffusion.py:12)."""

    def __init__(self, backbone, timesteps=1000, beta_schedule="linear"):
        self.backbone = backbone
        self.timesteps = timesteps
        self.beta_schedule = beta_schedule
        self.backbone_args = edict(inspect.signature(_unwrap(backbone).forward).parameters)
        betas = get_betas_by_name(self.beta_schedule, self.timesteps).astype(np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all(), "betas must be in (0, 1]"
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)

    # --- q(x_t | x_0) helpers (gaussian_diffusion.py:45-74); plain torch, not on the hot path -------------------
    def diffuse(self, x_0, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_0)
        assert noise.shape == x_0.shape, "noise must have same shape as x_0"
        return (_extract(self.sqrt_alphas_cumprod, t, x_0.shape) * x_0
                + _extract(self.sqrt_one_minus_alphas_cumprod, t, x_0.shape) * noise)

    def reverse_diffuse(self, x_t, t, noise):
        assert noise.shape == x_t.shape, "noise must have same shape as x_t"
        return ((x_t - _extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * noise)
                / _extract(self.sqrt_alphas_cumprod, t, x_t.shape))

    # --- native helpers -------------------------------------------------------------------------------------------
    def _cfg_forward(self, x, t, classes, strength):
        """(1+s)*eps(x,t,c) - s*eps(x,t,None) with both halves in one batch-2N forward sharing x."""
        net = _unwrap(self.backbone)
        if not (strength > 0):
            return (1 + strength) * net(x, t, classes)
        assert net.num_classes is not None, "this model is not class-conditioned"
        assert net.has_null_class, "this model does not have a null class"
        N = x.shape[0]
        net._ensure_packed()
        xx = x.to(torch.float32).contiguous()
        t2 = torch.cat([t, t]).to(device=x.device, dtype=torch.int64).contiguous()
        if classes is None:
            c2 = torch.full((2 * N,), -1, dtype=torch.int64, device=x.device)
        else:
            c2 = torch.cat([classes.to(x.device, torch.int64), torch.full((N,), -1, dtype=torch.int64, device=x.device)]).contiguous()
        out = torch.empty((2 * N, net.out_channels, net.image_size, net.image_size), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ivid_unet_forward(net._handle, _lib.ptr(xx), N, _lib.ptr(t2), _lib.ptr(c2), _lib.ptr(out),
                                                    2 * N, _lib.cur_stream(x.device)))
        return (1 + strength) * out[:N] - strength * out[N:]

    @torch.no_grad()
    def model_inference(self, x, t, classes=None, **kwargs):
        """Predicted noise (gaussian_diffusion.py:76-91)."""
        kwargs = {k: v for k, v in kwargs.items() if k in self.backbone_args}
        return self.backbone(x, t, classes, **kwargs)

    def training_losses(self, *a, **k):
        raise NotImplementedError("training is outside the sampling hot path this package implements")


class ClassifierFreeGuidance(GaussianDiffusion):
    """Diffusion model with classifier-free guidance (classifier_free_guidance.py:12)."""

    def __init__(self, backbone, *, p_uncond=0.1, **kwargs):
        super().__init__(backbone, **kwargs)
        self.p_uncond = p_uncond

    @torch.no_grad()
    def model_inference(self, x, t, classes=None, strength=3.0, **kwargs):
        # classifier_free_guidance.py:38-42
        return self._cfg_forward(x, t, classes, strength)


class InpaintCFG(GaussianDiffusion):
    """Image inpainting with classifier-free guidance (inpaint_cfg.py:11)."""

    def __init__(self, backbone, *, p_uncond=0.1, p_uncond_img=0.0, **kwargs):
        super().__init__(backbone, **kwargs)
        self.p_uncond = p_uncond
        self.p_uncond_img = p_uncond_img

    def make_cond_inputs(self, x, y, mask, **kwargs):
        # inpaint_cfg.py:24-49 (torch RNG semantics preserved: rgb noise is drawn before depth noise)
        y_rgb = y[:, :3]
        y_depth = y[:, 3:]
        in_list = [x]
        if "mask_rgb" in kwargs:
            mask_rgb = kwargs["mask_rgb"]
            in_list.append(mask_rgb)
        else:
            mask_rgb = mask
        y_rgb = y_rgb * mask_rgb + torch.randn_like(y_rgb) * (1 - mask_rgb)
        in_list.append(y_rgb)
        y_depth = y_depth * mask + torch.randn_like(y_depth) * (1 - mask)
        in_list.append(y_depth)
        in_list.append(mask)
        return torch.cat(in_list, dim=1)

    def make_uncond_inputs(self, x):
        return torch.cat([x, torch.randn_like(x), torch.zeros_like(x[:, :1])], dim=1)

    @torch.no_grad()
    def model_inference(self, x, t, y, mask, classes=None, strength=3.0, **kwargs):
        # inpaint_cfg.py:61-83
        cond_inputs = self.make_cond_inputs(x, y, mask, **kwargs)
        if classes is None:
            return self.backbone(cond_inputs, t, None)
        return self._cfg_forward(cond_inputs, t, classes, strength)


class SuperResCFG(GaussianDiffusion):
    """Image super-resolution with classifier-free guidance (sr_cfg.py:11)."""

    def __init__(self, backbone, *, p_uncond=0.1, **kwargs):
        super().__init__(backbone, **kwargs)
        self.p_uncond = p_uncond

    def make_cond_inputs(self, x, y, **kwargs):
        # sr_cfg.py:23-36
        scale = x.shape[-1] // y.shape[-1]
        y = F.interpolate(y, scale_factor=scale, mode="bilinear", align_corners=False)
        return torch.cat([x, y], dim=1)

    @torch.no_grad()
    def model_inference(self, x, t, y, classes=None, strength=3.0, **kwargs):
        # sr_cfg.py:39-60
        cond_inputs = self.make_cond_inputs(x, y, **kwargs)
        if classes is None:
            return self.backbone(cond_inputs, t, None)
        return self._cfg_forward(cond_inputs, t, classes, strength)
