from .samplers import DdpmSampler, DdimSampler
