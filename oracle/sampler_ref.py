"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's diffusion frameworks and samplers with *injected* noise (the reference draws
torch.randn_like inside each step; the restatement takes the same draws as arguments so that device and CPU can be
compared step by step):
    reference: diffusion/frameworks/utils.py:7-37              get_betas_by_name
               diffusion/frameworks/classifier_free_guidance.py:23-42
               diffusion/frameworks/inpaint_cfg.py:24-49,61-83  make_cond_inputs / model_inference
               diffusion/frameworks/sr_cfg.py:23-36,39-60
               diffusion/samplers/ddpm.py:20-41,66-131          tables, p_mean_variance, sample_once
               diffusion/samplers/ddim.py:19-31,48-103          tables, sample_once (replace / constrain guidance)
Pinned against the reference classes by tests/golden/make_golden.py (schedule known-answer values of SURVEY.md
Appendix C + seeded step fixtures).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.nn.functional as F


def get_betas(schedule: str, T: int) -> np.ndarray:
    if schedule == "linear":
        scale = 1000 / T
        return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
    if schedule == "cosine":
        f = lambda t: np.cos((t + 0.008) / 1.008 * np.pi / 2) ** 2
        return np.array([min(1 - f((i + 1) / T) / f(i / T), 0.999) for i in range(T)])
    raise NotImplementedError(f"unknown beta schedule: {schedule}")


class Tables:
    """float64 tables of DdpmSampler.__init__ (ddpm.py:26-41) / DdimSampler.__init__ (ddim.py:26-31)."""

    def __init__(self, betas: np.ndarray):
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def _ex(arr: np.ndarray, t: torch.Tensor, ndim: int) -> torch.Tensor:
    # samplers/utils.py:20-23: float64 table -> index -> float32 -> broadcast
    r = torch.from_numpy(arr).to(t.device)[t].float()      # (the reference copies the table to t's device on every call)
    return r.view(-1, *([1] * (ndim - 1)))


def cfg_eps(model: Callable, x, t, classes, strength: float):
    """ClassifierFreeGuidance.model_inference (classifier_free_guidance.py:39-42)."""
    return (1 + strength) * model(x, t, classes) - (strength * model(x, t, None) if strength > 0 else 0)


def make_inpaint_inputs(x, y, mask, mask_rgb, noise_rgb, noise_depth):
    """InpaintCFG.make_cond_inputs (inpaint_cfg.py:33-49) with the two randn_like draws injected."""
    parts = [x]
    if mask_rgb is not None:
        parts.append(mask_rgb)
        m_rgb = mask_rgb
    else:
        m_rgb = mask
    parts.append(y[:, :3] * m_rgb + noise_rgb * (1 - m_rgb))
    parts.append(y[:, 3:] * mask + noise_depth * (1 - mask))
    parts.append(mask)
    return torch.cat(parts, dim=1)


def make_sr_inputs(x, y):
    """SuperResCFG.make_cond_inputs (sr_cfg.py:31-36)."""
    scale = x.shape[-1] // y.shape[-1]
    return torch.cat([x, F.interpolate(y, scale_factor=scale, mode="bilinear", align_corners=False)], dim=1)


def cond_eps(model: Callable, cond_inputs, t, classes, strength: float):
    """InpaintCFG / SuperResCFG.model_inference tail (inpaint_cfg.py:77-83, sr_cfg.py:53-60)."""
    if classes is None:
        return model(cond_inputs, t, None)
    return (1 + strength) * model(cond_inputs, t, classes) - (strength * model(cond_inputs, t, None) if strength > 0 else 0)


def ddpm_step(tb: Tables, x_t, t: torch.Tensor, eps, noise, clip_denoised=False):
    """DdpmSampler.p_mean_variance + sample_once (ddpm.py:85-100, 127-131). t is the step minus 1."""
    nd = x_t.dim()
    x0 = _ex(tb.sqrt_recip_alphas_cumprod, t, nd) * x_t - _ex(tb.sqrt_recipm1_alphas_cumprod, t, nd) * eps
    if clip_denoised:
        x0 = x0.clamp(-1, 1)
    mean = _ex(tb.posterior_mean_coef1, t, nd) * x0 + _ex(tb.posterior_mean_coef2, t, nd) * x_t
    logvar = _ex(tb.posterior_log_variance_clipped, t, nd)
    nz = (t != 0).float().view(-1, *([1] * (nd - 1)))
    return mean + nz * torch.exp(0.5 * logvar) * noise, x0


def ddim_step(tb: Tables, x_t, t: torch.Tensor, t_prev: torch.Tensor, eps, noise, clip_denoised=False, eta=0.0,
              replace_rgb=None, replace_depth=None, constrain_depth=None):
    """DdimSampler.sample_once after the model call (ddim.py:82-103).  t is the actual step (1-indexed)."""
    nd = x_t.dim()
    srac = _ex(tb.sqrt_recip_alphas_cumprod, t - 1, nd)
    srm1 = _ex(tb.sqrt_recipm1_alphas_cumprod, t - 1, nd)
    x0 = srac * x_t - srm1 * eps
    nz = (t_prev != 0).float().view(-1, *([1] * (nd - 1)))
    if clip_denoised:
        x0 = torch.clamp(x0, -1.0, 1.0)
    x0 = x0.clone()
    if replace_rgb is not None:
        w, rgb, m = replace_rgb
        x0[:, :3] = (1 - nz) * x0[:, :3] + nz * ((w * rgb + (1 - w) * x0[:, :3]) * m + x0[:, :3] * (1 - m))
    if replace_depth:
        w, d, m = replace_depth
        x0[:, 3:] = (w * d + (1 - w) * x0[:, 3:]) * m + x0[:, 3:] * (1 - m)
        if constrain_depth:
            cw, convex = constrain_depth
            x0[:, 3:] = x0[:, 3:] * m + (cw * torch.maximum(x0[:, 3:], convex) + (1 - cw) * x0[:, 3:]) * (1 - m)
    eps2 = (srac * x_t - x0) / srm1
    ab = _ex(tb.alphas_cumprod, t - 1, nd)
    abp = _ex(tb.alphas_cumprod_prev, t_prev, nd)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    mean = torch.sqrt(abp) * x0 + torch.sqrt(1 - abp - sigma ** 2) * eps2
    return mean + nz * sigma * noise, x0


def ddim_schedule(T: int, steps: int):
    jump = T // steps
    return [(jump * (i + 1), jump * i) for i in reversed(range(steps))]   # ddim.py:153-154
