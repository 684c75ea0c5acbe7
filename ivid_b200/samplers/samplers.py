"""DdpmSampler / DdimSampler — host-side mirrors of the reference samplers
(diffusion/samplers/ddpm.py:12-187, diffusion/samplers/ddim.py:12-165).

Same constructor (`Sampler(framework)`), same float64 numpy table attributes, same `.sample(...)` / `.sample_once(...)`
signatures and return dict (`samples`, `pred_x_t`, `pred_x_0`).  The step itself — UNet forward with both
classifier-free-guidance halves, eps mix, x_{t-1} update, multiview replace/constrain guidance — runs natively behind
the C ABI; `.sample()` keeps the whole reverse process on the device (no per-step host round trips).

RNG: the reference draws `torch.randn_like` inside every step.  `rng="philox"` (default) draws in-kernel
(Philox4x32-10, seeded from torch's generator); `rng="torch"` draws with torch exactly where the reference does
(one `randn_like(x_t)` per step; for InpaintCFG additionally rgb then depth noise before the model call), which keeps
the torch RNG stream consumption identical to the reference.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from ..frameworks.gaussian_diffusion import ClassifierFreeGuidance, GaussianDiffusion, InpaintCFG, SuperResCFG
from ..utils import edict

__all__ = ["DdpmSampler", "DdimSampler"]


def _unwrap(backbone):
    return backbone.module if hasattr(backbone, "module") else backbone


def _f32(t, device):
    return None if t is None else t.to(device=device, dtype=torch.float32).contiguous()


class _NativeSampler:
    KIND = 0

    def __init__(self, framework):
        self.framework = framework
        betas = np.ascontiguousarray(self.framework.betas, dtype=np.float64)
        alphas = 1.0 - betas
        # attribute parity with the reference (ddpm.py:26-41, ddim.py:26-31)
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.lib().ivid_sampler_create(betas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(betas),
                                                  ctypes.byref(self._handle)))
        self._keep = []   # tensors referenced by raw pointer during a native call

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ivid_sampler_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def native_table(self, which: int) -> np.ndarray:
        out = np.empty(len(self.framework.betas), dtype=np.float64)
        _lib.check(_lib.lib().ivid_sampler_table(self._handle, which, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(out)))
        return out

    # ------------------------------------------------------------------------------------------------------------
    def _step_args(self, device, classes, clip_denoised, eta, kwargs, step_noise=None, cond_noise=None, seed=0):
        fw = self.framework
        a = _lib.StepArgsT()
        keep = []

        def P(t):
            if t is None:
                return None
            t = _f32(t, device)
            keep.append(t)
            return t.data_ptr()

        a.kind = self.KIND
        uses_cfg = isinstance(fw, (ClassifierFreeGuidance, InpaintCFG, SuperResCFG))
        a.use_cfg = 1 if uses_cfg else 0
        a.strength = float(kwargs.get("strength", 3.0)) if uses_cfg else 0.0
        a.clip_denoised = 1 if clip_denoised else 0
        a.eta = float(eta)
        if classes is not None:
            c = classes.to(device=device, dtype=torch.int64).contiguous()
            keep.append(c)
            a.classes_dev = c.data_ptr()
        if isinstance(fw, InpaintCFG):
            assert "y" in kwargs and "mask" in kwargs, "InpaintCFG.model_inference needs y and mask"
            a.cond.kind = 1
            a.cond.y_dev = P(kwargs["y"])
            a.cond.mask_dev = P(kwargs["mask"])
            a.cond.mask_rgb_dev = P(kwargs.get("mask_rgb"))
            a.cond.noise_dev = P(cond_noise)
        elif isinstance(fw, SuperResCFG):
            assert "y" in kwargs, "SuperResCFG.model_inference needs y"
            a.cond.kind = 2
            a.cond.y_dev = P(kwargs["y"])
        rr = kwargs.get("replace_rgb")
        if rr is not None:
            assert self.KIND == 1, "replace_rgb is a DdimSampler argument"
            a.replace_rgb_weight = float(rr[0]); a.replace_rgb_dev = P(rr[1]); a.replace_rgb_mask_dev = P(rr[2])
        rd = kwargs.get("replace_depth")
        if rd:
            a.replace_depth_weight = float(rd[0]); a.replace_depth_dev = P(rd[1]); a.replace_depth_mask_dev = P(rd[2])
            cd = kwargs.get("constrain_depth")
            if cd:
                a.constrain_depth_weight = float(cd[0]); a.constrain_depth_dev = P(cd[1])
        a.step_noise_dev = P(step_noise)
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        return a, keep

    def _net(self):
        net = _unwrap(self.framework.backbone)
        net._ensure_packed()
        return net

    def _native_step(self, x_t, t_int, t_prev_int, classes, clip_denoised, eta, kwargs, noise, cond_noise):
        net = self._net()
        dev = x_t.device
        x_t = _f32(x_t, dev)
        a, keep = self._step_args(dev, classes, clip_denoised, eta, kwargs, step_noise=noise, cond_noise=cond_noise)
        x_prev = torch.empty_like(x_t)
        x0 = torch.empty_like(x_t)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ivid_sampler_step(self._handle, net._handle, _lib.ptr(x_t), _lib.ptr(x_prev), _lib.ptr(x0),
                                                    x_t.shape[0], int(t_int), int(t_prev_int), ctypes.byref(a),
                                                    _lib.cur_stream(dev)))
        del keep
        return edict({"pred_x_prev": x_prev, "pred_x_0": x0})

    def _native_step_dev(self, x_t, t, t_prev, classes, clip_denoised, eta, kwargs, noise, cond_noise):
        """Same step with the timestep taken on the device from the [N] tensors the caller passed (no host sync)."""
        net = self._net()
        dev = x_t.device
        x_t = _f32(x_t, dev)
        a, keep = self._step_args(dev, classes, clip_denoised, eta, kwargs, step_noise=noise, cond_noise=cond_noise)
        td = t.to(device=dev, dtype=torch.int64).contiguous()
        tp = t_prev.to(device=dev, dtype=torch.int64).contiguous() if t_prev is not None else None
        x_prev = torch.empty_like(x_t)
        x0 = torch.empty_like(x_t)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ivid_sampler_step_dev(self._handle, net._handle, _lib.ptr(x_t), _lib.ptr(x_prev), _lib.ptr(x0),
                                                        x_t.shape[0], _lib.ptr(td), _lib.ptr(tp), ctypes.byref(a),
                                                        _lib.cur_stream(dev)))
        del keep
        return edict({"pred_x_prev": x_prev, "pred_x_0": x0})

    def _draw_step_noise(self, x_t, kwargs):
        """torch draws in the reference's order: InpaintCFG rgb, depth (inside model_inference), then randn_like(x_t)."""
        cond_noise = None
        if isinstance(self.framework, InpaintCFG):
            y = kwargs["y"]
            n_rgb = torch.randn_like(y[:, :3])
            n_d = torch.randn_like(y[:, 3:])
            cond_noise = torch.cat([n_rgb, n_d], dim=1)
        return torch.randn_like(x_t), cond_noise

    def _run(self, num, image_size, noise, classes, steps, clip_denoised, eta, verbose, rng, return_trajectory, kwargs):
        net = self._net()
        net.eval()
        if image_size is None:
            image_size = net.image_size
        assert image_size == net.image_size, "image_size must match the backbone"
        shape = (num, net.out_channels, image_size, image_size)
        device = net.device
        img = noise if noise is not None else torch.randn(shape, device=device)
        assert tuple(img.shape) == shape, f"noise must have shape {shape}"
        img = _f32(img, device).clone()
        T = self.framework.timesteps
        nsteps = T if self.KIND == 0 else (steps if steps is not None else T)
        ret = edict({"samples": None, "pred_x_t": [], "pred_x_0": []})
        if rng == "torch":
            if self.KIND == 0:
                sched = [(i, 0) for i in range(T)][::-1]
            else:
                jump = T // nsteps
                sched = [(jump * (i + 1), jump * i) for i in reversed(range(nsteps))]
            for (t, t_prev) in sched:
                # the reference draws the model-input noise first (inside model_inference), then randn_like(x_t)
                cond_noise = None
                if isinstance(self.framework, InpaintCFG):
                    y = kwargs["y"]
                    cond_noise = torch.cat([torch.randn_like(y[:, :3]), torch.randn_like(y[:, 3:])], dim=1)
                z = torch.randn_like(img)
                out = self._native_step(img, t, t_prev, classes, clip_denoised, eta, kwargs, z, cond_noise)
                img = out.pred_x_prev
                if return_trajectory:
                    ret.pred_x_t.append(out.pred_x_prev)
                    ret.pred_x_0.append(out.pred_x_0)
        elif rng == "philox":
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            a, keep = self._step_args(device, classes, clip_denoised, eta, kwargs, seed=seed)
            traj0 = trajt = None
            if return_trajectory:
                traj0 = torch.empty((nsteps,) + shape, dtype=torch.float32, device=device)
                trajt = torch.empty((nsteps,) + shape, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                _lib.check(_lib.lib().ivid_sampler_run(self._handle, net._handle, _lib.ptr(img), num, int(nsteps), ctypes.byref(a),
                                                       None, None, _lib.ptr(traj0), _lib.ptr(trajt), _lib.cur_stream(device)))
            del keep
            if return_trajectory:
                ret.pred_x_t = list(trajt.unbind(0))
                ret.pred_x_0 = list(traj0.unbind(0))
        else:
            raise ValueError("rng must be 'philox' or 'torch'")
        ret.samples = img
        net.train()   # the reference toggles eval()/train() around sampling (ddpm.py:166,186)
        return ret


class DdpmSampler(_NativeSampler):
    """Generate samples with the DDPM ancestral schedule (reference ddpm.py:12)."""
    KIND = 0

    def __init__(self, framework):
        super().__init__(framework)
        betas = self.framework.betas
        alphas = 1.0 - betas
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    @torch.no_grad()
    def sample_once(self, x_t, t, classes=None, clip_denoised=False, noise=None, **kwargs):
        """x_{t-1} from x_t (ddpm.py:111-131).  `t` is the [N] tensor of steps minus 1 (all equal).
        `noise` (extension) injects the randn_like draw; default draws it with torch like the reference."""
        B = x_t.shape[0]
        assert t.shape == (B,), "t must be a 1D tensor of shape (B,)"
        # all samples of a batch share the timestep (both reference samplers are driven that way: ddpm.py:177-179);
        # element 0 is read on the device, so this call does not synchronise with the host
        if noise is None:
            noise, cond_noise = self._draw_step_noise(x_t, kwargs)
        else:
            cond_noise = kwargs.pop("cond_noise", None)
        return self._native_step_dev(x_t, t, None, classes, clip_denoised, 0.0, kwargs, noise, cond_noise)

    @torch.no_grad()
    def sample(self, num, steps=None, image_size=None, noise=None, classes=None, clip_denoised=False, verbose=True,
               rng="philox", return_trajectory=False, **kwargs):
        """Run the full reverse process (ddpm.py:134-187).  `steps` is accepted and ignored exactly as in the reference.
        pred_x_t / pred_x_0 are only materialised with return_trajectory=True (the reference keeps 2x1000 tensors alive;
        its callers read `.samples` only: inference/sample.py:82)."""
        return self._run(num, image_size, noise, classes, None, clip_denoised, 0.0, verbose, rng, return_trajectory, kwargs)


class DdimSampler(_NativeSampler):
    """Generate samples using DDIM (reference ddim.py:12), including the multiview replace/constrain guidance."""
    KIND = 1

    @torch.no_grad()
    def sample_once(self, x_t, t, t_prev, classes=None, clip_denoised=False, eta=0.0, replace_rgb=None,
                    replace_depth=None, constrain_depth=None, noise=None, **kwargs):
        """x_{t_prev} from x_t (ddim.py:48-103).  t / t_prev are [N] tensors of actual steps (1 means one step)."""
        B = x_t.shape[0]
        assert t.shape == (B,) and t_prev.shape == (B,)
        # element 0 of t / t_prev is read on the device (all samples share the step, ddim.py:154-158): no host sync
        kw = dict(kwargs, replace_rgb=replace_rgb, replace_depth=replace_depth, constrain_depth=constrain_depth)
        if noise is None:
            noise, cond_noise = self._draw_step_noise(x_t, kw)
        else:
            cond_noise = kw.pop("cond_noise", None)
        return self._native_step_dev(x_t, t, t_prev, classes, clip_denoised, eta, kw, noise, cond_noise)

    @torch.no_grad()
    def sample(self, num, image_size=None, noise=None, classes=None, steps=None, clip_denoised=False, eta=0.0,
               verbose=True, rng="philox", return_trajectory=False, **kwargs):
        """Run `steps` DDIM steps (ddim.py:106-165)."""
        return self._run(num, image_size, noise, classes, steps, clip_denoised, eta, verbose, rng, return_trajectory, kwargs)
