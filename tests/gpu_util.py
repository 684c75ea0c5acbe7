"""Helpers for the -m gpu parity tests (all calls go through the C ABI via ctypes)."""
import ctypes

import numpy as np
import torch

from ivid_b200 import _lib


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def report(name, got, want):
    got = got.detach().float().cpu(); want = want.detach().float().cpu()
    d = (got - want).abs()
    r = rel(got, want)
    print(f"[parity] {name}: rel_l2={r:.3e} max_abs={float(d.max()):.3e} ref_rms={float(want.pow(2).mean().sqrt()):.3e} "
          f"nan={int(torch.isnan(got).sum())} shape={tuple(got.shape)}")
    return r


def conv2d(act_nhwc_f16, w, b, ksize, act2=None, w2=None, b2=None, residual=None, out_fp16=False):
    N, H, W, Cin = act_nhwc_f16.shape
    Cout = w.shape[0]
    out = torch.empty((N, H, W, Cout), dtype=torch.float16 if out_fp16 else torch.float32, device="cuda")
    wc = w.detach().float().cpu().contiguous(); bc = b.detach().float().cpu().contiguous()
    w2c = w2.detach().float().cpu().contiguous() if w2 is not None else None
    b2c = b2.detach().float().cpu().contiguous() if b2 is not None else None
    _lib.check(_lib.lib().ivid_op_conv2d(_lib.ptr(act_nhwc_f16), N, H, W, Cin, _lib.ptr(wc), _lib.ptr(bc), Cout, ksize,
                                         _lib.ptr(act2), act2.shape[-1] if act2 is not None else 0, _lib.ptr(w2c), _lib.ptr(b2c),
                                         _lib.ptr(residual), _lib.ptr(out), 1 if out_fp16 else 0, _lib.cur_stream()))
    return out


def group_norm(x0, x1, groups, gamma, beta, film, silu, mode):
    N, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    Ho = H * 2 if mode == 1 else (H // 2 if mode == 2 else H)
    Wo = W * 2 if mode == 1 else (W // 2 if mode == 2 else W)
    out = torch.empty((N, Ho, Wo, C0 + C1), dtype=torch.float16, device="cuda")
    g = gamma.float().cpu().contiguous(); bt = beta.float().cpu().contiguous()
    _lib.check(_lib.lib().ivid_op_group_norm(_lib.ptr(x0), C0, _lib.ptr(x1), C1, N, H, W, groups, 1e-5, _lib.ptr(g), _lib.ptr(bt),
                                             _lib.ptr(film), 1 if silu else 0, mode, _lib.ptr(out), _lib.cur_stream()))
    return out


def attention(qkv_f16, C):
    N, T, _ = qkv_f16.shape
    out = torch.empty((N, T, C), dtype=torch.float16, device="cuda")
    _lib.check(_lib.lib().ivid_op_attention(_lib.ptr(qkv_f16), N, T, C, _lib.ptr(out), _lib.cur_stream()))
    return out
