"""rgbd_3d.utils — host-side mirror of the functions inference/sample.py calls (reference rgbd_3d/utils.py):
linearize_depth :38-58, project_depth :61-67, depth_to_mesh :144-260, aggregate_conditions :420-477 (numpy HWC in/out).
depth_to_mesh and aggregate_conditions run on the CUDA kernels behind the C ABI (mesh build, rasterise + aggregate,
LANCZOS / votes / depth-edge / erosion post-filters); the two depth formulas are plain elementwise numpy as in the
reference.  The sampling loop itself uses rgbd_3d.DeviceWarp and never leaves the GPU."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from ..utils import edict
from .glm_compat import as_matrix
from .renderer import AggregationRenderer, warp_params

__all__ = ["linearize_depth", "project_depth", "depth_to_mesh", "aggregate_conditions"]


def linearize_depth(depth, near=0.5, far=100, mode="z_buffer"):
    if mode == "z_buffer":
        depth = np.clip(depth, 1e-6, 1.0 - 1e-6)
        depth = near * far / (far - (far - near) * depth)
    elif mode == "linear":
        depth = near + (far - near) * depth
    return depth


def project_depth(depth, near=0.5, far=100, mode="z_buffer"):
    if mode == "z_buffer":
        depth = np.clip(depth, near, far)
        depth = (1 / near - 1 / depth) / (1 / near - 1 / far)
    elif mode == "linear":
        depth = (depth - near) / (far - near)
    return depth


_scratch = {}


def _scratch_renderer(n, device=0):
    key = (n, device)
    if key not in _scratch:
        _scratch[key] = AggregationRenderer(n, n, device=device, max_views=1)
    return _scratch[key]


def depth_to_mesh(depth, padding=None, fov=45, modelview=None, atol=None, rtol=None, erode_rgb=None, cal_normal=False):
    """Convert a linearised depth image [H,W,1] to the padded, flagged triangle mesh of the reference (utils.py:144-260).
    Implemented: padding='frustum' (inference/sample.py) or a positive number of pixels (inference/utils.py:load_scene),
    cal_normal=True, with a modelview."""
    numeric = isinstance(padding, (int, float)) and not isinstance(padding, bool) and padding > 0
    if not (padding == "frustum" or numeric) or not cal_normal or modelview is None:
        raise NotImplementedError("depth_to_mesh: padding must be 'frustum' or a positive number, with cal_normal=True and a modelview")
    d = np.ascontiguousarray(np.asarray(depth, dtype=np.float32).reshape(depth.shape[0], depth.shape[1]))
    n = d.shape[0]
    r = _scratch_renderer(n, torch.cuda.current_device())
    V, F = (n + 2) ** 2, 2 * (n + 1) ** 2
    vb = np.empty((V, 9), np.float32)
    faces = np.empty((F, 3), np.uint32)
    mv = np.ascontiguousarray(as_matrix(modelview), dtype=np.float32)
    p = warp_params(fov, 1.0, 2.0, atol, rtol, erode_rgb, padding=float(padding) if numeric else 0.0)   # near/far unused: depth is linear
    _lib.check(_lib.lib().ivid_warp_mesh_from_depth(r._handle, d.ctypes.data, mv.ctypes.data, ctypes.byref(p), vb.ctypes.data,
                                                    faces.ctypes.data, r._stream()))
    return edict({
        "depth": depth, "fov": fov, "modelview": modelview, "faces": faces.astype(np.int64),
        "vertices": edict({"position": vb[:, 0:3], "normal": vb[:, 3:6], "uv": vb[:, 6:8], "flag": vb[:, 8:9]}),
    })


def aggregate_conditions(renderer, meshes, colors, modelview, fov=45, near=0.5, mode="z_buffer", far=100, atol=0.02,
                         rtol=0.02, erode_rgb=2):
    """Aggregate the partial RGBD condition of a target view from all previous views (utils.py:420-477)."""
    if mode != "z_buffer":
        raise NotImplementedError("aggregate_conditions: only mode='z_buffer' is on the sampling path")
    renderer.render(meshes, colors, modelview, fov, is_autoregressive=True)
    color, depth, mc, md = renderer._last_raw
    n = colors[0].shape[0]
    out = torch.empty((1, 7, n, n), dtype=torch.float32, device=color.device)
    p = warp_params(fov, near, far, atol, rtol, erode_rgb)
    with torch.cuda.device(color.device):
        _lib.check(_lib.lib().ivid_warp_postfilter(renderer._handle, _lib.ptr(color), _lib.ptr(depth), _lib.ptr(mc), _lib.ptr(md),
                                                   ctypes.byref(p), _lib.ptr(out), renderer._stream()))
    o = out[0].permute(1, 2, 0).cpu().numpy()
    return edict({"color": o[:, :, 0:3].copy(), "depth": o[:, :, 3:4].copy(), "mask": o[:, :, 4:5].copy(),
                  "mask_rgb": o[:, :, 5:6].copy(), "depth_convex": o[:, :, 6:7].copy()})
