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

__all__ = ["linearize_depth", "project_depth", "depth_to_mesh", "aggregate_conditions", "forward_backward_warp"]


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
    """Convert a linearised depth image [H,W,1] to the flagged triangle mesh of the reference (utils.py:144-260): padding None
    (H*W vertices; forward_backward_warp's second mesh), 'frustum' (inference/sample.py) or a number of pixels
    (inference/utils.py:load_scene, datasets/base.py:238).  modelview None leaves the points in camera space.  The mesh is
    built on the GPU; `normal` is only returned with cal_normal=True, as in the reference."""
    numeric = isinstance(padding, (int, float)) and not isinstance(padding, bool)
    if not (padding is None or padding == "frustum" or (numeric and padding > 0)):
        raise NotImplementedError("depth_to_mesh: padding must be None, 'frustum' or a positive number")
    d = np.ascontiguousarray(np.asarray(depth, dtype=np.float32).reshape(depth.shape[0], depth.shape[1]))
    n = d.shape[0]
    r = _scratch_renderer(n, torch.cuda.current_device())
    m = n + (0 if padding is None else 2)
    vb = np.empty((m * m, 9), np.float32)
    faces = np.empty((2 * (m - 1) * (m - 1), 3), np.uint32)
    mv = np.ascontiguousarray(as_matrix(modelview) if modelview is not None else np.eye(4), dtype=np.float32)
    pad = -1.0 if padding is None else (float(padding) if numeric else 0.0)
    p = warp_params(fov, 1.0, 2.0, atol, rtol, erode_rgb, padding=pad)   # near/far unused: depth is linear
    _lib.check(_lib.lib().ivid_warp_mesh_from_depth(r._handle, d.ctypes.data, mv.ctypes.data, ctypes.byref(p), vb.ctypes.data,
                                                    faces.ctypes.data, r._stream()))
    verts = edict({"position": vb[:, 0:3], "uv": vb[:, 6:8], "flag": vb[:, 8:9]})
    if cal_normal:
        verts["normal"] = vb[:, 3:6]
    return edict({"depth": depth, "fov": fov, "modelview": modelview, "faces": faces.astype(np.int64), "vertices": verts})


def forward_backward_warp(renderer, rgbd, modelview1, modelview0=None, padding=None, fov=45, near=0.5, far=100, mode="z_buffer",
                          atol=0.02, rtol=0.02):
    """Warp an RGBD image [H,W,4] (values in [0,1], z-buffer depth) to view 1 and back: what survives both trips and is not a
    depth edge is the partial condition of a training pair (reference utils.py:335-417, called by datasets/base.py:238).
    `renderer` is an rgbd_3d.SimpleRenderer; both meshes, both renders and all resolves stay on the device.
    Returns edict(color [H,W,3], depth [H,W,1], mask [H,W,1]) like the reference."""
    if mode != "z_buffer":
        raise NotImplementedError("forward_backward_warp: only mode='z_buffer' is implemented")
    from . import glm_compat as glm
    if modelview0 is None:
        modelview0 = glm.lookAt(glm.vec3(0.0, 0.0, 1.0), glm.vec3(0.0, 0.0, 0.0), glm.vec3(0.0, 1.0, 0.0))
    rgbd = np.asarray(rgbd)
    out = renderer.forward_backward(linearize_depth(rgbd[:, :, 3:], near, far, mode), rgbd[:, :, :3], modelview1, modelview0, padding,
                                    fov, near, far, atol, rtol)
    o = out.transpose(1, 2, 0)
    return edict({"color": o[:, :, 0:3].copy(), "depth": o[:, :, 3:4].copy(), "mask": o[:, :, 4:5].copy()})


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
