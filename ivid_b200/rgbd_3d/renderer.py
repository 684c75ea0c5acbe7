"""AggregationRenderer — host-side mirror of rgbd_3d.AggregationRenderer (reference moderngl_renderer.py:151-340) on
the CUDA rasteriser behind the C ABI (no OpenGL / EGL), plus DeviceWarp: the device-resident multiview warp used by
the sampling loop (all source views of a batch stay in HBM; no per-view host round trips).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from ..utils import edict
from .glm_compat import as_matrix

__all__ = ["AggregationRenderer", "SimpleRenderer", "DeviceWarp", "warp_params"]


def warp_params(fov=45, near=0.5, far=100, atol=0.02, rtol=0.02, erode_rgb=2, padding=0.0) -> _lib.WarpParamsT:
    """padding: 0 = 'frustum' (sampling path), > 0 = numeric padding in pixels (free-view rendering, load_scene; the
    training-pair warp), < 0 = None (no padding ring)."""
    p = _lib.WarpParamsT()
    p.padding = float(padding)
    p.fov_deg, p.near, p.far = float(fov), float(near), float(far)
    # None travels as a negative value: depth_to_mesh skips the discontinuity test when BOTH are None and uses 0 for a single
    # None (reference utils.py:227-229)
    p.atol = -1.0 if atol is None else float(atol)
    p.rtol = -1.0 if rtol is None else float(rtol)
    p.erode_rgb = int(erode_rgb or 0)
    return p


def _mv(m) -> np.ndarray:
    return np.ascontiguousarray(as_matrix(m), dtype=np.float32)


class _Native:
    def __init__(self, image_size, render_size, max_views, batch, near, far, device):
        assert torch.cuda.is_available(), "ivid_b200.rgbd_3d runs on CUDA only (no CPU path)"
        self.image_size, self.render_size, self.max_views, self.batch = image_size, render_size, max_views, batch
        self.near, self.far, self.device = near, far, device
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.lib().ivid_warp_create(image_size, render_size, max_views, batch, float(near), float(far), int(device),
                                               ctypes.byref(self._handle)))

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ivid_warp_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _stream(self):
        return _lib.cur_stream(torch.device("cuda", self.device))


class AggregationRenderer(_Native):
    """Renderer of 3D meshes with multi-view aggregation (reference moderngl_renderer.py:151).

    Args (as the reference): render_size, image_size, near, far, device, max_views.
    """

    def __init__(self, render_size=128, image_size=128, near=0.01, far=200.0, device=0, max_views=27):
        # one spare slot is used by utils.depth_to_mesh as scratch
        super().__init__(image_size, render_size, max_views + 1, 1, near, far, device)
        self._uploaded = 0

    def render(self, meshes, colors, modelview, fov=45.0, is_autoregressive=False, verbose=False, tqdm_args={}):
        """Render the aggregated view(s) (reference moderngl_renderer.py:260-340).  numpy HWC in / out."""
        L = _lib.lib()
        if not is_autoregressive:
            _lib.check(L.ivid_warp_reset(self._handle))      # a fresh set of source views (e.g. the next scene of render.py)
        for i, mesh in enumerate(meshes):
            if is_autoregressive and i != len(meshes) - 1:
                continue      # earlier views were uploaded by earlier calls (moderngl_renderer.py:281-283)
            v = mesh.vertices if hasattr(mesh, "vertices") else mesh["vertices"]
            vb = np.ascontiguousarray(np.concatenate([v["position"], v["normal"], v["uv"], v["flag"]], axis=-1).astype(np.float32))
            faces = np.ascontiguousarray((mesh.faces if hasattr(mesh, "faces") else mesh["faces"]).astype(np.uint32))
            col = np.ascontiguousarray(colors[i].astype(np.float32))
            mv = _mv(mesh.modelview if hasattr(mesh, "modelview") else mesh["modelview"])
            _lib.check(L.ivid_warp_set_mesh(self._handle, 0, i, vb.ctypes.data, faces.ctypes.data, col.ctypes.data, mv.ctypes.data))
        # the native view count is max(uploaded)+1; make it match len(meshes)
        targets = modelview if isinstance(modelview, list) else [modelview]
        S = self.render_size
        dev = torch.device("cuda", self.device)
        ret = []
        for t in targets:
            color = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
            depth = torch.empty((S, S), dtype=torch.float32, device=dev)
            mc = torch.empty((S, S), dtype=torch.float32, device=dev)
            md = torch.empty((S, S), dtype=torch.float32, device=dev)
            tm = _mv(t)
            with torch.cuda.device(dev):
                self._set_views(len(meshes))
                _lib.check(L.ivid_warp_render(self._handle, tm.ctypes.data, 1, float(fov), _lib.ptr(color), _lib.ptr(depth),
                                              _lib.ptr(mc), _lib.ptr(md), self._stream()))
            self._last_raw = (color, depth, mc, md)
            ret.append(edict({
                "color": color.cpu().numpy(),
                "depth": depth.cpu().numpy()[..., None],
                "mask_color": mc.cpu().numpy()[..., None] > 0.5,
                "mask_depth": md.cpu().numpy()[..., None] > 0.5,
            }))
        return ret if len(ret) > 1 else ret[0]

    def render_resolved(self, meshes, colors, modelviews, fov=45.0, project_near=0.5, project_far=100, lut=None):
        """Free-view frames resolved on the device (inference/render.py:74-84): for every target modelview the aggregated
        render at render_size is down-sampled by the 8-bit LANCZOS kernels and its centre-sampled, projected depth is colour
        mapped through `lut` (uint8 [256,3]); only the two uint8 [image_size, image_size, 3] images cross PCIe.
        Returns (colors uint8 [F,n,n,3], depths uint8 [F,n,n,3])."""
        L = _lib.lib()
        _lib.check(L.ivid_warp_reset(self._handle))
        for i, mesh in enumerate(meshes):
            v = mesh.vertices if hasattr(mesh, "vertices") else mesh["vertices"]
            vb = np.ascontiguousarray(np.concatenate([v["position"], v["normal"], v["uv"], v["flag"]], axis=-1).astype(np.float32))
            faces = np.ascontiguousarray((mesh.faces if hasattr(mesh, "faces") else mesh["faces"]).astype(np.uint32))
            col = np.ascontiguousarray(colors[i].astype(np.float32))
            mv = _mv(mesh.modelview if hasattr(mesh, "modelview") else mesh["modelview"])
            _lib.check(L.ivid_warp_set_mesh(self._handle, 0, i, vb.ctypes.data, faces.ctypes.data, col.ctypes.data, mv.ctypes.data))
        n = self.image_size
        lut = np.ascontiguousarray(lut, dtype=np.uint8).reshape(256, 3)
        cols, deps = [], []
        with torch.cuda.device(torch.device("cuda", self.device)):
            for t in (modelviews if isinstance(modelviews, list) else [modelviews]):
                tm = _mv(t)
                _lib.check(L.ivid_warp_render(self._handle, tm.ctypes.data, 1, float(fov), None, None, None, None, self._stream()))
                c8 = np.empty((n, n, 3), np.uint8); d8 = np.empty((n, n, 3), np.uint8)
                _lib.check(L.ivid_warp_resolve_frame(self._handle, float(project_near), float(project_far), lut.ctypes.data, c8.ctypes.data,
                                                     d8.ctypes.data, self._stream()))
                cols.append(c8); deps.append(d8)
        return np.stack(cols, axis=0), np.stack(deps, axis=0)

    def _set_views(self, n):
        cur = ctypes.c_int()
        _lib.check(_lib.lib().ivid_warp_num_views(self._handle, ctypes.byref(cur)))
        assert cur.value >= n, "render: meshes were not uploaded (is_autoregressive=True needs every view rendered once)"
        if cur.value != n:
            raise AssertionError(f"renderer holds {cur.value} views but {n} meshes were passed; create a new renderer per sample")


class SimpleRenderer(_Native):
    """Renderer of one textured grid mesh with raw texture colours (reference moderngl_renderer.py:11-148 + shaders/simple.*):
    alpha = 0 on back faces and discontinuity edges.  Used by rgbd_3d.utils.forward_backward_warp (training-pair synthesis,
    datasets/base.py:219).  Same constructor and .render() contract as the reference class; numpy HWC in / out."""

    def __init__(self, render_size=128, image_size=128, near=0.01, far=200.0, device=0):
        super().__init__(image_size, render_size, 2, 1, near, far, device)

    def render(self, mesh, color, modelview, fov=45.0):
        L = _lib.lib()
        v = mesh.vertices if hasattr(mesh, "vertices") else mesh["vertices"]
        pos = np.asarray(v["position"], np.float32)
        nrm = np.asarray(v["normal"], np.float32) if "normal" in v else np.tile(np.float32([0, 0, 1]), (pos.shape[0], 1))
        vb = np.ascontiguousarray(np.concatenate([pos, nrm, np.asarray(v["uv"], np.float32), np.asarray(v["flag"], np.float32)], axis=-1))
        faces = np.ascontiguousarray(np.asarray(mesh.faces if hasattr(mesh, "faces") else mesh["faces"]).astype(np.uint32))
        col = np.ascontiguousarray(np.asarray(color).astype(np.float32))
        S = self.render_size
        targets = modelview if isinstance(modelview, list) else [modelview]
        ret = []
        for t in targets:
            c = np.empty((S, S, 3), np.float32); d = np.empty((S, S, 1), np.float32); m = np.empty((S, S, 1), np.float32)
            tm = _mv(t)
            with torch.cuda.device(self.device):
                _lib.check(L.ivid_warp_render_simple(self._handle, vb.ctypes.data, vb.shape[0], faces.ctypes.data, faces.shape[0],
                                                     col.ctypes.data, tm.ctypes.data, float(fov), c.ctypes.data, d.ctypes.data,
                                                     m.ctypes.data, self._stream()))
            ret.append(edict({"color": c, "depth": d, "mask": m > 0.5}))
        return ret if len(ret) > 1 else ret[0]

    def forward_backward(self, lin_depth0, color0, modelview1, modelview0, padding, fov, near, far, atol, rtol):
        """The whole forward_backward_warp on the device (one upload, one download); see rgbd_3d.utils.forward_backward_warp."""
        n = self.image_size
        d0 = np.ascontiguousarray(np.asarray(lin_depth0, np.float32).reshape(1, n, n))
        c0 = np.ascontiguousarray(np.asarray(color0, np.float32).reshape(1, n, n, 3))
        out = np.empty((1, 7, n, n), np.float32)
        p = warp_params(fov, near, far, atol, rtol, 0, padding=-1.0 if padding is None else (0.0 if padding == "frustum" else float(padding)))
        m1, m0 = _mv(modelview1), _mv(modelview0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ivid_warp_forward_backward(self._handle, d0.ctypes.data, c0.ctypes.data, m1.ctypes.data, m0.ctypes.data, 1,
                                                             ctypes.byref(p), out.ctypes.data, self._stream()))
        return out[0]


class DeviceWarp(_Native):
    """Device-resident replacement of the per-sample `[AggregationRenderer(...)]` list + mesh lists of
    inference/sample.py:50,57-58: every sample's source views (RGBD) stay on the GPU."""

    def __init__(self, batch, image_size=128, ssaa=3, max_views=27, near=0.01, far=200.0, device=None):
        device = torch.cuda.current_device() if device is None else device
        super().__init__(image_size, image_size * ssaa, max_views, batch, near, far, device)

    def reset(self):
        _lib.check(_lib.lib().ivid_warp_reset(self._handle))

    @property
    def num_views(self):
        n = ctypes.c_int()
        _lib.check(_lib.lib().ivid_warp_num_views(self._handle, ctypes.byref(n)))
        return n.value

    def _mvs(self, modelviews):
        """-> (float32 [k,4,4] array, shared flag).  A list of `batch` matrices is per-sample, anything else is shared."""
        if isinstance(modelviews, (list, tuple)) and len(modelviews) == self.batch:
            try:
                per = [_mv(m) for m in modelviews]
                return np.ascontiguousarray(np.stack(per)), 0
            except (AssertionError, TypeError, ValueError, IndexError):
                pass
        return _mv(modelviews)[None].copy(), 1

    def add_view(self, rgbd, modelviews, fov=45, near=0.6, far=5, atol=0.03, rtol=0.03, erode_rgb=3):
        """rgbd: cuda fp32 [B,4,H,W] sampler output in [-1,1]; modelviews: one 4x4 (shared) or a list of B."""
        assert rgbd.is_cuda and rgbd.shape == (self.batch, 4, self.image_size, self.image_size)
        r = rgbd.to(torch.float32).contiguous()
        mv, shared = self._mvs(modelviews)
        p = warp_params(fov, near, far, atol, rtol, erode_rgb)
        with torch.cuda.device(r.device):
            _lib.check(_lib.lib().ivid_warp_add_view(self._handle, _lib.ptr(r), mv.ctypes.data, shared, ctypes.byref(p), self._stream()))

    def aggregate(self, modelviews, fov=45, near=0.6, far=5, atol=0.03, rtol=0.03, erode_rgb=3):
        """aggregate_conditions for every sample -> cuda fp32 [B,7,H,W]: color(3), depth, mask, mask_rgb, depth_convex."""
        dev = torch.device("cuda", self.device)
        out = torch.empty((self.batch, 7, self.image_size, self.image_size), dtype=torch.float32, device=dev)
        mv, shared = self._mvs(modelviews)
        p = warp_params(fov, near, far, atol, rtol, erode_rgb)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ivid_warp_aggregate(self._handle, mv.ctypes.data, shared, ctypes.byref(p), _lib.ptr(out), self._stream()))
        return out

    def render_raw(self, modelviews, fov=45):
        dev = torch.device("cuda", self.device)
        S, B = self.render_size, self.batch
        color = torch.empty((B, S, S, 3), dtype=torch.float32, device=dev)
        depth = torch.empty((B, S, S), dtype=torch.float32, device=dev)
        mc = torch.empty((B, S, S), dtype=torch.float32, device=dev)
        md = torch.empty((B, S, S), dtype=torch.float32, device=dev)
        mv, shared = self._mvs(modelviews)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ivid_warp_render(self._handle, mv.ctypes.data, shared, float(fov), _lib.ptr(color), _lib.ptr(depth),
                                                   _lib.ptr(mc), _lib.ptr(md), self._stream()))
        return color, depth, mc, md

    def get_mesh(self, sample, view):
        n = self.image_size
        V, F = (n + 2) ** 2, 2 * (n + 1) ** 2
        vb = np.empty((V, 9), np.float32); faces = np.empty((F, 3), np.uint32); col = np.empty((n, n, 3), np.float32)
        _lib.check(_lib.lib().ivid_warp_get_mesh(self._handle, sample, view, vb.ctypes.data, faces.ctypes.data, col.ctypes.data))
        return vb, faces, col
