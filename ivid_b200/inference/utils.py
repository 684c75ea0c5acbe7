"""Scene / list helpers of the sampling driver (reference inference/utils.py): parse_int_list :13-22, reorder :44-55,
save_scene :74-101, load_scene :104-113.  The scene container is the reference's: an .npz whose `data` entry is a list of
{color: PNG bytes (uint8 RGB), depth: PNG bytes of the float32 linear depth reinterpreted as RGBA8, fov, modelview}."""
from __future__ import annotations

import io

import numpy as np
import torch
from PIL import Image

from ..utils import edict


def parse_int_list(int_list_str):
    ints = []
    for s in int_list_str.split(","):
        if "-" in s:
            a, b = s.split("-")
            ints += list(range(int(a), int(b) + 1))
        else:
            ints.append(int(s))
    return ints


def reorder(data, order="3x9"):
    data = list(data)
    if order != "3x9":
        raise NotImplementedError
    if len(data) == 26:
        data.insert(0, -torch.ones_like(data[0]))
    idx = [23, 17, 11, 5, 2, 8, 14, 20, 26, 21, 15, 9, 3, 0, 6, 12, 18, 24, 22, 16, 10, 4, 1, 7, 13, 19, 25]
    return torch.stack([data[i] for i in idx], dim=0)


def _png(arr: np.ndarray) -> bytes:
    with io.BytesIO() as f:
        Image.fromarray(arr).save(f, format="png")
        return f.getvalue()


def _unpng(b: bytes) -> np.ndarray:
    return np.array(Image.open(io.BytesIO(b)))


def _store_modelview(mv):
    from ..rgbd_3d.glm_compat import as_matrix
    m = as_matrix(mv)          # PyGLM mat4 (column-major indexing) or numpy / nested lists (row-major) -> m[row][col]
    try:                       # keep scenes loadable by the reference's render.py when PyGLM is installed
        import glm
        return glm.mat4(*[float(m[r][c]) for c in range(4) for r in range(4)])
    except ImportError:
        return m


def save_scene(path, meshes, colors):
    data = []
    for mesh, col in zip(meshes, colors):
        c8 = np.clip(np.asarray(col) * 255, 0, 255).astype(np.uint8)
        n = mesh.depth.shape[0]
        d = np.ascontiguousarray(np.asarray(mesh.depth).astype(np.float32))
        d8 = np.frombuffer(d.tobytes(), dtype=np.uint8).reshape(n, n, 4)
        data.append({"color": _png(c8), "depth": _png(d8), "fov": mesh.fov, "modelview": _store_modelview(mesh.modelview)})
    np.savez_compressed(path, data=np.array(data, dtype=object))


def load_scene_views(path):
    """Host-only decode of a scene file -> list of edict(color [H,W,3] float in [0,1], depth [H,W,1] float32 linear, fov,
    modelview)."""
    data = np.load(path, allow_pickle=True)["data"]
    out = []
    for d in data:
        col = _unpng(d["color"]) / 255
        n = col.shape[0]
        depth = np.frombuffer(np.ascontiguousarray(_unpng(d["depth"])).tobytes(), dtype=np.float32).reshape(n, n, 1)
        out.append(edict(color=col, depth=depth, fov=d["fov"], modelview=d["modelview"]))
    return out


def load_scene(path, atol=0.03, rtol=0.03, erode_rgb=3):
    """(meshes, colors) as the reference's load_scene (inference/utils.py:104-113): every stored view is re-meshed with
    depth_to_mesh(depth, 32, fov, modelview, atol, rtol, erode_rgb, cal_normal=True) — numeric padding, for free-view
    rendering.  The meshing runs on the GPU (rgbd_3d.utils.depth_to_mesh)."""
    from ..rgbd_3d import utils as r3d
    views = load_scene_views(path)
    meshes = [r3d.depth_to_mesh(v.depth, 32, v.fov, v.modelview, atol=atol, rtol=rtol, erode_rgb=erode_rgb, cal_normal=True) for v in views]
    return meshes, [v.color for v in views]


def colorize_depth(depth, min=-1, max=1):      # noqa: A002 - the reference's argument names (inference/utils.py:25)
    """Inferno colour map of a depth image or batch (numpy HW / NHW or torch), near = bright; output in [min, max]."""
    import cv2
    is_tensor = isinstance(depth, torch.Tensor)
    d = depth.detach().cpu().numpy() if is_tensor else np.asarray(depth)
    d = d.squeeze()
    if d.ndim == 2:
        d = d[None]
    d = np.clip(1 - (d - min) / (max - min), 0, 1)
    maps = [cv2.cvtColor(cv2.applyColorMap((img * 255).astype(np.uint8), cv2.COLORMAP_INFERNO), cv2.COLOR_BGR2RGB) for img in d]
    out = np.stack(maps, axis=0) / 255
    if is_tensor:
        out = torch.from_numpy(out).permute(0, 3, 1, 2).float()
    return (out * (max - min) + min).squeeze()
