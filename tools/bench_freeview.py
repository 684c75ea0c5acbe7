"""Measurement of the free-view fusion rendering row (reference inference/render.py): one synthetic 27-view scene
(viewset 3x9), re-meshed by load_scene and rendered from a 60-frame swing trajectory at 640x640 (5x SSAA) + resolve to
128x128, through the public API (numpy in / out per frame, as the reference's renderer returns them).
Also times the CPU oracle (software rasteriser) on the same scene for 2 frames.   -> one JSON line."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ivid_b200.rgbd_3d as rgbd_3d
from ivid_b200.inference import build_modelviews, load_scene, load_scene_views, save_scene, swing_trajectory
from ivid_b200.inference.render import SSAA, resolve_frame
from ivid_b200.utils import edict
from oracle import warp_ref                      # measurement tooling only (CPU leg)

def synth(rng, n=128):
    yy, xx = np.mgrid[0:n, 0:n] / n
    z = 0.55 + 0.08 * np.sin(6.0 * xx + rng.uniform(0, 6)) * np.cos(5.0 * yy + rng.uniform(0, 6))
    cx, cy, r = rng.uniform(0.35, 0.65), rng.uniform(0.35, 0.65), rng.uniform(0.15, 0.25)
    z = np.where((xx - cx) ** 2 + (yy - cy) ** 2 < r ** 2, z - 0.18, z)
    rgb = np.stack([0.5 + 0.5 * np.sin(9 * xx + i) * np.cos(7 * yy - i) for i in range(3)], axis=-1)
    return rgb.astype(np.float32), z[..., None].astype(np.float32)

frames = int(os.environ.get("FRAMES", 60))
rng = np.random.default_rng(0)
mvs = build_modelviews("3x9", 1)
views, colors = [], []
for mv in mvs:
    rgb, z = synth(rng)
    views.append(edict(depth=warp_ref.linearize_depth(z, 0.6, 5.0).astype(np.float32), fov=45, modelview=np.asarray(mv, dtype=np.float32)))
    colors.append(rgb)
path = os.path.join(tempfile.mkdtemp(), "scene.npz")
save_scene(path, views, colors)
t0 = time.perf_counter(); meshes, cols = load_scene(path); torch.cuda.synchronize(); t_load = time.perf_counter() - t0
r = rgbd_3d.AggregationRenderer(128 * SSAA, 128, near=0.1, far=200)
traj = swing_trajectory(frames)
r.render(meshes, cols, traj[:2])                 # warm-up (uploads, first launches)
torch.cuda.synchronize(); t0 = time.perf_counter()
res = r.render(meshes, cols, traj)
torch.cuda.synchronize(); t_render = time.perf_counter() - t0
t0 = time.perf_counter(); out = [resolve_frame(f, 128) for f in res]; t_resolve = time.perf_counter() - t0
from ivid_b200.inference.render import depth_colour_table
r.render_resolved(meshes, cols, traj[:2], lut=depth_colour_table())
torch.cuda.synchronize(); t0 = time.perf_counter()
cd, dd = r.render_resolved(meshes, cols, traj, lut=depth_colour_table())
torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
same = bool(np.array_equal(cd[1], out[1][0]) and np.array_equal(dd[1], out[1][1]))
# CPU oracle leg: same scene, 2 frames
stored = load_scene_views(path)
t0 = time.perf_counter()
ms_ref = [warp_ref.depth_to_mesh(v.depth, fov=v.fov, modelview=np.asarray(v.modelview), atol=0.03, rtol=0.03, erode_rgb=3, padding=32) for v in stored]
t_mesh_cpu = time.perf_counter() - t0
sw = warp_ref.SoftwareAggregationRenderer(128 * SSAA, 128, near=0.1, far=200)
t0 = time.perf_counter()
for t in traj[:2]:
    ref = sw.render(ms_ref, [v.color for v in stored], t)
t_cpu = (time.perf_counter() - t0) / 2
agree = float((ref.mask_color == res[1]["mask_color"]).mean())
print(json.dumps({"workload": f"free-view rendering, 27 source views, {frames}-frame swing, 640x640 (5x SSAA) -> 128x128",
                  "load_scene_remesh_ms": t_load * 1e3, "render_ms_per_frame": t_render / frames * 1e3,
                  "resolve_ms_per_frame_host": t_resolve / frames * 1e3, "frames_per_s_render": frames / t_render,
                  "render_plus_device_resolve_ms_per_frame": t_dev / frames * 1e3, "frames_per_s_resolved_on_device": frames / t_dev,
                  "device_resolve_equals_host": same,
                  "cpu_oracle": {"remesh_ms": t_mesh_cpu * 1e3, "render_ms_per_frame": t_cpu * 1e3, "cores": 1, "kind": "port"},
                  "mask_agreement_frame1": agree, "coverage": float(res[1]["mask_color"].mean())}))
