"""Micro-benchmark of the device warp (config 4 shape: batch 16, viewset 3x9): per target view j, time
DeviceWarp.aggregate with j source views (rasterise + deferred shade/aggregate + post-filters) and add_view."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ivid_b200.rgbd_3d import DeviceWarp
from ivid_b200.inference import build_modelviews
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))

def synth(rng, n=128):
    yy, xx = np.mgrid[0:n, 0:n] / n
    z = 0.55 + 0.08 * np.sin(6.0 * xx + rng.uniform(0, 6)) * np.cos(5.0 * yy + rng.uniform(0, 6))
    cx, cy, r = rng.uniform(0.35, 0.65), rng.uniform(0.35, 0.65), rng.uniform(0.15, 0.25)
    blob = (xx - cx) ** 2 + (yy - cy) ** 2 < r ** 2
    z = np.where(blob, z - 0.18, z)
    rgb = np.stack([0.5 + 0.5 * np.sin(9 * xx + i) * np.cos(7 * yy - i) for i in range(3)], axis=-1)
    return np.concatenate([rgb, z[..., None]], axis=-1).astype(np.float32)

B = int(os.environ.get("B", 16)); V = 27
rng = np.random.default_rng(0)
views = build_modelviews("3x9", 1)
x = torch.from_numpy(np.stack([synth(rng).transpose(2, 0, 1) * 2 - 1 for _ in range(B)])).float().cuda()
kw = dict(fov=45, near=0.6, far=5, atol=0.03, rtol=0.03, erode_rgb=3)
w = DeviceWarp(B, image_size=128, ssaa=3, max_views=V)
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for j in range(V):
    if j > 0:
        a, b = ev(), ev(); a.record(); cond = w.aggregate(views[j], **kw); b.record(); torch.cuda.synchronize()
        agg_ms = a.elapsed_time(b)
    else:
        agg_ms = 0.0
    a, b = ev(), ev(); a.record(); w.add_view(x, views[j], **kw); b.record(); torch.cuda.synchronize()
    rows.append((j, agg_ms, a.elapsed_time(b)))
tot_agg = sum(r[1] for r in rows); tot_add = sum(r[2] for r in rows)
# algorithmic bytes: per (sample, target) with j sources: j*(4*128^2*4 + 384^2*8*2) + 7*128^2*4
alg = sum(B * (j * (4 * 128 * 128 * 4 + 384 * 384 * 16) + 7 * 128 * 128 * 4) for j in range(1, V))
out = {"batch": B, "views": V, "aggregate_ms_total": tot_agg, "add_view_ms_total": tot_add,
       "aggregate_ms_by_sources": {str(r[0]): round(r[1], 3) for r in rows if r[0] in (1, 2, 4, 8, 13, 20, 26)},
       "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / 1e9 / (tot_agg / 1e3),
       "warp_ms_per_sample_3x9": (tot_agg + tot_add) / B, "mask_coverage_last": float(cond[:, 4].mean())}
print(json.dumps(out))
