"""Small-batch latency with and without the CUDA graph of the forward: BASELINE configs[0] (single-category
small model, batch 1, 10 DDIM steps) and one guided DDIM step of the conditional large model at batch 1 (a batch-2 forward).

    [IVID_NO_GRAPH=1] python tools/micro/latency_small_batch.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import ivid_b200.backbones as backbones, ivid_b200.frameworks as frameworks, ivid_b200.samplers as samplers
from oracle import unet_ref   # synthetic weights only

g = np.load(os.path.join(ROOT, "tests", "golden", "unet_sampler_golden.npz"))
cfgS = json.loads(bytes(g["schemacfg_rgbd_singlecategory_adm_128_small"]).decode())


def net_of(cfg, seed):
    net = backbones.AdmUnet2d(**cfg); net.load_state_dict(unet_ref.make_synthetic_state_dict(cfg, seed=seed)); net = net.cuda(); net.repack(); return net


def timed(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"graph": os.environ.get("IVID_NO_GRAPH") is None}
fwS = frameworks.GaussianDiffusion(net_of(cfgS, 1), timesteps=1000, beta_schedule="linear")
sS = samplers.DdimSampler(fwS)
res["config1_small_batch1_ddim10_ms_per_sample"] = timed(lambda: sS.sample(1, steps=10, verbose=False), 10)
fwC = frameworks.InpaintCFG(net_of(bench.MODELS["Lc"], 2), timesteps=1000, beta_schedule="linear")
sC = samplers.DdimSampler(fwC)
x = torch.randn(1, 4, 128, 128, device="cuda"); y = torch.randn(1, 4, 128, 128, device="cuda"); m = (torch.rand(1, 1, 128, 128, device="cuda") > 0.5).float()
t = torch.tensor([500], device="cuda"); tp = torch.tensor([480], device="cuda"); c = torch.tensor([7], device="cuda")
res["cond_large_batch1_guided_ddim_step_ms"] = timed(lambda: sC.sample_once(x, t, tp, c, strength=0.5, y=y, mask=m, mask_rgb=m), 30)
print(json.dumps(res))
