"""Is a sample's eps independent of the batch it is computed in?  L model: batch of N against N/2 halves and single samples."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import ivid_b200.backbones as backbones
from oracle import unet_ref   # synthetic weights only

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = bench.MODELS[sys.argv[2] if len(sys.argv) > 2 else "L"]
sd = unet_ref.make_synthetic_state_dict(cfg, seed=1234)
net = backbones.AdmUnet2d(**cfg); net.load_state_dict(sd); net = net.cuda(); net.repack()
S = cfg["image_size"]
g = torch.Generator().manual_seed(3)
x = torch.randn(N, cfg["in_channels"], S, S, generator=g).cuda(); t = torch.full((N,), 500, device="cuda")
c = (torch.arange(N, device="cuda") % 1000) if cfg.get("num_classes") else None
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
whole = net(x, t, c).clone()
again = net(x, t, c).clone()
res = {"N": N, "eps_std": float(whole.std()), "eps_absmax": float(whole.abs().max()), "rerun_max_abs": float((whole - again).abs().max())}
h = N // 2
halves = torch.cat([net(x[:h].contiguous(), t[:h], c[:h] if c is not None else None), net(x[h:].contiguous(), t[h:], c[h:] if c is not None else None)])
res["halves_rel"] = rel(halves, whole); res["halves_max_abs"] = float((halves - whole).abs().max())
per = {}
for i in (0, 1, h - 1, h, N - 1):
    one = net(x[i:i + 1].contiguous(), t[i:i + 1], c[i:i + 1] if c is not None else None)
    per[i] = {"rel": rel(one, whole[i:i + 1]), "max_abs": float((one - whole[i:i + 1]).abs().max())}
res["single_vs_whole"] = per
print(json.dumps(res))
