"""Per-block difference between the GroupNorm-fold path (IVID_FOLD=1) and the separate-apply path on the large model."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import ivid_b200.backbones as backbones
from ivid_b200 import _lib
from oracle import unet_ref

cfg = bench.MODELS["L"]
sd = unet_ref.make_synthetic_state_dict(cfg, seed=1234)
N = 2
g = torch.Generator().manual_seed(9)
x = torch.randn(N, 4, 128, 128, generator=g).cuda(); t = torch.tensor([700, 20]).cuda(); c = torch.tensor([5, -1]).cuda()
blocks, _ = unet_ref._topology(cfg)
names = ["input_blocks.0.0"] + [l[1] for b in blocks for l in b["layers"] if l[0] in ("res", "attn")]


def tap(net, name):
    L = _lib.lib()
    C, H, W = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.ivid_unet_debug_tap(net._handle, N, name.encode(), None, 0, ctypes.byref(C), ctypes.byref(H), ctypes.byref(W)))
    out = torch.empty((N, C.value, H.value, W.value), dtype=torch.float32)
    _lib.check(L.ivid_unet_debug_tap(net._handle, N, name.encode(), _lib.ptr(out), out.numel(), None, None, None))
    return out


def run(fold):
    if fold: os.environ["IVID_FOLD"] = "1"
    else: os.environ.pop("IVID_FOLD", None)
    net = backbones.AdmUnet2d(**cfg); net.load_state_dict(sd); net = net.cuda(); net.repack()
    e = net(x, t, c).cpu()
    return e, {n: tap(net, n) for n in names}


e0, t0 = run(False)
e1, t1 = run(True)
e2, t2 = run(True)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
out = {"eps_fold_vs_apply": rel(e1, e0), "eps_fold_vs_fold": rel(e2, e1), "blocks": []}
for n in names:
    d = (t1[n] - t0[n]).abs()
    out["blocks"].append([n, rel(t1[n], t0[n]), float(d.max()), int((d > 0).sum()), d.numel()])
print(json.dumps(out))
