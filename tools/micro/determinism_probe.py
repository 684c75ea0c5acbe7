"""Run-to-run determinism of the L forward at batch N, eagerly replayed (IVID_NO_GRAPH=1) or through the CUDA graph, with
per-block taps to localise the first layer whose output differs between two runs.

    [IVID_NO_GRAPH=1] python tools/micro/determinism_probe.py [N=32] [runs=4] [taps=1] [model=L|Lc|SR|S]
"""
import ctypes, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import ivid_b200.backbones as backbones
from ivid_b200 import _lib
from oracle import unet_ref   # synthetic weights + block names only

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
want_taps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = bench.MODELS[sys.argv[4] if len(sys.argv) > 4 else "L"]
net = backbones.AdmUnet2d(**cfg); net.load_state_dict(unet_ref.make_synthetic_state_dict(cfg, seed=1234)); net = net.cuda(); net.repack()
g = torch.Generator().manual_seed(3)
S = cfg["image_size"]
x = torch.randn(N, cfg["in_channels"], S, S, generator=g).cuda(); t = torch.full((N,), 500, device="cuda")
c = (torch.arange(N, device="cuda") % 1000) if cfg.get("num_classes") else None
blocks, _ = unet_ref._topology(cfg)
names = ["input_blocks.0.0"] + [l[1] for b in blocks for l in b["layers"] if l[0] in ("res", "attn")]


def tap(name):
    L = _lib.lib()
    C, H, W = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.ivid_unet_debug_tap(net._handle, N, name.encode(), None, 0, ctypes.byref(C), ctypes.byref(H), ctypes.byref(W)))
    out = torch.empty((N, C.value, H.value, W.value), dtype=torch.float32)
    _lib.check(L.ivid_unet_debug_tap(net._handle, N, name.encode(), _lib.ptr(out), out.numel(), None, None, None))
    return out


def all_taps(keep):
    h, k = {}, {}
    for nm in names:
        tt = tap(nm)
        h[nm] = hashlib.md5(tt.numpy().tobytes()).hexdigest()
        if keep and (tt.numel() * 4 <= 140e6 or nm in names[:3]):
            k[nm] = tt
    return h, k


e0 = net(x, t, c).cpu()
h0, k0 = all_taps(True) if want_taps else ({}, {})
res = {"N": N, "graph": os.environ.get("IVID_NO_GRAPH") is None, "slab": os.environ.get("IVID_SLAB"), "runs": []}
for r in range(1, runs):
    e = net(x, t, c).cpu()
    d = float((e - e0).abs().max())
    res["runs"].append(d)
    if d > 0:
        res.setdefault("differing_md5", []).append(hashlib.md5(e.numpy().tobytes()).hexdigest()[:8])
    if d > 0 and want_taps:
        h1, _ = all_taps(False)
        diff = [nm for nm in names if h0[nm] != h1[nm]]
        res["differing_run"] = r
        res["samples_differing"] = [int(i) for i in range(N) if not torch.equal(e[i], e0[i])]
        res["first_differing_taps"] = diff[:8]
        res["n_differing_taps"] = len(diff)
        for nm in diff[:4]:
            if nm in k0:
                a, b = k0[nm], tap(nm)
                dd = (a - b).abs()
                idx = torch.nonzero(dd.flatten(1).amax(1) > 0).flatten().tolist()
                chans = torch.nonzero(dd.amax((0, 2, 3)) > 0).flatten().tolist()
                res["tap_" + nm] = {"max_abs": float(dd.max()), "n_diff": int((dd > 0).sum()), "numel": dd.numel(), "samples": idx[:8],
                                    "n_channels": len(chans), "channels": chans[:16], "rel": float((a.double() - b.double()).norm() / a.double().norm())}
        break
print(json.dumps(res))
