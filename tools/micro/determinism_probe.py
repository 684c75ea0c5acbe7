"""Run-to-run determinism of the L forward at batch N, eagerly replayed (IVID_NO_GRAPH=1) or through the CUDA graph, with
per-block taps to localise the first layer whose output differs between two runs.

    [IVID_NO_GRAPH=1] python tools/micro/determinism_probe.py [N=32] [runs=4] [taps=1]
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
cfg = bench.MODELS["L"]
net = backbones.AdmUnet2d(**cfg); net.load_state_dict(unet_ref.make_synthetic_state_dict(cfg, seed=1234)); net = net.cuda(); net.repack()
g = torch.Generator().manual_seed(3)
x = torch.randn(N, 4, 128, 128, generator=g).cuda(); t = torch.full((N,), 500, device="cuda"); c = torch.arange(N, device="cuda") % 1000
blocks, _ = unet_ref._topology(cfg)
names = ["input_blocks.0.0"] + [l[1] for b in blocks for l in b["layers"] if l[0] in ("res", "attn")]


def tap(name):
    L = _lib.lib()
    C, H, W = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.ivid_unet_debug_tap(net._handle, N, name.encode(), None, 0, ctypes.byref(C), ctypes.byref(H), ctypes.byref(W)))
    out = torch.empty((N, C.value, H.value, W.value), dtype=torch.float32)
    _lib.check(L.ivid_unet_debug_tap(net._handle, N, name.encode(), _lib.ptr(out), out.numel(), None, None, None))
    return out


outs, hashes, kept = [], [], []
out_buf = torch.empty(N, 4, 128, 128, device="cuda")
for r in range(runs):
    e = net(x, t, c)
    outs.append(e.cpu())
    if want_taps and r in (0, runs - 1):
        h, k = {}, {}
        for nm in names:
            tt = tap(nm)
            h[nm] = hashlib.md5(tt.numpy().tobytes()).hexdigest()
            if tt.numel() * 4 <= 80e6 or nm in names[:4]:
                k[nm] = tt
        hashes.append(h); kept.append(k)
res = {"N": N, "graph": os.environ.get("IVID_NO_GRAPH") is None, "slab": os.environ.get("IVID_SLAB"),
       "eps_max_abs_diff_vs_run0": [float((o - outs[0]).abs().max()) for o in outs],
       "eps_rel_vs_run0": [float((o.double() - outs[0].double()).norm() / outs[0].double().norm()) for o in outs],
       "samples_differing_last_vs_run0": [int(i) for i in range(N) if not torch.equal(outs[-1][i], outs[0][i])]}
if want_taps:
    diff = [nm for nm in names if hashes[0][nm] != hashes[1][nm]]
    res["first_differing_taps"] = diff[:6]
    res["n_differing_taps"] = len(diff)
    for nm in diff[:3]:
        if nm in kept[0]:
            a, b = kept[0][nm], kept[1][nm]
            d = (a - b).abs()
            idx = torch.nonzero(d.flatten(1).amax(1) > 0).flatten().tolist()
            res["tap_" + nm] = {"max_abs": float(d.max()), "n_diff": int((d > 0).sum()), "numel": d.numel(), "samples": idx[:8],
                                "rel": float((a.double() - b.double()).norm() / a.double().norm())}
print(json.dumps(res))
