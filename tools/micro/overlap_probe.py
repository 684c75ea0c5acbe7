"""Probe: does splitting the CFG batch of the L model into independent halves on two streams hide the HBM-bound GroupNorm
passes of one half under the tensor-bound convs of the other?  Times one N-sample forward against `parts` forwards of N/parts
samples issued on separate streams (separate handles: each owns its plan workspace).

    python tools/micro/overlap_probe.py [N=32] [parts=2]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import ivid_b200.backbones as backbones
from oracle import unet_ref   # synthetic weights only

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench.MODELS["L"]
sd = unet_ref.make_synthetic_state_dict(cfg, seed=1234)
nets = []
for _ in range(parts + 1):
    net = backbones.AdmUnet2d(**cfg); net.load_state_dict(sd); net = net.cuda(); net.repack(); nets.append(net)
x = torch.randn(N, 4, 128, 128, device="cuda"); t = torch.full((N,), 500, device="cuda"); c = torch.arange(N, device="cuda") % 1000
streams = [torch.cuda.Stream() for _ in range(parts)]
h = N // parts
xs = [x[i * h:(i + 1) * h].contiguous() for i in range(parts)]
ts = [t[i * h:(i + 1) * h].contiguous() for i in range(parts)]
cs = [c[i * h:(i + 1) * h].contiguous() for i in range(parts)]


def whole():
    return nets[parts](x, t, c)


def split(stagger=False):
    cur = torch.cuda.current_stream()
    outs = []
    for i, s in enumerate(streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(nets[i](xs[i], ts[i], cs[i]))
    for s in streams:
        cur.wait_stream(s)
    return outs


def serial():
    return [nets[i](xs[i], ts[i], cs[i]) for i in range(parts)]


def timeit(fn, iters=20, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ref = whole()
got = torch.cat(split())
res = {"N": N, "parts": parts, "max_abs_diff_split_vs_whole": float((ref - got).abs().max()),
       "whole_ms": timeit(whole), "split_streams_ms": timeit(split), "split_serial_ms": timeit(serial), "whole_ms_again": timeit(whole)}
print(json.dumps(res))
