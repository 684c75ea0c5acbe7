"""Does tcgen05.mma read a 128-byte-swizzled K-major operand correctly when the descriptor start is 1 or 2 rows (128 / 256 B) into
an 8-row swizzle atom?  Runs one 3x3 conv through the tap-reuse kernel in mode 1 (aligned starts), 2 (matrix base offset = row
offset) and 3 (base offset 0) and prints the error against fp32 torch on the same fp16-rounded operands, plus timing.

    python tools/micro/slab_probe.py
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.nn.functional as F
import gpu_util as G

res = []
for (N, H, W, Cin, Cout) in [(4, 64, 64, 128, 512), (2, 128, 128, 256, 256)]:
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((N, Cin, H, W)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((Cout, Cin, 3, 3)) / math.sqrt(9 * Cin)).astype(np.float32))
    b = torch.from_numpy((0.1 * rng.standard_normal(Cout)).astype(np.float32))
    ref = F.conv2d(x.half().float(), w.half().float(), b, padding=1)
    xn = x.half().permute(0, 2, 3, 1).contiguous().cuda()
    for mode in ("0", "1", "2", "3"):
        os.environ["IVID_SLAB"] = mode
        out = G.conv2d(xn, w, b, 3).permute(0, 3, 1, 2).cpu()
        rel = float((out.double() - ref.double()).norm() / ref.double().norm())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): G.conv2d(xn, w, b, 3)
        ev0.record()
        for _ in range(10): G.conv2d(xn, w, b, 3)
        ev1.record(); torch.cuda.synchronize()
        res.append({"shape": [N, H, W, Cin, Cout], "IVID_SLAB": mode, "rel_err": rel, "ms_per_call_incl_plan": ev0.elapsed_time(ev1) / 10})
print(json.dumps(res, indent=1))
