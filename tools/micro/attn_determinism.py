"""Run-to-run determinism of the attention kernel alone: same qkv, `iters` launches, every output compared bitwise with the first.

    [IVID_ATTN_VARIANT=k | IVID_ATTN_V1=1] python tools/micro/attn_determinism.py [iters=3000]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as G

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
res = {"variant": os.environ.get("IVID_ATTN_VARIANT", "0"), "v1": os.environ.get("IVID_ATTN_V1") is not None, "cases": []}
for (N, T, C) in [(32, 256, 768), (32, 1024, 512), (32, 64, 1024), (16, 4096, 256)]:
    g = torch.Generator().manual_seed(T)
    qkv = (torch.randn(N, T, 3 * C, generator=g) * 1.5).half().cuda()
    ref = G.attention(qkv, C).clone()
    bad, worst, rows = 0, 0.0, set()
    n_it = iters if T < 4096 else max(iters // 10, 50)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(n_it):
        out = G.attention(qkv, C)
        if not torch.equal(out, ref):
            bad += 1
            d = (out.float() - ref.float()).abs()
            worst = max(worst, float(d.max()))
            idx = torch.nonzero(d.amax(2) > 0)       # (n, t) pairs
            for n, t in idx[:64].tolist():
                rows.add((n, t // 32))
    ev1.record(); torch.cuda.synchronize()
    res["cases"].append({"N": N, "T": T, "C": C, "iters": n_it, "differing_runs": bad, "max_abs": worst,
                         "row_quarters(n, t//32)": sorted(rows)[:12], "ms_per_call_incl_compare": ev0.elapsed_time(ev1) / n_it})
print(json.dumps(res))
