"""GPU parity of the individual sm_100a kernels against fp32 torch on the CPU (through the op-level C ABI).

Tolerances: tensor-core operands are fp16 (10-bit mantissa, like the TF32 path the reference's cuDNN convs took on
A100) with fp32 accumulation, so a single conv/GEMM is compared at 2e-3 relative L2 against the fp32 result of the SAME
fp16-rounded operands at 2e-5 (accumulation-order only)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import gpu_util as G

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.default_rng(seed)


def _t(rng, *shape, scale=1.0):
    return torch.from_numpy((rng.standard_normal(shape) * scale).astype(np.float32))


CONV_CASES = [
    # N, H, W, Cin, Cout, k
    (2, 16, 16, 64, 64, 3),
    (1, 32, 32, 128, 256, 3),
    (3, 8, 8, 256, 128, 3),      # TN=2 tiles with odd batch (batch tail masked)
    (2, 16, 16, 192, 384, 1),    # 1x1, BN=128
    (1, 4, 4, 64, 64, 3),        # tiny spatial: TN=8 with N=1
    (2, 64, 64, 64, 16, 3),      # narrow Cout (BN=16 path, output head shape)
    (1, 128, 128, 256, 256, 3),  # the dominant shape of the large model
    (2, 64, 64, 64, 768, 3),     # 96 tile pairs on 74 SM pairs: the last round runs as half-width items
    # cluster-multicast kernel (low-resolution levels, N = 128 tiles): clusters of mc_m pixel tiles x mc_n column blocks
    (4, 8, 8, 256, 512, 3),      # 2 x 4 cluster (8 CTAs): activation slices of 32 pixels, weight slices of 64 rows
    (8, 8, 8, 128, 256, 3),      # 2 x 2 cluster, 2 cluster items per cluster row
    (32, 8, 8, 1024, 1024, 3),   # the 8x8 level of the large model at the benchmark batch: 16 clusters of 8
    (6, 16, 16, 128, 384, 1),    # 16x16 tiles (slices along rows), 3 column blocks: 2 x 1 cluster
    (4, 16, 16, 64, 640, 3),     # 5 column blocks (odd): 2 x 1 cluster, weight multicast only
    # 3x3 tap-reuse kernel (CTA pairs, 8 x 16 pixel tiles)
    (5, 32, 32, 192, 512, 3),    # odd batch, 3 channel chunks, 80 tiles = 40 pairs x 2 column blocks
    (13, 16, 16, 128, 768, 3),   # 16x16 images: one tile row per image (slab rows -1 and 16 are zero fill), 3 column blocks
]


@pytest.fixture(params=["default", "multicast", "slab", "contig"])
def conv_mode(request, monkeypatch):
    """Every conv case runs through the default kernels, through the opt-in cluster-multicast kernel (IVID_MC=1 is read when
    a launch is created; it only takes effect on low-resolution N = 128 layers) and through the 3x3 tap-reuse ("slab") kernel
    (IVID_SLAB=1, CTA-pair layers with H >= 16); "contig" = contiguous work ranges per CTA (IVID_CONV_CONTIG_ALL=1) instead of the
    default round-robin schedule."""
    monkeypatch.delenv("IVID_MC", raising=False)
    monkeypatch.delenv("IVID_SLAB", raising=False)
    monkeypatch.delenv("IVID_CONV_CONTIG_ALL", raising=False)
    if request.param == "multicast":
        monkeypatch.setenv("IVID_MC", "1")
    elif request.param == "slab":
        monkeypatch.setenv("IVID_SLAB", "1")
    elif request.param == "contig":
        monkeypatch.setenv("IVID_CONV_CONTIG_ALL", "1")
    return request.param


@pytest.mark.parametrize("N,H,W,Cin,Cout,k", CONV_CASES)
def test_conv_matches_torch(N, H, W, Cin, Cout, k, conv_mode):
    if conv_mode == "multicast" and H > 16:
        pytest.skip("multicast mode only changes low-resolution layers")
    if conv_mode == "slab" and (k != 3 or H < 16 or Cout % 256 != 0):
        pytest.skip("slab mode only changes 3x3 CTA-pair layers")
    rng = _rng(hash((N, H, W, Cin, Cout, k)) % 2**31)
    x = _t(rng, N, Cin, H, W)
    w = _t(rng, Cout, Cin, k, k, scale=1 / math.sqrt(Cin * k * k))
    b = _t(rng, Cout, scale=0.1)
    xh = x.half()
    ref16 = F.conv2d(xh.float(), w.half().float(), b, padding=k // 2)       # same rounded operands, fp32 math
    ref32 = F.conv2d(x, w, b, padding=k // 2)
    out = G.conv2d(xh.permute(0, 2, 3, 1).contiguous().cuda(), w, b, k)
    got = out.permute(0, 3, 1, 2).cpu()
    r16 = G.report(f"conv N{N} {H}x{W} {Cin}->{Cout} k{k} (vs fp16-rounded operands)", got, ref16)
    r32 = G.report(f"conv N{N} {H}x{W} {Cin}->{Cout} k{k} (vs fp32)", got, ref32)
    assert r16 < 2e-5
    assert r32 < 2e-3


def test_conv_skip_segment_residual_and_fp16_out():
    """out_layers conv + 1x1 skip conv as extra K slabs (ResBlock tail, adm.py:222), identity residual, fp16 output."""
    rng = _rng(5)
    N, H, W, C, Cx = 2, 16, 16, 128, 192
    a = _t(rng, N, C, H, W); x = _t(rng, N, Cx, H, W)
    w = _t(rng, C, C, 3, 3, scale=1 / math.sqrt(9 * C)); b = _t(rng, C, scale=0.1)
    ws = _t(rng, C, Cx, 1, 1, scale=1 / math.sqrt(Cx)); bs = _t(rng, C, scale=0.1)
    ref = F.conv2d(a.half().float(), w.half().float(), b, padding=1) + F.conv2d(x.half().float(), ws.half().float(), bs)
    out = G.conv2d(a.half().permute(0, 2, 3, 1).contiguous().cuda(), w, b, 3,
                   act2=x.half().permute(0, 2, 3, 1).contiguous().cuda(), w2=ws, b2=bs)
    assert G.report("conv3x3 + 1x1 skip segment", out.permute(0, 3, 1, 2), ref) < 2e-5
    res = _t(rng, N, C, H, W)
    ref2 = F.conv2d(a.half().float(), w.half().float(), b, padding=1) + res
    out2 = G.conv2d(a.half().permute(0, 2, 3, 1).contiguous().cuda(), w, b, 3, residual=res.permute(0, 2, 3, 1).contiguous().cuda())
    assert G.report("conv3x3 + identity residual", out2.permute(0, 3, 1, 2), ref2) < 2e-5
    out3 = G.conv2d(a.half().permute(0, 2, 3, 1).contiguous().cuda(), w, b, 3, out_fp16=True)
    assert G.report("conv3x3 fp16 out", out3.float().permute(0, 3, 1, 2), F.conv2d(a.half().float(), w.half().float(), b, padding=1)) < 5e-4


def test_conv_slab_segments_residual_and_fp16_out(monkeypatch):
    """3x3 tap-reuse kernel with a 1x1 skip segment over a second tensor (mixed 9-tap / 1-tap segments share the rings), a
    two-segment 3x3 input (virtual concat), the identity residual and the fp16 output; same shapes through the default kernel."""
    rng = _rng(31)
    N, H, W, C, Cx = 4, 64, 64, 128, 192
    Co = 512
    a = _t(rng, N, C, H, W); x = _t(rng, N, Cx, H, W); res = _t(rng, N, Co, H, W)
    w = _t(rng, Co, C, 3, 3, scale=1 / math.sqrt(9 * C)); b = _t(rng, Co, scale=0.1)
    ws = _t(rng, Co, Cx, 1, 1, scale=1 / math.sqrt(Cx)); bs = _t(rng, Co, scale=0.1)
    base = F.conv2d(a.half().float(), w.half().float(), b, padding=1)
    skip = F.conv2d(x.half().float(), ws.half().float(), bs)
    an = a.half().permute(0, 2, 3, 1).contiguous().cuda()
    xn = x.half().permute(0, 2, 3, 1).contiguous().cuda()
    rn = res.permute(0, 2, 3, 1).contiguous().cuda()
    for mode in ("1", None):
        if mode:
            monkeypatch.setenv("IVID_SLAB", mode)
        else:
            monkeypatch.delenv("IVID_SLAB", raising=False)
        tag = "slab" if mode else "default"
        out = G.conv2d(an, w, b, 3, act2=xn, w2=ws, b2=bs)
        assert G.report(f"{tag}: conv3x3 + 1x1 skip segment", out.permute(0, 3, 1, 2), base + skip) < 2e-5
        out2 = G.conv2d(an, w, b, 3, residual=rn)
        assert G.report(f"{tag}: conv3x3 + identity residual", out2.permute(0, 3, 1, 2), base + res) < 2e-5
        out3 = G.conv2d(an, w, b, 3, out_fp16=True)
        assert G.report(f"{tag}: conv3x3 fp16 out", out3.float().permute(0, 3, 1, 2), base) < 5e-4


@pytest.mark.parametrize("N,H,W,Cin,C", [(2, 64, 64, 64, 768), (4, 64, 64, 128, 512), (2, 16, 16, 128, 128), (1, 128, 128, 64, 256)])
def test_conv_residual_three_tiles_in_flight(monkeypatch, N, H, W, Cin, C):
    """IVID_RES3=1: the fp32 epilogue keeps three residual tiles in flight per warp (single output staging tile): CTA pairs with a
    split tail, plain pairs, single-CTA N = 128 tiles; identical bits to the default epilogue."""
    rng = _rng(N * 1000 + C)
    a = _t(rng, N, Cin, H, W); res = _t(rng, N, C, H, W)
    w = _t(rng, C, Cin, 3, 3, scale=1 / math.sqrt(9 * Cin)); b = _t(rng, C, scale=0.1)
    an = a.half().permute(0, 2, 3, 1).contiguous().cuda(); rn = res.permute(0, 2, 3, 1).contiguous().cuda()
    monkeypatch.delenv("IVID_RES3", raising=False)
    base = G.conv2d(an, w, b, 3, residual=rn).clone()
    monkeypatch.setenv("IVID_RES3", "1")
    got = G.conv2d(an, w, b, 3, residual=rn)
    assert torch.equal(got, base)
    ref = F.conv2d(a.half().float(), w.half().float(), b, padding=1) + res
    assert G.report(f"res3: conv3x3 + residual N{N} {H}x{W} {Cin}->{C}", got.permute(0, 3, 1, 2), ref) < 2e-5


def test_conv_multicast_residual_stats_paths(monkeypatch):
    """Cluster-multicast kernel through the residual-prefetch epilogue, the fp16-output epilogue and the 1x1 skip segment."""
    monkeypatch.setenv("IVID_MC", "1")
    rng = _rng(21)
    N, H, W, C, Cx = 8, 8, 8, 512, 256
    a = _t(rng, N, C, H, W); x = _t(rng, N, Cx, H, W); res = _t(rng, N, C, H, W)
    w = _t(rng, C, C, 3, 3, scale=1 / math.sqrt(9 * C)); b = _t(rng, C, scale=0.1)
    ws = _t(rng, C, Cx, 1, 1, scale=1 / math.sqrt(Cx)); bs = _t(rng, C, scale=0.1)
    base = F.conv2d(a.half().float(), w.half().float(), b, padding=1)
    an = a.half().permute(0, 2, 3, 1).contiguous().cuda()
    out = G.conv2d(an, w, b, 3, residual=res.permute(0, 2, 3, 1).contiguous().cuda())
    assert G.report("multicast: conv3x3 + residual", out.permute(0, 3, 1, 2), base + res) < 2e-5
    out16 = G.conv2d(an, w, b, 3, out_fp16=True)
    assert G.report("multicast: conv3x3 fp16 out", out16.float().permute(0, 3, 1, 2), base) < 5e-4
    outs = G.conv2d(an, w, b, 3, act2=x.half().permute(0, 2, 3, 1).contiguous().cuda(), w2=ws, b2=bs)
    assert G.report("multicast: conv3x3 + 1x1 skip segment", outs.permute(0, 3, 1, 2),
                    base + F.conv2d(x.half().float(), ws.half().float(), bs)) < 2e-5


def test_conv_split_tail_residual_and_fp16_out():
    """Layer whose persistent grid ends in a partial round (3 x 32 tile pairs on 74 SM pairs): half-width tail items
    through the residual-prefetch epilogue and the fp16-output epilogue."""
    rng = _rng(11)
    N, H, W, Cin, C = 2, 64, 64, 64, 768
    a = _t(rng, N, Cin, H, W)
    w = _t(rng, C, Cin, 3, 3, scale=1 / math.sqrt(9 * Cin)); b = _t(rng, C, scale=0.1)
    res = _t(rng, N, C, H, W)
    base = F.conv2d(a.half().float(), w.half().float(), b, padding=1)
    an = a.half().permute(0, 2, 3, 1).contiguous().cuda()
    out = G.conv2d(an, w, b, 3, residual=res.permute(0, 2, 3, 1).contiguous().cuda())
    assert G.report("split tail: conv3x3 + residual", out.permute(0, 3, 1, 2), base + res) < 2e-5
    out16 = G.conv2d(an, w, b, 3, out_fp16=True)
    assert G.report("split tail: conv3x3 fp16 out", out16.float().permute(0, 3, 1, 2), base) < 5e-4


@pytest.mark.parametrize("N,T,C", [(32, 256, 768), (32, 1024, 512), (8, 4096, 256)])
def test_attention_is_deterministic_under_load(N, T, C):
    """Same qkv, many launches with every SM busy: all outputs bitwise equal and equal to the fp32 reference within tolerance.
    Regression test for a barrier-parity alias of the double-buffered kernel (a softmax warp two key blocks ahead of the P V
    pipe passed its wait for the accumulator one phase early: ~10 % of the launches returned a few wrong 32-row groups)."""
    g = torch.Generator().manual_seed(T)
    qkv = (torch.randn(N, T, 3 * C, generator=g) * 1.5).half().cuda()
    ref = G.attention(qkv, C).clone()
    iters = 400 if T < 4096 else 60
    bad = 0
    for _ in range(iters):
        bad += 0 if torch.equal(G.attention(qkv, C), ref) else 1
    assert bad == 0, f"{bad} of {iters} launches differ from the first"
    # and the first one is right (one sample is enough: the op parity test covers the numerics)
    q, k, v = qkv[:1].float().cpu().reshape(1, T, C // 64, 3, 64).permute(3, 0, 2, 1, 4)
    w = torch.softmax(torch.einsum("bhtd,bhsd->bhts", q, k) / 8.0, dim=-1)
    want = torch.einsum("bhts,bhsd->bhtd", w, v).permute(0, 2, 1, 3).reshape(1, T, C)
    assert G.report(f"attention T={T} under load", ref[:1].float().cpu(), want) < 2e-3


GN_CASES = [
    # N, H, W, C0, C1, groups, silu, mode, film
    (2, 16, 16, 64, 0, 32, True, 0, False),
    (2, 16, 16, 128, 0, 32, True, 0, True),
    (3, 8, 8, 1024, 768, 32, True, 0, False),    # 1792 = 1024 (+) 768: 56 channels / group straddles the seam
    (2, 16, 16, 512, 256, 32, True, 0, False),   # 768 = 512 (+) 256: 24 / group straddles
    (2, 16, 16, 128, 0, 32, True, 1, False),     # nearest 2x upsample after GN+SiLU
    (2, 16, 16, 128, 0, 32, True, 2, False),     # 2x2 average pool after GN+SiLU
    (1, 32, 32, 512, 0, 32, False, 0, False),    # attention norm (no SiLU)
]


@pytest.mark.parametrize("N,H,W,C0,C1,groups,silu,mode,film", GN_CASES)
def test_group_norm_matches_torch(N, H, W, C0, C1, groups, silu, mode, film):
    rng = _rng(hash((N, H, W, C0, C1, mode)) % 2**31)
    C = C0 + C1
    x0 = _t(rng, N, C0, H, W) * 1.7 + 0.3
    x1 = (_t(rng, N, C1, H, W) * 0.6 - 0.2) if C1 else None
    gamma = 1 + 0.1 * _t(rng, C); beta = 0.1 * _t(rng, C)
    x = torch.cat([x0, x1], 1) if C1 else x0
    y = F.group_norm(x, groups, gamma, beta, 1e-5)
    fl = None
    if film:
        fl = 0.3 * _t(rng, N, 2 * C)
        y = y * (1 + fl[:, :C, None, None]) + fl[:, C:, None, None]
    if silu:
        y = F.silu(y)
    if mode == 1:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    elif mode == 2:
        y = F.avg_pool2d(y, 2)
    out = G.group_norm(x0.permute(0, 2, 3, 1).contiguous().cuda(), x1.permute(0, 2, 3, 1).contiguous().cuda() if C1 else None,
                       groups, gamma, beta, fl.cuda() if film else None, silu, mode)
    r = G.report(f"group_norm N{N} {H}x{W} C{C0}+{C1} silu{silu} mode{mode} film{film}", out.float().permute(0, 3, 1, 2), y)
    assert r < 6e-4     # output is rounded to fp16 (2^-11 relative)


@pytest.mark.parametrize("N,T,C", [(2, 64, 128), (1, 256, 192), (2, 1024, 128), (1, 4096, 64)])
def test_attention_matches_torch(N, T, C):
    rng = _rng(T + C)
    qkv = _t(rng, N, 3 * C, T)                     # reference layout [N, 3C, T]
    qh = qkv.half()
    heads = C // 64
    q, k, v = qh.float().reshape(N * heads, 3 * 64, T).split(64, dim=1)
    s = 1 / math.sqrt(math.sqrt(64))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ref = torch.einsum("bts,bcs->bct", w, v).reshape(N, C, T)
    out = G.attention(qh.permute(0, 2, 1).contiguous().cuda(), C)     # [N, T, 3C]
    r = G.report(f"attention N{N} T{T} C{C}", out.float().permute(0, 2, 1), ref)
    assert r < 2e-3
