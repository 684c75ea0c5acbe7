"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of the reference's ADM UNet forward, written against a plain state-dict (no nn.Module tree):
    reference: diffusion/backbones/adm.py
        PosEncoding.forward        :28-33      time_embed / label_emb   :357-365, 545-555
        GroupNorm32                :36-41      ResBlock2d.forward       :192-222
        QKVAttention.forward       :233-253    AttentionBlock.forward   :280-286
        AdmUnet2d.__init__ (block topology) :367-487    AdmUnet2d.forward :526-566
It is pinned against the reference modules themselves (imported from /root/reference on the build container) by
tests/golden/make_golden.py; the resulting fixtures are committed under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["unet_param_shapes", "make_synthetic_state_dict", "unet_forward", "UnetOracle"]


def _cfg_defaults(cfg: dict) -> dict:
    c = dict(
        dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, num_classes=None, has_null_class=False,
        use_fp16=False, num_groups=32, num_heads=1, num_head_channels=-1, use_scale_shift_norm=True,
        resblock_updown=True,
    )
    c.update({k: v for k, v in cfg.items() if v is not None or k in ("num_classes",)})
    if c.get("num_heads") is None:
        c["num_heads"] = 1
    return c


def _topology(cfg: dict):
    """Block structure of AdmUnet2d.__init__ (adm.py:367-487) as plain data.

    Returns (blocks, final_ch) where blocks is a list of dicts:
        {"group": "input"|"middle"|"output", "prefix": str, "layers": [("res", prefix, cin, cout, mode) | ("attn", prefix, ch)]}
    """
    c = _cfg_defaults(cfg)
    mc = c["model_channels"]
    mult = list(c["channel_mult"])
    nres = c["num_res_blocks"]
    attn_res = list(c["attention_resolutions"])
    blocks = [{"group": "input", "prefix": "input_blocks.0", "layers": [("conv", "input_blocks.0.0", c["in_channels"], int(mult[0] * mc))]}]
    ch = int(mult[0] * mc)
    chs = [ch]
    ds = c["image_size"]
    ib = 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            layers = [("res", f"input_blocks.{ib}.0", ch, int(m * mc), "same")]
            ch = int(m * mc)
            if ds in attn_res:
                layers.append(("attn", f"input_blocks.{ib}.1", ch))
            blocks.append({"group": "input", "prefix": f"input_blocks.{ib}", "layers": layers})
            chs.append(ch)
            ib += 1
        if level != len(mult) - 1:
            # adm.py:398-414: a "down" ResBlock, or (resblock_updown=False) a plain Downsample2d
            down = ("res", f"input_blocks.{ib}.0", ch, ch, "down") if c["resblock_updown"] else ("down", f"input_blocks.{ib}.0", ch)
            blocks.append({"group": "input", "prefix": f"input_blocks.{ib}", "layers": [down]})
            chs.append(ch)
            ib += 1
            ds //= 2
    blocks.append({"group": "middle", "prefix": "middle_block", "layers": [
        ("res", "middle_block.0", ch, ch, "same"), ("attn", "middle_block.1", ch), ("res", "middle_block.2", ch, ch, "same")]})
    ob = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nres + 1):
            ich = chs.pop()
            li = 0
            layers = [("res", f"output_blocks.{ob}.{li}", ch + ich, int(mc * m), "same")]
            li += 1
            ch = int(mc * m)
            if ds in attn_res:
                layers.append(("attn", f"output_blocks.{ob}.{li}", ch))
                li += 1
            if level and i == nres:
                # adm.py:463-480: an "up" ResBlock, or (resblock_updown=False) a plain Upsample2d
                layers.append(("res", f"output_blocks.{ob}.{li}", ch, ch, "up") if c["resblock_updown"] else ("up", f"output_blocks.{ob}.{li}", ch))
                ds *= 2
            blocks.append({"group": "output", "prefix": f"output_blocks.{ob}", "layers": layers})
            ob += 1
    return blocks, ch


def unet_param_shapes(cfg: dict) -> Dict[str, tuple]:
    """state_dict key -> shape, in the reference's registration order (494 keys for the large model)."""
    c = _cfg_defaults(cfg)
    mc = c["model_channels"]
    E = 4 * mc
    out: Dict[str, tuple] = {}
    out["time_embed.0.freqs"] = (mc // 2,)
    out["time_embed.1.weight"] = (E, mc)
    out["time_embed.1.bias"] = (E,)
    out["time_embed.3.weight"] = (E, E)
    out["time_embed.3.bias"] = (E,)
    if c["num_classes"] is not None:
        out["label_emb.weight"] = (c["num_classes"], E)
    blocks, final_ch = _topology(cfg)
    for b in blocks:
        for l in b["layers"]:
            if l[0] == "conv":
                _, p, cin, cout = l
                out[p + ".weight"] = (cout, cin, 3, 3)
                out[p + ".bias"] = (cout,)
            elif l[0] == "res":
                _, p, cin, cout, _mode = l
                out[p + ".in_layers.0.weight"] = (cin,)
                out[p + ".in_layers.0.bias"] = (cin,)
                out[p + ".in_layers.2.weight"] = (cout, cin, 3, 3)
                out[p + ".in_layers.2.bias"] = (cout,)
                ew = 2 * cout if c["use_scale_shift_norm"] else cout         # adm.py:176
                out[p + ".emb_layers.1.weight"] = (ew, E)
                out[p + ".emb_layers.1.bias"] = (ew,)
                out[p + ".out_layers.0.weight"] = (cout,)
                out[p + ".out_layers.0.bias"] = (cout,)
                out[p + ".out_layers.3.weight"] = (cout, cout, 3, 3)
                out[p + ".out_layers.3.bias"] = (cout,)
                if cin != cout:
                    out[p + ".skip_connection.weight"] = (cout, cin, 1, 1)
                    out[p + ".skip_connection.bias"] = (cout,)
            elif l[0] in ("down", "up"):
                _, p, ch = l
                if c["conv_resample"]:      # Downsample2d.op (adm.py:111) / Upsample2d.conv (adm.py:81)
                    sub = ".op" if l[0] == "down" else ".conv"
                    out[p + sub + ".weight"] = (ch, ch, 3, 3)
                    out[p + sub + ".bias"] = (ch,)
            else:
                _, p, ch = l
                out[p + ".norm.weight"] = (ch,)
                out[p + ".norm.bias"] = (ch,)
                out[p + ".qkv.weight"] = (3 * ch, ch, 1)
                out[p + ".qkv.bias"] = (3 * ch,)
                out[p + ".proj_out.weight"] = (ch, ch, 1)
                out[p + ".proj_out.bias"] = (ch,)
    out["out.0.weight"] = (final_ch,)
    out["out.0.bias"] = (final_ch,)
    out["out.2.weight"] = (c["out_channels"], int(list(c["channel_mult"])[0] * mc), 3, 3)
    out["out.2.bias"] = (c["out_channels"],)
    return out


def make_synthetic_state_dict(cfg: dict, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (numpy PCG64 — stable across machines and torch versions).

    The reference zero-initialises the second conv of every ResBlock, every attention proj_out and the final conv
    (adm.py:182,278,486), which makes the default-init network output identically zero.  Parity on that is
    vacuous, so every >=2-D tensor is drawn N(0, 1/fan_in), biases N(0, 0.02^2), norm scales 1 + N(0, 0.1^2).
    """
    rng = np.random.default_rng(seed)
    c = _cfg_defaults(cfg)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in unet_param_shapes(cfg).items():
        if name == "time_embed.0.freqs":
            half = shape[0]
            # adm.py:26 — exp(-log(max_freq) * arange(half) / half) in float32
            v = torch.exp(-np.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
        elif name == "label_emb.weight":
            v = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            v = torch.from_numpy((rng.standard_normal(shape) / math.sqrt(fan_in)).astype(np.float32))
        elif name.endswith("bias"):
            v = torch.from_numpy((0.02 * rng.standard_normal(shape)).astype(np.float32))
        else:  # GroupNorm weight
            v = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32))
        sd[name] = v
    return sd


def _group_norm(x, sd, p, groups):
    # GroupNorm32 (adm.py:36-41): fp32, eps 1e-5
    return F.group_norm(x.float(), groups, sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _resblock(x, emb, sd, p, mode, groups, taps: Optional[dict] = None):
    # ResBlock2d.forward (adm.py:192-222); use_scale_shift_norm is read off the width of emb_layers (adm.py:176)
    h = F.silu(_group_norm(x, sd, p + ".in_layers.0", groups))
    if mode == "up":       # Upsample2d without conv: nearest x2 on both branches (adm.py:89, 203-207)
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif mode == "down":   # Downsample2d without conv: AvgPool2d(2) (adm.py:113)
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])[:, :, None, None]
    if emb_out.shape[1] == 2 * h.shape[1]:          # use_scale_shift_norm (adm.py:214-218)
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = _group_norm(h, sd, p + ".out_layers.0", groups) * (1 + scale) + shift
    else:                                           # adm.py:219-221: h = h + emb_out; out_layers(h)
        h = _group_norm(h + emb_out, sd, p + ".out_layers.0", groups)
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attention(x, sd, p, groups, head_ch):
    # AttentionBlock.forward (adm.py:280-286) + QKVAttention.forward (adm.py:233-253)
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_group_norm(xf, sd, p + ".norm", groups), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    heads = c // head_ch
    T = xf.shape[-1]
    q, k, v = qkv.reshape(b * heads, head_ch * 3, T).split(head_ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(head_ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, T)
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


@torch.no_grad()
def unet_forward(cfg: dict, sd: Dict[str, torch.Tensor], x: torch.Tensor, times: torch.Tensor,
                 classes: Optional[torch.Tensor] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """AdmUnet2d.forward (adm.py:526-566), fp32.  `taps` (optional dict) receives named intermediate tensors."""
    c = _cfg_defaults(cfg)
    groups = c["num_groups"]
    head_ch = c["num_head_channels"]
    assert classes is None or c["num_classes"] is not None, "this model is not class-conditioned"
    if classes is not None:
        assert torch.all(classes >= 0) or c["has_null_class"], "this model does not have a null class"
    # PosEncoding (adm.py:30-33): [cos | sin]
    args = times[:, None] * sd["time_embed.0.freqs"][None, :]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    emb = F.linear(emb, sd["time_embed.1.weight"], sd["time_embed.1.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.3.weight"], sd["time_embed.3.bias"])
    if c["num_classes"] is not None:
        if classes is not None:
            assert classes.shape == (x.shape[0],)
            ce = sd["label_emb.weight"][classes * (classes >= 0).long()]
            if c["has_null_class"]:
                ce = ce * (classes >= 0).unsqueeze(1)
            emb = emb + ce
        # classes None -> zeros (adm.py:554-555)
    if taps is not None:
        taps["emb"] = emb
    blocks, _ = _topology(cfg)
    hs: List[torch.Tensor] = []
    h = x.float()
    for b in blocks:
        if b["group"] == "output":
            h = torch.cat([h, hs.pop()], dim=1)
        for l in b["layers"]:
            if l[0] == "conv":
                h = F.conv2d(h, sd[l[1] + ".weight"], sd[l[1] + ".bias"], padding=1)
            elif l[0] == "res":
                h = _resblock(h, emb, sd, l[1], l[4], groups)
            elif l[0] == "down":      # Downsample2d.forward (adm.py:115-117): conv 3x3 stride 2 pad 1, or AvgPool2d(2)
                h = (F.conv2d(h, sd[l[1] + ".op.weight"], sd[l[1] + ".op.bias"], stride=2, padding=1) if c["conv_resample"]
                     else F.avg_pool2d(h, 2))
            elif l[0] == "up":        # Upsample2d.forward (adm.py:86-91): nearest x2, then conv 3x3 if conv_resample
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                if c["conv_resample"]:
                    h = F.conv2d(h, sd[l[1] + ".conv.weight"], sd[l[1] + ".conv.bias"], padding=1)
            else:
                hc = head_ch if head_ch != -1 else l[2] // c["num_heads"]
                h = _attention(h, sd, l[1], groups, hc)
            if taps is not None:
                taps[l[1]] = h
        if b["group"] == "input":
            hs.append(h)
    h = F.silu(_group_norm(h, sd, "out.0", groups))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


class UnetOracle:
    """Callable with the reference backbone's surface (forward(x, times, classes), image_size, out_channels)."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor]):
        self.cfg = dict(cfg)
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.image_size = cfg["image_size"]
        self.out_channels = cfg["out_channels"]
        self.in_channels = cfg["in_channels"]
        self.num_classes = cfg.get("num_classes")

    def __call__(self, x, times, classes=None):
        return unet_forward(self.cfg, self.sd, x, times, classes)

    forward = __call__
