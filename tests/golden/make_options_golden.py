"""Golden eps of the UNMODIFIED reference UNet for backbone options no shipped config sets (adm.py:214-221 use_scale_shift_norm=False;
adm.py:409,475 resblock_updown=False with Downsample2d / Upsample2d, with and without conv_resample),
on the tiny test architecture with the oracle's synthetic weights; pins the oracle's branch for them.

    python tests/golden/make_options_golden.py        # needs /root/reference; writes tests/golden/options_golden.npz
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402  (easydict shim + reference imports; does not regenerate anything on import)
from oracle import unet_ref       # noqa: E402

out = {}
for tag, extra in (("noshift", dict(use_scale_shift_norm=False)), ("plainconv", dict(resblock_updown=False, conv_resample=True)),
                   ("plainpool", dict(resblock_updown=False, conv_resample=False, use_scale_shift_norm=False))):
    cfg = dict(mg.TINY, **extra)
    sd = unet_ref.make_synthetic_state_dict(cfg, seed=77)
    net = mg.ref_model(cfg, sd)
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((2, 4, 32, 32)).astype(np.float32))
    t = torch.tensor([700, 3]); c = torch.tensor([4, -1])
    with torch.no_grad():
        ref = net(x, t, c)
    ora = unet_ref.unet_forward(cfg, sd, x, t, c)
    assert torch.equal(ref, ora), f"{tag}: oracle differs from the reference by {float((ref - ora).abs().max())}"
    out[f"{tag}_cfg"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    out[f"{tag}_x"] = x.numpy(); out[f"{tag}_t"] = t.numpy(); out[f"{tag}_c"] = c.numpy(); out[f"{tag}_eps"] = ref.numpy()
np.savez_compressed(os.path.join(HERE, "options_golden.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
