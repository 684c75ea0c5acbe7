"""CPU: the oracle restatement (oracle/*.py) against the committed golden vectors that tests/golden/make_golden.py
produced by running the unmodified reference (so the oracle stays pinned where /root/reference does not exist)."""
import json

import numpy as np
import pytest
import torch

from oracle import sampler_ref, unet_ref


def _cfg(golden, tag):
    return json.loads(bytes(golden[f"{tag}_cfg"]).decode())


@pytest.mark.parametrize("tag", ["tiny", "tiny_cond", "tiny_sr"])
def test_unet_oracle_matches_reference_golden(golden, tag):
    cfg = _cfg(golden, tag)
    sd = unet_ref.make_synthetic_state_dict(cfg, seed=1234)
    x = torch.from_numpy(golden[f"{tag}_x"]); t = torch.from_numpy(golden[f"{tag}_t"]); c = torch.from_numpy(golden[f"{tag}_classes"])
    y = unet_ref.unet_forward(cfg, sd, x, t, c)
    ref = torch.from_numpy(golden[f"{tag}_eps"])
    assert float((y - ref).norm() / ref.norm()) < 1e-5          # same machine: 0; other CPUs: oneDNN summation order
    y0 = unet_ref.unet_forward(cfg, sd, x, t, None)
    ref0 = torch.from_numpy(golden[f"{tag}_eps_none"])
    assert float((y0 - ref0).norm() / ref0.norm()) < 1e-5


@pytest.mark.parametrize("name", ["rgbd_imagenet_adm_128_large_cfg", "rgbd_imagenet_adm_128_large_cond",
                                  "rgbd_singlecategory_adm_128_small", "rgbd_imagenet_adm_256_128_small_sr"])
def test_state_dict_schema(golden, name):
    cfg = json.loads(bytes(golden[f"schemacfg_{name}"]).decode())
    want = [(k, tuple(s)) for k, s in json.loads(bytes(golden[f"schema_{name}"]).decode())]
    assert list(unet_ref.unet_param_shapes(cfg).items()) == want
    assert len(want) in (493, 494)


def test_schedule_known_answers(golden):
    """SURVEY.md Appendix C values + full-table equality with the reference constructors."""
    tb = sampler_ref.Tables(sampler_ref.get_betas("linear", 1000))
    assert tb.betas[0] == 1e-4 and tb.betas[999] == 0.02
    assert abs(tb.alphas_cumprod[499] - 0.07858724288177824) < 1e-15
    assert abs(tb.sqrt_recip_alphas_cumprod[999] - 157.41045725150062) < 1e-9
    assert abs(tb.posterior_log_variance_clipped[0] - (-9.81672513529567)) < 1e-12
    for name in ["alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]:
        assert np.array_equal(getattr(tb, name), golden["sched_" + name]), name
    with pytest.raises(NotImplementedError):
        sampler_ref.get_betas("nope", 10)


def test_sampler_oracle_steps(golden):
    cfg = _cfg(golden, "tiny")
    sd = unet_ref.make_synthetic_state_dict(cfg, seed=1234)
    tb = sampler_ref.Tables(sampler_ref.get_betas("linear", 1000))
    x_t = torch.from_numpy(golden["step_x_t"]); classes = torch.from_numpy(golden["step_classes"])
    model = lambda x, t, c: unet_ref.unet_forward(cfg, sd, x, t, c)
    for ti in [999, 1, 0]:
        t = torch.tensor([ti] * x_t.shape[0])
        eps = sampler_ref.cfg_eps(model, x_t, t, classes, 0.5)
        xp, x0 = sampler_ref.ddpm_step(tb, x_t, t, eps, torch.from_numpy(golden[f"ddpm_t{ti}_noise"]))
        ref = torch.from_numpy(golden[f"ddpm_t{ti}_xprev"])
        assert float((xp - ref).norm() / ref.norm()) < 1e-5
    assert torch.equal(sampler_ref.make_sr_inputs(torch.from_numpy(golden["sr_x"]), torch.from_numpy(golden["sr_y"])),
                       torch.from_numpy(golden["sr_cond_inputs"]))


@pytest.mark.parametrize("tag", ["noshift", "plainconv", "plainpool"])
def test_unet_oracle_backbone_options(tag):
    """Backbone options no shipped config sets, against the unmodified reference's eps (tests/golden/options_golden.npz, made by
    tests/golden/make_options_golden.py): use_scale_shift_norm=False (adm.py:176, 219-221: h = out_layers(h + emb_out)),
    resblock_updown=False with Downsample2d / Upsample2d (adm.py:60-117, 409, 475) with and without conv_resample."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options_golden.npz"))
    cfg = json.loads(bytes(g[f"{tag}_cfg"]).decode())
    sd = unet_ref.make_synthetic_state_dict(cfg, seed=77)
    if not cfg.get("use_scale_shift_norm", True):
        assert sd["input_blocks.1.0.emb_layers.1.weight"].shape[0] == sd["input_blocks.1.0.out_layers.0.weight"].shape[0]
    if tag == "plainconv":
        assert "input_blocks.2.0.op.weight" in sd and "output_blocks.1.2.conv.weight" in sd
    if tag == "plainpool":
        assert not any(".op." in k or k.endswith(".conv.weight") for k in sd)
    y = unet_ref.unet_forward(cfg, sd, torch.from_numpy(g[f"{tag}_x"]), torch.from_numpy(g[f"{tag}_t"]), torch.from_numpy(g[f"{tag}_c"]))
    ref = torch.from_numpy(g[f"{tag}_eps"])
    assert float((y - ref).norm() / ref.norm()) < 1e-5
