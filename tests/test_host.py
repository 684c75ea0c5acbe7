"""CPU: host logic of the mirrored classes — state-dict schema from the native topology builder, float64 tables,
reference-compatible constructor surface.  No GPU calls."""
import json

import numpy as np
import pytest
import torch

import ivid_b200.backbones as backbones
import ivid_b200.frameworks as frameworks
import ivid_b200.samplers as samplers


@pytest.mark.parametrize("name", ["rgbd_imagenet_adm_128_large_cfg", "rgbd_imagenet_adm_128_large_cond",
                                  "rgbd_singlecategory_adm_128_small", "rgbd_imagenet_adm_256_128_small_sr"])
def test_native_schema_matches_reference(golden, name):
    cfg = json.loads(bytes(golden[f"schemacfg_{name}"]).decode())
    want = [(k, tuple(s)) for k, s in json.loads(bytes(golden[f"schema_{name}"]).decode())]
    with torch.device("meta"):
        pass
    net = backbones.AdmUnet2d(**cfg)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == want


def test_zero_init_convention():
    cfg = dict(image_size=32, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
               attention_resolutions=[16], channel_mult=[1, 2], num_head_channels=64)
    net = backbones.AdmUnet2d(**cfg)
    sd = net.state_dict()
    for k, v in sd.items():
        if k.endswith(("out_layers.3.weight", "proj_out.weight")) or k.startswith("out.2."):
            assert float(v.abs().max()) == 0.0, k          # zero_module (adm.py:182,278,486)
    assert net.image_size == 32 and net.out_channels == 4 and net.num_classes is None and net.has_null_class is False


def test_framework_and_sampler_tables(golden):
    cfg = dict(image_size=32, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
               attention_resolutions=[16], channel_mult=[1, 2], num_head_channels=64)
    net = backbones.AdmUnet2d(**cfg)
    fw = frameworks.ClassifierFreeGuidance(net, timesteps=1000, beta_schedule="linear", p_uncond=0.1)
    assert isinstance(fw, frameworks.GaussianDiffusion) and fw.timesteps == 1000
    assert np.array_equal(fw.betas, golden["sched_betas"])
    assert set(fw.backbone_args.keys()) >= {"x", "times", "classes"}
    ddpm = samplers.DdpmSampler(fw)
    ddim = samplers.DdimSampler(fw)
    names = ["alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
             "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
    for i, n in enumerate(names):
        assert np.array_equal(getattr(ddpm, n), golden["sched_" + n]), n
        nat = ddpm.native_table(i)                       # float64 tables computed inside the C++ sampler
        assert np.allclose(nat, golden["sched_" + n], rtol=1e-14, atol=0), n
    assert np.array_equal(ddim.alphas_cumprod, golden["sched_alphas_cumprod"])
    with pytest.raises(NotImplementedError):
        frameworks.GaussianDiffusion(net, timesteps=10, beta_schedule="bogus")


def test_cosine_schedule_tables():
    """beta_schedule="cosine" (frameworks/utils.py:31-35, betas_for_alpha_bar :40-60) against the unmodified reference's betas
    (tests/golden/schedule_golden.npz); oracle restatement, host mirror and the C++ sampler's float64 tables."""
    import os
    from oracle import sampler_ref
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schedule_golden.npz"))
    cfg = dict(image_size=32, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
               attention_resolutions=[16], channel_mult=[1, 2], num_head_channels=64)
    net = backbones.AdmUnet2d(**cfg)
    for name, T in (("cosine", 1000), ("cosine", 50), ("linear", 50)):
        want = g[f"{name}_{T}"]
        assert np.array_equal(sampler_ref.get_betas(name, T), want)
        fw = frameworks.GaussianDiffusion(net, timesteps=T, beta_schedule=name)
        assert np.array_equal(fw.betas, want)
        tb = sampler_ref.Tables(want)
        ddpm = samplers.DdpmSampler(fw)
        names = ["alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
        for i, n in enumerate(names):
            assert np.array_equal(getattr(ddpm, n), getattr(tb, n)), (name, T, n)
            # alphas_cumprod reaches 1e-9 at the end of the cosine schedule: products in a different order differ in the last bits
            assert np.allclose(ddpm.native_table(i), getattr(tb, n), rtol=1e-12, atol=0), (name, T, n)


def test_no_cpu_fallback():
    cfg = dict(image_size=32, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
               attention_resolutions=[16], channel_mult=[1, 2], num_head_channels=64)
    net = backbones.AdmUnet2d(**cfg)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 4, 32, 32), torch.zeros(1, dtype=torch.long))
