"""Multiview sampling driver — mirror of the reference's inference/sample.py (sample_all :30-147, per-rank sharding
:199-202, view sets :304-338) with the per-view loop kept on the device:

    view 0:  unconditional sampler  (DDPM 1000 / DDIM)                               -> RGBD on the GPU
    view j:  DeviceWarp.aggregate (CUDA mesh + rasterise + aggregate + post-filters)  -> condition maps on the GPU
             conditional DDIM sampler with replace / constrain guidance               -> RGBD on the GPU

Nothing crosses PCIe inside the loop; samples are copied to the host once per batch for saving.

    python -m ivid_b200.inference.sample --config_uncond ... --ckpt_uncond ... (same flags as the reference CLI)
"""
from __future__ import annotations

import argparse
import json
import os
import threading

import numpy as np
import torch

from .. import backbones, frameworks, samplers
from ..rgbd_3d import DeviceWarp, glm_compat as glm
from ..rgbd_3d import utils as rgbd_utils
from ..utils import edict
from .utils import colorize_depth, parse_int_list, reorder, save_scene


def shard(items, rank, world_size):
    """seeds / classes / per-sample modelviews of this rank (sample.py:199-202: [rank::world_size])."""
    return items[rank::world_size] if items is not None else None


def build_modelviews(viewset, num_samples, rng=None):
    """View sets of sample.py:304-338.  'random' draws yaw ~ N(0, 0.3^2), pitch ~ N(0, 0.15^2) per sample; the reference
    uses the unseeded global numpy RNG (sample.py:317-318) — pass `rng` for reproducible runs."""
    origin = lambda: glm.lookAt(glm.vec3(0, 0, 1), glm.vec3(0, 0, 0), glm.vec3(0, 1, 0))
    on_sphere = lambda yaw, pitch: glm.lookAt(glm.vec3(np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch)),
                                              glm.vec3(0, 0, 0), glm.vec3(0, 1, 0))
    if viewset == "uncond":
        return [origin()]
    if viewset == "random":
        normal = (rng.normal if rng is not None else np.random.normal)
        mvs = []
        for _ in range(num_samples):
            yaw = 0.3 * normal()
            pitch = 0.15 * normal()
            mvs.append([origin(), on_sphere(yaw, pitch)])
        return mvs
    if viewset == "3x9":
        yaws, pitches = [0.0], [0.0]
        for i in range(4):
            yaws += [(i + 1) * 0.15, -(i + 1) * 0.15]
        for i in range(1):
            pitches += [(i + 1) * 0.15, -(i + 1) * 0.15]
        return [on_sphere(y, p) for y in yaws for p in pitches]
    raise NotImplementedError


@torch.no_grad()
def sample_all(framework_uncond, framework_cond, seeds_or_num_samples, steps_uncond, steps_cond, modelviews, fov=45, near=0.6,
               far=5, atol=0.03, rtol=0.03, erode_rgb=2, classes=None, guidance=3.0, batchsize=10, rng="philox"):
    """Generator over finished samples: (meshes, colors, samples [V,4,H,W], conds) — signature of sample.py:30-46.
    `meshes[v]` carries what save_scene needs (linear depth, fov, modelview)."""
    sampler_uncond = samplers.DdimSampler(framework_uncond) if steps_uncond < 1000 else samplers.DdpmSampler(framework_uncond)
    sampler_cond = samplers.DdimSampler(framework_cond) if framework_cond is not None else None
    num_samples = seeds_or_num_samples if not isinstance(seeds_or_num_samples, list) else len(seeds_or_num_samples)
    seeds = seeds_or_num_samples if isinstance(seeds_or_num_samples, list) else None
    net = framework_uncond.backbone
    S = net.image_size
    dev = net.device
    per_sample_views = isinstance(modelviews[0], list)
    wparams = dict(fov=fov, near=near, far=far, atol=atol, rtol=rtol, erode_rgb=erode_rgb)
    warps = {}

    for i in range(0, num_samples, batchsize):
        bs = min(batchsize, num_samples - i)
        if seeds is not None:
            noise = []
            for j in range(bs):
                torch.manual_seed(seeds[i + j])
                noise.append(torch.randn(1, 4, S, S, device=dev))       # sample.py:66-69
            noise = torch.cat(noise, dim=0)
        else:
            noise = None
        b_classes = torch.tensor(classes[i: i + bs]).long().to(dev) if classes is not None else None
        views_of = (lambda k: modelviews[i + k]) if per_sample_views else (lambda k: modelviews)
        n_views = len(views_of(0))
        warp = None
        if framework_cond is not None and n_views > 1:
            if bs not in warps:
                warps[bs] = DeviceWarp(bs, image_size=S, ssaa=3, max_views=max(n_views, 2), device=dev.index)
            warp = warps[bs]
            warp.reset()
        samples, cond_color, cond_depth = [], [], []
        cfg_u = isinstance(framework_uncond, frameworks.ClassifierFreeGuidance)
        for j in range(n_views):
            mv_j = [views_of(k)[j] for k in range(bs)] if per_sample_views else views_of(0)[j]
            if j == 0:
                kw = dict(strength=guidance) if cfg_u else {}
                res = sampler_uncond.sample(bs, noise=noise, classes=b_classes, steps=steps_uncond, verbose=False, rng=rng, **kw)
            else:
                cond = warp.aggregate(mv_j, **wparams)                   # [bs,7,S,S] in [0,1]
                y = cond[:, 0:4] * 2 - 1                                 # sample.py:103
                mask, mask_rgb = cond[:, 4:5], cond[:, 5:6]
                cond_color.append(cond[:, 0:3] * 2 - 1)
                cond_depth.append(cond[:, 3:4] * 2 - 1)
                args = dict(y=y, mask=mask, mask_rgb=mask_rgb, replace_rgb=(0.1, y[:, :3], mask_rgb),
                            replace_depth=(0.2, y[:, 3:], mask), constrain_depth=(0.5, cond[:, 6:7] * 2 - 1))   # sample.py:104-119
                kw = dict(strength=guidance) if cfg_u else {}
                res = sampler_cond.sample(bs, classes=b_classes, steps=steps_cond, verbose=False, rng=rng, **args, **kw)
            samples.append(res.samples)
            if warp is not None:
                warp.add_view(res.samples, mv_j, **wparams)
        samples = torch.stack(samples, dim=1)                           # [bs, V, 4, S, S]
        conds = {"color": torch.stack(cond_color, dim=1), "depth": torch.stack(cond_depth, dim=1)} if cond_color else None
        rgbd = samples.permute(0, 1, 3, 4, 2).cpu().numpy() * 0.5 + 0.5   # one D2H per batch
        for k in range(bs):
            meshes = [edict(depth=rgbd_utils.linearize_depth(rgbd[k, v, :, :, 3:], near, far), fov=fov,
                            modelview=(views_of(k)[v])) for v in range(n_views)]
            colors = [rgbd[k, v, :, :, :3] for v in range(n_views)]
            yield meshes, colors, samples[k], ({n: t[k] for n, t in conds.items()} if conds is not None else None)


def image_grid_u8(images, nrow, value_range=(-1, 1), padding=2):
    """uint8 [H', W', 3] grid of `images` [K,3,H,W] — the arithmetic of torchvision.utils.save_image(make_grid(images, nrow,
    normalize=True, value_range=value_range)) that the reference calls (sample.py:160-166): clamp to the range, scale to
    [0,1] with the 1e-5 guard, tiles separated by `padding` black pixels, then *255 + 0.5 and truncation.  Runs on the
    tensor's device (the GPU in the sampling loop), so only the packed uint8 image crosses PCIe."""
    lo, hi = value_range
    t = images.detach().to(torch.float32).clamp(lo, hi).sub(lo).div(max(hi - lo, 1e-5))
    K, C, H, W = t.shape
    if K == 1:                                  # make_grid returns a single image unpadded
        grid = t[0]
    else:
        xm = min(nrow, K)
        ym = -(-K // xm)
        hh, ww = H + padding, W + padding
        grid = t.new_zeros((C, hh * ym + padding, ww * xm + padding))
        for k in range(K):
            y, x = divmod(k, xm)
            grid[:, y * hh + padding: y * hh + padding + H, x * ww + padding: x * ww + padding + W] = t[k]
    return grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)


def to_u8(rgb):
    """(np.clip(x*0.5+0.5, 0, 1) * 255).astype(uint8) of a [3,H,W] model-space image (sample.py:156,161-162)."""
    return (rgb.detach().to(torch.float32).mul(0.5).add(0.5).clamp(0, 1).mul(255)).to(torch.uint8).permute(1, 2, 0)


def async_save(meshes, colors, samples, conds, suffix, cfg):
    """Writes the outputs of the reference's async_save (sample.py:150-176) for one finished sample:
         viewset uncond: results/rgb_*.png + scenes/scene_*.npz
         viewset random: grids/rgb_*.png (both views), conds/rgb_*.png (view 0), results/rgb_*.png (view 1)
         viewset 3x9   : grids/{rgb,depth}_*.png, conds/{rgb_cond,depth_cond}_*.png (3x9 mosaics, `reorder`), scenes/scene_*.npz
    The 8-bit images are packed on the device on the caller's stream and copied to pinned host memory asynchronously; a
    worker thread waits for that copy, encodes the PNGs and writes the scene, while the main thread goes on sampling."""
    from PIL import Image
    out = cfg.output_dir
    jobs = []       # (relative path, device uint8 HWC tensor)
    vs = cfg.viewset
    if vs == "uncond":
        jobs.append((os.path.join("results", f"rgb_{suffix}.png"), to_u8(samples[0, :3])))
    elif vs == "random":
        jobs.append((os.path.join("grids", f"rgb_{suffix}.png"), image_grid_u8(samples[:, :3], 2)))
        jobs.append((os.path.join("conds", f"rgb_{suffix}.png"), to_u8(samples[0, :3])))
        jobs.append((os.path.join("results", f"rgb_{suffix}.png"), to_u8(samples[1, :3])))
    elif vs == "3x9":
        dev = samples.device
        jobs.append((os.path.join("grids", f"rgb_{suffix}.png"), image_grid_u8(reorder(samples[:, :3], vs), 9)))
        jobs.append((os.path.join("grids", f"depth_{suffix}.png"), image_grid_u8(reorder(colorize_depth(samples[:, 3:]).to(dev), vs), 9)))
        jobs.append((os.path.join("conds", f"rgb_cond_{suffix}.png"), image_grid_u8(reorder(conds["color"][:, :3], vs), 9)))
        jobs.append((os.path.join("conds", f"depth_cond_{suffix}.png"), image_grid_u8(reorder(colorize_depth(conds["depth"]).to(dev), vs), 9)))
    else:
        raise NotImplementedError
    host = []
    for rel, t in jobs:
        h = torch.empty(t.shape, dtype=torch.uint8, pin_memory=t.is_cuda)
        h.copy_(t, non_blocking=True)
        host.append((rel, h))
    done = torch.cuda.Event() if samples.is_cuda else None
    if done is not None:
        done.record()

    def worker():
        if done is not None:
            done.synchronize()
        for rel, h in host:
            Image.fromarray(h.numpy()).save(os.path.join(out, rel))
        if vs in ("uncond", "3x9"):
            save_scene(os.path.join(out, "scenes", f"scene_{suffix}.npz"), meshes, colors)

    th = threading.Thread(target=worker)
    th.start()
    return th


def _load_model(cfg, ckpt, device):
    net = getattr(backbones, cfg["backbone"]["name"])(**cfg["backbone"]["args"])
    if ckpt is not None:
        net.load_state_dict(torch.load(ckpt, map_location="cpu"))
    net = net.to(device)
    fw = getattr(frameworks, cfg["framework"]["name"])(net, **cfg["framework"]["args"])
    return fw


def main(rank, world_size, opt):
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    cfg_u = json.load(open(opt.config_uncond))
    fw_u = _load_model(cfg_u, opt.ckpt_uncond, dev)
    fw_c = _load_model(json.load(open(opt.config_cond)), opt.ckpt_cond, dev) if opt.viewset != "uncond" else None
    seeds = parse_int_list(opt.seeds) if opt.num_samples is None else None
    num = len(seeds) if seeds is not None else opt.num_samples
    ncls = cfg_u["backbone"]["args"].get("num_classes")
    classes = None
    if ncls is not None:
        if opt.classes == "mod":
            classes = [seeds[i] % ncls for i in range(num)]
        elif opt.classes == "uniform":
            classes = [i % ncls for i in range(num)]
        elif opt.classes == "random":
            classes = [int(np.random.randint(ncls)) for _ in range(num)]
        else:
            classes = parse_int_list(opt.classes)
    mvs = build_modelviews(opt.viewset, num)
    seeds_r, classes_r = shard(seeds, rank, world_size), shard(classes, rank, world_size)
    idx = list(range(num))[rank::world_size]
    mvs_r = shard(mvs, rank, world_size) if isinstance(mvs[0], list) else mvs
    out_dir = os.path.join(opt.output_dir, f"viewset_{opt.viewset}_steps_u{opt.steps_uncond}_c{opt.steps_cond}_guidance{opt.guidance}")
    for sub in ("results", "grids", "conds", "scenes"):                 # sample.py:283-286
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    save_cfg = edict(output_dir=out_dir, viewset=opt.viewset)
    gen = sample_all(fw_u, fw_c, seeds_r if seeds_r is not None else len(idx), opt.steps_uncond, opt.steps_cond, mvs_r, classes=classes_r,
                     guidance=opt.guidance, batchsize=opt.batchsize, fov=opt.fov, near=opt.near, far=opt.far, atol=opt.atol,
                     rtol=opt.rtol, erode_rgb=opt.erode_rgb, rng=opt.rng)
    threads = []
    for i, (meshes, colors, samples, conds) in enumerate(gen):
        tag = (f"class{classes_r[i]:03d}_" if classes_r is not None else "") + (f"seed{seeds_r[i]:05d}" if seeds_r is not None else f"{idx[i]:05d}")
        threads.append(async_save(meshes, colors, samples, conds, tag, save_cfg))
    for th in threads:
        th.join()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_uncond", default="configs/rgbd_imagenet_adm_128_large_cfg.json")
    ap.add_argument("--config_cond", default="configs/rgbd_imagenet_adm_128_large_cond.json")
    ap.add_argument("--ckpt_uncond", default=None)
    ap.add_argument("--ckpt_cond", default=None)
    ap.add_argument("--output_dir", default="samples/imagenet128")
    ap.add_argument("--seeds", default="0-8")
    ap.add_argument("--num_samples", type=int, default=None)
    ap.add_argument("--classes", default="mod")
    ap.add_argument("--viewset", default="3x9")
    ap.add_argument("--steps_uncond", type=int, default=1000)
    ap.add_argument("--steps_cond", type=int, default=50)
    ap.add_argument("--guidance", type=float, default=3.0)
    ap.add_argument("--batchsize", type=int, default=10)
    ap.add_argument("--fov", type=float, default=45)
    ap.add_argument("--near", type=float, default=0.6)
    ap.add_argument("--far", type=float, default=5)
    ap.add_argument("--atol", type=float, default=0.03)
    ap.add_argument("--rtol", type=float, default=0.03)
    ap.add_argument("--erode_rgb", type=int, default=3)
    ap.add_argument("--rng", choices=["philox", "torch"], default="philox",
                    help="per-step noise: 'philox' draws in-kernel (fast, default); 'torch' draws with the torch generator exactly "
                         "where the reference does (seed-for-seed reproduction of the reference's images needs this)")
    o = ap.parse_args()
    n = torch.cuda.device_count()
    if n <= 1:
        main(0, 1, o)
    else:
        import torch.multiprocessing as mp
        mp.spawn(main, args=(n, o), nprocs=n)
