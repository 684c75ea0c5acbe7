"""CPU: host logic of the sampling driver — view sets, per-rank sharding (incl. a world_size-2 gloo run), scene format."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ivid_b200.inference import (build_modelviews, colorize_depth, load_scene_views, parse_int_list, random_views, reorder,
                                 save_scene, shard, swing_trajectory)
from ivid_b200.utils import edict


def test_parse_int_list_and_reorder():
    assert parse_int_list("0-3,7,9-10") == [0, 1, 2, 3, 7, 9, 10]
    data = [torch.full((1,), float(i)) for i in range(27)]
    out = reorder(data, "3x9")
    assert out.shape == (27, 1) and out[13].item() == 0.0 and out[0].item() == 23.0     # view 0 sits in the grid centre
    assert reorder(data[1:], "3x9")[13].item() == -1.0                                  # 26 cond views: blank first cell
    with pytest.raises(NotImplementedError):
        reorder(data, "bogus")


def test_viewsets():
    v = build_modelviews("3x9", 1)
    assert len(v) == 27 and np.allclose(v[0], build_modelviews("uncond", 1)[0])
    eye = np.linalg.inv(v[1].astype(np.float64))[:3, 3]
    assert np.allclose(eye, [0, np.sin(0.15), np.cos(0.15)], atol=1e-6)                 # yaw 0, pitch +0.15 (sample.py:324-336)
    r = build_modelviews("random", 3, rng=np.random.default_rng(0))
    assert len(r) == 3 and len(r[0]) == 2
    for m in v:
        assert np.allclose(np.linalg.norm(np.linalg.inv(m.astype(np.float64))[:3, 3]), 1.0, atol=1e-6)   # eyes on the unit sphere
    with pytest.raises(NotImplementedError):
        build_modelviews("nope", 1)


def test_shard_partition():
    items = list(range(23))
    for w in (1, 2, 4, 8):
        parts = [shard(items, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == items and max(map(len, parts)) - min(map(len, parts)) <= 1
    assert shard(None, 0, 2) is None


def test_scene_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    meshes = [edict(depth=rng.uniform(0.6, 5, (16, 16, 1)).astype(np.float32), fov=45, modelview=build_modelviews("uncond", 1)[0]) for _ in range(2)]
    colors = [rng.uniform(0, 1, (16, 16, 3)).astype(np.float32) for _ in range(2)]
    p = os.path.join(tmp_path, "scene.npz")
    save_scene(p, meshes, colors)
    back = load_scene_views(p)
    for m, c, b in zip(meshes, colors, back):
        assert np.array_equal(b.depth, m.depth)                                          # float32 bits survive the RGBA8 PNG
        assert np.array_equal((np.clip(c * 255, 0, 255)).astype(np.uint8), np.round(b.color * 255).astype(np.uint8))
        assert b.fov == 45


def test_free_view_trajectories():
    """inference/render.py:43-61: cameras on the unit sphere looking at the origin."""
    tr = swing_trajectory(9)
    assert len(tr) == 9
    for mv in tr:
        m = np.asarray(mv, dtype=np.float64)
        R, t = m[:3, :3], m[:3, 3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
        eye = -R.T @ t
        assert abs(np.linalg.norm(eye) - 1) < 1e-6                     # unit sphere
        assert np.allclose(R @ (0 - eye), [0, 0, -1], atol=1e-6)        # looks at the origin down -z
    assert np.allclose(tr[0], tr[-1], atol=1e-6)                        # t = 0 and t = 2 pi coincide
    e0 = -np.asarray(tr[0])[:3, :3].T @ np.asarray(tr[0])[:3, 3]
    assert np.allclose(e0, [np.sin(0.6), 0, np.cos(0.6)], atol=1e-6)    # yaw 0.6, pitch 0 at t = 0
    a, b = random_views(4, seed=3), random_views(4, seed=3)
    assert len(a) == 4 and all(len(v) == 1 for v in a)
    assert all(np.array_equal(x[0], y[0]) for x, y in zip(a, b))
    for v in a:
        m = np.asarray(v[0], dtype=np.float64)
        eye = -m[:3, :3].T @ m[:3, 3]
        yaw, pitch = np.arctan2(eye[0], eye[2]), np.arcsin(eye[1])
        assert abs(yaw) <= 0.6 + 1e-6 and abs(pitch) <= 0.15 + 1e-6


def test_colorize_depth():
    import cv2
    d = np.linspace(0, 1, 32 * 32, dtype=np.float32).reshape(32, 32)
    c = colorize_depth(d, min=0, max=1)
    assert c.shape == (32, 32, 3) and c.min() >= 0 and c.max() <= 1
    ref = cv2.cvtColor(cv2.applyColorMap((np.clip(1 - d, 0, 1) * 255).astype(np.uint8), cv2.COLORMAP_INFERNO), cv2.COLOR_BGR2RGB) / 255
    assert np.array_equal(c, ref)
    assert c[0, 0].sum() > c[-1, -1].sum()                 # near is bright
    t = colorize_depth(torch.from_numpy(d)[None] * 2 - 1)   # torch in [-1,1] -> torch CHW in [-1,1]
    assert tuple(t.shape) == (3, 32, 32) and float(t.min()) >= -1 and float(t.max()) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = list(range(11))
    mine = shard(seeds, rank, world)
    # weights: rank 0 holds the packed arena, everyone else receives it with ONE broadcast (bench.py / sample driver)
    arena = torch.arange(1024, dtype=torch.uint8) if rank == 0 else torch.zeros(1024, dtype=torch.uint8)
    dist.broadcast(arena, src=0)
    got = [None] * world
    dist.all_gather_object(got, mine)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                                             # max-over-ranks timing reduction
    q.put((rank, mine, bool(torch.equal(arena, torch.arange(1024, dtype=torch.uint8))), got, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mine, ok, got, tmax in res:
        assert ok and tmax == 2.0
        assert sorted(got[0] + got[1]) == list(range(11)) and not set(got[0]) & set(got[1])
        assert mine == list(range(11))[rank::2]


def test_image_grid_matches_torchvision():
    """image_grid_u8 == torchvision.utils.save_image(make_grid(..., normalize=True, value_range=(-1, 1))) (sample.py:160-166)."""
    tv = pytest.importorskip("torchvision.utils")
    from ivid_b200.inference import image_grid_u8
    g = torch.Generator().manual_seed(0)
    for K, nrow in ((2, 2), (27, 9), (5, 4), (1, 9)):
        img = torch.randn(K, 3, 12, 10, generator=g) * 0.9
        want = tv.make_grid(img, nrow=nrow, normalize=True, value_range=(-1, 1))
        want = want.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)
        got = image_grid_u8(img, nrow)
        assert got.shape == want.shape and torch.equal(got, want), (K, nrow)


def test_async_save_outputs(tmp_path):
    """The directory contract of the reference's async_save (sample.py:150-176) per view set."""
    from PIL import Image
    from ivid_b200.inference import async_save
    g = torch.Generator().manual_seed(1)
    mv = build_modelviews("uncond", 1)[0]

    def scene(V):
        meshes = [edict(depth=np.full((8, 8, 1), 1.5, np.float32), fov=45, modelview=mv) for _ in range(V)]
        return meshes, [np.full((8, 8, 3), 0.5, np.float32) for _ in range(V)]
    for vs, V in (("uncond", 1), ("random", 2), ("3x9", 27)):
        out = os.path.join(tmp_path, vs)
        for sub in ("results", "grids", "conds", "scenes"):
            os.makedirs(os.path.join(out, sub))
        samples = torch.rand(V, 4, 8, 8, generator=g) * 2 - 1
        conds = {"color": torch.rand(V - 1, 3, 8, 8, generator=g) * 2 - 1, "depth": torch.rand(V - 1, 1, 8, 8, generator=g) * 2 - 1} if V > 1 else None
        meshes, colors = scene(V)
        async_save(meshes, colors, samples, conds, "class001_seed00005", edict(output_dir=out, viewset=vs)).join()
        files = sorted(os.path.relpath(os.path.join(d, f), out) for d, _, fs in os.walk(out) for f in fs)
        want = {"uncond": ["results/rgb_class001_seed00005.png", "scenes/scene_class001_seed00005.npz"],
                "random": ["conds/rgb_class001_seed00005.png", "grids/rgb_class001_seed00005.png", "results/rgb_class001_seed00005.png"],
                "3x9": ["conds/depth_cond_class001_seed00005.png", "conds/rgb_cond_class001_seed00005.png", "grids/depth_class001_seed00005.png",
                        "grids/rgb_class001_seed00005.png", "scenes/scene_class001_seed00005.npz"]}[vs]
        assert files == sorted(want), (vs, files)
        if vs == "uncond":
            px = np.array(Image.open(os.path.join(out, want[0])))
            ref = (np.clip(samples[0, :3].numpy().transpose(1, 2, 0) * 0.5 + 0.5, 0, 1) * 255).astype(np.uint8)
            assert np.array_equal(px, ref)
        if vs == "3x9":
            assert np.array(Image.open(os.path.join(out, "grids/rgb_class001_seed00005.png"))).shape == (3 * 10 + 2, 9 * 10 + 2, 3)
            assert len(load_scene_views(os.path.join(out, want[-1]))) == 27


def test_scene_modelview_accepts_column_major_glm_objects(tmp_path):
    """A PyGLM mat4 indexes as m[col][row]; save_scene must store the same camera either way."""
    from ivid_b200.inference.utils import _store_modelview
    m = build_modelviews("3x9", 1)[5]

    class FakeMat4:                       # quacks like glm.mat4: module name 'glm', column-major indexing
        __module__ = "glm"
        def __init__(self, a): self.a = np.asarray(a)
        def __getitem__(self, c): return [float(self.a[r][c]) for r in range(4)]
    stored = np.asarray(_store_modelview(FakeMat4(m)), dtype=np.float32)
    assert np.allclose(stored, np.asarray(_store_modelview(m), dtype=np.float32)) and np.allclose(stored, m)
