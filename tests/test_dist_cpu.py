"""Host-side logic of the multi-GPU paths on CPU: row tiling + halo exchange + gather with the
gloo backend (world_size 2 and 3), frame sharding of the video driver."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _truth(C, H, W):
    y = torch.arange(H, dtype=torch.float32)[None, :, None]
    x = torch.arange(W, dtype=torch.float32)[None, None, :]
    c = torch.arange(C, dtype=torch.float32)[:, None, None]
    return 1000 * c + y + 0.001 * x


def _worker(rank, world, port, H, W, radius, result_dir, bounds=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from black_hole_renderer_b200.dist import equal_bounds, exchange_halos, gather_rows
    C = 3
    truth = _truth(C, H, W)
    b = bounds or equal_bounds(H, world)
    row0, row1 = b[rank], b[rank + 1]
    plane = torch.full((C, H, W), -1.0)
    plane[:, row0:row1] = truth[:, row0:row1]
    exchange_halos(plane, H, radius, rank, world, bounds=bounds)
    lo, hi = max(row0 - radius, 0), min(row1 + radius, H)
    ok = bool(torch.equal(plane[:, lo:hi], truth[:, lo:hi]))
    # rows outside tile + halo must be untouched
    untouched = bool((plane[:, :lo] == -1).all() and (plane[:, hi:] == -1).all())
    tile = truth[0, row0:row1].unsqueeze(-1).repeat(1, 1, 3).contiguous()
    want = truth[0].unsqueeze(-1).repeat(1, 1, 3)
    full = gather_rows(tile, H, rank, world, 0, bounds=bounds)
    gathered = bool(torch.equal(full, want)) if rank == 0 else full is None
    # receiving straight into the frame whose own rows the tile aliases (no concatenation)
    frame = torch.full((H, W, 3), -2.0)
    frame[row0:row1] = tile
    full = gather_rows(frame[row0:row1], H, rank, world, 0, bounds=bounds, out=frame if rank == 0 else None)
    gathered = gathered and (bool(torch.equal(full, want)) if rank == 0 else full is None)
    with open(os.path.join(result_dir, f"r{rank}"), "w") as f:
        f.write(f"{int(ok)}{int(untouched)}{int(gathered)}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,radius", [(2, 64, 7), (3, 50, 12), (3, 20, 9)])
def test_halo_exchange_and_gather_gloo(tmp_path, world, H, radius):
    """(3, 20, 9): tiles of 6-7 rows are shorter than the radius, so halos span two ranks."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, 16, radius, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}").read() == "111", r


@pytest.mark.parametrize("world,H,radius,bounds", [(3, 48, 10, [0, 30, 36, 48]), (2, 40, 6, [0, 9, 40])])
def test_halo_exchange_and_gather_with_cost_balanced_tiles(tmp_path, world, H, radius, bounds):
    """Uneven tile heights (cost-balanced split), one of them shorter than the halo radius."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, 16, radius, str(tmp_path), bounds), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}").read() == "111", r


def test_balanced_bounds_equalise_the_cost():
    from black_hole_renderer_b200.dist import balanced_bounds, equal_bounds
    rng = np.random.default_rng(0)
    H = 2160
    y = np.arange(H)
    cost = 3840 * (70 + 60 * np.exp(-((y - 1080) / 300.0) ** 2)) + rng.integers(0, 1000, H)
    for world in (1, 2, 4, 8):
        b = balanced_bounds(cost, world, min_rows=8)
        assert b[0] == 0 and b[-1] == H and len(b) == world + 1 and all(b[i + 1] - b[i] >= 8 for i in range(world))
        per = [cost[b[i]:b[i + 1]].sum() for i in range(world)]
        eq = equal_bounds(H, world)
        per_eq = [cost[eq[i]:eq[i + 1]].sum() for i in range(world)]
        assert max(per) <= max(per_eq) * 1.002
        assert max(per) / (sum(per) / world) < 1.02
    # degenerate profiles still give valid, ordered bounds
    b = balanced_bounds(np.zeros(64), 4, min_rows=8)
    assert b[0] == 0 and b[-1] == 64 and all(b[i + 1] - b[i] >= 8 for i in range(4))
    b = balanced_bounds(np.r_[np.zeros(60), 1e9, np.zeros(3)], 4, min_rows=8)
    assert b[0] == 0 and b[-1] == 64 and all(b[i + 1] - b[i] >= 8 for i in range(4))


def test_tile_rows_partition():
    from black_hole_renderer_b200.dist import owners_of_rows, tile_rows
    for H in (2160, 1080, 37, 8):
        for world in (1, 2, 3, 4, 8):
            rows = [tile_rows(H, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == H
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1
            cover = owners_of_rows(H, world, -5, H + 5)
            assert sum(b - a for _, a, b in cover) == H


def test_frame_sharding_blocks_of_60():
    from black_hole_renderer_b200.driver import STATS_PERIOD, frame_owner, orbit_camera
    n = 3600
    for world in (1, 2, 4, 8):
        owners = [frame_owner(f, world) for f in range(n)]
        assert set(owners) == set(range(world))
        # a rank owns whole stats blocks, so the frame that recomputes the statistics
        # (frame % 60 == 0, render.py:4457) is rendered by the rank that uses them
        for f in range(n):
            assert owners[f] == owners[f - f % STATS_PERIOD]
        counts = np.bincount(owners, minlength=world)
        assert counts.max() - counts.min() <= STATS_PERIOD
    # orbit camera: radius = |pov| (3-D norm), height = pov.z
    cam = orbit_camera([6, 0, 0.5], 900, 3600, 360.0)
    assert abs(np.hypot(cam[0], cam[1]) - np.sqrt(36.25)) < 1e-12 and cam[2] == 0.5
    assert abs(cam[0]) < 1e-9 and cam[1] > 0


def _video_worker(rank, world, port, out_dir):
    """One rank of a sharded `render.py --video` run with a fake renderer: the control plane (gloo
    barriers, per-rank progress files, rank 0 merging and muxing) is the real one (render.py:91-103)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["BHR_MUX"] = "png"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_video_driver import FakeRenderer
    from black_hole_renderer_b200.driver import render_video
    r = FakeRenderer()
    render_video(r, 8, 4, n_frames=150, fps=30, output_path=os.path.join(out_dir, "v.mp4"), fov=90.0,
                 static_cam_pos=[6, 0, 0.5], orbit=True, resume=False, disk_rotation_speed=0.1,
                 orbit_degrees=360.0, rank=rank, world_size=world, barrier=dist.barrier)
    np.save(os.path.join(out_dir, f"cams{rank}.npy"), np.array(r.cams))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_video_run_over_gloo_ranks(tmp_path):
    """Two processes share one 150-frame job in 60-frame blocks (rank 0: frames 0-59 and 120-149,
    rank 1: 60-119), no data-path collective; rank 0 merges the progress lists and writes the movie
    with every frame in order."""
    import json
    from black_hole_renderer_b200 import mov
    mp.spawn(_video_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    frames_dir = [d for d in os.listdir(tmp_path) if d.startswith(".frames_")][0]
    d = tmp_path / frames_dir
    assert sorted(json.load(open(d / "progress.json"))["completed"]) == list(range(150))
    assert len(np.load(tmp_path / "cams0.npy")) == 90 and len(np.load(tmp_path / "cams1.npy")) == 60
    # rank 1's first camera is the orbit position of frame 60
    a = np.radians(60 * 360.0 / 150)
    np.testing.assert_allclose(np.load(tmp_path / "cams1.npy")[0], [np.sqrt(36.25) * np.cos(a), np.sqrt(36.25) * np.sin(a), 0.5], atol=1e-12)
    w, h, fps, index = mov.read_png_movie_index(tmp_path / "v.mp4")
    assert (w, h, fps, len(index)) == (8, 4, 30.0, 150)
    blob = open(tmp_path / "v.mp4", "rb").read()
    for f in (0, 59, 60, 119, 120, 149):
        off, size = index[f]
        assert blob[off:off + size] == open(d / f"frame_{f:04d}.png", "rb").read()
