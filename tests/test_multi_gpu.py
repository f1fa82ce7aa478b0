"""Row-tiled single frame over several GPUs (NCCL halo exchange + flare all-reduce + gather) must
reproduce the one-GPU frame.  Skipped on boxes with a single GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from util import synthetic_disk_texture, synthetic_skybox
    from black_hole_renderer_b200 import Renderer
    from black_hole_renderer_b200.dist import render_tiled
    W, H = 640, 360
    sky, tex = synthetic_skybox(256, 512), synthetic_disk_texture(144, 976)
    r = Renderer(W, H, sky, tex, anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True, cuda_device=rank)
    from black_hole_renderer_b200.dist import attach_shared_frame, balance_tiles
    frame = render_tiled(r, [6, 0, 0.5], 90, rank=rank, world_size=world)
    if rank == 0:
        np.save(os.path.join(out_dir, "tiled.npy"), frame)
        single = r.render([6, 0, 0.5], 90)
        np.save(os.path.join(out_dir, "single.npy"), single)
        np.save(os.path.join(out_dir, "single_u8.npy"), r.render_u8([6, 0, 0.5], 90))
    dist.barrier()
    # cost-balanced tile heights (RK4 evaluations per row, all-gathered) + u8 frames
    bounds = balance_tiles(r, [6, 0, 0.5], 90, rank, world)
    frame = render_tiled(r, [6, 0, 0.5], 90, rank=rank, world_size=world, want_u8=True, bounds=bounds)
    if rank == 0:
        np.save(os.path.join(out_dir, "tiled_balanced_u8.npy"), frame)
        np.save(os.path.join(out_dir, "bounds.npy"), np.array(bounds))
    # ... and with every rank's own D2H into the shared host frame instead of the gather
    attach_shared_frame(r, rank, world)
    frame = render_tiled(r, [6, 0, 0.5], 90, rank=rank, world_size=world, want_u8=True, bounds=bounds)
    if rank == 0:
        np.save(os.path.join(out_dir, "tiled_shared_u8.npy"), frame)
    dist.barrier()
    r.close()
    dist.destroy_process_group()


def _peer_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from util import synthetic_disk_texture, synthetic_skybox
    from black_hole_renderer_b200 import Renderer
    from black_hole_renderer_b200.dist import attach_peers, render_tiled_peer
    from black_hole_renderer_b200.driver import orbit_camera
    W, H = 640, 360
    sky, tex = synthetic_skybox(256, 512), synthetic_disk_texture(144, 976)
    r = Renderer(W, H, sky, tex, anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True, cuda_device=rank)
    attach_peers(r, rank, world)
    cams = [orbit_camera([6, 0, 0.5], f, 12, 360.0) for f in range(4)]
    # several frames back to back: exercises the frame serials and the back-pressure flags
    frames = [render_tiled_peer(r, c, 90) for c in cams]
    if rank == 0:
        np.save(os.path.join(out_dir, "peer.npy"), np.stack([f.copy() for f in [frames[-1]]]))
        got = []
        for c in cams:
            got.append(render_tiled_peer(r, c, 90).copy())
        np.save(os.path.join(out_dir, "peer_all.npy"), np.stack(got))
    else:
        for c in cams:
            render_tiled_peer(r, c, 90)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "single_all.npy"), np.stack([r.render_u8(c, 90) for c in cams]))
    dist.barrier()
    # distributed egress: every rank copies its rows into one shared, page-locked host frame
    from black_hole_renderer_b200.dist import attach_shared_frame
    attach_shared_frame(r, rank, world)
    shared = []
    for c in cams + cams[::-1]:
        f = render_tiled_peer(r, c, 90)
        if rank == 0:
            shared.append(f.copy())
    if rank == 0:
        np.save(os.path.join(out_dir, "shared_all.npy"), np.stack(shared))
    dist.barrier()
    # cost-balanced tile heights installed on every rank (bhr_peer_set_tiles), frames back to back
    from black_hole_renderer_b200.dist import balance_tiles
    bounds = balance_tiles(r, cams[0], 90, rank, world)
    bal = []
    for c in cams:
        f = render_tiled_peer(r, c, 90)
        if rank == 0:
            bal.append(f.copy())
    if rank == 0:
        np.save(os.path.join(out_dir, "balanced_all.npy"), np.stack(bal))
        np.save(os.path.join(out_dir, "bounds.npy"), np.array(bounds))
    dist.barrier()
    # pipelined frames: egress on the copy streams, two shared host frames, the caller collects frame s - 1 after
    # it has enqueued frame s (two passes over the cameras: the frame serials wrap around the four completion events)
    from black_hole_renderer_b200.dist import render_tiled_peer_async, wait_tiled_frame
    attach_shared_frame(r, rank, world)
    piped = []
    seq = cams + cams[::-1] + cams
    for c in seq:
        render_tiled_peer_async(r, c, 90)
        f = wait_tiled_frame(r, back=1)
        if rank == 0 and f is not None:
            piped.append(f.copy())
    f = wait_tiled_frame(r, back=0)
    if rank == 0:
        piped.append(f.copy())
        assert len(piped) == len(seq)
        np.save(os.path.join(out_dir, "pipelined_all.npy"), np.stack(piped))
    dist.barrier()
    # ... and the synchronous call still works afterwards
    f = render_tiled_peer(r, cams[1], 90)
    if rank == 0:
        np.save(os.path.join(out_dir, "after_pipelined.npy"), f.copy())
    dist.barrier()
    r.close()
    dist.destroy_process_group()


def _failure_worker(rank, world, port, out_dir):
    """A rank that fails or goes missing must not hang its peers (csrc/peer.cu: bounded waits, poison)."""
    import time
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from util import synthetic_disk_texture, synthetic_skybox
    from black_hole_renderer_b200 import Renderer
    from black_hole_renderer_b200._lib import BhrError
    from black_hole_renderer_b200.dist import attach_peers, render_tiled_peer
    sky, tex = synthetic_skybox(256, 512), synthetic_disk_texture(144, 976)
    log = []
    for case in ("poison", "missing"):
        r = Renderer(320, 180, sky, tex, cuda_device=rank)
        r.set_option("peer_timeout_ms", 1500)
        attach_peers(r, rank, world)
        render_tiled_peer(r, [6, 0, 0.5], 90)                   # one good frame
        dist.barrier()
        t0 = time.perf_counter()
        try:
            if case == "poison":
                # rank 1 is handed a camera the library rejects (escape radius >= 1e6) after its first
                # wait is enqueued: it must poison the frame so that rank 0 drains at once
                render_tiled_peer(r, [6e5, 0, 0] if rank == 1 else [6, 0, 0.5], 90)
            elif rank == 0:
                render_tiled_peer(r, [6, 0, 0.5], 90)           # rank 1 never calls: the wait must time out
            log.append(f"{case}:ok:{time.perf_counter() - t0:.2f}")
        except BhrError as e:
            log.append(f"{case}:error:{time.perf_counter() - t0:.2f}:{str(e)[:60]}")
        dist.barrier()
        r.close()
    with open(os.path.join(out_dir, f"fail_{rank}.txt"), "w") as f:
        f.write("\n".join(log))
    dist.destroy_process_group()


def test_peer_failure_does_not_hang(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_failure_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = open(tmp_path / "fail_0.txt").read().splitlines()
    r1 = open(tmp_path / "fail_1.txt").read().splitlines()
    # poisoned frame: rank 1 reports its own error, rank 0 a state error, well before the 1.5 s budget
    assert r1[0].startswith("poison:error") and r0[0].startswith("poison:error"), (r0, r1)
    assert float(r0[0].split(":")[2]) < 1.0, r0
    # missing rank: rank 0 gives up after the budget instead of spinning for ever
    assert r0[1].startswith("missing:error") and 1.0 < float(r0[1].split(":")[2]) < 10.0, r0


@pytest.mark.parametrize("world", [2, 4])
def test_peer_memory_tiled_frame_equals_single_gpu(tmp_path, world):
    """csrc/peer.cu: halo rows loaded from the neighbours' HBM, rows stored into rank 0's buffers,
    flare sums exchanged by kernels -- the frames must equal the one-GPU frames."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    peer, single = np.load(tmp_path / "peer_all.npy"), np.load(tmp_path / "single_all.npy")
    d = np.abs(peer.astype(int) - single.astype(int))
    # the flare centroid is summed per tile (f64): last-bit differences of a few pixels at most
    assert d.max() <= 1 and (d.max(axis=-1) > 0).mean() < 1e-3
    assert np.array_equal(np.load(tmp_path / "peer.npy")[0], peer[-1])
    shared = np.load(tmp_path / "shared_all.npy")
    assert np.array_equal(shared[:4], peer) and np.array_equal(shared[4:], peer[::-1])
    bal, bounds = np.load(tmp_path / "balanced_all.npy"), np.load(tmp_path / "bounds.npy")
    assert bounds[0] == 0 and bounds[-1] == 360 and len(bounds) == world + 1 and (np.diff(bounds) >= 8).all()
    d = np.abs(bal.astype(int) - single.astype(int))
    assert d.max() <= 1 and (d.max(axis=-1) > 0).mean() < 1e-3
    # pipelined frames (egress on the copy streams, two host frames) are the same frames, in order
    piped = np.load(tmp_path / "pipelined_all.npy")
    assert np.array_equal(piped, np.concatenate([bal, bal[::-1], bal]))
    assert np.array_equal(np.load(tmp_path / "after_pipelined.npy"), bal[1])


@pytest.mark.parametrize("world", [2, 4])
def test_tiled_frame_equals_single_gpu(tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    tiled, single = np.load(tmp_path / "tiled.npy"), np.load(tmp_path / "single.npy")
    # the flare centroid is summed per tile and all-reduced in f64: last-bit differences only
    assert np.abs(tiled - single).max() <= 2e-6
    assert (tiled == single).mean() > 0.999
    u8 = np.load(tmp_path / "single_u8.npy").astype(int)
    for name in ("tiled_balanced_u8.npy", "tiled_shared_u8.npy"):
        d = np.abs(np.load(tmp_path / name).astype(int) - u8)
        assert d.max() <= 1 and (d.max(axis=-1) > 0).mean() < 1e-3, name
