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
    frame = render_tiled(r, [6, 0, 0.5], 90, rank=rank, world_size=world)
    if rank == 0:
        np.save(os.path.join(out_dir, "tiled.npy"), frame)
        single = r.render([6, 0, 0.5], 90)
        np.save(os.path.join(out_dir, "single.npy"), single)
    dist.barrier()
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
    r.close()
    dist.destroy_process_group()


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
