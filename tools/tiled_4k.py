"""BASELINE.json configs[2]: one 4K frame (anti_alias lod_radius, tilt 20, flare) split into row
tiles over the ranks of a torchrun job (NCCL halo exchange + flare all-reduce + gather on rank 0).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/tiled_4k.py
Prints ms/frame (max over ranks, host clock around synchronised frames) and checks the tiled
frame against the one-GPU frame on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
from util import synthetic_disk_texture
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.dist import render_tiled
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
from black_hole_renderer_b200.skybox import generate_skybox

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, pov, fov = 3840, 2160, [6.0, 0.0, 0.5], 90.0
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
sky = generate_skybox(2048, 1024, seed=42, n_stars=6000).astype(np.float32)
r = Renderer(W, H, sky, synthetic_disk_texture(n_r, n_phi), anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True,
             cuda_device=local)
for _ in range(3):
    frame = render_tiled(r, pov, fov, rank=rank, world_size=world, want_u8=True, copy=False)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 10
t0 = time.perf_counter()
for _ in range(K):
    frame = render_tiled(r, pov, fov, rank=rank, world_size=world, want_u8=True, copy=False)
torch.cuda.synchronize()
ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / K], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    single = r.render_u8(pov, fov)
    d = np.abs(frame.astype(int) - single.astype(int))
    print(f"4K AA+tilt+flare tiled over {world} GPU(s): {ms.item():.3f} ms/frame (u8 frame gathered to rank 0's host), "
          f"{W * H / ms.item() / 1e3:.1f} Mrays/s; vs one-GPU frame: max|d| {d.max()}, differing pixels {(d.max(-1) > 0).sum()}")
# ---- the same frame through peer memory (csrc/peer.cu) ----
if world > 1:
    from black_hole_renderer_b200.dist import attach_peers, render_tiled_peer
    attach_peers(r, rank, world)
    for _ in range(3):
        f2 = render_tiled_peer(r, pov, fov)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        f2 = render_tiled_peer(r, pov, fov)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / K], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        d = np.abs(f2.astype(int) - single.astype(int))
        print(f"  peer-memory path (no NCCL on the data path): {ms.item():.3f} ms/frame, {W * H / ms.item() / 1e3:.1f} Mrays/s; "
              f"vs one-GPU frame: max|d| {d.max()}, differing pixels {(d.max(-1) > 0).sum()}")
    # ---- peer memory + distributed egress: every rank copies its own rows into a shared host frame ----
    from black_hole_renderer_b200.dist import attach_shared_frame
    attach_shared_frame(r, rank, world)
    for _ in range(3):
        f3 = render_tiled_peer(r, pov, fov)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        f3 = render_tiled_peer(r, pov, fov)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / K], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        d = np.abs(f3.astype(int) - single.astype(int))
        print(f"  peer memory + distributed egress (each rank's rows over its own PCIe link): {ms.item():.3f} ms/frame, "
              f"{W * H / ms.item() / 1e3:.1f} Mrays/s; vs one-GPU frame: max|d| {d.max()}, differing pixels {(d.max(-1) > 0).sum()}")
    dist.barrier(); dist.destroy_process_group()
