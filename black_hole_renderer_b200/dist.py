"""Multi-GPU partitioning of the render path (one process per GPU, torch.distributed).

Two workloads shard:
  * orbit video: frames are independent -> `driver.frame_owner`, no data-path collective;
  * one large frame: row tiles.  The ray march is independent per row; the bloom's vertical pass
    needs `radius` rows of the horizontally blurred layer from each neighbour (one exchange
    step, NCCL send/recv over NVLink), the flare needs three global sums (one all-reduce), and
    the finished tiles are gathered on rank 0.

The functions below work on torch tensors of either backend: CUDA tensors aliasing the context's
device buffers under NCCL, CPU tensors under gloo (tests/test_dist_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def tile_rows(height, world_size, rank):
    """Row range [row0, row1) of `rank`: contiguous tiles, sizes differing by at most one row."""
    base, extra = divmod(height, world_size)
    row0 = rank * base + min(rank, extra)
    return row0, row0 + base + (1 if rank < extra else 0)


def owners_of_rows(height, world_size, lo, hi):
    """[(rank, a, b)]: which ranks own the rows [lo, hi) (clipped to the frame)."""
    lo, hi = max(lo, 0), min(hi, height)
    out = []
    for r in range(world_size):
        r0, r1 = tile_rows(height, world_size, r)
        a, b = max(lo, r0), min(hi, r1)
        if a < b:
            out.append((r, a, b))
    return out


def exchange_halos(plane, height, radius, rank, world_size, group=None):
    """Fill rows [row0 - radius, row0) and [row1, row1 + radius) of `plane` (C, H, W) with the
    neighbours' data; every rank sends the rows of its own tile that others need.  Handles tiles
    shorter than the radius (a halo may span several ranks)."""
    if world_size == 1:
        return
    row0, row1 = tile_rows(height, world_size, rank)
    ops, keep = [], []
    # what I need from others
    for lo, hi in ((row0 - radius, row0), (row1, row1 + radius)):
        for r, a, b in owners_of_rows(height, world_size, lo, hi):
            if r == rank:
                continue
            buf = torch.empty_like(plane[:, a:b, :]).contiguous()
            keep.append((buf, a, b))
            ops.append(dist.P2POp(dist.irecv, buf, r, group=group))
    # what others need from me
    for r in range(world_size):
        if r == rank:
            continue
        p0, p1 = tile_rows(height, world_size, r)
        for lo, hi in ((p0 - radius, p0), (p1, p1 + radius)):
            a, b = max(lo, row0, 0), min(hi, row1, height)
            if a < b:
                ops.append(dist.P2POp(dist.isend, plane[:, a:b, :].contiguous(), r, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, a, b in keep:
        plane[:, a:b, :].copy_(buf)


def gather_rows(tile, height, rank, world_size, dst=0, group=None):
    """Gather the (rows, W, C) tiles on `dst`; returns the (H, W, C) frame there, None elsewhere."""
    if world_size == 1:
        return tile
    if rank == dst:
        parts = []
        for r in range(world_size):
            r0, r1 = tile_rows(height, world_size, r)
            parts.append(tile if r == dst else torch.empty((r1 - r0,) + tuple(tile.shape[1:]),
                                                           dtype=tile.dtype, device=tile.device))
        reqs = [dist.irecv(parts[r], r, group=group) for r in range(world_size) if r != dst]
        for q in reqs:
            q.wait()
        return torch.cat(parts, dim=0)
    dist.send(tile.contiguous(), dst, group=group)
    return None


class _CudaView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def device_tensor(renderer, buf_id, shape, dtype=torch.float32):
    """torch tensor aliasing one of the renderer's device buffers (no copy)."""
    ptr, _ = renderer.device_buffer(buf_id)
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device=f"cuda:{renderer.cuda_device}")


def render_tiled(renderer, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False,
                 rank=0, world_size=1, group=None, want_u8=False, copy=True):
    """One frame split into row tiles over `world_size` GPUs (render.py:3865-3923 semantics).

    stage 1 (ray march + horizontal bloom pass on my rows) -> halo exchange of the H-blurred layer
    -> all-reduce of the flare sums -> stage 2 (vertical pass + composite + flare on my rows)
    -> gather on rank 0.  Returns the (H, W, 3) frame (numpy) on rank 0, None elsewhere; with
    copy=False the array aliases a per-renderer pinned buffer that the next call overwrites.
    """
    import ctypes as C
    H, W = renderer.height, renderer.width
    row0, row1 = tile_rows(H, world_size, rank)
    lib, ctx = renderer._lib, renderer._ctx
    stream = torch.cuda.current_stream(renderer.cuda_device)
    if stream.cuda_stream == 0:           # legacy default stream: give torch and the kernels a real one
        stream = torch.cuda.Stream(renderer.cuda_device)
        torch.cuda.set_stream(stream)
    renderer.set_stream(stream.cuda_stream)
    cam = renderer._camera(cam_pos, fov, frame)
    flags = renderer._flags(skip_differentials, skip_bloom)
    L.check(ctx, lib.bhr_render_rows_stage1(ctx, C.byref(cam), flags, row0, row1))
    if not skip_bloom:
        hblur = device_tensor(renderer, L.BUF_HBLUR, (3, H, W))
        exchange_halos(hblur, H, renderer._lib.bhr_bloom_radius(ctx), rank, world_size, group)
    sums_ptr = None
    if renderer.lens_flare:
        sums = (C.c_double * 3)()
        L.check(ctx, lib.bhr_flare_sums(ctx, row0, row1, sums))
        t = torch.tensor(list(sums), dtype=torch.float64, device=f"cuda:{renderer.cuda_device}")
        if world_size > 1:
            dist.all_reduce(t, group=group)
        vals = t.cpu().tolist()
        sums_ptr = (C.c_double * 3)(*vals)
    L.check(ctx, lib.bhr_render_rows_stage2(ctx, flags, row0, row1, sums_ptr))
    if want_u8:
        full = device_tensor(renderer, L.BUF_FINAL_U8, (H, W, 3), torch.uint8)
    else:
        full = device_tensor(renderer, L.BUF_FINAL, (H, W, 3))
    frame_t = gather_rows(full[row0:row1], H, rank, world_size, 0, group)
    if frame_t is None:
        return None
    # page-locked landing buffer (cached per renderer and dtype): the D2H copy runs at PCIe speed
    cache = renderer.__dict__.setdefault("_tiled_host", {})
    host = cache.get(frame_t.dtype)
    if host is None or host.shape != frame_t.shape:
        host = cache[frame_t.dtype] = torch.empty(frame_t.shape, dtype=frame_t.dtype, pin_memory=True)
    host.copy_(frame_t, non_blocking=True)
    torch.cuda.current_stream(renderer.cuda_device).synchronize()
    return host.numpy().copy() if copy else host.numpy()     # copy=False: valid until the next call


# ----------------------------------------------------------------------------------------------
# The same split with peer memory instead of NCCL (csrc/peer.cu): the kernels read their halo rows
# from the neighbours' HBM and store the finished rows into rank 0's buffers over NVLink.
# torch.distributed is used once, to pass the CUDA IPC handles around.
# ----------------------------------------------------------------------------------------------
def attach_peers(renderer, rank, world_size, group=None):
    """Exchange the IPC handles of the renderer's shared buffers between the ranks and map the
    peers' buffers (once per renderer; every rank must call it)."""
    import ctypes as C
    handles = (C.c_ubyte * 256)()
    L.check(renderer._ctx, renderer._lib.bhr_peer_export(renderer._ctx, handles))
    everyone = [None] * world_size
    dist.all_gather_object(everyone, bytes(handles), group=group)
    blob = b"".join(everyone)
    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    L.check(renderer._ctx, renderer._lib.bhr_peer_attach(renderer._ctx, rank, world_size, buf))
    renderer._peer_rank, renderer._peer_world = rank, world_size


def attach_shared_frame(renderer, rank, world_size, group=None, dtype=np.uint8):
    """Distributed egress: one (H, W, 3) host frame in POSIX shared memory, mapped and page-locked
    by every rank, so that each rank copies its own rows to the host over its own PCIe link.
    Returns the numpy view (rank 0's result buffer).  Call once per renderer after attach_peers."""
    from multiprocessing import shared_memory
    nbytes = renderer.height * renderer.width * 3 * np.dtype(dtype).itemsize
    names = [None]
    if rank == 0:
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        names = [shm.name]
    dist.broadcast_object_list(names, src=0, group=group)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=names[0])
        try:        # Python < 3.13 registers attached segments with the resource tracker, which would
            from multiprocessing import resource_tracker      # unlink them again (noisily) at exit
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    frame = np.ndarray((renderer.height, renderer.width, 3), dtype=dtype, buffer=shm.buf)
    L.check(renderer._ctx, renderer._lib.bhr_host_register(frame.ctypes.data, nbytes))
    L.check(renderer._ctx, renderer._lib.bhr_peer_set_distributed_egress(renderer._ctx, 1))
    dist.barrier(group=group)
    if rank == 0:
        shm.unlink()                      # the mappings keep it alive; nothing is left behind in /dev/shm
    renderer._shared_frame, renderer._shared_shm = frame, shm
    return frame


def render_tiled_peer(renderer, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False, out=None):
    """One frame over all attached GPUs.  Rank 0 returns the (H, W, 3) frame -- 8-bit, or float32
    if the shared frame was created with that dtype -- in `out` (a `renderer.pinned_frame(np.uint8)`
    array, allocated on first use) or, after `attach_shared_frame`, in the shared host frame that
    every rank fills with its own rows; the other ranks return None as soon as their tile is
    enqueued.  No collective and no host synchronisation except rank 0's final wait."""
    import ctypes as C
    cam = renderer._camera(cam_pos, fov, frame)
    flags = renderer._flags(skip_differentials, skip_bloom)
    shared = renderer.__dict__.get("_shared_frame")
    if shared is not None:
        out = shared
    elif renderer._peer_rank == 0:
        if out is None:
            out = renderer.__dict__.get("_peer_out")
            if out is None:
                out = renderer._peer_out = renderer.pinned_frame(np.uint8)
    else:
        out = None
    f32 = u8 = None
    if out is not None:
        assert out.dtype in (np.uint8, np.float32) and out.flags.c_contiguous
        if out.dtype == np.uint8:
            u8 = out.ctypes.data
        else:
            f32 = out.ctypes.data
    L.check(renderer._ctx, renderer._lib.bhr_render_tiled_peer(renderer._ctx, C.byref(cam), flags, f32, u8))
    return out if renderer._peer_rank == 0 else None
