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


def equal_bounds(height, world_size):
    """Tile boundaries [b_0 = 0, ..., b_world = height] of the equal-height split."""
    return [tile_rows(height, world_size, r)[0] for r in range(world_size)] + [height]


def balanced_bounds(row_costs, world_size, min_rows=1):
    """Tile boundaries that give every rank (nearly) the same total cost: b_r is the first row at
    which the running cost reaches r / world of the total.  `row_costs`: one non-negative number
    per image row (RK4 evaluations + a per-pixel post-processing term, see `balance_tiles`)."""
    c = np.asarray(row_costs, dtype=np.float64)
    H = len(c)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1] if cum[-1] > 0 else 1.0
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(cum, target, side="left"))
        if 0 < b <= H and abs(cum[b - 1] - target) < abs(cum[min(b, H)] - target):
            b -= 1                        # the nearer of the two candidate rows
        b = min(max(b, bounds[-1] + min_rows), H - (world_size - r) * min_rows)
        bounds.append(b)
    bounds.append(H)
    return bounds


def owners_of_rows(height, world_size, lo, hi, bounds=None):
    """[(rank, a, b)]: which ranks own the rows [lo, hi) (clipped to the frame)."""
    bounds = bounds or equal_bounds(height, world_size)
    lo, hi = max(lo, 0), min(hi, height)
    out = []
    for r in range(world_size):
        a, b = max(lo, bounds[r]), min(hi, bounds[r + 1])
        if a < b:
            out.append((r, a, b))
    return out


def exchange_halos(plane, height, radius, rank, world_size, group=None, bounds=None):
    """Fill rows [row0 - radius, row0) and [row1, row1 + radius) of `plane` (C, H, W) with the
    neighbours' data; every rank sends the rows of its own tile that others need.  Handles tiles
    shorter than the radius (a halo may span several ranks).  Rows of one channel are contiguous,
    so every transfer goes straight out of / into `plane[c, a:b]` -- one send / recv per channel
    and neighbour piece, no staging copies."""
    if world_size == 1:
        return
    bounds = bounds or equal_bounds(height, world_size)
    row0, row1 = bounds[rank], bounds[rank + 1]
    n_ch = plane.shape[0]
    ops = []
    # what I need from others
    for lo, hi in ((row0 - radius, row0), (row1, row1 + radius)):
        for r, a, b in owners_of_rows(height, world_size, lo, hi, bounds):
            if r == rank:
                continue
            for c in range(n_ch):
                ops.append(dist.P2POp(dist.irecv, plane[c, a:b], r, group=group))
    # what others need from me
    for r in range(world_size):
        if r == rank:
            continue
        p0, p1 = bounds[r], bounds[r + 1]
        for lo, hi in ((p0 - radius, p0), (p1, p1 + radius)):
            a, b = max(lo, row0, 0), min(hi, row1, height)
            if a < b:
                for c in range(n_ch):
                    ops.append(dist.P2POp(dist.isend, plane[c, a:b], r, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def gather_rows(tile, height, rank, world_size, dst=0, group=None, bounds=None, out=None):
    """Gather the (rows, W, C) tiles on `dst`; returns the (H, W, C) frame there, None elsewhere.
    `out`: the full (H, W, C) tensor on `dst` whose own rows `tile` aliases -- the other ranks' tiles
    are then received straight into their rows (no concatenation)."""
    if world_size == 1:
        return tile if out is None else out
    bounds = bounds or equal_bounds(height, world_size)
    if rank == dst:
        if out is not None:
            ops = [dist.P2POp(dist.irecv, out[bounds[r]:bounds[r + 1]], r, group=group) for r in range(world_size)
                   if r != dst and bounds[r + 1] > bounds[r]]
            for q in dist.batch_isend_irecv(ops) if ops else ():
                q.wait()
            return out
        parts = []
        for r in range(world_size):
            parts.append(tile if r == dst else torch.empty((bounds[r + 1] - bounds[r],) + tuple(tile.shape[1:]),
                                                           dtype=tile.dtype, device=tile.device))
        ops = [dist.P2POp(dist.irecv, parts[r], r, group=group) for r in range(world_size)
               if r != dst and parts[r].shape[0] > 0]
        for q in dist.batch_isend_irecv(ops) if ops else ():
            q.wait()
        return torch.cat(parts, dim=0)
    if tile.shape[0] > 0:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, tile.contiguous(), dst, group=group)]):
            q.wait()
    return None


class _CudaView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def device_tensor(renderer, buf_id, shape, dtype=torch.float32):
    """torch tensor aliasing one of the renderer's device buffers (no copy; cached per buffer)."""
    cache = renderer.__dict__.setdefault("_device_tensors", {})
    key = (buf_id, tuple(shape), dtype)
    t = cache.get(key)
    if t is None:
        ptr, _ = renderer.device_buffer(buf_id)
        typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4", torch.float64: "<f8"}[dtype]
        t = cache[key] = torch.as_tensor(_CudaView(ptr, shape, typestr), device=f"cuda:{renderer.cuda_device}")
    return t


def render_tiled(renderer, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False,
                 rank=0, world_size=1, group=None, want_u8=False, copy=True, bounds=None):
    """One frame split into row tiles over `world_size` GPUs (render.py:3865-3923 semantics).

    stage 1 (ray march + horizontal bloom pass on my rows) -> halo exchange of the H-blurred layer
    (NCCL send / recv straight out of and into the layer buffer) -> all-reduce of the three flare
    sums IN PLACE in the device buffer stage 2 reads (no host round trip) -> stage 2 (vertical pass
    + composite + flare on my rows) -> tiles received straight into rank 0's frame buffer -> one
    D2H on rank 0.  Returns the (H, W, 3) frame (numpy) on rank 0, None elsewhere; with copy=False
    the array aliases a per-renderer pinned buffer that the next call overwrites.  After
    `attach_shared_frame` (u8 frames) every rank copies its own rows into the shared host frame
    over its own PCIe link instead of the gather.
    """
    import ctypes as C
    H, W = renderer.height, renderer.width
    bounds = bounds or equal_bounds(H, world_size)
    row0, row1 = bounds[rank], bounds[rank + 1]
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
        exchange_halos(hblur, H, renderer._lib.bhr_bloom_radius(ctx), rank, world_size, group, bounds)
    if renderer.lens_flare:
        L.check(ctx, lib.bhr_flare_sums_device(ctx, row0, row1))
        if world_size > 1:
            dist.all_reduce(device_tensor(renderer, L.BUF_FLARE_SUMS, (3,), torch.float64), group=group)
        flags |= L.FLARE_FROM_DEVICE
    L.check(ctx, lib.bhr_render_rows_stage2(ctx, flags, row0, row1, None))
    renderer._disk_post_bloom = not skip_bloom
    shared = renderer.__dict__.get("_shared_frame")
    if shared is not None and world_size > 1:
        # distributed egress: my rows -> the shared, page-locked host frame over my own PCIe link
        assert want_u8 == (shared.dtype == np.uint8)
        buf = L.BUF_FINAL_U8 if want_u8 else L.BUF_FINAL
        full = device_tensor(renderer, buf, (H, W, 3), torch.uint8 if want_u8 else torch.float32)
        host = renderer.__dict__.get("_shared_frame_t")
        if host is None:
            host = renderer._shared_frame_t = torch.from_numpy(shared)
        host[row0:row1].copy_(full[row0:row1], non_blocking=True)
        stream.synchronize()
        dist.barrier(group=group)         # (host-side rendezvous: every rank's rows have landed)
        return (shared.copy() if copy else shared) if rank == 0 else None
    if want_u8:
        full = device_tensor(renderer, L.BUF_FINAL_U8, (H, W, 3), torch.uint8)
    else:
        full = device_tensor(renderer, L.BUF_FINAL, (H, W, 3))
    frame_t = gather_rows(full[row0:row1], H, rank, world_size, 0, group, bounds, out=full if rank == 0 else None)
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


def balance_tiles(renderer, cam_pos, fov, rank, world_size, group=None, post_cost_per_pixel=10.0,
                  skip_differentials=False, refine=2):
    """Cost-balanced tile boundaries for `render_tiled` / `render_tiled_peer`: rows through the hole
    and the disk cost more than sky rows.  Every rank traces its equal-height tile once with the
    per-pixel step counts on, reduces them to RK4 evaluations per row on the device
    (bhr_row_costs), the per-row costs are all-gathered, and every rank derives the same
    boundaries (`balanced_bounds`; cost of a row = its RK4 evaluations + `post_cost_per_pixel`
    step-equivalents per pixel for bloom / composite / egress).  For a camera path the boundaries
    of one frame serve its neighbours (the cost profile moves slowly).
    `refine` rounds of calibration follow: an RK4 evaluation does not cost the same everywhere (rays of
    the photon ring run in the strict integrator at about three times the cost, and a tile that holds
    them has the latency of a strict batch as its floor), so every rank times stage 1 on its tile, the
    times are all-gathered, the row costs of each tile are scaled by (measured time / modelled cost)
    and the bounds are cut again.  Returns the bounds list; with peers attached it is installed in the
    library as well."""
    import time
    import ctypes as C
    H, W = renderer.height, renderer.width
    eq = equal_bounds(H, world_size)
    row0, row1 = eq[rank], eq[rank + 1]
    lib, ctx = renderer._lib, renderer._ctx
    cam = renderer._camera(cam_pos, fov, 0)
    flags = renderer._flags(skip_differentials, True, aux=True)
    L.check(ctx, lib.bhr_render_rows_stage1(ctx, C.byref(cam), flags, row0, row1))
    costs = np.zeros(row1 - row0, dtype=np.uint64)
    L.check(ctx, lib.bhr_row_costs(ctx, row0, row1, costs.ctypes.data_as(C.POINTER(C.c_uint64))))
    everyone = [None] * world_size
    if world_size > 1:
        dist.all_gather_object(everyone, costs.tolist(), group=group)
    else:
        everyone = [costs.tolist()]
    rows = np.concatenate([np.asarray(e, dtype=np.float64) for e in everyone]) + post_cost_per_pixel * W
    bounds = balanced_bounds(rows, world_size, min_rows=8)
    def stage1_seconds(bounds):
        a, b = bounds[rank], bounds[rank + 1]
        best = float("inf")
        for _rep in range(3):                       # (the first repetition warms the tile's working set)
            renderer.synchronize()
            t0 = time.perf_counter()
            L.check(ctx, lib.bhr_render_rows_stage1(ctx, C.byref(cam), flags, a, b))
            renderer.synchronize()
            best = min(best, time.perf_counter() - t0)
        times = [None] * world_size
        dist.all_gather_object(times, best, group=group)
        return times

    for it in range(refine + 1 if world_size > 1 else 0):
        times = stage1_seconds(bounds)
        renderer._tile_stage1_ms = [round(1e3 * t, 4) for t in times]     # stage 1 per rank for `bounds` (diagnostics)
        if it == refine:
            break
        scaled = rows.copy()
        for r in range(world_size):
            model = rows[bounds[r]:bounds[r + 1]].sum()
            if model > 0:
                scaled[bounds[r]:bounds[r + 1]] *= times[r] / model
        rows = scaled
        bounds = balanced_bounds(rows, world_size, min_rows=8)
    if renderer.__dict__.get("_peer_world"):
        arr = (C.c_int * (world_size + 1))(*bounds)
        L.check(ctx, lib.bhr_peer_set_tiles(ctx, arr))
    renderer._tile_bounds = bounds
    return bounds


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
    if renderer.__dict__.get("_shared_frame") is None:
        renderer._shared_frame, renderer._shared_shm = frame, shm
    # every shared frame of this renderer: pipelined tiled frames alternate between the first two
    renderer.__dict__.setdefault("_shared_frames", []).append(frame)
    renderer.__dict__.setdefault("_shared_shms", []).append(shm)
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


def render_tiled_peer_async(renderer, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False):
    """Pipelined variant of `render_tiled_peer` (needs TWO shared host frames: call `attach_shared_frame` twice):
    enqueues the frame and returns; the rows leave every GPU on its copy stream while the next frame is ray
    marched.  Frames alternate between the two shared frames.  `wait_tiled_frame(renderer, back=1)` after the call
    for frame s returns frame s - 1 on rank 0; finish a sequence with `wait_tiled_frame(renderer, back=0)`."""
    import ctypes as C
    frames = renderer.__dict__.get("_shared_frames", [])
    assert len(frames) >= 2, "pipelined tiled frames alternate between two shared host frames"
    k = renderer.__dict__.get("_tiled_async_count", 0)
    out = frames[k & 1]
    renderer._tiled_async_count = k + 1
    cam = renderer._camera(cam_pos, fov, frame)
    flags = renderer._flags(skip_differentials, skip_bloom)
    f32 = out.ctypes.data if out.dtype == np.float32 else None
    u8 = out.ctypes.data if out.dtype == np.uint8 else None
    L.check(renderer._ctx, renderer._lib.bhr_render_tiled_peer_async(renderer._ctx, C.byref(cam), flags, f32, u8))


def wait_tiled_frame(renderer, back=1):
    """Block until the pipelined frame `back` calls ago is complete; rank 0 gets its host frame (valid until the next
    `render_tiled_peer_async` but one), the other ranks None.  Returns None when there is no such frame yet."""
    k = renderer.__dict__.get("_tiled_async_count", 0)
    L.check(renderer._ctx, renderer._lib.bhr_peer_wait_frame(renderer._ctx, int(back)))
    if k - 1 - back < 0 or renderer._peer_rank != 0:
        return None
    return renderer._shared_frames[(k - 1 - back) & 1]
