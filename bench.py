#!/usr/bin/env python3
"""Benchmark of the per-pixel null-geodesic render path (BASELINE.json metric: Mrays/s and
ms/frame at 1080p; orbit frames/s at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--resolution fhd]

A step = one frame of BASELINE.json configs[1] (render.py -r fhd, default scene pov 6 0 0.5,
fov 90, step 0.1, r_max 10, anti_alias disabled; procedural skybox seed 42 and the lifecycle
disk texture at t = 0) through ray march + bloom + composite.  With N > 1 (torchrun, one process
per GPU) every rank renders its own frame of the orbit video per step (configs[4]'s camera path,
frames dealt round-robin, no collective): weak scaling, value = rays of all ranks / max-over-ranks
time.  `value` is timed with CUDA events on the launching stream with the result left in HBM;
`e2e` is the same frame through Renderer.render() into a pinned HOST buffer (D2H inside the
timed region).  L2 is flushed between timed steps (outside the event brackets).

Further blocks of the same JSON line (each timed on its own, after the headline):
  orbit_video_full  BASELINE.json configs[4] as a WHOLE JOB: all 3600 frames of `render.py --video
                    --orbit -r fhd` through driver.run_video_frames (the CLI's frame loop: 60-frame
                    blocks dealt round-robin, every rank ticks every frame, PNG / x264 off);
  tiled_4k          (N > 1) BASELINE.json configs[2]: one 4K frame (ray differentials + mip LOD, tilt
                    20, flare) split into row tiles -- NCCL path (halo send/recv, flare all-reduce,
                    gather to rank 0) and peer-memory path (csrc/peer.cu), equal and cost-balanced
                    tiles, rank-0 and distributed egress -- each compared with the one-GPU frame;
                    a tiled frame that differs beyond the flare-sum rounding makes the run FAIL;
  e2e.d2h           raw pinned D2H bandwidth of one frame-sized copy per rank, alone and with all
                    ranks copying at once (separates the host fabric from the renderer).

`--impl reference` times the reference's CPU implementation of the same path on the host cores:
Taichi is not installable here, so it is the oracle port (oracle/bhr_oracle.c,
OpenMP over all host threads) -- labelled kind "port".  Both arms render the same scene: skybox
seed 42 and the lifecycle disk texture at t = 0 (the CPU arm builds it with the oracle's
restatement of the texture pipeline).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = {"4k": (3840, 2160), "fhd": (1920, 1080), "hd": (1280, 720), "sd": (640, 360)}
FLOP_PER_STEP = 185          # algorithmic flop per RK4 step without differentials
FLOP_PER_STEP_DIFF = 521
POV, FOV = [6.0, 0.0, 0.5], 90.0


def scene_inputs(width, height):
    """Procedural inputs of the default scene: skybox(seed 42) and the disk-texture size."""
    from black_hole_renderer_b200.driver import compute_disk_texture_resolution
    from black_hole_renderer_b200.skybox import generate_skybox
    sky = generate_skybox(2048, 1024, seed=42, n_stars=6000).astype(np.float32)
    n_phi, n_r = compute_disk_texture_resolution(width, height, POV, FOV, 2.0, 15.0)
    return sky, n_r, n_phi


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def cpu_port_frame(width, height, sky, tex, threads=None):
    """One frame of the same path on the host cores with the oracle port; returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    if threads:
        O.lib().orc_set_num_threads(threads)
    t0 = time.perf_counter()
    r = O.render(width, height, POV, FOV, sky, tex)
    return time.perf_counter() - t0, r["total_steps"], O.lib().orc_num_threads()


def lifecycle_texture_cpu(n_r, n_phi, r_inner=2.0, r_outer=15.0):
    """The disk texture of `render.py` single frames (lifecycle system, seed 42, t = 0) built WITHOUT
    the GPU: host factories + the oracle's restatement of background / entity layer / statistics /
    compose (render.py:4079-4153).  For the CPU arms, so that both arms render the same scene."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from black_hole_renderer_b200.lifecycle import make_factories
    rng = np.random.default_rng(42)
    az_freq, az_shear = int(rng.integers(2, 5)), float(rng.uniform(2.0, 4.0))
    factories = make_factories(r_inner, r_outer, n_r, n_phi, seed=42)      # (seeded and pre-populated)
    edge, omega = O.edge_alpha(n_r), O.omega_rows(n_r, r_inner, r_outer)
    comp = np.zeros((13, n_r, n_phi), dtype=np.float32)
    for first in (True, False):              # _init_lifecycle_system, then _advance_lifecycle_frame(t=0, dt=0)
        if not first:
            for f in factories.values():
                f.tick(now=0.0, dt=0.0)
        O.generate_background(comp, az_freq, az_shear, r_inner, r_outer, 0.0)
        comp[5:11] = O.accumulate_entities(factories, 0.0, n_r, n_phi, omega)
        stats, rows = O.interactive_stats(comp, edge)
        tex = O.compose_texture(comp, omega, edge, stats, rows)
    return tex


def orbit_video_full(r, rank, world, dist, torch, n_frames=3600, block=60, png=False):
    """BASELINE.json configs[4] as a whole job: `render.py --video --orbit --n_frames 3600 --fps 36
    -r fhd` through the CLI's own frame loop (driver.run_video_frames, render.py:4437-4458): frames
    dealt to ranks in 60-frame blocks (the statistics cadence), every rank replays the host
    lifecycle ticks of ALL frames and, for its own frames, runs background + entity layer +
    [statistics on a block's first frame] + compose + mips + ray march + bloom + composite and
    copies the 8-bit frame to pinned host memory.  PNG / x264 encoding off (host I/O).  Timed from
    a barrier to the moment the slowest rank has its last frame in host memory.
    png=True: the frames leave the device as the deflate streams of their PNG files (csrc/png.cu) -- what the CLI
    writes to disk after adding the chunk framing; the D2H per frame shrinks to the stream's length."""
    from black_hole_renderer_b200.driver import frame_owner, png_ring, run_video_frames, video_ring
    per_rank = [sum(1 for f in range(n_frames) if frame_owner(f, world, block) == k) for k in range(world)]
    r.set_option("stage_timing", 0)
    t_setup = time.perf_counter()
    from black_hole_renderer_b200.lifecycle import init_lifecycle_system
    factories = init_lifecycle_system(r, r.dtex_h, r.dtex_w, seed=42)
    (png_ring if png else video_ring)(r, 48)      # the page-locked frame ring (one-off, ~0.3 s)
    r.synchronize()
    setup_s = time.perf_counter() - t_setup
    launches0 = r.launch_count()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    timing = {}
    rendered = run_video_frames(r, n_frames, FOV, POV, True, 360.0, 0.1, rank, world, factories=factories, timing=timing,
                                png=png)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    launches = r.launch_count() - launches0
    assert rendered == per_rank[rank], (rendered, per_rank)
    tt = torch.tensor([sec, setup_s, timing["host_foreign_ticks_s"], timing["host_own_frames_s"],
                       timing["host_blocked_on_device_s"]], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec, setup_s, h_tick, h_own, h_wait = (float(v) for v in tt.tolist())
    if png:
        nb = torch.tensor([float(timing["png_stream_bytes"])], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(nb)
        return {"frames": n_frames, "frames_per_s": n_frames / sec, "seconds_max_over_ranks": sec,
                "ms_per_frame_per_gpu": 1e3 * sec / max(per_rank),
                "mean_stream_bytes_per_frame": float(nb.item()) / n_frames, "raw_bytes_per_frame": 3 * r.width * r.height,
                "host_seconds_max_over_ranks": {"foreign_ticks": h_tick, "own_frames_texture_pass_and_render_calls": h_own,
                                                "blocked_on_device": h_wait},
                "gpu_launches_rank0": launches,
                "includes": "the orbit_video_full job with each frame Sub-filtered, run-length matched and Huffman coded ON THE "
                            "DEVICE (3 kernels after the composite); the host receives each frame's complete zlib stream "
                            "(copy sized from the previous frames' streams)",
                "excludes": "CRC-32 + file write + x264 mux (host I/O)"}
    return {"frames": n_frames, "frames_per_s": n_frames / sec, "seconds_max_over_ranks": sec,
            "per_rank_frames": per_rank, "ms_per_frame_per_gpu": 1e3 * sec / max(per_rank),
            "balance_ceiling": n_frames / (world * max(per_rank)),
            "setup_seconds": setup_s, "frames_per_s_including_setup": n_frames / (sec + setup_s),
            "host_seconds_max_over_ranks": {"foreign_ticks": h_tick, "own_frames_texture_pass_and_render_calls": h_own,
                                            "blocked_on_device": h_wait},
            "gpu_launches_rank0": launches,
            "includes": "the whole 3600-frame job of render.py --video --orbit -r fhd: host lifecycle ticks of all frames on every "
                        "rank; for a rank's own frames background + entity + compose + mips kernels, statistics on each block's "
                        "first frame, render, 8-bit frame D2H to pinned memory",
            "excludes": "PNG / x264 encoding (host I/O); setup (factory seeding, first texture) reported separately",
            "sharding": f"{block}-frame blocks round-robin, no collective"}


def d2h_probe(r, torch, dist, rank, world, nbytes, reps=8):
    """Raw D2H bandwidth of one frame-sized pinned copy (GB/s): every rank alone (the others idle)
    and all ranks at once.  Separates what the host fabric can take from what the renderer does."""
    from black_hole_renderer_b200 import _lib as L
    from black_hole_renderer_b200.dist import device_tensor
    dev = device_tensor(r, L.BUF_FINAL, (r.height, r.width, 3))
    host = torch.empty((r.height, r.width, 3), dtype=torch.float32, pin_memory=True)
    assert host.numel() * 4 == nbytes
    stream = torch.cuda.current_stream()

    def one_round():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            host.copy_(dev, non_blocking=True)
        e1.record(stream)
        e1.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    one_round()
    alone = [0.0] * world
    for k in range(world):
        if dist:
            dist.barrier()
        if k == rank:
            alone[k] = one_round()
    if dist:
        dist.barrier()
    together = one_round()
    t = torch.tensor(alone + [0.0] * world, dtype=torch.float64, device="cuda")
    t[world + rank] = together
    if dist:
        dist.all_reduce(t)
    v = t.tolist()
    return {"bytes": nbytes, "alone_gbs_per_rank": [round(x, 2) for x in v[:world]],
            "concurrent_gbs_per_rank": [round(x, 2) for x in v[world:]],
            "concurrent_gbs_total": round(sum(v[world:]), 2)}


def nvlink_kib(index):
    """(tx, rx) KiB moved over all NVLink links of GPU `index` since driver load (NVML throughput counters, payload
    bytes), or None when NVML does not expose them."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        v = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xffffffff),
                                                (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xffffffff)])
        if any(x.nvmlReturn != 0 for x in v):
            return None
        return int(v[0].value.ullVal), int(v[1].value.ullVal)
    except Exception:
        return None


def tiled_4k(Renderer, sky, rank, world, local, dist, torch, frames=8):
    """BASELINE.json configs[2]: one 4K frame (anti_alias lod_radius, tilt 20, lens flare, lifecycle
    texture 832 x 5824) row-tiled over the ranks; 8-bit frame in rank 0's host memory at the end of
    every frame.  Times (ms/frame): CUDA events on rank 0's stream around `frames` back-to-back
    frames -- rank 0's stream ends with the gather / the wait for every tile, so its event span
    covers all ranks -- next to the host clock from a barrier to the last synchronised frame (max
    over ranks).  Every variant's last frame is compared with the one-GPU frame; the flare centroid
    is summed per tile, so a handful of last-bit differences is the allowance, anything more FAILS
    the run."""
    from black_hole_renderer_b200 import dist as D
    from black_hole_renderer_b200.driver import compute_disk_texture_resolution
    from black_hole_renderer_b200.lifecycle import advance_lifecycle_frame, init_lifecycle_system
    W, H = RES["4k"]
    n_phi, n_r = compute_disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
    r = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32), anti_alias="lod_radius", disk_tilt=20.0,
                 lens_flare=True, cuda_device=local)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    r.set_option("stage_timing", 0)
    factories = init_lifecycle_system(r, n_r, n_phi, seed=42)
    advance_lifecycle_frame(r, factories, t=0.0, dt=0.0, recompute_stats=True)
    r.synchronize()
    out = {"workload": "render.py -r 4k --anti_alias lod_radius --disk_tilt 20 --lens_flare (BASELINE.json configs[2]), "
                       "lifecycle disk texture at t=0, 8-bit frame to rank 0's host memory every frame",
           "resolution": [W, H], "disk_texture": [n_r, n_phi], "frames_timed": frames, "n_gpus": world}

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(frames):
            frame = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        host_ms = (time.perf_counter() - t0) * 1e3 / frames
        t = torch.tensor([e0.elapsed_time(e1) / frames, host_ms], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, host_ms = t.tolist()
        return frame, dev_ms, host_ms

    # the one-GPU frame (rank 0 alone; the others wait) -- reference pixels and the strong-scaling base
    single = None
    if rank == 0:
        pinned = r.pinned_frame(np.uint8)
        for _ in range(2):
            r.render_u8(POV, FOV, out=pinned)
        t0 = time.perf_counter()
        for _ in range(frames):
            r.render_u8(POV, FOV, out=pinned)
        out["single_gpu_ms"] = (time.perf_counter() - t0) * 1e3 / frames
        single = pinned.copy()
    if world == 1:
        r.close()
        return out
    dist.barrier()

    def compare(name, frame):
        if rank != 0:
            return
        d = np.abs(frame.astype(np.int16) - single.astype(np.int16))
        out[name]["max_abs_diff"] = int(d.max())
        out[name]["differing_pixels_vs_single_gpu"] = int((d.max(axis=-1) > 0).sum())

    variants = [("nccl", lambda: D.render_tiled(r, POV, FOV, rank=rank, world_size=world, want_u8=True, copy=False))]
    frame, dev_ms, host_ms = timed(variants[0][1])
    out["nccl"] = {"ms": dev_ms, "host_clock_ms": host_ms,
                   "path": "stage 1 -> ncclSend/Recv halos of the H-blurred layer -> ncclAllReduce of the flare sums (device) "
                           "-> stage 2 -> tiles received into rank 0's frame buffer -> one D2H"}
    compare("nccl", frame)
    D.attach_peers(r, rank, world)
    frame, dev_ms, host_ms = timed(lambda: D.render_tiled_peer(r, POV, FOV))
    # NVLink evidence for the peer path: rate at which every rank reads its right-hand neighbour's H-blurred layer
    # (99.5 MB at 4K) through the CUDA-IPC mapping with the halo pull's access pattern, one rank at a time and all
    # together; PCIe 5 x16 cannot carry more than 64 GB/s
    import ctypes as C
    def probe():
        g = C.c_double()
        r._check(r._lib.bhr_peer_probe_read(r._ctx, (rank + 1) % world, 4, C.byref(g)))
        return g.value
    alone = 0.0
    for k in range(world):
        dist.barrier()
        if k == rank:
            alone = probe()
    dist.barrier()
    together = probe()
    pr = torch.tensor([alone, together], dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(pr) for _ in range(world)]
    dist.all_gather(gathered, pr)
    R = r._lib.bhr_bloom_radius(r._ctx)
    rows = H // world
    nv = {"peer_read_gbs_alone_per_rank": [round(float(g[0]), 1) for g in gathered],
          "peer_read_gbs_all_ranks_at_once": [round(float(g[1]), 1) for g in gathered],
          "halo_bytes_pulled_per_frame_inner_rank": 2 * min(R, rows) * W * 3 * 4,
          "tile_bytes_stored_to_rank0_per_frame": rows * W * 3,
          "nvml_link_counters": "not exposed in this VM (nvmlDeviceGetFieldValues NVLINK_THROUGHPUT returns not supported)"
                                if nvlink_kib(local) is None else "available"}
    out["peer"] = {"ms": dev_ms, "host_clock_ms": host_ms,
                   "nvlink_evidence": nv,
                   "path": "csrc/peer.cu: V pass loads halo rows from the neighbours' HBM, composite stores into rank 0's "
                           "buffers, release/acquire flags; equal-height tiles; rank 0 copies the frame out"}
    compare("peer", frame)
    bounds = D.balance_tiles(r, POV, FOV, rank, world)
    frame, dev_ms, host_ms = timed(lambda: D.render_tiled_peer(r, POV, FOV))
    out["peer_balanced"] = {"ms": dev_ms, "host_clock_ms": host_ms, "tile_bounds": bounds,
                            "stage1_ms_per_rank": getattr(r, "_tile_stage1_ms", None),
                            "path": "the same with tile heights balanced by RK4 evaluations per row, calibrated by the measured "
                                    "stage-1 time of every rank's tile (dist.balance_tiles)"}
    compare("peer_balanced", frame)
    D.attach_shared_frame(r, rank, world)
    frame, dev_ms, host_ms = timed(lambda: D.render_tiled_peer(r, POV, FOV))
    out["peer_balanced_egress"] = {"ms": dev_ms, "host_clock_ms": host_ms,
                                   "path": "balanced tiles + distributed egress: every rank copies its own rows into one shared, "
                                           "page-locked host frame over its own PCIe link"}
    compare("peer_balanced_egress", frame)
    # where a tiled frame's time goes on every rank: the library's stage events around one more frame
    # (ray march, H pass, [h_ready barrier + halo pull +] V pass + composite, flare; total = first to last event)
    r.set_option("stage_timing", 1)
    D.render_tiled_peer(r, POV, FOV)
    torch.cuda.synchronize()
    st = r.last_stage_ms()
    r.set_option("stage_timing", 0)
    stages = [None] * world
    dist.all_gather_object(stages, {k: round(float(v), 4) for k, v in st.items()})
    out["peer_balanced_egress"]["stage_ms_per_rank"] = stages
    # pipelined: a second shared host frame; the call for frame s returns frame s - 1
    D.attach_shared_frame(r, rank, world)

    def pipelined_frames(n):
        last = None
        for _ in range(n):
            D.render_tiled_peer_async(r, POV, FOV)
            last = D.wait_tiled_frame(r, back=1)
        return D.wait_tiled_frame(r, back=0)

    pipelined_frames(3)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    frame = pipelined_frames(frames)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / frames], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["peer_balanced_egress_pipelined"] = {
        "ms": float(t.item()), "host_clock_ms": float(t.item()),
        "path": "balanced tiles + distributed egress on the copy streams: the rows of frame s leave over PCIe while frame "
                "s + 1 is ray marched (two shared host frames; the caller collects frame s - 1 after enqueueing frame s); "
                "host clock from a barrier to the last frame in host memory, max over ranks"}
    compare("peer_balanced_egress_pipelined", frame)
    frame, dev_ms, host_ms = timed(lambda: D.render_tiled(r, POV, FOV, rank=rank, world_size=world, want_u8=True, copy=False,
                                                          bounds=bounds))
    out["nccl_balanced_egress"] = {"ms": dev_ms, "host_clock_ms": host_ms,
                                   "path": "NCCL halos + all-reduce, balanced tiles, every rank's own D2H into the shared host frame"}
    compare("nccl_balanced_egress", frame)
    dist.barrier()
    ok = True
    if rank == 0:
        names = ("nccl", "peer", "peer_balanced", "peer_balanced_egress", "peer_balanced_egress_pipelined", "nccl_balanced_egress")
        best = min(out[k]["ms"] for k in names)
        out["best_ms"] = best
        out["strong_scaling_efficiency_vs_single_gpu"] = out["single_gpu_ms"] / (world * best)
        limit = max(8, int(1e-5 * W * H))
        for k in names:
            ok = ok and out[k]["max_abs_diff"] <= 1 and out[k]["differing_pixels_vs_single_gpu"] <= limit
        out["parity_ok"] = bool(ok)
        out["parity_rule"] = f"per variant: max |delta| <= 1 (8-bit) and <= {limit} differing pixels vs the one-GPU frame"
    r.close()
    return out, ok


def other_configs(Renderer, sky, peak, stream, frames=5):
    """configs[2] (4K, ray differentials + mip LOD, tilt 20, flare) and configs[3] (fhd, step 0.02,
    r_max 30): best-of-`frames` device time of one frame with the result left in HBM, the ray
    march's share of it and its fraction of the measured FP32 peak (SURVEY.md 8d flop counts)."""
    import torch
    from black_hole_renderer_b200.driver import compute_disk_texture_resolution
    from black_hole_renderer_b200.lifecycle import advance_lifecycle_frame, init_lifecycle_system
    out = {}
    cases = {"configs[2] 4k anti_alias=lod_radius disk_tilt=20 lens_flare": ("4k", dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True), FLOP_PER_STEP_DIFF),
             "configs[3] fhd step_size=0.02 r_max=30": ("fhd", dict(step_size=0.02, r_max=30.0), FLOP_PER_STEP)}
    for name, (res, kw, flop) in cases.items():
        W, H = RES[res]
        n_phi, n_r = compute_disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
        r = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32), cuda_device=torch.cuda.current_device(), **kw)
        r.set_stream(stream.cuda_stream)
        advance_lifecycle_frame(r, init_lifecycle_system(r, n_r, n_phi, seed=42), t=0.0, dt=0.0, recompute_stats=True)
        best = None
        for i in range(frames + 2):
            r.render_device(POV, FOV)
            r.synchronize()
            ms = r.last_stage_ms()
            if i >= 2 and (best is None or ms["total"] < best["total"]):
                best = ms
        steps = r.last_total_steps()
        tf = flop * steps / (best["ray_march"] * 1e-3) / 1e12
        out[name] = {"ms_per_frame": best["total"], "Mrays_per_s": W * H / (best["total"] * 1e-3) / 1e6,
                     "stage_ms": {k: best[k] for k in ("ray_march", "bloom_h", "bloom_v_composite", "flare")},
                     "rk4_steps_per_frame": steps, "ray_march_tflops": tf, "frac_of_fp32_peak": (tf / peak) if peak else None,
                     "flop_per_step": flop, "disk_texture": [n_r, n_phi]}
        r.close()
    return out


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the newest committed `ncu --set full`
    capture (profiles/*_raymarch_ncu.txt: dram__bytes_read.sum + dram__bytes_write.sum)."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_raymarch_ncu.txt")), reverse=True):
        blocks = open(path).read().split("=====")
        for b in blocks:
            if "raymarch_persistent<0" not in b:
                continue
            tot = 0.0
            for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                m = re.search(re.escape(name) + r" \[(\w+)\] = ([0-9.]+)", b)
                if not m:
                    break
                tot += float(m.group(2)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(1)]
            else:
                return tot, os.path.relpath(path, ROOT)
    return None, None


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, H = RES[args.resolution]
    sky, n_r, n_phi = scene_inputs(W, H)
    tex = lifecycle_texture_cpu(n_r, n_phi)       # the same scene as the GPU arm: lifecycle texture at t = 0
    times = []
    all_threads = os.cpu_count() or 1         # torchrun exports OMP_NUM_THREADS=1: use every host thread
    for i in range(args.warmup + args.steps):
        dt, steps, threads = cpu_port_frame(W, H, sky, tex, threads=all_threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = W * H / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.resolution, W, H, n_r, n_phi, 1, None),
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "every step is one full frame on the host cores (OpenMP oracle "
                                       "port; Taichi not installable, SURVEY.md 8c)"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(resolution, W, H, n_r, n_phi, world, mode):
    """The `config` object: identical in both arms (same workload string, same scene)."""
    return {"workload": (f"render.py -r {resolution} default scene (pov 6 0 0.5, fov 90, step 0.1, r_max 10), single "
                         "frame, anti_alias disabled = BASELINE.json configs[1]; procedural skybox seed 42 + "
                         "lifecycle disk texture at t=0; ray march + bloom + composite"),
            "resolution": [W, H], "rays_per_frame": W * H, "disk_texture": [n_r, n_phi]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--resolution", default="fhd", choices=list(RES))
    ap.add_argument("--mode", default=None, help="raymarch mode override: fast | strict")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-orbit", action="store_true")
    ap.add_argument("--no-tiled", action="store_true")
    ap.add_argument("--orbit-frames", type=int, default=3600)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # host placement first: threads and the pinned frame buffers allocated below follow the affinity
    from black_hole_renderer_b200 import hostmem
    placement = hostmem.bind_to_gpu(local)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"

    from black_hole_renderer_b200 import Renderer, _lib as L
    from black_hole_renderer_b200.driver import orbit_camera
    from black_hole_renderer_b200.lifecycle import advance_lifecycle_frame, init_lifecycle_system

    W, H = RES[args.resolution]
    sky, n_r, n_phi = scene_inputs(W, H)
    r = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32), cuda_device=local)
    if args.mode:
        r.set_option("raymarch_mode", {"fast": 0, "strict": 2}[args.mode])
    stream = torch.cuda.Stream()          # a real (non-legacy) stream: the kernels and the timing events share it
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    factories = init_lifecycle_system(r, n_r, n_phi, seed=42)
    advance_lifecycle_frame(r, factories, t=0.0, dt=0.0, recompute_stats=True)
    r.synchronize()

    n_frames_orbit = 3600
    def camera_of(step):
        if world == 1:
            return POV                              # configs[1]: the default still
        return orbit_camera(POV, (step * world + rank) % n_frames_orbit, n_frames_orbit, 360.0)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    peak = None
    if rank == 0:
        import ctypes as C
        v = C.c_double()
        if L.load().bhr_measure_fp32_peak(local, 0, C.byref(v)) == 0:
            peak = v.value

    for s in range(args.warmup):
        r.render_device(camera_of(s), FOV)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = {"ray_march": 0.0, "bloom_h": 0.0, "bloom_v_composite": 0.0}
    total_steps = 0
    torch.cuda.synchronize()
    launches0 = r.launch_count()
    for s in range(args.steps):
        flush.zero_()
        # the library's own per-stage timing events (instrumentation, five per frame) are recorded on
        # every fourth frame only -- they feed stage_ms / the roofline; the step timing is evs[]
        sampled = s % 4 == 3 or s == args.steps - 1
        r.set_option("stage_timing", 1 if sampled else 0)
        evs[s][0].record(stream)
        r.render_device(camera_of(args.warmup + s), FOV)
        evs[s][1].record(stream)
        if sampled:                                # stage timers of the latest frame (events already recorded)
            evs[s][1].synchronize()
            ms = r.last_stage_ms()
            for k in ("ray_march", "bloom_h", "bloom_v_composite"):
                stage[k] += ms[k]
            stage["n"] = stage.get("n", 0) + 1
            total_steps = r.last_total_steps()
    torch.cuda.synchronize()
    launches = r.launch_count() - launches0
    if world > 1:
        dist.barrier()
    r.set_option("stage_timing", 0)               # off for the end-to-end and video blocks below
    ms_step = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    t = torch.tensor([ms_step, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_step, launches = float(tmax[0].item()), int(t[1].item())
    else:
        ms_step, launches = float(t[0].item()), int(t[1].item())

    # ---- end to end through the public API: camera in, frame in host memory out ----
    out = r.pinned_frame(np.float32)
    for s in range(2):
        r.render(camera_of(s), FOV, out=out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        r.render(camera_of(args.warmup + s), FOV, out=out)      # synchronises: the frame is on the host
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    # the same frames through the pipelined public API (render_async / wait_frame): every float
    # frame still lands in host memory inside the timed region, but its copy overlaps the next frame
    bufs = [r.pinned_frame(np.float32) for _ in range(2)]
    r.render_async(camera_of(0), FOV, bufs[0], 0); r.wait_frame(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        r.render_async(camera_of(args.warmup + s), FOV, bufs[s % 2], s % 2)
        if s > 0:
            r.wait_frame((s - 1) % 2)
    r.wait_frame((args.steps - 1) % 2)
    pipe_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    # ... and as 8-bit frames (what render_image / render_video save: 6.2 MB instead of 24.9 MB)
    u8 = r.pinned_frame(np.uint8)
    r.render_u8(camera_of(0), FOV, out=u8)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        r.render_u8(camera_of(args.warmup + s), FOV, out=u8)
    u8_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([pipe_ms, u8_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pipe_ms, u8_ms = (float(v) for v in t.tolist())
    d2h = d2h_probe(r, torch, dist if world > 1 else None, rank, world, W * H * 12)

    # ---- orbit video (BASELINE.json configs[4]): the whole 3600-frame job ----
    orbit = orbit_png = None
    if not args.no_orbit:
        orbit = orbit_video_full(r, rank, world, dist if world > 1 else None, torch, n_frames=args.orbit_frames)
        orbit_png = orbit_video_full(r, rank, world, dist if world > 1 else None, torch, n_frames=args.orbit_frames, png=True)
    clocks = sampler.stop() if sampler else None      # sampled every 20 ms over all the timed regions above

    # ---- the other BASELINE.json configurations, device-resident, rank 0 at N = 1 (parity-tested in
    #      tests/; timed here so that every config has a number next to its roofline) ----
    other = None
    if world == 1 and not args.no_other_configs:
        other = other_configs(Renderer, sky, peak, stream)
    # ---- configs[2] row-tiled over the ranks (N > 1); its one-GPU end-to-end number at N = 1 ----
    tiled, tiled_ok = None, True
    if not args.no_tiled:
        res = tiled_4k(Renderer, sky, rank, world, local, dist if world > 1 else None, torch)
        tiled, tiled_ok = res if isinstance(res, tuple) else (res, True)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    n = stage.pop("n", 1)
    stage = {k: v / n for k, v in stage.items()}
    rays = W * H * world
    flops = FLOP_PER_STEP * total_steps
    achieved = flops / (stage["ray_march"] * 1e-3) / 1e12
    traffic, traffic_src = profiled_traffic()
    nominal = 148 * 128 * 2 * 1.965e9 / 1e12          # SMs x FP32 lanes x 2 flop x boost clock
    t_copy_ms = W * H * 12 / (min(d2h["concurrent_gbs_per_rank"]) * 1e9) * 1e3
    cfg = workload_config(args.resolution, W, H, n_r, n_phi, world, args.mode)
    cfg.update({"rays_per_step": rays,
                "multi_gpu": None if world == 1 else "N>1: one orbit-video frame (configs[4] camera path) per rank per "
                                                     "step, frames sharded, no collective",
                "l2": "flushed (256 MiB memset) between timed steps, outside the event brackets",
                "raymarch_mode": args.mode or "default", "host_placement": placement})
    line = {
        "metric": "Mrays/s", "value": rays / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "ms_per_frame": ms_step, "frames_per_s": world / (ms_step * 1e-3),
        "stage_ms": stage, "rk4_steps_per_frame": total_steps,
        "e2e": {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
                "h2d_bytes_per_step": 64, "d2h_bytes_per_step": W * H * 12,
                "note": "Renderer.render(cam, fov, out=pinned (H,W,3) f32): camera struct H2D, frame D2H, host sync; the "
                        "frame is finished in 3 row bands (photon-ring rows first) so that a band's D2H overlaps the next band's ray march",
                "pipelined": {"value": rays / (pipe_ms * 1e-3) / 1e6, "ms_per_frame": pipe_ms,
                              "note": "same frames and bytes through Renderer.render_async / wait_frame (two pinned "
                                      "buffers): the D2H of frame i overlaps the ray march of frame i + 1"},
                "u8": {"value": rays / (u8_ms * 1e-3) / 1e6, "ms_per_frame": u8_ms, "d2h_bytes_per_step": W * H * 3,
                       "note": "Renderer.render_u8: the 8-bit frame the drivers save (render.py:423, 4463)"},
                "d2h": d2h,
                "limiter": ("host D2H bandwidth" if t_copy_ms > ms_step else "render"),
                "limiter_arithmetic": (f"one {W * H * 12 / 1e6:.1f} MB frame at the slowest rank's concurrent D2H rate "
                                       f"({min(d2h['concurrent_gbs_per_rank']):.1f} GB/s with all {world} rank(s) copying) = "
                                       f"{t_copy_ms:.3f} ms vs {ms_step:.3f} ms to render it")},
        "gpu_launches": launches,                 # counted by the library (bhr_launch_count), all ranks, timed region of `value`
        "roofline": {"bound": "fp32", "kernel": "raymarch_persistent (+ band_list, retrace)", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": (achieved / peak) if peak else None,
                     "frac_of_nominal": achieved / nominal, "nominal_peak": nominal, "traffic": traffic,
                     "traffic_note": (f"dram read + write bytes per launch from {traffic_src} (ncu --set full); the kernel is "
                                      "FP32-bound, its algorithmic DRAM traffic is the two 24.9 MB layers it writes") if traffic else None,
                     "algorithmic": f"{FLOP_PER_STEP} flop/RK4 step x {total_steps} steps (SURVEY.md 8d)",
                     "peak_source": "scalar FFMA microbenchmark measured in this run (bhr_measure_fp32_peak); "
                                    "MEASURED_PEAKS.json holds no FP32 figure; nominal = 148 SM x 128 lanes x 2 x 1.965 GHz"},
        "clocks": clocks,
    }
    if orbit:
        line["orbit_video_full"] = orbit
        if orbit_png is not None:
            line["orbit_video_device_png"] = orbit_png
    if other:
        line["other_configs"] = other
    if tiled:
        line["tiled_4k"] = tiled
    if not args.no_cpu_baseline and world == 1:      # CPU baseline: rank 0 at N = 1 only
        tex = r.disk_texture_field.to_numpy()
        times = []
        for _ in range(3):
            dt, _, threads = cpu_port_frame(W, H, sky, tex, threads=os.cpu_count() or 1)
            times.append(dt)
        line["cpu_baseline"] = {"value": W * H / min(times[1:]) / 1e6, "unit": "Mrays/s", "cores": threads,
                                "kind": "port", "ms_per_frame": 1e3 * min(times[1:]),
                                "sample": "3 full frames of the same workload (best of the last 2) with the "
                                          "OpenMP oracle port on all host threads; Taichi is not installable"}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not tiled_ok:
        raise SystemExit("tiled_4k: a row-tiled frame differs from the one-GPU frame beyond the flare-sum rounding")


if __name__ == "__main__":
    main()
