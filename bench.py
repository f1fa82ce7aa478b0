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

`--impl reference` times the reference's CPU implementation of the same path on the host cores:
Taichi is not installable here, so it is the oracle port (oracle/bhr_oracle.c,
OpenMP over all host threads) -- labelled kind "port".
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


def orbit_video_block(r, n_r, n_phi, rank, world, dist, torch, block=60, n_frames_total=3600):
    """Steady-state frames/s of the orbit video path (render.py --video --orbit --n_frames 3600
    -r fhd).  Frames are dealt to ranks in 60-frame blocks (the statistics cadence); every rank
    replays the host lifecycle ticks of ALL frames.  Timed: one full cycle of 60 x world frames
    per rank, starting at the rank's own block (the ticks up to there run before the clock starts):
    60 frames of background + entity layer + [statistics on the block's first frame] + compose +
    mips + ray march + bloom + composite with the 8-bit frame copied to pinned host memory, then
    the ticks of the other ranks' 60 x (world - 1) frames, which the host runs while the device
    drains the pipeline.  PNG / x264 encoding is excluded."""
    from black_hole_renderer_b200.driver import frame_owner, orbit_camera
    from black_hole_renderer_b200.lifecycle import advance_lifecycle_frame, init_lifecycle_system
    factories = init_lifecycle_system(r, n_r, n_phi, seed=42)
    depth = 7                                         # frames the host may run ahead (8 completion slots)
    bufs = [r.pinned_frame(np.uint8) for _ in range(depth + 1)]
    dt = 0.1
    for frame in range(block * rank):                 # untimed: bring the lifecycle to my block
        for f in factories.values():
            f.tick(now=frame * dt, dt=dt)
    r.render_u8(POV, FOV, out=bufs[0])
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    issued, in_flight = 0, []
    for frame in range(block * rank, block * rank + block * world):
        t = frame * dt
        if frame_owner(frame, world, block) != rank:
            for f in factories.values():
                f.tick(now=t, dt=dt)
            continue
        # pipelined like driver.render_video: the host runs up to `depth` frames ahead
        advance_lifecycle_frame(r, factories, t, dt, recompute_stats=(frame % block == 0))
        slot = issued % (depth + 1)
        r.render_u8_async(orbit_camera(POV, frame, n_frames_total, 360.0), FOV, bufs[slot], slot)
        in_flight.append(slot)
        if len(in_flight) > depth:
            r.wait_frame(in_flight.pop(0))
        issued += 1
    for slot in in_flight:
        r.wait_frame(slot)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec = float(tt.item())
    return {"frames_per_s": block * world / sec, "frames": block * world, "seconds_max_over_ranks": sec,
            "ms_per_frame_per_gpu": 1e3 * sec / block,
            "includes": "one steady-state cycle per rank: host lifecycle ticks of all 60 x N frames, and for the rank's own "
                        "60 frames background + entity + compose + mips kernels, statistics on the block's first frame, "
                        "render, 8-bit frame D2H to pinned memory",
            "excludes": "PNG / x264 encoding (host I/O)", "sharding": f"{block}-frame blocks round-robin, no collective"}


def other_configs(Renderer, sky, peak, stream, frames=5):
    """configs[2] (4K, ray differentials + mip LOD, tilt 20, flare) and configs[3] (fhd, step 0.02,
    r_max 30): best-of-`frames` device time of one frame with the result left in HBM, the ray
    march's share of it and its fraction of the measured FP32 peak (SURVEY.md 8d flop counts)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import synthetic_disk_texture
    from black_hole_renderer_b200.driver import compute_disk_texture_resolution
    out = {}
    cases = {"configs[2] 4k anti_alias=lod_radius disk_tilt=20 lens_flare": ("4k", dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True), FLOP_PER_STEP_DIFF),
             "configs[3] fhd step_size=0.02 r_max=30": ("fhd", dict(step_size=0.02, r_max=30.0), FLOP_PER_STEP)}
    for name, (res, kw, flop) in cases.items():
        W, H = RES[res]
        n_phi, n_r = compute_disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
        r = Renderer(W, H, sky, synthetic_disk_texture(n_r, n_phi), cuda_device=torch.cuda.current_device(), **kw)
        r.set_stream(stream.cuda_stream)
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
                     "stage_ms": {k: best[k] for k in ("ray_march", "bloom_h", "bloom_v_composite", "gap")},
                     "rk4_steps_per_frame": steps, "ray_march_tflops": tf, "frac_of_fp32_peak": (tf / peak) if peak else None,
                     "flop_per_step": flop, "disk_texture": "synthetic (texel values do not affect the timing)"}
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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import synthetic_disk_texture
    tex = synthetic_disk_texture(n_r, n_phi)      # texel values do not affect the CPU timing
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
            "config": {"workload": f"render.py -r {args.resolution} default scene, single frame, anti_alias "
                                   "disabled (BASELINE.json configs[1]); ray march + bloom + composite"},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "every step is one full frame on the host cores (OpenMP oracle "
                                       "port; Taichi not installable, SURVEY.md 8c)"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    if world > 1:
        dist.barrier()
    r.set_option("stage_timing", 0)               # off for the end-to-end and video blocks below
    ms_step = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    t = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())

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
    t = torch.tensor([pipe_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pipe_ms = float(t.item())

    # ---- orbit video (BASELINE.json configs[4]): one 60-frame block per rank, whole per-frame path ----
    orbit = orbit_video_block(r, n_r, n_phi, rank, world, dist if world > 1 else None, torch)
    clocks = sampler.stop() if sampler else None      # sampled every 20 ms over all the timed regions above

    # ---- the other BASELINE.json configurations, device-resident, rank 0 at N = 1 (parity-tested in
    #      tests/; timed here so that every config has a number next to its roofline) ----
    other = None
    if world == 1 and not args.no_other_configs:
        other = other_configs(Renderer, sky, peak, stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n = stage.pop("n", 1)
    stage = {k: v / n for k, v in stage.items()}
    rays = W * H * world
    flops = FLOP_PER_STEP * total_steps
    achieved = flops / (stage["ray_march"] * 1e-3) / 1e12
    traffic, traffic_src = profiled_traffic()
    line = {
        "metric": "Mrays/s", "value": rays / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"render.py -r {args.resolution} default scene (pov 6 0 0.5, fov 90, step 0.1, "
                                "r_max 10), single frame, anti_alias disabled = BASELINE.json configs[1]; "
                                "procedural skybox seed 42 + lifecycle disk texture at t=0"
                                + ("" if world == 1 else "; N>1: one orbit-video frame (configs[4] camera path) "
                                                        "per rank per step, frames sharded, no collective")),
                   "resolution": [W, H], "rays_per_step": rays, "disk_texture": [n_r, n_phi],
                   "l2": "flushed (256 MiB memset) between timed steps, outside the event brackets",
                   "raymarch_mode": args.mode or "default"},
        "ms_per_frame": ms_step, "frames_per_s": world / (ms_step * 1e-3),
        "stage_ms": stage, "rk4_steps_per_frame": total_steps,
        "e2e": {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
                "h2d_bytes_per_step": 64, "d2h_bytes_per_step": W * H * 12,
                "note": "Renderer.render(cam, fov, out=pinned (H,W,3) f32): camera struct H2D, frame D2H, host sync; the "
                        "frame is finished in 3 row bands (photon-ring rows first) so that a band's D2H overlaps the next band's ray march",
                "pipelined": {"value": rays / (pipe_ms * 1e-3) / 1e6, "ms_per_frame": pipe_ms,
                              "note": "same frames and bytes through Renderer.render_async / wait_frame (two pinned "
                                      "buffers): the D2H of frame i overlaps the ray march of frame i + 1"}},
        "gpu_launches": 6 * args.steps * world,   # band_list, raymarch_persistent, retrace, bloom_h, bloom_v, composite per frame
        "roofline": {"bound": "fp32", "kernel": "raymarch_persistent (+ band_list, retrace)", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": (achieved / peak) if peak else None, "traffic": traffic,
                     "traffic_note": (f"dram read + write bytes per launch from {traffic_src} (ncu --set full); the kernel is "
                                      "FP32-bound, its algorithmic DRAM traffic is the two 24.9 MB layers it writes") if traffic else None,
                     "algorithmic": f"{FLOP_PER_STEP} flop/RK4 step x {total_steps} steps (SURVEY.md 8d)",
                     "peak_source": "scalar FFMA microbenchmark measured in this run (bhr_measure_fp32_peak); "
                                    "MEASURED_PEAKS.json holds no FP32 figure"},
        "clocks": clocks,
        "orbit_video": orbit,
    }
    if other:
        line["other_configs"] = other
    if not args.no_cpu_baseline and world == 1:      # CPU baseline: rank 0 at N = 1 only
        sys.path.insert(0, os.path.join(ROOT, "tests"))
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
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
