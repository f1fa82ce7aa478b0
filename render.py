#!/usr/bin/env python3
"""Schwarzschild black-hole ray tracer -- B200 (sm_100a) build.

Same entry points and flags as the reference's render.py (render.py:4518-4694): single frame,
`--video [--orbit] [--resume]`; `--interactive` needs a display and is not provided.  With
`torchrun --nproc-per-node N render.py --video ...` the frames are sharded over N GPUs.
"""
import argparse
import math
import os

from black_hole_renderer_b200.driver import (make_renderer, render_image, render_video, save_image)
from black_hole_renderer_b200.renderer import R_DISK_INNER_DEFAULT, R_DISK_OUTER_DEFAULT

DISK_GENERATION_SCALE_CHOICES = (1, 2, 4)
RESOLUTIONS = {"4k": (3840, 2160), "fhd": (1920, 1080), "hd": (1280, 720), "sd": (640, 360)}


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Schwarzschild black-hole ray tracer (B200)")
    p.add_argument("--pov", type=float, nargs=3, default=[6, 0, 0.5], metavar=("X", "Y", "Z"))
    p.add_argument("--fov", type=float, default=90)
    p.add_argument("--resolution", "-r", type=str, default="fhd", choices=list(RESOLUTIONS))
    p.add_argument("--texture", "-t", type=str, default=None)
    p.add_argument("--output", "-o", type=str, default="output/blackhole.png")
    p.add_argument("--step_size", "-s", type=float, default=0.1)
    p.add_argument("--r_max", type=float, default=10)
    p.add_argument("--n_stars", type=int, default=6000)
    p.add_argument("--disk_texture", type=str, default=None)
    p.add_argument("--disk_generation_scale", type=int, default=2, choices=DISK_GENERATION_SCALE_CHOICES,
                   help="[deprecated, ignored]")
    p.add_argument("--force_regenerate_disk_texture", action="store_true", help="[deprecated, ignored]")
    p.add_argument("--disk_inner_radius", "--ar1", dest="disk_inner_radius", type=float,
                   default=R_DISK_INNER_DEFAULT)
    p.add_argument("--disk_outer_radius", "--ar2", dest="disk_outer_radius", type=float,
                   default=R_DISK_OUTER_DEFAULT)
    p.add_argument("--disk_tilt", type=float, default=0.0)
    p.add_argument("--lens_flare", action="store_true")
    p.add_argument("--anti_alias", type=str, default="disabled", choices=["disabled", "lod_radius"])
    p.add_argument("--aa_strength", type=float, default=1.0)
    p.add_argument("--device", "-d", type=str, default="cpu", choices=["cpu", "gpu"],
                   help="accepted for compatibility; the kernels always run on the CUDA device")
    p.add_argument("--ignore_taichi_cache", action="store_true", help="accepted, no effect")
    p.add_argument("--video", action="store_true")
    p.add_argument("--interactive", action="store_true")
    p.add_argument("--orbit", action="store_true")
    p.add_argument("--orbit_degrees", type=float, default=360.0)
    p.add_argument("--n_frames", type=int, default=3600)
    p.add_argument("--fps", type=int, default=36)
    p.add_argument("--resume", action="store_true")
    p.add_argument("--disk_rotation_algorithm", type=str, default="baseline",
                   choices=["baseline", "parametric", "keyframes"], help="[deprecated, ignored]")
    p.add_argument("--disk_rotation_speed", type=float, default=0.1)
    p.add_argument("--keyframes_count", type=int, default=10, help="[deprecated, ignored]")
    return p.parse_args(argv)


def validate_args(args):
    if not (0 < args.fov < 180):
        raise ValueError(f"FOV must be between 0 and 180 degrees, got {args.fov}")
    if args.disk_inner_radius >= args.disk_outer_radius:
        raise ValueError(f"disk_inner_radius ({args.disk_inner_radius}) must be less than "
                         f"disk_outer_radius ({args.disk_outer_radius})")
    if args.step_size <= 0:
        raise ValueError(f"step_size must be positive, got {args.step_size}")
    if not (0.5 <= args.aa_strength <= 2.0):
        raise ValueError(f"aa_strength must be between 0.5 and 2.0, got {args.aa_strength}")
    if args.n_frames <= 0:
        raise ValueError(f"n_frames must be positive, got {args.n_frames}")
    if args.fps <= 0:
        raise ValueError(f"fps must be positive, got {args.fps}")
    if not math.isfinite(args.orbit_degrees):
        raise ValueError(f"orbit_degrees must be finite, got {args.orbit_degrees}")
    if args.disk_texture and (args.video or args.interactive):
        raise ValueError("--disk_texture is only supported for single frames; video/interactive "
                         "modes use the lifecycle system")


def main(argv=None):
    args = parse_args(argv)
    validate_args(args)
    width, height = RESOLUTIONS[args.resolution]
    fov = args.fov % 180
    if args.interactive:
        raise SystemExit("--interactive needs a display (ti.GUI in the reference); not provided here")
    common = dict(step_size=args.step_size, r_max=args.r_max, device=args.device,
                  r_disk_inner=args.disk_inner_radius, r_disk_outer=args.disk_outer_radius,
                  disk_tilt=args.disk_tilt, lens_flare=args.lens_flare, anti_alias=args.anti_alias,
                  aa_strength=args.aa_strength, disk_rotation_speed=args.disk_rotation_speed)
    if args.video:
        rank, world, barrier = 0, 1, None
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            dist.init_process_group("gloo")      # control plane only: barriers, no frame data
            rank, world, barrier = dist.get_rank(), dist.get_world_size(), dist.barrier
        renderer, _ = make_renderer(width, height, args.pov, fov, args.texture, args.n_stars, **common)
        if rank == 0:
            print(f"Rendering video: {args.n_frames} frames at {width}x{height}, orbit={args.orbit}, "
                  f"{world} GPU(s)")
        render_video(renderer, width, height, n_frames=args.n_frames, fps=args.fps,
                     output_path=args.output, fov=fov, static_cam_pos=args.pov, orbit=args.orbit,
                     resume=args.resume, disk_rotation_speed=args.disk_rotation_speed,
                     orbit_degrees=args.orbit_degrees, rank=rank, world_size=world, barrier=barrier)
    else:
        img = render_image(width=width, height=height, cam_pos=args.pov, fov=fov,
                           skybox_path=args.texture, n_stars=args.n_stars,
                           disk_texture_path=args.disk_texture, **common)
        save_image(img, args.output)


if __name__ == "__main__":
    main()
