"""A/B: entity layer on its own stream beside the background kernel (option entity_stream) x resident background blocks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution, run_video_frames
from black_hole_renderer_b200.lifecycle import init_lifecycle_system
W, H = RESOLUTIONS["fhd"]; pov, fov = [6.0, 0.0, 0.5], 90.0
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), np.zeros((n_r, n_phi, 4), np.float32))
r.set_option("stage_timing", 0)
frames = {}
for rep in range(2):
    for es, early, blocks in ((0, 0, 2), (1, 0, 1), (1, 1, 1), (1, 1, 2), (1, 0, 2)):
        F = init_lifecycle_system(r, n_r, n_phi, seed=42)
        r.set_option("entity_stream", es)
        r.set_option("entity_early", early)
        r.set_option("background_blocks_per_sm", blocks)
        r.synchronize()
        keep = {}
        t0 = time.perf_counter()
        run_video_frames(r, 600, fov, pov, True, 360.0, 0.1, factories=F,
                         sink=lambda f, img: keep.__setitem__(f, img.copy()) if f in (0, 59, 60, 599) else None)
        r.synchronize()
        ms = (time.perf_counter() - t0) / 600 * 1e3
        same = all(np.array_equal(keep[f], frames.setdefault(f, keep[f])) for f in keep)
        print(f"entity_stream {es} early {early} bg blocks/SM {blocks}: {ms:.4f} ms/frame, frames identical to the first run: {same}", flush=True)
