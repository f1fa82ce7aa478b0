"""cProfile of the host side of the orbit-video frame loop (development aid)."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution, run_video_frames
from black_hole_renderer_b200.skybox import generate_skybox
W, H, POV, FOV = 1920, 1080, [6.0, 0.0, 0.5], 90.0
n_phi, n_r = compute_disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
r = Renderer(W, H, generate_skybox(2048, 1024, seed=42, n_stars=6000).astype(np.float32), np.zeros((n_r, n_phi, 4), np.float32))
r.set_option("stage_timing", 0)
run_video_frames(r, 120, FOV, POV, True, 360.0, 0.1)
pr = cProfile.Profile(); pr.enable()
t = {}
run_video_frames(r, 600, FOV, POV, True, 360.0, 0.1, timing=t)
pr.disable()
print(t)
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
