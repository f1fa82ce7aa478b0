"""Device time of the ray march / bloom H of single row bands (stage calls through the C-ABI)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer, _lib as L
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
res = sys.argv[1] if len(sys.argv) > 1 else "fhd"
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi))
for _ in range(3): r.render_device(pov, fov)
r.synchronize()
cam = r._camera(pov, fov, 0)
s = H / 1080.0
bands = [(0, H), (int(299 * s), int(781 * s)), (0, int(299 * s)), (int(781 * s), H), (0, int(150 * s)), (int(150 * s), int(299 * s)),
         (int(400 * s), int(680 * s)), (int(299 * s), int(540 * s)), (0, int(75 * s)), (0, int(30 * s)), (0, 8)]
for (a, b) in bands:
    best = None
    for _ in range(5):
        L.check(r._ctx, r._lib.bhr_render_rows_stage1(r._ctx, C.byref(cam), 0, a, b))
        L.check(r._ctx, r._lib.bhr_render_rows_stage2(r._ctx, 0, a, b, None))
        ms = r.last_stage_ms()
        if best is None or ms["ray_march"] < best["ray_march"]: best = ms
    print(f"rows [{a:4d}, {b:4d}) {b - a:5d} rows: ray march {best['ray_march']*1e3:7.1f} us  bloom_h {best['bloom_h']*1e3:6.1f} us  "
          f"bloom_v+composite {best['bloom_v_composite']*1e3:6.1f} us   ({best['ray_march']*1e3/(b-a):.3f} us/row)", flush=True)
