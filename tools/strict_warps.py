"""Ray-march time of the full frame and of the photon-ring band against the number of strict warps per block."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer, _lib as L
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
res = sys.argv[1] if len(sys.argv) > 1 else "fhd"
kw = dict(anti_alias="lod_radius", disk_tilt=20.0) if len(sys.argv) > 2 and sys.argv[2] == "aa" else {}
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi), **kw)
for _ in range(3): r.render_device(pov, fov)
r.synchronize()
cam = r._camera(pov, fov, 0)
s = H / 1080.0
bands = [(0, H), (int(278 * s), int(802 * s)), (int(270 * s), int(405 * s)), (int(405 * s), int(540 * s))]
for sw in (32, 20, 16, 12, 8, 6, 4, 2):
    r.set_option("strict_warps", sw)
    line = f"{res} {'aa ' if kw else ''}strict_warps={sw:2d}:"
    for (a, b) in bands:
        best = 1e9
        for _ in range(6):
            L.check(r._ctx, r._lib.bhr_render_rows_stage1(r._ctx, C.byref(cam), 0, a, b))
            L.check(r._ctx, r._lib.bhr_render_rows_stage2(r._ctx, 0, a, b, None))
            best = min(best, r.last_stage_ms()["ray_march"])
        line += f"  [{a},{b}) {best*1e3:7.1f} us"
    print(line, flush=True)
