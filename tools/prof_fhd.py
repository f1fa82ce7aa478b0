"""Profiling target: a few fhd frames of the default scene (synthetic textures), one raymarch mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
res = sys.argv[1] if len(sys.argv) > 1 else "fhd"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
retrace = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kw = {}
if len(sys.argv) > 4 and sys.argv[4] == "aa":
    kw = dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True)
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi), **kw)
r.set_option("raymarch_mode", mode); r.set_option("retrace_min_cross", retrace)
band = float(os.environ.get("BAND", "0.02")) if retrace else 0.0
r.set_option("retrace_band", band)
r.set_option("persistent", int(os.environ.get("PERSISTENT", "1")))
r.set_option("pblock_big", int(os.environ.get("BIG", "0")))
for _ in range(4): r.render_device(pov, fov)
r.synchronize()
print(res, "mode", mode, "retrace", retrace, kw, r.last_stage_ms(), "steps", r.last_total_steps(), "retraced", r.last_retrace_count())
