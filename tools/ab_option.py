"""A/B of one bhr_set_option knob on the full-frame stage times: python tools/ab_option.py key v0 v1 [res]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
key, vals = sys.argv[1], [float(v) for v in sys.argv[2:4]]
res = sys.argv[4] if len(sys.argv) > 4 else "fhd"
kw = dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True) if len(sys.argv) > 5 and sys.argv[5] == "aa" else {}
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi), **kw)
for _ in range(3): r.render_device(pov, fov)
for rep in range(3):
    for v in vals:
        r.set_option(key, v)
        ts = []
        for _ in range(8):
            r.render_device(pov, fov); r.synchronize()
            ts.append(r.last_stage_ms())
        best = min(ts, key=lambda d: d["total"])
        print(f"{key}={v:g}: ray_march {best['ray_march']*1e3:.1f} us  bloom_h {best['bloom_h']*1e3:.1f}  bloom_v+composite {best['bloom_v_composite']*1e3:.1f}  total {best['total']*1e3:.1f}", flush=True)
