"""Pixels per lane of the TMA H bloom pass (option bloom_h_p): stage time and pixels at fhd and 4K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import RESOLUTIONS, synthetic_disk_texture, synthetic_skybox
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
for res in ("fhd", "4k", "hd", "sd"):
    W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
    n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
    r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi))
    base = None
    for p in (0, 10, 5, 0):
        r.set_option("bloom_h_p", p)
        for _ in range(3):
            r.render_device(pov, fov)
        r.synchronize()
        h, v = [], []
        for _ in range(12):
            r.render_device(pov, fov); r.synchronize(); st = r.last_stage_ms(); h.append(st["bloom_h"]); v.append(st["bloom_v_composite"])
        img = r.render(pov, fov)
        if base is None:
            base = img
        print(f"{res} bloom_h_p {p}: H {1e3 * np.median(h):.1f} us (min {1e3 * min(h):.1f}), V {1e3 * np.median(v):.1f} us, identical to P=default: {np.array_equal(img, base)}", flush=True)
    r.close()
