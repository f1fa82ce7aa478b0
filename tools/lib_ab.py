"""A/B of two BUILDS of libbhr.so on the ray-march stage time: python tools/lib_ab.py libA.so libB.so [...]
Each library runs in its own process (BHR_LIB); configurations: fhd default, fhd fine step, 4K AA + tilt + flare.
Also writes the 8-bit fhd frame of every library and reports whether they are identical."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from util import RESOLUTIONS, synthetic_disk_texture, synthetic_skybox
    from black_hole_renderer_b200 import Renderer
    from black_hole_renderer_b200.driver import compute_disk_texture_resolution
    pov, fov = [6, 0, 0.5], 90
    for name, res, kw in (("fhd", "fhd", {}), ("fhd_fine", "fhd", dict(step_size=0.02, r_max=30.0)),
                          ("4k_aa", "4k", dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True))):
        W, H = RESOLUTIONS[res]
        n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
        r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi), **kw)
        for _ in range(3):
            r.render_device(pov, fov)
        best = None
        for _ in range(8 if res == "fhd" and not kw else 4):
            r.render_device(pov, fov); r.synchronize()
            t = r.last_stage_ms()
            best = t if best is None or t["ray_march"] < best["ray_march"] else best
        img = r.render_u8(pov, fov)
        print(f"{name}: ray_march {best['ray_march'] * 1e3:.1f} us  total {best['total'] * 1e3:.1f} us  steps {r.last_total_steps()}  "
              f"frame sha {hashlib.sha1(img.tobytes()).hexdigest()[:12]}", flush=True)
        r.close()
    sys.exit(0)

for rep in range(int(os.environ.get("PASSES", "1"))):
    for lib in sys.argv[1:]:
        env = dict(os.environ, BHR_LIB=os.path.abspath(lib))
        print(f"== {lib} (pass {rep})", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=True)
