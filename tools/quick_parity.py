"""Ad-hoc GPU parity / timing probe (development aid; the real checks live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
import oracle as O
from black_hole_renderer_b200 import Renderer

def golden_case(name, mode):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = d["params"]; w, h = int(p[0]), int(p[1])
    aa = "lod_radius" if p[12] else "disabled"
    r = Renderer(w, h, d["skybox"], d["disk_tex"], step_size=p[6], r_max=p[7], r_disk_inner=p[8], r_disk_outer=p[9],
                 disk_tilt=p[10], lens_flare=bool(p[11]), anti_alias=aa, aa_strength=p[13])
    r.set_option("raymarch_mode", mode)
    img = r.render(list(p[2:5]), p[5])
    print(f"  {name} mode={mode}: final max|d|={np.abs(img - d['final']).max():.3g} "
          f"bg {np.abs(r.image_field.to_numpy().transpose(1,0,2) - d['bg']).max():.3g} "
          f"disk {np.abs(r._planar(1).transpose(1,2,0) - d['disk_layer']).max():.3g} "
          f"blur {np.abs(r.blur_field.to_numpy().transpose(1,0,2) - d['blur']).max():.3g}")
    img2 = r.render(list(p[2:5]), p[5], skip_bloom=True)
    print(f"      skip_bloom max|d|={np.abs(img2 - d['final_skip_bloom']).max():.3g}")

def scene(res, mode, **kw):
    W, H = RESOLUTIONS[res]
    pov, fov = [6, 0, 0.5], 90
    n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
    sky = synthetic_skybox(); tex = synthetic_disk_texture(n_r, n_phi)
    r = Renderer(W, H, sky, tex, **kw)
    r.set_option("raymarch_mode", mode)
    img = r.render(pov, fov, aux=True)
    cls, steps = r.last_aux()
    for _ in range(3): r.render_device(pov, fov)
    r.synchronize()
    ms = r.last_stage_ms(); tot = r.last_total_steps()
    print(f"  {res} mode={mode} {kw}: stage ms {ms}  total steps {tot} mean {tot/(W*H):.2f}")
    t0 = time.time()
    ref = O.render(W, H, pov, fov, sky, tex, lens_flare_on=kw.get('lens_flare', False),
                   **{k: v for k, v in kw.items() if k != 'lens_flare'})
    print(f"      oracle {time.time()-t0:.1f}s steps {ref['total_steps']}")
    rep = parity_report(img, ref["final"], cls & 7, ref["term"] | (4 * (ref["nhits"] > 0)))
    print("      parity", rep, "steps equal frac", float((steps == ref["steps"]).mean()))

if __name__ == "__main__":
    for mode in (2, 0):
        for n in ["raymarch_default", "raymarch_aa_tilt_flare", "raymarch_e2e_like", "raymarch_offaxis_fine"]:
            golden_case(n, mode)
    for mode in (2, 0):
        scene("sd", mode)
    scene("sd", 0, anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True)
    for mode in (0,):
        scene("fhd", mode)
