"""Which rays are ill-conditioned?  Compare the fast integrator (no retrace) with the oracle and
print the impact parameter band eps = b / b_crit - 1 of the pixels whose 8-bit value differs by > 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
import oracle as O
from black_hole_renderer_b200 import Renderer
BC = 1.5 * np.sqrt(3.0)
def run(res, pov, fov, **kw):
    W, H = RESOLUTIONS[res] if isinstance(res, str) else res
    n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, kw.get("r_disk_inner", 2.0), kw.get("r_disk_outer", 15.0))
    sky = synthetic_skybox(); tex = synthetic_disk_texture(n_r, n_phi)
    r = Renderer(W, H, sky, tex, **kw); r.set_option("raymarch_mode", 0); r.set_option("retrace_min_cross", int(os.environ.get("MINCROSS", "0"))); r.set_option("retrace_band", float(os.environ.get("BAND", "0.02"))); r.set_option("band_lo_auto", int(os.environ.get("LOAUTO", "1")))
    img = r.render(pov, fov, aux=True, skip_bloom=True); cls, steps = r.last_aux()
    okw = dict(step_size=kw.get("step_size", 0.1), r_max=kw.get("r_max", 10.0), r_inner=kw.get("r_disk_inner", 2.0),
               r_outer=kw.get("r_disk_outer", 15.0), disk_tilt=kw.get("disk_tilt", 0.0))
    ref = O.render(W, H, pov, fov, sky, tex, skip_bloom=True, **okw)
    g8 = (np.clip(img,0,1)*np.float32(255)).astype(np.uint8).astype(int); r8 = (np.clip(ref['final'],0,1)*np.float32(255)).astype(np.uint8).astype(int)
    d = np.abs(g8-r8).max(-1)
    p, right, up, fwd, pw, ph = O.build_camera(pov, fov, W, H)
    xs = (np.arange(W) + 0.5 - W/2) * pw; ys = -(np.arange(H) + 0.5 - H/2) * ph
    dirs = fwd[None,None,:] + xs[None,:,None]*right[None,None,:] + ys[:,None,None]*up[None,None,:]
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    b = np.linalg.norm(np.cross(dirs, p[None,None,:]), axis=-1)
    rc = np.linalg.norm(p)
    eps = b/np.sqrt(np.maximum(1 - b*b/rc**3, 1e-6))/BC - 1     # impact parameter at infinity, as in the kernel
    nc = cls >> 5
    steps_diff = steps != ref['steps']
    clsd = (cls & 31) != (ref['term'] | (np.minimum(ref['nhits'],7) << 2))
    same = ~steps_diff & ~clsd
    print(f"{res} pov={pov} fov={fov} {kw}: bad(d>1)={int((d>1).sum())} bad(d>2)={int((d>2).sum())} clsdiff={int(clsd.sum())} stepdiff={int(steps_diff.sum())}  | same steps+class: d>1 {int((d>1)[same].sum())} d>2 {int((d>2)[same].sum())} retraced {r.last_retrace_count()}")
    if os.environ.get("DETAIL"):
        bg = r.image_field.to_numpy().transpose(1,0,2); dk = r._planar(1).transpose(1,2,0)
        ys_, xs_ = np.nonzero(d > 1)
        for y, x in list(zip(ys_, xs_))[:12]:
            print(f"     px ({x},{y}) d={d[y,x]} eps={eps[y,x]:.4f} cls gpu {cls[y,x]&31} ref {ref['term'][y,x] | (min(ref['nhits'][y,x],7)<<2)} steps gpu {steps[y,x]} ref {ref['steps'][y,x]} ncross {nc[y,x]}"
                  f"\n        gpu bg {bg[y,x]} disk {dk[y,x]}\n        ref bg {ref['bg'][y,x]} disk {ref['disk'][y,x]}")
    for name, mask in (("d>1", d>1), ("d>2", d>2), ("cls", clsd), ("steps", steps_diff)):
        if mask.any():
            e = eps[mask]; print(f"   {name}: eps range [{e.min():.4f}, {e.max():.4f}]  ncross min {nc[mask].min()}  |eps|max {np.abs(e).max():.4f}")
    for thr in (0.005, 0.01, 0.015, 0.02):
        band = np.abs(eps) < thr
        print(f"   band |eps|<{thr}: {band.mean()*100:.2f}% of pixels, covers d>1: {(d>1)[band].sum()}/{(d>1).sum()}  d>2: {(d>2)[band].sum()}/{(d>2).sum()}, nc>=3 inside: {(nc>=3)[band].sum()}/{(nc>=3).sum()}")
run("sd", [6,0,0.5], 90)
run("fhd", [6,0,0.5], 90)
run("sd", [6,0,0.5], 90, disk_tilt=20.0)
run("sd", [4,3,2], 75, disk_tilt=-35.0, r_disk_inner=1.5, r_disk_outer=9.0)
run("sd", [0,0,8], 60)
run("sd", [20,0,3], 40)
run((320,180), [6,0,0.5], 90, step_size=0.02, r_max=30.0)
run("sd", [2.5,0,0.3], 100)
