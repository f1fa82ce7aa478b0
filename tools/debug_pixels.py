import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
import oracle as O
from black_hole_renderer_b200 import Renderer
W, H = RESOLUTIONS["sd"]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
sky = synthetic_skybox(); tex = synthetic_disk_texture(n_r, n_phi)
r = Renderer(W, H, sky, tex); r.set_option("raymarch_mode", 0)
img = r.render(pov, fov, aux=True, skip_bloom=True); cls, steps = r.last_aux()
bg = r.image_field.to_numpy().transpose(1,0,2); dk = r._planar(1).transpose(1,2,0)
ref = O.render(W, H, pov, fov, sky, tex, skip_bloom=True)
g8 = (np.clip(img,0,1)*np.float32(255)).astype(np.uint8).astype(int); r8 = (np.clip(ref['final'],0,1)*np.float32(255)).astype(np.uint8).astype(int)
d = np.abs(g8-r8).max(-1)
ys, xs = np.nonzero(d > 2)
print("n bad", len(ys)); nc = cls >> 5; print("ncross hist", np.bincount(nc.ravel(), minlength=8)); cls = cls & 31
for k in range(8): print(k, "steps mean", steps[nc==k].mean() if (nc==k).any() else None, "bad", int((d[nc==k]>2).sum()), "d>0", int((d[nc==k]>0).sum()))
for y, x in zip(ys, xs):
    print(f"({x},{y}) d={d[y,x]} nc={nc[y,x]} cls gpu={cls[y,x]:#x} ref term={ref['term'][y,x]} nh={ref['nhits'][y,x]} steps {steps[y,x]} {ref['steps'][y,x]}")
    print("    bg gpu", bg[y,x], "ref", ref['bg'][y,x])
    print("    disk gpu", dk[y,x], "ref", ref['disk'][y,x])
