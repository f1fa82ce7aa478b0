"""Does the entity layer (own stream, FP64 / XU heavy) overlap the background kernel (FP32 heavy)?  Times background + entity
layer + compose per frame with the entity stream off / on and 1 / 2 resident background blocks per SM."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
from black_hole_renderer_b200.lifecycle import init_lifecycle_system
W, H = RESOLUTIONS["fhd"]; pov, fov = [6.0, 0.0, 0.5], 90.0
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), np.zeros((n_r, n_phi, 4), np.float32))
F = init_lifecycle_system(r, n_r, n_phi, seed=42)
for t in np.arange(0.0, 3.0, 0.1):
    for f in F.values():
        f.tick(now=float(t), dt=0.1)
def run(parts, n=60):
    r.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        if "bg" in parts: r.generate_background(3.0)
        if "ent" in parts: r.accumulate_entity_layer(F, 3.0)
        if "comp" in parts: r.compose_interactive_texture()
    r.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for es, blocks in ((0, 2), (1, 2), (1, 1), (0, 1), (0, 2)):
    r.set_option("entity_stream", es); r.set_option("background_blocks_per_sm", blocks)
    run(("bg", "ent", "comp"), 5)
    print(f"entity_stream {es} bg blocks/SM {blocks}: bg {run(('bg',)):.4f}  ent {run(('ent',)):.4f}  comp {run(('comp',)):.4f}  "
          f"bg+ent+comp {run(('bg', 'ent', 'comp')):.4f} ms", flush=True)
