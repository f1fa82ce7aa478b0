"""A/B of the packed background kernel against the scalar one (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from black_hole_renderer_b200 import Renderer
n_r, n_phi = int(sys.argv[1]), int(sys.argv[2])
r = Renderer(16, 8, np.zeros((8, 16, 3), np.float32), np.zeros((n_r, n_phi, 4), np.float32))
r.init_background_layer(n_r, n_phi, seed=42)
import time
rng = np.random.default_rng(1)
c = rng.uniform(-900, 900, size=(4001, 3)).astype(np.float32)
a, b = r.eval_noise(c, "simplex"), r.eval_noise(c, "simplex_packed")
d = a != b
print("packed noise: differing", int(d.sum()), "even idx", int(d[0::2].sum()), "odd idx", int(d[1::2].sum()), "max", float(np.abs(a - b).max()))
for t in (0.0, 6.0):
    r.set_option("background_scalar", 1); r.generate_background(t); want = r._comp_field.to_numpy()
    r.set_option("background_scalar", 0); r.generate_background(t); got = r._comp_field.to_numpy()
    for pl in (0, 3, 11, 12):
        d = got[pl] != want[pl]
        print("t", t, "plane", pl, "differing", int(d.sum()), "even cols", int(d[:, 0::2].sum()), "odd cols", int(d[:, 1::2].sum()),
              "max", float(np.abs(got[pl] - want[pl]).max()), "rows with diffs", int(d.any(axis=1).sum()))
for sc in (1, 0):
    r.set_option("background_scalar", sc)
    r.generate_background(1.0); r.synchronize()
    t0 = time.perf_counter()
    for i in range(20): r.generate_background(1.0 + i)
    r.synchronize()
    print("scalar" if sc else "packed", (time.perf_counter() - t0) / 20 * 1e3, "ms")
