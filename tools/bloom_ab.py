"""A/B of the TMA bloom kernels against the generic ones on one size (debug aid).
    python tools/bloom_ab.py W H [flare]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import synthetic_disk_texture, synthetic_skybox
from black_hole_renderer_b200 import Renderer
W, H = int(sys.argv[1]), int(sys.argv[2])
flare = len(sys.argv) > 3
r = Renderer(W, H, synthetic_skybox(256, 512), synthetic_disk_texture(144, 976), lens_flare=flare)
out = {}
for g in (3, 2, 1, 0):
    r.set_option("bloom_generic", g)
    try:
        out[g] = r.render([6, 0, 0.5], 90).copy()
    except Exception as e:
        print({3: "generic", 2: "tma H only", 1: "tma V only", 0: "tma"}[g], "FAILED", str(e)[-60:], flush=True)
        break
    print({3: "generic", 2: "tma H only", 1: "tma V only", 0: "tma"}[g], "ok", out[g].mean(), "equal:", np.array_equal(out[g], out[3]), flush=True)
