"""Synchronous Renderer.render() into pinned host memory with the frame finished in row bands
(api.cu: render_sync_banded): ms/frame for sync_bands = 0 (one shot), 1, 2, 3, 4 and a bit-exactness
check against the one-shot frame.  Usage: python tools/sync_bands.py [fhd|4k] [n_frames]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
res = sys.argv[1] if len(sys.argv) > 1 else "fhd"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi))
out = r.pinned_frame(np.float32)
r.set_option("sync_bands", 0)
ref = r.render(pov, fov, out=out).copy()
steps0 = r.last_total_steps()
for dtype in (np.float32, np.uint8):
    buf = r.pinned_frame(dtype)
    call = r.render if dtype == np.float32 else r.render_u8
    for bands in (0, 1, 2, 3, 4, 6):
        r.set_option("sync_bands", bands)
        for _ in range(3): call(pov, fov, out=buf)
        t0 = time.perf_counter()
        for _ in range(n): call(pov, fov, out=buf)
        ms = (time.perf_counter() - t0) * 1e3 / n
        if dtype == np.float32:
            same = np.array_equal(buf, ref) and r.last_total_steps() == steps0
        else:
            same = np.array_equal(buf, (np.clip(ref, 0, 1) * np.float32(255)).astype(np.uint8))
        print(f"{res} {np.dtype(dtype).name} sync_bands={bands}: {ms:.4f} ms/frame  identical={same}", flush=True)
