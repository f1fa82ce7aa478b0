"""Profiling target: the device PNG encoder on an fhd frame of the default scene; prints stream size and kernel time."""
import sys, os, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
from black_hole_renderer_b200 import Renderer, png_codec
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
res = sys.argv[1] if len(sys.argv) > 1 else "fhd"
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi))
u8 = r.render_u8(pov, fov).copy()
for _ in range(3):
    png = r.encode_png_current()
t0 = time.perf_counter()
for _ in range(20):
    png = r.encode_png_current()
ms = (time.perf_counter() - t0) / 20 * 1e3
filt = png_codec.sub_filter(u8).tobytes()
t0 = time.perf_counter(); z1 = zlib.compress(filt, 1); z1_ms = (time.perf_counter() - t0) * 1e3
print(f"{res}: raw {u8.size} B, device PNG {len(png)} B ({u8.size / len(png):.2f}x), sync encode+copy {ms:.3f} ms; "
      f"host zlib-1 of the same rows {len(z1)} B in {z1_ms:.1f} ms")
# pipelined: raw u8 frames vs PNG streams, static scene, 200 frames, 8 slots
import time as _t
def loop(png):
    bufs = [r.pinned_bytes(8 + r.png_stream_capacity()) if png else r.pinned_frame(np.uint8) for _ in range(8)]
    host = 0.0
    t0 = _t.perf_counter()
    for i in range(208):
        s = i % 8
        if i >= 8:
            r.wait_frame(s)
        h0 = _t.perf_counter()
        if png:
            r.render_png_async(pov, fov, bufs[s], s, copy_bytes=2_600_000)
        else:
            r.render_u8_async(pov, fov, bufs[s], s)
        host += _t.perf_counter() - h0
    for s in range(8):
        r.wait_frame(s)
    return (_t.perf_counter() - t0) / 208 * 1e3, host / 208 * 1e3
r.set_option("stage_timing", 0)
for png in (False, True, False, True):
    ms, host = loop(png)
    print(f"pipelined {'png' if png else 'u8 '}: {ms:.4f} ms/frame, host enqueue {host:.4f} ms/frame")
