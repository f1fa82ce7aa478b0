"""Per-SM timeline of the persistent ray march (option "timeline"): how long the strict role keeps its SMs, and
how much of the launch is tail (SMs idle while the last ones finish).   python tools/tail_report.py [res] [aa]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import RESOLUTIONS, synthetic_disk_texture, synthetic_skybox
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution
res = sys.argv[1] if len(sys.argv) > 1 else "fhd"
kw = dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True) if len(sys.argv) > 2 else {}
W, H = RESOLUTIONS[res]; pov, fov = [6, 0, 0.5], 90
n_phi, n_r = compute_disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
r = Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi), **kw)
for k, v in [a.split("=") for a in os.environ.get("BHR_OPTS", "").split(",") if a]:
    r.set_option(k, float(v))
r.set_option("timeline", 1)
for _ in range(4):
    r.render_device(pov, fov)
r.synchronize()
ms = r.last_stage_ms()
buf = (C.c_uint64 * (3 * 256))()
n = r._lib.bhr_last_raymarch_timeline(r._ctx, buf, 256)
t = np.array(buf[:3 * n], dtype=np.float64).reshape(n, 3)
t0 = t[:, 0].min()
start, strict_done, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
span = end.max()
strict = (strict_done - start) > 5.0
print(f"{res} {kw or ''}: ray march {ms['ray_march'] * 1e3:.1f} us (events), persistent kernel span {span:.1f} us over {n} blocks")
print(f"  block start spread {start.max():.1f} us; blocks with a strict role: {int(strict.sum())}, strict role ends at "
      f"{strict_done[strict].mean() if strict.any() else 0:.1f} us (mean) / {strict_done[strict].max() if strict.any() else 0:.1f} us (max)")
print(f"  block end: min {end.min():.1f}  p10 {np.percentile(end, 10):.1f}  median {np.median(end):.1f}  p90 {np.percentile(end, 90):.1f}  max {end.max():.1f} us")
idle = (span - end).sum() / (n * span)
print(f"  tail: SM-time idle after a block's last tile = {100 * idle:.2f} % of the launch ({(span - end).mean():.1f} us per SM on average)")
