"""Host-side time breakdown of one orbit-video block (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from black_hole_renderer_b200 import Renderer
from black_hole_renderer_b200.driver import compute_disk_texture_resolution, orbit_camera
from black_hole_renderer_b200.skybox import generate_skybox
from black_hole_renderer_b200.lifecycle import init_lifecycle_system, pack_entities_array
W, H, POV, FOV = 1920, 1080, [6.0, 0.0, 0.5], 90.0
n_phi, n_r = compute_disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
r = Renderer(W, H, generate_skybox(2048, 1024, seed=42, n_stars=6000).astype(np.float32), np.zeros((n_r, n_phi, 4), np.float32))
F = init_lifecycle_system(r, n_r, n_phi, seed=42)
out = r.pinned_frame(np.uint8)
T = {}
def tic(name, fn):
    t0 = time.perf_counter(); v = fn(); r.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return v
N = 120
t_all = time.perf_counter()
for f in range(N):
    t = f * 0.1
    tic("tick", lambda: [x.tick(now=t, dt=0.1) for x in F.values()])
    tic("background", lambda: r.generate_background(t=t))
    tic("pack", lambda: pack_entities_array(F, t, n_r))
    tic("entities(pack+kernel)", lambda: r.accumulate_entity_layer(F, now=t))
    if f % 60 == 0:
        tic("stats", lambda: r.recompute_interactive_stats())
    tic("compose+mips", lambda: r.compose_interactive_texture())
    tic("render_u8", lambda: r.render_u8(orbit_camera(POV, f, 3600, 360.0), FOV, out=out))
tot = time.perf_counter() - t_all
for k, v in T.items(): print(f"{k:24s} {1e3 * v / N:8.3f} ms/frame")
print(f"total {1e3 * tot / N:.3f} ms/frame (every stage synchronised)")
