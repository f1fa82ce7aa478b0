"""Two rays per thread (option raymarch_pair) against one ray per thread at fhd: ray-march time, pixels, parity vs the oracle.
Measurement tool (uses the tests' oracle frame)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import test_parity_gpu as T
c = T._fhd_default_case()
r, ref = c["r"], c["ref"]
ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
base = None
for pair in (0, 512, 448, 384, 0):
    r.set_option("raymarch_pair", pair)
    for _ in range(3):
        r.render_device(c["pov"], c["fov"])
    r.synchronize()
    ms = []
    for _ in range(10):
        r.render_device(c["pov"], c["fov"]); r.synchronize(); ms.append(r.last_stage_ms()["ray_march"])
    img = r.render(c["pov"], c["fov"], aux=True)
    cls, steps = r.last_aux()
    u8 = r.render_u8(c["pov"], c["fov"]).copy()
    if base is None:
        base = u8
    rep = T.parity_report(img, ref["final"], cls & 31, ref_cls)
    escaped = ref["term"] == 2
    boundary = ((cls & 31) != ref_cls) | (escaped & (steps != ref["steps"]))
    d = np.abs(u8.astype(int) - base.astype(int)).max(axis=-1)
    print(f"pair {pair}: ray march {np.median(ms):.4f} ms (min {min(ms):.4f}), steps {r.last_total_steps()}, vs one-ray frame: "
          f"{int((d > 0).sum())} px differ (max {int(d.max())}); vs oracle: class flips {rep['class_flips']}, boundary px {int(boundary.sum())}, "
          f"px>2 {rep['n_gt2']}, max u8 {rep['max_u8']}, psnr {rep['psnr']:.2f}", flush=True)
