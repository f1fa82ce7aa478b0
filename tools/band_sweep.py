"""Width of the strict-integrator band (|b / b_c - 1| < retrace_band) against parity and ray-march time at fhd.
Uses the tests' oracle frame (this is a measurement tool, not product code)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import test_parity_gpu as T
c = T._fhd_default_case()
r, ref = c["r"], c["ref"]
ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
for band in (0.02, 0.015, 0.01, 0.0075, 0.005, 0.0025):
    r.set_option("retrace_band", band)
    for _ in range(3):
        r.render_device(c["pov"], c["fov"])
    r.synchronize()
    ms = []
    for _ in range(10):
        r.render_device(c["pov"], c["fov"]); r.synchronize(); ms.append(r.last_stage_ms()["ray_march"])
    img = r.render(c["pov"], c["fov"], aux=True)
    cls, steps = r.last_aux()
    rep = T.parity_report(img, ref["final"], cls & 31, ref_cls)
    escaped = ref["term"] == 2
    boundary = ((cls & 31) != ref_cls) | (escaped & (steps != ref["steps"]))
    g8 = (np.clip(img, 0, 1) * np.float32(255)).astype(np.uint8).astype(np.int32)
    r8 = (np.clip(ref["final"], 0, 1) * np.float32(255)).astype(np.uint8).astype(np.int32)
    d = np.abs(g8 - r8).max(axis=-1)
    print(f"band {band}: ray march {np.median(ms):.4f} ms, retraced {r.last_retrace_count()}, class flips {rep['class_flips']}, "
          f"boundary px {int(boundary.sum())}, px>2 {rep['n_gt2']}, px>1 non-boundary {int((d[~boundary] > 1).sum())}, "
          f"max u8 {rep['max_u8']}, psnr {rep['psnr']:.2f}", flush=True)
