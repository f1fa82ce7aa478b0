"""Shared helpers of the test-suite: synthetic inputs and the parity metrics of the north star."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

RESOLUTIONS = {"4k": (3840, 2160), "fhd": (1920, 1080), "hd": (1280, 720), "sd": (640, 360)}


def synthetic_skybox(h=1024, w=2048, seed=0):
    """Smooth large-scale structure + sparse bright 'stars' (timing is insensitive to values)."""
    rng = np.random.default_rng(seed)
    v = np.linspace(0, np.pi, h)[:, None]
    u = np.linspace(0, 2 * np.pi, w, endpoint=False)[None, :]
    sky = np.zeros((h, w, 3), dtype=np.float32)
    for c in range(3):
        sky[..., c] = 0.08 + 0.06 * np.sin(3 * u + c) * np.sin(2 * v + 0.5 * c) ** 2
    stars = rng.random((h, w)) > 0.9985
    sky[stars] += rng.uniform(0.3, 1.0, (int(stars.sum()), 1)).astype(np.float32)
    return np.clip(sky, 0, 1).astype(np.float32)


def synthetic_disk_texture(n_r, n_phi, seed=11):
    """Smooth RGBA polar texture with a wide alpha range and soft radial edges."""
    rng = np.random.default_rng(seed)
    r = np.linspace(0, 1, n_r)[:, None]
    p = np.linspace(0, 2 * np.pi, n_phi, endpoint=False)[None, :]
    tex = np.zeros((n_r, n_phi, 4), dtype=np.float32)
    for c in range(4):
        a, b = rng.uniform(1, 4, 2)
        k = int(rng.integers(1, 9))
        ph = rng.uniform(0, 6.28)
        tex[..., c] = 0.5 + 0.5 * np.sin(a * r * 9.0 + k * p + ph) * np.cos(b * r * 5.0)
    tex += rng.uniform(0, 0.05, tex.shape).astype(np.float32)
    edge = np.minimum(np.clip(r / 0.1, 0, 1) ** 3, np.clip((1 - r) / 0.3, 0, 1) ** 2)
    tex[..., 3] *= edge
    return np.clip(tex, 0, 1).astype(np.float32)


def psnr_u8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def class_map(term, nhits):
    """north-star classes: horizon / disk-hit / sky-only (+ exhausted), from term and hit flag."""
    return term.astype(np.int32) + 4 * (nhits > 0).astype(np.int32)


def parity_report(gpu_f32, ref_f32, gpu_cls=None, ref_cls=None):
    """Metrics of BASELINE.json's parity gate on 8-bit (truncated) frames."""
    g8 = (np.clip(gpu_f32, 0, 1) * np.float32(255)).astype(np.uint8)
    r8 = (np.clip(ref_f32, 0, 1) * np.float32(255)).astype(np.uint8)
    d = np.abs(g8.astype(np.int32) - r8.astype(np.int32))
    rep = dict(max_u8=int(d.max()), n_gt2=int((d.max(axis=-1) > 2).sum()),
               frac_gt2=float((d.max(axis=-1) > 2).mean()), psnr=float(psnr_u8(g8, r8)),
               max_f32=float(np.abs(gpu_f32 - ref_f32).max()))
    if gpu_cls is not None:
        flips = gpu_cls != ref_cls
        rep["class_flips"] = int(flips.sum())
        rep["class_flip_frac"] = float(flips.mean())
        ok = ~flips
        dd = d.max(axis=-1)
        rep["max_u8_same_class"] = int(dd[ok].max()) if ok.any() else 0
    return rep
