"""Test-infrastructure script (not imported by the product): derives the static Huffman code lengths of the device PNG
encoder (black_hole_renderer_b200/png_codec.py: _FITTED_LENGTHS) from frames rendered by the CPU oracle, and reports
how the code does on held-out frames.

    python oracle/fit_png_code.py          # ~3 min on 8 cores: four orbit frames with the lifecycle texture + two test frames

Fit set: orbit-video frames 0 and 1350 (lifecycle disk texture, `render.py --video --orbit -r fhd`) and the default frame
with the tests' synthetic texture.  Held out: orbit frames 450 and 2400, the synthetic texture from another camera angle.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as O                                                         # noqa: E402
from util import RESOLUTIONS, synthetic_disk_texture, synthetic_skybox     # noqa: E402
from black_hole_renderer_b200 import png_codec as pc                       # noqa: E402
from black_hole_renderer_b200.driver import orbit_camera                   # noqa: E402
from black_hole_renderer_b200.lifecycle import make_factories              # noqa: E402

W, H = RESOLUTIONS["fhd"]
POV, FOV, DT, N_TOTAL = [6.0, 0.0, 0.5], 90.0, 0.1, 3600


def to_u8(ref):
    return (np.clip(ref["final"], 0, 1) * np.float32(255)).astype(np.uint8)


def orbit_frames(wanted):
    """The oracle's run of the video lifecycle (tests/test_parity_gpu.py::test_config5_*), frames `wanted`."""
    n_phi, n_r = O.disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
    sky = synthetic_skybox()
    rng = np.random.default_rng(42)
    az_freq, az_shear = int(rng.integers(2, 5)), float(rng.uniform(2.0, 4.0))
    F = make_factories(2.0, 15.0, n_r, n_phi, seed=42)
    edge, omega = O.edge_alpha(n_r), O.omega_rows(n_r, 2.0, 15.0)
    comp = np.zeros((13, n_r, n_phi), dtype=np.float32)
    out, stats, rows = {}, None, None
    for frame in range(max(wanted) + 1):
        t = frame * DT
        for f in F.values():
            f.tick(now=t, dt=DT)
        if frame % 60 == 0 or frame in wanted:
            O.generate_background(comp, az_freq, az_shear, 2.0, 15.0, t)
            comp[5:11] = O.accumulate_entities(F, t, n_r, n_phi, omega)
            if frame % 60 == 0:
                stats, rows = O.interactive_stats(comp, edge)
        if frame in wanted:
            tex = O.compose_texture(comp, omega, edge, stats, rows)
            cam = orbit_camera(POV, frame, N_TOTAL, 360.0)
            out[frame] = to_u8(O.render(W, H, cam, FOV, sky, tex, mips=O.build_mips(tex, 5, numpy_order=False)))
    return out


def token_histograms(u8):
    """Literal counts per byte value and match counts per length of the encoder's tokens (runs inside 256-byte segments)."""
    filt = pc.sub_filter(u8)
    n = len(filt)
    same = np.zeros(n, bool)
    same[1:] = filt[1:] == filt[:-1]
    same[(np.arange(n) % pc.SEGMENT) == 0] = False
    starts = np.flatnonzero(~same)
    lens = np.diff(np.append(starts, n))
    lit, mlen, r = np.zeros(256), np.zeros(pc.MAX_MATCH + 1), lens - 1
    np.add.at(lit, filt[starts], 1 + np.where(r < pc.MIN_RUN, r, 0))
    np.add.at(mlen, r[r >= pc.MIN_RUN], 1)
    return lit, mlen, filt


def length_symbol(length):
    return 28 if length == 258 else max(i for i, b in enumerate(pc._LEN_BASE) if b <= length)


def symbol_freqs(lit, mlen):
    f = np.zeros(286)
    f[:256] = lit
    f[256] = 1
    for length in np.flatnonzero(mlen):
        f[257 + length_symbol(length)] += mlen[length]
    return f


def stream_bytes(code_len, lit, mlen):
    bits = (lit * code_len[:256]).sum()
    for length in np.flatnonzero(mlen):
        k = length_symbol(length)
        bits += mlen[length] * (code_len[257 + k] + pc._LEN_EXTRA[k] + 1)
    return int(bits / 8)


if __name__ == "__main__":
    n_phi, n_r = O.disk_texture_resolution(W, H, POV, FOV, 2.0, 15.0)
    sky, tex = synthetic_skybox(), synthetic_disk_texture(n_r, n_phi)
    frames = {f"orbit {k}": v for k, v in orbit_frames((0, 450, 1350, 2400)).items()}
    for f in (0, 900):
        frames[f"test texture, orbit angle of frame {f}"] = to_u8(O.render(W, H, orbit_camera(POV, f, N_TOTAL, 360.0), FOV, sky, tex))
    S = {k: token_histograms(v) for k, v in frames.items()}
    fit_set = ("orbit 0", "orbit 1350", "test texture, orbit angle of frame 0")
    freq = sum(symbol_freqs(*S[k][:2]) for k in fit_set)
    freq = freq + freq.sum() * 2e-5                       # floor: every symbol gets a code
    lengths = pc._huffman_lengths(freq, 13)
    cur = pc.static_code().lit_len
    for k, (lit, mlen, filt) in S.items():
        own = pc._huffman_lengths(symbol_freqs(lit, mlen) + 0.5, 13)
        print(f"{k:45s} {'(fit set)' if k in fit_set else '(held out)':10s} fitted {stream_bytes(lengths, lit, mlen):8d}  in the tree now "
              f"{stream_bytes(cur, lit, mlen):8d}  own fit {stream_bytes(own, lit, mlen):8d}  zlib-1 {len(zlib.compress(filt.tobytes(), 1)):8d}")
    print("lengths:", [int(x) for x in lengths])
    print("identical to png_codec._FITTED_LENGTHS:", [int(x) for x in lengths] == list(pc._FITTED_LENGTHS))
