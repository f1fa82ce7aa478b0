"""Host-side glue against fixtures produced by the reference's own host code (host.npz,
texture_pipeline.npz): camera, texture sizing, skybox, lifecycle RNG streams, CLI."""
import hashlib
import os

import numpy as np
import pytest

from util import GOLDEN


def test_build_camera():
    from black_hole_renderer_b200 import build_camera
    h = np.load(os.path.join(GOLDEN, "host.npz"))
    for row in h["cameras"]:
        pov, fov, w, hh = row[0:3], row[3], int(row[4]), int(row[5])
        got = build_camera(pov, fov, w, hh)
        flat = np.concatenate([got[0], got[1], got[2], got[3], [got[4], got[5]]])
        assert np.array_equal(flat, row[6:])


def test_scalar_camera_uploads_the_same_float32_values():
    """Renderer._camera forms the per-frame camera with build_camera_scalar (no numpy calls on the
    frame's latency path); what reaches the device -- the float32 casts of render.py:3880-3892 --
    is bit-identical to build_camera's on the reference's golden cameras, the 3600 orbit cameras,
    random cameras and the on-axis fallback."""
    from black_hole_renderer_b200 import build_camera
    from black_hole_renderer_b200.camera import build_camera_scalar
    from black_hole_renderer_b200.driver import orbit_camera
    h = np.load(os.path.join(GOLDEN, "host.npz"))
    cases = [(list(row[0:3]), float(row[3]), int(row[4]), int(row[5])) for row in h["cameras"]]
    cases += [(orbit_camera([6.0, 0.0, 0.5], f, 3600, 360.0), 90.0, 1920, 1080) for f in range(0, 3600, 7)]
    cases += [([0, 0, 8], 60.0, 160, 90), ([0, 0, -3], 45.0, 640, 360), ([1e-9, 0, 5], 90.0, 640, 360)]
    rng = np.random.default_rng(7)
    for _ in range(3000):
        cases.append((list(rng.normal(size=3) * rng.uniform(0.5, 20)), float(rng.uniform(10, 150)), 1283, 727))
    bits = lambda v: np.asarray(v, dtype=np.float64).astype(np.float32).view(np.uint32)
    for pov, fov, w, hh in cases:
        a = build_camera(np.array(pov, dtype=np.float64), fov, w, hh)
        b = build_camera_scalar(pov, fov, w, hh)
        for k in range(4):
            assert np.array_equal(bits(a[k]), bits(b[k])), (pov, fov, k)
        assert np.float32(a[4]) == np.float32(b[4]) and np.float32(a[5]) == np.float32(b[5])
        assert np.float32(max(10.0, float(np.linalg.norm(a[0])) * 2)) == np.float32(max(10.0, b[6] * 2))


def test_disk_texture_resolution():
    from black_hole_renderer_b200.driver import compute_disk_texture_resolution
    h = np.load(os.path.join(GOLDEN, "host.npz"))
    for row in h["tex_resolution"]:
        w, hh, pov, fov, ri, ro = int(row[0]), int(row[1]), list(row[2:5]), row[5], row[6], row[7]
        assert compute_disk_texture_resolution(w, hh, pov, fov, ri, ro) == (int(row[8]), int(row[9]))
    # the sizes SURVEY.md section 8 quotes
    assert compute_disk_texture_resolution(1920, 1080, [6, 0, 0.5], 90, 2.0, 15.0) == (2912, 416)


def test_edge_alpha():
    from black_hole_renderer_b200 import compute_edge_alpha
    h = np.load(os.path.join(GOLDEN, "host.npz"))
    assert np.array_equal(compute_edge_alpha(37), h["edge_alpha_37"])
    assert np.array_equal(compute_edge_alpha(416), h["edge_alpha_416"])


def test_skybox_bit_exact():
    from black_hole_renderer_b200.skybox import generate_skybox
    h = np.load(os.path.join(GOLDEN, "host.npz"))
    small = generate_skybox(256, 128, seed=42, n_stars=200)
    assert small.dtype == np.float32 and np.array_equal(small, h["skybox_256x128_s42_n200"])
    full = generate_skybox(2048, 1024, seed=42, n_stars=6000)
    assert hashlib.md5(full.astype(np.float32).tobytes()).hexdigest() == str(h["skybox_full_md5"][0])


def _dump(f):
    rows = []
    for e in f.entities:
        rows.append([e.birth_time, e.lifetime, e.omega, e.fade_in, e.fade_out, e.source_phi,
                     e.alpha_shear, e.tau_cool, e.blob_base_r, e.blob_sigma_r, e.blob_sigma_phi0,
                     e.blob_peak_density, e.blob_peak_temp, float(len(e.row_indices)),
                     float(e.row_indices[0]), float(np.sum(e.phi_density, dtype=np.float64)),
                     float(np.sum(e.phi_temp, dtype=np.float64)),
                     float(np.sum(e.fade_noise, dtype=np.float64))])
    return np.array(rows, dtype=np.float64).reshape(len(rows), 18)


def test_lifecycle_rng_streams_match_reference():
    """900 frames of spawn / cull: every entity parameter equals the reference's, i.e. the three
    PCG64 streams are consumed in the same order."""
    from black_hole_renderer_b200 import lifecycle as LC
    t = np.load(os.path.join(GOLDEN, "texture_pipeline.npz"))
    F = LC.make_factories(2.0, 15.0, 32, 128, 42)
    for k in F:
        assert np.array_equal(_dump(F[k]), t["init_factory_" + k]), k
    for f in F.values():
        f.tick(now=0.0, dt=0.0)
    for frame in range(1, 900):
        for f in F.values():
            f.tick(now=frame * 0.1, dt=0.1)
        if frame in (7, 25, 899):
            for k in F:
                assert np.array_equal(_dump(F[k]), t[f"f{frame}_factory_{k}"]), (frame, k)


def test_entity_lifecycle_properties():
    """Ported from the reference's tests/unit/test_entity_lifecycle.py: fade envelope, death."""
    from black_hole_renderer_b200 import lifecycle as LC
    F = LC.make_factories(2.0, 15.0, 64, 256, 7)
    hs = F["hotspot"].entities[0]
    b = hs.birth_time
    assert hs.fade_factor(b - 1.0) == 0.0
    assert abs(hs.fade_factor(b + hs.fade_in / 2) - 0.5) < 1e-9
    assert hs.fade_factor(b + hs.fade_in + hs.lifetime / 2) == 1.0
    assert hs.fade_factor(b + hs.total_duration + 0.1) == 0.0
    assert hs.is_dead(b + hs.total_duration) and not hs.is_dead(b + hs.total_duration - 1e-6)
    fl = F["filament"].entities[0]
    assert fl.density_factor(0.0) == 1.0
    assert fl.density_factor(10.0) > fl.density_factor(20.0) > 0
    assert fl.is_dead(fl.birth_time + LC.FILAMENT_MAX_LIFETIME)
    # steady state: the population stays at the target
    for frame in range(600):
        for f in F.values():
            f.tick(now=frame * 0.1, dt=0.1)
    assert 150 <= len(F["filament"].entities) <= 200
    assert len(F["hotspot"].entities) <= 30 and len(F["rt_spike"].entities) <= 15


def test_cli_surface():
    import render
    a = render.parse_args([])
    assert (a.pov, a.fov, a.resolution, a.step_size, a.r_max, a.n_stars) == ([6, 0, 0.5], 90, "fhd", 0.1, 10, 6000)
    assert (a.disk_inner_radius, a.disk_outer_radius, a.disk_tilt, a.anti_alias, a.aa_strength) == (2.0, 15.0, 0.0, "disabled", 1.0)
    assert (a.n_frames, a.fps, a.orbit_degrees, a.disk_rotation_speed) == (3600, 36, 360.0, 0.1)
    a = render.parse_args("--pov 4 3 2 --fov 75 -r 4k --ar1 1.5 --ar2 9 --disk_tilt 20 -s 0.02 --r_max 30 "
                          "--anti_alias lod_radius --lens_flare --video --orbit --resume "
                          "--disk_generation_scale 4 --force_regenerate_disk_texture "
                          "--disk_rotation_algorithm keyframes --keyframes_count 3 -d gpu".split())
    render.validate_args(a)
    assert a.video and a.orbit and a.resume and a.lens_flare and a.disk_outer_radius == 9
    for bad in (["--fov", "180"], ["--ar1", "5", "--ar2", "3"], ["-s", "0"], ["--aa_strength", "3"],
                ["--n_frames", "0"], ["--fps", "0"], ["--orbit_degrees", "inf"],
                ["--disk_texture", "x.png", "--video"]):
        with pytest.raises(ValueError):
            render.validate_args(render.parse_args(bad))


def test_numpy_quantile_helpers_match_numpy_bit_for_bit():
    """The device statistics return order statistics; the host applies numpy's interpolation to
    them with these helpers, which must reproduce np.percentile / np.quantile exactly (float32
    virtual index and all)."""
    from black_hole_renderer_b200.renderer import numpy_linear_lerp, numpy_quantile_neighbours
    rng = np.random.default_rng(1)
    for n in [1, 2, 3, 7, 100, 2912, 5824, 416 * 2912]:
        a = rng.random(n).astype(np.float32) ** 3
        srt = np.sort(a)
        for q in [98, 95, 50, 0, 100]:
            lo, hi, g = numpy_quantile_neighbours(n, np.true_divide(q, np.float32(100)))
            got, want = numpy_linear_lerp(srt[lo], srt[hi], g), np.percentile(a, q)
            assert got == want and got.dtype == want.dtype, (n, q)
        lo, hi, g = numpy_quantile_neighbours(n, np.asanyarray(0.7, dtype=np.float32))
        assert numpy_linear_lerp(srt[lo], srt[hi], g) == np.quantile(a, 0.7)
    rows = rng.random((50, 2912)).astype(np.float32)
    lo, hi, g = numpy_quantile_neighbours(2912, np.asanyarray(0.7, dtype=np.float32))
    srt = np.sort(rows, axis=1)
    assert np.array_equal(numpy_linear_lerp(srt[:, lo], srt[:, hi], g), np.quantile(rows, 0.7, axis=1))
