"""The CPU oracle against the golden fixtures produced by the reference itself
(oracle/make_golden.py: unmodified reference kernels executed through oracle/ti_shim.py).
This is what pins the oracle; tolerances are a few f32 ulps (the shim evaluates some constant
products in double, the oracle in float like Taichi)."""
import os

import numpy as np
import pytest

import oracle as O
from util import GOLDEN

CASES = ["raymarch_default", "raymarch_aa_tilt_flare", "raymarch_e2e_like", "raymarch_offaxis_fine",
         "raymarch_frame_rot", "raymarch_frame_rot_aa"]      # the last two: render(frame != 0), rotated texture lookups
TOL = 2e-6


def _render_case(d, **over):
    p = d["params"]
    kw = dict(step_size=p[6], r_max=p[7], r_inner=p[8], r_outer=p[9], disk_tilt=p[10],
              anti_alias="lod_radius" if p[12] else "disabled", aa_strength=p[13])
    flare = bool(p[11])
    if len(p) > 14:                       # t_offset = float(frame) * disk_rotation_speed (render.py:3897)
        kw["t_offset"] = float(p[14]) * float(p[15])
    kw.update(over)
    flare = kw.pop("lens_flare_on", flare)
    return O.render(int(p[0]), int(p[1]), p[2:5], p[5], d["skybox"], d["disk_tex"], mips=d["mips"],
                    lens_flare_on=flare, **kw)


@pytest.mark.parametrize("name", CASES)
def test_render_matches_reference(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    r = _render_case(d)
    assert np.abs(r["bg"] - d["bg"]).max() <= TOL
    assert np.abs(r["disk"] - d["disk_layer"]).max() <= TOL
    assert np.abs(r["blur"] - d["blur"]).max() <= TOL
    assert np.abs(r["final"] - d["final"]).max() <= TOL
    # the in-place `disk += 0.4 * blur` tail of _bloom_kernel (dead in render(), T6)
    post = np.clip(r["disk"] + r["blur"] * np.float32(0.4), 0, 1)
    assert np.abs(post - d["disk_layer_after_bloom"]).max() <= TOL
    # most values are bit-identical
    assert (r["final"] == d["final"]).mean() > 0.97


@pytest.mark.parametrize("name", CASES)
def test_skip_flags_match_reference(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    r = _render_case(d, skip_bloom=True)
    assert np.abs(r["final"] - d["final_skip_bloom"]).max() <= TOL
    if "final_skip_diff" in d.files:
        r = _render_case(d, skip_bloom=True, skip_differentials=True, lens_flare_on=False)
        assert np.abs(r["final"] - d["final_skip_diff"]).max() <= TOL


@pytest.mark.parametrize("name", ["raymarch_default", "raymarch_aa_tilt_flare"])
def test_render_to_field_matches_reference(name):
    """render_to_field (render.py:3819-3863) run by the reference itself: final_field is
    clamp(bg + clamp(disk + 0.4 blur) + blur), y-flipped (W, H, 3), never flared; the disk layer
    field is left in its post-bloom state."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = np.load(os.path.join(GOLDEN, "render_to_field.npz"))
    p = d["params"]
    kw = dict(step_size=p[6], r_max=p[7], r_inner=p[8], r_outer=p[9], disk_tilt=p[10],
              anti_alias="lod_radius" if p[12] else "disabled", aa_strength=p[13])
    for skip, suffix in ((False, ""), (True, "_skip_bloom")):
        r = O.render_to_field(int(p[0]), int(p[1]), p[2:5], p[5], d["skybox"], d["disk_tex"], mips=d["mips"],
                              skip_bloom=skip, **kw)
        assert r["final_field"].shape == g[name + "/final_field" + suffix].shape == (int(p[0]), int(p[1]), 3)
        assert np.abs(r["final_field"] - g[name + "/final_field" + suffix]).max() <= TOL
        assert np.abs(r["disk_layer_field"] - g[name + "/disk_layer_field" + suffix]).max() <= TOL
    # it is NOT render()'s frame: the disk layer enters with 1.4 x its bloom
    assert np.abs(g[name + "/final_field"][:, ::-1].transpose(1, 0, 2) - d["final"]).max() > 1e-3


def test_shifted_compose_matches_the_reference_numpy_generator():
    """Legacy parametric path (f4): the oracle's compose with a Keplerian row shift against the
    reference's numpy _generate_disk_texture_rotating_from_state (render.py:988-1024) at the offsets
    its own test uses (tests/unit/test_gpu_texture_compose.py:98-112, tolerance 1e-4 there)."""
    g = np.load(os.path.join(GOLDEN, "shifted_compose.npz"))
    names = ["temp_base", "spiral", "spiral_temp", "turbulence", "turb_temp", "arcs", "arcs_temp",
             "rt_spikes", "rt_temp", "hotspot", "hotspot_temp", "az_hotspot", "disturb_mod"]
    comp = np.stack([g["state_" + k] for k in names]).astype(np.float32)
    stats, rows = parametric_stats(comp, g["edge"], bool(g["enable_rt"][0]))
    for t in (0.0, 5.0, 50.0, 180.0):
        want = g[f"tex_t{t:g}"]
        got = O.compose_texture(comp, g["omega_rows"], g["edge"], stats, rows, t_offset=t,
                                enable_rt=int(g["enable_rt"][0]), color_temp=float(g["color_temp"][0]))
        assert np.abs(got - want).max() < 1e-4, t
    assert np.abs(g["tex_t50"] - g["tex_t0"]).max() > 0.05       # (the offsets do move the texture)


def parametric_stats(comp, edge, enable_rt=True):
    """upload_parametric_state's statistics, render.py:2363-2379: raw percentiles of the unrotated
    state (no floors, no base-temperature term)."""
    rt_w = 0.20 if enable_rt else 0.0
    density = (0.15 + 0.10 * comp[1] + 0.30 * comp[3] + 0.20 * comp[9] + 0.30 * comp[5] + rt_w * comp[7]) * comp[12]
    density *= edge[:, None]
    ts = (comp[2] + comp[4] + comp[6] + comp[8] + comp[10]) * comp[12]
    scale = float(np.percentile(ts[ts > 0], 95))
    tss = np.clip(ts / (scale + 1e-6) * 0.8, 0, 1.2)
    stats = np.array([float(np.percentile(density, 98)), scale], dtype=np.float32)
    rows = np.stack([np.max(tss, axis=1), np.quantile(tss, 0.7, axis=1)], axis=1).astype(np.float32)
    return stats, rows


def test_mip_pyramid_matches_numpy_generator():
    d = np.load(os.path.join(GOLDEN, "raymarch_aa_tilt_flare.npz"))
    assert np.array_equal(O.build_mips(d["disk_tex"], 5, numpy_order=True), d["mips"])


def test_noise_bit_exact():
    d = np.load(os.path.join(GOLDEN, "noise.npz"))
    assert np.array_equal(O.eval_noise(d["coords"], "simplex"), d["simplex"])
    for k in d.files:
        if k.startswith("fbm_"):
            _, o, p, l = k.split("_")
            assert np.array_equal(O.eval_noise(d["coords"], "fbm", int(o), float(p), float(l)), d[k]), k


def test_noise_properties():
    """The reference's own property tests (tests/unit/test_simplex_noise.py) on the oracle."""
    rng = np.random.RandomState(123)
    c = rng.uniform(-100, 100, size=(5000, 3)).astype(np.float32)
    v = O.eval_noise(c, "simplex")
    assert v.min() >= -1.01 and v.max() <= 1.01 and v.std() > 0.05
    c = rng.uniform(-10, 10, size=(500, 3)).astype(np.float32)
    np.testing.assert_allclose(O.eval_noise(c, "simplex"),
                               O.eval_noise(c, "fbm", octaves=1, persistence=1.0), atol=1e-5)
    v = O.eval_noise(rng.uniform(-50, 50, size=(3000, 3)).astype(np.float32), "fbm", 4, 0.5)
    assert np.abs(v).max() <= 1.875 + 0.1


def test_background_layer():
    t = np.load(os.path.join(GOLDEN, "texture_pipeline.npz"))
    az_f, az_s = int(t["az_freq"][0]), float(t["az_shear"][0])
    for key, time_ in (("init", 0.0), ("f7", 7 * 0.1), ("f25", 25 * 0.1)):
        want = t[key + "_comp"]
        got = want.copy()
        got[[0, 1, 2, 3, 4, 11, 12]] = -1.0
        O.generate_background(got, az_f, az_s, 2.0, 15.0, time_)
        assert np.abs(got - want).max() <= 1e-6, key
        assert np.array_equal(got[5:11], want[5:11])          # entity planes untouched
        np.testing.assert_allclose(got[4], np.float32(0.05) * got[3], atol=1e-7)


def test_entity_layer_compose_mips_stats():
    from black_hole_renderer_b200 import lifecycle as LC
    t = np.load(os.path.join(GOLDEN, "texture_pipeline.npz"))
    n_r, n_phi = 32, 128
    assert np.array_equal(O.edge_alpha(n_r), t["edge"])
    om = O.omega_rows(n_r, 2.0, 15.0)
    assert np.array_equal(om, t["omega_rows"])
    F = LC.make_factories(2.0, 15.0, n_r, n_phi, 42)
    assert np.array_equal(O.accumulate_entities(F, 0.0, n_r, n_phi, om), t["init_comp"][5:11])
    s, rs = O.interactive_stats(t["init_comp"], t["edge"])
    assert np.array_equal(s, t["init_stats"]) and np.array_equal(rs, t["init_row_stats"])
    for f in F.values():
        f.tick(now=0.0, dt=0.0)
    for frame in range(1, 26):
        for f in F.values():
            f.tick(now=frame * 0.1, dt=0.1)
        if frame in (7, 25):
            k = f"f{frame}"
            assert np.array_equal(O.accumulate_entities(F, frame * 0.1, n_r, n_phi, om), t[k + "_comp"][5:11])
            tex = O.compose_texture(t[k + "_comp"], om, t["edge"], t[k + "_stats"], t[k + "_row_stats"])
            assert np.abs(tex - t[k + "_tex"]).max() <= 2e-7
            assert np.abs(O.build_mips(tex, 5, numpy_order=False) - t[k + "_mips"]).max() <= 2e-7
    s, rs = O.interactive_stats(t["f25_comp"], t["edge"])
    assert np.array_equal(s, t["f25_stats"]) and np.array_equal(rs, t["f25_row_stats"])


def test_physics_known_answers():
    """Capture iff impact parameter b < 3*sqrt(3)/2 (rs = 1); weak-field deflection -> 2/b."""
    sky = np.zeros((8, 16, 3), np.float32)
    tex = np.zeros((16, 64, 4), np.float32)
    W = 257
    # camera far away on the x axis, narrow fov: pixel offset maps to impact parameter
    dist, fov = 40.0, 12.0
    sc = O.make_scene(W, 1, [dist, 0.0, 1e-9], fov, sky.shape, tex.shape, step_size=0.05, r_max=80.0)
    rm = O.ray_march(sc, sky, tex)
    term = rm["term"][0]
    pw = 2 * np.tan(np.radians(fov) / 2) / 1 * (W / 1) / W
    x = (np.arange(W) + 0.5 - W / 2) * pw           # tan(angle)
    b = dist * np.abs(x) / np.sqrt(1 + x * x)
    captured = term == 1
    b_crit = 3 * np.sqrt(3) / 2
    assert captured[b < b_crit - 0.05].all()
    assert (~captured[b > b_crit + 0.05]).all()
