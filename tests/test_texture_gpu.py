"""The on-device disk-texture pipeline (noise, background, entity layer, compose, mips) against
the reference goldens and the oracle, plus the reference's own property tests
(tests/unit/test_simplex_noise.py, test_background_layer.py, test_entity_accumulate.py)."""
import os

import numpy as np
import pytest

import oracle as O
from util import GOLDEN

pytestmark = pytest.mark.gpu


def _renderer(n_r=32, n_phi=128, **kw):
    from black_hole_renderer_b200 import Renderer
    return Renderer(16, 8, np.zeros((8, 16, 3), np.float32), np.zeros((n_r, n_phi, 4), np.float32), **kw)


def test_noise_bit_exact_with_reference():
    d = np.load(os.path.join(GOLDEN, "noise.npz"))
    r = _renderer()
    assert np.array_equal(r.eval_noise(d["coords"], "simplex"), d["simplex"])
    for k in d.files:
        if k.startswith("fbm_"):
            _, o, p, l = k.split("_")
            got = r.eval_noise(d["coords"], "fbm", int(o), float(p), float(l))
            assert np.abs(got - d[k]).max() <= 3e-7, k


def test_noise_reference_properties():
    r = _renderer()
    rng = np.random.RandomState(123)
    c = rng.uniform(-100, 100, size=(5000, 3)).astype(np.float32)
    v = r.eval_noise(c, mode="simplex")
    assert np.all(v >= -1.01) and np.all(v <= 1.01) and np.std(v) > 0.05
    assert np.array_equal(v, r.eval_noise(c, mode="simplex"))
    c = rng.uniform(-10, 10, size=(500, 3)).astype(np.float32)
    np.testing.assert_allclose(r.eval_noise(c, "simplex"),
                               r.eval_noise(c, "fbm", octaves=1, persistence=1.0, lacunarity=2.0), atol=1e-5)
    n = 50
    rv = np.linspace(0.1, 1.0, n)
    c0 = np.column_stack([np.cos(0.0) * 8 * np.ones(n), np.sin(0.0) * 8 * np.ones(n), rv * 8 + 0.5]).astype(np.float32)
    c1 = np.column_stack([np.cos(2 * np.pi) * 8 * np.ones(n), np.sin(2 * np.pi) * 8 * np.ones(n), rv * 8 + 0.5]).astype(np.float32)
    np.testing.assert_allclose(r.eval_noise(c0), r.eval_noise(c1), atol=1e-5)
    assert np.array_equal(r.eval_noise(c, "simplex"), O.eval_noise(c, "simplex"))
    assert r.eval_noise(np.zeros((0, 3), np.float32)).shape == (0,)


def test_lifecycle_pipeline_matches_reference_goldens():
    """_init_lifecycle_system + 25 video frames: comp planes, stats, RGBA texture, mips."""
    from black_hole_renderer_b200 import lifecycle as LC
    t = np.load(os.path.join(GOLDEN, "texture_pipeline.npz"))
    n_r, n_phi = 32, 128
    r = _renderer(n_r, n_phi)
    F = LC.init_lifecycle_system(r, n_r, n_phi, seed=42)
    assert r._bg_az_freq == int(t["az_freq"][0]) and r._bg_az_shear == float(t["az_shear"][0])
    comp = r._comp_field.to_numpy()
    assert np.abs(comp - t["init_comp"]).max() <= 1e-6
    assert np.abs(comp[5:11] - t["init_comp"][5:11]).max() <= 1.2e-7      # entity planes
    np.testing.assert_allclose(r._param_stats_field.to_numpy(), t["init_stats"], rtol=1e-5)
    np.testing.assert_allclose(r._param_row_stats_field.to_numpy(), t["init_row_stats"], atol=1e-6)
    assert np.abs(r.disk_texture_field.to_numpy() - t["init_tex"]).max() <= 2e-5
    LC.advance_lifecycle_frame(r, F, t=0.0, dt=0.0, recompute_stats=True)
    assert np.abs(r.disk_texture_field.to_numpy() - t["f0_tex"]).max() <= 2e-5
    for frame in range(1, 26):
        tt = frame * 0.1
        for f in F.values():
            f.tick(now=tt, dt=0.1)
        if frame in (7, 25):
            r.generate_background(t=tt)
            r.accumulate_entity_layer(F, now=tt)
            if frame == 25:
                r.recompute_interactive_stats()
            r.compose_interactive_texture()
            k = f"f{frame}"
            comp = r._comp_field.to_numpy()
            assert np.abs(comp - t[k + "_comp"]).max() <= 1e-6, k
            assert (comp[5:11] == t[k + "_comp"][5:11]).mean() > 0.999
            np.testing.assert_allclose(r._param_stats_field.to_numpy(), t[k + "_stats"], rtol=1e-5)
            assert np.abs(r.disk_texture_field.to_numpy() - t[k + "_tex"]).max() <= 2e-5, k
            assert np.abs(r.disk_mips_field.to_numpy() - t[k + "_mips"]).max() <= 2e-5, k


def test_background_reference_properties():
    """tests/unit/test_background_layer.py: plane ranges / relations, untouched entity planes,
    temporal evolution."""
    r = _renderer(64, 256)
    r.init_background_layer(64, 256, seed=42)
    with_entities = np.random.default_rng(0).random((13, 64, 256)).astype(np.float32)
    r._check(r._lib.bhr_upload_comp(r._ctx, with_entities.ctypes.data_as(
        __import__("ctypes").POINTER(__import__("ctypes").c_float))))
    r.generate_background(0.0)
    c0 = r._comp_field.to_numpy()
    assert np.array_equal(c0[5:11], with_entities[5:11])
    assert np.all(c0[1] == 0) and np.all(c0[2] == 0)
    assert c0[0].min() >= 0 and c0[0].max() <= 0.25 + 1e-6
    assert c0[3].min() >= 0 and c0[3].max() <= 1
    np.testing.assert_allclose(c0[4], 0.05 * c0[3], atol=1e-7)
    assert c0[11].min() >= 0 and c0[11].max() <= 1 and c0[12].min() >= 0.1 and c0[12].max() <= 1
    r.generate_background(5.0)
    c5 = r._comp_field.to_numpy()
    assert np.abs(c5[3] - c0[3]).mean() > 1e-3
    r.generate_background(5.01)
    assert np.abs(r._comp_field.to_numpy()[0] - c5[0]).mean() < 1e-3
    want = c0.copy()
    O.generate_background(want, r._bg_az_freq, r._bg_az_shear, 2.0, 15.0, 0.0)
    assert np.abs(want - c0).max() <= 1e-6


def test_entity_layer_against_oracle_larger_texture():
    from black_hole_renderer_b200 import lifecycle as LC
    n_r, n_phi = 144, 976          # the sd texture of BASELINE.json configs[0]
    r = _renderer(n_r, n_phi)
    r.init_background_layer(n_r, n_phi, seed=42)
    F = LC.make_factories(2.0, 15.0, n_r, n_phi, 42)
    for frame in range(40):
        for f in F.values():
            f.tick(now=frame * 0.1, dt=0.1)
    now = 3.9
    r.accumulate_entity_layer(F, now)
    got = r._comp_field.to_numpy()[5:11]
    want = O.accumulate_entities(F, now, n_r, n_phi, r._bg_omega_all_np)
    assert np.abs(got - want).max() <= 2.4e-7
    assert (got == want).mean() > 0.999
    assert got.min() >= 0 and got[0].max() > 0.1
    # empty factories zero the planes (tests/unit/test_entity_accumulate.py)
    for f in F.values():
        f.entities = []
    r.accumulate_entity_layer(F, now)
    assert np.all(r._comp_field.to_numpy()[5:11] == 0)
    with pytest.raises(AssertionError):
        _renderer(32, 128).generate_background(0.0)


def test_compose_against_oracle():
    r = _renderer(64, 256)
    r.init_background_layer(64, 256, seed=3)
    comp = np.random.default_rng(1).random((13, 64, 256)).astype(np.float32)
    import ctypes as C
    r._check(r._lib.bhr_upload_comp(r._ctx, comp.ctypes.data_as(C.POINTER(C.c_float))))
    r.recompute_interactive_stats()
    s, rs = O.interactive_stats(comp, r._bg_edge_np)
    assert np.array_equal(s, r._param_stats_field.to_numpy())
    assert np.array_equal(rs, r._param_row_stats_field.to_numpy())
    r.compose_interactive_texture()
    want = O.compose_texture(comp, r._bg_omega_all_np, r._bg_edge_np, s, rs)
    got = r.disk_texture_field.to_numpy()
    assert np.abs(got - want).max() <= 2e-5
    assert np.abs(r.disk_mips_field.to_numpy() - O.build_mips(got, 5, numpy_order=False)).max() <= 1e-7


@pytest.mark.parametrize("case", ["ties", "no_structure", "negative", "fhd_texture"])
def test_device_statistics_equal_numpy_on_hard_inputs(case):
    """recompute_interactive_stats on the device (radix select + row sort + numpy's interpolation
    on the exact order statistics) against the numpy restatement, bit for bit: heavy ties, no
    positive structure at all (scale falls back to 1.0), negative values, the fhd texture size."""
    import ctypes as C
    n_r, n_phi = (416, 2912) if case == "fhd_texture" else (48, 208)
    r = _renderer(n_r, n_phi)
    r.init_background_layer(n_r, n_phi, seed=5)
    rng = np.random.default_rng(7)
    comp = rng.random((13, n_r, n_phi)).astype(np.float32)
    if case == "ties":
        comp = (np.round(comp * 4) / 4).astype(np.float32)         # five distinct values per plane
    if case == "no_structure":
        comp[[2, 4, 6, 8, 10]] = 0.0
    if case == "negative":
        comp[[2, 4, 6, 8, 10]] -= 0.45                               # structure sums of both signs
        comp[1] -= 0.5
    r._check(r._lib.bhr_upload_comp(r._ctx, comp.ctypes.data_as(C.POINTER(C.c_float))))
    r.recompute_interactive_stats()
    s, rs = O.interactive_stats(comp, r._bg_edge_np)
    assert np.array_equal(s, r._param_stats_field.to_numpy()), (s, r._param_stats_field.to_numpy())
    assert np.array_equal(rs, r._param_row_stats_field.to_numpy())


def test_legacy_parametric_rotation_path():
    """upload_parametric_state + update_disk_texture_gpu (render.py:2314-2387, 3792-3817): the
    shifted compose against the oracle for several rotation offsets, statistics = raw numpy
    percentiles of the unrotated state."""
    from types import SimpleNamespace
    n_r, n_phi = 64, 256
    rng = np.random.default_rng(3)
    names = ["temp_base", "spiral", "spiral_temp", "turbulence", "turb_temp", "arcs", "arcs_temp",
             "rt_spikes", "rt_temp", "hotspot", "hotspot_temp", "az_hotspot", "disturb_mod"]
    planes = {k: rng.random((n_r, n_phi)).astype(np.float32) for k in names}
    r_vals = 2.0 + 13.0 * np.linspace(0, 1, n_r)
    from black_hole_renderer_b200.renderer import compute_edge_alpha
    state = SimpleNamespace(n_r=n_r, n_phi=n_phi, enable_rt=True, color_temp=6000.0,
                            omega_rows=np.sqrt(0.5 / (r_vals ** 3 + 1e-6)).astype(np.float32),
                            edge=compute_edge_alpha(n_r).astype(np.float32), **planes)
    r = _renderer(n_r, n_phi)
    with pytest.raises(AssertionError):
        r.update_disk_texture_gpu(0.0)
    r.upload_parametric_state(state)
    comp = np.stack([planes[k] for k in names])
    density = (0.15 + 0.10 * comp[1] + 0.30 * comp[3] + 0.20 * comp[9] + 0.30 * comp[5] + 0.20 * comp[7]) * comp[12]
    density *= state.edge[:, None]
    ts = (comp[2] + comp[4] + comp[6] + comp[8] + comp[10]) * comp[12]
    scale = float(np.percentile(ts[ts > 0], 95))
    tss = np.clip(ts / (scale + 1e-6) * 0.8, 0, 1.2)
    stats = np.array([float(np.percentile(density, 98)), scale], dtype=np.float32)
    rows = np.stack([np.max(tss, axis=1), np.quantile(tss, 0.7, axis=1)], axis=1).astype(np.float32)
    assert np.array_equal(stats, r._param_stats_field.to_numpy())
    assert np.array_equal(rows, r._param_row_stats_field.to_numpy())
    for t_offset in (0.0, 7.3, 250.0):
        r.update_disk_texture_gpu(t_offset)
        want = O.compose_texture(comp, state.omega_rows, state.edge, stats, rows, t_offset=t_offset)
        got = r.disk_texture_field.to_numpy()
        assert np.abs(got - want).max() <= 2e-5, t_offset
        assert np.abs(r.disk_mips_field.to_numpy() - O.build_mips(got, 5, numpy_order=False)).max() <= 1e-7
    assert not np.array_equal(got, O.compose_texture(comp, state.omega_rows, state.edge, stats, rows))


def test_shifted_compose_against_the_reference_numpy_generator():
    """f4: upload_parametric_state + update_disk_texture_gpu against the reference's numpy
    _generate_disk_texture_rotating_from_state (tests/golden/shifted_compose.npz, written by
    oracle/make_golden.py from the reference itself) at the offsets and the tolerance of the
    reference's own test (tests/unit/test_gpu_texture_compose.py:98-112: < 1e-4)."""
    from types import SimpleNamespace
    g = np.load(os.path.join(GOLDEN, "shifted_compose.npz"))
    names = ["temp_base", "spiral", "spiral_temp", "turbulence", "turb_temp", "arcs", "arcs_temp",
             "rt_spikes", "rt_temp", "hotspot", "hotspot_temp", "az_hotspot", "disturb_mod"]
    n_r, n_phi = int(g["n_r"][0]), int(g["n_phi"][0])
    state = SimpleNamespace(n_r=n_r, n_phi=n_phi, enable_rt=bool(g["enable_rt"][0]), color_temp=float(g["color_temp"][0]),
                            omega_rows=g["omega_rows"], edge=g["edge"], **{k: g["state_" + k] for k in names})
    r = _renderer(n_r, n_phi)
    r.upload_parametric_state(state)
    np.testing.assert_allclose(r._comp_field.to_numpy(), np.stack([g["state_" + k] for k in names]), atol=1e-6)
    for t in (0.0, 5.0, 50.0, 180.0):
        r.update_disk_texture_gpu(t)
        got = r.disk_texture_field.to_numpy()
        assert np.abs(got - g[f"tex_t{t:g}"]).max() < 1e-4, t
        # mip kernels against generate_disk_mipmaps on the same texture (reference test: < 1e-3)
        assert np.abs(r.disk_mips_field.to_numpy() - O.build_mips(got, 5, numpy_order=True)).max() < 1e-6


@pytest.mark.parametrize("n_r,n_phi", [(32, 128), (144, 976), (416, 2912)])
def test_packed_background_kernel_equals_the_scalar_kernel(n_r, n_phi):
    """background_kernel (two texels per thread in f32x2 lanes, floor / int conversions without the XU
    pipe, row quantities tabulated at init) against background_scalar_kernel (one texel per thread,
    the bit-exact-with-the-reference simplex3 that eval_noise exposes): every packed operation is
    IEEE-rounded per lane in the same order, so all seven planes must be BIT-identical, at several
    times and at the production texture size; and within 1e-6 of the oracle."""
    r = _renderer(n_r, n_phi)
    r.init_background_layer(n_r, n_phi, seed=42)
    for t in (0.0, 0.1, 6.0, 359.9):
        r.set_option("background_scalar", 1)
        r.generate_background(t)
        want = r._comp_field.to_numpy()
        r.set_option("background_scalar", 0)
        r._check(r._lib.bhr_upload_comp(r._ctx, np.full((13, n_r, n_phi), -7.0, np.float32).ctypes.data_as(
            __import__("ctypes").POINTER(__import__("ctypes").c_float))))
        r.generate_background(t)
        got = r._comp_field.to_numpy()
        for pl in (0, 1, 2, 3, 4, 11, 12):
            assert np.array_equal(got[pl], want[pl]), (t, pl, float(np.abs(got[pl] - want[pl]).max()))
        assert (got[5:11] == -7.0).all()                  # the entity planes are not touched
    if n_r <= 144:
        ref = np.zeros((13, n_r, n_phi), np.float32)
        O.generate_background(ref, r._bg_az_freq, r._bg_az_shear, 2.0, 15.0, 359.9)
        assert np.abs(ref[[0, 1, 2, 3, 4, 11, 12]] - got[[0, 1, 2, 3, 4, 11, 12]]).max() <= 1e-6


def test_reference_style_entity_objects_are_accepted():
    """Drop-in boundary: accumulate_entity_layer takes the CALLER'S factory objects -- the reference's EntityFactory /
    EntityInstance (render.py:499-792) carry tabulated phi_density / phi_temp / row_indices arrays, not this package's
    analytic parameters.  Duck-typed copies of such objects (plain namespaces with the reference dataclass's fields and
    methods) go through the tabulated-profile kernel path, which does numpy's float32 `+= roll(row, -shift) * alpha`:
    hotspot / RT-spike planes bit-equal to the reference's numpy loop (oracle.accumulate_entities), filaments <= 1e-6."""
    from types import SimpleNamespace
    from black_hole_renderer_b200 import lifecycle as LC
    n_r, n_phi = 64, 256
    r = _renderer(n_r, n_phi)
    r.init_background_layer(n_r, n_phi, seed=42)
    own = LC.make_factories(2.0, 15.0, n_r, n_phi, seed=42)
    for frame in range(40):
        for f in own.values():
            f.tick(now=frame * 0.1, dt=0.1)
    now = 3.9

    def foreign(e):
        ns = SimpleNamespace(entity_type=e.entity_type, birth_time=e.birth_time, lifetime=e.lifetime, fade_in=e.fade_in,
                             fade_out=e.fade_out, omega=e.omega, row_indices=np.array(e.row_indices),
                             source_phi=e.source_phi, alpha_shear=e.alpha_shear, tau_cool=e.tau_cool,
                             blob_base_r=e.blob_base_r, blob_sigma_r=e.blob_sigma_r, blob_sigma_phi0=e.blob_sigma_phi0,
                             blob_peak_density=e.blob_peak_density, blob_peak_temp=e.blob_peak_temp,
                             phi_density=np.array(e.phi_density) if e.entity_type != "filament" else np.zeros((0, 0), np.float32),
                             phi_temp=np.array(e.phi_temp) if e.entity_type != "filament" else np.zeros((0, 0), np.float32))
        ns.density_factor = e.density_factor
        ns.fade_factor = e.fade_factor
        return ns

    theirs = {k: SimpleNamespace(alive_entities=[foreign(e) for e in f.alive_entities]) for k, f in own.items()}
    want = O.accumulate_entities(theirs, now, n_r, n_phi, r._bg_omega_all_np)
    for _ in range(2):                                   # second call: the table cache is reused
        r.accumulate_entity_layer(theirs, now)
        got = r._comp_field.to_numpy()[5:11]
        assert np.array_equal(got[2:6], want[2:6])       # rt_spike (7, 8) and hotspot (9, 10) planes: bit-equal
        assert np.abs(got[0:2] - want[0:2]).max() <= 1e-6
    assert want[2:6].max() > 0 and want[0:2].max() > 0
    # ... and the package's own factories still take the analytic path, to the same planes within float32 noise
    r.accumulate_entity_layer(own, now)
    assert np.abs(r._comp_field.to_numpy()[5:11] - want).max() <= 2e-6


def test_noise_continuity_and_fbm_bound():
    """tests/unit/test_simplex_noise.py: Lipschitz continuity of the simplex noise and the bound
    sum(persistence^k) of the FBM."""
    r = _renderer()
    rng = np.random.default_rng(5)
    c = rng.uniform(-50, 50, size=(4000, 3)).astype(np.float32)
    step = (rng.standard_normal((4000, 3)) * 1e-3).astype(np.float32)
    d = np.abs(r.eval_noise(c + step) - r.eval_noise(c))
    assert d.max() < 12.0 * np.linalg.norm(step, axis=1).max()          # |grad| of 32 sum t^4 g.x stays below ~10
    v = r.eval_noise(c, "fbm", octaves=4, persistence=0.5, lacunarity=2.0)
    assert np.abs(v).max() <= 1.875 + 1e-5 and np.abs(v).max() > 0.5


def test_pipeline_stage_time_budgets():
    """tests/unit/test_lifecycle_perf.py budgets (background < 500 ms, entities < 200 ms, compose +
    mips < 50 ms, stats < 100 ms, total < 800 ms at a 128 x 784 texture on Taichi-CPU), held here
    at the 1080p texture (416 x 2912) with a hundredth of those budgets -- an order of magnitude
    above what the kernels take, so only a fallen-off-the-device path can trip them."""
    import time
    from black_hole_renderer_b200 import lifecycle as LC
    n_r, n_phi = 416, 2912
    r = _renderer(n_r, n_phi)
    F = LC.init_lifecycle_system(r, n_r, n_phi, seed=42)

    def timed(fn, reps=5):
        fn(); r.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        r.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    ms = dict(background=timed(lambda: r.generate_background(1.0)),
              entities=timed(lambda: r.accumulate_entity_layer(F, 1.0)),
              compose=timed(lambda: r.compose_interactive_texture()),
              stats=timed(lambda: r.recompute_interactive_stats()))
    assert ms["background"] < 5.0 and ms["entities"] < 2.0 and ms["compose"] < 0.5 and ms["stats"] < 2.0, ms
    assert sum(ms.values()) < 8.0, ms
    assert np.isfinite(r.disk_texture_field.to_numpy()).all() and r.disk_texture_field.to_numpy()[..., :3].max() > 0.05


def test_entity_stream_orders_the_planes_like_one_stream():
    """The entity layer runs on its own stream beside the background kernel (options entity_stream / entity_early): the
    component planes, the statistics and the composed texture of a run of video frames must not depend on it -- back to
    back frames without any host synchronisation in between, which is where a missing event would show."""
    from black_hole_renderer_b200 import lifecycle as LC
    n_r, n_phi = 96, 640
    outs = {}
    for mode in ((0, 0), (1, 0), (1, 1)):
        r = _renderer(n_r, n_phi)
        r.set_option("entity_stream", mode[0])
        r.set_option("entity_early", mode[1])
        F = LC.init_lifecycle_system(r, n_r, n_phi, seed=7)
        for frame in range(1, 70):                     # crosses a statistics block boundary (frame 60)
            LC.advance_lifecycle_frame(r, F, t=0.1 * frame, dt=0.1, recompute_stats=(frame % 60 == 0))
            r.render_device([6, 0, 0.5], 90)
        outs[mode] = (r._comp_field.to_numpy(), r.disk_texture_field.to_numpy(), r._param_stats_field.to_numpy())
        r.close()
    for mode in ((1, 0), (1, 1)):
        for a, b in zip(outs[(0, 0)], outs[mode]):
            assert np.array_equal(a, b), mode
