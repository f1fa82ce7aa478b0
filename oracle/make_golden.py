#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by running the REFERENCE ITSELF.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference, which does not
exist on the GPU box):

    python oracle/make_golden.py [--only NAME]

The reference's render.py is imported unmodified with ``oracle/ti_shim.py`` standing in for
``taichi`` (not installable here), so every number written below is produced by
the reference's own kernels / host code:

  raymarch_*.npz   TaichiRenderer.render() at tiny frames (render.py:3865-3923): inputs + the
                   bg / disk / blur fields and the final image
  noise.npz        eval_noise (render.py:3769) simplex + fbm samples
  texture_pipeline.npz  _init_lifecycle_system + _advance_lifecycle_frame (render.py:4079-4153):
                   comp field, stats, composed RGBA texture, mip pyramid
  render_to_field.npz   TaichiRenderer.render_to_field() (render.py:3819-3863) on two of the
                   raymarch_* cases: final_field (W, H, 3) y-flipped, with and without bloom, and
                   the disk layer field it leaves behind
  shifted_compose.npz  the legacy parametric path's numpy side, _generate_disk_texture_rotating_from_state
                   (render.py), at t_offset in {0, 5, 50, 180} -- what the reference's
                   tests/unit/test_gpu_texture_compose.py:55-105 compares the compose kernel with
  host.npz         build_camera, compute_disk_texture_resolution, generate_skybox,
                   EntityFactory parameter streams (render.py:93, 1128, 153, 624)
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BHR_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, HERE)
import ti_shim  # noqa: E402

ti_shim.install()
sys.path.insert(0, REF)
import render as ref  # noqa: E402  (the reference, unmodified)


def smooth_disk_texture(n_r, n_phi, seed):
    """A small synthetic RGBA disk texture with smooth structure and a wide alpha range."""
    rng = np.random.default_rng(seed)
    r = np.linspace(0, 1, n_r)[:, None]
    p = np.linspace(0, 2 * np.pi, n_phi, endpoint=False)[None, :]
    tex = np.zeros((n_r, n_phi, 4), dtype=np.float32)
    for c in range(4):
        a, b = rng.uniform(1, 4, 2)
        k = int(rng.integers(1, 5))
        ph = rng.uniform(0, 6.28)
        tex[..., c] = 0.5 + 0.5 * np.sin(a * r * 6.0 + k * p + ph) * np.cos(b * r * 3.0)
    tex += rng.uniform(0, 0.08, tex.shape).astype(np.float32)
    tex[..., 3] *= 0.9
    return np.clip(tex, 0, 1).astype(np.float32)


RAY_CASES = {
    # name: dict(width, height, pov, fov, step, r_max, r_in, r_out, tilt, flare, aa, aa_strength, dtex)
    "raymarch_default": dict(w=48, h=27, pov=[6, 0, 0.5], fov=90, step=0.1, r_max=10.0,
                             r_in=2.0, r_out=15.0, tilt=0.0, flare=False, aa="disabled",
                             aa_strength=1.0, dtex=(32, 128)),
    "raymarch_aa_tilt_flare": dict(w=40, h=24, pov=[6, 0, 0.5], fov=90, step=0.1, r_max=10.0,
                                   r_in=2.0, r_out=15.0, tilt=20.0, flare=True, aa="lod_radius",
                                   aa_strength=1.0, dtex=(64, 256)),
    "raymarch_e2e_like": dict(w=32, h=18, pov=[6, 0, 0.5], fov=60, step=0.1, r_max=10.0,
                              r_in=2.0, r_out=3.5, tilt=15.0, flare=False, aa="disabled",
                              aa_strength=1.0, dtex=(16, 64)),
    "raymarch_offaxis_fine": dict(w=20, h=12, pov=[4, 3, 2], fov=75, step=0.05, r_max=30.0,
                                  r_in=1.5, r_out=9.0, tilt=-35.0, flare=True, aa="lod_radius",
                                  aa_strength=1.5, dtex=(32, 128)),
    # render(frame != 0): the samplers rotate the texture by t_offset * Omega(r), t_offset = frame * disk_rotation_speed
    # (render.py:3897, 2569-2575, 2601-2607); no live caller of the reference does this, the API does
    "raymarch_frame_rot": dict(w=36, h=20, pov=[6, 0, 0.5], fov=90, step=0.1, r_max=10.0,
                               r_in=2.0, r_out=15.0, tilt=0.0, flare=False, aa="disabled",
                               aa_strength=1.0, dtex=(32, 128), frame=12, speed=0.25),
    "raymarch_frame_rot_aa": dict(w=32, h=18, pov=[5, -2, 1.5], fov=80, step=0.1, r_max=10.0,
                                  r_in=2.0, r_out=12.0, tilt=10.0, flare=False, aa="lod_radius",
                                  aa_strength=1.0, dtex=(64, 256), frame=37, speed=0.1),
}


def gen_raymarch(name, c):
    skybox = ref.generate_skybox(tex_w=128, tex_h=64, seed=7, n_stars=300)
    disk = smooth_disk_texture(c["dtex"][0], c["dtex"][1], seed=11)
    r = ref.TaichiRenderer(c["w"], c["h"], skybox, disk, step_size=c["step"], r_max=c["r_max"],
                           device="cpu", r_disk_inner=c["r_in"], r_disk_outer=c["r_out"],
                           disk_tilt=c["tilt"], lens_flare=c["flare"], anti_alias=c["aa"],
                           aa_strength=c["aa_strength"], disk_rotation_speed=c.get("speed", 0.1))
    frame = c.get("frame", 0)
    t0 = time.time()
    final = r.render(c["pov"], c["fov"], frame=frame)
    out = dict(
        skybox=skybox.astype(np.float32), disk_tex=disk,
        params=np.array([c["w"], c["h"], *c["pov"], c["fov"], c["step"], c["r_max"], c["r_in"],
                         c["r_out"], c["tilt"], float(c["flare"]),
                         float(c["aa"] != "disabled"), c["aa_strength"], frame, c.get("speed", 0.1)], dtype=np.float64),
        # reference fields are (W, H, 3); stored transposed to (H, W, 3) like render()'s return
        bg=r.image_field.to_numpy().transpose(1, 0, 2),
        # NB: read after render(), i.e. after _bloom_kernel's in-place `+= 0.4*blur`
        # (render.py:3112-3114); render() itself uses the pre-bloom copy (render.py:3909)
        disk_layer_after_bloom=r.disk_layer_field.to_numpy().transpose(1, 0, 2),
        blur=r.blur_field.to_numpy().transpose(1, 0, 2),
        final=np.asarray(final, dtype=np.float32),
        mips=r.disk_mips_field.to_numpy(),
    )
    # skip_bloom / skip_differentials variants of the same call (render.py:3911-3912, 3900)
    out["final_skip_bloom"] = np.asarray(
        r.render(c["pov"], c["fov"], frame=frame, skip_bloom=True), dtype=np.float32)
    out["disk_layer"] = r.disk_layer_field.to_numpy().transpose(1, 0, 2)   # pre-bloom
    if c["aa"] != "disabled":
        lf = r.lens_flare
        r.lens_flare = False
        out["final_skip_diff"] = np.asarray(
            r.render(c["pov"], c["fov"], frame=frame, skip_differentials=True, skip_bloom=True),
            dtype=np.float32)
        r.lens_flare = lf
    print(f"  {name}: {time.time() - t0:.1f}s  final mean={final.mean():.5f}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def gen_render_to_field():
    """render_to_field (render.py:3819-3863): ray march -> _bloom_kernel (in-place
    disk = clamp(disk + 0.4 blur), render.py:3112-3114) -> _compose_final_kernel
    (bg + disk + blur, y-flipped, render.py:3285-3300); no lens flare."""
    out = {}
    for name in ("raymarch_default", "raymarch_aa_tilt_flare"):
        c = RAY_CASES[name]
        skybox = ref.generate_skybox(tex_w=128, tex_h=64, seed=7, n_stars=300)
        disk = smooth_disk_texture(c["dtex"][0], c["dtex"][1], seed=11)
        r = ref.TaichiRenderer(c["w"], c["h"], skybox, disk, step_size=c["step"], r_max=c["r_max"],
                               device="cpu", r_disk_inner=c["r_in"], r_disk_outer=c["r_out"],
                               disk_tilt=c["tilt"], lens_flare=c["flare"], anti_alias=c["aa"],
                               aa_strength=c["aa_strength"])
        t0 = time.time()
        r.render_to_field(c["pov"], c["fov"], frame=0)
        out[name + "/final_field"] = r.final_field.to_numpy()                  # (W, H, 3)
        out[name + "/disk_layer_field"] = r.disk_layer_field.to_numpy()        # (W, H, 3) post-bloom
        r.render_to_field(c["pov"], c["fov"], frame=0, skip_bloom=True)
        out[name + "/final_field_skip_bloom"] = r.final_field.to_numpy()
        out[name + "/disk_layer_field_skip_bloom"] = r.disk_layer_field.to_numpy()
        print(f"  render_to_field {name}: {time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(OUT, "render_to_field.npz"), **out)


def gen_shifted_compose():
    """Legacy parametric path, numpy side (render.py:988-1024): the state's 13 component planes
    rolled per row by the Keplerian shift and composed -- the array the reference's
    tests/unit/test_gpu_texture_compose.py:55-105 holds the compose kernel to (< 1e-4)."""
    n_phi, n_r = 256, 64
    st = ref.build_disk_texture_rotating_state(n_phi=n_phi, n_r=n_r, seed=42, r_inner=2.0, r_outer=15.0,
                                               enable_rt=True, generation_scale=1)
    out = dict(n_phi=np.array([n_phi]), n_r=np.array([n_r]), enable_rt=np.array([int(st.enable_rt)]),
               color_temp=np.array([st.color_temp]), omega_rows=st.omega_rows, edge=st.edge)
    for k in ("temp_base", "spiral", "spiral_temp", "turbulence", "turb_temp", "arcs", "arcs_temp",
              "rt_spikes", "rt_temp", "hotspot", "hotspot_temp", "az_hotspot", "disturb_mod"):
        out["state_" + k] = getattr(st, k)
    for t in (0.0, 5.0, 50.0, 180.0):
        out[f"tex_t{t:g}"] = ref._generate_disk_texture_rotating_from_state(st, t_offset=t)
    np.savez_compressed(os.path.join(OUT, "shifted_compose.npz"), **out)


def _tiny_renderer(n_r=16, n_phi=64, r_in=2.0, r_out=15.0):
    skybox = np.zeros((8, 16, 3), dtype=np.float32)
    disk = np.zeros((n_r, n_phi, 4), dtype=np.float32)
    return ref.TaichiRenderer(16, 8, skybox, disk, device="cpu", r_disk_inner=r_in,
                              r_disk_outer=r_out)


def gen_noise():
    r = _tiny_renderer()
    rng = np.random.RandomState(123)
    coords = np.concatenate([
        rng.uniform(-100, 100, size=(300, 3)),
        rng.uniform(-3, 3, size=(200, 3)),
        rng.uniform(-900, 900, size=(100, 3)),
        np.array([[0, 0, 0], [1, 1, 1], [-1, -1, -1], [0.5, 0.5, 0.5], [255.7, 256.2, -256.4]]),
    ]).astype(np.float32)
    out = dict(coords=coords, simplex=r.eval_noise(coords, mode="simplex"))
    for (o, p, l) in [(4, 0.5, 2.0), (1, 1.0, 2.0), (5, 0.45, 2.0), (3, 0.35, 2.0), (4, 0.6, 2.0)]:
        out[f"fbm_{o}_{p}_{l}"] = r.eval_noise(coords, mode="fbm", octaves=o, persistence=p,
                                                lacunarity=l)
    np.savez_compressed(os.path.join(OUT, "noise.npz"), **out)


def _dump_factory(f):
    rows = []
    for e in f.entities:
        rows.append([e.birth_time, e.lifetime, e.omega, e.fade_in, e.fade_out, e.source_phi,
                     e.alpha_shear, e.tau_cool, e.blob_base_r, e.blob_sigma_r, e.blob_sigma_phi0,
                     e.blob_peak_density, e.blob_peak_temp, float(len(e.row_indices)),
                     float(e.row_indices[0]), float(np.sum(e.phi_density, dtype=np.float64)),
                     float(np.sum(e.phi_temp, dtype=np.float64)),
                     float(np.sum(e.fade_noise, dtype=np.float64))])
    return np.array(rows, dtype=np.float64).reshape(len(rows), 18)


def gen_texture_pipeline():
    n_r, n_phi = 32, 128
    r = _tiny_renderer(n_r, n_phi)
    out = {}
    t0 = time.time()
    factories = ref._init_lifecycle_system(r, n_r, n_phi, seed=42)
    out["az_freq"] = np.array([r._bg_az_freq])
    out["az_shear"] = np.array([r._bg_az_shear])
    out["edge"] = r._edge_field.to_numpy()
    out["omega_rows"] = r._omega_rows_field.to_numpy()
    out["init_comp"] = r._comp_field.to_numpy()
    out["init_stats"] = r._param_stats_field.to_numpy()
    out["init_row_stats"] = r._param_row_stats_field.to_numpy()
    out["init_tex"] = r.disk_texture_field.to_numpy()
    out["init_mips"] = r.disk_mips_field.to_numpy()
    for k in ("filament", "hotspot", "rt_spike"):
        out["init_factory_" + k] = _dump_factory(factories[k])
    # single-frame path: t = 0, dt = 0, recompute (render.py:4068)
    ref._advance_lifecycle_frame(r, factories, t=0.0, dt=0.0, recompute_stats=True)
    out["f0_tex"] = r.disk_texture_field.to_numpy()
    # video path: dt = 0.1 per frame; stats only on frame % 60 == 0 (render.py:4456)
    dt = 0.1
    for frame in range(1, 26):
        t = frame * dt
        for f in factories.values():
            f.tick(now=t, dt=dt)
        if frame in (7, 25):
            r.generate_background(t=t)
            r.accumulate_entity_layer(factories, now=t)
            if frame == 25:
                r.recompute_interactive_stats()
            r.compose_interactive_texture()
            out[f"f{frame}_comp"] = r._comp_field.to_numpy()
            out[f"f{frame}_stats"] = r._param_stats_field.to_numpy()
            out[f"f{frame}_row_stats"] = r._param_row_stats_field.to_numpy()
            out[f"f{frame}_tex"] = r.disk_texture_field.to_numpy()
            out[f"f{frame}_mips"] = r.disk_mips_field.to_numpy()
            for k in ("filament", "hotspot", "rt_spike"):
                out[f"f{frame}_factory_" + k] = _dump_factory(factories[k])
    # long host-only run of the factories (RNG call order across culls/spawns)
    for frame in range(26, 900):
        t = frame * dt
        for f in factories.values():
            f.tick(now=t, dt=dt)
    for k in ("filament", "hotspot", "rt_spike"):
        out["f899_factory_" + k] = _dump_factory(factories[k])
    print(f"  texture_pipeline: {time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(OUT, "texture_pipeline.npz"), **out)


def gen_host():
    out = {}
    cams = []
    for pov, fov, w, h in [([6, 0, 0.5], 90, 1920, 1080), ([4, 3, 2], 75, 640, 360),
                           ([0, 0, 8], 60, 320, 180), ([-5.5, 2.25, -1.0], 110, 3840, 2160),
                           ([6.0208, 1e-7, 0.5], 90, 1280, 720)]:
        cp, cr, cu, cf, pw, ph = ref.build_camera(np.array(pov, dtype=np.float64), fov, w, h)
        cams.append(np.concatenate([pov, [fov, w, h], cp, cr, cu, cf, [pw, ph]]))
    out["cameras"] = np.array(cams, dtype=np.float64)
    res = []
    for (w, h) in [(640, 360), (1280, 720), (1920, 1080), (3840, 2160), (320, 180)]:
        for pov, fov, ri, ro in [([6, 0, 0.5], 90, 2.0, 15.0), ([6, 0, 0.5], 60, 2.0, 3.5),
                                 ([12, 5, 1], 45, 3.0, 8.0)]:
            n_phi, n_r = ref.compute_disk_texture_resolution(w, h, pov, fov, ri, ro)
            res.append([w, h, *pov, fov, ri, ro, n_phi, n_r])
    out["tex_resolution"] = np.array(res, dtype=np.float64)
    sky_small = ref.generate_skybox(tex_w=256, tex_h=128, seed=42, n_stars=200)
    out["skybox_256x128_s42_n200"] = sky_small.astype(np.float32)
    t0 = time.time()
    sky = ref.generate_skybox(tex_w=2048, tex_h=1024, seed=42, n_stars=6000)
    out["skybox_full_md5"] = np.array([hashlib.md5(sky.astype(np.float32).tobytes()).hexdigest()])
    out["skybox_full_dtype"] = np.array([str(sky.dtype)])
    out["skybox_full_rowsum"] = sky.astype(np.float64).sum(axis=(1, 2))
    out["skybox_full_sample"] = sky[::64, ::64].astype(np.float32)
    print(f"  skybox full: {time.time() - t0:.1f}s dtype={sky.dtype}")
    out["edge_alpha_37"] = ref.compute_edge_alpha(37)
    out["edge_alpha_416"] = ref.compute_edge_alpha(416)
    np.savez_compressed(os.path.join(OUT, "host.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    jobs = {name: (lambda n=name, c=c: gen_raymarch(n, c)) for name, c in RAY_CASES.items()}
    jobs["noise"] = gen_noise
    jobs["texture_pipeline"] = gen_texture_pipeline
    jobs["host"] = gen_host
    jobs["render_to_field"] = gen_render_to_field
    jobs["shifted_compose"] = gen_shifted_compose
    for name, fn in jobs.items():
        if a.only and a.only != name:
            continue
        print("generating", name)
        fn()


if __name__ == "__main__":
    main()
