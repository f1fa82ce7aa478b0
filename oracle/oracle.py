"""ctypes front-end of the CPU oracle (oracle/bhr_oracle.c) + the host-side numpy restatements.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Every entry point cites the reference lines (render.py in /root/reference) it restates.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


class Scene(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("cam_pos", C.c_float * 3), ("cam_right", C.c_float * 3),
        ("cam_up", C.c_float * 3), ("cam_fwd", C.c_float * 3),
        ("pixel_w", C.c_float), ("pixel_h", C.c_float), ("r_escape", C.c_float),
        ("h_base", C.c_float), ("r_inner", C.c_float), ("r_outer", C.c_float),
        ("t_offset", C.c_float), ("disk_tilt_deg", C.c_float),
        ("skip_diff", C.c_int32), ("aa_mode", C.c_int32), ("aa_strength", C.c_float),
        ("sky_w", C.c_int32), ("sky_h", C.c_int32),
        ("dtex_w", C.c_int32), ("dtex_h", C.c_int32), ("num_mip_levels", C.c_int32),
    ]


def build(force=False):
    """Compile the oracle with gcc (strict IEEE f32, no FMA contraction)."""
    src = os.path.join(_HERE, "bhr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.orc_ray_march.restype = C.c_int64
        L.orc_ray_march.argtypes = [C.POINTER(Scene), fp, fp, fp, C.c_int, C.c_int, fp, fp,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bloom.restype = None
        L.orc_bloom.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, fp]
        L.orc_composite.restype = None
        L.orc_composite.argtypes = [fp, fp, fp, C.c_size_t, fp]
        L.orc_to_u8.restype = None
        L.orc_to_u8.argtypes = [fp, C.c_size_t, C.c_void_p]
        L.orc_lens_flare.restype = None
        L.orc_lens_flare.argtypes = [fp, fp, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_eval_noise.restype = None
        L.orc_eval_noise.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, fp]
        L.orc_generate_background.restype = None
        L.orc_generate_background.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float,
                                              C.c_float, C.c_float, C.c_float]
        L.orc_compose_texture.restype = None
        L.orc_compose_texture.argtypes = [fp, fp, fp, fp, fp, C.c_int, C.c_int, C.c_float,
                                          C.c_int, C.c_float, fp]
        L.orc_build_mips.restype = None
        L.orc_build_mips.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp]
        L.orc_set_escape_dir_out.restype = None
        L.orc_set_escape_dir_out.argtypes = [C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ----------------------------------------------------------------------------------------------
# host glue restated from the reference
# ----------------------------------------------------------------------------------------------
def build_camera(cam_pos, fov_deg, width, height):
    """render.py:93-127 build_camera (float64): look-at-origin pinhole basis + pixel pitch."""
    p = np.array(cam_pos, dtype=np.float64)
    fwd = -p / np.linalg.norm(p)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    rn = np.linalg.norm(right)
    right = np.array([1.0, 0.0, 0.0]) if rn < 1e-6 else right / rn
    up = np.cross(right, fwd)
    up = up / np.linalg.norm(up)
    plane_h = 2.0 * np.tan(np.radians(fov_deg) / 2)
    plane_w = plane_h * (width / height)
    return p, right, up, fwd, plane_w / width, plane_h / height


def make_scene(width, height, cam_pos, fov, sky_shape, dtex_shape, step_size=0.1, r_max=10.0,
               r_inner=2.0, r_outer=15.0, disk_tilt=0.0, anti_alias="disabled", aa_strength=1.0,
               skip_differentials=False, t_offset=0.0, num_mip_levels=5):
    """Scalar uploads of TaichiRenderer.render, render.py:3880-3900."""
    p, right, up, fwd, pw, ph = build_camera(cam_pos, fov, width, height)
    s = Scene()
    s.width, s.height = width, height
    for k in range(3):
        s.cam_pos[k] = np.float32(p[k]); s.cam_right[k] = np.float32(right[k])
        s.cam_up[k] = np.float32(up[k]); s.cam_fwd[k] = np.float32(fwd[k])
    s.pixel_w, s.pixel_h = float(pw), float(ph)
    s.r_escape = max(r_max, float(np.linalg.norm(p)) * 2)
    s.h_base, s.r_inner, s.r_outer = step_size, r_inner, r_outer
    s.t_offset, s.disk_tilt_deg = t_offset, disk_tilt
    s.skip_diff = 1 if skip_differentials else 0
    s.aa_mode = 0 if anti_alias == "disabled" else 1
    s.aa_strength = aa_strength
    s.sky_h, s.sky_w = sky_shape[0], sky_shape[1]
    s.dtex_h, s.dtex_w = dtex_shape[0], dtex_shape[1]
    s.num_mip_levels = num_mip_levels
    return s


def build_mips(base, levels=5, numpy_order=True, zero_pad=True):
    """Padded mip pyramid (levels, n_r, n_phi, 4): render.py:1113-1125 + 2242-2251 (numpy order)
    or the mip kernels render.py:3261-3281 (kernel order)."""
    base = _f32c(base)
    n_r, n_phi = base.shape[:2]
    mips = np.zeros((levels, n_r, n_phi, 4), dtype=np.float32)
    lib().orc_build_mips(_fp(base), n_r, n_phi, levels, int(zero_pad), int(numpy_order), _fp(mips))
    return mips


def ray_march(scene, skybox, disk_tex, mips=None, rows=None, want_aux=True, want_escape_dir=False):
    """_ray_march_kernel, render.py:2787-3018.  Returns dict(bg, disk, term, nhits, steps, total
    [, escape_dir (H, W, 3): the direction the sky lookup used, zeros for captured rays])."""
    skybox, disk_tex = _f32c(skybox), _f32c(disk_tex)
    if mips is None:
        mips = build_mips(disk_tex, scene.num_mip_levels)
    mips = _f32c(mips)
    W, H = scene.width, scene.height
    r0, r1 = rows if rows is not None else (0, H)
    bg = np.zeros((H, W, 3), dtype=np.float32)
    disk = np.zeros((H, W, 3), dtype=np.float32)
    term = np.zeros((H, W), dtype=np.uint8)
    nhits = np.zeros((H, W), dtype=np.uint8)
    steps = np.zeros((H, W), dtype=np.int32)
    esc = np.zeros((H, W, 3), dtype=np.float32) if want_escape_dir else None
    lib().orc_set_escape_dir_out(esc.ctypes.data if esc is not None else None)
    try:
        total = lib().orc_ray_march(C.byref(scene), _fp(skybox), _fp(disk_tex), _fp(mips), r0, r1,
                                    _fp(bg), _fp(disk), term.ctypes.data, nhits.ctypes.data,
                                    steps.ctypes.data)
    finally:
        lib().orc_set_escape_dir_out(None)
    out = dict(bg=bg, disk=disk, term=term, nhits=nhits, steps=steps, total_steps=int(total))
    if esc is not None:
        out["escape_dir"] = esc
    return out


def bloom(disk_layer, width=None):
    """_bloom_kernel as called by render(): render.py:3914-3916, 3022-3110."""
    disk_layer = _f32c(disk_layer)
    H, W = disk_layer.shape[:2]
    width = W if width is None else width
    radius = int(width * 0.02)
    sigma_scale = (width / 640.0) ** 2
    out = np.zeros_like(disk_layer)
    lib().orc_bloom(_fp(disk_layer), W, H, radius, sigma_scale, 0.0, _fp(out))
    return out


def composite(bg, disk, blur=None):
    """render.py:3912 / 3918."""
    bg, disk = _f32c(bg), _f32c(disk)
    out = np.zeros_like(bg)
    lib().orc_composite(_fp(bg), _fp(disk), _fp(_f32c(blur)) if blur is not None else None,
                        bg.size, _fp(out))
    return out


def lens_flare(final, disk):
    """_apply_lens_flare, render.py:3925-4028.  Returns (image, centroid[4])."""
    out = _f32c(final).copy()
    H, W = out.shape[:2]
    cen = (C.c_double * 4)()
    lib().orc_lens_flare(_fp(out), _fp(_f32c(disk)), W, H, cen)
    return out, np.array(list(cen))


def to_u8(img):
    """render.py:423 / 4463 (truncating conversion)."""
    img = _f32c(img)
    out = np.zeros(img.shape, dtype=np.uint8)
    lib().orc_to_u8(_fp(img), img.size, out.ctypes.data)
    return out


def render(width, height, cam_pos, fov, skybox, disk_tex, mips=None, lens_flare_on=False,
           skip_bloom=False, want_escape_dir=False, **kw):
    """TaichiRenderer.render, render.py:3865-3923.  Returns dict with every intermediate."""
    sc = make_scene(width, height, cam_pos, fov, skybox.shape, disk_tex.shape, **kw)
    rm = ray_march(sc, skybox, disk_tex, mips, want_escape_dir=want_escape_dir)
    if skip_bloom:
        rm["blur"] = None
        final = composite(rm["bg"], rm["disk"])
    else:
        rm["blur"] = bloom(rm["disk"], width)
        final = composite(rm["bg"], rm["disk"], rm["blur"])
    if lens_flare_on:
        final, rm["centroid"] = lens_flare(final, rm["disk"])
    rm["final"] = final
    return rm


def render_to_field(width, height, cam_pos, fov, skybox, disk_tex, mips=None, skip_bloom=False, **kw):
    """TaichiRenderer.render_to_field, render.py:3819-3863: ray march, _bloom_kernel INCLUDING its
    in-place tail `disk = clamp(disk + 0.4 blur, 0, 1)` (render.py:3112-3114; f32 multiply, then
    add), then _compose_final_kernel (render.py:3285-3300): clamp((bg + disk) + blur, 0, 1) -- or
    clamp(bg + disk) without bloom -- stored y-flipped as (W, H, 3).  No lens flare on this path.
    Returns dict(final_field (W, H, 3), disk_layer_field (W, H, 3), frame (H, W, 3))."""
    sc = make_scene(width, height, cam_pos, fov, skybox.shape, disk_tex.shape, **kw)
    rm = ray_march(sc, skybox, disk_tex, mips)
    bg, disk = rm["bg"], rm["disk"]
    if skip_bloom:
        frame = np.clip(bg + disk, np.float32(0), np.float32(1))
    else:
        blur = bloom(disk, width)
        disk = np.clip(disk + blur * np.float32(0.4), np.float32(0), np.float32(1))
        frame = np.clip((bg + disk) + blur, np.float32(0), np.float32(1))
    return dict(final_field=np.ascontiguousarray(frame[::-1].transpose(1, 0, 2)),
                disk_layer_field=np.ascontiguousarray(disk.transpose(1, 0, 2)), frame=frame)


def eval_noise(coords, mode="simplex", octaves=4, persistence=0.5, lacunarity=2.0):
    """eval_noise / _noise_eval_kernel, render.py:3769-3790, 3305-3326."""
    coords = _f32c(coords)
    out = np.zeros(coords.shape[0], dtype=np.float32)
    lib().orc_eval_noise(_fp(coords), coords.shape[0], 0 if mode == "simplex" else 1, octaves,
                         persistence, lacunarity, _fp(out))
    return out


def generate_background(comp, az_freq, az_shear, r_inner, r_outer, t):
    """_generate_background_kernel, render.py:3332-3451 (in place on comp (13, n_r, n_phi))."""
    assert comp.dtype == np.float32 and comp.flags.c_contiguous and comp.shape[0] == 13
    lib().orc_generate_background(_fp(comp), comp.shape[1], comp.shape[2], int(az_freq),
                                  float(az_shear), float(r_inner), float(r_outer), float(t))
    return comp


def compose_texture(comp, omega, edge, stats, row_stats, t_offset=0.0, enable_rt=1,
                    color_temp=6000.0):
    """_compose_disk_texture_kernel, render.py:3169-3257."""
    comp = _f32c(comp)
    n_r, n_phi = comp.shape[1:]
    tex = np.zeros((n_r, n_phi, 4), dtype=np.float32)
    lib().orc_compose_texture(_fp(comp), _fp(_f32c(omega)), _fp(_f32c(edge)), _fp(_f32c(stats)),
                              _fp(_f32c(row_stats)), n_r, n_phi, t_offset, enable_rt, color_temp,
                              _fp(tex))
    return tex


# ----------------------------------------------------------------------------------------------
# numpy restatements of the reference's host-side texture-pipeline stages
# ----------------------------------------------------------------------------------------------
def edge_alpha(n, inner_soft=0.1, outer_soft=0.3):
    """compute_edge_alpha, render.py:437-445."""
    v = np.linspace(0, 1, n).astype(np.float32)
    a = np.ones_like(v)
    m_in = v < inner_soft
    m_out = v > (1 - outer_soft)
    a[m_in] = (v[m_in] / inner_soft) ** 3.0
    a[m_out] = ((1 - v[m_out]) / outer_soft) ** 2
    return a


def omega_rows(n_r, r_inner, r_outer):
    """render.py:3517-3519."""
    r = r_inner + (r_outer - r_inner) * np.linspace(0, 1, n_r)
    return np.sqrt(0.5 / (r ** 3 + 1e-6)).astype(np.float32)


def interactive_stats(comp, edge, enable_rt=True):
    """recompute_interactive_stats, render.py:3655-3712.  Returns (stats[2], row_stats[n_r, 2])."""
    rt_w = 0.20 if enable_rt else 0.0
    dm = comp[12]
    density = (0.15 + 0.10 * comp[1] + 0.30 * comp[3] + 0.20 * comp[9] + 0.30 * comp[5]
               + rt_w * comp[7]) * dm
    density *= edge[:, None]
    p98 = max(float(np.percentile(density, 98)), 0.01)
    ts = (comp[2] + comp[4] + comp[6] + comp[8] + comp[10]) * dm
    pos = ts > 0
    scale = float(np.percentile(ts[pos], 95)) if np.any(pos) else 1.0
    scale = max(scale, 0.01)
    tss = np.clip(ts / (scale + 1e-6) * 0.8, 0, 1.2)
    smax = np.max(tss, axis=1).astype(np.float32)
    sp70 = np.quantile(tss, 0.7, axis=1).astype(np.float32)
    tbmax = np.max(comp[0], axis=1).astype(np.float32)
    smax = np.maximum(smax, tbmax)
    sp70 = np.maximum(sp70, tbmax * 0.8)
    return (np.array([p98, scale], dtype=np.float32),
            np.column_stack([smax, sp70]).astype(np.float32))


def initial_stats(n_r):
    """init_background_layer's initial stats, render.py:3532-3542."""
    tb = np.clip(1.0 - np.linspace(0, 1, n_r), 0, 1) ** 1.3 * 0.25
    rows = np.column_stack([np.maximum(tb, 0.25).astype(np.float32),
                            np.maximum(tb * 0.8, 0.10).astype(np.float32)])
    return np.array([0.5, 0.5], dtype=np.float32), rows


def disk_texture_resolution(width, height, cam_pos, fov, r_inner, r_outer):
    """compute_disk_texture_resolution, render.py:1128-1149.  Returns (n_phi, n_r)."""
    d = math.sqrt(cam_pos[0] ** 2 + cam_pos[1] ** 2 + cam_pos[2] ** 2)
    ang = math.atan(r_outer / d)
    frac = fov * math.pi / 180.0
    n_phi = max(256, int(width * (2 * ang / frac)))
    n_r = max(128, int(height * (ang / frac) * 0.5))
    n_phi += (16 - n_phi % 16) % 16
    n_r += (16 - n_r % 16) % 16
    return n_phi, n_r


def accumulate_entities(factories, now, n_r, n_phi, omega_np, r_norm_all=None):
    """accumulate_entity_layer's numpy body, render.py:3585-3649: returns staging (6, n_r, n_phi)
    = comp[5:11].  Works on any objects with the reference's EntityInstance attributes."""
    import math as _m
    staging = np.zeros((6, n_r, n_phi), dtype=np.float32)
    if r_norm_all is None:
        r_norm_all = np.linspace(0, 1, n_r)
    phi_arr = np.linspace(0, 2 * np.pi, n_phi, endpoint=False)
    two_pi = 2 * np.pi
    for key, d_idx, t_idx in (("filament", 0, 1), ("rt_spike", 2, 3), ("hotspot", 4, 5)):
        factory = factories.get(key)
        if factory is None:
            continue
        for e in factory.alive_entities:
            age = now - e.birth_time
            if e.entity_type == "filament":
                if e.density_factor(age) < 0.008:
                    continue
                s0 = max(e.blob_sigma_phi0, 1e-6)
                sig = s0 + e.alpha_shear * age
                amp_d = e.blob_peak_density * s0 / sig
                amp_t = e.blob_peak_temp * s0 / sig
                birth = min(age / 5.0, 1.0)
                cool = _m.exp(-age / e.tau_cool) if e.tau_cool > 0 else 1.0
                sc_d, sc_t = amp_d * birth * cool, amp_t * birth * cool
                i2p = 0.5 / (sig * sig)
                sr = max(e.blob_sigma_r, 1e-6)
                i2r = 0.5 / (sr * sr)
                for ri in e.row_indices:
                    if 0 <= ri < n_r:
                        r_w = _m.exp(-(r_norm_all[ri] - e.blob_base_r) ** 2 * i2r)
                        center = (e.source_phi - omega_np[ri] * age) % two_pi
                        d_phi = phi_arr - center
                        d_phi = d_phi - two_pi * np.round(d_phi / two_pi)
                        prof = np.exp(-d_phi * d_phi * i2p)
                        staging[d_idx, ri] += prof * (sc_d * r_w)
                        staging[t_idx, ri] += prof * (sc_t * r_w)
            else:
                alpha = e.fade_factor(now)
                if alpha <= 0:
                    continue
                for k, ri in enumerate(e.row_indices):
                    if 0 <= ri < n_r:
                        shift = int(age * omega_np[ri] / (2 * np.pi) * n_phi)
                        staging[d_idx, ri] += np.roll(e.phi_density[k], -shift) * alpha
                        staging[t_idx, ri] += np.roll(e.phi_temp[k], -shift) * alpha
    return staging
