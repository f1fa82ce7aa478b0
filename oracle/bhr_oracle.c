/*
 * bhr_oracle.c -- CPU oracle for the per-pixel null-geodesic render path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (black_hole_renderer_b200/, render.py) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and only as the checker / reported CPU baseline.
 *
 * It is a plain-C, strict-IEEE float32 restatement of the arithmetic the reference performs in
 * its Taichi kernels (reference = /root/reference/render.py; every function cites the lines it
 * follows).  Taichi itself (requirements.txt:3, "taichi>=1.6", unpinned, not vendored) is not
 * installable here, so the restatement honours the Taichi semantics listed in SURVEY.md App. D:
 * f32 everywhere, left-to-right dot products, normalized() = v * (1/|v|), integer powers as
 * multiplications, cast-to-int truncation, Python-style integer modulo.  Transcendentals are
 * evaluated in double and rounded to float (an ideal f32 libm).
 *
 * Parity status: PINNED against the reference's own code -- tests/golden/ (npz files) holds outputs of
 * the unmodified reference kernels executed through oracle/ti_shim.py (see make_golden.py), and
 * tests/test_oracle_golden.py checks this file against them.  The reference's only shipped golden
 * (tests/e2e_baseline.txt, an md5 of Taichi-CPU float bytes) is not reproducible without Taichi.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile).
 * Image layout: all images here are (H, W, 3) row-major, i.e. the reference's (W, H) fields
 * transposed exactly as render() does on return (render.py:3923).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------- ideal f32 libm: evaluate in double, round once ---------- */
static inline float f_sqrt(float x) { return sqrtf(x); } /* IEEE exact */
static inline float f_exp(float x) { return (float)exp((double)x); }
static inline float f_log(float x) { return (float)log((double)x); }
static inline float f_sin(float x) { return (float)sin((double)x); }
static inline float f_cos(float x) { return (float)cos((double)x); }
static inline float f_tan(float x) { return (float)tan((double)x); }
static inline float f_acos(float x) { return (float)acos((double)x); }
static inline float f_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
static inline float f_pow(float x, float y) { return (float)pow((double)x, (double)y); }
static inline float f_floor(float x) { return floorf(x); }
static inline float f_min(float a, float b) { return a <= b ? a : b; }
static inline float f_max(float a, float b) { return a >= b ? a : b; }
static inline float f_clamp(float x, float lo, float hi) { return f_min(f_max(x, lo), hi); }
static inline int i_min(int a, int b) { return a < b ? a : b; }
static inline int i_max(int a, int b) { return a > b ? a : b; }
static inline int py_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

#define PI_F 3.14159265358979323846f
#define TWO_PI_F 6.28318530717958647692f

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vs(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }      /* s * v  */
static inline v3 vsr(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }     /* v * s  */
static inline v3 vdiv(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float vnorm(v3 a) { return f_sqrt(vdot(a, a)); }
static inline v3 vnormalized(v3 a) { float inv = 1.0f / vnorm(a); return vs(inv, a); }

/* ------------------------------------------------------------------------------------------
 * Scene description shared by the ray-march entry points (mirrors the kernel arguments and
 * compile-time captures of _ray_march_kernel, render.py:2787-2794 and 2391-2405).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t width, height;
    float cam_pos[3], cam_right[3], cam_up[3], cam_fwd[3];
    float pixel_w, pixel_h, r_escape;
    float h_base, r_inner, r_outer, t_offset, disk_tilt_deg;
    int32_t skip_diff;      /* kernel arg skip_diff (1 => no differentials)            */
    int32_t aa_mode;        /* anti_alias != "disabled"                                */
    float aa_strength;
    int32_t sky_w, sky_h;   /* skybox (sky_h, sky_w, 3)                                */
    int32_t dtex_w, dtex_h; /* disk texture (dtex_h, dtex_w, 4); mips (L, dtex_h, dtex_w, 4) */
    int32_t num_mip_levels;
} orc_scene;

/* render.py:2541-2566  _sample_skybox */
static v3 sample_skybox(const orc_scene *s, const float *sky, v3 d)
{
    const int tw = s->sky_w, th = s->sky_h;
    float theta = f_acos(f_min(f_max(d.z, -1.0f), 1.0f));
    float phi = f_atan2(d.y, d.x);
    if (phi < 0) phi += TWO_PI_F;
    float u = phi / TWO_PI_F * (float)tw;
    float v = theta / PI_F * (float)th;
    int u0 = (int)f_floor(u), v0 = (int)f_floor(v);
    float fu = u - (float)u0, fv = v - (float)v0;
    int u0w = py_mod(u0, tw), u1w = py_mod(u0 + 1, tw);
    int v0h = i_min(i_max(v0, 0), th - 1), v1h = i_min(i_max(v0 + 1, 0), th - 1);
    const float *c00 = sky + ((size_t)v0h * tw + u0w) * 3, *c10 = sky + ((size_t)v0h * tw + u1w) * 3;
    const float *c01 = sky + ((size_t)v1h * tw + u0w) * 3, *c11 = sky + ((size_t)v1h * tw + u1w) * 3;
    float o[3];
    for (int k = 0; k < 3; ++k)
        o[k] = ((c00[k] * (1 - fu) * (1 - fv) + c10[k] * fu * (1 - fv)) + c01[k] * (1 - fu) * fv)
               + c11[k] * fu * fv;
    return V(o[0], o[1], o[2]);
}

/* shared (phi, r) part of render.py:2568-2581 / 2600-2611 */
static inline void disk_polar(float hx, float hy, float t_offset, float *r_out, float *phi_out)
{
    float r = f_sqrt(hx * hx + hy * hy);
    float phi = f_atan2(hy, hx);
    float r_safe = f_max(r, 1e-3f);
    float omega = f_sqrt(0.5f / (r_safe * r_safe * r_safe + 1e-6f));
    phi = phi + t_offset * omega;
    while (phi < 0) phi += TWO_PI_F;
    while (phi >= TWO_PI_F) phi -= TWO_PI_F;
    *r_out = r; *phi_out = phi;
}

static inline void bilerp4(const float *c00, const float *c10, const float *c01, const float *c11,
                           float fu, float fv, float out[4])
{
    for (int k = 0; k < 4; ++k)
        out[k] = ((c00[k] * (1 - fu) * (1 - fv) + c10[k] * fu * (1 - fv)) + c01[k] * (1 - fu) * fv)
                 + c11[k] * fu * fv;
}

/* render.py:2568-2598  _sample_disk */
static void sample_disk(const orc_scene *s, const float *tex, float hx, float hy, float out[4])
{
    const int tw = s->dtex_w, th = s->dtex_h;
    float r, phi;
    disk_polar(hx, hy, s->t_offset, &r, &phi);
    float u = phi / TWO_PI_F * (float)tw;
    float v = (r - s->r_inner) / (s->r_outer - s->r_inner) * (float)th;
    int u0 = (int)f_floor(u), v0 = (int)f_floor(v);
    float fu = u - (float)u0, fv = v - (float)v0;
    int u0w = py_mod(u0, tw), u1w = py_mod(u0 + 1, tw);
    int v0h = i_min(i_max(v0, 0), th - 1), v1h = i_min(i_max(v0 + 1, 0), th - 1);
    bilerp4(tex + ((size_t)v0h * tw + u0w) * 4, tex + ((size_t)v0h * tw + u1w) * 4,
            tex + ((size_t)v1h * tw + u0w) * 4, tex + ((size_t)v1h * tw + u1w) * 4, fu, fv, out);
}

/* render.py:2600-2637  _sample_disk_mip (mips padded to base size, level in the top-left) */
static void sample_disk_mip(const orc_scene *s, const float *mips, float hx, float hy, float lod,
                            float out[4])
{
    const int tw = s->dtex_w, th = s->dtex_h;
    float r, phi;
    disk_polar(hx, hy, s->t_offset, &r, &phi);
    int lod_i = (int)f_min(f_max(lod, 0.0f), (float)(s->num_mip_levels - 1));
    float scale = f_pow(2.0f, (float)lod_i);
    float tw_l = (float)tw / scale, th_l = (float)th / scale;
    float u = phi / TWO_PI_F * tw_l;
    float v = (r - s->r_inner) / (s->r_outer - s->r_inner) * th_l;
    int u0 = (int)f_floor(u), v0 = (int)f_floor(v);
    float fu = u - (float)u0, fv = v - (float)v0;
    int twi = (int)tw_l;
    int u0w = py_mod(u0, twi), u1w = py_mod(u0 + 1, twi);
    int vmax = (int)(th_l - 1);
    int v0h = i_min(i_max(v0, 0), vmax), v1h = i_min(i_max(v0 + 1, 0), vmax);
    const float *base = mips + (size_t)lod_i * th * tw * 4;
    bilerp4(base + ((size_t)v0h * tw + u0w) * 4, base + ((size_t)v0h * tw + u1w) * 4,
            base + ((size_t)v1h * tw + u0w) * 4, base + ((size_t)v1h * tw + u1w) * 4, fu, fv, out);
}

/* render.py:2407-2437  _color_temp_to_tint (Tanner Helland) */
static v3 color_temp_to_tint(float temp)
{
    float t = temp / 100.0f;
    float r = 1.0f, g, b = 1.0f;
    if (t > 66.0f)
        r = f_min(f_max(1.292936f * f_pow(f_max(t - 60.0f, 0.0001f), -0.1332047592f), 0.0f), 1.0f);
    if (t <= 66.0f)
        g = f_min(f_max(0.390082f * f_log(f_max(t, 0.0001f)) - 0.631841f, 0.0f), 1.0f);
    else
        g = f_min(f_max(1.129891f * f_pow(f_max(t - 60.0f, 0.0001f), -0.0755148492f), 0.0f), 1.0f);
    if (t < 66.0f) {
        if (t <= 19.0f) b = 0.0f;
        else b = f_min(f_max(0.543207f * f_log(f_max(t - 10.0f, 0.0001f)) - 1.19625f, 0.0f), 1.0f);
    }
    return V(r, g, b);
}

/* render.py:2439-2516  _apply_g_factor; constants render.py:42-59 */
static v3 apply_g_factor(v3 base, v3 hit, float hit_r, v3 ray_to_cam, v3 cam_pos, float r_inner,
                         float r_outer, float tilt_rad)
{
    const float g_cap = 1.5f, lum_power = 1.5f, gain = 0.38f, color_temp = 6000.0f;
    float r_obs = vnorm(cam_pos);
    float r_em = vnorm(hit);
    float r_safe = f_max(r_em, 1.0f + 1e-3f);
    float omega = f_sqrt(0.5f / (r_safe * r_safe * r_safe + 1e-6f));
    float lorentz = f_sqrt(f_max(1.0f - 1.0f / r_safe, 1e-6f));
    float beta = f_min(r_safe * omega / f_max(lorentz, 1e-6f), 0.99f);
    float gamma = 1.0f / f_sqrt(f_max(1.0f - beta * beta, 1e-6f));
    float sin_t = f_sin(tilt_rad), cos_t = f_cos(tilt_rad);
    v3 normal = V(0.0f, -sin_t, cos_t);
    v3 r_hat = vnormalized(hit);
    v3 v_hat = vcross(r_hat, normal);
    float v_norm = vnorm(v_hat);
    if (v_norm > 1e-6f) v_hat = vdiv(v_hat, v_norm);
    else v_hat = V(0.0f, 1.0f, 0.0f);
    v3 ray_hat = vnormalized(ray_to_cam);
    float cos_theta = vdot(v_hat, ray_hat);
    float denom = f_max(1.0f - beta * cos_theta, 1e-3f);
    float g_doppler = 1.0f / (gamma * denom);
    float grav_num = f_sqrt(f_max(1.0f - 1.0f / f_max(r_obs, 1.0f + 1e-3f), 1e-6f));
    float grav_den = f_sqrt(f_max(1.0f - 1.0f / f_max(r_em, 1.0f + 1e-3f), 1e-6f));
    float g_grav = grav_num / grav_den;
    float g = f_min(g_doppler * g_grav, g_cap);
    float intensity = f_max(f_pow(g, lum_power), 0.0f);
    float brightness = gain * intensity / (1.0f + intensity / g_cap);
    float span = f_max(r_outer - r_inner, 1e-3f);
    float radial_t = (f_max(hit_r, r_inner) - r_inner) / span;
    radial_t = f_min(f_max(radial_t, 0.0f), 1.0f);
    float profile = f_pow(1.0f - radial_t, 1.2f);
    const float min_boost = 0.2f, max_boost = 8.0f;
    float boost = min_boost + (max_boost - min_boost) * profile;
    brightness *= boost;
    float g_safe = f_max(g, 0.1f);
    float wien = 1.0f - 1.0f / g_safe;
    float rs = f_exp(2.21f * wien), gs = f_exp(2.72f * wien), bs = f_exp(3.13f * wien);
    rs = f_min(rs / gs, 3.0f);
    bs = f_min(bs / gs, 3.0f);
    v3 shifted = V(base.x * rs, base.y * 1.0f, base.z * bs);
    v3 tint = color_temp_to_tint(color_temp);
    v3 c = vsr(vmul(shifted, tint), brightness);
    return V(f_clamp(c.x, 0.0f, 10.0f), f_clamp(c.y, 0.0f, 10.0f), f_clamp(c.z, 0.0f, 10.0f));
}

/* render.py:2518-2524  _compute_acceleration */
static inline v3 accel(v3 p, float L2)
{
    float r2 = vdot(p, p);
    float r = f_sqrt(r2);
    float r5 = r2 * r2 * r;
    return vs(-1.5f * L2 / r5, p);
}

/* render.py:2526-2539  _compute_acc_jacobian applied to d */
static inline v3 accel_jac(v3 p, v3 d, float L2)
{
    float r2 = vdot(p, p);
    float r = f_sqrt(r2);
    float r5 = r2 * r2 * r;
    float factor = -1.5f * L2 / r5;
    float proj = vdot(p, d) / r2;
    return vs(factor, vsub(d, vsr(vs(5.0f, p), proj)));
}

/* one variational RK4 step (render.py:2889-2899): returns new (dpos, ddir) */
static inline void rk4_diff(v3 pos, v3 k1p, v3 k2p, v3 k3p, float h, float L2, v3 dp, v3 dd,
                            v3 *ndp, v3 *ndd)
{
    v3 a1p = vs(h, dd);
    v3 a1d = vs(h, accel_jac(pos, dp, L2));
    v3 a2p = vs(h, vadd(dd, vs(0.5f, a1d)));
    v3 a2d = vs(h, accel_jac(vadd(pos, vs(0.5f, k1p)), vadd(dp, vs(0.5f, a1p)), L2));
    v3 a3p = vs(h, vadd(dd, vs(0.5f, a2d)));
    v3 a3d = vs(h, accel_jac(vadd(pos, vs(0.5f, k2p)), vadd(dp, vs(0.5f, a2p)), L2));
    v3 a4p = vs(h, vadd(dd, a3d));
    v3 a4d = vs(h, accel_jac(vadd(pos, k3p), vadd(dp, a3p), L2));
    *ndp = vadd(dp, vdiv(vadd(vadd(vadd(a1p, vs(2.0f, a2p)), vs(2.0f, a3p)), a4p), 6.0f));
    *ndd = vadd(dd, vdiv(vadd(vadd(vadd(a1d, vs(2.0f, a2d)), vs(2.0f, a3d)), a4d), 6.0f));
}

/* termination codes written to out_term */
enum { TERM_EXHAUSTED = 0, TERM_HORIZON = 1, TERM_ESCAPED = 2 };

/*
 * render.py:2787-3018  _ray_march_kernel for one pixel (i = x, j = y).
 * Outputs: bg[3], disk[3]; *term, *nhits (crossings inside [r_inner, r_outer]), *steps = number
 * of RK4 evaluations including the terminating one (SURVEY.md 8d flop accounting).
 */
/* diagnostic output for the parity tests: when set, trace_pixel stores the escape direction of
 * every pixel ((H, W, 3), zeros for rays that do not escape) -- the sky lookup's input, which
 * tells a pole-direction outlier (d phi = d dir / sin theta) from a real difference */
static float *g_escape_dir_out = NULL;
void orc_set_escape_dir_out(float *p) { g_escape_dir_out = p; }

static void trace_pixel(const orc_scene *s, const float *sky, const float *tex, const float *mips,
                        int i, int j, float bg[3], float disk[3], uint8_t *term, uint8_t *nhits,
                        int32_t *steps)
{
    const v3 cp = V(s->cam_pos[0], s->cam_pos[1], s->cam_pos[2]);
    const v3 cr = V(s->cam_right[0], s->cam_right[1], s->cam_right[2]);
    const v3 cu = V(s->cam_up[0], s->cam_up[1], s->cam_up[2]);
    const v3 cf = V(s->cam_fwd[0], s->cam_fwd[1], s->cam_fwd[2]);
    const float pw = s->pixel_w, ph = s->pixel_h;
    const float tilt_rad = s->disk_tilt_deg * PI_F / 180.0f;
    const float min_fac = 0.2f, max_fac = 10.0f, r_cap = 1.0f;
    const v3 center = vadd(cp, vsr(cf, 1.0f));
    const v3 tl = vadd(vsub(center, vsr(cr, pw * (float)s->width / 2.0f)),
                       vsr(cu, ph * (float)s->height / 2.0f));
    const float r_esc = s->r_escape;
    const int max_iter = (int)(r_esc * 40.0f / s->h_base);
    const float max_affine = r_esc * 40.0f;
    const float h_base = s->h_base, r_inner = s->r_inner, r_outer = s->r_outer;
    const int skip_diff = s->skip_diff;

    float px_f = (float)i, py_f = (float)j;
    v3 pixel_pos = vsub(vadd(tl, vs((px_f + 0.5f) * pw, cr)), vs((py_f + 0.5f) * ph, cu));
    v3 ray_dir = vnormalized(vsub(pixel_pos, cp));
    v3 pos = cp, dir = ray_dir;
    float nrm = vnorm(vcross(dir, pos));
    float L2 = nrm * nrm;

    v3 dpx = V(0, 0, 0), ddx = V(0, 0, 0), dpy = V(0, 0, 0), ddy = V(0, 0, 0);
    if (skip_diff == 0) {
        v3 px1 = vsub(vadd(tl, vs((px_f + 1.5f) * pw, cr)), vs((py_f + 0.5f) * ph, cu));
        ddx = vsub(vnormalized(vsub(px1, cp)), ray_dir);
        v3 py1 = vsub(vadd(tl, vs((px_f + 0.5f) * pw, cr)), vs((py_f + 1.5f) * ph, cu));
        ddy = vsub(vnormalized(vsub(py1, cp)), ray_dir);
    }

    int escaped = 0, horizon = 0;
    v3 esc_dir = V(0, 0, 0);
    v3 accum = V(0, 0, 0);
    float alpha_total = 0.0f, affine = 0.0f;
    int step_count = 0, evals = 0, hits = 0;
    v3 hit_dpx = V(0, 0, 0), hit_dpy = V(0, 0, 0);
    const float tan_t = f_tan(tilt_rad);

    while (step_count < max_iter) {
        v3 old_pos = pos;
        float r_cur = vnorm(pos);
        float r_safe = f_max(r_cur, r_cap + 1e-3f);
        float far_scale = f_sqrt(r_safe / r_cap);
        if (far_scale > max_fac) far_scale = max_fac;
        float q = r_cap / r_safe;
        float near_damp = 1.0f / (1.0f + 2.0f * (q * q * q));
        float dt_fac = far_scale * near_damp;
        if (dt_fac < min_fac) dt_fac = min_fac;
        if (dt_fac > max_fac) dt_fac = max_fac;
        float h = h_base * dt_fac;

        v3 k1p = vs(h, dir);
        v3 k1d = vs(h, accel(pos, L2));
        v3 k2p = vs(h, vadd(dir, vs(0.5f, k1d)));
        v3 k2d = vs(h, accel(vadd(pos, vs(0.5f, k1p)), L2));
        v3 k3p = vs(h, vadd(dir, vs(0.5f, k2d)));
        v3 k3d = vs(h, accel(vadd(pos, vs(0.5f, k2p)), L2));
        v3 k4p = vs(h, vadd(dir, k3d));
        v3 k4d = vs(h, accel(vadd(pos, k3p), L2));
        v3 new_pos = vadd(pos, vdiv(vadd(vadd(vadd(k1p, vs(2.0f, k2p)), vs(2.0f, k3p)), k4p), 6.0f));
        v3 new_dir = vadd(dir, vdiv(vadd(vadd(vadd(k1d, vs(2.0f, k2d)), vs(2.0f, k3d)), k4d), 6.0f));

        v3 ndpx = dpx, nddx = ddx, ndpy = dpy, nddy = ddy;
        if (skip_diff == 0) {
            rk4_diff(pos, k1p, k2p, k3p, h, L2, dpx, ddx, &ndpx, &nddx);
            rk4_diff(pos, k1p, k2p, k3p, h, L2, dpy, ddy, &ndpy, &nddy);
        }

        float r = vnorm(new_pos);
        affine += h;
        ++evals;

        if (r < r_cap) { horizon = 1; break; }
        else if (r > r_esc) { escaped = 1; esc_dir = vnormalized(new_dir); break; }
        else if (affine > max_affine) { escaped = 1; esc_dir = vnormalized(new_dir); break; }

        if (skip_diff == 0) { dpx = ndpx; ddx = nddx; dpy = ndpy; ddy = nddy; }

        float f_old = old_pos.z - old_pos.y * tan_t;
        float f_new = new_pos.z - new_pos.y * tan_t;
        if (f_old * f_new < 0) {
            float t = f_old / (f_old - f_new + 1e-8f);
            float hx = old_pos.x + t * (new_pos.x - old_pos.x);
            float hy = old_pos.y + t * (new_pos.y - old_pos.y);
            float hr = f_sqrt(hx * hx + hy * hy);
            if (skip_diff == 0) {
                /* render.py:2947-2949: dpx was already overwritten with ndpx (2928-2932), so this
                 * "lerp" yields the end-of-step differential. */
                hit_dpx = vadd(dpx, vs(t, vsub(ndpx, dpx)));
                hit_dpy = vadd(dpy, vs(t, vsub(ndpy, dpy)));
            }
            if (r_outer >= hr && hr >= r_inner) {
                float hz = hy * tan_t;
                float rgba[4];
                if (s->aa_mode == 0 || skip_diff == 1) {
                    sample_disk(s, tex, hx, hy, rgba);
                } else {
                    float rc = f_sqrt(hx * hx + hy * hy + 1e-6f);
                    float dr_dx = (hx * hit_dpx.x + hy * hit_dpx.y) / rc;
                    float dphi_dx = (-hy * hit_dpx.x + hx * hit_dpx.y) / (rc * rc + 1e-6f);
                    float dudx = dphi_dx * (float)s->dtex_w / TWO_PI_F;
                    float dvdx = dr_dx * (float)s->dtex_h / (r_outer - r_inner);
                    float dr_dy = (hx * hit_dpy.x + hy * hit_dpy.y) / rc;
                    float dphi_dy = (-hy * hit_dpy.x + hx * hit_dpy.y) / (rc * rc + 1e-6f);
                    float dudy = dphi_dy * (float)s->dtex_w / TWO_PI_F;
                    float dvdy = dr_dy * (float)s->dtex_h / (r_outer - r_inner);
                    float gx = dudx * dudx + dvdx * dvdx;
                    float gy = dudy * dudy + dvdy * dvdy;
                    float g2 = f_max(gx, gy);
                    float lod = f_log(f_max(g2, 1.0f)) / f_log(2.0f) * s->aa_strength;
                    lod = f_min(f_max(lod, 0.0f), 3.0f);
                    sample_disk_mip(s, mips, hx, hy, lod, rgba);
                }
                float base_alpha = f_min(rgba[3], 0.999f);
                float a = 1.0f - f_pow(1.0f - base_alpha, 6.0f);
                v3 col = apply_g_factor(V(rgba[0], rgba[1], rgba[2]), V(hx, hy, hz), hr, vneg(dir),
                                        cp, r_inner, r_outer, tilt_rad);
                float front = 1.0f - alpha_total;
                accum = vadd(accum, vsr(vsr(col, a), front));
                alpha_total = 1.0f - front * (1.0f - a);
                ++hits;
            }
        }
        pos = new_pos;
        dir = new_dir;
        ++step_count;
    }

    v3 bgc = V(0, 0, 0);
    if (horizon) bgc = V(0, 0, 0);
    else if (escaped) bgc = sample_skybox(s, sky, esc_dir);
    bgc = vsr(bgc, 1.0f - alpha_total);
    bg[0] = bgc.x; bg[1] = bgc.y; bg[2] = bgc.z;
    disk[0] = f_clamp(accum.x, 0.0f, 1.0f);
    disk[1] = f_clamp(accum.y, 0.0f, 1.0f);
    disk[2] = f_clamp(accum.z, 0.0f, 1.0f);
    if (g_escape_dir_out) {
        float *e = g_escape_dir_out + ((size_t)j * s->width + i) * 3;
        e[0] = esc_dir.x; e[1] = esc_dir.y; e[2] = esc_dir.z;
    }
    *term = horizon ? TERM_HORIZON : (escaped ? TERM_ESCAPED : TERM_EXHAUSTED);
    *nhits = (uint8_t)(hits > 255 ? 255 : hits);
    *steps = evals;
}

/*
 * Ray-march rows [row0, row1) of the frame.  bg/disk are (H, W, 3); term/nhits (H, W) u8;
 * steps (H, W) i32.  Any of term/nhits/steps may be NULL.  Returns the sum of steps.
 */
int64_t orc_ray_march(const orc_scene *s, const float *sky, const float *tex, const float *mips,
                      int row0, int row1, float *bg, float *disk, uint8_t *term, uint8_t *nhits,
                      int32_t *steps)
{
    const int W = s->width;
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int j = row0; j < row1; ++j) {
        for (int i = 0; i < W; ++i) {
            size_t p = (size_t)j * W + i;
            uint8_t t, n;
            int32_t st;
            trace_pixel(s, sky, tex, mips, i, j, bg + p * 3, disk + p * 3, &t, &n, &st);
            if (term) term[p] = t;
            if (nhits) nhits[p] = n;
            if (steps) steps[p] = st;
            total += st;
        }
    }
    return total;
}

/* ------------------------------------------------------------------------------------------
 * Bloom: render.py:3022-3114 as called from render() (3914-3916): threshold 0, so the live
 * result is H pass (along x) then V pass (along y) of the disk layer, per-channel Gaussian
 * weights, renormalised by the in-bounds weight sum.  `src` and `out` are (H, W, 3).
 * The bright pass (lum > threshold ? c : 0) is kept for fidelity.
 * ---------------------------------------------------------------------------------------- */
void orc_bloom(const float *src, int W, int H, int radius, float sigma_scale, float threshold,
               float *out)
{
    size_t n = (size_t)W * H;
    float *bright = (float *)malloc(n * 3 * sizeof(float));
    float *tmp = (float *)malloc(n * 3 * sizeof(float));
    int taps = 2 * radius + 1;
    float *wt = (float *)malloc((size_t)taps * 3 * sizeof(float));
    for (int d = -radius; d <= radius; ++d) {
        float dist_sq = (float)(d * d);
        wt[(d + radius) * 3 + 0] = f_exp(-dist_sq / (25.0f * sigma_scale));
        wt[(d + radius) * 3 + 1] = f_exp(-dist_sq / (80.0f * sigma_scale));
        wt[(d + radius) * 3 + 2] = f_exp(-dist_sq / (1600.0f * sigma_scale));
    }
#pragma omp parallel for
    for (long p = 0; p < (long)n; ++p) {
        const float *c = src + p * 3;
        float lum = (c[0] * 0.2126f + c[1] * 0.7152f) + c[2] * 0.0722f;
        for (int k = 0; k < 3; ++k) bright[p * 3 + k] = lum > threshold ? c[k] : 0.0f;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const float *in = pass == 0 ? bright : tmp;
        float *o = pass == 0 ? tmp : out;
#pragma omp parallel for
        for (int j = 0; j < H; ++j) {
            for (int i = 0; i < W; ++i) {
                float sum[3] = {0, 0, 0}, wsum[3] = {0, 0, 0};
                for (int d = -radius; d <= radius; ++d) {
                    int ni = pass == 0 ? i + d : i, nj = pass == 0 ? j : j + d;
                    if (ni < 0 || ni >= W || nj < 0 || nj >= H) continue;
                    const float *c = in + ((size_t)nj * W + ni) * 3;
                    const float *w = wt + (d + radius) * 3;
                    for (int k = 0; k < 3; ++k) { sum[k] += c[k] * w[k]; wsum[k] += w[k]; }
                }
                float *dst = o + ((size_t)j * W + i) * 3;
                if (wsum[0] > 0.0f) for (int k = 0; k < 3; ++k) dst[k] = sum[k] / wsum[k];
                else dst[0] = dst[1] = dst[2] = 0.0f;
            }
        }
    }
    free(bright); free(tmp); free(wt);
}

/* render.py:3912 / 3918: final = clip(bg + disk [+ blur], 0, 1); blur may be NULL (skip_bloom) */
void orc_composite(const float *bg, const float *disk, const float *blur, size_t n3, float *out)
{
    for (size_t k = 0; k < n3; ++k) {
        float v = bg[k] + disk[k];
        if (blur) v = v + blur[k];
        out[k] = f_clamp(v, 0.0f, 1.0f);
    }
}

/* render.py:423 / 4463: (clip(img, 0, 1) * 255).astype(uint8) -- truncation */
void orc_to_u8(const float *img, size_t n3, uint8_t *out)
{
    for (size_t k = 0; k < n3; ++k) out[k] = (uint8_t)(f_clamp(img[k], 0.0f, 1.0f) * 255.0f);
}

/* numpy's pairwise float32 summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum),
 * used by np.sum(disk_brightness) at render.py:3933 on a contiguous f32 array. */
static float np_pairwise_sum_f32(const float *a, size_t n)
{
    if (n < 8) {
        float res = 0.0f;
        for (size_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        size_t i;
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        size_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
    }
}

static double np_pairwise_sum_f64(const double *a, size_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (size_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        size_t i;
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        size_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f64(a, n2) + np_pairwise_sum_f64(a + n2, n - n2);
    }
}

/*
 * Lens flare: render.py:3925-4028 (_apply_lens_flare, host numpy with float64 temporaries).
 * `final` (in/out) and `disk` are (H, W, 3).  The reference works on (W, H) arrays with
 * x = first index; the brightness reductions below therefore run in x-major order.
 * centroid_out (optional) receives {total_brightness, light_x, light_y, intensity}.
 */
void orc_lens_flare(float *final, const float *disk, int W, int H, double *centroid_out)
{
    size_t n = (size_t)W * H;
    double scale = (double)(W < H ? W : H) / 360.0;
    float *br = (float *)malloc(n * sizeof(float));
    double *tmp = (double *)malloc(n * sizeof(double));
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) {
            const float *c = disk + ((size_t)y * W + x) * 3;
            br[(size_t)x * H + y] = f_max(f_max(c[0], c[1]), c[2]);
        }
    float total_f = np_pairwise_sum_f32(br, n);
    if (centroid_out) { centroid_out[0] = total_f; centroid_out[1] = centroid_out[2] = centroid_out[3] = 0; }
    if (total_f < 0.01f) { free(br); free(tmp); return; }
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) tmp[(size_t)x * H + y] = (double)x * (double)br[(size_t)x * H + y];
    double light_x = np_pairwise_sum_f64(tmp, n) / (double)total_f;
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) tmp[(size_t)x * H + y] = (double)y * (double)br[(size_t)x * H + y];
    double light_y = np_pairwise_sum_f64(tmp, n) / (double)total_f;
    free(br); free(tmp);
    double scx = W / 2.0, scy = H / 2.0;
    /* total_brightness is np.float32, so `total / (w*h*0.3)`, `* 1.5` and `intensity * 0.3` are
     * float32 operations (NEP 50 weak Python scalars) unless min() returned the Python float 1.0 */
    float qf = total_f / (float)((double)(W * H) * 0.3);
    double intensity, streak_alpha;
    if (1.0f < qf) { intensity = 1.0 * 1.5; streak_alpha = intensity * 0.3; }
    else { float it = qf * 1.5f; intensity = it; streak_alpha = it * 0.3f; }
    if (centroid_out) { centroid_out[1] = light_x; centroid_out[2] = light_y; centroid_out[3] = intensity; }

    static const double ghost_color[3] = {1.0, 0.9, 0.7};
    static const double ring_colors[3][3] = {{0.3, 0.4, 1.0}, {0.5, 0.5, 0.9}, {0.7, 0.5, 0.8}};
    static const double hex_color[3] = {0.6, 0.7, 1.0};
    static const double streak_color[3] = {1.0, 0.95, 0.9};
    const double main_angles[4] = {0.0, M_PI / 2, M_PI, 3 * M_PI / 2};
    double streak_len = (double)(W < H ? W : H) * 0.4;

#pragma omp parallel for
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float fl[3] = {0.0f, 0.0f, 0.0f};
            for (int g = 0; g < 8; ++g) {
                double t = (g + 1) * 0.15;
                double gx = light_x + (scx - light_x) * t, gy = light_y + (scy - light_y) * t;
                double size = (25 + g * 30) * scale;
                double dx = x - gx, dy = y - gy;
                double dist = sqrt(dx * dx + dy * dy);
                float alpha = 0.0f;
                if (dist < size) {
                    double u = 1 - dist / size;
                    alpha = (float)(u * u * (1 - g * 0.08) * intensity);
                }
                for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + (double)alpha * ghost_color[c]);
            }
            for (int k = 0; k < 3; ++k) {
                double t = 0.35 + k * 0.15;
                double rx = light_x + (scx - light_x) * t, ry = light_y + (scy - light_y) * t;
                double rr = (60 + k * 40) * scale, rw = (6 + k * 3) * scale;
                double dx = x - rx, dy = y - ry;
                double dist = sqrt(dx * dx + dy * dy);
                double u = 1 - fabs(dist - rr) / rw;
                u = u < 0 ? 0 : (u > 1 ? 1 : u);
                double ra = u * u * 0.5 * intensity * (1 - k * 0.25);
                for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + ra * ring_colors[k][c]);
            }
            {
                double hx = light_x + (scx - light_x) * 0.5, hy = light_y + (scy - light_y) * 0.5;
                double hr = 100 * scale;
                double dx = x - hx, dy = y - hy;
                double angle = atan2(dy, dx);
                double dist = sqrt(dx * dx + dy * dy);
                double m = fmod(angle, M_PI / 3);
                if (m != 0 && m < 0) m += M_PI / 3; /* np.mod: result has the sign of the divisor */
                double edge = fabs(m - M_PI / 6);
                double hf = 1 - edge / 0.2;
                hf = hf < 0 ? 0 : (hf > 1 ? 1 : hf);
                double u = 1 - fabs(dist - hr) / (15 * scale);
                u = u < 0 ? 0 : (u > 1 ? 1 : u);
                double ra = u * u * hf * 0.3 * intensity;
                for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + ra * hex_color[c]);
            }
            {
                double dx = x - light_x, dy = y - light_y;
                double dist = sqrt(dx * dx + dy * dy);
                double angle = atan2(dy, dx);
                double falloff = exp(-dist / streak_len);
                for (int a = 0; a < 4; ++a) {
                    double m = fmod(angle - main_angles[a] + M_PI, 2 * M_PI);
                    if (m != 0 && m < 0) m += 2 * M_PI;
                    double diff = fabs(m - M_PI);
                    for (int c = 0; c < 3; ++c) {
                        double add = diff < 0.05 ? falloff * streak_alpha * streak_color[c] : 0.0;
                        fl[c] = (float)((double)fl[c] + add);
                    }
                }
            }
            float *px = final + ((size_t)y * W + x) * 3;
            for (int c = 0; c < 3; ++c) px[c] = f_clamp(px[c] + fl[c], 0.0f, 1.0f);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Simplex noise / FBM: render.py:2642-2785; permutation table render.py:2269-2288.
 * ---------------------------------------------------------------------------------------- */
static const uint8_t PERM[256] = {
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69,
    142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148, 247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219,
    203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175,
    74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230,
    220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73, 209,
    76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198,
    173, 186, 3, 64, 52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212,
    207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119, 248, 152, 2, 44,
    154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79,
    113, 224, 232, 178, 185, 112, 104, 218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12,
    191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106,
    157, 184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67,
    29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180};

static inline int perm(int i) { return PERM[i & 255]; } /* doubled table: index < 512 */

/* render.py:2642-2660: h = hash % 12, so the (h == 12 || h == 14) arm is unreachable */
static inline float grad3_dot(int hash, float x, float y, float z)
{
    int h = hash % 12;
    float u = h < 8 ? x : y;
    float v = h < 4 ? y : z;
    float r1 = (h & 1) == 0 ? u : -u;
    float r2 = (h & 2) == 0 ? v : -v;
    return r1 + r2;
}

float orc_simplex3(float x, float y, float z)
{
    const float F3 = (float)(1.0 / 3.0), G3 = (float)(1.0 / 6.0);
    float s = (x + y + z) * F3;
    int i = (int)f_floor(x + s), j = (int)f_floor(y + s), k = (int)f_floor(z + s);
    float t = (float)(i + j + k) * G3;
    float x0 = x - ((float)i - t), y0 = y - ((float)j - t), z0 = z - ((float)k - t);
    int i1, j1, k1, i2, j2, k2;
    if (x0 >= y0) {
        if (y0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
        else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
        else { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
    } else {
        if (y0 < z0) { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
        else if (x0 < z0) { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
        else { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
    }
    const float G3x2 = (float)(2.0 * (1.0 / 6.0)), G3x3 = (float)(3.0 * (1.0 / 6.0));
    float x1 = x0 - (float)i1 + G3, y1 = y0 - (float)j1 + G3, z1 = z0 - (float)k1 + G3;
    float x2 = x0 - (float)i2 + G3x2, y2 = y0 - (float)j2 + G3x2, z2 = z0 - (float)k2 + G3x2;
    float x3 = x0 - 1.0f + G3x3, y3 = y0 - 1.0f + G3x3, z3 = z0 - 1.0f + G3x3;
    int ii = i & 255, jj = j & 255, kk = k & 255;
    int gi0 = perm(ii + perm(jj + perm(kk)));
    int gi1 = perm(ii + i1 + perm(jj + j1 + perm(kk + k1)));
    int gi2 = perm(ii + i2 + perm(jj + j2 + perm(kk + k2)));
    int gi3 = perm(ii + 1 + perm(jj + 1 + perm(kk + 1)));
    float n = 0.0f;
    float t0 = 0.6f - x0 * x0 - y0 * y0 - z0 * z0;
    if (t0 >= 0.0f) { t0 = t0 * t0; n += t0 * t0 * grad3_dot(gi0, x0, y0, z0); }
    float t1 = 0.6f - x1 * x1 - y1 * y1 - z1 * z1;
    if (t1 >= 0.0f) { t1 = t1 * t1; n += t1 * t1 * grad3_dot(gi1, x1, y1, z1); }
    float t2 = 0.6f - x2 * x2 - y2 * y2 - z2 * z2;
    if (t2 >= 0.0f) { t2 = t2 * t2; n += t2 * t2 * grad3_dot(gi2, x2, y2, z2); }
    float t3 = 0.6f - x3 * x3 - y3 * y3 - z3 * z3;
    if (t3 >= 0.0f) { t3 = t3 * t3; n += t3 * t3 * grad3_dot(gi3, x3, y3, z3); }
    return 32.0f * n;
}

float orc_fbm3(float x, float y, float z, int octaves, float persistence, float lacunarity)
{
    float value = 0.0f, amplitude = 1.0f, freq = 1.0f;
    for (int o = 0; o < octaves; ++o) {
        value += amplitude * orc_simplex3(x * freq, y * freq, z * freq);
        amplitude *= persistence;
        freq *= lacunarity;
    }
    return value;
}

/* render.py:3305-3326 _noise_eval_kernel: mode 0 simplex, 1 fbm */
void orc_eval_noise(const float *coords, int n, int mode, int octaves, float persistence,
                    float lacunarity, float *out)
{
    for (int i = 0; i < n; ++i) {
        const float *c = coords + (size_t)i * 3;
        out[i] = mode == 0 ? orc_simplex3(c[0], c[1], c[2])
                           : orc_fbm3(c[0], c[1], c[2], octaves, persistence, lacunarity);
    }
}

static inline float unit_fbm(float x, float y, float z, int o, float p)
{
    return f_min(f_max(0.5f + 0.5f * orc_fbm3(x, y, z, o, p, 2.0f), 0.0f), 1.0f);
}

/* render.py:3332-3451 _generate_background_kernel: writes comp planes 0,1,2,3,4,11,12 of
 * comp (13, n_r, n_phi); planes 5..10 are left untouched. */
void orc_generate_background(float *comp, int n_r, int n_phi, int az_freq, float az_shear,
                             float r_inner, float r_outer, float t)
{
    const size_t plane = (size_t)n_r * n_phi;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ri = 0; ri < n_r; ++ri) {
        for (int pi = 0; pi < n_phi; ++pi) {
            size_t o = (size_t)ri * n_phi + pi;
            float r = (float)ri / (float)n_r;
            float phi = (float)pi / (float)n_phi * TWO_PI_F;
            float r_phys = r_inner + (r_outer - r_inner) * r;
            float omega = f_sqrt(0.5f / (r_phys * r_phys * r_phys + 1e-6f));
            float phi_rot = phi + omega * t;
            float cx = f_cos(phi_rot), cy = f_sin(phi_rot);

            float decay = f_pow(f_max(1.0f - r, 0.0f), 1.3f);
            float tb_noise = unit_fbm(cx * 8.0f, cy * 8.0f, r * 8.0f + t * 0.05f, 4, 0.6f);
            comp[0 * plane + o] = decay * (0.85f + 0.15f * tb_noise) * 0.25f;
            comp[1 * plane + o] = 0.0f;
            comp[2 * plane + o] = 0.0f;

            float t_coarse = unit_fbm(cx * 8.0f, cy * 8.0f, r * 4.0f + t * 0.06f, 3, 0.45f) * 0.08f;
            float t_mid = unit_fbm(cx * 24.0f, cy * 24.0f, r * 12.0f + t * 0.08f, 4, 0.45f) * 0.15f;
            float t_fine = unit_fbm(cx * 80.0f, cy * 80.0f, r * 40.0f + t * 0.1f, 5, 0.45f) * 0.25f;
            float t_extra = unit_fbm(cx * 200.0f, cy * 200.0f, r * 100.0f + t * 0.12f, 4, 0.4f) * 0.22f;
            float t_ultra = unit_fbm(cx * 400.0f, cy * 400.0f, r * 200.0f + t * 0.15f, 3, 0.35f) * 0.18f;
            float t_pixel = f_min(f_max(orc_simplex3(cx * 800.0f, cy * 800.0f, r * 400.0f + t * 0.2f),
                                        0.0f), 1.0f) * 0.12f;
            float turb = f_min(f_max(((((t_coarse + t_mid) + t_fine) + t_extra) + t_ultra) + t_pixel,
                                     0.0f), 1.0f);
            comp[3 * plane + o] = turb;
            comp[4 * plane + o] = 0.05f * turb;

            float shear = f_pow(r, 1.2f) * az_shear;
            float az_wave = 0.5f + 0.5f * f_sin((phi_rot + shear) * (float)az_freq);
            float az_n = unit_fbm(cx * 3.0f, cy * 3.0f, r * 3.0f + t * 0.04f, 3, 0.5f);
            comp[11 * plane + o] = az_wave * az_n;

            float d_coarse = unit_fbm(cx * 8.0f, cy * 8.0f, r * 4.0f + t * 0.003f, 3, 0.5f) * 0.05f;
            float d_mid = unit_fbm(cx * 32.0f, cy * 32.0f, r * 16.0f + t * 0.005f, 3, 0.5f) * 0.15f;
            float d_fine = unit_fbm(cx * 100.0f, cy * 100.0f, r * 50.0f + t * 0.006f, 4, 0.45f) * 0.30f;
            float d_extra = unit_fbm(cx * 250.0f, cy * 250.0f, r * 125.0f + t * 0.008f, 4, 0.4f) * 0.30f;
            float d_pixel = f_min(f_max(orc_simplex3(cx * 500.0f, cy * 500.0f, r * 250.0f + t * 0.01f),
                                        0.0f), 1.0f) * 0.20f;
            float raw = ((((d_coarse + d_mid) + d_fine) + d_extra) + d_pixel) * 1.4f;
            raw = f_min(f_max(raw, 0.05f), 1.0f);
            float preserve = 0.6f + 0.4f * r;
            comp[12 * plane + o] = f_min(f_max(raw * preserve, 0.1f), 1.0f);
        }
    }
}

/* render.py:3169-3257 _compose_disk_texture_kernel: comp (13, n_r, n_phi) -> tex (n_r, n_phi, 4).
 * stats = {density_p98, struct_scale}; row_stats (n_r, 2) = {max_r, p70_r}. */
void orc_compose_texture(const float *comp, const float *omega, const float *edge,
                         const float *stats, const float *row_stats, int n_r, int n_phi,
                         float t_offset, int enable_rt, float color_temp, float *tex)
{
    const size_t plane = (size_t)n_r * n_phi;
    const float p98 = stats[0], sscale = stats[1];
    const float t_factor = (color_temp - 4500.0f) / (6500.0f - 2700.0f);
    const float T_min = 2000.0f + t_factor * 1000.0f, T_max = 9000.0f + t_factor * 3000.0f;
    const float rt_w = enable_rt == 0 ? 0.0f : 0.20f;
#pragma omp parallel for
    for (int ri = 0; ri < n_r; ++ri) {
        for (int pi = 0; pi < n_phi; ++pi) {
            int shift = (int)(t_offset * omega[ri] / TWO_PI_F * (float)n_phi);
            int src = py_mod(pi + shift, n_phi);
            size_t o = (size_t)ri * n_phi + src;
            float tb = comp[0 * plane + o], sp = comp[1 * plane + o], sp_t = comp[2 * plane + o];
            float turb = comp[3 * plane + o], turb_t = comp[4 * plane + o];
            float arc = comp[5 * plane + o], arc_t = comp[6 * plane + o];
            float rt = comp[7 * plane + o], rt_t = comp[8 * plane + o];
            float hs = comp[9 * plane + o], hs_t = comp[10 * plane + o];
            float az = comp[11 * plane + o], dm = comp[12 * plane + o];
            float density = (((((0.15f + 0.10f * sp) + 0.30f * turb) + 0.20f * hs) + 0.30f * arc)
                             + rt_w * rt) * dm * edge[ri];
            density = f_min(f_max(density / (p98 + 1e-6f), 0.0f), 1.0f);
            float ts = ((((sp_t + turb_t) + arc_t) + rt_t) + hs_t) * dm;
            float ts_scaled = f_min(f_max(ts / (sscale + 1e-6f) * 0.8f, 0.0f), 1.2f);
            float max_r = row_stats[ri * 2 + 0], p70_r = row_stats[ri * 2 + 1];
            float ceiling = f_max(p70_r, 0.05f);
            float tbc = f_min(f_min(tb, ceiling), max_r);
            float temperature = f_min(f_max(f_max(tbc, ts_scaled), 0.0f), 1.0f);
            float ta = f_min(f_max(temperature * (0.9f + 0.25f * az), 0.0f), 1.0f);
            float T_K = T_min + ta * (T_max - T_min);
            v3 bb = color_temp_to_tint(T_K);
            float bb_b = f_min(bb.z, bb.x);
            float lum = f_min(f_max(f_sqrt(ta), 0.0f), 1.0f);
            float *dst = tex + ((size_t)ri * n_phi + pi) * 4;
            dst[0] = f_min(f_max(bb.x * lum, 0.0f), 1.0f);
            dst[1] = f_min(f_max(bb.y * lum, 0.0f), 1.0f);
            dst[2] = f_min(f_max(bb_b * lum, 0.0f), 1.0f);
            dst[3] = density;
        }
    }
}

/* render.py:3261-3281 + driver 3761-3767: mips (levels, n_r, n_phi, 4), level l in the top-left
 * (n_r >> l, n_phi >> l) corner; level 0 = copy of base.  Also equals generate_disk_mipmaps
 * (render.py:1113-1125) padded as in render.py:2246-2251 when `zero_pad` is set. */
void orc_build_mips(const float *base, int n_r, int n_phi, int levels, int zero_pad,
                    int numpy_order, float *mips)
{
    const size_t plane = (size_t)n_r * n_phi * 4;
    if (zero_pad) memset(mips, 0, plane * levels * sizeof(float));
    memcpy(mips, base, plane * sizeof(float));
    int h = n_r, w = n_phi;
    for (int lev = 1; lev < levels; ++lev) {
        int dh = h / 2, dw = w / 2;
        const float *src = mips + (size_t)(lev - 1) * plane;
        float *dst = mips + (size_t)lev * plane;
        for (int r = 0; r < dh; ++r)
            for (int c = 0; c < dw; ++c)
                for (int k = 0; k < 4; ++k) {
                    float a = src[((size_t)(2 * r) * n_phi + 2 * c) * 4 + k];
                    float b = src[((size_t)(2 * r) * n_phi + 2 * c + 1) * 4 + k];
                    float cc = src[((size_t)(2 * r + 1) * n_phi + 2 * c) * 4 + k];
                    float d = src[((size_t)(2 * r + 1) * n_phi + 2 * c + 1) * 4 + k];
                    /* kernel: (0,0)+(0,1)+(1,0)+(1,1); numpy generate_disk_mipmaps:
                     * (0,0)+(1,0)+(0,1)+(1,1) (render.py:1122-1123) */
                    dst[((size_t)r * n_phi + c) * 4 + k] =
                        numpy_order ? (((a + cc) + b) + d) / 4.0f : (((a + b) + cc) + d) / 4.0f;
                }
        h = dh; w = dw;
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
