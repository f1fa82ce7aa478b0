// texture.cu -- skybox / disk-texture uploads, mip pyramid, and the on-device disk-texture
// pipeline: simplex-FBM background layer, entity layer, compose, noise test hook.
//
// Replaces (reference = render.py):
//   generate_disk_mipmaps + padded upload (1113-1125, 2239-2251) and the mip kernels (3261-3281)
//   _simplex_noise_3d / _fbm_3d / _generate_background_kernel (2642-2785, 3332-3451)
//   accumulate_entity_layer's numpy loops + staging copy kernel (3564-3653, 3455-3471)
//   _compose_disk_texture_kernel (3169-3257), _noise_eval_kernel (3305-3326)
#include <math.h>

#include "common.cuh"
#include "noise.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// uploads
// ---------------------------------------------------------------------------------------------
__global__ void pack_rgb_kernel(const float* __restrict__ rgb, float4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.0f);
}

// one mip level from the previous one: 2x2 mean.  numpy_order selects the summation order of
// generate_disk_mipmaps ((0,0)+(1,0)+(0,1)+(1,1), render.py:1122-1123) instead of the mip
// kernel's ((0,0)+(0,1)+(1,0)+(1,1), render.py:3277-3280).
__global__ void mip_down_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int dh, int dw,
                                int src_pitch, int numpy_order) {
    int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= dw || r >= dh) return;
    float4 a = src[(size_t)(2 * r) * src_pitch + 2 * c], b = src[(size_t)(2 * r) * src_pitch + 2 * c + 1];
    float4 cc = src[(size_t)(2 * r + 1) * src_pitch + 2 * c], d = src[(size_t)(2 * r + 1) * src_pitch + 2 * c + 1];
    float4 o;
    if (numpy_order) {
        o.x = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.x, cc.x), b.x), d.x), 4.0f);
        o.y = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.y, cc.y), b.y), d.y), 4.0f);
        o.z = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.z, cc.z), b.z), d.z), 4.0f);
        o.w = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.w, cc.w), b.w), d.w), 4.0f);
    } else {
        o.x = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.x, b.x), cc.x), d.x), 4.0f);
        o.y = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.y, b.y), cc.y), d.y), 4.0f);
        o.z = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.z, b.z), cc.z), d.z), 4.0f);
        o.w = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.w, b.w), cc.w), d.w), 4.0f);
    }
    dst[(size_t)r * dw + c] = o;
}

__global__ void __launch_bounds__(256) noise_eval_kernel(const float* __restrict__ coords, int n, int mode, int octaves,
                                                         float persistence, float lacunarity, float* __restrict__ out) {
    __shared__ __align__(16) unsigned char sperm[kPermSmem];
    Perm perm;
    load_perm(sperm, perm);
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
    out[i] = mode == 0 ? simplex3(perm, x, y, z) : fbm3(perm, x, y, z, octaves, persistence, lacunarity);
}

// render.py:3332-3451; writes comp planes 0,1,2,3,4,11,12.  42 simplex evaluations per texel.
// The cos/sin of the Keplerian-rotated angle feed noise coordinates scaled by up to 800, so
// they are evaluated in double and rounded once (the oracle's ideal-libm convention).
__global__ void __launch_bounds__(256) background_scalar_kernel(float* __restrict__ comp, int n_r, int n_phi, int az_freq,
                                                         float az_shear, float r_inner, float r_outer, float t) {
    __shared__ __align__(16) unsigned char sperm[kPermSmem];
    Perm perm;
    load_perm(sperm, perm);
    const size_t plane = (size_t)n_r * n_phi;
    // block = 256 columns of row blockIdx.y; the quantities that depend on the row only (two
    // double-precision pow, omega) are evaluated by one thread and shared
    __shared__ float row_decay, row_shear, row_omega;
    const int ri = blockIdx.y, pi = blockIdx.x * blockDim.x + threadIdx.x;
    const float r = __fdiv_rn((float)ri, (float)n_r);
    if (threadIdx.x == 0) {
        const float r_phys = a_(r_inner, m_(__fsub_rn(r_outer, r_inner), r));
        row_omega = __fsqrt_rn(__fdiv_rn(0.5f, a_(m_(m_(r_phys, r_phys), r_phys), 1e-6f)));
        row_decay = (float)pow((double)fmaxf(__fsub_rn(1.0f, r), 0.0f), (double)1.3f);
        row_shear = m_((float)pow((double)r, (double)1.2f), az_shear);
    }
    __syncthreads();
    if (pi >= n_phi) return;
    const size_t o = (size_t)ri * n_phi + pi;
    const float phi = m_(__fdiv_rn((float)pi, (float)n_phi), 6.2831855f);
    const float omega = row_omega;
    const float phi_rot = a_(phi, m_(omega, t));
    const float cx = (float)cos((double)phi_rot), cy = (float)sin((double)phi_rot);

    const float decay = row_decay;
    float tb = unit_fbm(perm, m_(cx, 8.0f), m_(cy, 8.0f), a_(m_(r, 8.0f), m_(t, 0.05f)), 4, 0.6f);
    comp[0 * plane + o] = m_(m_(decay, a_(0.85f, m_(0.15f, tb))), 0.25f);
    comp[1 * plane + o] = 0.0f;
    comp[2 * plane + o] = 0.0f;

    float t_coarse = m_(unit_fbm(perm, m_(cx, 8.0f), m_(cy, 8.0f), a_(m_(r, 4.0f), m_(t, 0.06f)), 3, 0.45f), 0.08f);
    float t_mid = m_(unit_fbm(perm, m_(cx, 24.0f), m_(cy, 24.0f), a_(m_(r, 12.0f), m_(t, 0.08f)), 4, 0.45f), 0.15f);
    float t_fine = m_(unit_fbm(perm, m_(cx, 80.0f), m_(cy, 80.0f), a_(m_(r, 40.0f), m_(t, 0.1f)), 5, 0.45f), 0.25f);
    float t_extra = m_(unit_fbm(perm, m_(cx, 200.0f), m_(cy, 200.0f), a_(m_(r, 100.0f), m_(t, 0.12f)), 4, 0.4f), 0.22f);
    float t_ultra = m_(unit_fbm(perm, m_(cx, 400.0f), m_(cy, 400.0f), a_(m_(r, 200.0f), m_(t, 0.15f)), 3, 0.35f), 0.18f);
    float t_pixel = m_(fminf(fmaxf(simplex3(perm, m_(cx, 800.0f), m_(cy, 800.0f), a_(m_(r, 400.0f), m_(t, 0.2f))), 0.0f), 1.0f), 0.12f);
    float turb = fminf(fmaxf(a_(a_(a_(a_(a_(t_coarse, t_mid), t_fine), t_extra), t_ultra), t_pixel), 0.0f), 1.0f);
    comp[3 * plane + o] = turb;
    comp[4 * plane + o] = m_(0.05f, turb);

    const float shear = row_shear;
    float az_wave = a_(0.5f, m_(0.5f, (float)sin((double)m_(a_(phi_rot, shear), (float)az_freq))));
    float az_n = unit_fbm(perm, m_(cx, 3.0f), m_(cy, 3.0f), a_(m_(r, 3.0f), m_(t, 0.04f)), 3, 0.5f);
    comp[11 * plane + o] = m_(az_wave, az_n);

    float d_coarse = m_(unit_fbm(perm, m_(cx, 8.0f), m_(cy, 8.0f), a_(m_(r, 4.0f), m_(t, 0.003f)), 3, 0.5f), 0.05f);
    float d_mid = m_(unit_fbm(perm, m_(cx, 32.0f), m_(cy, 32.0f), a_(m_(r, 16.0f), m_(t, 0.005f)), 3, 0.5f), 0.15f);
    float d_fine = m_(unit_fbm(perm, m_(cx, 100.0f), m_(cy, 100.0f), a_(m_(r, 50.0f), m_(t, 0.006f)), 4, 0.45f), 0.30f);
    float d_extra = m_(unit_fbm(perm, m_(cx, 250.0f), m_(cy, 250.0f), a_(m_(r, 125.0f), m_(t, 0.008f)), 4, 0.4f), 0.30f);
    float d_pixel = m_(fminf(fmaxf(simplex3(perm, m_(cx, 500.0f), m_(cy, 500.0f), a_(m_(r, 250.0f), m_(t, 0.01f))), 0.0f), 1.0f), 0.20f);
    float raw = m_(a_(a_(a_(a_(d_coarse, d_mid), d_fine), d_extra), d_pixel), 1.4f);
    raw = fminf(fmaxf(raw, 0.05f), 1.0f);
    float preserve = a_(0.6f, m_(0.4f, r));
    comp[12 * plane + o] = fminf(fmaxf(m_(raw, preserve), 0.1f), 1.0f);
}

// ---------------------------------------------------------------------------------------------
// entity layer, render.py:3564-3653: one thread per texel sums, in list order, every entity
// whose row range contains the texel's row (so the f32 accumulation order equals numpy's).
//   filament p[]: 0 source_phi, 1 base_r, 2 inv_2sigma_r_sq, 3 inv_2sigma_phi_sq, 4 scale_d, 5 scale_t
//   hotspot  p[]: 0 h_phi, 1 h_r, 2 h_phi_width, 3 h_r_width, 4 h_intensity        (temp = 0.12 * density)
//   rt_spike p[]: 0 rt_phi, 1 rt_r_base, 2 rt_phi_width, 3 rt_r_length, 4 rt_intensity, 5 rt_delta_T
// planes: filament -> comp[5], comp[6]; rt_spike -> comp[7], comp[8]; hotspot -> comp[9], comp[10]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float py_modf32(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.0f && ((m < 0.0f) != (b < 0.0f))) m += b;
    return m;
}

// Hotspot / rt_spike azimuthal profiles exp(kappa (cos(phi - phi0) - 1)) depend on the (rolled)
// source column only: one table row of n_phi doubles per such entity.
__global__ void __launch_bounds__(256) entity_coltab_kernel(const bhr_entity* __restrict__ ents, const int* __restrict__ slot_ent,
                                                            int n_phi, double* __restrict__ coltab) {
    const int src = blockIdx.x * 256 + threadIdx.x;
    if (src >= n_phi) return;
    const bhr_entity& E = ents[slot_ent[blockIdx.y]];
    const double phi = (double)src * (6.283185307179586 / (double)n_phi);
    const double kappa = 1.5 / (E.p[2] * E.p[2]);
    coltab[(size_t)blockIdx.y * n_phi + src] = exp(kappa * (cos(phi - E.p[0]) - 1.0));
}

// one entity that touches the block's row, with everything that depends on (entity, row) only
struct RowEntity {
    double a, b, c;          // filament: scale_d * r_w, scale_t * r_w, 1 / (2 sigma_phi^2); others: r_prof, intensity, -
    const double* col;       // hotspot / rt_spike: azimuthal profile table
    const float *tab_d, *tab_t;   // kind 3 / 4: this row of the caller's tabulated density / temperature profiles
    float ctr;               // filament: blob centre (float32, render.py:3633); others: fade alpha
    float tfac;              // rt_spike: delta_T
    int kind, shift;
};

constexpr int kEntCols = 1024;     // columns per block (4 per thread)

__global__ void __launch_bounds__(256) entity_accumulate_kernel(float* __restrict__ comp, int n_r, int n_phi,
                                                                const bhr_entity* __restrict__ ents, int n_ent,
                                                                const float* __restrict__ omega_rows,
                                                                const int* __restrict__ ent_slot,
                                                                const double* __restrict__ coltab,
                                                                const float* __restrict__ tables) {
    extern __shared__ __align__(16) unsigned char ent_smem[];
    RowEntity* list = reinterpret_cast<RowEntity*>(ent_smem);
    __shared__ int warp_count[8];
    __shared__ int n_list;
    const int ri = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)n_r * n_phi;
    const double TWO_PI = 6.283185307179586, PI = 3.141592653589793;
    const double r_norm = (ri == n_r - 1) ? 1.0 : (double)ri * (1.0 / (double)(n_r - 1));   // np.linspace(0, 1, n_r)
    const double phi_step = TWO_PI / (double)n_phi;                                           // linspace(endpoint=False)
    const float om = omega_rows[ri];
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    // ---- the entities of this row, in list order (ordered compaction, 256 candidates at a time) ----
    for (int e0 = 0; e0 < n_ent; e0 += 256) {
        const int e = e0 + threadIdx.x;
        const bool hit = e < n_ent && ri >= ents[e].row_begin && ri < ents[e].row_end;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_count[warp] = __popc(m);
        __syncthreads();
        int base = n_list;
        for (int w = 0; w < warp; ++w) base += warp_count[w];
        if (hit) {
            const bhr_entity& E = ents[e];
            RowEntity R;
            R.kind = E.kind; R.col = nullptr; R.shift = 0; R.tfac = 0.0f; R.c = 0.0;
            const double rd = r_norm - E.p[1];
            R.tab_d = R.tab_t = nullptr;
            if (E.kind >= 3) {
                // a reference EntityInstance with tabulated (rows, n_phi) float32 profiles (render.py:3640-3649):
                // p[0] = offset of its density rows in `tables`, its temperature rows follow them
                const size_t rows = (size_t)(E.row_end - E.row_begin);
                R.tab_d = tables + (size_t)E.p[0] + (size_t)(ri - E.row_begin) * n_phi;
                R.tab_t = R.tab_d + rows * n_phi;
                R.shift = (int)__fmul_rn(__fdiv_rn(__fmul_rn((float)E.age, om), 6.2831855f), (float)n_phi);
                R.ctr = (float)E.scale;
            } else if (E.kind == 0) {
                const double r_w = exp(-(rd * rd) * E.p[2]);
                R.a = E.p[4] * r_w; R.b = E.p[5] * r_w; R.c = E.p[3];
                // center = (source_phi - omega[ri] * age) % 2pi evaluates in float32 in the reference
                // (np.float32 scalar with weak Python floats), render.py:3633
                R.ctr = py_modf32(__fsub_rn((float)E.p[0], __fmul_rn(om, (float)E.age)), 6.2831855f);
            } else {
                // shift = int(age * omega[ri] / (2 pi) * n_phi) in float32, render.py:3645
                R.shift = (int)__fmul_rn(__fdiv_rn(__fmul_rn((float)E.age, om), 6.2831855f), (float)n_phi);
                R.col = coltab + (size_t)ent_slot[e] * n_phi;
                R.ctr = (float)E.scale;
                R.b = E.p[4];
                if (E.kind == 1) {
                    const double q = rd / (E.p[3] + 1e-8);
                    R.a = exp(-0.5 * (q * q));
                } else {
                    const double fo = fmin(fmax(E.p[3] * 2 - rd, 0.0), 1.0);
                    const double fi = fmin(fmax(rd / (E.p[3] * 0.3 + 1e-8), 0.0), 1.0);
                    const double q = rd / (E.p[3] * 0.4 + 1e-8);
                    R.a = exp(-0.5 * (q * q)) * fo * fi;
                    R.tfac = (float)E.p[5];
                }
            }
            list[base + __popc(m & ((1u << lane) - 1u))] = R;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = n_list; for (int w = 0; w < 8; ++w) t += warp_count[w]; n_list = t; }
        __syncthreads();
    }
    const int n = n_list;
    // ---- accumulate: f32 sums in list order, float64 profiles (numpy's order and precision) ----
    for (int pi = blockIdx.x * kEntCols + threadIdx.x; pi < min(n_phi, (int)(blockIdx.x + 1) * kEntCols); pi += 256) {
        float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const double phi_tex = (double)pi * phi_step;
        for (int k = 0; k < n; ++k) {
            const RowEntity& R = list[k];
            if (R.kind == 0) {
                double d_phi = phi_tex - (double)R.ctr;
                // d_phi - 2 pi rint(d_phi / (2 pi)) for |d_phi| < 3 pi: the quotient rounds above 0.5
                // exactly when d_phi > pi (pi = 2 pi / 2 in double), so two comparisons replace the division
                if (d_phi > PI) d_phi -= TWO_PI; else if (d_phi < -PI) d_phi += TWO_PI;
                const double prof = exp(-d_phi * d_phi * R.c);
                acc[0] = (float)((double)acc[0] + prof * R.a);
                acc[1] = (float)((double)acc[1] + prof * R.b);
            } else if (R.kind >= 3) {
                // staging[d, ri] += np.roll(phi_density[k], -shift) * alpha: float32 product, float32 sum
                int src = (pi + R.shift) % n_phi;
                if (src < 0) src += n_phi;
                const int d = R.kind == 3 ? 4 : 2;              // hotspot -> comp[9], comp[10]; rt_spike -> comp[7], comp[8]
                acc[d] = __fadd_rn(acc[d], __fmul_rn(R.tab_d[src], R.ctr));
                acc[d + 1] = __fadd_rn(acc[d + 1], __fmul_rn(R.tab_t[src], R.ctr));
            } else {
                int src = (pi + R.shift) % n_phi;
                if (src < 0) src += n_phi;
                const double phi_prof = R.col[src];
                float dens = (float)(phi_prof * R.a * R.b);
                if (R.kind == 1) {
                    float temp = __fmul_rn(dens, 0.12f);
                    dens = fminf(fmaxf(dens, 0.0f), 1.0f);
                    temp = fminf(fmaxf(temp, 0.0f), 1.0f);
                    acc[4] = __fadd_rn(acc[4], __fmul_rn(dens, R.ctr));
                    acc[5] = __fadd_rn(acc[5], __fmul_rn(temp, R.ctr));
                } else {
                    const float temp = __fmul_rn(dens, R.tfac);
                    dens = fminf(fmaxf(dens, 0.0f), 1.0f);
                    acc[2] = __fadd_rn(acc[2], __fmul_rn(dens, R.ctr));
                    acc[3] = __fadd_rn(acc[3], __fmul_rn(temp, R.ctr));
                }
            }
        }
        const size_t o = (size_t)ri * n_phi + pi;
#pragma unroll
        for (int k = 0; k < 6; ++k) comp[(5 + k) * plane + o] = acc[k];
    }
}

// ---------------------------------------------------------------------------------------------
// compose, render.py:3169-3257 (+ _color_temp_to_tint, 2407-2437); writes mip level 0 directly
// ---------------------------------------------------------------------------------------------
__device__ void color_temp_to_tint(float temp, float& r, float& g, float& b) {
    float t = __fdiv_rn(temp, 100.0f);
    r = 1.0f; b = 1.0f;
    if (t > 66.0f) r = fminf(fmaxf(m_(1.292936f, (float)pow((double)fmaxf(__fsub_rn(t, 60.0f), 0.0001f), (double)-0.1332047592f)), 0.0f), 1.0f);
    if (t <= 66.0f) g = fminf(fmaxf(__fsub_rn(m_(0.390082f, (float)log((double)fmaxf(t, 0.0001f))), 0.631841f), 0.0f), 1.0f);
    else g = fminf(fmaxf(m_(1.129891f, (float)pow((double)fmaxf(__fsub_rn(t, 60.0f), 0.0001f), (double)-0.0755148492f)), 0.0f), 1.0f);
    if (t < 66.0f) {
        if (t <= 19.0f) b = 0.0f;
        else b = fminf(fmaxf(__fsub_rn(m_(0.543207f, (float)log((double)fmaxf(__fsub_rn(t, 10.0f), 0.0001f))), 1.19625f), 0.0f), 1.0f);
    }
}

__global__ void __launch_bounds__(256) compose_kernel(const float* __restrict__ comp, const float* __restrict__ omega,
                                                      const float* __restrict__ edge, float p98, float sscale,
                                                      const float* __restrict__ row_stats, int n_r, int n_phi,
                                                      float t_offset, int enable_rt, float color_temp,
                                                      float4* __restrict__ tex) {
    const size_t plane = (size_t)n_r * n_phi;
    const size_t idx = blockIdx.x * (size_t)256 + threadIdx.x;
    if (idx >= plane) return;
    const int ri = (int)(idx / n_phi), pi = (int)(idx % n_phi);
    const float t_factor = __fdiv_rn(__fsub_rn(color_temp, 4500.0f), 3800.0f);
    const float T_min = a_(2000.0f, m_(t_factor, 1000.0f)), T_max = a_(9000.0f, m_(t_factor, 3000.0f));
    const float rt_w = enable_rt == 0 ? 0.0f : 0.20f;
    int shift = (int)m_(__fdiv_rn(m_(t_offset, omega[ri]), 6.2831855f), (float)n_phi);
    int src = (pi + shift) % n_phi;
    if (src < 0) src += n_phi;
    const size_t o = (size_t)ri * n_phi + src;
    float tb = comp[o], sp = comp[plane + o], sp_t = comp[2 * plane + o], turb = comp[3 * plane + o];
    float turb_t = comp[4 * plane + o], arc = comp[5 * plane + o], arc_t = comp[6 * plane + o];
    float rt = comp[7 * plane + o], rt_t = comp[8 * plane + o], hs = comp[9 * plane + o], hs_t = comp[10 * plane + o];
    float az = comp[11 * plane + o], dm = comp[12 * plane + o];
    float density = m_(m_(a_(a_(a_(a_(a_(0.15f, m_(0.10f, sp)), m_(0.30f, turb)), m_(0.20f, hs)), m_(0.30f, arc)), m_(rt_w, rt)), dm), edge[ri]);
    density = fminf(fmaxf(__fdiv_rn(density, a_(p98, 1e-6f)), 0.0f), 1.0f);
    float ts = m_(a_(a_(a_(a_(sp_t, turb_t), arc_t), rt_t), hs_t), dm);
    float ts_scaled = fminf(fmaxf(m_(__fdiv_rn(ts, a_(sscale, 1e-6f)), 0.8f), 0.0f), 1.2f);
    float max_r = row_stats[2 * ri], p70_r = row_stats[2 * ri + 1];
    float tbc = fminf(fminf(tb, fmaxf(p70_r, 0.05f)), max_r);
    float temperature = fminf(fmaxf(fmaxf(tbc, ts_scaled), 0.0f), 1.0f);
    float ta = fminf(fmaxf(m_(temperature, a_(0.9f, m_(0.25f, az))), 0.0f), 1.0f);
    float T_K = a_(T_min, m_(ta, __fsub_rn(T_max, T_min)));
    float br, bg, bb;
    color_temp_to_tint(T_K, br, bg, bb);
    float bb_b = fminf(bb, br);
    float lum = fminf(fmaxf(__fsqrt_rn(ta), 0.0f), 1.0f);
    tex[idx] = make_float4(fminf(fmaxf(m_(br, lum), 0.0f), 1.0f), fminf(fmaxf(m_(bg, lum), 0.0f), 1.0f),
                           fminf(fmaxf(m_(bb_b, lum), 0.0f), 1.0f), density);
}

}  // namespace

int bhr_launch_build_mips(bhr_ctx* ctx, int numpy_order) {
    int h = ctx->n_r, w = ctx->n_phi;
    for (int lev = 1; lev < BHR_NUM_MIPS; ++lev) {
        int dh = h / 2, dw = w / 2;
        dim3 block(32, 8), grid(bhr_div_up(dw, 32), bhr_div_up(dh, 8));
        mip_down_kernel<<<grid, block, 0, ctx->stream>>>(ctx->mips + ctx->level_off[lev - 1], ctx->mips + ctx->level_off[lev],
                                                         dh, dw, w, numpy_order);
        ++ctx->launches;
        h = dh; w = dw;
    }
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

extern "C" int bhr_upload_skybox(bhr_ctx* ctx, const float* rgb, int h, int w) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !rgb || h <= 0 || w <= 0) return BHR_ERR_INVALID;
    size_t n = (size_t)h * w;
    float* staging = nullptr;
    if (ctx->sky) { cudaFree(ctx->sky); ctx->sky = nullptr; }
    BHR_CUDA(ctx, cudaMalloc(&ctx->sky, n * sizeof(float4)));
    BHR_CUDA(ctx, cudaMalloc(&staging, n * 3 * sizeof(float)));
    BHR_CUDA(ctx, cudaMemcpyAsync(staging, rgb, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    pack_rgb_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(staging, ctx->sky, n);
    BHR_CUDA(ctx, cudaGetLastError());
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(staging);
    ctx->sky_w = w; ctx->sky_h = h;
    return BHR_OK;
}

static int ensure_disk_storage(bhr_ctx* ctx, int n_r, int n_phi) {
    if (ctx->mips) {
        if (n_r != ctx->n_r || n_phi != ctx->n_phi)
            BHR_FAIL(ctx, BHR_ERR_INVALID, "Texture size mismatch: expected %dx%d, got %dx%d", ctx->n_r, ctx->n_phi, n_r, n_phi);
        return BHR_OK;
    }
    if (n_r < 16 || n_phi < 16) BHR_FAIL(ctx, BHR_ERR_INVALID, "disk texture must be at least 16x16 (5 mip levels)");
    ctx->n_r = n_r; ctx->n_phi = n_phi;
    unsigned int off = 0;
    int h = n_r, w = n_phi;
    for (int lev = 0; lev < BHR_NUM_MIPS; ++lev) { ctx->level_off[lev] = off; off += (unsigned)(h * w); h /= 2; w /= 2; }
    ctx->level_off[BHR_NUM_MIPS] = off;
    BHR_CUDA(ctx, cudaMalloc(&ctx->mips, (size_t)off * sizeof(float4)));
    BHR_CUDA(ctx, cudaMemsetAsync(ctx->mips, 0, (size_t)off * sizeof(float4), ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_upload_disk_texture(bhr_ctx* ctx, const float* rgba, int n_r, int n_phi) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !rgba) return BHR_ERR_INVALID;
    int rc = ensure_disk_storage(ctx, n_r, n_phi);
    if (rc) return rc;
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->mips, rgba, (size_t)n_r * n_phi * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    rc = bhr_launch_build_mips(ctx, 1);
    if (rc) return rc;
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_init_background(bhr_ctx* ctx, int n_r, int n_phi, int az_freq, float az_shear, const float* edge,
                                   const float* omega_rows) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !edge || !omega_rows) return BHR_ERR_INVALID;
    int rc = ensure_disk_storage(ctx, n_r, n_phi);
    if (rc) return rc;
    const size_t plane = (size_t)n_r * n_phi;
    if (!ctx->comp) {
        BHR_CUDA(ctx, cudaMalloc(&ctx->comp, plane * BHR_N_COMP * sizeof(float)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->edge, n_r * sizeof(float)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->omega_rows, n_r * sizeof(float)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->row_stats, 2 * n_r * sizeof(float)));
    }
    if ((rc = bhr_join_entities(ctx))) return rc;            // (an entity layer still in flight on its own stream)
    BHR_CUDA(ctx, cudaMemsetAsync(ctx->comp, 0, plane * BHR_N_COMP * sizeof(float), ctx->stream));
    if ((rc = bhr_mark_comp_read(ctx))) return rc;           // (... and the next one starts behind this clear)
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->edge, edge, n_r * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->omega_rows, omega_rows, n_r * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->az_freq = az_freq; ctx->az_shear = az_shear;
    {
        int rc = bhr_setup_background(ctx);          // row tables of the packed kernel (background.cu)
        if (rc) return rc;
    }
    ctx->bg_ready = 1;
    return BHR_OK;
}

extern "C" int bhr_generate_background(bhr_ctx* ctx, float t) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return BHR_ERR_INVALID;
    if (!ctx->bg_ready) BHR_FAIL(ctx, BHR_ERR_STATE, "Must call init_background_layer() first");
    if (ctx->entity_stream_on) {
        // the entity layer of the same frame (bhr_accumulate_entities, its own stream) may start with this kernel: the two
        // write different planes of `comp`, and everything that read the old planes is ahead of this point on the stream
        if (!ctx->bg_start_ev) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bg_start_ev, cudaEventDisableTiming));
        BHR_CUDA(ctx, cudaEventRecord(ctx->bg_start_ev, ctx->stream));
        ctx->bg_start_armed = 1;
    }
    if (!ctx->background_scalar) return bhr_launch_background(ctx, t);       // background.cu: two texels per thread, packed
    // one texel per thread, row quantities per block (option "background_scalar": A/B reference of the packed kernel);
    // one block per (row, column chunk), the block width that wastes the fewest lanes on the ragged last chunk
    int best = 256, best_waste = 1 << 30;
    for (int b = 256; b >= 128; b -= 32) {
        const int waste = bhr_div_up(ctx->n_phi, b) * b - ctx->n_phi;
        if (waste < best_waste) { best_waste = waste; best = b; }
    }
    background_scalar_kernel<<<dim3(bhr_div_up(ctx->n_phi, best), ctx->n_r), best, 0, ctx->stream>>>(
        ctx->comp, ctx->n_r, ctx->n_phi, ctx->az_freq, ctx->az_shear, ctx->cfg.r_disk_inner, ctx->cfg.r_disk_outer, t);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

extern "C" int bhr_accumulate_entities(bhr_ctx* ctx, const bhr_entity* entities, int n) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || (n > 0 && !entities)) return BHR_ERR_INVALID;
    if (!ctx->bg_ready) BHR_FAIL(ctx, BHR_ERR_STATE, "Must call init_background_layer() first");
    // host staging ring (pinned): the caller's array may be reused as soon as we return, and the
    // upload must not wait for the frames still in flight on the stream
    const int n_slots_ring = BHR_FRAME_SLOTS;
    // Stream: the entity kernels are FP64-bound, the background kernel FP32-bound; on their own stream, released by the
    // event bhr_generate_background records in front of its kernel (or, without one, by everything enqueued so far),
    // they share the SMs with it.  Readers of the entity planes join through bhr_join_entities.
    cudaStream_t es = ctx->stream;
    if (ctx->entity_stream_on) {
        if (!ctx->ent_stream) {
            BHR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->ent_stream, cudaStreamNonBlocking));
            BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ent_done_ev, cudaEventDisableTiming));
        }
        es = ctx->ent_stream;
        if (ctx->entity_early && ctx->comp_read_valid) {
            // as early as the data allow: behind the last reader of the planes (the previous frame's compose / statistics),
            // i.e. beside the previous frame's bloom passes as well as beside this frame's background kernel
            BHR_CUDA(ctx, cudaStreamWaitEvent(es, ctx->comp_read_ev, 0));
            ctx->bg_start_armed = 0;
        } else {
            if (!ctx->bg_start_ev) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bg_start_ev, cudaEventDisableTiming));
            if (!ctx->bg_start_armed) BHR_CUDA(ctx, cudaEventRecord(ctx->bg_start_ev, ctx->stream));
            ctx->bg_start_armed = 0;
            BHR_CUDA(ctx, cudaStreamWaitEvent(es, ctx->bg_start_ev, 0));
        }
    }
    if (n > ctx->entities_cap) {
        BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ent_stream) BHR_CUDA(ctx, cudaStreamSynchronize(ctx->ent_stream));
        if (ctx->d_entities) cudaFree(ctx->d_entities);
        if (ctx->h_entities) cudaFreeHost(ctx->h_entities);
        if (ctx->d_coltab) cudaFree(ctx->d_coltab);
        ctx->entities_cap = n + 256;
        const size_t per = (size_t)ctx->entities_cap * (sizeof(bhr_entity) + 2 * sizeof(int));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_entities, per * n_slots_ring));
        BHR_CUDA(ctx, cudaMallocHost(&ctx->h_entities, per * n_slots_ring));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_coltab, (size_t)ctx->entities_cap * ctx->n_phi * sizeof(double) * 2));
        for (int k = 0; k < n_slots_ring; ++k)
            if (!ctx->ent_ev[k]) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ent_ev[k], cudaEventDisableTiming));
        ctx->ent_ring = 0;
    }
    for (int e = 0; e < n; ++e) {
        const bhr_entity& E = entities[e];
        if (E.kind < 0 || E.kind > 4 || E.row_begin < 0 || E.row_end > ctx->n_r || E.row_begin > E.row_end)
            BHR_FAIL(ctx, BHR_ERR_INVALID, "entity %d: bad kind / row range", e);
        if (E.kind >= 3) {
            const double need = E.p[0] + 2.0 * (double)(E.row_end - E.row_begin) * ctx->n_phi;
            if (E.p[0] < 0 || need > (double)ctx->entity_tables_n)
                BHR_FAIL(ctx, BHR_ERR_INVALID, "entity %d references profile tables that were not uploaded (bhr_upload_entity_tables)", e);
        }
    }
    const size_t plane = (size_t)ctx->n_r * ctx->n_phi;
    if (n == 0) {
        BHR_CUDA(ctx, cudaMemsetAsync(ctx->comp + 5 * plane, 0, 6 * plane * sizeof(float), es));
        if (es != ctx->stream) { BHR_CUDA(ctx, cudaEventRecord(ctx->ent_done_ev, es)); ctx->ent_pending = 1; }
        return BHR_OK;
    }
    const int ring = ctx->ent_ring;
    ctx->ent_ring = (ring + 1) % n_slots_ring;
    const size_t per = (size_t)ctx->entities_cap * (sizeof(bhr_entity) + 2 * sizeof(int));
    BHR_CUDA(ctx, cudaEventSynchronize(ctx->ent_ev[ring]));       // the kernels that last read this slot are done
    char* h = (char*)ctx->h_entities + per * ring;
    char* d = (char*)ctx->d_entities + per * ring;
    bhr_entity* h_ent = (bhr_entity*)h;
    int* h_slot = (int*)(h + (size_t)ctx->entities_cap * sizeof(bhr_entity));    // entity -> table row
    int* h_slot_ent = h_slot + ctx->entities_cap;                                // table row -> entity
    memcpy(h_ent, entities, (size_t)n * sizeof(bhr_entity));
    int n_tab = 0;
    for (int e = 0; e < n; ++e) {
        if (entities[e].kind == 1 || entities[e].kind == 2) { h_slot[e] = n_tab; h_slot_ent[n_tab++] = e; } else h_slot[e] = -1;
    }
    BHR_CUDA(ctx, cudaMemcpyAsync(d, h, per, cudaMemcpyHostToDevice, es));
    const bhr_entity* d_ent = (const bhr_entity*)d;
    const int* d_slot = (const int*)(d + (size_t)ctx->entities_cap * sizeof(bhr_entity));
    const int* d_slot_ent = d_slot + ctx->entities_cap;
    double* coltab = ctx->d_coltab + (size_t)(ring & 1) * ctx->entities_cap * ctx->n_phi;
    if (n_tab > 0) {
        dim3 g(bhr_div_up(ctx->n_phi, 256), n_tab);
        entity_coltab_kernel<<<g, 256, 0, es>>>(d_ent, d_slot_ent, ctx->n_phi, coltab);
        ++ctx->launches;
    }
    const size_t smem = (size_t)n * sizeof(RowEntity);
    if (smem > 200 * 1024) BHR_FAIL(ctx, BHR_ERR_INVALID, "too many entities (%d) for the per-row list", n);
    if (ctx->entity_stream_on && !ctx->entity_carveout_set) {      // the same carve-out as the background kernel they run beside
        BHR_CUDA(ctx, cudaFuncSetAttribute(entity_accumulate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
        BHR_CUDA(ctx, cudaFuncSetAttribute(entity_coltab_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
        ctx->entity_carveout_set = 1;
    }
    if (smem > 48 * 1024)
        BHR_CUDA(ctx, cudaFuncSetAttribute(entity_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(bhr_div_up(ctx->n_phi, kEntCols), ctx->n_r);
    entity_accumulate_kernel<<<grid, 256, smem, es>>>(ctx->comp, ctx->n_r, ctx->n_phi, d_ent, n, ctx->omega_rows,
                                                               d_slot, coltab, ctx->d_entity_tables);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    BHR_CUDA(ctx, cudaEventRecord(ctx->ent_ev[ring], es));
    if (es != ctx->stream) { BHR_CUDA(ctx, cudaEventRecord(ctx->ent_done_ev, es)); ctx->ent_pending = 1; }
    return BHR_OK;
}

// Tabulated profiles of caller-owned entities (the reference's EntityInstance.phi_density / phi_temp): one float32
// buffer, referenced by kind 3 / 4 entities through p[0].  Replaces the previous upload.
extern "C" int bhr_upload_entity_tables(bhr_ctx* ctx, const float* data, size_t n_floats) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || (n_floats > 0 && !data)) return BHR_ERR_INVALID;
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));           // kernels of earlier frames may still read the old buffer
    if (ctx->ent_stream) BHR_CUDA(ctx, cudaStreamSynchronize(ctx->ent_stream));
    if (n_floats > ctx->entity_tables_cap) {
        if (ctx->d_entity_tables) cudaFree(ctx->d_entity_tables);
        ctx->d_entity_tables = nullptr;
        ctx->entity_tables_cap = n_floats + n_floats / 2;
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_entity_tables, ctx->entity_tables_cap * sizeof(float)));
    }
    if (n_floats) BHR_CUDA(ctx, cudaMemcpyAsync(ctx->d_entity_tables, data, n_floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->entity_tables_n = n_floats;
    return BHR_OK;
}

extern "C" int bhr_set_stats(bhr_ctx* ctx, float density_p98, float struct_scale, const float* row_stats) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !row_stats) return BHR_ERR_INVALID;
    if (!ctx->bg_ready) BHR_FAIL(ctx, BHR_ERR_STATE, "Must call init_background_layer() first");
    ctx->stats[0] = density_p98; ctx->stats[1] = struct_scale;
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->row_stats, row_stats, 2 * ctx->n_r * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_upload_comp(bhr_ctx* ctx, const float* comp) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !comp) return BHR_ERR_INVALID;
    if (!ctx->bg_ready) BHR_FAIL(ctx, BHR_ERR_STATE, "Must call init_background_layer() first");
    const size_t bytes = (size_t)ctx->n_r * ctx->n_phi * BHR_N_COMP * sizeof(float);
    if (int rc = bhr_join_entities(ctx)) return rc;
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->comp, comp, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = bhr_mark_comp_read(ctx)) return rc;
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_compose_texture(bhr_ctx* ctx, float t_offset, int enable_rt, float color_temp) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return BHR_ERR_INVALID;
    if (!ctx->bg_ready) BHR_FAIL(ctx, BHR_ERR_STATE, "Must call init_background_layer() first");
    const size_t plane = (size_t)ctx->n_r * ctx->n_phi;
    if (int rc = bhr_join_entities(ctx)) return rc;
    compose_kernel<<<(unsigned)((plane + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->comp, ctx->omega_rows, ctx->edge, ctx->stats[0], ctx->stats[1], ctx->row_stats, ctx->n_r, ctx->n_phi,
        t_offset, enable_rt, color_temp, ctx->mips);
    ++ctx->launches;
    if (int rc = bhr_mark_comp_read(ctx)) return rc;
    BHR_CUDA(ctx, cudaGetLastError());
    return bhr_launch_build_mips(ctx, 0);
}

extern "C" int bhr_eval_noise(bhr_ctx* ctx, const float* coords, int n, int mode, int octaves, float persistence,
                              float lacunarity, float* out) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !coords || !out || n < 0) return BHR_ERR_INVALID;
    if (n == 0) return BHR_OK;
    float *dc = nullptr, *dout = nullptr;
    BHR_CUDA(ctx, cudaMalloc(&dc, (size_t)n * 3 * sizeof(float)));
    BHR_CUDA(ctx, cudaMalloc(&dout, (size_t)n * sizeof(float)));
    BHR_CUDA(ctx, cudaMemcpyAsync(dc, coords, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (mode == 2) {          // the background kernel's packed simplex noise (two points per thread)
        int rc = bhr_launch_noise_packed(ctx, dc, n, dout);
        if (rc) { cudaFree(dc); cudaFree(dout); return rc; }
    } else {
        noise_eval_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dc, n, mode, octaves, persistence, lacunarity, dout);
    }
    BHR_CUDA(ctx, cudaGetLastError());
    BHR_CUDA(ctx, cudaMemcpyAsync(out, dout, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(dc); cudaFree(dout);
    return BHR_OK;
}
