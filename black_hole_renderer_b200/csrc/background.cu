// background.cu -- the simplex-FBM background layer of the disk texture, two texels per thread.
//
// Replaces _generate_background_kernel (render.py:3332-3451): 42 simplex evaluations per texel
// -> comp planes 0, 1, 2, 3, 4, 11, 12.  The kernel has to reproduce the reference's separately
// rounded float operations bit for bit; see fma2 below for how the packed arithmetic guarantees it.
//
// The one-texel-per-thread kernel (texture.cu: background_scalar_kernel, simplex3) runs at 82 % issue
// utilisation with the half-rate ALU pipe (compares, selects, integer hashing) as its busiest unit.
// This version
//   * evaluates the noise of two neighbouring texels in the two lanes of f32x2 registers: every
//     float operation of the pair is one FADD2 / FMUL2 issue slot, IEEE-rounded per lane in the
//     reference's order, so the planes are bit-identical to the scalar kernel's;
//   * takes work off the ALU and XU pipes: floor(v) and its integer come from one round-down magic
//     add (add.rm: MAGIC + floor(v) exactly for |v| < 2^22; the low byte of the sum's bit pattern
//     IS the hashed lattice index), float(i + j + k) is fi + fj + fk, and a corner outside its
//     kernel contributes max(t, 0)^4 * (g . p) = +-0 instead of a compare + select;
//   * runs ONE copy of the noise code in a table-driven loop over the 13 noise terms (the scalar
//     kernel inlines it 15 times, 74 KB of code: instruction-cache misses were 0.5 stall cycles per
//     issue in the packed kernel's first form);
//   * reads the row-only quantities (omega, two float64 pow) from a table filled when the layer is
//     initialised, and fills the hash / gradient tables once per persistent block.
#include <math.h>

#include "common.cuh"
#include "noise.cuh"

namespace {

// ---- packed, exactly rounded arithmetic (lanes = the thread's two texels) ----
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return ((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x); }
__device__ __forceinline__ float2 bits_f2(unsigned long long b) { return make_float2(__uint_as_float((unsigned)b), __uint_as_float((unsigned)(b >> 32))); }
// Every packed operation is an explicit fma.rn.f32x2: a * b + (-0) for a product, a * 1 + b for a sum, b * (-1) + a
// for a difference -- each rounds exactly like the plain operation (one rounding of the exact result; the -0 addend
// keeps the sign of a zero product).  The constants 1, -1 and -0 are RUN-TIME values (kernel arguments held in
// registers): ptxas contracts everything it can prove to be a plain packed add or multiply -- mul.rn.f32x2 feeding
// add.rn / sub.rn.f32x2 becomes one FFMA2, with --fmad=false too, and so does fma(a, 1.0, fma(b, c, -0.0)) -- one
// rounding instead of two, although scalar operations with an explicit rounding modifier are never fused.  With
// opaque constants an FMA is just an FMA.
struct Consts { float2 one, minus_one, neg_zero; };
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#define a2(a, b) fma2((a), K.one, (b))
#define s2(a, b) fma2((b), K.minus_one, (a))
#define m2(a, b) fma2((a), (b), K.neg_zero)
__device__ __forceinline__ float2 k2(float c) { return make_float2(c, c); }
// a >= b ? 1.0f : 0.0f.  (Not PTX set.ge.f32.f32: ptxas turned one lane's set into a predicate + SEL of the INTEGER
// 1 -- good enough for the `bits >> 29` consumer below, wrong for the FFMA2 that reads the same register as a float.)
__device__ __forceinline__ float ge01(float a, float b) { return a >= b ? 1.0f : 0.0f; }
__device__ __forceinline__ float2 clamp01_2(float2 v) { return make_float2(fminf(fmaxf(v.x, 0.0f), 1.0f), fminf(fmaxf(v.y, 0.0f), 1.0f)); }

// hash / gradient tables in shared memory: perm twice (512 bytes) and the gradient components of perm[i] % 12 as
// three float arrays, so that a corner's (g.x, g.y, g.z) of the two lanes load straight into f32x2 register pairs
struct Tables {
    const unsigned char* t;
    const float *gx, *gy, *gz;
};
constexpr int kTableBytes = 512 + 3 * 512 * 4;
__device__ __forceinline__ void load_tables(unsigned char* smem /* 16-byte aligned */, Tables& T) {
    float* g = reinterpret_cast<float*>(smem);
    unsigned char* t = smem + 3 * 512 * 4;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const int p = c_perm[i & 255];
        const float4 gr = grad_of(p % 12);
        t[i] = (unsigned char)p; g[i] = gr.x; g[512 + i] = gr.y; g[1024 + i] = gr.z;
    }
    __syncthreads();
    T.t = t; T.gx = g; T.gy = g + 512; T.gz = g + 1024;
}

// lattice cell of v (both lanes): f = floor(v) as floats, and the sums' bit patterns, whose low bytes are the
// lattice indices modulo 256 (MAGIC + n has the integer pattern 0x4B400000 + n for |n| < 2^22)
__device__ __forceinline__ void floor2(const Consts& K, const float2 v, float2& f, unsigned& bits_a, unsigned& bits_b) {
    const float MAGIC = 12582912.0f;                       // 1.5 * 2^23
    // (FRND + F2I on the XU pipe instead of these two FFMA2 was measured: no difference)
    const float2 r = __ffma2_rd(v, K.one, k2(MAGIC));      // rounded DOWN: MAGIC + floor(v)
    bits_a = __float_as_uint(r.x); bits_b = __float_as_uint(r.y);
    f = s2(r, k2(MAGIC));
}

// gradient index of one corner of one lane: perm[ii + o + perm[jj + p + pk]] (all offsets already added in)
__device__ __forceinline__ int corner_index(const Tables& T, int ii_o, int jj_p, int pk) { return ii_o + T.t[jj_p + pk]; }

// one simplex corner of both lanes: n += max(t, 0)^4 * (g . p),  t = 0.6 - x^2 - y^2 - z^2.  (For t < 0 the reference
// adds nothing; here the term is (+-0) and n + (+-0) == n bit for bit -- n is never -0.)
__device__ __forceinline__ float2 corner2(const Consts& K, const Tables& T, const float2 n, const int ga, const int gb, const float2 x, const float2 y, const float2 z) {
    const float2 t = s2(s2(s2(k2(0.6f), m2(x, x)), m2(y, y)), m2(z, z));
    const float2 tp = make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f));
    const float2 t2 = m2(tp, tp);
    // g . p = (gx x + gy y) + gz z, components in {-1, 0, 1} (render.py:2642-2660 as grad3_dot of noise.cuh)
    const float2 gx = make_float2(T.gx[ga], T.gx[gb]), gy = make_float2(T.gy[ga], T.gy[gb]), gz = make_float2(T.gz[ga], T.gz[gb]);
    const float2 dot = a2(a2(m2(gx, x), m2(gy, y)), m2(gz, z));
    return a2(n, m2(m2(t2, t2), dot));
}

__device__ __forceinline__ float2 simplex3x2(const Consts& K, const Tables& T, const float2 x, const float2 y, const float2 z) {
    const float F3 = (float)(1.0 / 3.0), G3 = (float)(1.0 / 6.0);
    const float G3x2 = (float)(2.0 * (1.0 / 6.0)), G3x3 = (float)(3.0 * (1.0 / 6.0));
    const float2 s = m2(a2(a2(x, y), z), k2(F3));
    float2 fi, fj, fk;
    unsigned ia, ib, ja, jb, ka, kb;
    floor2(K, a2(x, s), fi, ia, ib);
    floor2(K, a2(y, s), fj, ja, jb);
    floor2(K, a2(z, s), fk, ka, kb);
    const float2 t = m2(a2(a2(fi, fj), fk), k2(G3));      // float(i + j + k) * G3: the sum of three small integers is exact
    const float2 x0 = s2(x, s2(fi, t)), y0 = s2(y, s2(fj, t)), z0 = s2(z, s2(fk, t));
    // simplex traversal order (render.py:2694-2712) as predicates of a = x0 >= y0, b = y0 >= z0, c = x0 >= z0, per
    // lane, on the ALU pipe (compares, predicate logic, selects): the FMA pipe is this kernel's busiest unit
    float2 i1, j1, k1, i2, j2, kk2;
    int o1[2], p1[2], q1[2], o2[2], p2[2], q2[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const float xx = l ? x0.y : x0.x, yy = l ? y0.y : y0.x, zz = l ? z0.y : z0.x;
        const bool a = xx >= yy, b = yy >= zz, c = xx >= zz;
        const bool bi1 = a && (b || c), bj1 = !a && b, bk1 = !b && !(a && c);
        const bool bi2 = a || (b && c), bj2 = b || !a, bk2 = !b || (!a && !c);
        (l ? i1.y : i1.x) = bi1 ? 1.0f : 0.0f; (l ? j1.y : j1.x) = bj1 ? 1.0f : 0.0f; (l ? k1.y : k1.x) = bk1 ? 1.0f : 0.0f;
        (l ? i2.y : i2.x) = bi2 ? 1.0f : 0.0f; (l ? j2.y : j2.x) = bj2 ? 1.0f : 0.0f; (l ? kk2.y : kk2.x) = bk2 ? 1.0f : 0.0f;
        o1[l] = bi1; p1[l] = bj1; q1[l] = bk1; o2[l] = bi2; p2[l] = bj2; q2[l] = bk2;
    }
    const float2 one = k2(1.0f);
    const float2 x1 = a2(s2(x0, i1), k2(G3)), y1 = a2(s2(y0, j1), k2(G3)), z1 = a2(s2(z0, k1), k2(G3));
    const float2 x2 = a2(s2(x0, i2), k2(G3x2)), y2 = a2(s2(y0, j2), k2(G3x2)), z2 = a2(s2(z0, kk2), k2(G3x2));
    const float2 x3 = a2(s2(x0, one), k2(G3x3)), y3 = a2(s2(y0, one), k2(G3x3)), z3 = a2(s2(z0, one), k2(G3x3));
    // hashing, per lane
    int g0a, g1a, g2a, g3a, g0b, g1b, g2b, g3b;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const unsigned bi = l ? ib : ia, bj = l ? jb : ja, bk = l ? kb : ka;
        const int ii = bi & 255, jj = bj & 255, kk = bk & 255;
        const int pk0 = T.t[kk], pk1 = T.t[kk + 1];
        const int c0 = corner_index(T, ii, jj, pk0);
        const int c1 = corner_index(T, ii + o1[l], jj + p1[l], q1[l] ? pk1 : pk0);
        const int c2 = corner_index(T, ii + o2[l], jj + p2[l], q2[l] ? pk1 : pk0);
        const int c3 = corner_index(T, ii + 1, jj + 1, pk1);
        if (l) { g0b = c0; g1b = c1; g2b = c2; g3b = c3; } else { g0a = c0; g1a = c1; g2a = c2; g3a = c3; }
    }
    float2 n = make_float2(0.0f, 0.0f);
    n = corner2(K, T, n, g0a, g0b, x0, y0, z0);
    n = corner2(K, T, n, g1a, g1b, x1, y1, z1);
    n = corner2(K, T, n, g2a, g2b, x2, y2, z2);
    n = corner2(K, T, n, g3a, g3b, x3, y3, z3);
    return m2(k2(32.0f), n);
}

// The 13 noise terms of a texel (render.py:3383-3447): coordinates (f cx, f cy, zr r + zt t), `oct` octaves of
// persistence `pers` (lacunarity 2), weight `w` in its sum.  kind 0: clamp(0.5 + 0.5 fbm, 0, 1); kind 1: clamp(simplex, 0, 1).
struct NoiseTerm { float f, zr, zt, pers, w; int oct, kind; };
__constant__ NoiseTerm c_terms[13] = {
    {8.0f, 8.0f, 0.05f, 0.6f, 1.0f, 4, 0},          //  0  temp_base
    {8.0f, 4.0f, 0.06f, 0.45f, 0.08f, 3, 0},        //  1  turbulence: coarse
    {24.0f, 12.0f, 0.08f, 0.45f, 0.15f, 4, 0},      //  2              mid
    {80.0f, 40.0f, 0.1f, 0.45f, 0.25f, 5, 0},       //  3              fine
    {200.0f, 100.0f, 0.12f, 0.4f, 0.22f, 4, 0},     //  4              extra
    {400.0f, 200.0f, 0.15f, 0.35f, 0.18f, 3, 0},    //  5              ultra
    {800.0f, 400.0f, 0.2f, 1.0f, 0.12f, 1, 1},      //  6              pixel (plain simplex)
    {3.0f, 3.0f, 0.04f, 0.5f, 1.0f, 3, 0},          //  7  azimuthal hotspot noise
    {8.0f, 4.0f, 0.003f, 0.5f, 0.05f, 3, 0},        //  8  disturbance: coarse
    {32.0f, 16.0f, 0.005f, 0.5f, 0.15f, 3, 0},      //  9               mid
    {100.0f, 50.0f, 0.006f, 0.45f, 0.30f, 4, 0},    // 10               fine
    {250.0f, 125.0f, 0.008f, 0.4f, 0.30f, 4, 0},    // 11               extra
    {500.0f, 250.0f, 0.01f, 1.0f, 0.20f, 1, 1},     // 12               pixel (plain simplex)
};

// Quantities of _generate_background_kernel that depend on the row only (render.py:3362-3380): evaluated once when
// the layer is initialised (two float64 pow per row) instead of by one thread of every block with the others waiting.
//   rows[0][ri] = omega(r_phys), rows[1][ri] = max(1 - r, 0)^1.3, rows[2][ri] = r^1.2 * az_shear
__global__ void background_rows_kernel(float* __restrict__ rows, int n_r, float az_shear, float r_inner, float r_outer) {
    const int ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= n_r) return;
    const float r = __fdiv_rn((float)ri, (float)n_r);
    const float r_phys = a_(r_inner, m_(__fsub_rn(r_outer, r_inner), r));
    rows[ri] = __fsqrt_rn(__fdiv_rn(0.5f, a_(m_(m_(r_phys, r_phys), r_phys), 1e-6f)));
    rows[n_r + ri] = (float)pow((double)fmaxf(__fsub_rn(1.0f, r), 0.0f), (double)1.3f);
    rows[2 * n_r + ri] = m_((float)pow((double)r, (double)1.2f), az_shear);
}

// A thread owns two neighbouring texels of a row; blocks are persistent and walk the texel pairs grid-stride.
// The cos / sin of the Keplerian-rotated angle feed noise coordinates scaled by up to 800, so they are evaluated
// in double and rounded once (the oracle's ideal-libm convention).
__global__ void __launch_bounds__(256, 2) background_kernel(float* __restrict__ comp, const float* __restrict__ rows, int n_r, int n_phi,
                                                         int az_freq, float t, const Consts K) {
    __shared__ __align__(16) unsigned char stab[kTableBytes];
    Tables T;
    load_tables(stab, T);
    const size_t plane = (size_t)n_r * n_phi;
    const int half = n_phi >> 1;                                  // (n_phi is a multiple of 16)
    const int n_pairs = n_r * half;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += gridDim.x * blockDim.x) {
        const int ri = q / half, pi = (q - ri * half) * 2;
        const size_t o = (size_t)ri * n_phi + pi;
        const float r = __fdiv_rn((float)ri, (float)n_r);
        const float omega = __ldg(rows + ri), decay = __ldg(rows + n_r + ri), shear = __ldg(rows + 2 * n_r + ri);
        const float wt = m_(omega, t);
        const float rot_a = a_(m_(__fdiv_rn((float)pi, (float)n_phi), 6.2831855f), wt);
        const float rot_b = a_(m_(__fdiv_rn((float)(pi + 1), (float)n_phi), 6.2831855f), wt);
        const float2 cx = make_float2((float)cos((double)rot_a), (float)cos((double)rot_b));
        const float2 cy = make_float2((float)sin((double)rot_a), (float)sin((double)rot_b));
        auto st2 = [&](int pl, float2 v) { *reinterpret_cast<float2*>(comp + pl * plane + o) = v; };
        st2(1, k2(0.0f));
        st2(2, k2(0.0f));
        float2 sum = k2(0.0f);
#pragma unroll 1
        for (int term = 0; term < 13; ++term) {
            const NoiseTerm P = c_terms[term];
            const float z = a_(m_(r, P.zr), m_(t, P.zt));
            const float2 x = m2(cx, k2(P.f)), y = m2(cy, k2(P.f));
            float2 value = k2(0.0f);
            float amplitude = 1.0f, freq = 1.0f;
#pragma unroll 1
            for (int oc = 0; oc < P.oct; ++oc) {
                const float2 f2 = k2(freq);
                value = a2(value, m2(k2(amplitude), simplex3x2(K, T, m2(x, f2), m2(y, f2), k2(m_(z, freq)))));
                amplitude = m_(amplitude, P.pers);
                freq = m_(freq, 2.0f);
            }
            // (kind 1: value = 0 + 1 * simplex = simplex exactly)
            const float2 u = clamp01_2(P.kind ? value : a2(k2(0.5f), m2(k2(0.5f), value)));
            const float2 wu = m2(u, k2(P.w));
            if (term == 0) {
                st2(0, m2(m2(k2(decay), a2(k2(0.85f), m2(k2(0.15f), u))), k2(0.25f)));
            } else if (term == 7) {
                const float faz = (float)az_freq;
                const float2 wave = a2(k2(0.5f), m2(k2(0.5f), make_float2((float)sin((double)m_(a_(rot_a, shear), faz)),
                                                                       (float)sin((double)m_(a_(rot_b, shear), faz)))));
                st2(11, m2(wave, u));
            } else {
                sum = (term == 1 || term == 8) ? wu : a2(sum, wu);
                if (term == 6) {
                    const float2 turb = clamp01_2(sum);
                    st2(3, turb);
                    st2(4, m2(k2(0.05f), turb));
                } else if (term == 12) {
                    float2 raw = m2(sum, k2(1.4f));
                    raw = make_float2(fminf(fmaxf(raw.x, 0.05f), 1.0f), fminf(fmaxf(raw.y, 0.05f), 1.0f));
                    const float2 pres = m2(raw, k2(a_(0.6f, m_(0.4f, r))));
                    st2(12, make_float2(fminf(fmaxf(pres.x, 0.1f), 1.0f), fminf(fmaxf(pres.y, 0.1f), 1.0f)));
                }
            }
        }
    }
}

// test hook (bhr_eval_noise mode 2): the packed simplex noise on consecutive point pairs, lane .x = point 2i, lane .y =
// point 2i + 1 (the z coordinate of a pair is common in the background kernel; here each lane keeps its own)
__global__ void __launch_bounds__(256) noise_eval_packed_kernel(const float* __restrict__ coords, int n, float* __restrict__ out, const Consts K) {
    __shared__ __align__(16) unsigned char stab[kTableBytes];
    Tables T;
    load_tables(stab, T);
    const int i = 2 * (blockIdx.x * 256 + threadIdx.x);
    if (i >= n) return;
    const int j = i + 1 < n ? i + 1 : i;
    const float2 v = simplex3x2(K, T, make_float2(coords[3 * i], coords[3 * j]), make_float2(coords[3 * i + 1], coords[3 * j + 1]),
                                make_float2(coords[3 * i + 2], coords[3 * j + 2]));
    out[i] = v.x;
    if (i + 1 < n) out[i + 1] = v.y;
}

}  // namespace

int bhr_launch_noise_packed(bhr_ctx* ctx, const float* d_coords, int n, float* d_out) {
    Consts K;
    K.one = make_float2(1.0f, 1.0f); K.minus_one = make_float2(-1.0f, -1.0f); K.neg_zero = make_float2(-0.0f, -0.0f);
    noise_eval_packed_kernel<<<bhr_div_up((n + 1) / 2, 256), 256, 0, ctx->stream>>>(d_coords, n, d_out, K);
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

int bhr_setup_background(bhr_ctx* ctx) {
    if (ctx->bg_rows) { cudaFree(ctx->bg_rows); ctx->bg_rows = nullptr; }
    BHR_CUDA(ctx, cudaMalloc(&ctx->bg_rows, 3 * (size_t)ctx->n_r * sizeof(float)));
    background_rows_kernel<<<bhr_div_up(ctx->n_r, 128), 128, 0, ctx->stream>>>(ctx->bg_rows, ctx->n_r, ctx->az_shear, ctx->cfg.r_disk_inner,
                                                                             ctx->cfg.r_disk_outer);
    BHR_CUDA(ctx, cudaGetLastError());
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // the entity kernels may share the SMs with this kernel (option "entity_stream"): kernels whose shared-memory carve-outs
    // differ cannot be resident on one SM together, so ask for a carve-out that has room for their blocks as well
    BHR_CUDA(ctx, cudaFuncSetAttribute(background_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    int per_sm = 0;
    BHR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, background_kernel, 256, 0));
    ctx->bg_blocks_per_sm = per_sm < 1 ? 1 : per_sm;
    return BHR_OK;
}

int bhr_launch_background(bhr_ctx* ctx, float t) {
    // persistent blocks (hash / gradient tables filled once each): as many as stay resident
    const int want = bhr_div_up(ctx->n_r * (ctx->n_phi / 2), 256);
    // With the entity stream released by this kernel's start (option entity_early = 0), ONE block per SM: the two blocks
    // that fit hold the whole register file, and the entity kernels (FP64 / XU heavy, this one FP32 heavy) are to run beside
    // it -- alone that costs 0.330 instead of 0.307 ms, together with the entity layer 0.391 instead of 0.435
    // (tools/entity_overlap.py).  With the early release (default) the entity layer has mostly run beside the previous
    // frame's bloom passes by now, and two blocks are better again (video frame 1.030 vs 1.036 ms, tools/entity_stream_ab.py).
    const int per_sm = ctx->bg_blocks_override > 0 ? ctx->bg_blocks_override : (ctx->entity_stream_on && !ctx->entity_early) ? 1 : ctx->bg_blocks_per_sm;
    const int cap = per_sm * ctx->num_sms;
    Consts K;
    K.one = make_float2(1.0f, 1.0f); K.minus_one = make_float2(-1.0f, -1.0f); K.neg_zero = make_float2(-0.0f, -0.0f);
    background_kernel<<<want < cap ? want : cap, 256, 0, ctx->stream>>>(ctx->comp, ctx->bg_rows, ctx->n_r, ctx->n_phi, ctx->az_freq, t, K);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}
