// noise.cuh -- 3-D simplex noise + FBM of the disk-texture pipeline (render.py:2642-2785), shared by texture.cu
// (noise test hook, one-texel-per-thread background kernel) and background.cu (packed background kernel).
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ float m_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float a_(float a, float b) { return __fadd_rn(a, b); }

// ---------------------------------------------------------------------------------------------
// 3-D simplex noise + FBM, render.py:2642-2785
// ---------------------------------------------------------------------------------------------
__constant__ unsigned char c_perm[256] = {
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

// the kernels read the permutation from shared memory (divergent byte lookups: constant memory
// would serialise them); t12[i] = t[i] % 12 serves the last lookup of each corner, whose result
// is only used modulo 12 (render.py:2643)
struct Perm {
    const unsigned char* t;      // 512 entries: the table twice, so ii + i1 + perm(...) <= 511 needs no mask (render.py:2269-2288)
    const float4* g;             // 512 entries: gradient direction of t[i] % 12 as (gx, gy, gz, 0), components in {-1, 0, 1}
    __device__ __forceinline__ int operator()(int i) const { return t[i]; }
};
constexpr int kPermSmem = 512 + 512 * 16;     // bytes
// gradient direction h in [0, 12): (+-u) + (+-v) with u = h < 8 ? x : y, v = h < 4 ? y : z and the
// signs from bits 0 / 1 of h (render.py:2642-2660; the h == 12 / 14 arm is unreachable)
__device__ __forceinline__ float4 grad_of(int h) {
    const float su = (h & 1) ? -1.0f : 1.0f, sv = (h & 2) ? -1.0f : 1.0f;
    if (h < 4) return make_float4(su, sv, 0.0f, 0.0f);
    if (h < 8) return make_float4(su, 0.0f, sv, 0.0f);
    return make_float4(0.0f, su, sv, 0.0f);
}
__device__ __forceinline__ void load_perm(unsigned char* smem /* kPermSmem bytes, 16-byte aligned */, Perm& perm) {
    float4* g = reinterpret_cast<float4*>(smem);
    unsigned char* t = smem + 512 * 16;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { t[i] = c_perm[i & 255]; g[i] = grad_of(c_perm[i & 255] % 12); }
    __syncthreads();
    perm.t = t; perm.g = g;
}

// The gradient dot product as gx x + gy y + gz z (left to right) with the tabulated direction:
// multiplying by +-1 is a sign flip and the one zero term adds +-0, so this is (+-u) + (+-v)
// exactly (up to the sign of an exact zero) -- on the FMA pipe instead of ~10 select / logic
// instructions on the half-rate ALU pipe, which is this kernel's bottleneck.
__device__ __forceinline__ float grad3_dot(const float4 g, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(g.x, x), __fmul_rn(g.y, y)), __fmul_rn(g.z, z));
}

// one corner: t = 0.6 - x^2 - y^2 - z^2; contributes t^4 * grad when t >= 0.  Branch-free: a corner
// outside the kernel adds +0.0f, which leaves the (never negative-zero) running sum unchanged.
__device__ __forceinline__ float corner(float n, const float4 h, float x, float y, float z) {
    float t = __fsub_rn(__fsub_rn(__fsub_rn(0.6f, __fmul_rn(x, x)), __fmul_rn(y, y)), __fmul_rn(z, z));
    const float t2 = __fmul_rn(t, t);
    const float c = __fmul_rn(__fmul_rn(t2, t2), grad3_dot(h, x, y, z));
    return __fadd_rn(n, t >= 0.0f ? c : 0.0f);
}

__device__ float simplex3(const Perm& perm, float x, float y, float z) {
    const float F3 = (float)(1.0 / 3.0), G3 = (float)(1.0 / 6.0);
    const float G3x2 = (float)(2.0 * (1.0 / 6.0)), G3x3 = (float)(3.0 * (1.0 / 6.0));
    // the skew / unskew arithmetic decides which lattice cell a point falls in; keep it exactly
    // rounded (no FMA contraction) so the cell choice matches the reference's f32 evaluation
    float s = __fmul_rn(__fadd_rn(__fadd_rn(x, y), z), F3);
    float fi = floorf(__fadd_rn(x, s)), fj = floorf(__fadd_rn(y, s)), fk = floorf(__fadd_rn(z, s));
    int i = (int)fi, j = (int)fj, k = (int)fk;
    float t = __fmul_rn((float)(i + j + k), G3);
    float x0 = __fsub_rn(x, __fsub_rn(fi, t)), y0 = __fsub_rn(y, __fsub_rn(fj, t)), z0 = __fsub_rn(z, __fsub_rn(fk, t));
    // simplex traversal order (render.py:2694-2712) as predicates of a = x0 >= y0, b = y0 >= z0, c = x0 >= z0
    const bool a = x0 >= y0, b = y0 >= z0, c = x0 >= z0;
    const int i1 = a && (b || c), j1 = !a && b, k1 = !b && !(a && c);
    const int i2 = a || (b && c), j2 = b || !a, k2 = !b || (!a && !c);
    float x1 = __fadd_rn(__fsub_rn(x0, (float)i1), G3), y1 = __fadd_rn(__fsub_rn(y0, (float)j1), G3), z1 = __fadd_rn(__fsub_rn(z0, (float)k1), G3);
    float x2 = __fadd_rn(__fsub_rn(x0, (float)i2), G3x2), y2 = __fadd_rn(__fsub_rn(y0, (float)j2), G3x2), z2 = __fadd_rn(__fsub_rn(z0, (float)k2), G3x2);
    float x3 = __fadd_rn(__fsub_rn(x0, 1.0f), G3x3), y3 = __fadd_rn(__fsub_rn(y0, 1.0f), G3x3), z3 = __fadd_rn(__fsub_rn(z0, 1.0f), G3x3);
    const int ii = i & 255, jj = j & 255, kk = k & 255;
    const float4 gi0 = perm.g[ii + perm(jj + perm(kk))];
    const float4 gi1 = perm.g[ii + i1 + perm(jj + j1 + perm(kk + k1))];
    const float4 gi2 = perm.g[ii + i2 + perm(jj + j2 + perm(kk + k2))];
    const float4 gi3 = perm.g[ii + 1 + perm(jj + 1 + perm(kk + 1))];
    float n = 0.0f;
    n = corner(n, gi0, x0, y0, z0);
    n = corner(n, gi1, x1, y1, z1);
    n = corner(n, gi2, x2, y2, z2);
    n = corner(n, gi3, x3, y3, z3);
    return __fmul_rn(32.0f, n);
}

__device__ float fbm3(const Perm& perm, float x, float y, float z, int octaves, float persistence, float lacunarity) {
    float value = 0.0f, amplitude = 1.0f, freq = 1.0f;
    for (int o = 0; o < octaves; ++o) {
        value = __fadd_rn(value, __fmul_rn(amplitude, simplex3(perm, __fmul_rn(x, freq), __fmul_rn(y, freq), __fmul_rn(z, freq))));
        amplitude = __fmul_rn(amplitude, persistence);
        freq = __fmul_rn(freq, lacunarity);
    }
    return value;
}

__device__ __forceinline__ float unit_fbm(const Perm& perm, float x, float y, float z, int o, float p) {
    return fminf(fmaxf(__fadd_rn(0.5f, __fmul_rn(0.5f, fbm3(perm, x, y, z, o, p, 2.0f))), 0.0f), 1.0f);
}


}  // namespace
