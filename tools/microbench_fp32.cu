// FP32 issue/throughput microbenchmarks for B200 (sm_100a): scalar FFMA vs packed FFMA2
// (PTX fma.rn.f32x2), and mixes with ALU-pipe (FMNMX) and XU-pipe (MUFU.RSQ) instructions.
// Used to fix the roofline denominator of the geodesic integrator and to decide
// the integrator's instruction mix.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

#define CHAINS 8
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters) {
    float x[CHAINS]; u64 v[CHAINS];
    for (int j = 0; j < CHAINS; ++j) { x[j] = threadIdx.x * 1e-3f + j; v[j] = pk(x[j], x[j] + 0.5f); }
    float m = 0.999f, c = 1e-3f; u64 m2 = pk(0.999f, 1.001f), c2 = pk(1e-3f, 2e-3f);
    float mm[4] = {1.5f, 2.5f, 3.5f, 4.5f}; float q[4] = {1.1f, 1.2f, 1.3f, 1.4f};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { // scalar FFMA
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) x[j] = fmaf(x[j], m, c);
            } else if (MODE == 1) { // FFMA2
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) v[j] = fma2(v[j], m2, c2);
            } else if (MODE == 2) { // 8 FFMA + 2 FMNMX
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) x[j] = fmaf(x[j], m, c);
                mm[u & 3] = fminf(mm[u & 3], x[u & 7]); mm[(u + 1) & 3] = fmaxf(mm[(u + 1) & 3], x[(u + 3) & 7]);
            } else if (MODE == 3) { // 8 FFMA2 + 2 FMNMX
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) v[j] = fma2(v[j], m2, c2);
                float a, b; up(v[u & 7], a, b);
                mm[u & 3] = fminf(mm[u & 3], a); mm[(u + 1) & 3] = fmaxf(mm[(u + 1) & 3], b);
            } else if (MODE == 4) { // 8 FFMA + 1 MUFU
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) x[j] = fmaf(x[j], m, c);
                q[u & 3] = rsq(q[u & 3]);
            } else if (MODE == 5) { // 8 FFMA2 + 1 MUFU
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) v[j] = fma2(v[j], m2, c2);
                q[u & 3] = rsq(q[u & 3]);
            } else if (MODE == 6) { // MUFU only
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = rsq(q[j]);
            } else if (MODE == 8) { // 6 FFMA + 1 FFMA2 interleaved
#pragma unroll
                for (int j = 0; j < 6; ++j) x[j] = fmaf(x[j], m, c);
                v[u & 7] = fma2(v[u & 7], m2, c2);
            } else if (MODE == 9) { // 4 FFMA + 2 FFMA2 interleaved (F F P F F P)
                x[0] = fmaf(x[0], m, c); x[1] = fmaf(x[1], m, c); v[u & 3] = fma2(v[u & 3], m2, c2);
                x[2] = fmaf(x[2], m, c); x[3] = fmaf(x[3], m, c); v[4 + (u & 3)] = fma2(v[4 + (u & 3)], m2, c2);
            } else if (MODE == 10) { // 4 FFMA + 4 FFMA2 alternating
#pragma unroll
                for (int j = 0; j < 4; ++j) { x[j] = fmaf(x[j], m, c); v[j] = fma2(v[j], m2, c2); }
            } else if (MODE == 11) { // 5 FFMA + 1 FFMA2 + 1 FMNMX (odd scalar count between packed ops)
#pragma unroll
                for (int j = 0; j < 5; ++j) x[j] = fmaf(x[j], m, c);
                v[u & 7] = fma2(v[u & 7], m2, c2);
                mm[u & 3] = fminf(mm[u & 3], x[u & 3]);
            } else if (MODE == 12) { // integrator-like: 10 FFMA + 2 FFMA2 + 1 MUFU + 2 FMNMX
#pragma unroll
                for (int j = 0; j < 5; ++j) x[j] = fmaf(x[j], m, c);
                v[u & 7] = fma2(v[u & 7], m2, c2);
#pragma unroll
                for (int j = 0; j < 5; ++j) x[j] = fmaf(x[j], m, c);
                v[(u + 1) & 7] = fma2(v[(u + 1) & 7], m2, c2);
                q[u & 3] = rsq(q[u & 3]);
                mm[u & 3] = fminf(mm[u & 3], x[u & 3]); mm[(u + 1) & 3] = fmaxf(mm[(u + 1) & 3], x[(u + 3) & 3]);
            } else if (MODE == 13) { // all packed with the same non-FMA load per 2 rays: 6 FFMA2 + 1 MUFU + 2 FMNMX... x2 rays
#pragma unroll
                for (int j = 0; j < 7; ++j) v[j] = fma2(v[j], m2, c2);
                q[u & 3] = rsq(q[u & 3]); q[(u + 1) & 3] = rsq(q[(u + 1) & 3]);
                float a, b; up(v[u & 7], a, b);
                mm[u & 3] = fminf(mm[u & 3], a); mm[(u + 1) & 3] = fmaxf(mm[(u + 1) & 3], b);
                mm[(u + 2) & 3] = fminf(mm[(u + 2) & 3], b); mm[(u + 3) & 3] = fmaxf(mm[(u + 3) & 3], a);
            } else if (MODE == 7) { // 8 FFMA + 4 MUFU
#pragma unroll
                for (int j = 0; j < CHAINS; ++j) x[j] = fmaf(x[j], m, c);
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = rsq(q[j]);
            }
        }
    }
    float s = 0;
    for (int j = 0; j < CHAINS; ++j) { float a, b; up(v[j], a, b); s += x[j] + a + b; }
    for (int j = 0; j < 4; ++j) s += mm[j] + q[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, double fma_per_iter, double other_per_iter, int blocks_per_sm) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * blocks_per_sm, iters = 20000;
    float* out; cudaMalloc(&out, sizeof(float) * grid * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) k<MODE><<<grid, 256>>>(out, iters);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double threads = (double)grid * 256, per_thread_iter = 8.0 * iters;
    double flops = threads * per_thread_iter * fma_per_iter * 2.0;
    double others = threads * per_thread_iter * other_per_iter;
    printf("%-28s occ=%d blk/SM  %.3f ms  %.2f TFLOP/s  other-instr %.2f Tinstr/s  err=%s\n", name, blocks_per_sm, best,
           flops / best / 1e9, others / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s  SMs %d  clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int occ : {2, 4, 8}) {
        run<0>("FFMA x8", 8, 0, occ);
        run<1>("FFMA2 x8 (16 fma)", 16, 0, occ);
        run<2>("FFMA x8 + FMNMX x2", 8, 2, occ);
        run<3>("FFMA2 x8 + FMNMX x2", 16, 2, occ);
        run<4>("FFMA x8 + MUFU x1", 8, 1, occ);
        run<5>("FFMA2 x8 + MUFU x1", 16, 1, occ);
        run<7>("FFMA x8 + MUFU x4", 8, 4, occ);
        run<8>("FFMA x6 + FFMA2 x1", 8, 0, occ);
        run<9>("FFMA x4 + FFMA2 x2", 8, 0, occ);
        run<10>("FFMA x4 + FFMA2 x4", 12, 0, occ);
        run<11>("FFMA x5 + FFMA2 x1 + FMNMX", 7, 1, occ);
        run<12>("FFMA x10+FFMA2 x2+MUFU+2FMNMX", 14, 3, occ);
        run<13>("FFMA2 x7 + 2 MUFU + 4 FMNMX", 14, 6, occ);
        run<6>("MUFU x4", 0, 4, occ);
    }
    return 0;
}
