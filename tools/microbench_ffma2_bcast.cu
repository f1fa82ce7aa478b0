// Does FFMA2 with a broadcast scalar operand (SASS "Rb.F32") run at the rate of FFMA2 with a register-pair operand?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_ffma2_bcast.cu -o tools/microbench_ffma2_bcast
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
    float2 a[12];
    for (int i = 0; i < 12; ++i) a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 32 + i]);
    const float w0 = in[threadIdx.x + 64], w1 = in[threadIdx.x + 65];
    float2 wp = make_float2(w0, w1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                if (MODE == 0) a[i] = __ffma2_rn(a[i], wp, a[(i + 1) % 12]);                       // pair operand
                else if (MODE == 1) a[i] = __ffma2_rn(a[i], make_float2(w0, w0), a[(i + 1) % 12]);   // broadcast scalar
                else { a[i].x = fmaf(a[i].x, w0, a[(i + 1) % 12].x); a[i].y = fmaf(a[i].y, w0, a[(i + 1) % 12].y); }   // scalar FFMA
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 12; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out, float* in) {
    const int iters = 20000, blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, in, 100);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)blocks * 256 * iters * 4 * 12 * 4;      // 2 lanes x 2 flop per packed op
    printf("%-34s %.3f ms  %.1f TFLOP/s\n", name, ms, flop / ms / 1e9);
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run<0>("FFMA2 pair operand", out, in);
    run<1>("FFMA2 broadcast scalar operand", out, in);
    run<2>("scalar FFMA x2", out, in);
    run<0>("FFMA2 pair operand", out, in);
    run<1>("FFMA2 broadcast scalar operand", out, in);
    return 0;
}
