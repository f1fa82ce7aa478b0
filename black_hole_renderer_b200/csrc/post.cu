// post.cu -- separable chromatic bloom, composite (+ lens flare) and the u8 conversion.
//
// Replaces _bloom_kernel (render.py:3022-3114) as render() drives it (render.py:3914-3918:
// threshold 0, so the live result is the H pass followed by the V pass of the disk layer), the
// host-side numpy composite `clip(img + disk + blur, 0, 1)` + transpose (render.py:3918-3923),
// the numpy lens flare (render.py:3925-4028) and the truncating u8 conversion (render.py:4463).
//
// Layers are planar (3 x H x W).  Both passes keep a sliding window of weights and P = 10 running
// outputs in registers so that every loaded sample feeds 10 FMAs:
//   H pass: one warp per row segment of 320 outputs, samples staged transposed in shared memory;
//   V pass: lanes along x, 10 consecutive rows per thread, a 32 x (80 + 2R) tile in shared memory;
//   composite: streaming kernel, four pixels per thread, flare (float64) fused in.
// Taps outside the image are skipped and the sum is renormalised by the in-bounds weight sum,
// which is tabulated per x / per y in the reference's sequential f32 summation order.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "post.cuh"

namespace {

constexpr int P_OUT = 10;   // outputs per thread: 32 * 10 = 320 divides every standard frame width

__host__ __device__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------
// H pass (one channel plane per blockIdx.z).  block = 8 warps; warp w blurs outputs
// x0 .. x0 + 319 of row blockIdx.y * 8 + w.  Thread `lane` owns the 10 consecutive outputs
// x0 + 10 lane + p.  The row segment (sample s <-> x = x0 - R + s) is staged in shared memory
// TRANSPOSED: s lives at (s % 10) * pitch + s / 10, so that window position k = 10 j + kk of all
// lanes reads the consecutive words kk * pitch + j + lane (no bank conflicts, no index math).
// Each loaded sample feeds 10 FMAs through a sliding register window of weights.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bloom_h_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                      int W, int row0, int row1, int R,
                                                      const float* __restrict__ wtab, int wtab_stride,
                                                      const float* __restrict__ wsum_x, size_t plane) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.z;
    const int nk = round_up(2 * R + P_OUT, P_OUT);          // window positions, padded to P_OUT
    const int seg = 31 * P_OUT + nk;                        // samples any lane may touch
    const int pitch = (seg / P_OUT + 1) | 1;                // odd: spreads the staging stores over banks
    float* wsh = smem;                                      // zero-padded weights of this channel
    float* row = smem + wtab_stride + warp * (P_OUT * pitch);
    for (int k = threadIdx.x; k < wtab_stride; k += 256) wsh[k] = wtab[ch * wtab_stride + k];
    const int y = row0 + blockIdx.y * 8 + warp;
    const int x0 = blockIdx.x * (32 * P_OUT);
    const bool row_ok = y < row1;
    if (row_ok) {
        const float* srow = src + ch * plane + (size_t)y * W;
        // s = lane + 32 i  ->  (s % 10, s / 10) kept incrementally: 32 = 3 * 10 + 2
        int sr = lane % P_OUT, sq = lane / P_OUT;
        for (int s = lane; s < seg; s += 32) {
            const int x = x0 - R + s;
            row[sr * pitch + sq] = (x >= 0 && x < W) ? __ldg(srow + x) : 0.0f;
            sr += 32 % P_OUT; sq += 32 / P_OUT;
            if (sr >= P_OUT) { sr -= P_OUT; ++sq; }
        }
    }
    __syncthreads();
    if (!row_ok) return;

    float acc[P_OUT], wr[P_OUT];
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) { acc[p] = 0.0f; wr[p] = 0.0f; }
    const float* rp = row + lane;
    for (int k0 = 0; k0 < nk; k0 += P_OUT, ++rp) {
#pragma unroll
        for (int kk = 0; kk < P_OUT; ++kk) {
            // rotate the weight window: wr[p] = w[k - p]  (zero outside [0, 2R])
#pragma unroll
            for (int p = P_OUT - 1; p > 0; --p) wr[p] = wr[p - 1];
            wr[0] = wsh[k0 + kk];
            const float v = rp[kk * pitch];
#pragma unroll
            for (int p = 0; p < P_OUT; ++p) acc[p] = fmaf(v, wr[p], acc[p]);
        }
    }
    // transpose the results through the (now free) row buffer for coalesced stores
    __syncwarp();
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) row[p * pitch + lane] = acc[p];
    __syncwarp();
    float* drow = dst + ch * plane + (size_t)y * W;
    const float* ws = wsum_x + ch * W;
    int ir = lane % P_OUT, iq = lane / P_OUT;
#pragma unroll
    for (int j = 0; j < P_OUT; ++j) {
        const int x = x0 + lane + 32 * j;
        if (x < W) drow[x] = row[ir * pitch + iq] * ws[x];       // ws = 1 / in-bounds weight sum (0 if none)
        ir += 32 % P_OUT; iq += 32 / P_OUT;
        if (ir >= P_OUT) { ir -= P_OUT; ++iq; }
    }
}

// ---------------------------------------------------------------------------------------------
// V pass (one channel plane per blockIdx.z).  block = 32 (x) x 8 (row groups) threads, tile =
// 32 columns x 80 rows.  The (80 + 2R) x 32 samples are staged in shared memory (coalesced rows,
// zero outside the image); each thread slides its 10-output window down its column.
// ---------------------------------------------------------------------------------------------
constexpr int V_TILE_ROWS = 8 * P_OUT;

template <bool PEER>
__global__ void __launch_bounds__(256) bloom_v_kernel(const float* __restrict__ hblur, float* __restrict__ blur,
                                                      int W, int H, int row0, int row1, int R,
                                                      const float* __restrict__ wtab, int wtab_stride,
                                                      const float* __restrict__ wsum_y, size_t plane,
                                                      const float* const* __restrict__ row_src) {
    extern __shared__ float vsm[];
    const int ch = blockIdx.z;
    const int nk = round_up(2 * R + P_OUT, P_OUT);
    const int tile_rows = V_TILE_ROWS - P_OUT + nk;
    float* wsh = vsm;
    float* tile = vsm + wtab_stride;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k < wtab_stride; k += 256) wsh[k] = wtab[ch * wtab_stride + k];
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int ty0 = row0 + blockIdx.y * V_TILE_ROWS;
    const bool col_ok = x < W;
    const float* src = hblur + ch * plane;
    for (int r = threadIdx.y; r < tile_rows; r += 8) {
        const int y = ty0 - R + r;
        float v = 0.0f;
        if (col_ok && y >= 0 && y < H) {
            // PEER (row-tiled frame over several GPUs): row y of the H-blurred layer is read from the
            // HBM of the rank that produced it (same offset in every rank's buffer; peer loads over
            // NVLink).  Plain loads there: the read-only path must not cache another GPU's live data.
            if (PEER) v = row_src[y][ch * plane + (size_t)y * W + x];
            else v = __ldg(src + (size_t)y * W + x);
        }
        tile[r * 32 + threadIdx.x] = v;
    }
    __syncthreads();
    float acc[P_OUT], wr[P_OUT];
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) { acc[p] = 0.0f; wr[p] = 0.0f; }
    const float* colp = tile + threadIdx.y * P_OUT * 32 + threadIdx.x;
    for (int k0 = 0; k0 < nk; k0 += P_OUT, colp += P_OUT * 32) {
#pragma unroll
        for (int kk = 0; kk < P_OUT; ++kk) {
#pragma unroll
            for (int p = P_OUT - 1; p > 0; --p) wr[p] = wr[p - 1];
            wr[0] = wsh[k0 + kk];
            const float v = colp[kk * 32];
#pragma unroll
            for (int p = 0; p < P_OUT; ++p) acc[p] = fmaf(v, wr[p], acc[p]);
        }
    }
    if (!col_ok) return;
    const int y0 = ty0 + threadIdx.y * P_OUT;
    float* dst = blur + ch * plane;
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) {
        const int y = y0 + p;
        if (y < row1) {
            dst[(size_t)y * W + x] = acc[p] * wsum_y[ch * H + y];   // 1 / in-bounds weight sum (0 if none)
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Composite: final = clip(bg + disk [+ blur], 0, 1) [+ flare], written as (H, W, 3) f32 and u8
// (render.py:3912-3923, 4463).  Four pixels per thread: float4 plane loads, 3 x float4 + 3 x u32
// stores when the width allows it.
// ---------------------------------------------------------------------------------------------
template <bool BLOOM, int VEC, bool FLARE>
__global__ void __launch_bounds__(256) composite_kernel(const float* __restrict__ bg, const float* __restrict__ disk,
                                                        const float* __restrict__ blur, float* __restrict__ final_f32,
                                                        uint8_t* __restrict__ final_u8, int W, int row0, int row1,
                                                        size_t plane, FlareParams F_arg, const FlareParams* __restrict__ F_dev,
                                                        float field_gain) {
    // (peer path: the flare parameters are reduced on the device from every rank's partial sums)
    const FlareParams F = F_dev ? *F_dev : F_arg;
    const int groups_per_row = (W + VEC - 1) / VEC;
    const size_t n = (size_t)(row1 - row0) * groups_per_row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = row0 + (int)(i / groups_per_row), x = (int)(i % groups_per_row) * VEC;
        const size_t o = (size_t)y * W + x;
        float v[3][VEC];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a[VEC], d[VEC], b[VEC];
            if (VEC == 4) {
                const float4 a4 = *reinterpret_cast<const float4*>(bg + c * plane + o);
                const float4 d4 = *reinterpret_cast<const float4*>(disk + c * plane + o);
                a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
                d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
                if (BLOOM) {
                    const float4 b4 = *reinterpret_cast<const float4*>(blur + c * plane + o);
                    b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
                }
            } else {
                a[0] = bg[c * plane + o]; d[0] = disk[c * plane + o];
                if (BLOOM) b[0] = blur[c * plane + o];
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float dd = d[j];
                // render_to_field: the disk layer after _bloom_kernel's in-place add (render.py:3112-3114)
                if (BLOOM && field_gain != 0.0f) dd = fminf(fmaxf(__fadd_rn(dd, __fmul_rn(b[j], field_gain)), 0.0f), 1.0f);
                float t = a[j] + dd;
                if (BLOOM) t = t + b[j];
                v[c][j] = fminf(fmaxf(t, 0.0f), 1.0f);
            }
        }
        if (FLARE && F.enabled) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float fl[3];
                flare_pixel(F, x + j, y, fl);
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c][j] = fminf(fmaxf(v[c][j] + fl[c], 0.0f), 1.0f);
            }
        }
        if (VEC == 4) {
            float4* of = reinterpret_cast<float4*>(final_f32 + o * 3);
            of[0] = make_float4(v[0][0], v[1][0], v[2][0], v[0][1]);
            of[1] = make_float4(v[1][1], v[2][1], v[0][2], v[1][2]);
            of[2] = make_float4(v[2][2], v[0][3], v[1][3], v[2][3]);
            unsigned char q[12];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) q[j * 3 + c] = (unsigned char)(v[c][j] * 255.0f);
            uint32_t* ou = reinterpret_cast<uint32_t*>(final_u8 + o * 3);
#pragma unroll
            for (int w = 0; w < 3; ++w)
                ou[w] = (uint32_t)q[4 * w] | ((uint32_t)q[4 * w + 1] << 8) | ((uint32_t)q[4 * w + 2] << 16) | ((uint32_t)q[4 * w + 3] << 24);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                final_f32[o * 3 + c] = v[c][0];
                final_u8[o * 3 + c] = (unsigned char)(v[c][0] * 255.0f);
            }
        }
    }
}

// brightness sums for the flare centroid: {sum B, sum x*B, sum y*B}, B = max(r, g, b) of the disk layer
__global__ void __launch_bounds__(256) flare_sums_kernel(const float* __restrict__ disk, int W, int row0, int row1,
                                                         size_t plane, double* __restrict__ sums) {
    double sb = 0.0, sx = 0.0, sy = 0.0;
    const size_t n = (size_t)(row1 - row0) * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int y = row0 + (int)(i / W), x = (int)(i % W);
        size_t o = (size_t)y * W + x;
        float b = fmaxf(fmaxf(disk[o], disk[o + plane]), disk[o + 2 * plane]);
        sb += (double)b; sx += (double)x * (double)b; sy += (double)y * (double)b;
    }
    __shared__ double sh[3][8];
    for (int off = 16; off > 0; off >>= 1) {
        sb += __shfl_down_sync(0xffffffffu, sb, off);
        sx += __shfl_down_sync(0xffffffffu, sx, off);
        sy += __shfl_down_sync(0xffffffffu, sy, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = sb; sh[1][warp] = sx; sh[2][warp] = sy; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        sums[blockIdx.x * 3 + threadIdx.x] = t;          // per-block partial: summed in block order below (deterministic)
    }
}

__global__ void flare_sums_final_kernel(const double* __restrict__ parts, int n_blocks, double* __restrict__ sums) {
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int b = 0; b < n_blocks; ++b) t += parts[b * 3 + threadIdx.x];
        sums[threadIdx.x] = t;
    }
}

// disk_layer_field as _bloom_kernel leaves it: clamp(disk + 0.4 blur, 0, 1), render.py:3112-3114
__global__ void __launch_bounds__(256) disk_post_kernel(const float* __restrict__ disk, const float* __restrict__ blur,
                                                        float* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fminf(fmaxf(__fadd_rn(disk[i], __fmul_rn(blur[i], 0.4f)), 0.0f), 1.0f);
}

}  // namespace

int bhr_launch_disk_post(bhr_ctx* ctx, float* out) {
    disk_post_kernel<<<148 * 8, 256, 0, ctx->stream>>>(ctx->disk, ctx->blur, out, (size_t)ctx->W * ctx->H * 3);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// weights exp(-d^2 / (sigma2 * sigma_scale)) (render.py:3058-3060) evaluated like the oracle
// (double exp rounded once), plus the reciprocals of the in-bounds weight sums (summed in the
// reference's tap order; the kernels multiply by the reciprocal instead of dividing: <= 1 ulp)
int bhr_setup_bloom_tables(bhr_ctx* ctx) {
    const int W = ctx->W, H = ctx->H;
    const int R = (int)(W * 0.02);                          // render.py:3914
    const float sigma_scale = (float)((W / 640.0) * (W / 640.0));   // render.py:3915 (f64 -> f32 kernel arg)
    ctx->bloom_R = R; ctx->sigma_scale = sigma_scale;
    const int stride = round_up(2 * R + P_OUT, P_OUT) + P_OUT;   // zero padded past the last tap
    ctx->wtab_stride = stride;
    float* wt = (float*)calloc((size_t)3 * stride, sizeof(float));
    float* wx = (float*)malloc((size_t)3 * W * sizeof(float));
    float* wy = (float*)malloc((size_t)3 * H * sizeof(float));
    if (!wt || !wx || !wy) BHR_FAIL(ctx, BHR_ERR_NOMEM, "host allocation failed");
    const float s2[3] = {25.0f, 80.0f, 1600.0f};
    for (int c = 0; c < 3; ++c)
        for (int d = -R; d <= R; ++d) {
            float dist_sq = (float)(d * d);
            wt[c * stride + d + R] = (float)exp((double)(-dist_sq / (s2[c] * sigma_scale)));
        }
    for (int c = 0; c < 3; ++c) {
        for (int x = 0; x < W; ++x) {
            float s = 0.0f;
            for (int d = -R; d <= R; ++d) if (x + d >= 0 && x + d < W) s += wt[c * stride + d + R];
            wx[c * W + x] = s > 0.0f ? 1.0f / s : 0.0f;
        }
        for (int y = 0; y < H; ++y) {
            float s = 0.0f;
            for (int d = -R; d <= R; ++d) if (y + d >= 0 && y + d < H) s += wt[c * stride + d + R];
            wy[c * H + y] = s > 0.0f ? 1.0f / s : 0.0f;
        }
    }
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_wtab, (size_t)3 * stride * sizeof(float)));
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_wsum_x, (size_t)3 * W * sizeof(float)));
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_wsum_y, (size_t)3 * H * sizeof(float)));
    BHR_CUDA(ctx, cudaMemcpy(ctx->d_wtab, wt, (size_t)3 * stride * sizeof(float), cudaMemcpyHostToDevice));
    BHR_CUDA(ctx, cudaMemcpy(ctx->d_wsum_x, wx, (size_t)3 * W * sizeof(float), cudaMemcpyHostToDevice));
    BHR_CUDA(ctx, cudaMemcpy(ctx->d_wsum_y, wy, (size_t)3 * H * sizeof(float), cudaMemcpyHostToDevice));
    free(wt); free(wx); free(wy);
    return BHR_OK;
}

int bhr_launch_bloom_h(bhr_ctx* ctx, int row0, int row1) {
    if (row1 <= row0) return BHR_OK;
    if (ctx->bloom_tma && !(ctx->bloom_generic & 1)) return bhr_launch_bloom_h_tma(ctx, row0, row1);
    const int R = ctx->bloom_R;
    const int nk = round_up(2 * R + P_OUT, P_OUT);
    const int seg = 31 * P_OUT + nk, pitch = (seg / P_OUT + 1) | 1;
    size_t smem = (size_t)(ctx->wtab_stride + 8 * P_OUT * pitch) * sizeof(float);
    dim3 grid(bhr_div_up(ctx->W, 32 * P_OUT), bhr_div_up(row1 - row0, 8), 3);
    bloom_h_kernel<<<grid, 256, smem, ctx->stream>>>(ctx->disk, ctx->hblur, ctx->W, row0, row1, R, ctx->d_wtab,
                                                     ctx->wtab_stride, ctx->d_wsum_x, (size_t)ctx->W * ctx->H);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// {sum B, sum x*B, sum y*B} of rows [row0, row1) -> ctx->d_flare_sums (device; stays there unless
// the caller asks for it).  Per-block partials summed in block order: the same bits on every run.
int bhr_launch_flare_sums(bhr_ctx* ctx, int row0, int row1) {
    if (row1 <= row0) {
        BHR_CUDA(ctx, cudaMemsetAsync(ctx->d_flare_sums, 0, 3 * sizeof(double), ctx->stream));
        return BHR_OK;
    }
    flare_sums_kernel<<<BHR_FLARE_BLOCKS, 256, 0, ctx->stream>>>(ctx->disk, ctx->W, row0, row1, (size_t)ctx->W * ctx->H,
                                                                 ctx->d_flare_parts);
    flare_sums_final_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_flare_parts, BHR_FLARE_BLOCKS, ctx->d_flare_sums);
    ctx->launches += 2;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// flare_sums_host: {sum B, sum x*B, sum y*B} over the WHOLE frame, or NULL for no flare
namespace {
// Device-side version of the host code below, for the peer path: total = sum of the ranks' partial
// sums in rank order (f64), then the flare parameters of render.py:3929-3942.
__global__ void flare_params_kernel(const double* __restrict__ parts /* [world][3] */, int world, int W, int H,
                                    FlareParams* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int r = 0; r < world; ++r) { s0 += parts[3 * r]; s1 += parts[3 * r + 1]; s2 += parts[3 * r + 2]; }
    FlareParams F;
    F.enabled = 0; F.light_x = F.light_y = F.scx = F.scy = F.scale = F.intensity = F.streak_alpha = F.streak_len = 0.0;
    const float total_f = (float)s0;
    if (!(total_f < 0.01f)) {
        F.enabled = 1;
        F.scale = (double)(W < H ? W : H) / 360.0;
        F.light_x = s1 / (double)total_f;
        F.light_y = s2 / (double)total_f;
        F.scx = W / 2.0; F.scy = H / 2.0;
        const float qf = __fdiv_rn(total_f, (float)((double)(W * H) * 0.3));
        if (1.0f < qf) { F.intensity = 1.0 * 1.5; F.streak_alpha = F.intensity * 0.3; }
        else { const float it = __fmul_rn(qf, 1.5f); F.intensity = it; F.streak_alpha = __fmul_rn(it, 0.3f); }
        F.streak_len = (double)(W < H ? W : H) * 0.4;
    }
    *out = F;
}

}  // namespace

int bhr_launch_flare_params(bhr_ctx* ctx, const double* d_parts, int world, void* d_flare_params) {
    flare_params_kernel<<<1, 32, 0, ctx->stream>>>(d_parts, world, ctx->W, ctx->H, (FlareParams*)d_flare_params);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}
size_t bhr_flare_params_size() { return sizeof(FlareParams); }

// flare_sums_host: {sum B, sum x*B, sum y*B} over the WHOLE frame, or NULL for no flare.
// peer (row-tiled frame over several GPUs, peer.cu): halo rows of the H-blurred layer come from the
// neighbours' buffers (row_src), the flare parameters from device memory, and the finished rows are
// written into rank 0's final buffers.
int bhr_launch_bloom_v_composite_ex(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const double* sums,
                                    const bhr_post_peer* peer) {
    if (row1 <= row0) return BHR_OK;
    const int W = ctx->W, H = ctx->H;
    FlareParams F;
    memset(&F, 0, sizeof(F));
    if (sums) {
        // render.py:3929-3942; total_brightness is np.float32 there, see the oracle for the
        // float32 / float64 mix this reproduces
        float total_f = (float)sums[0];
        if (!(total_f < 0.01f)) {
            F.enabled = 1;
            F.scale = (double)(W < H ? W : H) / 360.0;
            F.light_x = sums[1] / (double)total_f;
            F.light_y = sums[2] / (double)total_f;
            F.scx = W / 2.0; F.scy = H / 2.0;
            float qf = total_f / (float)((double)(W * H) * 0.3);
            if (1.0f < qf) { F.intensity = 1.0 * 1.5; F.streak_alpha = F.intensity * 0.3; }
            else { float it = qf * 1.5f; F.intensity = it; F.streak_alpha = it * 0.3f; }
            F.streak_len = (double)(W < H ? W : H) * 0.4;
        }
    }
    const size_t plane = (size_t)W * H;
    const bool bloom = !(flags & BHR_SKIP_BLOOM);
    if (bloom && ctx->bloom_tma && !(ctx->bloom_generic & 2)) {
        // fused V pass + composite (bloom.cu); everything that must precede stores into the final buffers first
        if (ctx->copy_pending) {
            BHR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
            ctx->copy_pending = 0;
        }
        if (peer && peer->before_composite) {
            int rc = peer->before_composite(ctx);
            if (rc) return rc;
        }
        const void* F_dev = peer ? peer->flare_params : nullptr;
        if (!peer && (flags & BHR_FLARE_FROM_DEVICE) && !sums) {
            int rc = bhr_launch_flare_params(ctx, ctx->d_flare_sums, 1, ctx->d_flare_params_own);
            if (rc) return rc;
            F_dev = ctx->d_flare_params_own;
        }
        return bhr_launch_bloom_v_fused(ctx, flags, row0, row1, &F, F_dev, peer ? peer->row_src : nullptr,
                                        peer && peer->final_f32 ? peer->final_f32 : ctx->final_f32,
                                        peer && peer->final_u8 ? peer->final_u8 : ctx->final_u8);
    }
    if (bloom) {
        const int nk = round_up(2 * ctx->bloom_R + P_OUT, P_OUT);
        size_t smem = ((size_t)ctx->wtab_stride + (size_t)(V_TILE_ROWS - P_OUT + nk) * 32) * sizeof(float);
        if (smem > 48 * 1024) {
            BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_v_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_v_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        dim3 block(32, 8), grid(bhr_div_up(W, 32), bhr_div_up(row1 - row0, V_TILE_ROWS), 3);
        if (peer && peer->row_src)
            bloom_v_kernel<true><<<grid, block, smem, ctx->stream>>>(ctx->hblur, ctx->blur, W, H, row0, row1, ctx->bloom_R,
                                                                    ctx->d_wtab, ctx->wtab_stride, ctx->d_wsum_y, plane, peer->row_src);
        else
            bloom_v_kernel<false><<<grid, block, smem, ctx->stream>>>(ctx->hblur, ctx->blur, W, H, row0, row1, ctx->bloom_R,
                                                                     ctx->d_wtab, ctx->wtab_stride, ctx->d_wsum_y, plane, nullptr);
        BHR_CUDA(ctx, cudaGetLastError());
        ++ctx->launches;
    }
    if (ctx->copy_pending) {      // a frame is still being copied out of the final buffers (bhr_render_async)
        BHR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
        ctx->copy_pending = 0;
    }
    if (peer && peer->before_composite) {
        int rc = peer->before_composite(ctx);
        if (rc) return rc;
    }
    float* dst_f32 = peer && peer->final_f32 ? peer->final_f32 : ctx->final_f32;
    uint8_t* dst_u8 = peer && peer->final_u8 ? peer->final_u8 : ctx->final_u8;
    const FlareParams* F_dev = peer ? (const FlareParams*)peer->flare_params : nullptr;
    if (!peer && (flags & BHR_FLARE_FROM_DEVICE) && !sums) {
        // the frame-wide sums are in ctx->d_flare_sums (single GPU: just reduced; several GPUs: all-reduced
        // in place by the caller): form the flare parameters on the device, no host round trip
        int rc = bhr_launch_flare_params(ctx, ctx->d_flare_sums, 1, ctx->d_flare_params_own);
        if (rc) return rc;
        F_dev = (const FlareParams*)ctx->d_flare_params_own;
    }
    const int cgrid = 148 * 8;
    const float field_gain = (flags & BHR_FIELD_COMPOSITE) ? 0.4f : 0.0f;
#define BHR_COMPOSITE(B, V, FL) composite_kernel<B, V, FL><<<cgrid, 256, 0, ctx->stream>>>( \
        ctx->bg, ctx->disk, ctx->blur, dst_f32, dst_u8, W, row0, row1, plane, F, F_dev, field_gain)
    const bool vec = (W % 4 == 0), fl = F.enabled != 0 || F_dev != nullptr;
    if (vec) {
        if (bloom) { if (fl) BHR_COMPOSITE(true, 4, true); else BHR_COMPOSITE(true, 4, false); }
        else { if (fl) BHR_COMPOSITE(false, 4, true); else BHR_COMPOSITE(false, 4, false); }
    } else {
        if (bloom) { if (fl) BHR_COMPOSITE(true, 1, true); else BHR_COMPOSITE(true, 1, false); }
        else { if (fl) BHR_COMPOSITE(false, 1, true); else BHR_COMPOSITE(false, 1, false); }
    }
#undef BHR_COMPOSITE
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// blur_field on demand: the V pass alone over the whole frame (the fused kernel keeps `blur` in registers)
int bhr_launch_blur_only(bhr_ctx* ctx) {
    const int W = ctx->W, H = ctx->H;
    const int nk = round_up(2 * ctx->bloom_R + P_OUT, P_OUT);
    size_t smem = ((size_t)ctx->wtab_stride + (size_t)(V_TILE_ROWS - P_OUT + nk) * 32) * sizeof(float);
    if (smem > 48 * 1024)
        BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_v_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, 8), grid(bhr_div_up(W, 32), bhr_div_up(H, V_TILE_ROWS), 3);
    bloom_v_kernel<false><<<grid, block, smem, ctx->stream>>>(ctx->hblur, ctx->blur, W, H, 0, H, ctx->bloom_R, ctx->d_wtab,
                                                             ctx->wtab_stride, ctx->d_wsum_y, (size_t)W * H, nullptr);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

int bhr_launch_bloom_v_composite(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const double* sums) {
    return bhr_launch_bloom_v_composite_ex(ctx, flags, row0, row1, sums, nullptr);
}
