// post.cu -- separable chromatic bloom, composite (+ lens flare) and the u8 conversion.
//
// Replaces _bloom_kernel (render.py:3022-3114) as render() drives it (render.py:3914-3918:
// threshold 0, so the live result is the H pass followed by the V pass of the disk layer), the
// host-side numpy composite `clip(img + disk + blur, 0, 1)` + transpose (render.py:3918-3923),
// the numpy lens flare (render.py:3925-4028) and the truncating u8 conversion (render.py:4463).
//
// Layers are planar (3 x H x W).  Both passes keep a sliding window of weights and P running
// outputs in registers so that every loaded sample feeds P FMAs:
//   H pass: one warp per row segment of 32*P outputs, samples staged in shared memory;
//   V pass: lanes along x (coalesced), P consecutive rows per thread, samples straight from
//           L1/L2; the composite, the optional flare and the f32/u8 stores are fused in.
// Taps outside the image are skipped and the sum is renormalised by the in-bounds weight sum,
// which is tabulated per x / per y in the reference's sequential f32 summation order.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int P_OUT = 8;   // outputs per thread

// ---------------------------------------------------------------------------------------------
// H pass.  block = 8 warps; warp w handles row (blockIdx.y * 8 + w), outputs x0 .. x0 + 255.
// Shared row segment: s in [0, 256 + 2R) <-> x = x0 - R + s, stored at s + s / 32 (one pad word
// per 32) so that the stride-8 lane access pattern is bank-conflict free.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bloom_h_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                      int W, int row0, int row1, int R,
                                                      const float* __restrict__ wtab, int wtab_stride,
                                                      const float* __restrict__ wsum_x, size_t plane) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.z;
    const int nk = (2 * R + P_OUT + P_OUT - 1) / P_OUT * P_OUT;   // window positions, padded to P_OUT
    const int seg = 32 * P_OUT - P_OUT + nk;                      // samples any lane may touch
    const int seg_pad = seg + seg / 32 + 1;
    float* wsh = smem;                                  // zero-padded weights of this channel
    float* row = smem + wtab_stride + warp * seg_pad;
    for (int k = threadIdx.x; k < wtab_stride; k += 256) wsh[k] = wtab[ch * wtab_stride + k];
    const int y = row0 + blockIdx.y * 8 + warp;
    const int x0 = blockIdx.x * 256;
    const bool row_ok = y < row1;
    if (row_ok) {
        const float* srow = src + ch * plane + (size_t)y * W;
        for (int s = lane; s < seg; s += 32) {
            int x = x0 - R + s;
            row[s + (s >> 5)] = (x >= 0 && x < W) ? __ldg(srow + x) : 0.0f;
        }
    }
    __syncthreads();
    if (!row_ok) return;

    float acc[P_OUT], wr[P_OUT];
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) { acc[p] = 0.0f; wr[p] = 0.0f; }
    const int base = lane * P_OUT;
    for (int k0 = 0; k0 < nk; k0 += P_OUT) {
#pragma unroll
        for (int kk = 0; kk < P_OUT; ++kk) {
            const int k = k0 + kk;
            const int s = base + k;
            // rotate the weight window: wr[p] = w[k - p]  (zero outside [0, 2R])
#pragma unroll
            for (int p = P_OUT - 1; p > 0; --p) wr[p] = wr[p - 1];
            wr[0] = wsh[k];
            const float v = row[s + (s >> 5)];
#pragma unroll
            for (int p = 0; p < P_OUT; ++p) acc[p] = fmaf(v, wr[p], acc[p]);
        }
    }
    float* drow = dst + ch * plane + (size_t)y * W;
    const float* ws = wsum_x + ch * W;
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) {
        int x = x0 + base + p;
        if (x < W) {
            float wsum = ws[x];
            drow[x] = wsum > 0.0f ? acc[p] / wsum : 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// lens flare, render.py:3925-4028 (float64 like the numpy original, f32 accumulation of `flare`)
// ---------------------------------------------------------------------------------------------
struct FlareParams {
    int enabled;
    double light_x, light_y, scx, scy, scale, intensity, streak_alpha, streak_len;
};

__device__ __forceinline__ double np_mod(double a, double b) {   // numpy.mod for b > 0
    double m = fmod(a, b);
    if (m != 0.0 && m < 0.0) m += b;
    return m;
}

// Every term is zero for most pixels; the squared-distance / slope pre-tests below skip the
// sqrt / atan2 / exp of terms that cannot contribute (a skipped term adds exactly 0, as in numpy).
__device__ void flare_pixel(const FlareParams& F, int x, int y, float fl[3]) {
    const double PI = 3.14159265358979323846;
    const double GUARD = 1.0 + 1e-9;
    fl[0] = fl[1] = fl[2] = 0.0f;
    const double gc[3] = {1.0, 0.9, 0.7};
    for (int g = 0; g < 8; ++g) {
        double t = (g + 1) * 0.15;
        double gx = F.light_x + (F.scx - F.light_x) * t, gy = F.light_y + (F.scy - F.light_y) * t;
        double size = (25 + g * 30) * F.scale;
        double dx = x - gx, dy = y - gy;
        double d2 = dx * dx + dy * dy;
        if (d2 >= size * size * GUARD) continue;
        double dist = sqrt(d2);
        float alpha = 0.0f;
        if (dist < size) { double u = 1 - dist / size; alpha = (float)(u * u * (1 - g * 0.08) * F.intensity); }
        for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + (double)alpha * gc[c]);
    }
    const double rc[3][3] = {{0.3, 0.4, 1.0}, {0.5, 0.5, 0.9}, {0.7, 0.5, 0.8}};
    for (int k = 0; k < 3; ++k) {
        double t = 0.35 + k * 0.15;
        double rx = F.light_x + (F.scx - F.light_x) * t, ry = F.light_y + (F.scy - F.light_y) * t;
        double rr = (60 + k * 40) * F.scale, rw = (6 + k * 3) * F.scale;
        double dx = x - rx, dy = y - ry;
        double d2 = dx * dx + dy * dy;
        double lo = rr - rw, hi = rr + rw;
        if (d2 >= hi * hi * GUARD || (lo > 0 && d2 * GUARD <= lo * lo)) continue;
        double dist = sqrt(d2);
        double u = fmin(fmax(1 - fabs(dist - rr) / rw, 0.0), 1.0);
        double ra = u * u * 0.5 * F.intensity * (1 - k * 0.25);
        for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + ra * rc[k][c]);
    }
    {
        const double hc[3] = {0.6, 0.7, 1.0};
        double hx = F.light_x + (F.scx - F.light_x) * 0.5, hy = F.light_y + (F.scy - F.light_y) * 0.5;
        double hr = 100 * F.scale, hw = 15 * F.scale;
        double dx = x - hx, dy = y - hy;
        double d2 = dx * dx + dy * dy;
        double lo = hr - hw, hi = hr + hw;
        if (!(d2 >= hi * hi * GUARD || (lo > 0 && d2 * GUARD <= lo * lo))) {
            double angle = atan2(dy, dx);
            double dist = sqrt(d2);
            double edge = fabs(np_mod(angle, PI / 3) - PI / 6);
            double hf = fmin(fmax(1 - edge / 0.2, 0.0), 1.0);
            double u = fmin(fmax(1 - fabs(dist - hr) / hw, 0.0), 1.0);
            double ra = u * u * hf * 0.3 * F.intensity;
            for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + ra * hc[c]);
        }
    }
    {
        const double sc[3] = {1.0, 0.95, 0.9};
        const double main_angles[4] = {0.0, PI / 2, PI, 3 * PI / 2};
        double dx = x - F.light_x, dy = y - F.light_y;
        // within 0.05 rad of one of the four axes through the light?  tan(0.05) = 0.050041708
        const double T = 0.0500418 * GUARD;
        if (fabs(dy) <= T * fabs(dx) || fabs(dx) <= T * fabs(dy)) {
            double dist = sqrt(dx * dx + dy * dy);
            double angle = atan2(dy, dx);
            double falloff = exp(-dist / F.streak_len);
            for (int a = 0; a < 4; ++a) {
                double diff = fabs(np_mod(angle - main_angles[a] + PI, 2 * PI) - PI);
                for (int c = 0; c < 3; ++c) {
                    double add = diff < 0.05 ? falloff * F.streak_alpha * sc[c] : 0.0;
                    fl[c] = (float)((double)fl[c] + add);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// V pass + composite.  block = 32 (x) x 8 (row groups) threads, tile = 32 columns x 64 rows.
// Per channel the block stages the (64 + 2R) x 32 samples it needs in shared memory (coalesced
// rows, zero outside the image; 2.2x re-read instead of 10x), then each thread slides its
// 8-output window down the column.  Composite, flare and the f32 / u8 stores are fused in.
// ---------------------------------------------------------------------------------------------
constexpr int V_TILE_ROWS = 8 * P_OUT;   // 64

template <bool BLOOM>
__global__ void __launch_bounds__(256) bloom_v_composite_kernel(
    const float* __restrict__ hblur, const float* __restrict__ bg, const float* __restrict__ disk,
    float* __restrict__ blur_out, float* __restrict__ final_f32, uint8_t* __restrict__ final_u8,
    int W, int H, int row0, int row1, int R, const float* __restrict__ wtab, int wtab_stride,
    const float* __restrict__ wsum_y, size_t plane, FlareParams F) {
    extern __shared__ float vsm[];
    const int nk = (2 * R + P_OUT + P_OUT - 1) / P_OUT * P_OUT;    // window positions, padded
    const int tile_rows = V_TILE_ROWS - P_OUT + nk;                // rows any thread may touch
    float* wsh = vsm;                                              // 3 x wtab_stride weights
    float* tile = vsm + 3 * wtab_stride;                           // tile_rows x 32
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (BLOOM)
        for (int k = tid; k < 3 * wtab_stride; k += 256) wsh[k] = wtab[k];
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int ty0 = row0 + blockIdx.y * V_TILE_ROWS;               // first output row of the tile
    const int y0 = ty0 + threadIdx.y * P_OUT;                      // first output row of this thread
    const bool col_ok = x < W;

    float val[3][P_OUT];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc[P_OUT], wr[P_OUT];
#pragma unroll
        for (int p = 0; p < P_OUT; ++p) { acc[p] = 0.0f; wr[p] = 0.0f; }
        if (BLOOM) {
            __syncthreads();                                       // previous channel's tile is consumed
            const float* src = hblur + c * plane;
            for (int r = threadIdx.y; r < tile_rows; r += 8) {
                const int y = ty0 - R + r;
                tile[r * 32 + threadIdx.x] = (col_ok && y >= 0 && y < H) ? __ldg(src + (size_t)y * W + x) : 0.0f;
            }
            __syncthreads();
            const float* wc = wsh + c * wtab_stride;
            const float* colp = tile + threadIdx.y * P_OUT * 32 + threadIdx.x;
            for (int k0 = 0; k0 < nk; k0 += P_OUT) {
#pragma unroll
                for (int kk = 0; kk < P_OUT; ++kk) {
                    const int k = k0 + kk;
#pragma unroll
                    for (int p = P_OUT - 1; p > 0; --p) wr[p] = wr[p - 1];
                    wr[0] = wc[k];
                    const float v = colp[k * 32];
#pragma unroll
                    for (int p = 0; p < P_OUT; ++p) acc[p] = fmaf(v, wr[p], acc[p]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < P_OUT; ++p) {
            const int y = y0 + p;
            float b = 0.0f;
            val[c][p] = 0.0f;
            if (col_ok && y < row1) {
                const size_t o = c * plane + (size_t)y * W + x;
                float v = bg[o] + disk[o];
                if (BLOOM) {
                    float wsum = wsum_y[c * H + y];
                    b = wsum > 0.0f ? acc[p] / wsum : 0.0f;
                    if (blur_out) blur_out[o] = b;
                    v = v + b;
                }
                val[c][p] = fminf(fmaxf(v, 0.0f), 1.0f);
            }
        }
    }
    if (!col_ok) return;
#pragma unroll
    for (int p = 0; p < P_OUT; ++p) {
        const int y = y0 + p;
        if (y >= row1) break;
        float r = val[0][p], g = val[1][p], b = val[2][p];
        if (F.enabled) {
            float fl[3];
            flare_pixel(F, x, y, fl);
            r = fminf(fmaxf(r + fl[0], 0.0f), 1.0f);
            g = fminf(fmaxf(g + fl[1], 0.0f), 1.0f);
            b = fminf(fmaxf(b + fl[2], 0.0f), 1.0f);
        }
        const size_t o = ((size_t)y * W + x) * 3;
        final_f32[o] = r; final_f32[o + 1] = g; final_f32[o + 2] = b;
        final_u8[o] = (uint8_t)(r * 255.0f); final_u8[o + 1] = (uint8_t)(g * 255.0f); final_u8[o + 2] = (uint8_t)(b * 255.0f);
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
        atomicAdd(sums + threadIdx.x, t);
    }
}

}  // namespace

// weights exp(-d^2 / (sigma2 * sigma_scale)) (render.py:3058-3060) evaluated like the oracle
// (double exp rounded once), plus in-bounds weight sums in the reference's tap order
int bhr_setup_bloom_tables(bhr_ctx* ctx) {
    const int W = ctx->W, H = ctx->H;
    const int R = (int)(W * 0.02);                          // render.py:3914
    const float sigma_scale = (float)((W / 640.0) * (W / 640.0));   // render.py:3915 (f64 -> f32 kernel arg)
    ctx->bloom_R = R; ctx->sigma_scale = sigma_scale;
    const int taps = 2 * R + 1;
    const int stride = (taps + 2 * P_OUT) / P_OUT * P_OUT;  // zero padded past the last tap
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
            wx[c * W + x] = s;
        }
        for (int y = 0; y < H; ++y) {
            float s = 0.0f;
            for (int d = -R; d <= R; ++d) if (y + d >= 0 && y + d < H) s += wt[c * stride + d + R];
            wy[c * H + y] = s;
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
    const int R = ctx->bloom_R;
    const int nk = (2 * R + P_OUT + P_OUT - 1) / P_OUT * P_OUT;
    const int seg = 32 * P_OUT - P_OUT + nk, seg_pad = seg + seg / 32 + 1;
    size_t smem = (size_t)(ctx->wtab_stride + 8 * seg_pad) * sizeof(float);
    dim3 grid(bhr_div_up(ctx->W, 256), bhr_div_up(row1 - row0, 8), 3);
    bloom_h_kernel<<<grid, 256, smem, ctx->stream>>>(ctx->disk, ctx->hblur, ctx->W, row0, row1, R, ctx->d_wtab,
                                                     ctx->wtab_stride, ctx->d_wsum_x, (size_t)ctx->W * ctx->H);
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

int bhr_launch_flare_sums(bhr_ctx* ctx, int row0, int row1) {
    BHR_CUDA(ctx, cudaMemsetAsync(ctx->d_flare_sums, 0, 3 * sizeof(double), ctx->stream));
    if (row1 <= row0) return BHR_OK;
    flare_sums_kernel<<<592, 256, 0, ctx->stream>>>(ctx->disk, ctx->W, row0, row1, (size_t)ctx->W * ctx->H,
                                                    ctx->d_flare_sums);
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// flare_sums_host: {sum B, sum x*B, sum y*B} over the WHOLE frame, or NULL for no flare
int bhr_launch_bloom_v_composite(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const double* sums) {
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
    dim3 block(32, 8);
    dim3 grid(bhr_div_up(W, 32), bhr_div_up(row1 - row0, 8 * P_OUT));
    const size_t plane = (size_t)W * H;
    if (flags & BHR_SKIP_BLOOM) {
        bloom_v_composite_kernel<false><<<grid, block, 0, ctx->stream>>>(
            ctx->hblur, ctx->bg, ctx->disk, nullptr, ctx->final_f32, ctx->final_u8, W, H, row0, row1, ctx->bloom_R,
            ctx->d_wtab, ctx->wtab_stride, ctx->d_wsum_y, plane, F);
    } else {
        const int nk = (2 * ctx->bloom_R + P_OUT + P_OUT - 1) / P_OUT * P_OUT;
        size_t smem = ((size_t)3 * ctx->wtab_stride + (size_t)(V_TILE_ROWS - P_OUT + nk) * 32) * sizeof(float);
        if (smem > 48 * 1024)
            BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_v_composite_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bloom_v_composite_kernel<true><<<grid, block, smem, ctx->stream>>>(
            ctx->hblur, ctx->bg, ctx->disk, ctx->blur, ctx->final_f32, ctx->final_u8, W, H, row0, row1, ctx->bloom_R,
            ctx->d_wtab, ctx->wtab_stride, ctx->d_wsum_y, plane, F);
    }
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}
