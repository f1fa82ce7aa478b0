// stats.cu -- normalisation statistics of the disk-texture pipeline on the device.
//
// Replaces the numpy body of recompute_interactive_stats (render.py:3655-3712), which reads the
// whole component field back (63 MB at fhd) and runs np.percentile / np.quantile on the host
// (~50 ms per call).  The device computes exact ORDER STATISTICS -- the two neighbours of each
// quantile's virtual index -- and the host applies numpy's own interpolation formula to those
// neighbours, so the results are the reference's numbers, not an approximation:
//   stats_prepare : density and structure-temperature planes in numpy's float32 operation order,
//                   count of positive structure texels, per-row max of the base temperature;
//   stats_select  : radix select (4 passes of 8 bits on order-preserving keys) of two ranks at
//                   once: density over all texels, structure over the positive texels;
//   stats_rows    : per row, sort the clipped / scaled structure values in shared memory and
//                   return the row maximum and the two neighbours of the 70 % quantile.
#include <math_constants.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float m_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float a_(float a, float b) { return __fadd_rn(a, b); }

// order-preserving map float -> uint32 (and back)
__device__ __forceinline__ unsigned f2key(float x) {
    const unsigned b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// density = (0.15 + 0.10 sp + 0.30 turb + 0.20 hs + 0.30 arc + rt_w rt) * dm * edge[row]
// struct  = (sp_t + turb_t + arc_t + rt_t + hs_t) * dm        (render.py:3671-3684; numpy evaluates
// every operator in float32, left to right)
__global__ void __launch_bounds__(256) stats_prepare_kernel(const float* __restrict__ comp, const float* __restrict__ edge,
                                                            int n_r, int n_phi, float rt_w, float* __restrict__ dens,
                                                            float* __restrict__ strc, float* __restrict__ row_tb_max,
                                                            unsigned long long* __restrict__ n_positive) {
    const int ri = blockIdx.x;
    const size_t plane = (size_t)n_r * n_phi;
    const float e = edge[ri];
    float tb_max = -CUDART_INF_F;
    unsigned pos = 0;
    for (int pi = threadIdx.x; pi < n_phi; pi += 256) {
        const size_t o = (size_t)ri * n_phi + pi;
        const float dm = comp[12 * plane + o];
        float d = a_(0.15f, m_(0.10f, comp[1 * plane + o]));
        d = a_(d, m_(0.30f, comp[3 * plane + o]));
        d = a_(d, m_(0.20f, comp[9 * plane + o]));
        d = a_(d, m_(0.30f, comp[5 * plane + o]));
        d = a_(d, m_(rt_w, comp[7 * plane + o]));
        dens[o] = m_(m_(d, dm), e);
        float s = a_(comp[2 * plane + o], comp[4 * plane + o]);
        s = a_(s, comp[6 * plane + o]);
        s = a_(s, comp[8 * plane + o]);
        s = a_(s, comp[10 * plane + o]);
        s = m_(s, dm);
        strc[o] = s;
        pos += s > 0.0f ? 1u : 0u;
        tb_max = fmaxf(tb_max, comp[o]);
    }
    __shared__ float sh_max[8];
    __shared__ unsigned sh_pos[8];
    for (int off = 16; off > 0; off >>= 1) {
        tb_max = fmaxf(tb_max, __shfl_down_sync(0xffffffffu, tb_max, off));
        pos += __shfl_down_sync(0xffffffffu, pos, off);
    }
    if ((threadIdx.x & 31) == 0) { sh_max[threadIdx.x >> 5] = tb_max; sh_pos[threadIdx.x >> 5] = pos; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { tb_max = fmaxf(tb_max, sh_max[w]); pos += sh_pos[w]; }
        row_tb_max[ri] = tb_max;
        if (pos) atomicAdd(n_positive, (unsigned long long)pos);
    }
}

// select state of the two simultaneous problems (0: density, all texels; 1: structure, positive)
struct SelectState {
    unsigned prefix[2];              // key bits decided so far
    unsigned long long rank[2];      // rank still to descend inside the current prefix
    unsigned hist[2][256];
    unsigned same_next[2];           // 1: the next order statistic equals the selected one
    unsigned min_above[2];           // smallest eligible key above the selected one
};

__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ dens, const float* __restrict__ strc,
                                                          size_t n, int shift, SelectState* __restrict__ st) {
    __shared__ unsigned h[2][256];
    h[0][threadIdx.x] = 0; h[1][threadIdx.x] = 0;
    __syncthreads();
    const unsigned p0 = st->prefix[0], p1 = st->prefix[1];
    const unsigned hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const unsigned k0 = f2key(dens[i]);
        if ((k0 & hi_mask) == p0) atomicAdd(&h[0][(k0 >> shift) & 255u], 1u);
        const float s = strc[i];
        if (s > 0.0f) {
            const unsigned k1 = f2key(s);
            if ((k1 & hi_mask) == p1) atomicAdd(&h[1][(k1 >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (h[0][threadIdx.x]) atomicAdd(&st->hist[0][threadIdx.x], h[0][threadIdx.x]);
    if (h[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], h[1][threadIdx.x]);
}

// one thread per problem walks the 256 bins (tiny), fixes the next 8 key bits and clears the bins
__global__ void select_scan_kernel(int shift, SelectState* __restrict__ st) {
    const int p = threadIdx.x;
    if (p >= 2) return;
    unsigned long long rank = st->rank[p], cum = 0;
    int d = 0;
    for (; d < 256; ++d) {
        const unsigned c = st->hist[p][d];
        if (rank < cum + c) break;
        cum += c;
    }
    if (d == 256) d = 255;        // (empty problem: harmless)
    const unsigned c = st->hist[p][d];
    st->prefix[p] |= (unsigned)d << shift;
    st->rank[p] = rank - cum;
    if (shift == 0) st->same_next[p] = (rank - cum + 1 < c) ? 1u : 0u;
    for (int k = 0; k < 256; ++k) st->hist[p][k] = 0;
}

__global__ void __launch_bounds__(256) select_next_kernel(const float* __restrict__ dens, const float* __restrict__ strc,
                                                          size_t n, SelectState* __restrict__ st) {
    const unsigned p0 = st->prefix[0], p1 = st->prefix[1];
    unsigned m0 = 0xffffffffu, m1 = 0xffffffffu;
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const unsigned k0 = f2key(dens[i]);
        if (k0 > p0) m0 = min(m0, k0);
        const float s = strc[i];
        if (s > 0.0f) { const unsigned k1 = f2key(s); if (k1 > p1) m1 = min(m1, k1); }
    }
    m0 = __reduce_min_sync(0xffffffffu, m0);
    m1 = __reduce_min_sync(0xffffffffu, m1);
    if ((threadIdx.x & 31) == 0) { atomicMin(&st->min_above[0], m0); atomicMin(&st->min_above[1], m1); }
}

__global__ void select_result_kernel(const SelectState* __restrict__ st, float* __restrict__ out) {
    const int p = threadIdx.x;
    if (p >= 2) return;
    const float v = key2f(st->prefix[p]);
    out[2 * p] = v;
    out[2 * p + 1] = (st->same_next[p] || st->min_above[p] == 0xffffffffu) ? v : key2f(st->min_above[p]);
}

// per row: scaled = clip(struct / denom * 0.8, 0, 1.2) (render.py:3692), sorted ascending in
// shared memory (bitonic, padded with +inf); out[row] = {max, sorted[lo], sorted[hi]}
__global__ void __launch_bounds__(256) stats_rows_kernel(const float* __restrict__ strc, int n_phi, int n_pad, float denom,
                                                         int lo, int hi, float* __restrict__ out) {
    extern __shared__ float v[];
    const int ri = blockIdx.x;
    for (int i = threadIdx.x; i < n_pad; i += 256) {
        float x = CUDART_INF_F;
        if (i < n_phi) x = fminf(fmaxf(m_(__fdiv_rn(strc[(size_t)ri * n_phi + i], denom), 0.8f), 0.0f), 1.2f);
        v[i] = x;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += 256) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = v[i], b = v[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { v[i] = b; v[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        out[3 * ri] = v[n_phi - 1];
        out[3 * ri + 1] = v[lo];
        out[3 * ri + 2] = v[hi];
    }
}

}  // namespace

static int ensure_stats_storage(bhr_ctx* ctx) {
    if (ctx->stats_scratch) return BHR_OK;
    const size_t plane = (size_t)ctx->n_r * ctx->n_phi;
    // [dens plane][struct plane][row out: 4 * n_r floats][select result: 4 floats] + state
    BHR_CUDA(ctx, cudaMalloc(&ctx->stats_scratch, (2 * plane + 4 * (size_t)ctx->n_r + 4) * sizeof(float)));
    BHR_CUDA(ctx, cudaMalloc(&ctx->stats_state, sizeof(SelectState) + sizeof(unsigned long long)));
    return BHR_OK;
}

extern "C" int bhr_stats_prepare(bhr_ctx* ctx, int enable_rt, uint64_t* n_total, uint64_t* n_positive) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !n_total || !n_positive) return BHR_ERR_INVALID;
    if (!ctx->bg_ready) BHR_FAIL(ctx, BHR_ERR_STATE, "Must call init_background_layer() first");
    int rc = ensure_stats_storage(ctx);
    if (rc) return rc;
    if ((rc = bhr_join_entities(ctx))) return rc;
    const size_t plane = (size_t)ctx->n_r * ctx->n_phi;
    float* dens = ctx->stats_scratch;
    float* strc = dens + plane;
    float* row_out = strc + plane;
    unsigned long long* d_pos = (unsigned long long*)((char*)ctx->stats_state + sizeof(SelectState));
    BHR_CUDA(ctx, cudaMemsetAsync(d_pos, 0, sizeof(unsigned long long), ctx->stream));
    stats_prepare_kernel<<<ctx->n_r, 256, 0, ctx->stream>>>(ctx->comp, ctx->edge, ctx->n_r, ctx->n_phi,
                                                            enable_rt ? 0.20f : 0.0f, dens, strc,
                                                            row_out + 3 * (size_t)ctx->n_r, d_pos);
    if ((rc = bhr_mark_comp_read(ctx))) return rc;
    BHR_CUDA(ctx, cudaGetLastError());
    unsigned long long pos = 0;
    BHR_CUDA(ctx, cudaMemcpyAsync(&pos, d_pos, sizeof(pos), cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_total = plane;
    *n_positive = pos;
    return BHR_OK;
}

extern "C" int bhr_stats_select(bhr_ctx* ctx, uint64_t rank_density, uint64_t rank_struct, float out[4]) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    if (!ctx->stats_scratch) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_stats_prepare has not run");
    const size_t plane = (size_t)ctx->n_r * ctx->n_phi;
    const float* dens = ctx->stats_scratch;
    const float* strc = dens + plane;
    float* d_out = ctx->stats_scratch + 2 * plane + 4 * (size_t)ctx->n_r;
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.rank[0] = rank_density; init.rank[1] = rank_struct;
    init.min_above[0] = init.min_above[1] = 0xffffffffu;
    SelectState* st = (SelectState*)ctx->stats_state;
    BHR_CUDA(ctx, cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = ctx->num_sms * 4;
    for (int shift = 24; shift >= 0; shift -= 8) {
        select_hist_kernel<<<grid, 256, 0, ctx->stream>>>(dens, strc, plane, shift, st);
        select_scan_kernel<<<1, 32, 0, ctx->stream>>>(shift, st);
    }
    select_next_kernel<<<grid, 256, 0, ctx->stream>>>(dens, strc, plane, st);
    select_result_kernel<<<1, 32, 0, ctx->stream>>>(st, d_out);
    BHR_CUDA(ctx, cudaGetLastError());
    BHR_CUDA(ctx, cudaMemcpyAsync(out, d_out, 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_stats_rows(bhr_ctx* ctx, float denom, int lo, int hi, float* out) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    if (!ctx->stats_scratch) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_stats_prepare has not run");
    if (lo < 0 || hi < lo || hi >= ctx->n_phi) BHR_FAIL(ctx, BHR_ERR_INVALID, "quantile neighbours out of range");
    const size_t plane = (size_t)ctx->n_r * ctx->n_phi;
    const float* strc = ctx->stats_scratch + plane;
    float* row_out = ctx->stats_scratch + 2 * plane;
    int n_pad = 1;
    while (n_pad < ctx->n_phi) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(float);
    if (smem > 200 * 1024) BHR_FAIL(ctx, BHR_ERR_INVALID, "n_phi too large for the in-shared-memory row sort");
    if (smem > 48 * 1024)
        BHR_CUDA(ctx, cudaFuncSetAttribute(stats_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stats_rows_kernel<<<ctx->n_r, 256, smem, ctx->stream>>>(strc, ctx->n_phi, n_pad, denom, lo, hi, row_out);
    BHR_CUDA(ctx, cudaGetLastError());
    // out (n_r, 4): max of scaled, sorted[lo], sorted[hi], max of the base temperature
    BHR_CUDA(ctx, cudaMemcpy2DAsync(out, 4 * sizeof(float), row_out, 3 * sizeof(float), 3 * sizeof(float), ctx->n_r,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaMemcpy2DAsync(out + 3, 4 * sizeof(float), row_out + 3 * (size_t)ctx->n_r, sizeof(float),
                                    sizeof(float), ctx->n_r, cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}
