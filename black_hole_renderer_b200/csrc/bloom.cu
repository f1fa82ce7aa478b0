// bloom.cu -- the separable chromatic bloom as two TMA-fed, FFMA2-bound stencil kernels, the second
// one fused with the composite (+ lens flare) and the u8 conversion.
//
// Replaces _bloom_kernel as render() drives it (render.py:3022-3114, 3914-3918: threshold 0, so
// the live result is the H pass followed by the V pass of the disk layer) and the host-side
// composite `clip(img + disk + blur, 0, 1)` + transpose + u8 truncation (render.py:3918-3923, 4463).
//
// Arithmetic per output and channel: 2R + 1 taps, summed in tap order with one FMA each (the order
// of the reference's `while dx <= kernel_radius` loop), then one multiplication by the reciprocal
// of the in-bounds weight sum.  Taps outside the image read zeros (TMA fills out-of-bounds box
// elements with zeros), which is the same as skipping them.
//
// Formulation: tap-outer, output-inner.  A thread owns P consecutive outputs along the blurred
// axis of TWO adjacent lines of the other axis, packed in the two lanes of f32x2 registers:
//     acc[p] += w[j] * window[(j + p) mod P]        j = 0 .. 2R,  p = 0 .. P-1
// with a sliding window of P sample pairs in registers (one new pair enters per tap, register
// rotation resolved at compile time by unrolling P taps).  Every FFMA2 does two useful FMAs -- there
// are no zero-weight positions to pad or peel -- and a tap costs P FFMA2 + one (H: two) shared-memory
// load + one broadcast weight load, so the FMA pipe, not the issue slots, is the limit
// (profiles/r01_microbench_fp32.txt: a pure FFMA2 stream runs at 74 TFLOP/s in half the issue slots).
//   H pass: the packed pair is (row y, row y + 1); lanes own consecutive runs of P = 15 (or 5)
//           pixels of a row, so lane strides are odd and the row-major tile is read conflict-free.
//           Every warp is its own double-buffered pipeline: row segments arrive as 1-row TMA boxes
//           (boxes start at multiples of 4 pixels: a box whose first element is not 16-byte aligned
//           in global memory raises "illegal instruction" on this device -- measured, sizes with
//           R mod 4 != 0 faulted until the left margin was rounded up).
//   V pass: the packed pair is (column x, column x + 1), i.e. one aligned LDS.64 of the row-major
//           tile; a block shares a 64-column x (80 + 2R)-row tile (one TMA box) between 8 warps.
//           The three channels of a tile are blurred one after the other (next channel's tile in
//           flight meanwhile), each is composited at once -- clamp(bg + disk + blur) -- and after
//           the third one the RGB pixels leave through a per-warp staging row as full-line stores of
//           the (H, W, 3) f32 and u8 frames.  `blur` never goes to memory (option "keep_blur").
#include <cuda.h>

#include "common.cuh"
#include "post.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (PTX ISA 8.x: cp.async.bulk.tensor, mbarrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "LAB_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one box of the (x, y, channel) tensor -> shared memory; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

__device__ __forceinline__ float2 splat2(float s) { return make_float2(s, s); }
// The stencil of one thread: P packed outputs, `ntaps` taps.  `ld(base, k)` returns sample pair base + k of the
// thread's window (k a compile-time offset, base advanced once per P taps, so every load is pointer + immediate);
// `wsh[j]` is the weight of tap j.  Samples and weights are fetched one tap ahead of their use: a tap's new
// sample is needed by its last FFMA2 already, and a shared-memory load takes longer than the P - 1 before it.
// Reads one sample and one weight past the window (never used; the callers' buffers have the room).
template <int P, typename Base, typename Load>
__device__ __forceinline__ void stencil(float2 (&acc)[P], Base base, const Load& ld, const float* __restrict__ wsh, const int ntaps) {
    float2 win[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int p = 0; p < P - 1; ++p) win[p] = ld(base, p);
    float2 pre = ld(base, P - 1);
    float wpre = wsh[0];
    int j = 0;
    for (; j + P <= ntaps; j += P) {
#pragma unroll
        for (int u = 0; u < P; ++u) {
            win[(u + P - 1) % P] = pre;
            const float w = wpre;
            pre = ld(base, u + P);
            wpre = wsh[j + u + 1];
#pragma unroll
            for (int p = 0; p < P; ++p) acc[p] = __ffma2_rn(win[(u + p) % P], make_float2(w, w), acc[p]);
        }
        base = base + P;
    }
    // the last ntaps mod P taps (the rotation is back at offset 0 here)
#pragma unroll
    for (int u = 0; u < P - 1; ++u) {
        if (j + u >= ntaps) break;
        win[(u + P - 1) % P] = pre;
        const float w = wpre;
        pre = ld(base, u + P);
        wpre = wsh[j + u + 1];
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = __ffma2_rn(win[(u + p) % P], make_float2(w, w), acc[p]);
    }
}

// window bases of the two passes: H = one float pointer per packed row, V = a float2 pointer with a row stride
struct HBase { const float *r0, *r1; __device__ __forceinline__ HBase operator+(int k) const { return {r0 + k, r1 + k}; } };
template <int STRIDE> struct VBase { const float2* c; __device__ __forceinline__ VBase operator+(int k) const { return {c + k * STRIDE}; } };

// ---------------------------------------------------------------------------------------------
// H pass.  Warp task = (channel, row pair, run of 32 P pixels).  Per warp: two stages of 2 rows x
// segp samples (sample s <-> x = x0 - xlead + s), filled by nb one-row TMA boxes of bw samples per row.
// ---------------------------------------------------------------------------------------------
struct HArgs {
    float* dst;                      // hblur, planar 3 x (H, W)
    const float* wtab; int wtab_stride;
    const float* wsum_x;             // 3 x W reciprocal in-bounds weight sums
    int W, H, row0, row1, R;
    int bw, nb, segp;                // box width, boxes per row, samples per staged row (= nb * bw)
    int xlead;                       // samples staged left of x0: R rounded up to a multiple of 4
    size_t plane;
};

template <int P>
__global__ void __launch_bounds__(256, P <= 10 ? 3 : 2) bloom_h_tma_kernel(const __grid_constant__ CUtensorMap src_map, const HArgs a) {
    extern __shared__ __align__(128) unsigned char hsm[];
    constexpr int TX = 32 * P;
    // (the shuffle tells the compiler that the warp index is warp-uniform: the TMA operands derived from it then
    // live in uniform registers instead of being broadcast lane by lane in front of every UTMALDG)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int segp = a.segp, ntaps = 2 * a.R + 1;
    float* const my = reinterpret_cast<float*>(hsm) + (size_t)warp * (4 * segp);     // [stage][row][segp]
    float* const wsh = reinterpret_cast<float*>(hsm) + (size_t)8 * (4 * segp);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(wsh + ((3 * a.wtab_stride + 1) & ~1)) + warp * 2;
    for (int k = threadIdx.x; k < 3 * a.wtab_stride; k += 256) wsh[k] = a.wtab[k];
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    fence_mbar_init();
    __syncthreads();

    const int tiles_x = (a.W + TX - 1) / TX, n_pairs = (a.row1 - a.row0 + 1) >> 1;
    const int n_tasks = 3 * n_pairs * tiles_x;
    auto decode = [&](int task, int& ch, int& y, int& x0) {
        const int tx = task % tiles_x, t2 = task / tiles_x;
        x0 = tx * TX; y = a.row0 + 2 * (t2 % n_pairs); ch = t2 / n_pairs;
    };
    auto issue = [&](int task, int stage) {          // lane 0 only
        int ch, y, x0;
        decode(task, ch, y, x0);
        float* st = my + stage * (2 * segp);
        mbar_expect_tx(&bars[stage], (uint32_t)(2 * a.nb * a.bw * sizeof(float)));
        for (int r = 0; r < 2; ++r)
            for (int b = 0; b < a.nb; ++b)
                tma_load_3d(st + r * segp + b * a.bw, &src_map, &bars[stage], x0 - a.xlead + b * a.bw, y + r, ch);
    };

    // Static, SM-balanced assignment.  Every warp of the grid takes `base` tasks; the n_tasks mod (warps of the
    // grid) left over are dealt block by block (left-over task j -> block j mod gridDim, warp j / gridDim), so
    // that every SM -- it hosts the same number of blocks -- gets the same share of them.
    const int total_warps = gridDim.x * 8;
    const int base = n_tasks / total_warps, left = n_tasks - base * total_warps;
    const int my_left = warp * (int)gridDim.x + (int)blockIdx.x;
    const int my_n = base + (my_left < left ? 1 : 0);
    auto task_of = [&](int i) { return i < base ? i * total_warps + (int)blockIdx.x * 8 + warp : base * total_warps + my_left; };
    if (my_n > 0 && lane == 0) issue(task_of(0), 0);
    for (int it = 0; it < my_n; ++it) {
        const int stage = it & 1;
        const int task = task_of(it);
        if (it + 1 < my_n) {
            // the other stage was read and re-written (output staging) by this warp in the previous
            // iteration: order those generic-proxy accesses before the async-proxy refill
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) issue(task_of(it + 1), stage ^ 1);
        }
        mbar_wait(&bars[stage], (uint32_t)((it >> 1) & 1));
        int ch, y, x0;
        decode(task, ch, y, x0);
        float* st = my + stage * (2 * segp);
        const float* s0 = st + lane * P + (a.xlead - a.R);
        // 1 / in-bounds weight sum of the pixels this lane stores below (0 if none): in flight during the stencil
        float4 wsx[(TX / 4 + 31) / 32];
#pragma unroll
        for (int q = 0; q < (TX / 4 + 31) / 32; ++q) {
            const int xq = x0 + 4 * (lane + 32 * q);
            wsx[q] = __ldg(reinterpret_cast<const float4*>(a.wsum_x + (size_t)ch * a.W + (xq < a.W ? xq : 0)));
        }
        float2 acc[P];
        stencil<P>(acc, HBase{s0, s0 + segp}, [](const HBase& b, int k) { return make_float2(b.r0[k], b.r1[k]); },
                   wsh + ch * a.wtab_stride, ntaps);
        // through the (consumed) stage rows for full-line stores
        __syncwarp();
#pragma unroll
        for (int p = 0; p < P; ++p) { st[lane * P + p] = acc[p].x; st[segp + lane * P + p] = acc[p].y; }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (y + r >= a.row1) break;
            float* drow = a.dst + ch * a.plane + (size_t)(y + r) * a.W;
#pragma unroll
            for (int q = 0; q < (TX / 4 + 31) / 32; ++q) {
                const int i = lane + 32 * q, x = x0 + 4 * i;
                if (i < TX / 4 && x < a.W) {
                    float4 v = reinterpret_cast<const float4*>(st + r * segp)[i];
                    v.x *= wsx[q].x; v.y *= wsx[q].y; v.z *= wsx[q].z; v.w *= wsx[q].w;
                    *reinterpret_cast<float4*>(drow + x) = v;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// V pass + composite (+ flare) + u8.  Block tile = 64 columns x 8 P rows; unit of the pipeline =
// (tile, channel); the (8 P + 2R) x 64 samples of a unit are one TMA box.
// ---------------------------------------------------------------------------------------------
constexpr int VCOLS = 64;

struct VArgs {
    const float *bg, *disk;          // planar 3 x (H, W)
    float* blur;                     // planar, written when want_blur
    float* final_f32; uint8_t* final_u8;   // (H, W, 3); may live in a peer's memory
    const float* wtab; int wtab_stride;
    const float* wsum_y;             // 3 x H
    int W, H, row0, row1, R;
    size_t plane;
    float field_gain;                // 0.4 for render_to_field's compositing, else 0
    int want_blur;
};

// STAGED: the unit's bg / disk tiles arrive by TMA with its samples (one block of 8 warps per SM, everything the
// finalisation reads is in shared memory); !STAGED: they are read with one batch of global loads per channel after
// the stencil, which leaves room for two blocks per SM -- the other block's warps cover the batch's latency.
template <int P, bool STAGED>
__global__ void __launch_bounds__(256, STAGED ? 1 : 2) bloom_v_fused_kernel(const __grid_constant__ CUtensorMap hb_map,
                                                               const __grid_constant__ CUtensorMap bg_map,
                                                               const __grid_constant__ CUtensorMap disk_map, const VArgs a) {
    extern __shared__ __align__(128) unsigned char vsm[];
    constexpr int TR = 8 * P;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int rows = TR + 2 * a.R, ntaps = 2 * a.R + 1;
    // one stage = the unit's H-blurred samples (rows x 64) + its bg and disk tiles (TR x 64 each)
    const size_t hb_floats = (size_t)rows * VCOLS, lay_floats = STAGED ? (size_t)TR * VCOLS : 0;
    const size_t stage_floats = hb_floats + 2 * lay_floats;
    float* const stages = reinterpret_cast<float*>(vsm);
    float* const wsh = stages + 2 * stage_floats;
    float* const ostage = wsh + ((3 * a.wtab_stride + 3) & ~3) + warp * (3 * VCOLS);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(wsh + ((3 * a.wtab_stride + 3) & ~3) + 8 * (3 * VCOLS));
    for (int k = threadIdx.x; k < 3 * a.wtab_stride; k += 256) wsh[k] = a.wtab[k];
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    fence_mbar_init();
    __syncthreads();

    const int tiles_x = (a.W + VCOLS - 1) / VCOLS, tiles_y = (a.row1 - a.row0 + TR - 1) / TR;
    const int n_tiles = tiles_x * tiles_y;
    const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int n_units = 3 * my_tiles;
    auto issue = [&](int u) {                        // thread 0 only
        const int tile = blockIdx.x + (u / 3) * gridDim.x, ch = u % 3;
        const int tx = tile % tiles_x, ty = tile / tiles_x;
        float* st = stages + (u & 1) * stage_floats;
        mbar_expect_tx(&bars[u & 1], (uint32_t)(stage_floats * sizeof(float)));
        tma_load_3d(st, &hb_map, &bars[u & 1], tx * VCOLS, a.row0 + ty * TR - a.R, ch);
        if (STAGED) {
            tma_load_3d(st + hb_floats, &bg_map, &bars[u & 1], tx * VCOLS, a.row0 + ty * TR, ch);
            tma_load_3d(st + hb_floats + lay_floats, &disk_map, &bars[u & 1], tx * VCOLS, a.row0 + ty * TR, ch);
        }
    };
    if (n_units > 0 && threadIdx.x == 0) issue(0);
    float2 rgb[3][P];
    for (int k = 0; k < my_tiles; ++k) {
        const int tile = blockIdx.x + k * gridDim.x;
        const int tx = tile % tiles_x, ty = tile / tiles_x;
        const int x = tx * VCOLS + 2 * lane, y0 = a.row0 + ty * TR + warp * P;
        const bool col_ok = x < a.W;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {             // (static channel index: rgb[][] stays in registers)
        const int u = 3 * k + ch;
        if (u + 1 < n_units && threadIdx.x == 0) issue(u + 1);       // (its stage was released by the barrier below)
        // 1 / in-bounds weight sum of my rows (0 if none): in flight while the stencil runs
        float ws[P];
#pragma unroll
        for (int p = 0; p < P; ++p) ws[p] = __ldg(a.wsum_y + ch * a.H + min(y0 + p, a.H - 1));
        mbar_wait(&bars[u & 1], (uint32_t)((u >> 1) & 1));
        const float* st = stages + (u & 1) * stage_floats;
        const float2* col = reinterpret_cast<const float2*>(st + (size_t)(warp * P) * VCOLS) + lane;
        float2 acc[P];
        stencil<P>(acc, VBase<VCOLS / 2>{col}, [](const VBase<VCOLS / 2>& b, int k) { return b.c[k * (VCOLS / 2)]; },
                   wsh + ch * a.wtab_stride, ntaps);
        // ---- this channel of the composite: clamp(bg + disk + blur), render.py:3918 (3857-3863 with field_gain) ----
        float2 gl[P], dl[P];
        if (STAGED) {
            const float2* gcol = reinterpret_cast<const float2*>(st + hb_floats + (size_t)(warp * P) * VCOLS) + lane;
            const float2* dcol = gcol + lay_floats / 2;
#pragma unroll
            for (int p = 0; p < P; ++p) { gl[p] = gcol[p * (VCOLS / 2)]; dl[p] = dcol[p * (VCOLS / 2)]; }
        } else {
            // one batch of loads (clamped addresses: rows / columns outside the range are discarded below)
            const float* gp = a.bg + ch * a.plane + (size_t)(col_ok ? x : 0);
            const float* dp = a.disk + ch * a.plane + (size_t)(col_ok ? x : 0);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const size_t ro = (size_t)min(y0 + p, a.H - 1) * a.W;
                gl[p] = __ldg(reinterpret_cast<const float2*>(gp + ro));
                dl[p] = __ldg(reinterpret_cast<const float2*>(dp + ro));
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2 b = make_float2(acc[p].x * ws[p], acc[p].y * ws[p]);
            const float2 g = gl[p];
            float2 d = dl[p];
            acc[p] = b;
            if (a.field_gain != 0.0f) {          // the disk layer after _bloom_kernel's in-place add (render.py:3112-3114)
                d.x = fminf(fmaxf(__fadd_rn(d.x, __fmul_rn(b.x, a.field_gain)), 0.0f), 1.0f);
                d.y = fminf(fmaxf(__fadd_rn(d.y, __fmul_rn(b.y, a.field_gain)), 0.0f), 1.0f);
            }
            rgb[ch][p] = make_float2(fminf(fmaxf((g.x + d.x) + b.x, 0.0f), 1.0f), fminf(fmaxf((g.y + d.y) + b.y, 0.0f), 1.0f));
        }
        if (a.want_blur && col_ok) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (y0 + p < a.row1) *reinterpret_cast<float2*>(a.blur + ch * a.plane + (size_t)(y0 + p) * a.W + x) = acc[p];
        }
        if (ch == 2) {
            // ---- (H, W, 3) f32 and u8 rows through the warp's staging row: full-line stores ----
            const int xt = tx * VCOLS;
            const bool full = xt + VCOLS <= a.W;
            float2* o2 = reinterpret_cast<float2*>(ostage) + 3 * lane;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int y = y0 + p;
                if (y >= a.row1) break;          // (warp-uniform)
                __syncwarp();
                o2[0] = make_float2(rgb[0][p].x, rgb[1][p].x);
                o2[1] = make_float2(rgb[2][p].x, rgb[0][p].y);
                o2[2] = make_float2(rgb[1][p].y, rgb[2][p].y);
                __syncwarp();
                const size_t pix = (size_t)y * a.W + xt;
                float* of = a.final_f32 + pix * 3;
                uint8_t* ou = a.final_u8 + pix * 3;
                if (full) {
#pragma unroll
                    for (int i = lane; i < 3 * VCOLS / 4; i += 32) {
                        const float4 q = reinterpret_cast<const float4*>(ostage)[i];
                        reinterpret_cast<float4*>(of)[i] = q;
                        reinterpret_cast<uint32_t*>(ou)[i] = (uint32_t)(unsigned char)(q.x * 255.0f) | ((uint32_t)(unsigned char)(q.y * 255.0f) << 8) |
                                                             ((uint32_t)(unsigned char)(q.z * 255.0f) << 16) | ((uint32_t)(unsigned char)(q.w * 255.0f) << 24);
                    }
                } else {
                    const int n = 3 * (a.W - xt);
                    for (int i = lane; i < n; i += 32) {
                        const float q = ostage[i];
                        of[i] = q;
                        ou[i] = (unsigned char)(q * 255.0f);
                    }
                }
            }
        }
        __syncthreads();
      }
    }
}

// Lens flare on top of the composited rows (render.py:3920-3922): a streaming pass of its own.  Fused into the V
// pass it ran at 8 warps per SM and its dependent float64 chains were latency-bound (4K: 0.55 ms for V + flare
// against 0.16 + 0.05 ms apart); here every SM holds 64 warps.  Reads the flare-less frame from `src` (local)
// and writes the f32 and u8 frames to `dst_*` (local, or rank 0's memory on the peer path).  Four pixels per thread.
__global__ void __launch_bounds__(256) flare_add_kernel(const float* __restrict__ src, float* __restrict__ dst_f32,
                                                        uint8_t* __restrict__ dst_u8, int W, int row0, int row1,
                                                        FlareParams F_arg, const FlareParams* __restrict__ F_dev) {
    const FlareParams F = F_dev ? *F_dev : F_arg;
    const int groups = W / 4;                      // (the TMA path requires W % 4 == 0)
    const size_t n = (size_t)(row1 - row0) * groups;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = row0 + (int)(i / groups), x = (int)(i % groups) * 4;
        const size_t o = ((size_t)y * W + x) * 3;
        const float4* s4 = reinterpret_cast<const float4*>(src + o);
        float4 q[3] = {s4[0], s4[1], s4[2]};
        float* v = reinterpret_cast<float*>(q);
        if (F.enabled) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float fl[3];
                flare_pixel(F, x + j, y, fl);
#pragma unroll
                for (int c = 0; c < 3; ++c) v[3 * j + c] = fminf(fmaxf(v[3 * j + c] + fl[c], 0.0f), 1.0f);
            }
        }
        float4* d4 = reinterpret_cast<float4*>(dst_f32 + o);
        d4[0] = q[0]; d4[1] = q[1]; d4[2] = q[2];
        uint32_t* du = reinterpret_cast<uint32_t*>(dst_u8 + o);
#pragma unroll
        for (int w = 0; w < 3; ++w)
            du[w] = (uint32_t)(unsigned char)(v[4 * w] * 255.0f) | ((uint32_t)(unsigned char)(v[4 * w + 1] * 255.0f) << 8) |
                    ((uint32_t)(unsigned char)(v[4 * w + 2] * 255.0f) << 16) | ((uint32_t)(unsigned char)(v[4 * w + 3] * 255.0f) << 24);
    }
}

// Peer path (row tiles over several GPUs): the R halo rows above and below my tile live in the
// neighbours' H-blurred buffers; pull them into my own buffer (plain loads from peer memory over
// NVLink, coalesced 16-byte accesses) so that the TMA boxes of the V pass read local HBM.
__global__ void __launch_bounds__(256) peer_halo_pull_kernel(float* __restrict__ own, const float* const* __restrict__ row_src,
                                                             int W, int H, int row0, int row1, int R, size_t plane) {
    const int k = blockIdx.x;                     // halo row index: [0, R) above, [R, 2R) below
    const int y = k < R ? row0 - R + k : row1 + (k - R);
    if (y < 0 || y >= H) return;
    const float* src = row_src[y];
    if (src == own) return;
    const int ch = blockIdx.y;
    const float4* s = reinterpret_cast<const float4*>(src + ch * plane + (size_t)y * W);
    float4* d = reinterpret_cast<float4*>(own + ch * plane + (size_t)y * W);
    for (int i = threadIdx.x; i < W / 4; i += 256) d[i] = s[i];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (x, y, channel) view of a planar 3 x (H, W) f32 layer with a box of bx x by x 1 elements
int make_layer_map(bhr_ctx* ctx, PFN_encodeTiled enc, void* base, int bx, int by, unsigned char out[128]) {
    alignas(64) CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)ctx->W, (cuuint64_t)ctx->H, 3};
    const cuuint64_t strides[2] = {(cuuint64_t)ctx->W * sizeof(float), (cuuint64_t)ctx->W * ctx->H * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) BHR_FAIL(ctx, BHR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    memcpy(out, &m, 128);
    return BHR_OK;
}

constexpr int V_P = 10;     // rows per thread of the V pass: tiles of 80 rows divide 360 / 720 / 1080 / 2160 well

template <int P> size_t h_smem(const bhr_ctx* ctx) {
    return (size_t)8 * 4 * ctx->bloom_h_segp * sizeof(float) + (size_t)((3 * ctx->wtab_stride + 1) & ~1) * sizeof(float) + 16 * sizeof(uint64_t);
}
size_t v_smem(const bhr_ctx* ctx, bool staged) {
    const size_t rows = 8 * V_P + 2 * ctx->bloom_R + (staged ? 2 * 8 * V_P : 0);   // H-blurred samples [+ bg tile + disk tile]
    return 2 * rows * VCOLS * sizeof(float) + (size_t)((3 * ctx->wtab_stride + 3) & ~3) * sizeof(float) +
           (size_t)8 * 3 * VCOLS * sizeof(float) + 2 * sizeof(uint64_t);
}

}  // namespace

// Tensor maps and launch geometry of the TMA bloom kernels; leaves ctx->bloom_tma = 0 (the generic
// kernels of post.cu are used instead) when the frame does not meet TMA's layout rules.
int bhr_setup_bloom_tma(bhr_ctx* ctx) {
    ctx->bloom_tma = 0;
    const int W = ctx->W, R = ctx->bloom_R;
    if (W % 4 != 0 || W < 64) return BHR_OK;                   // row pitch must be a multiple of 16 bytes
    if (8 * V_P + 2 * R > 256) return BHR_OK;                  // V tile = one box (<= 256 rows)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return BHR_OK;
    }
    const PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    // H pass: P = 15 pixels per lane when the width is a multiple of 480 (fhd, 4K), 10 for multiples of 320 (hd, sd:
    // 14.4 vs 18.5 us at hd with P = 5; at fhd / 4K P = 10 with three blocks per SM is no faster than 15 with two), else 5
    ctx->bloom_h_P = (W % 480 == 0) ? 15 : (W % 320 == 0) ? 10 : 5;
    if (ctx->bloom_h_P_option == 10 && W % 320 == 0) ctx->bloom_h_P = 10;      // option "bloom_h_p": A/B
    if (ctx->bloom_h_P_option == 5) ctx->bloom_h_P = 5;
    // (boxes start 16-byte aligned in global memory; + 1: the stencil reads one sample past its window)
    const int seg = 32 * ctx->bloom_h_P + R + (R + 3) / 4 * 4 + 1;
    ctx->bloom_h_nb = bhr_div_up(seg, 256);
    ctx->bloom_h_bw = bhr_div_up(bhr_div_up(seg, ctx->bloom_h_nb), 32) * 32;      // 128-byte aligned box destinations
    if (ctx->bloom_h_bw > 256) { ctx->bloom_h_nb += 1; ctx->bloom_h_bw = bhr_div_up(bhr_div_up(seg, ctx->bloom_h_nb), 32) * 32; }
    ctx->bloom_h_segp = ctx->bloom_h_nb * ctx->bloom_h_bw;
    int rc = make_layer_map(ctx, enc, ctx->disk, ctx->bloom_h_bw, 1, ctx->tmap_disk_row);
    if (rc) return rc;
    rc = make_layer_map(ctx, enc, ctx->hblur, VCOLS, 8 * V_P + 2 * R, ctx->tmap_hblur_tile);
    if (rc) return rc;
    rc = make_layer_map(ctx, enc, ctx->bg, VCOLS, 8 * V_P, ctx->tmap_bg_tile);
    if (rc) return rc;
    rc = make_layer_map(ctx, enc, ctx->disk, VCOLS, 8 * V_P, ctx->tmap_disk_tile);
    if (rc) return rc;
    // Variant of the V pass: the kernel is FMA-bound, so the tiles per SM decide -- ceil(n / SMs) with one staged
    // block per SM against 2 ceil(n / 2 SMs) with two unstaged ones (measured: fhd 37 vs 45 us, hd 23 vs 21 us);
    // the unstaged variant also needs its two sample stages to fit twice (not at 4K).
    {
        const int n_tiles = bhr_div_up(W, VCOLS) * bhr_div_up(ctx->H, 8 * V_P);
        const int per_sm_staged = bhr_div_up(n_tiles, ctx->num_sms), per_sm_unstaged = 2 * bhr_div_up(n_tiles, 2 * ctx->num_sms);
        ctx->bloom_v_staged = v_smem(ctx, false) > 112 * 1024 || per_sm_staged < per_sm_unstaged;
    }
    const size_t hs = h_smem<15>(ctx), vs = v_smem(ctx, ctx->bloom_v_staged);      // (h_smem does not depend on P but through segp)
    if (hs > 110 * 1024 || vs > 225 * 1024) return BHR_OK;
    if (ctx->bloom_h_P == 15) BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_h_tma_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
    else if (ctx->bloom_h_P == 10) BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_h_tma_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
    else BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_h_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
    if (ctx->bloom_v_staged) BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_v_fused_kernel<V_P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs));
    else BHR_CUDA(ctx, cudaFuncSetAttribute(bloom_v_fused_kernel<V_P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs));
    ctx->bloom_tma = 1;
    return BHR_OK;
}

int bhr_launch_bloom_h_tma(bhr_ctx* ctx, int row0, int row1) {
    HArgs a;
    a.dst = ctx->hblur; a.wtab = ctx->d_wtab; a.wtab_stride = ctx->wtab_stride; a.wsum_x = ctx->d_wsum_x;
    a.W = ctx->W; a.H = ctx->H; a.row0 = row0; a.row1 = row1; a.R = ctx->bloom_R;
    a.bw = ctx->bloom_h_bw; a.nb = ctx->bloom_h_nb; a.segp = ctx->bloom_h_segp; a.plane = (size_t)ctx->W * ctx->H;
    a.xlead = (ctx->bloom_R + 3) / 4 * 4;
    alignas(64) CUtensorMap map;
    memcpy(&map, ctx->tmap_disk_row, 128);
    const int P = ctx->bloom_h_P;
    const int n_tasks = 3 * ((row1 - row0 + 1) / 2) * bhr_div_up(ctx->W, 32 * P);
    const int per_sm = P <= 10 ? 3 : 2;
    const int grid = bhr_div_up(n_tasks, 8) < per_sm * ctx->num_sms ? bhr_div_up(n_tasks, 8) : per_sm * ctx->num_sms;
    if (P == 15) bloom_h_tma_kernel<15><<<grid, 256, h_smem<15>(ctx), ctx->stream>>>(map, a);
    else if (P == 10) bloom_h_tma_kernel<10><<<grid, 256, h_smem<10>(ctx), ctx->stream>>>(map, a);
    else bloom_h_tma_kernel<5><<<grid, 256, h_smem<5>(ctx), ctx->stream>>>(map, a);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// V pass + composite of rows [row0, row1).  F_host / F_dev: flare parameters (host struct or device pointer, both may
// be absent); row_src != NULL: pull the halo rows from the peers' buffers first.
int bhr_launch_bloom_v_fused(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const void* F_host, const void* F_dev,
                             const float* const* row_src, float* dst_f32, uint8_t* dst_u8) {
    const int W = ctx->W, H = ctx->H, R = ctx->bloom_R;
    if (row_src) {
        peer_halo_pull_kernel<<<dim3(2 * R, 3), 256, 0, ctx->stream>>>(ctx->hblur, row_src, W, H, row0, row1, R, (size_t)W * H);
        ++ctx->launches;
        BHR_CUDA(ctx, cudaGetLastError());
    }
    VArgs a;
    memset(&a, 0, sizeof(a));
    FlareParams F;
    memset(&F, 0, sizeof(F));
    if (F_host) memcpy(&F, F_host, sizeof(FlareParams));
    const bool flare = F.enabled || F_dev;
    // with the flare on, the V pass leaves the flare-less rows in the local final buffers and the flare pass
    // carries them to their destination
    a.bg = ctx->bg; a.disk = ctx->disk; a.blur = ctx->blur;
    a.final_f32 = flare ? ctx->final_f32 : dst_f32; a.final_u8 = flare ? ctx->final_u8 : dst_u8;
    a.wtab = ctx->d_wtab; a.wtab_stride = ctx->wtab_stride; a.wsum_y = ctx->d_wsum_y;
    a.W = W; a.H = H; a.row0 = row0; a.row1 = row1; a.R = R; a.plane = (size_t)W * H;
    a.field_gain = (flags & BHR_FIELD_COMPOSITE) ? 0.4f : 0.0f;
    a.want_blur = ctx->keep_blur;
    alignas(64) CUtensorMap map, map_bg, map_disk;
    memcpy(&map, ctx->tmap_hblur_tile, 128);
    memcpy(&map_bg, ctx->tmap_bg_tile, 128);
    memcpy(&map_disk, ctx->tmap_disk_tile, 128);
    const int n_tiles = bhr_div_up(W, VCOLS) * bhr_div_up(row1 - row0, 8 * V_P);
    // The kernel is FMA-bound, so what counts is the number of tiles per SM: with b blocks per SM, a grid of
    // b * SMs blocks with ceil(n_tiles / grid) tiles each at most, trimmed to leave as few SMs idle as possible
    const int b = ctx->bloom_v_staged ? 1 : 2;
    const int rounds = bhr_div_up(n_tiles, b * ctx->num_sms);
    const int grid = bhr_div_up(n_tiles, rounds);
    if (ctx->bloom_v_staged) bloom_v_fused_kernel<V_P, true><<<grid, 256, v_smem(ctx, true), ctx->stream>>>(map, map_bg, map_disk, a);
    else bloom_v_fused_kernel<V_P, false><<<grid, 256, v_smem(ctx, false), ctx->stream>>>(map, map_bg, map_disk, a);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    if (flare) {
        flare_add_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->final_f32, dst_f32, dst_u8, W, row0, row1, F,
                                                                    (const FlareParams*)F_dev);
        ++ctx->launches;
        BHR_CUDA(ctx, cudaGetLastError());
    }
    return BHR_OK;
}
