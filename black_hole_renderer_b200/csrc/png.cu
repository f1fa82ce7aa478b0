// png.cu -- the deflate stream of a frame's PNG file, produced on the device.
//
// The reference hands every video frame to PIL on the host (render.py:4462-4467); zlib on the host
// is what limits a video run with files (tens of ms per 1080p frame and core).  Here the finished
// 8-bit frame is turned into the complete zlib stream of its IDAT chunk by three small kernels, so
// that a frame file costs the host one CRC-32 and one write.  Format and a CPU twin of every step:
// black_hole_renderer_b200/png_codec.py (tests hold the two to each other byte for byte, and the CPU
// twin to zlib / PIL).
//
//   encode   one warp per 256-byte segment of the Sub-filtered scanline stream (filter byte + RGB
//            differences to the pixel on the left, formed on the fly from the u8 frame), a lane per
//            8 bytes: runs of a repeated byte -> (length, distance 1) matches, the rest literals,
//            coded with a STATIC Huffman table (uploaded once); the segment's bits go to a staging
//            area, its bit count and Adler-32 partial sums to arrays;
//   scan     one block: exclusive prefix sum of the bit counts (+ zlib / block header bits), the
//            Adler-32 of the whole stream from the partial sums, header / end-of-block / checksum bytes;
//   merge    every segment ORs its bits into place (bit granularity, atomicOr on 32-bit words).
#include "common.cuh"

namespace {

constexpr int kSeg = 256;            // == png_codec.SEGMENT
constexpr int kMinRun = 4;           // == png_codec.MIN_RUN
constexpr int kSegWords = 112;       // staging words per segment: 256 bytes x <= 13 bits + slack (checked at upload)

struct PngTables {
    unsigned int lit_bits[256], lit_nbits[256];
    unsigned int match_bits[259], match_nbits[259];
    unsigned char header[64];
    unsigned int header_nbits, eob_bits, eob_nbits;
};

struct PngInfo { unsigned int total_bytes, adler; };

constexpr int kEncWarps = 8;         // segments (= warps) per block of the encode and merge kernels

// per block of kEncWarps segments: bits, byte sum, and the block's Adler "B" contribution with A counted from the block's start
struct PngBlockSums { unsigned int bits, sum; unsigned long long b; };

// One WARP per 256-byte segment, one lane per 8 consecutive bytes; no data-dependent loops.
//   * the lane forms its 8 filtered bytes from the u8 frame (row = g / (3W + 1), column 0 is the filter type 1 = Sub,
//     the others raw - left pixel);
//   * a byte continues a run when it equals its predecessor in the segment.  back[q] / cont[q] = bytes of the same run
//     before / after position q; across lanes they come from two segmented scans (lead / trail = bytes at the lane's
//     start / end that continue the neighbour's run; a lane made of one run passes the count through);
//   * a run of m bytes is coded as literal + (length m - 1, distance 1) match if m - 1 >= kMinRun (a segment is shorter
//     than the longest match, 258), else as m literals: position q emits its literal unless a match covers it, and the
//     run's first position also emits the match.  Bits go into the warp's buffer in shared memory at (exclusive scan
//     of the lanes' bit counts); the warp copies the buffer out coalesced.
__global__ void __launch_bounds__(32 * kEncWarps) png_encode_kernel(const uint8_t* __restrict__ img, int row_bytes, unsigned int n_bytes, int n_seg,
                                                                    const PngTables* __restrict__ T, unsigned int* __restrict__ staging,
                                                                    unsigned int* __restrict__ seg_bits, PngBlockSums* __restrict__ block_sums) {
    __shared__ unsigned int s_lit[256];                       // code | nbits << 16
    __shared__ unsigned int s_buf[kEncWarps][kSegWords];
    __shared__ unsigned int s_tot[kEncWarps][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += 32 * kEncWarps) s_lit[i] = T->lit_bits[i] | (T->lit_nbits[i] << 16);
    for (int i = lane; i < kSegWords; i += 32) s_buf[warp][i] = 0u;
    __syncthreads();
    const int s = blockIdx.x * kEncWarps + warp;
    const unsigned int g0 = (unsigned int)s * kSeg;
    const int n = s < n_seg ? (int)min(n_bytes - g0, (unsigned int)kSeg) : 0;
    // ---- the lane's 8 filtered bytes ----
    unsigned int x[8];
    {
        const unsigned int row_len = (unsigned int)row_bytes + 1u;
        const unsigned int g = g0 + 8u * lane;
        unsigned int row = g / row_len, col = g - row * row_len;
        const uint8_t* p = img + (size_t)row * row_bytes + col;      // raw byte of column col is p[-1]
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned int v = 1u;                                        // column 0: filter type
            if (8 * lane + q < n && col != 0) v = ((unsigned int)p[-1] - (col >= 4 ? (unsigned int)p[-4] : 0u)) & 255u;
            x[q] = v;
            if (++col == row_len) col = 0;                              // (p then is the next row's first byte already)
            else ++p;
        }
    }
    const int nv = max(0, min(8, n - 8 * lane));                        // valid bytes of this lane
    // ---- run structure ----
    const unsigned int prev_last = __shfl_up_sync(0xffffffffu, x[7], 1);
    unsigned int eqm = 0;                                               // bit q: byte q continues its predecessor's run
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (q < nv && (q > 0 ? x[q] == x[q - 1] : (lane > 0 && x[0] == prev_last))) eqm |= 1u << q;
    const int lead = __ffs(~eqm) - 1;                                   // leading ones of eqm (8 = the whole lane continues)
    // trail = bytes at the lane's end that belong to the run of its last byte, not counting a continuation from the left
    // (eq bits of positions 1..7 set from the top): a full lane has trail 8 only through lead
    const int trail = min(__clz(~(eqm << 24)), 7) + 1;                  // 1..8 bytes ending at position 7 in one run (within the lane)
    // F(i) = bytes from lane i's first byte on that continue lane i - 1's last run; Bk(i) = bytes up to lane i's last byte
    // that belong to its last run (reaching into earlier lanes)
    int F = lead, Bk = lead == 8 ? 8 : trail;
    bool openf = lead == 8, openb = lead == 8;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int f2 = __shfl_down_sync(0xffffffffu, F, off), o2 = __shfl_down_sync(0xffffffffu, (int)openf, off);
        const int b2 = __shfl_up_sync(0xffffffffu, Bk, off), p2 = __shfl_up_sync(0xffffffffu, (int)openb, off);
        if (lane + off < 32) { if (openf) F += f2; openf = openf && o2; }
        if (lane >= off) { if (openb) Bk += b2; openb = openb && p2; }
    }
    int ext = __shfl_down_sync(0xffffffffu, F, 1), before = __shfl_up_sync(0xffffffffu, Bk, 1);
    if (lane == 31) ext = 0;
    if (lane == 0) before = 0;
    // ---- tokens ----
    int cont[8], back[8];
    cont[7] = ext;
#pragma unroll
    for (int q = 6; q >= 0; --q) cont[q] = (eqm >> (q + 1)) & 1u ? 1 + cont[q + 1] : 0;
    back[0] = eqm & 1u ? before : 0;
#pragma unroll
    for (int q = 1; q < 8; ++q) back[q] = (eqm >> q) & 1u ? 1 + back[q - 1] : 0;
    unsigned long long tok[8];
    unsigned int nb[8], lane_bits = 0, sum = 0, wsum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const bool valid = q < nv;
        const int r = back[q] + cont[q];                                // the run's other bytes
        const bool matched = r >= kMinRun, first = back[q] == 0;
        const unsigned int lit = s_lit[x[q]], lc = lit & 0xffffu, ln = lit >> 16;
        unsigned long long t = 0;
        unsigned int bits = 0;
        if (valid && (first || !matched)) { t = lc; bits = ln; }
        if (valid && first && matched) {
            t |= (unsigned long long)T->match_bits[r] << ln;
            bits += T->match_nbits[r];
        }
        tok[q] = t;
        nb[q] = bits;
        lane_bits += bits;
        if (valid) { sum += x[q]; wsum += (unsigned int)(n - (8 * lane + q)) * x[q]; }
    }
    // ---- exclusive scan of the lanes' bit counts, emission ----
    unsigned int incl = lane_bits;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned int pos = incl - lane_bits;
    unsigned int* buf = s_buf[warp];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (nb[q]) {                                                    // (<= 13 + 20 bits: two words)
            const unsigned long long t = tok[q] << (pos & 31u);
            atomicOr(buf + (pos >> 5), (unsigned int)t);
            if (t >> 32) atomicOr(buf + (pos >> 5) + 1, (unsigned int)(t >> 32));
            pos += nb[q];
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        wsum += __shfl_xor_sync(0xffffffffu, wsum, off);
    }
    __syncwarp();
    const int words = (int)((total + 31) >> 5);
    if (s < n_seg) {
        for (int j = lane; j < words; j += 32) staging[(size_t)s * kSegWords + j] = buf[j];
        if (lane == 0) seg_bits[s] = total;
    }
    if (lane == 0) { s_tot[warp][0] = total; s_tot[warp][1] = sum; s_tot[warp][2] = wsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        // Adler-32 within the block: with A = 1 + (sum of the earlier bytes), a segment of n bytes adds n A + sum_k (n - k) x_k
        // to B; the "1 + bytes before the block" part of A is added by the scan kernel (n_block x that prefix)
        PngBlockSums bs = {0u, 0u, 0ull};
        for (int w = 0; w < kEncWarps; ++w) {
            const int sw = blockIdx.x * kEncWarps + w;
            if (sw >= n_seg) break;
            const unsigned long long nw = min(n_bytes - (unsigned int)sw * kSeg, (unsigned int)kSeg);
            bs.b += nw * bs.sum + s_tot[w][2];
            bs.bits += s_tot[w][0];
            bs.sum += s_tot[w][1];
        }
        block_sums[blockIdx.x] = bs;
    }
}

// one block of 1024 threads: exclusive scan of the encode blocks' bit counts and byte sums in tiles of 1024 (coalesced);
// header, end-of-block code and Adler-32 of the whole stream
__global__ void __launch_bounds__(1024) png_scan_kernel(const PngBlockSums* __restrict__ block_sums, int n_blocks, size_t n_bytes,
                                                        const PngTables* __restrict__ T, unsigned long long* __restrict__ block_off,
                                                        unsigned int* __restrict__ out, PngInfo* __restrict__ info) {
    __shared__ unsigned long long sh_bits[32], sh_sum[32], sh_b[32];
    __shared__ unsigned long long sh_tile[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long base_bits = 16ull + T->header_nbits;       // zlib header (2 bytes) + block header
    unsigned long long carry_bits = base_bits, carry_sum = 0, B = 0;    // running totals before the current tile
    for (int t0 = 0; t0 < n_blocks; t0 += 1024) {
        const int i = t0 + tid;
        PngBlockSums bs = {0u, 0u, 0ull};
        if (i < n_blocks) bs = block_sums[i];
        unsigned long long ib = bs.bits, is = bs.sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long vb = __shfl_up_sync(0xffffffffu, ib, off), vs = __shfl_up_sync(0xffffffffu, is, off);
            if (lane >= off) { ib += vb; is += vs; }
        }
        if (lane == 31) { sh_bits[warp] = ib; sh_sum[warp] = is; }
        __syncthreads();
        if (warp == 0) {                                                // exclusive scan of the 32 warp totals
            unsigned long long wb = sh_bits[lane], ws = sh_sum[lane];
            const unsigned long long wb0 = wb, ws0 = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long vb = __shfl_up_sync(0xffffffffu, wb, off), vs = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) { wb += vb; ws += vs; }
            }
            sh_bits[lane] = wb - wb0; sh_sum[lane] = ws - ws0;
            if (lane == 31) { sh_tile[0] = wb; sh_tile[1] = ws; }
        }
        __syncthreads();
        if (i < n_blocks) {
            block_off[i] = carry_bits + sh_bits[warp] + ib - bs.bits;
            const size_t g0 = (size_t)i * (kEncWarps * kSeg);
            const unsigned long long nblk = (n_bytes - g0) < (size_t)(kEncWarps * kSeg) ? (n_bytes - g0) : (size_t)(kEncWarps * kSeg);
            // (no reduction needed before the end: n A < 2^45, fewer than 2^17 blocks)
            B += nblk * (1ull + carry_sum + sh_sum[warp] + is - bs.sum) + bs.b;
        }
        carry_bits += sh_tile[0]; carry_sum += sh_tile[1];
        __syncthreads();
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) B += __shfl_xor_sync(0xffffffffu, B, off);
    if (lane == 0) sh_b[warp] = B;
    __syncthreads();
    if (tid == 0) {
        unsigned long long Bt = 0;
        for (int i = 0; i < 32; ++i) Bt += sh_b[i] % 65521ull;
        Bt %= 65521ull;
        const unsigned long long At = (1ull + carry_sum) % 65521ull;
        const unsigned int adler = (unsigned int)((Bt << 16) | At);
        const unsigned long long end_bits = carry_bits;                          // where the end-of-block code goes
        uint8_t* o8 = reinterpret_cast<uint8_t*>(out);
        // zlib header + block header (the buffer was zeroed; the merge kernel ORs the segments in afterwards)
        o8[0] = 0x78; o8[1] = 0x01;
        for (unsigned int k = 0; k < (T->header_nbits + 7) / 8; ++k) atomicOr(out + ((2 + k) >> 2), (unsigned int)T->header[k] << (8 * ((2 + k) & 3)));
        // end of block
        const unsigned long long eb = (unsigned long long)T->eob_bits << (end_bits & 31);
        atomicOr(out + (end_bits >> 5), (unsigned int)eb);
        if (eb >> 32) atomicOr(out + (end_bits >> 5) + 1, (unsigned int)(eb >> 32));
        const unsigned long long stream_bytes = (end_bits + T->eob_nbits + 7) >> 3;
        for (int k = 0; k < 4; ++k) {
            const unsigned long long p = stream_bytes + k;
            atomicOr(out + (p >> 2), ((adler >> (8 * (3 - k))) & 255u) << (8 * (p & 3)));       // big-endian checksum
        }
        info->total_bytes = (unsigned int)(stream_bytes + 4);
        info->adler = adler;
    }
}

// warp per segment: its words, shifted to the segment's bit offset (block offset + the block's earlier segments), are ORed
// into the stream
__global__ void __launch_bounds__(32 * kEncWarps) png_merge_kernel(const unsigned int* __restrict__ staging, const unsigned int* __restrict__ seg_bits,
                                                                   const unsigned long long* __restrict__ block_off, int n_seg,
                                                                   unsigned int* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, s = blockIdx.x * kEncWarps + warp;
    if (s >= n_seg) return;
    unsigned int mine = lane <= warp ? seg_bits[blockIdx.x * kEncWarps + lane] : 0u;       // (s < n_seg: those segments exist)
    const unsigned int nb = __shfl_sync(0xffffffffu, mine, warp);
    if (lane >= warp) mine = 0u;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);   // lanes 0..7 hold the values
    const unsigned long long off = block_off[blockIdx.x] + __shfl_sync(0xffffffffu, mine, 0);
    const int sh = (int)(off & 31);
    unsigned int* dst = out + (off >> 5);
    const int words = (int)((nb + 31) >> 5);
    const unsigned int* src = staging + (size_t)s * kSegWords;
    for (int j = lane; j <= words; j += 32) {                       // output word j = low part of word j | high part of word j - 1
        const unsigned int cur = j < words ? src[j] : 0u, prev = j > 0 ? src[j - 1] : 0u;
        const unsigned int v = sh ? (cur << sh) | (prev >> (32 - sh)) : cur;
        if (v) atomicOr(dst + j, v);
    }
}

}  // namespace

// tables of the static code (png_codec.StaticCode.device_tables()), once per context
extern "C" int bhr_png_setup(bhr_ctx* ctx, const uint32_t* lit_bits, const uint32_t* lit_nbits, const uint32_t* match_bits,
                             const uint32_t* match_nbits, const uint8_t* header, uint32_t header_nbits, uint32_t eob_bits,
                             uint32_t eob_nbits) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !lit_bits || !lit_nbits || !match_bits || !match_nbits || !header) return BHR_ERR_INVALID;
    if ((header_nbits + 7) / 8 > 64) BHR_FAIL(ctx, BHR_ERR_INVALID, "block header too long");
    PngTables T;
    memset(&T, 0, sizeof(T));
    unsigned int max_lit = 0;
    for (int i = 0; i < 256; ++i) { T.lit_bits[i] = lit_bits[i]; T.lit_nbits[i] = lit_nbits[i]; if (lit_nbits[i] > max_lit) max_lit = lit_nbits[i]; }
    for (int i = 0; i < 259; ++i) { T.match_bits[i] = match_bits[i]; T.match_nbits[i] = match_nbits[i]; }
    if (max_lit * kSeg > 32u * (kSegWords - 2)) BHR_FAIL(ctx, BHR_ERR_INVALID, "literal codes too long for the staging area");
    unsigned int max_match = 0;
    for (int i = kMinRun; i < 259; ++i) if (match_nbits[i] > max_match) max_match = match_nbits[i];
    if (max_lit + max_match > 33u) BHR_FAIL(ctx, BHR_ERR_INVALID, "literal + match code longer than 33 bits");   // (encode kernel: two words per token)
    memcpy(T.header, header, (header_nbits + 7) / 8);
    T.header_nbits = header_nbits; T.eob_bits = eob_bits; T.eob_nbits = eob_nbits;
    const size_t n_bytes = (size_t)ctx->H * (3 * (size_t)ctx->W + 1);
    if (n_bytes >= (1ull << 31)) BHR_FAIL(ctx, BHR_ERR_INVALID, "frame too large for the PNG encoder");
    const int n_seg = (int)((n_bytes + kSeg - 1) / kSeg);
    ctx->png_n_seg = n_seg;
    ctx->png_capacity = (2 + (header_nbits + (size_t)max_lit * n_bytes + 15 + 7) / 8 + 4 + 64 + 3) & ~(size_t)3;
    if (!ctx->d_png_tables) {
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_tables, sizeof(PngTables)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_staging, (size_t)kSegWords * n_seg * sizeof(unsigned int)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_seg, (size_t)n_seg * sizeof(unsigned int)));
        // per encode block: bit offset (u64), then its PngBlockSums (16 bytes)
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_off, (size_t)bhr_div_up(n_seg, kEncWarps) * (sizeof(unsigned long long) + sizeof(PngBlockSums))));
    }
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->d_png_tables, &T, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_png_capacity(bhr_ctx* ctx, size_t* stream_bytes) {
    if (!ctx || !stream_bytes) return BHR_ERR_INVALID;
    if (!ctx->d_png_tables) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_png_setup has not run");
    *stream_bytes = ctx->png_capacity;
    return BHR_OK;
}

// encode the current 8-bit frame (BHR_BUF_FINAL_U8) into slot's device stream; enqueued on `stream`
int bhr_launch_png_encode(bhr_ctx* ctx, int slot, cudaStream_t stream) {
    if (!ctx->d_png_tables) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_png_setup has not run");
    if (!ctx->d_png_stream[slot]) {
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_stream[slot], ctx->png_capacity + sizeof(PngInfo)));
    }
    PngInfo* info = reinterpret_cast<PngInfo*>(ctx->d_png_stream[slot]);
    unsigned int* out = reinterpret_cast<unsigned int*>(ctx->d_png_stream[slot] + sizeof(PngInfo));
    const size_t n_bytes = (size_t)ctx->H * (3 * (size_t)ctx->W + 1);
    const int n_seg = ctx->png_n_seg, n_blocks = bhr_div_up(n_seg, kEncWarps);
    PngBlockSums* block_sums = reinterpret_cast<PngBlockSums*>(ctx->d_png_off + n_blocks);
    BHR_CUDA(ctx, cudaMemsetAsync(out, 0, ctx->png_capacity, stream));
    png_encode_kernel<<<n_blocks, 32 * kEncWarps, 0, stream>>>(ctx->final_u8, 3 * ctx->W, (unsigned int)n_bytes, n_seg, (const PngTables*)ctx->d_png_tables,
                                                                    ctx->d_png_staging, ctx->d_png_seg, block_sums);
    png_scan_kernel<<<1, 1024, 0, stream>>>(block_sums, n_blocks, n_bytes, (const PngTables*)ctx->d_png_tables, ctx->d_png_off, out, info);
    png_merge_kernel<<<n_blocks, 32 * kEncWarps, 0, stream>>>(ctx->d_png_staging, ctx->d_png_seg, ctx->d_png_off, n_seg, out);
    ctx->launches += 3;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}
