// png.cu -- the deflate stream of a frame's PNG file, produced on the device.
//
// The reference hands every video frame to PIL on the host (render.py:4462-4467); zlib on the host
// is what limits a video run with files (tens of ms per 1080p frame and core).  Here the finished
// 8-bit frame is turned into the complete zlib stream of its IDAT chunk by three small kernels, so
// that a frame file costs the host one CRC-32 and one write.  Format and a CPU twin of every step:
// black_hole_renderer_b200/png_codec.py (tests hold the two to each other byte for byte, and the CPU
// twin to zlib / PIL).
//
//   encode   one thread per 256-byte segment of the Sub-filtered scanline stream (filter byte + RGB
//            differences to the pixel on the left, formed on the fly from the u8 frame): runs of a
//            repeated byte -> (length, distance 1) matches, the rest literals, coded with a STATIC
//            Huffman table (uploaded once); the segment's bits go to a staging area laid out
//            word-major (coalesced), its bit count and Adler-32 partial sums to arrays;
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

// walks the Sub-filtered scanline stream: row = g / (3W + 1); column 0 is the filter type (1 = Sub)
struct FilteredWalk {
    const uint8_t* __restrict__ img;
    int row_bytes, col;
    const uint8_t* p;                       // raw byte under the cursor (the row's first byte while col == 0)
    __device__ FilteredWalk(const uint8_t* image, int rb, size_t g) : img(image), row_bytes(rb) {
        const size_t row = g / (size_t)(rb + 1);
        col = (int)(g - row * (size_t)(rb + 1));
        p = image + row * (size_t)rb + (col > 0 ? col - 1 : 0);
    }
    __device__ __forceinline__ unsigned int next() {
        unsigned int v;
        if (col == 0) v = 1u;
        else { v = ((unsigned int)p[0] - (col >= 4 ? (unsigned int)p[-3] : 0u)) & 255u; ++p; }
        if (++col > row_bytes) col = 0;       // p then already points at the next row's first byte
        return v;
    }
};

__global__ void __launch_bounds__(128) png_encode_kernel(const uint8_t* __restrict__ img, int row_bytes, size_t n_bytes, int n_seg,
                                                         const PngTables* __restrict__ T, unsigned int* __restrict__ staging,
                                                         unsigned int* __restrict__ seg_bits, unsigned int* __restrict__ seg_sum,
                                                         unsigned int* __restrict__ seg_wsum) {
    __shared__ unsigned int s_lit_bits[256], s_lit_nbits[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_lit_bits[i] = T->lit_bits[i]; s_lit_nbits[i] = T->lit_nbits[i]; }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const size_t g0 = (size_t)s * kSeg;
    const int n = (int)((n_bytes - g0) < (size_t)kSeg ? (n_bytes - g0) : (size_t)kSeg);
    unsigned long long acc = 0;          // bit accumulator (LSB first)
    int nacc = 0, word = 0;
    unsigned int total = 0;
    auto put = [&](unsigned int v, unsigned int nb) {
        acc |= (unsigned long long)v << nacc;
        nacc += (int)nb;
        total += nb;
        if (nacc >= 32) {
            staging[(size_t)word * n_seg + s] = (unsigned int)acc;      // word-major: coalesced across the warp
            ++word;
            acc >>= 32;
            nacc -= 32;
        }
    };
    // a maximal run of m equal bytes: its first byte is a literal, the other m - 1 a match if there are >= kMinRun
    // of them (segments are shorter than the longest match, 258), else literals
    auto flush_run = [&](unsigned int b, int m) {
        put(s_lit_bits[b], s_lit_nbits[b]);
        const int r = m - 1;
        if (r >= kMinRun) put(T->match_bits[r], T->match_nbits[r]);
        else for (int k = 0; k < r; ++k) put(s_lit_bits[b], s_lit_nbits[b]);
    };
    FilteredWalk walk(img, row_bytes, g0);
    unsigned int cur = walk.next(), sum = cur, wsum = (unsigned int)n * cur;
    int m = 1;
    for (int k = 1; k < n; ++k) {
        const unsigned int b = walk.next();
        sum += b;
        wsum += (unsigned int)(n - k) * b;
        if (b == cur) { ++m; continue; }
        flush_run(cur, m);
        cur = b; m = 1;
    }
    flush_run(cur, m);
    if (nacc > 0) staging[(size_t)word * n_seg + s] = (unsigned int)acc;
    seg_bits[s] = total;
    seg_sum[s] = sum;
    seg_wsum[s] = wsum;
}

// one block of 1024 threads: exclusive scan of the segments' bit counts and byte sums; header, trailer, Adler-32
__global__ void __launch_bounds__(1024) png_scan_kernel(const unsigned int* __restrict__ seg_bits, const unsigned int* __restrict__ seg_sum,
                                                        const unsigned int* __restrict__ seg_wsum, int n_seg, size_t n_bytes,
                                                        const PngTables* __restrict__ T, unsigned long long* __restrict__ bit_off,
                                                        unsigned int* __restrict__ out, PngInfo* __restrict__ info) {
    __shared__ unsigned long long sh_bits[1024], sh_sum[1024];
    __shared__ unsigned long long sh_b[1024];
    const int tid = threadIdx.x;
    const int per = (n_seg + 1023) / 1024;
    const int a = tid * per, b = min(a + per, n_seg);
    unsigned long long bits = 0, sum = 0;
    for (int i = a; i < b; ++i) { bits += seg_bits[i]; sum += seg_sum[i]; }
    sh_bits[tid] = bits; sh_sum[tid] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned long long vb = 0, vs = 0;
        if (tid >= off) { vb = sh_bits[tid - off]; vs = sh_sum[tid - off]; }
        __syncthreads();
        sh_bits[tid] += vb; sh_sum[tid] += vs;
        __syncthreads();
    }
    const unsigned long long base_bits = 16ull + T->header_nbits;       // zlib header (2 bytes) + block header
    unsigned long long pb = base_bits + sh_bits[tid] - bits, ps = sh_sum[tid] - sum;
    unsigned long long B = 0;
    for (int i = a; i < b; ++i) {
        bit_off[i] = pb;
        pb += seg_bits[i];
        // Adler-32 over the stream: with A = 1 + (sum of all earlier bytes), a segment of n bytes adds n A + sum (n - k) x_k to B
        const size_t g0 = (size_t)i * kSeg;
        const unsigned long long n = (n_bytes - g0) < (size_t)kSeg ? (n_bytes - g0) : (size_t)kSeg;
        const unsigned long long A = (1ull + ps) % 65521ull;
        B = (B + n * A + seg_wsum[i]) % 65521ull;
        ps += seg_sum[i];
    }
    sh_b[tid] = B;
    __syncthreads();
    if (tid == 0) {
        unsigned long long Bt = 0;
        for (int i = 0; i < 1024; ++i) Bt += sh_b[i];
        Bt %= 65521ull;
        const unsigned long long At = (1ull + sh_sum[1023]) % 65521ull;
        const unsigned int adler = (unsigned int)((Bt << 16) | At);
        const unsigned long long end_bits = base_bits + sh_bits[1023];           // where the end-of-block code goes
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

__global__ void __launch_bounds__(128) png_merge_kernel(const unsigned int* __restrict__ staging, const unsigned int* __restrict__ seg_bits,
                                                        const unsigned long long* __restrict__ bit_off, int n_seg,
                                                        unsigned int* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const unsigned int nb = seg_bits[s];
    const unsigned long long off = bit_off[s];
    const int sh = (int)(off & 31);
    unsigned int* dst = out + (off >> 5);
    const int words = (int)((nb + 31) >> 5);
    unsigned int carry = 0;
    for (int j = 0; j < words; ++j) {
        unsigned int w = staging[(size_t)j * n_seg + s];
        if (j == words - 1 && (nb & 31)) w &= (1u << (nb & 31)) - 1u;       // (bits past the segment's end are garbage of the accumulator)
        const unsigned int lo = (w << sh) | carry;
        carry = sh ? (w >> (32 - sh)) : 0u;
        if (lo) atomicOr(dst + j, lo);
    }
    if (carry) atomicOr(dst + words, carry);
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
    if (max_lit * kSeg > 32u * (kSegWords - 1)) BHR_FAIL(ctx, BHR_ERR_INVALID, "literal codes too long for the staging area");
    memcpy(T.header, header, (header_nbits + 7) / 8);
    T.header_nbits = header_nbits; T.eob_bits = eob_bits; T.eob_nbits = eob_nbits;
    const size_t n_bytes = (size_t)ctx->H * (3 * (size_t)ctx->W + 1);
    const int n_seg = (int)((n_bytes + kSeg - 1) / kSeg);
    ctx->png_n_seg = n_seg;
    ctx->png_capacity = (2 + (header_nbits + (size_t)max_lit * n_bytes + 15 + 7) / 8 + 4 + 64 + 3) & ~(size_t)3;
    if (!ctx->d_png_tables) {
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_tables, sizeof(PngTables)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_staging, (size_t)kSegWords * n_seg * sizeof(unsigned int)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_seg, (size_t)3 * n_seg * sizeof(unsigned int)));
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_off, (size_t)n_seg * sizeof(unsigned long long)));
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

// encode the current 8-bit frame (BHR_BUF_FINAL_U8) into slot's device stream; enqueued on the context's stream
int bhr_launch_png_encode(bhr_ctx* ctx, int slot) {
    if (!ctx->d_png_tables) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_png_setup has not run");
    if (!ctx->d_png_stream[slot]) {
        BHR_CUDA(ctx, cudaMalloc(&ctx->d_png_stream[slot], ctx->png_capacity + sizeof(PngInfo)));
    }
    PngInfo* info = reinterpret_cast<PngInfo*>(ctx->d_png_stream[slot]);
    unsigned int* out = reinterpret_cast<unsigned int*>(ctx->d_png_stream[slot] + sizeof(PngInfo));
    const size_t n_bytes = (size_t)ctx->H * (3 * (size_t)ctx->W + 1);
    const int n_seg = ctx->png_n_seg;
    unsigned int* seg_bits = ctx->d_png_seg;
    unsigned int* seg_sum = seg_bits + n_seg;
    unsigned int* seg_wsum = seg_sum + n_seg;
    // only the part of the buffer the previous frame of this slot can have touched needs zeroing; its size is not known
    // on the host, so the first use clears everything and later ones the previous high-water mark
    BHR_CUDA(ctx, cudaMemsetAsync(out, 0, ctx->png_capacity, ctx->stream));
    png_encode_kernel<<<bhr_div_up(n_seg, 128), 128, 0, ctx->stream>>>(ctx->final_u8, 3 * ctx->W, n_bytes, n_seg, (const PngTables*)ctx->d_png_tables,
                                                                       ctx->d_png_staging, seg_bits, seg_sum, seg_wsum);
    png_scan_kernel<<<1, 1024, 0, ctx->stream>>>(seg_bits, seg_sum, seg_wsum, n_seg, n_bytes, (const PngTables*)ctx->d_png_tables,
                                                 ctx->d_png_off, out, info);
    png_merge_kernel<<<bhr_div_up(n_seg, 128), 128, 0, ctx->stream>>>(ctx->d_png_staging, seg_bits, ctx->d_png_off, n_seg, out);
    ctx->launches += 3;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}
