// common.cuh -- context layout and small helpers shared by the libbhr.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bhr.h"

#define BHR_NUM_MIPS 5          // base + 4 levels, render.py:2239
#define BHR_FLARE_BLOCKS 592     // blocks of the flare-sum reduction (4 per SM)
#define BHR_FRAME_SLOTS 32       // frames in flight of a pipelined video loop (completion events, entity staging ring)
#define BHR_N_COMP 13           // component planes, render.py:2328-2332

// Per-frame parameters of the ray-march kernel (kernel argument, lives in constant bank 0).
struct RayParams {
    int W, H;                   // full frame
    int row0, row1;             // rows traced by this launch
    float cp[3], cr[3], cu[3], cf[3];
    float tl[3];                // top-left corner of the image plane
    float pw, ph;
    float r_esc, r_esc2;
    float h_base, r_in, r_out, t_offset;
    float inv_span, inv_span_clamped;   // 1 / (r_out - r_in), 1 / max(r_out - r_in, 1e-3)
    float grav_num;             // sqrt(max(1 - 1 / max(|cam|, 1.001), 1e-6)), render.py:2480
    float tilt_rad, tan_t, sin_t, cos_t;
    int max_iter;
    float max_affine;
    int aa_mode;                // sample through the mip pyramid (anti_alias && !skip_diff)
    float aa_strength;
    float tint[3];              // _color_temp_to_tint(6000 K), render.py:2515
    const float4* sky;          // (sky_h, sky_w) RGBx
    int sky_w, sky_h;
    const float4* mips;         // compact pyramid of RGBA texels
    int dtex_w, dtex_h;
    unsigned int level_off[BHR_NUM_MIPS];
    float* bg;                  // planar 3 x (H, W)
    float* disk;                // planar 3 x (H, W)
    size_t plane;               // W * H
    uint8_t* cls;               // (H, W) or nullptr
    int* steps;                 // (H, W) or nullptr
    unsigned long long* total_steps; // or nullptr
    unsigned long long* queue;  // (serial << 32 | pixel) entries to re-trace in strict mode, or nullptr
    unsigned int* queue_count;  // tail: entries appended so far
    unsigned int queue_serial;  // tags the entries of this launch
    int retrace_min_cross;
    float retrace_band;         // rays with band_lo < b / b_crit - 1 < retrace_band go to the strict integrator
    float band_lo;              // (negative) lower edge, see bhr_launch_raymarch
    float inv_rcam3;            // 1 / |cam|^3
    int* band;                  // persistent kernel: pixels of the ill-conditioned band (built beforehand)
    unsigned int* band_count;
    unsigned int* band_head;    // batches of 32 band rays claimed so far
    unsigned int* tile_counter; // fast tiles claimed so far
    int band_prequeued;
    int band_x0, band_x1, band_y0, band_y1;   // pixels band_list_kernel scans (the ring's bounding box)
    int strict_warps;           // warps per strict block that trace band batches (the rest wait, then join the fast pool)
    unsigned long long* timeline;   // optional (option "timeline"): per block {start, strict role done, end} in globaltimer ns
};

struct PeerSync;
struct bhr_ctx;

struct bhr_ctx {
    bhr_config cfg;
    int W, H;
    cudaStream_t stream, own_stream;
    char err[512];

    float4* sky; int sky_w, sky_h;
    float4* mips; int n_r, n_phi; unsigned int level_off[BHR_NUM_MIPS + 1];
    float* tex_staging;         // (n_r, n_phi, 4) upload staging == disk_texture_field

    float *bg, *disk, *hblur, *blur;   // planar 3 x (H, W) each
    float* final_f32;                  // (H, W, 3)
    uint8_t* final_u8;                 // (H, W, 3)
    uint8_t* cls; int* steps;
    unsigned long long* d_total_steps;
    unsigned long long* retrace_queue; unsigned int* d_queue_count; int retrace_min_cross; unsigned int queue_serial; float retrace_band; int persistent, num_sms, pblock_big, band_lo_auto;
    double* d_flare_sums;              // {sum B, sum x*B, sum y*B}
    double* d_flare_parts;             // per-block partial sums (BHR_FLARE_BLOCKS x 3)
    void* d_flare_params_own;          // device FlareParams formed from d_flare_sums (post.cu)

    int bloom_R; float sigma_scale;
    // TMA bloom kernels (bloom.cu): tensor maps of the disk layer (1-row boxes, H pass) and of the H-blurred
    // layer (64-column tiles, V pass), launch geometry; bloom_tma = 0 -> the generic kernels of post.cu
    unsigned char tmap_disk_row[128], tmap_hblur_tile[128], tmap_bg_tile[128], tmap_disk_tile[128];
    int bloom_tma, bloom_generic, bloom_h_P, bloom_h_bw, bloom_h_nb, bloom_h_segp, bloom_v_staged;
    int keep_blur;                     // also write blur_field from the fused V pass (it stays in registers otherwise)
    float* d_wtab;                     // 3 x wtab_stride (2R+1 weights + zero padding)
    int wtab_stride;
    float* d_wsum_x;                   // 3 x W   1 / in-bounds weight sums (summed in sequential f32 order)
    float* d_wsum_y;                   // 3 x H

    // disk-texture pipeline
    float* comp;                       // (13, n_r, n_phi)
    float *edge, *omega_rows, *row_stats;
    float stats[2];
    int bg_ready, az_freq; float az_shear;
    float* bg_rows;                    // 3 x n_r row quantities of the background kernel (omega, decay, shear)
    int background_scalar;             // option: the one-texel-per-thread background kernel
    int bg_blocks_override;            // option "background_blocks_per_sm" (0 = automatic)
    int bg_blocks_per_sm;              // resident blocks of the packed background kernel (background.cu)
    bhr_entity* d_entities; int entities_cap;      // 8-slot ring: entities + slot maps (texture.cu)
    float* d_entity_tables; size_t entity_tables_cap, entity_tables_n;   // tabulated profiles of caller-owned entities (kind 3 / 4)
    void* h_entities; double* d_coltab; int ent_ring; cudaEvent_t ent_ev[BHR_FRAME_SLOTS];
    float* stats_scratch;              // device statistics: density / structure planes, row results (stats.cu)
    void* stats_state;

    cudaEvent_t ev[6];
    cudaEvent_t frame_ev[BHR_FRAME_SLOTS]; // completion events of bhr_render_async slots
    cudaStream_t copy_stream;          // D2H of finished frames (bhr_render_async), overlaps the next frame
    // peer-memory tiled frame (peer.cu)
    int peer_rank, peer_world, peer_distributed; unsigned peer_serial;
    PeerSync* peer_sync_own; PeerSync* peer_sync[16]; PeerSync** d_peer_sync;
    float* peer_hblur[16]; float* peer_final_f32[16]; uint8_t* peer_final_u8[16];
    const float** d_row_src; void* d_flare_params;
    int peer_bounds[17];               // rank r renders rows [peer_bounds[r], peer_bounds[r + 1])
    void* peer_rows_host;              // pinned staging of the row-source table
    int* peer_host_err;                // host-mapped word a wait kernel writes when it gives up (peer.cu)
    double peer_timeout_ms;
    cudaEvent_t copy_done; int copy_pending;
    // pipelined tiled frames (peer.cu): completion events of the last four frames, "my egress copy has read the final buffers"
    cudaEvent_t tiled_ev[4]; cudaEvent_t tiled_copy_done; int tiled_copy_pending;
    // synchronous frames finished in row bands (api.cu): pieces per side of the photon-ring band (0 = off),
    // completion event per band, and "this launch continues a frame: keep the RK4 step total"
    int sync_bands; double sync_min_bytes, sync_extend; cudaEvent_t band_ev[12]; int keep_step_total;
    int strict_warps, band_box, planar, planar_attr_set;
    int bloom_h_P_option;              // option "bloom_h_p": pixels per lane of the TMA H pass (0 = by width)
    int raymarch_pair;                 // option "raymarch_pair": threads per block of the two-rays-per-thread ray march (0 = one ray per thread)
    int timeline; unsigned long long* d_timeline;   // option "timeline": per-block timestamps of the persistent ray march
    unsigned long long launches;       // kernels this context has launched (bhr_launch_count)
    int stage_timing;                  // record the per-stage timing events (instrumentation; each costs ~1.5 us of stream time)
    cudaEvent_t frame_done;            // orders the copy stream behind the composite (no timing)
    // device-side PNG stream (png.cu): code tables, per-segment staging, one stream buffer per frame slot
    void* d_png_tables; unsigned int* d_png_staging; unsigned int* d_png_seg; unsigned long long* d_png_off;
    uint8_t* d_png_stream[BHR_FRAME_SLOTS]; int png_n_seg; size_t png_capacity;
    // the entity layer runs on its own stream next to the background kernel (texture.cu): FP64-bound beside FP32-bound
    cudaEvent_t comp_read_ev; int comp_read_valid, entity_early;   // "the last kernel that reads the entity planes has been enqueued"
    cudaStream_t ent_stream; cudaEvent_t bg_start_ev, ent_done_ev; int bg_start_armed, ent_pending, entity_stream_on, entity_carveout_set;
    int ev_valid;
    float tint[3];
};

extern char g_bhr_create_error[512];

#define BHR_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e_));                                                    \
            return BHR_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define BHR_FAIL(ctx, code, ...)                                   \
    do {                                                           \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);     \
        return (code);                                             \
    } while (0)

// After a kernel that reads the entity planes of `comp` has been enqueued on the context's stream: the next entity layer
// (entity stream) may start once it is done.
static inline int bhr_mark_comp_read(bhr_ctx* ctx) {
    if (!ctx->entity_stream_on) return BHR_OK;
    if (!ctx->comp_read_ev) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comp_read_ev, cudaEventDisableTiming));
    BHR_CUDA(ctx, cudaEventRecord(ctx->comp_read_ev, ctx->stream));
    ctx->comp_read_valid = 1;
    return BHR_OK;
}

// Work on the context's stream that reads or writes the entity planes of `comp` first waits for the entity stream.
static inline int bhr_join_entities(bhr_ctx* ctx) {
    if (ctx->ent_pending) {
        BHR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ent_done_ev, 0));
        ctx->ent_pending = 0;
    }
    return BHR_OK;
}

// Every entry point that touches the device runs with the context's GPU current and restores the
// caller's device afterwards: a process may hold contexts on several GPUs, and hosts such as torch
// switch the current device between calls.
struct BhrDeviceGuard {
    int prev = -1, dev = -1;
    explicit BhrDeviceGuard(int device) : dev(device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (dev >= 0 && prev != dev) cudaSetDevice(dev);
    }
    explicit BhrDeviceGuard(const bhr_ctx* ctx) : BhrDeviceGuard(ctx ? ctx->cfg.device : -2) {}
    ~BhrDeviceGuard() { if (prev >= 0 && prev != dev && dev >= 0) cudaSetDevice(prev); }
    BhrDeviceGuard(const BhrDeviceGuard&) = delete;
    BhrDeviceGuard& operator=(const BhrDeviceGuard&) = delete;
};

static inline int bhr_div_up(int a, int b) { return (a + b - 1) / b; }

// row-tiled frame over several GPUs through peer memory (peer.cu): what the V pass / composite need
struct bhr_post_peer {
    const float* const* row_src;     // device array [H]: the hblur buffer (own or a peer's) that holds each row
    float* final_f32;                // where the finished rows go (rank 0's buffers; NULL = own)
    uint8_t* final_u8;
    const void* flare_params;        // device FlareParams reduced from all ranks' sums, or NULL
    int (*before_composite)(bhr_ctx*);   // enqueue the back-pressure wait before the peer stores
};
// kernels / launchers implemented in the other translation units
int bhr_launch_raymarch(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, int row0, int row1);
int bhr_launch_bloom_h(bhr_ctx* ctx, int row0, int row1);
int bhr_launch_bloom_v_composite(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const double* flare_sums_host);
int bhr_launch_bloom_v_composite_ex(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const double* flare_sums_host,
                                    const bhr_post_peer* peer);
int bhr_launch_flare_params(bhr_ctx* ctx, const double* d_parts, int world, void* d_flare_params);
size_t bhr_flare_params_size();
int bhr_launch_flare_sums(bhr_ctx* ctx, int row0, int row1);
int bhr_launch_disk_post(bhr_ctx* ctx, float* out);
int bhr_setup_background(bhr_ctx* ctx);
int bhr_launch_background(bhr_ctx* ctx, float t);
int bhr_launch_noise_packed(bhr_ctx* ctx, const float* d_coords, int n, float* d_out);
int bhr_launch_build_mips(bhr_ctx* ctx, int numpy_order);
int bhr_setup_bloom_tables(bhr_ctx* ctx);
int bhr_setup_bloom_tma(bhr_ctx* ctx);
int bhr_launch_bloom_h_tma(bhr_ctx* ctx, int row0, int row1);
int bhr_launch_bloom_v_fused(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const void* F_host, const void* F_dev,
                             const float* const* row_src, float* dst_f32, uint8_t* dst_u8);
int bhr_launch_blur_only(bhr_ctx* ctx);
int bhr_launch_png_encode(bhr_ctx* ctx, int slot, cudaStream_t stream);
