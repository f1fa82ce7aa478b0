// api.cu -- context lifecycle and the C-ABI entry points of libbhr.so (see include/bhr.h).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

char g_bhr_create_error[512] = "";
extern int bhr_raymarch_mode_override;

#define CREATE_CHECK(call)                                                                              \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            snprintf(g_bhr_create_error, sizeof(g_bhr_create_error), "%s: %s", #call, cudaGetErrorString(e_)); \
            bhr_destroy(ctx);                                                                           \
            return BHR_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

// _color_temp_to_tint(DISK_COLOR_TEMPERATURE = 6000), render.py:2407-2437, ideal-libm convention
static void tint_6000(float out[3]) {
    float t = 6000.0f / 100.0f;
    out[0] = 1.0f;
    float g = 0.390082f * (float)log((double)fmaxf(t, 0.0001f)) - 0.631841f;
    out[1] = fminf(fmaxf(g, 0.0f), 1.0f);
    float b = 0.543207f * (float)log((double)fmaxf(t - 10.0f, 0.0001f)) - 1.19625f;
    out[2] = fminf(fmaxf(b, 0.0f), 1.0f);
}

extern "C" int bhr_version(void) { return 100; }

extern "C" int bhr_create(const bhr_config* cfg, bhr_ctx** out) {
    if (!cfg || !out) { snprintf(g_bhr_create_error, sizeof(g_bhr_create_error), "null argument"); return BHR_ERR_INVALID; }
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->step_size <= 0.0f || cfg->r_disk_inner >= cfg->r_disk_outer) {
        snprintf(g_bhr_create_error, sizeof(g_bhr_create_error), "invalid configuration");
        return BHR_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(g_bhr_create_error, sizeof(g_bhr_create_error), "no CUDA device available (%s); libbhr has no CPU fallback",
                 cudaGetErrorString(e));
        return BHR_ERR_CUDA;
    }
    bhr_ctx* ctx = (bhr_ctx*)calloc(1, sizeof(bhr_ctx));
    if (!ctx) return BHR_ERR_NOMEM;
    ctx->cfg = *cfg;
    ctx->W = cfg->width; ctx->H = cfg->height;
    BhrDeviceGuard device_guard_(cfg->device);          // (the caller's current device is restored on return)
    CREATE_CHECK(cudaSetDevice(cfg->device));
    CREATE_CHECK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    const size_t plane = (size_t)ctx->W * ctx->H;
    CREATE_CHECK(cudaMalloc(&ctx->bg, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMalloc(&ctx->disk, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMalloc(&ctx->hblur, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMalloc(&ctx->blur, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMalloc(&ctx->final_f32, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMalloc(&ctx->final_u8, plane * 3));
    CREATE_CHECK(cudaMalloc(&ctx->cls, plane));
    CREATE_CHECK(cudaMalloc(&ctx->steps, plane * sizeof(int)));
    CREATE_CHECK(cudaMalloc(&ctx->d_flare_sums, 3 * sizeof(double)));
    CREATE_CHECK(cudaMalloc(&ctx->d_flare_parts, BHR_FLARE_BLOCKS * 3 * sizeof(double)));
    CREATE_CHECK(cudaMalloc(&ctx->d_flare_params_own, bhr_flare_params_size()));
    // [0, plane) u64 re-trace entries of the safety net; then plane i32 band pixels
    CREATE_CHECK(cudaMalloc(&ctx->retrace_queue, plane * (sizeof(unsigned long long) + sizeof(int))));
    CREATE_CHECK(cudaMemset(ctx->retrace_queue, 0, plane * (sizeof(unsigned long long) + sizeof(int))));
    CREATE_CHECK(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
    ctx->persistent = 1;
    ctx->pblock_big = 0;        // 768 (512 with differentials) threads per persistent block: measured best at every size
    // [0] re-trace queue tail, [1] band count, [2] band head, [3] tile counter, [4..5] u64 step total
    CREATE_CHECK(cudaMalloc(&ctx->d_queue_count, 8 * sizeof(unsigned int)));
    CREATE_CHECK(cudaMemset(ctx->d_queue_count, 0, 8 * sizeof(unsigned int)));
    ctx->d_total_steps = (unsigned long long*)(ctx->d_queue_count + 4);
    ctx->retrace_min_cross = 3;
    ctx->entity_stream_on = 1;
    ctx->entity_early = 1;
    ctx->retrace_band = 0.02f;
    ctx->band_lo_auto = 1;
    ctx->sync_bands = 1;
    ctx->sync_min_bytes = 16e6;
    ctx->sync_extend = 0.1;
    ctx->strict_warps = 32;
    ctx->band_box = 1;
    ctx->stage_timing = 1;
    CREATE_CHECK(cudaMemset(ctx->bg, 0, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMemset(ctx->disk, 0, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMemset(ctx->hblur, 0, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMemset(ctx->blur, 0, plane * 3 * sizeof(float)));
    CREATE_CHECK(cudaMemset(ctx->cls, 0, plane));
    CREATE_CHECK(cudaMemset(ctx->steps, 0, plane * sizeof(int)));
    for (int k = 0; k < 6; ++k) CREATE_CHECK(cudaEventCreate(&ctx->ev[k]));
    tint_6000(ctx->tint);
    ctx->stats[0] = 0.5f; ctx->stats[1] = 0.5f;    // render.py:3533
    if (bhr_setup_bloom_tables(ctx) != BHR_OK || bhr_setup_bloom_tma(ctx) != BHR_OK) {
        snprintf(g_bhr_create_error, sizeof(g_bhr_create_error), "%s", ctx->err);
        bhr_destroy(ctx);
        return BHR_ERR_CUDA;
    }
    *out = ctx;
    return BHR_OK;
}

extern "C" void bhr_destroy(bhr_ctx* ctx) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return;
    bhr_peer_detach(ctx);
    if (ctx->peer_sync_own) cudaFree(ctx->peer_sync_own);
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->own_stream); }
    if (ctx->ent_stream) cudaStreamSynchronize(ctx->ent_stream);
    void* ptrs[] = {ctx->sky, ctx->mips, ctx->tex_staging, ctx->bg, ctx->disk, ctx->hblur, ctx->blur, ctx->final_f32,
                    ctx->final_u8, ctx->cls, ctx->steps, ctx->d_flare_sums, ctx->d_flare_parts, ctx->d_flare_params_own, ctx->d_timeline, ctx->retrace_queue, ctx->d_queue_count, ctx->d_wtab, ctx->d_wsum_x,
                    ctx->d_wsum_y, ctx->comp, ctx->bg_rows, ctx->edge, ctx->omega_rows, ctx->row_stats, ctx->d_entity_tables, ctx->d_entities, ctx->stats_scratch, ctx->stats_state, ctx->d_coltab};
    for (void* p : ptrs) if (p) cudaFree(p);
    void* png[] = {ctx->d_png_tables, ctx->d_png_staging, ctx->d_png_seg, ctx->d_png_off};
    for (void* p : png) if (p) cudaFree(p);
    for (int k = 0; k < BHR_FRAME_SLOTS; ++k) if (ctx->d_png_stream[k]) cudaFree(ctx->d_png_stream[k]);
    for (int k = 0; k < 6; ++k) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (int k = 0; k < BHR_FRAME_SLOTS; ++k) if (ctx->ent_ev[k]) cudaEventDestroy(ctx->ent_ev[k]);
    for (int k = 0; k < BHR_FRAME_SLOTS; ++k) if (ctx->frame_ev[k]) cudaEventDestroy(ctx->frame_ev[k]);
    for (int k = 0; k < 12; ++k) if (ctx->band_ev[k]) cudaEventDestroy(ctx->band_ev[k]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->frame_done) cudaEventDestroy(ctx->frame_done);
    if (ctx->ent_stream) { cudaStreamSynchronize(ctx->ent_stream); cudaStreamDestroy(ctx->ent_stream); }
    if (ctx->bg_start_ev) cudaEventDestroy(ctx->bg_start_ev);
    if (ctx->comp_read_ev) cudaEventDestroy(ctx->comp_read_ev);
    if (ctx->ent_done_ev) cudaEventDestroy(ctx->ent_done_ev);
    if (ctx->h_entities) cudaFreeHost(ctx->h_entities);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx);
}

extern "C" const char* bhr_last_error(const bhr_ctx* ctx) { return ctx ? ctx->err : g_bhr_create_error; }

extern "C" int bhr_set_stream(bhr_ctx* ctx, void* cuda_stream) {
    if (!ctx) return BHR_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return BHR_OK;
}

extern "C" int bhr_synchronize(bhr_ctx* ctx) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return BHR_ERR_INVALID;
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) BHR_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->ent_stream) BHR_CUDA(ctx, cudaStreamSynchronize(ctx->ent_stream));
    return BHR_OK;
}

extern "C" int bhr_set_lens_flare(bhr_ctx* ctx, int enabled) {
    if (!ctx) return BHR_ERR_INVALID;
    ctx->cfg.lens_flare = enabled ? 1 : 0;
    return BHR_OK;
}

extern "C" int bhr_set_option(bhr_ctx* ctx, const char* key, double value) {
    if (!key) return BHR_ERR_INVALID;
    if (!strcmp(key, "raymarch_mode")) { bhr_raymarch_mode_override = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "retrace_min_cross")) { ctx->retrace_min_cross = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "retrace_band")) { ctx->retrace_band = (float)value; return BHR_OK; }
    if (ctx && !strcmp(key, "persistent")) { ctx->persistent = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "band_lo_auto")) { ctx->band_lo_auto = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "pblock_big")) { ctx->pblock_big = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "sync_bands")) { ctx->sync_bands = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "sync_extend")) { ctx->sync_extend = value; return BHR_OK; }
    if (ctx && !strcmp(key, "sync_min_bytes")) { ctx->sync_min_bytes = value; return BHR_OK; }
    if (ctx && !strcmp(key, "planar")) { ctx->planar = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "stage_timing")) { ctx->stage_timing = (int)value != 0; return BHR_OK; }
    if (ctx && !strcmp(key, "band_box")) { ctx->band_box = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "strict_warps")) { ctx->strict_warps = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "peer_timeout_ms")) { ctx->peer_timeout_ms = value; return BHR_OK; }
    if (ctx && !strcmp(key, "bloom_generic")) { ctx->bloom_generic = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "keep_blur")) { ctx->keep_blur = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "timeline")) { ctx->timeline = (int)value; return BHR_OK; }
    if (ctx && !strcmp(key, "bloom_h_p")) {               // pixels per lane of the TMA H pass: 0 = by width, 5, 10
        ctx->bloom_h_P_option = (int)value;
        return bhr_setup_bloom_tma(ctx);
    }
    if (ctx && !strcmp(key, "raymarch_pair")) {          // 0 = one ray per thread; 384 / 448 / 512 = threads per block, two rays each
        const int v = (int)value;
        if (v != 0 && v != 384 && v != 448 && v != 512) BHR_FAIL(ctx, BHR_ERR_INVALID, "raymarch_pair must be 0, 384, 448 or 512");
        ctx->raymarch_pair = v;
        return BHR_OK;
    }
    if (ctx && !strcmp(key, "background_blocks_per_sm")) {      // resident blocks of the packed background kernel (experiments)
        if (value < 0) BHR_FAIL(ctx, BHR_ERR_INVALID, "background_blocks_per_sm must be >= 0 (0 = automatic)");
        ctx->bg_blocks_override = (int)value;
        return BHR_OK;
    }
    if (ctx && !strcmp(key, "entity_early")) { ctx->entity_early = value != 0.0; return BHR_OK; }
    if (ctx && !strcmp(key, "entity_stream")) {          // 1 (default): the entity layer runs on its own stream beside the background kernel
        if (int rc = bhr_join_entities(ctx)) return rc;
        ctx->entity_stream_on = value != 0.0;
        return bhr_mark_comp_read(ctx);          // (turned on: the next entity layer starts behind everything enqueued so far)
    }
    if (ctx && !strcmp(key, "background_scalar")) { ctx->background_scalar = (int)value; return BHR_OK; }
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "unknown option %s", key);
    return BHR_ERR_INVALID;
}

extern "C" int bhr_bloom_radius(const bhr_ctx* ctx) { return ctx ? ctx->bloom_R : -1; }

static int check_rows(bhr_ctx* ctx, int row0, int row1) {
    if (row0 < 0 || row1 > ctx->H || row0 > row1) BHR_FAIL(ctx, BHR_ERR_INVALID, "bad row range [%d, %d)", row0, row1);
    return BHR_OK;
}

extern "C" int bhr_render_rows_stage1(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, int row0, int row1) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !cam) return BHR_ERR_INVALID;
    int rc = check_rows(ctx, row0, row1);
    if (rc) return rc;
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = bhr_launch_raymarch(ctx, cam, flags, row0, row1);
    if (rc) return rc;
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (!(flags & BHR_SKIP_BLOOM)) {
        rc = bhr_launch_bloom_h(ctx, row0, row1);
        if (rc) return rc;
    }
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_flare_sums(bhr_ctx* ctx, int row0, int row1, double out[3]) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    int rc = check_rows(ctx, row0, row1);
    if (rc) return rc;
    rc = bhr_launch_flare_sums(ctx, row0, row1);
    if (rc) return rc;
    BHR_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_flare_sums, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_flare_sums_device(bhr_ctx* ctx, int row0, int row1) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return BHR_ERR_INVALID;
    int rc = check_rows(ctx, row0, row1);
    if (rc) return rc;
    return bhr_launch_flare_sums(ctx, row0, row1);
}

extern "C" int bhr_render_rows_stage2(bhr_ctx* ctx, uint32_t flags, int row0, int row1, const double* flare_sums) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return BHR_ERR_INVALID;
    int rc = check_rows(ctx, row0, row1);
    if (rc) return rc;
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    rc = bhr_launch_bloom_v_composite(ctx, flags, row0, row1, flare_sums);
    if (rc) return rc;
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    ctx->ev_valid = ctx->stage_timing;
    return BHR_OK;
}

// copy_stream != nullptr: the D2H copies run on that stream (after the composite), so that the
// next frame's kernels overlap them; the next composite waits for ctx->copy_done before it
// overwrites the final buffers (post.cu).
static int render_enqueue(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8,
                          cudaStream_t copy_stream) {
    int rc = bhr_render_rows_stage1(ctx, cam, flags, 0, ctx->H);
    if (rc) return rc;
    if (ctx->cfg.lens_flare && !(flags & BHR_SKIP_FLARE)) {
        // the flare needs the global brightness centroid before any pixel can be finished: the three
        // sums and the parameters derived from them stay on the device (no host round trip)
        rc = bhr_launch_flare_sums(ctx, 0, ctx->H);
        if (rc) return rc;
        flags |= BHR_FLARE_FROM_DEVICE;
    }
    rc = bhr_render_rows_stage2(ctx, flags, 0, ctx->H, nullptr);
    if (rc) return rc;
    cudaStream_t cs = ctx->stream;
    if (copy_stream && (out_f32 || out_u8)) {
        cs = copy_stream;
        if (!ctx->frame_done) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->frame_done, cudaEventDisableTiming));
        BHR_CUDA(ctx, cudaEventRecord(ctx->frame_done, ctx->stream));
        BHR_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->frame_done, 0));
    }
    const size_t n3 = (size_t)ctx->W * ctx->H * 3;
    if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32, ctx->final_f32, n3 * sizeof(float), cudaMemcpyDeviceToHost, cs));
    if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8, ctx->final_u8, n3, cudaMemcpyDeviceToHost, cs));
    if (cs != ctx->stream) {
        if (!ctx->copy_done) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
        BHR_CUDA(ctx, cudaEventRecord(ctx->copy_done, cs));
        ctx->copy_pending = 1;
    }
    return BHR_OK;
}

// ---------------------------------------------------------------------------------------------
// Synchronous frame into host memory, finished in row bands.  The D2H copy of a float fhd frame
// (24.9 MB, 0.44 ms over PCIe 5) costs almost as much as rendering it (0.67 ms); a frame whose
// rows are finished band by band can copy the first bands out while the later ones are still
// being traced.  Any split reproduces the one-shot frame bit for bit (every pixel is independent
// up to the bloom's +-R rows of the H-blurred layer, and a row is finished only when those are
// there), so the plan below is a pure scheduling decision:
//   * band 0 = the rows that contain the photon ring.  All ill-conditioned rays (strict
//     integrator) live there; a launch that holds any of them lasts at least one
//     strict batch (~0.19 ms), so they go into one launch, first, where that latency is hidden;
//   * the rows above and below it are traced afterwards in `sync_bands` further pieces each.
// The ring is a circle around the image centre (build_camera looks at the hole): a ray at angle
// theta from the forward axis has L = r sin(theta) and impact parameter b = L / sqrt(1 - L^2/r^3),
// and the band is |b / b_crit - 1| < retrace_band.
// ---------------------------------------------------------------------------------------------
static int plan_bands(const bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, size_t out_bytes, int bands[][2], int max_bands) {
    const int H = ctx->H, R = (flags & BHR_SKIP_BLOOM) ? 0 : ctx->bloom_R;
    bands[0][0] = 0; bands[0][1] = H;
    // every extra launch costs ~50 us (tile-queue tail, band list, launch gaps): bands pay when the
    // copy they hide is longer than that -- a float fhd frame or any 4K frame, not an 8-bit fhd one
    if (ctx->sync_bands <= 0 || (double)out_bytes < ctx->sync_min_bytes) return 1;
    if (ctx->cfg.lens_flare && !(flags & BHR_SKIP_FLARE)) return 1;              // the flare needs the whole disk layer first
    const double r = sqrt((double)cam->pos[0] * cam->pos[0] + (double)cam->pos[1] * cam->pos[1] + (double)cam->pos[2] * cam->pos[2]);
    const double b = 2.598076211353316 * (1.0 + (double)ctx->retrace_band + 2e-3);
    const double L2 = b * b / (1.0 + b * b / (r * r * r));
    const double s2 = L2 / (r * r);
    if (!(r > 1.6) || !(s2 < 0.98) || !(cam->pixel_h > 0.0f)) return 1;
    const double rho = sqrt(s2 / (1.0 - s2)) / (double)cam->pixel_h;          // ring radius in pixels
    int ya = (int)floor(0.5 * H - 0.5 - rho) - 1, yb = (int)ceil(0.5 * H - 0.5 + rho) + 2;
    // the ring band is bounded by the latency of its strict batches, not by its fast rays: rows
    // added below it are traced for free and shorten the last band (the one whose copy nothing hides)
    yb += (int)((yb - ya) * ctx->sync_extend);
    ya = ya < 0 ? 0 : ya; yb = yb > H ? H : yb;
    const int min_rows = R + 32;
    if (ya < min_rows && H - yb < min_rows) return 1;                           // the ring fills the frame
    if (ya < min_rows) ya = 0;
    if (H - yb < min_rows) yb = H;
    int n = 0;
    bands[n][0] = ya; bands[n][1] = yb; ++n;
    const int pieces = ctx->sync_bands < 5 ? ctx->sync_bands : 5;      // 1 + 2 * 5 bands fit max_bands
    // pieces next to the ring first: together with it they complete the rows around its edges
    for (int side = 0; side < 2; ++side) {
        const int lo = side == 0 ? 0 : yb, hi = side == 0 ? ya : H;
        const int rows = hi - lo;
        if (rows <= 0) continue;
        int k = pieces;
        while (k > 1 && rows / k < min_rows) --k;
        for (int i = 0; i < k && n < max_bands; ++i) {
            // side 0 runs from the ring upwards, side 1 from the ring downwards
            const int a = (int)((long long)rows * i / k), c = (int)((long long)rows * (i + 1) / k);
            if (side == 0) { bands[n][0] = hi - c; bands[n][1] = hi - a; }
            else { bands[n][0] = lo + a; bands[n][1] = lo + c; }
            ++n;
        }
    }
    return n;
}

static int render_sync_banded(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8) {
    enum { MAXB = 12 };
    int bands[MAXB][2];
    const size_t out_bytes = (size_t)ctx->W * ctx->H * 3 * ((out_f32 ? sizeof(float) : 0) + (out_u8 ? 1 : 0));
    const int nb = plan_bands(ctx, cam, flags, out_bytes, bands, MAXB);
    if (nb <= 1) {
        int rc = render_enqueue(ctx, cam, flags, out_f32, out_u8, nullptr);
        if (rc) return rc;
        BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return BHR_OK;
    }
    const int H = ctx->H, R = (flags & BHR_SKIP_BLOOM) ? 0 : ctx->bloom_R;
    if (!ctx->copy_stream) BHR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // rows traced / finished so far; runs[] = the row ranges finished after each band
    std::vector<uint8_t> traced((size_t)H, 0), finished((size_t)H, 0);
    std::vector<int> pre((size_t)H + 1, 0);               // prefix sums of traced[]
    int runs[MAXB][4][2], n_runs[MAXB];
    int rc = BHR_OK;
    const size_t row_f32 = (size_t)ctx->W * 3 * sizeof(float), row_u8 = (size_t)ctx->W * 3;
    // every band is enqueued before the first copy: a copy into pageable memory blocks the host
    for (int k = 0; k < nb && !rc; ++k) {
        ctx->keep_step_total = k > 0;
        // the ring band alone is bounded by the latency of its strict batches: spread them over more
        // SMs (20 instead of 28 strict warps per block: 321 -> 299 us for the fhd ring band)
        const int strict_warps = ctx->strict_warps;
        if (k == 0 && strict_warps > 20) ctx->strict_warps = 20;
        rc = bhr_render_rows_stage1(ctx, cam, flags, bands[k][0], bands[k][1]);
        ctx->strict_warps = strict_warps;
        ctx->keep_step_total = 0;
        if (rc) break;
        for (int y = bands[k][0]; y < bands[k][1]; ++y) traced[y] = 1;
        pre[0] = 0;
        for (int y = 0; y < H; ++y) pre[y + 1] = pre[y] + traced[y];
        // a row can be finished when it is not finished yet and its +-R neighbourhood is traced
        auto ready = [&](int yy) {
            const int a = yy - R < 0 ? 0 : yy - R, c = yy + R >= H ? H - 1 : yy + R;
            return !finished[yy] && pre[c + 1] - pre[a] == c - a + 1;
        };
        n_runs[k] = 0;
        int y = 0;
        while (y < H && !rc) {
            if (!ready(y)) { ++y; continue; }
            int e = y;
            while (e < H && ready(e)) ++e;
            if (n_runs[k] == 4) break;                                           // (cannot happen: <= 2 runs per band)
            rc = bhr_render_rows_stage2(ctx, flags, y, e, nullptr);
            for (int t = y; t < e; ++t) finished[t] = 1;
            runs[k][n_runs[k]][0] = y; runs[k][n_runs[k]][1] = e; ++n_runs[k];
            y = e;
        }
        if (rc) break;
        if (!ctx->band_ev[k]) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_ev[k], cudaEventDisableTiming));
        BHR_CUDA(ctx, cudaEventRecord(ctx->band_ev[k], ctx->stream));
    }
    int all = 1;
    for (int y = 0; y < H; ++y) all &= finished[y];
    if (rc) return rc;
    if (!all) BHR_FAIL(ctx, BHR_ERR_STATE, "band plan left rows unfinished");
    for (int k = 0; k < nb; ++k) {
        BHR_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->band_ev[k], 0));
        for (int i = 0; i < n_runs[k]; ++i) {
            const size_t y0 = (size_t)runs[k][i][0], rows = (size_t)(runs[k][i][1] - runs[k][i][0]);
            if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync((char*)out_f32 + y0 * row_f32, (const char*)ctx->final_f32 + y0 * row_f32,
                                                       rows * row_f32, cudaMemcpyDeviceToHost, ctx->copy_stream));
            if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8 + y0 * row_u8, ctx->final_u8 + y0 * row_u8, rows * row_u8,
                                                      cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
    }
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_render(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !cam) return BHR_ERR_INVALID;
    if (out_f32 || out_u8) return render_sync_banded(ctx, cam, flags, out_f32, out_u8);
    return render_enqueue(ctx, cam, flags, nullptr, nullptr, nullptr);
}

extern "C" int bhr_render_async(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8, int slot) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !cam || slot < 0 || slot >= BHR_FRAME_SLOTS) return BHR_ERR_INVALID;
    if (!ctx->copy_stream) BHR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    int rc = render_enqueue(ctx, cam, flags, out_f32, out_u8, ctx->copy_stream);
    if (rc) return rc;
    if (!ctx->frame_ev[slot]) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->frame_ev[slot], cudaEventDisableTiming));
    BHR_CUDA(ctx, cudaEventRecord(ctx->frame_ev[slot], (out_f32 || out_u8) ? ctx->copy_stream : ctx->stream));
    return BHR_OK;
}

// A frame whose 8-bit image leaves the device as the zlib stream of its PNG file (png.cu).  host receives
// {uint32 stream_bytes, uint32 adler32} followed by the first copy_bytes of the stream; a caller that guessed
// copy_bytes too small (stream_bytes > copy_bytes) fetches the rest with bhr_png_fetch before it reuses the slot.
extern "C" int bhr_render_async_png(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, void* host, size_t copy_bytes, int slot) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !cam || !host || slot < 0 || slot >= BHR_FRAME_SLOTS) return BHR_ERR_INVALID;
    if (!ctx->copy_stream) BHR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    int rc = render_enqueue(ctx, cam, flags, nullptr, nullptr, nullptr);
    if (rc) return rc;
    // The encoder stays on the context's stream.  (On the copy stream it would have to share the SMs with the next frame's
    // persistent ray march, which leaves no room: measured, the encoder then runs after that ray march and the next
    // composite -- which must wait until the 8-bit frame has been read -- stalls: 1.34 instead of 1.12 ms per video frame.)
    rc = bhr_launch_png_encode(ctx, slot, ctx->stream);
    if (rc) return rc;
    if (copy_bytes > ctx->png_capacity) copy_bytes = ctx->png_capacity;
    if (!ctx->frame_done) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->frame_done, cudaEventDisableTiming));
    BHR_CUDA(ctx, cudaEventRecord(ctx->frame_done, ctx->stream));
    BHR_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->frame_done, 0));
    BHR_CUDA(ctx, cudaMemcpyAsync(host, ctx->d_png_stream[slot], 8 + copy_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (!ctx->frame_ev[slot]) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->frame_ev[slot], cudaEventDisableTiming));
    BHR_CUDA(ctx, cudaEventRecord(ctx->frame_ev[slot], ctx->copy_stream));
    return BHR_OK;
}

// bytes [offset, offset + bytes) of slot's stream, synchronously (the slot's frame must have been waited for)
extern "C" int bhr_png_fetch(bhr_ctx* ctx, int slot, size_t offset, size_t bytes, void* host) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !host || slot < 0 || slot >= BHR_FRAME_SLOTS) return BHR_ERR_INVALID;
    if (!ctx->d_png_stream[slot]) BHR_FAIL(ctx, BHR_ERR_STATE, "no PNG stream was encoded in slot %d", slot);
    if (offset + bytes > ctx->png_capacity) BHR_FAIL(ctx, BHR_ERR_INVALID, "range past the stream buffer");
    BHR_CUDA(ctx, cudaMemcpy(host, ctx->d_png_stream[slot] + 8 + offset, bytes, cudaMemcpyDeviceToHost));
    return BHR_OK;
}

// the stream of the frame currently in BHR_BUF_FINAL_U8 (after a bhr_render), synchronously: tests and one-off files
extern "C" int bhr_png_encode_current(bhr_ctx* ctx, void* host, size_t host_bytes, uint32_t* stream_bytes) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !host || !stream_bytes) return BHR_ERR_INVALID;
    int rc = bhr_launch_png_encode(ctx, 0, ctx->stream);
    if (rc) return rc;
    uint32_t info[2];
    BHR_CUDA(ctx, cudaMemcpyAsync(info, ctx->d_png_stream[0], 8, cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *stream_bytes = info[0];
    if (info[0] > host_bytes) BHR_FAIL(ctx, BHR_ERR_INVALID, "stream of %u bytes does not fit the buffer", info[0]);
    BHR_CUDA(ctx, cudaMemcpy(host, ctx->d_png_stream[0] + 8, info[0], cudaMemcpyDeviceToHost));
    return BHR_OK;
}

extern "C" int bhr_wait_frame(bhr_ctx* ctx, int slot) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || slot < 0 || slot >= BHR_FRAME_SLOTS) return BHR_ERR_INVALID;
    if (!ctx->frame_ev[slot]) BHR_FAIL(ctx, BHR_ERR_STATE, "no frame was enqueued in slot %d", slot);
    BHR_CUDA(ctx, cudaEventSynchronize(ctx->frame_ev[slot]));
    return BHR_OK;
}

extern "C" int bhr_buffer(bhr_ctx* ctx, int id, void** dev_ptr, size_t* bytes) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !dev_ptr || !bytes) return BHR_ERR_INVALID;
    const size_t plane = (size_t)ctx->W * ctx->H;
    const size_t tex = (size_t)ctx->n_r * ctx->n_phi;
    switch (id) {
        case BHR_BUF_BG: *dev_ptr = ctx->bg; *bytes = plane * 12; break;
        case BHR_BUF_DISK: *dev_ptr = ctx->disk; *bytes = plane * 12; break;
        case BHR_BUF_HBLUR: *dev_ptr = ctx->hblur; *bytes = plane * 12; break;
        case BHR_BUF_BLUR: *dev_ptr = ctx->blur; *bytes = plane * 12; break;
        case BHR_BUF_FINAL: *dev_ptr = ctx->final_f32; *bytes = plane * 12; break;
        case BHR_BUF_FINAL_U8: *dev_ptr = ctx->final_u8; *bytes = plane * 3; break;
        case BHR_BUF_CLASS: *dev_ptr = ctx->cls; *bytes = plane; break;
        case BHR_BUF_STEPS: *dev_ptr = ctx->steps; *bytes = plane * 4; break;
        case BHR_BUF_DISK_TEX: *dev_ptr = ctx->mips; *bytes = tex * 16; break;
        case BHR_BUF_DISK_MIPS: *dev_ptr = ctx->mips; *bytes = (size_t)ctx->level_off[BHR_NUM_MIPS] * 16; break;
        case BHR_BUF_COMP: if (int rc = bhr_join_entities(ctx)) return rc; *dev_ptr = ctx->comp; *bytes = ctx->comp ? tex * BHR_N_COMP * 4 : 0; break;
        case BHR_BUF_FLARE_SUMS: *dev_ptr = ctx->d_flare_sums; *bytes = 3 * sizeof(double); break;
        default: BHR_FAIL(ctx, BHR_ERR_INVALID, "unknown buffer id %d", id);
    }
    if (!*dev_ptr) BHR_FAIL(ctx, BHR_ERR_STATE, "buffer %d not allocated yet", id);
    return BHR_OK;
}

extern "C" int bhr_download(bhr_ctx* ctx, int id, void* host, size_t bytes) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !host) return BHR_ERR_INVALID;
    void* d; size_t n;
    if ((id == BHR_BUF_BLUR || id == BHR_BUF_DISK_POST) && ctx->bloom_tma && !(ctx->bloom_generic & 2) && !ctx->keep_blur) {
        // the fused V pass keeps `blur` in registers: form blur_field now from the H-blurred layer of the last frame
        int rc = bhr_launch_blur_only(ctx);
        if (rc) return rc;
    }
    if (id == BHR_BUF_DISK_POST) {
        // formed on demand (an inspection path: the reference's disk_layer_field after a bloomed frame)
        n = (size_t)ctx->W * ctx->H * 12;
        if (bytes > n) BHR_FAIL(ctx, BHR_ERR_INVALID, "buffer %d holds %zu bytes, %zu requested", id, n, bytes);
        float* tmp = nullptr;
        BHR_CUDA(ctx, cudaMalloc(&tmp, n));
        int rc = bhr_launch_disk_post(ctx, tmp);
        cudaError_t e = rc ? cudaSuccess : cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(tmp);
        if (rc) return rc;
        BHR_CUDA(ctx, e);
        return BHR_OK;
    }
    int rc = bhr_buffer(ctx, id, &d, &n);
    if (rc) return rc;
    if (bytes > n) BHR_FAIL(ctx, BHR_ERR_INVALID, "buffer %d holds %zu bytes, %zu requested", id, n, bytes);
    BHR_CUDA(ctx, cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BHR_OK;
}

namespace {
// RK4 evaluations per image row (the cost model of the row tiles): block per row
__global__ void __launch_bounds__(256) row_cost_kernel(const int* __restrict__ steps, int W, int row0,
                                                       unsigned long long* __restrict__ out) {
    const int* row = steps + (size_t)(row0 + blockIdx.x) * W;
    unsigned long long s = 0;
    for (int x = threadIdx.x; x < W; x += 256) s += (unsigned long long)row[x];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    __shared__ unsigned long long sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        out[blockIdx.x] = t;
    }
}
}  // namespace

extern "C" int bhr_row_costs(bhr_ctx* ctx, int row0, int row1, uint64_t* out) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    int rc = check_rows(ctx, row0, row1);
    if (rc) return rc;
    if (row1 == row0) return BHR_OK;
    unsigned long long* d = nullptr;
    BHR_CUDA(ctx, cudaMalloc(&d, (size_t)(row1 - row0) * sizeof(unsigned long long)));
    row_cost_kernel<<<row1 - row0, 256, 0, ctx->stream>>>(ctx->steps, ctx->W, row0, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, (size_t)(row1 - row0) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    BHR_CUDA(ctx, e);
    return BHR_OK;
}

extern "C" int bhr_last_raymarch_timeline(bhr_ctx* ctx, uint64_t* out, int max_blocks) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out || !ctx->d_timeline) return 0;
    const int n = max_blocks < ctx->num_sms ? max_blocks : ctx->num_sms;
    if (cudaMemcpyAsync(out, ctx->d_timeline, (size_t)3 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return 0;
    return n;
}

extern "C" int bhr_launch_count(bhr_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return BHR_ERR_INVALID;
    *out = ctx->launches;
    return BHR_OK;
}

extern "C" int bhr_last_total_steps(bhr_ctx* ctx, uint64_t* out) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    unsigned long long v = 0;
    BHR_CUDA(ctx, cudaMemcpyAsync(&v, ctx->d_total_steps, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = v;
    return BHR_OK;
}

extern "C" int bhr_last_retrace_count(bhr_ctx* ctx, uint32_t* out) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    unsigned int v[2] = {0, 0};
    BHR_CUDA(ctx, cudaMemcpyAsync(v, ctx->d_queue_count, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = v[0];
    return BHR_OK;
}

extern "C" int bhr_last_stage_ms(bhr_ctx* ctx, float out[5]) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    if (!ctx->ev_valid) BHR_FAIL(ctx, BHR_ERR_STATE, "no frame rendered with stage timing on (option \"stage_timing\")");
    BHR_CUDA(ctx, cudaEventSynchronize(ctx->ev[4]));
    BHR_CUDA(ctx, cudaEventElapsedTime(&out[0], ctx->ev[0], ctx->ev[1]));   // ray march
    BHR_CUDA(ctx, cudaEventElapsedTime(&out[1], ctx->ev[1], ctx->ev[2]));   // bloom H
    BHR_CUDA(ctx, cudaEventElapsedTime(&out[2], ctx->ev[3], ctx->ev[4]));   // bloom V + composite (+ flare)
    BHR_CUDA(ctx, cudaEventElapsedTime(&out[3], ctx->ev[2], ctx->ev[3]));   // flare reduction / halo gap
    BHR_CUDA(ctx, cudaEventElapsedTime(&out[4], ctx->ev[0], ctx->ev[4]));   // total
    return BHR_OK;
}

extern "C" int bhr_device_pci_bus_id(int device, char* out, int len) {
    if (!out || len < 16) return BHR_ERR_INVALID;
    return cudaDeviceGetPCIBusId(out, len, device) == cudaSuccess ? BHR_OK : BHR_ERR_CUDA;
}

// pinned host memory for zero-staging D2H of frames
extern "C" int bhr_host_alloc(size_t bytes, void** out) {
    if (!out) return BHR_ERR_INVALID;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? BHR_OK : BHR_ERR_NOMEM;
}
extern "C" int bhr_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? BHR_OK : BHR_ERR_CUDA; }
// page-lock memory the caller owns (e.g. a POSIX shared-memory frame that several ranks copy into)
extern "C" int bhr_host_register(void* p, size_t bytes) {
    if (!p) return BHR_ERR_INVALID;
    return cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess ? BHR_OK : BHR_ERR_CUDA;
}
extern "C" int bhr_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? BHR_OK : BHR_ERR_CUDA; }

// ---------------------------------------------------------------------------------------------
// FP32 throughput probe (roofline denominator of the integrator)
// ---------------------------------------------------------------------------------------------
namespace {
template <int MODE> __global__ void __launch_bounds__(256) fma_probe(float* out, int iters) {
    float x[8]; float2 v[8];
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x * 1e-3f + j; v[j] = make_float2(x[j], x[j] + 0.5f); }
    const float m = 0.999f, c = 1e-3f; const float2 m2 = make_float2(0.999f, 1.001f), c2 = make_float2(1e-3f, 2e-3f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (MODE == 0) x[j] = fmaf(x[j], m, c);
                else v[j] = __ffma2_rn(v[j], m2, c2);
            }
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += x[j] + v[j].x + v[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int bhr_measure_fp32_peak(int device, int mode, double* tflops) {
    if (!tflops) return BHR_ERR_INVALID;
    BhrDeviceGuard device_guard_(device);
    if (cudaSetDevice(device) != cudaSuccess) return BHR_ERR_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = sms * 8, iters = 20000;
    float* out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * grid * 256) != cudaSuccess) return BHR_ERR_NOMEM;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        if (mode == 0) fma_probe<0><<<grid, 256>>>(out, iters); else fma_probe<1><<<grid, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (e != cudaSuccess) return BHR_ERR_CUDA;
    const double fma = (double)grid * 256 * iters * 64.0 * (mode == 0 ? 1.0 : 2.0);
    *tflops = fma * 2.0 / (best * 1e-3) / 1e12;
    return BHR_OK;
}
