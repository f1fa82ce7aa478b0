/*
 * bhr.h -- C ABI of libbhr.so, the B200 (sm_100a) implementation of hwuu/black-hole-renderer's
 * per-pixel null-geodesic render path.
 *
 * The reference has no FFI of its own: its device code is reached only through the Python class
 * `TaichiRenderer` (render.py:2189-4028), whose methods launch Taichi kernels.  Each entry point
 * below replaces one of those methods / kernels and cites it.  The Python class
 * black_hole_renderer_b200.Renderer (same constructor and method surface as TaichiRenderer)
 * binds these symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success and a non-zero bhr_status otherwise (the
 * message is available from bhr_last_error); no exceptions cross the ABI; host buffers are owned
 * by the caller, device buffers by the context; one CUDA stream per context (bhr_set_stream lets
 * the host supply it); a context is not thread-safe; one context per GPU.  Images are row-major
 * (H, W, 3) -- the reference's (W, H) fields transposed as render() does on return
 * (render.py:3923).  All device arithmetic is float32 (the lens flare uses float64 like the
 * reference's numpy implementation).
 */
#ifndef BHR_H_
#define BHR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bhr_ctx bhr_ctx;

typedef enum {
    BHR_OK = 0,
    BHR_ERR_INVALID = 1,   /* bad argument / size mismatch (the reference raises AssertionError) */
    BHR_ERR_CUDA = 2,      /* CUDA runtime failure                                                */
    BHR_ERR_STATE = 3,     /* call order violated (e.g. generate_background before init)          */
    BHR_ERR_NOMEM = 4
} bhr_status;

/* TaichiRenderer.__init__ arguments, render.py:2199-2208 */
typedef struct {
    int32_t width, height;
    float step_size;            /* h_base                                   */
    float r_max;
    float r_disk_inner, r_disk_outer;
    float disk_tilt_deg;
    int32_t lens_flare;         /* mutable later through bhr_set_lens_flare */
    int32_t anti_alias;         /* 0 = "disabled", 1 = "lod_radius"         */
    float aa_strength;
    float disk_rotation_speed;  /* t_offset = frame * speed (render.py:3897) */
    int32_t device;             /* CUDA device ordinal                      */
} bhr_config;

/* What TaichiRenderer.render uploads per frame (render.py:3880-3892): build_camera's float64
 * results cast to f32, the pixel pitch and r_escape = max(r_max, 2|cam|). */
typedef struct {
    float pos[3], right[3], up[3], forward[3];
    float pixel_w, pixel_h;
    float r_escape;
    float t_offset;
} bhr_camera;

/* bhr_render flags */
#define BHR_SKIP_DIFFERENTIALS 1u   /* render(skip_differentials=True), render.py:3900 */
#define BHR_SKIP_BLOOM 2u           /* render(skip_bloom=True), render.py:3911         */
#define BHR_WANT_AUX 4u             /* also fill the class / step-count buffers        */
#define BHR_SKIP_FLARE 8u           /* render_to_field (render.py:3819-3863) never applies the lens flare */
/* render_to_field's compositing (render.py:3857-3863 after _bloom_kernel's in-place add, 3112-3114):
 * final = clamp(bg + clamp(disk + 0.4 blur, 0, 1) + blur, 0, 1) instead of render()'s clamp(bg + disk + blur) */
#define BHR_FIELD_COMPOSITE 16u
/* bhr_render_rows_stage2: take the frame-wide flare sums from device buffer BHR_BUF_FLARE_SUMS (filled by
 * bhr_flare_sums_device and, with several GPUs, all-reduced in place by the caller) instead of a host array */
#define BHR_FLARE_FROM_DEVICE 32u

/* device buffers that can be inspected / exchanged (bhr_buffer, bhr_download) */
typedef enum {
    BHR_BUF_BG = 0,        /* image_field: planar 3 x (H, W) f32                         */
    BHR_BUF_DISK = 1,      /* disk_layer_field before bloom: planar 3 x (H, W) f32       */
    BHR_BUF_HBLUR = 2,     /* bright_field after the horizontal pass: planar 3 x (H, W)  */
    BHR_BUF_FINAL = 3,     /* (H, W, 3) f32, render()'s return value                     */
    BHR_BUF_FINAL_U8 = 4,  /* (H, W, 3) u8 = trunc(clip(final) * 255), render.py:4463    */
    BHR_BUF_CLASS = 5,     /* (H, W) u8: bits 0-1 termination (0 exhausted, 1 horizon, 2 escaped), bits 2-4 = min(disk hits, 7) */
    BHR_BUF_STEPS = 6,     /* (H, W) i32 RK4 evaluations per ray                         */
    BHR_BUF_DISK_TEX = 7,  /* disk_texture_field: (n_r, n_phi, 4) f32                    */
    BHR_BUF_DISK_MIPS = 8, /* compact pyramid, level l = (n_r>>l, n_phi>>l, 4) f32       */
    BHR_BUF_COMP = 9,      /* _comp_field: (13, n_r, n_phi) f32                          */
    BHR_BUF_BLUR = 10,     /* blur_field (after the vertical pass): planar 3 x (H, W)    */
    BHR_BUF_FLARE_SUMS = 12, /* 3 f64 {sum B, sum x*B, sum y*B} of the last bhr_flare_sums[_device] call       */
    BHR_BUF_DISK_POST = 11 /* disk_layer_field as _bloom_kernel leaves it (render.py:3112-3114): clamp(disk + 0.4 blur, 0, 1),
                              planar 3 x (H, W); formed on demand by bhr_download (bhr_buffer does not expose it) */
} bhr_buffer_id;

/* ---- lifecycle (TaichiRenderer.__init__, render.py:2199-2290) ---- */
int bhr_create(const bhr_config* cfg, bhr_ctx** out);
void bhr_destroy(bhr_ctx* ctx);
const char* bhr_last_error(const bhr_ctx* ctx); /* ctx may be NULL: error of a failed bhr_create */
int bhr_set_stream(bhr_ctx* ctx, void* cuda_stream); /* NULL = the context's own stream */
int bhr_synchronize(bhr_ctx* ctx);
int bhr_set_lens_flare(bhr_ctx* ctx, int enabled);  /* renderer.lens_flare attribute */
int bhr_version(void);
/* tuning knobs: "raymarch_mode" = 0 fast integrator (default; re-associated RK4, packed FFMA2),
 * 2 strict (reference operation order, exactly rounded; ~3x slower); "retrace_band" = eps: rays
 * whose impact parameter is within eps of the critical one (photon-ring rays, chaotic) are traced
 * by the strict integrator (default 0.02, 0 = off); "retrace_min_cross" = n: safety net, rays with
 * >= n disk-plane crossings are re-traced too (default 3, 0 = off); "band_lo_auto" = 1 (default): with the
 * disk's inner edge outside the photon sphere the band is (-0.005, eps) -- rays safely below the
 * critical impact parameter end in the horizon whatever they do near it; "persistent" = 1 (default):
 * one block per SM with work queues, strict and fast rays on disjoint SMs; "pblock_big";
 * "strict_warps" = n: warps per strict block that trace photon-ring batches (default: all; fewer
 * spreads them over more SMs -- lower latency of a small ring tile, idle warps meanwhile);
 * "stage_timing" = 1 (default): record the CUDA events bhr_last_stage_ms reads (five timing events
 * per frame, ~1.5 us of stream time each; video loops switch them off); "band_box" = 1 (default): the band-list kernel scans only the photon ring's bounding box;
 * "sync_bands" / "sync_min_bytes": row bands of synchronous host frames, see bhr_render;
 * "bloom_generic" = 3: use the generic bloom / composite kernels (any frame size) instead of the TMA-fed ones
 * (widths that are a multiple of 4; same arithmetic, bit-identical frames); 1 = generic H pass only, 2 = generic
 * V pass + composite only; "keep_blur" = 1: the fused V pass
 * also stores blur_field (it is otherwise formed on demand by bhr_download(BHR_BUF_BLUR)) */
int bhr_set_option(bhr_ctx* ctx, const char* key, double value);
/* pinned host memory so that frame read-back DMA needs no staging copy */
int bhr_host_alloc(size_t bytes, void** out);
int bhr_host_free(void* p);
/* "dddd:bb:dd.f" of a CUDA device: lets the host find the GPU's NUMA node in sysfs and place its
 * threads / pinned frame buffers there (black_hole_renderer_b200/hostmem.py) */
int bhr_device_pci_bus_id(int device, char* out, int len);
/* page-lock caller-owned memory, e.g. a shared-memory frame every rank copies its rows into */
int bhr_host_register(void* p, size_t bytes);
int bhr_host_unregister(void* p);

/* texture_field.from_numpy (render.py:2232-2233): skybox (h, w, 3) f32 host */
int bhr_upload_skybox(bhr_ctx* ctx, const float* rgb, int h, int w);
/* __init__/update_disk_texture (render.py:2235-2251, 2292-2312): (n_r, n_phi, 4) f32 host;
 * rebuilds the 5-level mip pyramid in generate_disk_mipmaps' summation order. The first call
 * fixes (n_r, n_phi); later calls must match (the reference asserts). */
int bhr_upload_disk_texture(bhr_ctx* ctx, const float* rgba, int n_r, int n_phi);

/* ---- hot path (TaichiRenderer.render, render.py:3865-3923) ---- */
/* Ray march + bloom + composite [+ flare]; if out_f32 / out_u8 are non-NULL the (H, W, 3)
 * result is copied to those HOST buffers (the call then synchronises).  Large host frames
 * (>= option "sync_min_bytes", 16 MB) without flare are finished in row bands -- the rows of the
 * photon ring first, then "sync_bands" pieces above and below -- so that the copy of one band
 * overlaps the ray march of the next; the result is bit-identical to the one-shot frame, and
 * bhr_last_stage_ms then describes the last band. */
int bhr_render(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8);
/* The same without the final synchronisation, for pipelined video loops: the frame is enqueued
 * (ray march ... composite, copies into the PINNED host buffers) and completion event `slot`
 * (0..31) is recorded; bhr_wait_frame(slot) blocks until that frame is in host memory.  The host
 * can prepare the next frame (lifecycle tick, entity packing) while the device works. */
int bhr_render_async(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8, int slot);
int bhr_wait_frame(bhr_ctx* ctx, int slot);

/* ---- frame files: the PNG's deflate stream produced on the device ----
 * Replaces the host-side PIL save of every video frame (render.py:4462-4467, render_video) by three small
 * kernels after the composite: Sub-filtered scanlines, byte runs as (length, distance 1) matches, a static
 * Huffman code (format + CPU twin: black_hole_renderer_b200/png_codec.py).  What reaches the host is the
 * complete zlib stream of the IDAT chunk; the host adds the chunk framing and CRC-32 and writes the file.
 * bhr_png_setup uploads the code tables once per context (png_codec.StaticCode.device_tables()). */
int bhr_png_setup(bhr_ctx* ctx, const uint32_t* lit_bits, const uint32_t* lit_nbits, const uint32_t* match_bits,
                  const uint32_t* match_nbits, const uint8_t* header, uint32_t header_nbits, uint32_t eob_bits,
                  uint32_t eob_nbits);
/* upper bound of a frame's stream length in bytes (size of the host buffers below, without the 8-byte prefix) */
int bhr_png_capacity(bhr_ctx* ctx, size_t* stream_bytes);
/* bhr_render_async whose result is the stream: `host` (pinned, >= 8 + copy_bytes) receives
 * {uint32 stream_bytes, uint32 adler32} and the first copy_bytes of the stream.  A caller may pass a
 * copy_bytes smaller than the capacity (e.g. the previous frame's length + margin); when stream_bytes turns
 * out larger it reads the remainder with bhr_png_fetch before the slot is reused. */
int bhr_render_async_png(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, void* host, size_t copy_bytes, int slot);
int bhr_png_fetch(bhr_ctx* ctx, int slot, size_t offset, size_t bytes, void* host);
/* stream of the frame currently in BHR_BUF_FINAL_U8, synchronously */
int bhr_png_encode_current(bhr_ctx* ctx, void* host, size_t host_bytes, uint32_t* stream_bytes);

/* Row-tile variants used when one frame is split over several GPUs.  Stage 1 ray
 * marches rows [row0, row1) and runs the horizontal bloom pass on them; the caller then exchanges
 * the radius-row halos of BHR_BUF_HBLUR between neighbours (NCCL) and, for the flare, all-reduces
 * the three brightness sums; stage 2 runs the vertical pass + composite (+ flare) on the rows. */
int bhr_render_rows_stage1(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, int row0, int row1);
int bhr_render_rows_stage2(bhr_ctx* ctx, uint32_t flags, int row0, int row1,
                           const double* flare_sums /* NULL or {sum B, sum x*B, sum y*B} */);
int bhr_flare_sums(bhr_ctx* ctx, int row0, int row1, double out[3]);
/* the same reduction, enqueued only: the sums stay in BHR_BUF_FLARE_SUMS (no host synchronisation) */
int bhr_flare_sums_device(bhr_ctx* ctx, int row0, int row1);
int bhr_bloom_radius(const bhr_ctx* ctx);

/* ---- the same split with peer memory instead of NCCL (one process per GPU on one node) ----
 * bhr_peer_export returns CUDA IPC handles of {H-blurred layer, final f32, final u8, sync block};
 * the caller exchanges them between the ranks (any transport; dist.py uses all_gather_object)
 * and hands all of them (world x 4, rank-major) to bhr_peer_attach.  bhr_render_tiled_peer then
 * renders this rank's row tile: the vertical bloom pass loads its halo rows straight from the
 * neighbours' HBM over NVLink, the flare sums are exchanged and reduced by kernels, the finished
 * rows are stored straight into rank 0's final buffers, and rank 0 copies the frame to the host.
 * Distributed egress: when EVERY rank passes the same host buffer (shared memory mapped in all
 * processes, page-locked with bhr_host_register) each rank copies its own rows to the host over
 * its own PCIe link instead, and rank 0 returns when all of them have landed.
 * No collective library, no host synchronisation between the stages; every rank must call it once
 * per frame, all with or all (but rank 0) without host buffers.
 * Failure behaviour: no wait spins for ever.  A rank that cannot finish a frame poisons it (its
 * peers drain and return BHR_ERR_STATE); a flag that does not arrive within option
 * "peer_timeout_ms" (default 20 000: a rank died or the ranks called an unequal number of times)
 * ends the wait, and the call that next synchronises returns BHR_ERR_STATE. */
typedef struct { unsigned char bytes[64]; } bhr_ipc_handle;
int bhr_peer_export(bhr_ctx* ctx, bhr_ipc_handle out[4]);
int bhr_peer_attach(bhr_ctx* ctx, int rank, int world, const bhr_ipc_handle* all);
int bhr_render_tiled_peer(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8);
int bhr_peer_set_distributed_egress(bhr_ctx* ctx, int enabled);   /* all ranks, before the first frame */
/* Pipelined row-tiled frames (distributed egress only): the call returns once frame s is enqueued; the egress copies
 * and their flags run on the context's copy stream, so the rows of frame s leave over PCIe while frame s + 1 is ray
 * marched.  Consecutive calls must alternate between TWO host frames.  bhr_peer_wait_frame(ctx, back) blocks until
 * frame (latest - back) is complete in its host frame (rank 0: every rank's rows; other ranks: their own rows) and
 * reports a failed / timed-out peer.  Typical loop: async(s); wait_frame(1) -> frame s - 1; ...; wait_frame(0). */
int bhr_render_tiled_peer_async(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8);
int bhr_peer_wait_frame(bhr_ctx* ctx, int back);
/* measurement hook: GB/s at which this rank reads rank `peer`'s H-blurred layer through the peer mapping (the halo
 * pull's access pattern); above PCIe's 64 GB/s the mapping is NVLink */
int bhr_peer_probe_read(bhr_ctx* ctx, int peer, int reps, double* gbs);
/* Tile boundaries (world + 1 ints, bounds[0] = 0, bounds[world] = H): rank r renders rows
 * [bounds[r], bounds[r + 1]).  Equal heights by default; every rank must install the same bounds
 * before the same frame.  Rows through the hole and the disk cost more than sky rows, so a caller
 * balances the tiles by cost: bhr_row_costs returns, for the rows of the last frame rendered with
 * BHR_WANT_AUX, the RK4 evaluations per row. */
int bhr_peer_set_tiles(bhr_ctx* ctx, const int* bounds);
int bhr_row_costs(bhr_ctx* ctx, int row0, int row1, uint64_t* out /* row1 - row0 */);
int bhr_peer_detach(bhr_ctx* ctx);

/* ---- device buffers ---- */
int bhr_buffer(bhr_ctx* ctx, int id, void** dev_ptr, size_t* bytes);
int bhr_download(bhr_ctx* ctx, int id, void* host, size_t bytes);
/* total RK4 evaluations of the last ray march (sum over pixels); feeds the flop count */
int bhr_last_total_steps(bhr_ctx* ctx, uint64_t* out);
/* kernels of this library launched by the context since bhr_create (render path, texture pipeline,
 * peer flags; the statistics / test hooks are not counted) -- what bench.py reports as gpu_launches */
int bhr_launch_count(bhr_ctx* ctx, uint64_t* out);
/* Instrumentation (option "timeline" = 1): {start, strict role done, end} of every block of the last persistent
 * ray-march launch in nanoseconds of the GPU's global timer, 3 x min(max_blocks, SMs) values.  Returns the number of
 * blocks written (0: nothing recorded / failure) -- a count, not a bhr_status. */
int bhr_last_raymarch_timeline(bhr_ctx* ctx, uint64_t* out, int max_blocks);
/* number of rays the last ray march re-traced with the exactly-rounded integrator */
int bhr_last_retrace_count(bhr_ctx* ctx, uint32_t* out);
/* device time (ms, CUDA events) of the stages of the last bhr_render: {ray march, bloom H,
 * bloom V + composite, flare, total}; valid after bhr_synchronize */
int bhr_last_stage_ms(bhr_ctx* ctx, float out[5]);

/* ---- disk-texture pipeline (render.py:3491-3767) ---- */
/* init_background_layer (render.py:3491-3547): allocates comp (13, n_r, n_phi), uploads edge /
 * omega rows, seeds the initial stats.  az_freq / az_shear are the two numbers the reference
 * draws from default_rng(seed) (render.py:3510-3511). */
int bhr_init_background(bhr_ctx* ctx, int n_r, int n_phi, int az_freq, float az_shear,
                        const float* edge, const float* omega_rows);
/* generate_background (render.py:3549-3562) -> _generate_background_kernel (3332-3451) */
int bhr_generate_background(bhr_ctx* ctx, float t);

/* One entity of the lifecycle system as the accumulate kernel consumes it
 * (accumulate_entity_layer, render.py:3564-3653).  kind 0 = filament (Gaussian blob sheared by
 * differential rotation), 1 = hotspot, 2 = rt_spike (analytic profiles of
 * _spawn_single_hotspot / _spawn_single_rt_spike, render.py:1725-1866, rolled per row);
 * kind 3 = hotspot, 4 = rt_spike given as the reference's own tabulated float32 profiles
 * (EntityInstance.phi_density / phi_temp, (rows, n_phi) each, uploaded with bhr_upload_entity_tables):
 * p[0] = offset (in floats) of the entity's density rows in that buffer, its temperature rows follow;
 * scale = fade alpha; the kernel does numpy's float32 `+= roll(row, -shift) * alpha`. */
typedef struct {
    int32_t kind;
    int32_t row_begin, row_end;   /* affected rows [row_begin, row_end)        */
    double age;                   /* now - birth_time                          */
    double scale;                 /* filament: birth_alpha*cool_factor*s0/sigma_t ; others: fade alpha */
    double p[8];                  /* kind-specific parameters, see texture.cu  */
} bhr_entity;
int bhr_accumulate_entities(bhr_ctx* ctx, const bhr_entity* entities, int n);
int bhr_upload_entity_tables(bhr_ctx* ctx, const float* data, size_t n_floats);
/* recompute_interactive_stats (render.py:3655-3712) on the device.  The reference runs
 * np.percentile / np.quantile on the host; here the device returns the exact order statistics next
 * to each quantile's virtual index and the caller applies numpy's interpolation to them:
 *   bhr_stats_prepare: density / structure planes (numpy's f32 operation order); returns the texel
 *                      count and the number of texels with structure > 0;
 *   bhr_stats_select:  out = {density[rank_d], density[rank_d + 1], struct+[rank_s], struct+[rank_s + 1]}
 *                      in sorted order (struct+ = the positive structure values; the upper neighbour
 *                      is clipped to the last element);
 *   bhr_stats_rows:    out (n_r, 4) = per row {max, sorted[lo], sorted[hi]} of
 *                      clip(structure / denom * 0.8, 0, 1.2) and the row max of the base temperature.
 * The results are pushed back with bhr_set_stats. */
int bhr_stats_prepare(bhr_ctx* ctx, int enable_rt, uint64_t* n_total, uint64_t* n_positive);
int bhr_stats_select(bhr_ctx* ctx, uint64_t rank_density, uint64_t rank_struct, float out[4]);
int bhr_stats_rows(bhr_ctx* ctx, float denom, int lo, int hi, float* out /* (n_r, 4) */);
int bhr_set_stats(bhr_ctx* ctx, float density_p98, float struct_scale, const float* row_stats /* (n_r, 2) */);
int bhr_upload_comp(bhr_ctx* ctx, const float* comp /* (13, n_r, n_phi) */);
/* compose_interactive_texture (render.py:3714-3767): compose kernel + mip kernels */
int bhr_compose_texture(bhr_ctx* ctx, float t_offset, int enable_rt, float color_temp);
/* eval_noise (render.py:3769-3790): mode 0 simplex, 1 fbm; coords (n, 3) host -> out (n) host.  Mode 2 (test hook):
 * simplex through the background kernel's packed two-points-per-thread code (csrc/background.cu) */
int bhr_eval_noise(bhr_ctx* ctx, const float* coords, int n, int mode, int octaves,
                   float persistence, float lacunarity, float* out);

/* ---- measurement helpers ---- */
/* FP32 FMA throughput microbenchmark (TFLOP/s): mode 0 scalar FFMA, 1 packed FFMA2 */
int bhr_measure_fp32_peak(int device, int mode, double* tflops);

/* self-test of the strict integrator's three-instruction x / 6: compares it with the IEEE division
 * on all 2^32 float bit patterns and returns the number of mismatches among the inputs whose
 * quotient is a normal number or zero (expected 0), and among all inputs */
int bhr_selftest_div6(int device, unsigned long long* mismatches_normal, unsigned long long* mismatches_all);

#ifdef __cplusplus
}
#endif
#endif /* BHR_H_ */
