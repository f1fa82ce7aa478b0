// peer.cu -- one frame split into row tiles over the GPUs of a node, with every exchange done by
// the kernels themselves through peer memory (CUDA IPC mappings over NVLink / NVSwitch); no NCCL
// and no host synchronisation on the data path.
//
// Every rank (one process per GPU) owns full-size layer buffers and fills its own rows.  Per frame:
//   stage 1   ray march + horizontal bloom pass on my rows                       (own HBM)
//   publish   my flare partial sums and "H pass of frame s done" -> every rank's PeerSync block
//   wait      until every rank has published frame s
//   stage 2   vertical bloom pass: halo rows are loaded straight from the neighbours' H-blurred
//             buffers (bloom_v_kernel's row_src table); flare parameters reduced on the device;
//             composite: finished rows are stored straight into rank 0's final buffers
//   publish   "my V pass of frame s is done" and "my rows of frame s are in place" -> every rank
//   rank 0    waits for all tiles, copies the frame to the host, publishes "consumed s"
//   (distributed egress: each rank copies its own rows into a shared, page-locked host frame over
//    its own PCIe link instead of storing them into rank 0's HBM; rank 0 only waits)
// Back-pressure: a rank starts frame s + 1 (overwrites its H-blurred rows) only when every rank's V
// pass of frame s is done, and touches rank 0's final buffers / the host frame only when rank 0's
// caller has come back for frame s + 1 (i.e. is done with frame s).
// Pipelined frames (bhr_render_tiled_peer_async, distributed egress only): the egress copies and
// their flags run on the copy stream, so the rows of frame s leave over PCIe while frame s + 1 is
// ray marched; frames alternate between TWO host frames, the caller collects frame s - 1
// (bhr_peer_wait_frame) after it has enqueued frame s, and "consumed" lags by two frames.
// Flags live in the memory of the rank that waits on them, so spinning is local.
#include "common.cuh"

struct PeerSync {
    unsigned h_ready[16];        // [r] = s: rank r's H pass (and flare sums) of frame s are complete
    unsigned tile_done[16];      // [r] = s: rank r's rows of frame s are stored in rank 0's final buffers
    unsigned consumed;           // = s: rank 0 has copied frame s out of its final buffers
    unsigned poison;             // = s: some rank failed while frame s was in flight; waits for frames <= s give up
    unsigned v_done[16];         // [r] = s: rank r's V pass of frame s no longer reads anybody's H-blurred rows
    unsigned pad[14];
    double flare_part[16][3];
};
// A wait never spins for ever: if a flag does not arrive within the budget (a rank died, raised, or
// the ranks called an unequal number of times) or the frame is poisoned, the waiting thread stores
// a code into a host-mapped error word and lets its stream run on (the frame it produces is then
// garbage); the host finds the code at its next synchronisation / call and returns BHR_ERR_STATE.
enum : int { PEER_ERR_TIMEOUT = 1, PEER_ERR_POISONED = 2 };

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// thread r publishes to rank r: field 0 = h_ready (+ flare partial sums), 1 = tile_done, 2 = consumed, 4 = v_done, 3 = poison
__global__ void peer_publish_kernel(PeerSync* const* __restrict__ peers, int rank, int world, const double* __restrict__ sums,
                                    unsigned serial, int field) {
    const int r = threadIdx.x;
    if (r >= world) return;
    PeerSync* p = peers[r];
    if (field == 0) {
        if (sums) { p->flare_part[rank][0] = sums[0]; p->flare_part[rank][1] = sums[1]; p->flare_part[rank][2] = sums[2]; }
        __threadfence_system();
        st_release_sys(&p->h_ready[rank], serial);
    } else if (field == 1) {
        __threadfence_system();
        st_release_sys(&p->tile_done[rank], serial);
    } else if (field == 2) {
        __threadfence_system();
        st_release_sys(&p->consumed, serial);
    } else if (field == 4) {
        __threadfence_system();
        st_release_sys(&p->v_done[rank], serial);
    } else {
        // poison frame `serial`: this rank cannot finish it.  Its flags are published as well so that
        // nobody waits for it on this frame.
        st_release_sys(&p->poison, serial);
        st_release_sys(&p->h_ready[rank], serial);
        st_release_sys(&p->v_done[rank], serial);
        st_release_sys(&p->tile_done[rank], serial);
        if (rank == 0) st_release_sys(&p->consumed, serial);
    }
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// thread i waits until flags[i] >= serial (flags are in this GPU's own memory), the frame is
// poisoned, or the budget runs out
__global__ void peer_wait_kernel(const unsigned* __restrict__ flags, int n, unsigned serial,
                                 const unsigned* __restrict__ poison, unsigned long long budget_ns,
                                 volatile int* __restrict__ host_err) {
    const int i = threadIdx.x;
    if (i < n) {
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(flags + i) - serial) < 0) {
            if ((int)(ld_acquire_sys(poison) - serial) >= 0) break;
            if (global_ns() - t0 > budget_ns) { *host_err = PEER_ERR_TIMEOUT; break; }
            __nanosleep(200);
        }
        // (a poisoning rank stores `poison` before its flags, so a wait that its flags released sees it too)
        if ((int)(ld_acquire_sys(poison) - serial) >= 0 && serial != 0) *host_err = PEER_ERR_POISONED;
    }
    __threadfence_system();
}

}  // namespace

static void tile_rows(int H, int world, int rank, int* r0, int* r1) {   // == dist.tile_rows
    const int base = H / world, extra = H % world;
    *r0 = rank * base + (rank < extra ? rank : extra);
    *r1 = *r0 + base + (rank < extra ? 1 : 0);
}

static int launch_wait(bhr_ctx* ctx, const unsigned* flags, int n, unsigned serial, cudaStream_t stream = nullptr) {
    const unsigned long long budget = (unsigned long long)(ctx->peer_timeout_ms > 0 ? ctx->peer_timeout_ms : 20000.0) * 1000000ull;
    peer_wait_kernel<<<1, 32, 0, stream ? stream : ctx->stream>>>(flags, n, serial, &ctx->peer_sync_own->poison, budget, ctx->peer_host_err);
    ++ctx->launches;
    BHR_CUDA(ctx, cudaGetLastError());
    return BHR_OK;
}

// which buffer holds row y of the H-blurred layer, for the current tile bounds (enqueued on the
// context's stream, i.e. behind the V pass of the previous frame that still reads the old table)
static int upload_row_sources(bhr_ctx* ctx) {
    const float** rows = (const float**)ctx->peer_rows_host;
    for (int r = 0; r < ctx->peer_world; ++r)
        for (int y = ctx->peer_bounds[r]; y < ctx->peer_bounds[r + 1]; ++y) rows[y] = ctx->peer_hblur[r];
    BHR_CUDA(ctx, cudaMemcpyAsync(ctx->d_row_src, rows, sizeof(float*) * ctx->H, cudaMemcpyHostToDevice, ctx->stream));
    return BHR_OK;
}

extern "C" int bhr_peer_export(bhr_ctx* ctx, bhr_ipc_handle out[4]) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !out) return BHR_ERR_INVALID;
    if (!ctx->peer_sync_own) {
        BHR_CUDA(ctx, cudaMalloc(&ctx->peer_sync_own, sizeof(PeerSync)));
        BHR_CUDA(ctx, cudaMemset(ctx->peer_sync_own, 0, sizeof(PeerSync)));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(bhr_ipc_handle), "IPC handle size");
    void* bufs[4] = {ctx->hblur, ctx->final_f32, ctx->final_u8, ctx->peer_sync_own};
    for (int k = 0; k < 4; ++k) BHR_CUDA(ctx, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)&out[k], bufs[k]));
    return BHR_OK;
}

extern "C" int bhr_peer_attach(bhr_ctx* ctx, int rank, int world, const bhr_ipc_handle* all) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !all || world < 1 || world > 16 || rank < 0 || rank >= world) return BHR_ERR_INVALID;
    if (!ctx->peer_sync_own) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_peer_export must run first");
    if (ctx->peer_world) BHR_FAIL(ctx, BHR_ERR_STATE, "peers are already attached");
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            ctx->peer_hblur[r] = ctx->hblur; ctx->peer_final_f32[r] = ctx->final_f32;
            ctx->peer_final_u8[r] = ctx->final_u8; ctx->peer_sync[r] = ctx->peer_sync_own;
            continue;
        }
        void* p[4];
        for (int k = 0; k < 4; ++k)
            BHR_CUDA(ctx, cudaIpcOpenMemHandle(&p[k], *(const cudaIpcMemHandle_t*)&all[4 * r + k], cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_hblur[r] = (float*)p[0]; ctx->peer_final_f32[r] = (float*)p[1];
        ctx->peer_final_u8[r] = (uint8_t*)p[2]; ctx->peer_sync[r] = (PeerSync*)p[3];
    }
    ctx->peer_rank = rank; ctx->peer_world = world; ctx->peer_serial = 0;
    for (int r = 0; r < world; ++r) {
        int r0, r1;
        tile_rows(ctx->H, world, r, &r0, &r1);
        ctx->peer_bounds[r] = r0; ctx->peer_bounds[r + 1] = r1;
    }
    // (pinned: the table upload of bhr_peer_set_tiles must not block behind frames in flight)
    BHR_CUDA(ctx, cudaMallocHost(&ctx->peer_rows_host, sizeof(float*) * ctx->H));
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_row_src, sizeof(float*) * ctx->H));
    {
        int rc = upload_row_sources(ctx);
        if (rc) return rc;
        BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (!ctx->peer_host_err) {
        BHR_CUDA(ctx, cudaHostAlloc((void**)&ctx->peer_host_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *ctx->peer_host_err = 0;
    }
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_peer_sync, sizeof(PeerSync*) * 16));
    BHR_CUDA(ctx, cudaMemcpy(ctx->d_peer_sync, ctx->peer_sync, sizeof(PeerSync*) * 16, cudaMemcpyHostToDevice));
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_flare_params, bhr_flare_params_size()));
    return BHR_OK;
}

// Tile boundaries: rank r renders rows [bounds[r], bounds[r + 1]).  Every rank must install the
// SAME bounds before the same frame (dist.balance_tiles derives them on every rank from all-gathered
// per-row costs); equal-height tiles until then.
extern "C" int bhr_peer_set_tiles(bhr_ctx* ctx, const int* bounds) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !bounds) return BHR_ERR_INVALID;
    if (!ctx->peer_world) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_peer_attach has not run");
    if (bounds[0] != 0 || bounds[ctx->peer_world] != ctx->H) BHR_FAIL(ctx, BHR_ERR_INVALID, "tile bounds must run from 0 to the frame height");
    for (int r = 0; r < ctx->peer_world; ++r)
        if (bounds[r + 1] < bounds[r]) BHR_FAIL(ctx, BHR_ERR_INVALID, "tile bounds must not decrease");
    // the pinned staging table may still be read by the previous upload: wait for this stream
    BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r <= ctx->peer_world; ++r) ctx->peer_bounds[r] = bounds[r];
    return upload_row_sources(ctx);
}

static int peer_check_error(bhr_ctx* ctx) {
    const int code = ctx->peer_host_err ? *(volatile int*)ctx->peer_host_err : 0;
    if (code == PEER_ERR_TIMEOUT) BHR_FAIL(ctx, BHR_ERR_STATE, "peer wait timed out: a rank failed, exited or fell out of step (option \"peer_timeout_ms\")");
    if (code == PEER_ERR_POISONED) BHR_FAIL(ctx, BHR_ERR_STATE, "a peer rank failed while the frame was in flight");
    return BHR_OK;
}

static int wait_consumed(bhr_ctx* ctx) {     // before the composite stores into rank 0's final buffers
    if (ctx->peer_rank != 0) return launch_wait(ctx, &ctx->peer_sync_own->consumed, 1, ctx->peer_serial - 1);
    return BHR_OK;
}

static int render_tiled_peer_body(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8, unsigned s,
                                  bool pipelined);

static int render_tiled_peer_entry(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8, bool pipelined) {
    if (!ctx || !cam) return BHR_ERR_INVALID;
    if (!ctx->peer_world) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_peer_attach has not run");
    int rc = peer_check_error(ctx);                  // a wait of an earlier frame gave up: the ranks are out of step
    if (rc) return rc;
    if (pipelined && !((out_f32 || out_u8) && ctx->peer_distributed))
        BHR_FAIL(ctx, BHR_ERR_STATE, "pipelined tiled frames need a host frame and distributed egress (bhr_peer_set_distributed_egress)");
    const unsigned s = ++ctx->peer_serial;
    rc = render_tiled_peer_body(ctx, cam, flags, out_f32, out_u8, s, pipelined);
    if (rc) {
        // this rank cannot finish frame s: poison it so that every peer's waits drain instead of spinning
        // (their frame is garbage and they report BHR_ERR_STATE); keep our own error message
        peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, ctx->peer_rank, ctx->peer_world, nullptr, s, 3);
        ++ctx->launches;
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    return BHR_OK;
}

extern "C" int bhr_render_tiled_peer(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8) {
    BhrDeviceGuard device_guard_(ctx);
    return render_tiled_peer_entry(ctx, cam, flags, out_f32, out_u8, false);
}

// The same frame without the final synchronisation on rank 0 and with the egress on the copy stream.  Consecutive
// calls must alternate between two host frames; bhr_peer_wait_frame(ctx, 1) after the call for frame s returns when
// frame s - 1 is complete in its host frame (rank 0: every rank's rows; other ranks: their own rows).  Do not mix
// with bhr_render_tiled_peer on the same context without a bhr_peer_wait_frame(ctx, 0) in between.
extern "C" int bhr_render_tiled_peer_async(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8) {
    BhrDeviceGuard device_guard_(ctx);
    return render_tiled_peer_entry(ctx, cam, flags, out_f32, out_u8, true);
}

extern "C" int bhr_peer_wait_frame(bhr_ctx* ctx, int back) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || back < 0 || back > 2) return BHR_ERR_INVALID;
    if (!ctx->peer_world) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_peer_attach has not run");
    if ((unsigned)back >= ctx->peer_serial) return BHR_OK;                 // no such frame yet
    const unsigned s = ctx->peer_serial - (unsigned)back;
    if (!ctx->tiled_ev[s & 3]) BHR_FAIL(ctx, BHR_ERR_STATE, "frame %u was not enqueued with bhr_render_tiled_peer_async", s);
    BHR_CUDA(ctx, cudaEventSynchronize(ctx->tiled_ev[s & 3]));
    return peer_check_error(ctx);
}

static int render_tiled_peer_body(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8, unsigned s,
                                  bool pipelined) {
    const int rank = ctx->peer_rank, world = ctx->peer_world;
    PeerSync* mine = ctx->peer_sync_own;
    // distributed egress: every rank copies its own rows into the (shared, page-locked) host frame
    const bool own_egress = rank != 0 && (out_f32 || out_u8);
    const bool host_out = out_f32 || out_u8;
    const int row0 = ctx->peer_bounds[rank], row1 = ctx->peer_bounds[rank + 1];
    const unsigned lag = pipelined ? 2u : 1u;        // the host frame of frame s last held frame s - lag
    int rc;
    cudaStream_t es = ctx->stream;                   // where the egress copies and their flags run
    if (pipelined) {
        if (!ctx->copy_stream) BHR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        if (!ctx->tiled_copy_done) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tiled_copy_done, cudaEventDisableTiming));
        if (!ctx->frame_done) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->frame_done, cudaEventDisableTiming));
        for (int k = 0; k < 4; ++k)
            if (!ctx->tiled_ev[k]) BHR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tiled_ev[k], cudaEventDisableTiming));
        es = ctx->copy_stream;
        // the host runs at most two frames ahead of this rank's egress
        if (s > 2) BHR_CUDA(ctx, cudaEventSynchronize(ctx->tiled_ev[(s - 2) & 3]));
    }
    if (rank == 0 && s >= lag) {
        // the caller is back for another frame: it is done with frame s - lag (host frame / final buffers)
        peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, nullptr, s - lag, 2);
        ++ctx->launches;
    }
    // every rank's V pass of frame s - 1 is done (it no longer reads my H-blurred rows)
    rc = launch_wait(ctx, mine->v_done, world, s - 1);
    if (rc) return rc;
    rc = bhr_render_rows_stage1(ctx, cam, flags, row0, row1);
    if (rc) return rc;
    const bool flare = ctx->cfg.lens_flare && !(flags & BHR_SKIP_FLARE);
    if (flare) {
        rc = bhr_launch_flare_sums(ctx, row0, row1);
        if (rc) return rc;
    }
    peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, flare ? ctx->d_flare_sums : nullptr, s, 0);
    ++ctx->launches;
    rc = launch_wait(ctx, mine->h_ready, world, s);
    if (rc) return rc;
    if (flare) {
        rc = bhr_launch_flare_params(ctx, &mine->flare_part[0][0], world, ctx->d_flare_params);
        if (rc) return rc;
    }
    // the egress copy of the previous pipelined frame has read my final buffers
    if (ctx->tiled_copy_pending) {
        BHR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->tiled_copy_done, 0));
        ctx->tiled_copy_pending = 0;
    }
    bhr_post_peer peer;
    peer.row_src = ctx->d_row_src;
    peer.final_f32 = own_egress ? nullptr : ctx->peer_final_f32[0];
    peer.final_u8 = own_egress ? nullptr : ctx->peer_final_u8[0];
    peer.flare_params = flare ? ctx->d_flare_params : nullptr;
    peer.before_composite = own_egress ? nullptr : wait_consumed;
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    rc = bhr_launch_bloom_v_composite_ex(ctx, flags, row0, row1, nullptr, &peer);
    if (rc) return rc;
    if (ctx->stage_timing) BHR_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    ctx->ev_valid = ctx->stage_timing;
    peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, nullptr, s, 4);
    ++ctx->launches;
    if (pipelined) {
        BHR_CUDA(ctx, cudaEventRecord(ctx->frame_done, ctx->stream));
        BHR_CUDA(ctx, cudaStreamWaitEvent(es, ctx->frame_done, 0));
    }
    const size_t row3 = (size_t)ctx->W * 3;
    if (own_egress) {
        // the host frame still holds frame s - lag until rank 0's caller has come back for more
        if (s >= lag) {
            rc = launch_wait(ctx, &mine->consumed, 1, s - lag, es);
            if (rc) return rc;
        }
        if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32 + row0 * row3, ctx->final_f32 + row0 * row3,
                                                   (row1 - row0) * row3 * sizeof(float), cudaMemcpyDeviceToHost, es));
        if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8 + row0 * row3, ctx->final_u8 + row0 * row3, (row1 - row0) * row3,
                                                  cudaMemcpyDeviceToHost, es));
    }
    if (rank != 0) {
        peer_publish_kernel<<<1, 32, 0, es>>>(ctx->d_peer_sync, rank, world, nullptr, s, 1);
        ++ctx->launches;
    }
    BHR_CUDA(ctx, cudaGetLastError());
    if (rank == 0) {
        const bool distributed = host_out && ctx->peer_distributed;
        if (distributed) {
            // my own rows go out right away; the other ranks' rows arrive over their own links
            if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32 + row0 * row3, ctx->final_f32 + row0 * row3,
                                                       (row1 - row0) * row3 * sizeof(float), cudaMemcpyDeviceToHost, es));
            if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8 + row0 * row3, ctx->final_u8 + row0 * row3, (row1 - row0) * row3,
                                                      cudaMemcpyDeviceToHost, es));
            peer_publish_kernel<<<1, 32, 0, es>>>(ctx->d_peer_sync, rank, world, nullptr, s, 1);
            ++ctx->launches;
            if (pipelined) { BHR_CUDA(ctx, cudaEventRecord(ctx->tiled_copy_done, es)); ctx->tiled_copy_pending = 1; }
            rc = launch_wait(ctx, mine->tile_done, world, s, es);
            if (rc) return rc;
        } else {
            peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, nullptr, s, 1);
            ++ctx->launches;
            rc = launch_wait(ctx, mine->tile_done, world, s);
            if (rc) return rc;
            const size_t n3 = (size_t)ctx->H * row3;
            if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32, ctx->final_f32, n3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8, ctx->final_u8, n3, cudaMemcpyDeviceToHost, ctx->stream));
        }
        BHR_CUDA(ctx, cudaGetLastError());
        if (pipelined) {
            BHR_CUDA(ctx, cudaEventRecord(ctx->tiled_ev[s & 3], es));
            return BHR_OK;
        }
        if (host_out) {
            BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return peer_check_error(ctx);
        }
    } else if (pipelined) {
        BHR_CUDA(ctx, cudaEventRecord(ctx->tiled_copy_done, es));
        ctx->tiled_copy_pending = 1;
        BHR_CUDA(ctx, cudaEventRecord(ctx->tiled_ev[s & 3], es));
    }
    return BHR_OK;
}

// Measurement hook: the H-blurred layer of `peer` (3 x H x W floats in that GPU's HBM, mapped with CUDA IPC) is read
// `reps` times by a plain 16-byte-load kernel -- the access pattern of the halo pull -- into this rank's blur
// scratch layer; GB/s by CUDA events.  A rate above what PCIe can carry (64 GB/s) shows the mapping goes over NVLink.
__global__ void __launch_bounds__(256) peer_probe_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

extern "C" int bhr_peer_probe_read(bhr_ctx* ctx, int peer, int reps, double* gbs) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx || !gbs || reps < 1) return BHR_ERR_INVALID;
    if (peer < 0 || peer >= ctx->peer_world || !ctx->peer_hblur[peer]) BHR_FAIL(ctx, BHR_ERR_STATE, "rank %d is not attached", peer);
    const size_t n4 = (size_t)ctx->W * ctx->H * 3 / 4;
    cudaEvent_t e0, e1;
    BHR_CUDA(ctx, cudaEventCreate(&e0));
    BHR_CUDA(ctx, cudaEventCreate(&e1));
    const int grid = ctx->num_sms * 8;
    peer_probe_kernel<<<grid, 256, 0, ctx->stream>>>((float4*)ctx->blur, (const float4*)ctx->peer_hblur[peer], n4);
    BHR_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    for (int k = 0; k < reps; ++k) peer_probe_kernel<<<grid, 256, 0, ctx->stream>>>((float4*)ctx->blur, (const float4*)ctx->peer_hblur[peer], n4);
    BHR_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    BHR_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.0f;
    BHR_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->launches += reps + 1;
    *gbs = (double)n4 * 16.0 * reps / (ms * 1e-3) / 1e9;
    return BHR_OK;
}

extern "C" int bhr_peer_set_distributed_egress(bhr_ctx* ctx, int enabled) {
    if (!ctx) return BHR_ERR_INVALID;
    ctx->peer_distributed = enabled ? 1 : 0;
    return BHR_OK;
}

extern "C" int bhr_peer_detach(bhr_ctx* ctx) {
    BhrDeviceGuard device_guard_(ctx);
    if (!ctx) return BHR_ERR_INVALID;
    if (!ctx->peer_world) return BHR_OK;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int r = 0; r < ctx->peer_world; ++r) {
        if (r == ctx->peer_rank) continue;
        cudaIpcCloseMemHandle(ctx->peer_hblur[r]); cudaIpcCloseMemHandle(ctx->peer_final_f32[r]);
        cudaIpcCloseMemHandle(ctx->peer_final_u8[r]); cudaIpcCloseMemHandle(ctx->peer_sync[r]);
    }
    cudaFree(ctx->d_row_src); cudaFree(ctx->d_peer_sync); cudaFree(ctx->d_flare_params);
    if (ctx->peer_rows_host) cudaFreeHost(ctx->peer_rows_host);
    if (ctx->peer_host_err) cudaFreeHost((void*)ctx->peer_host_err);
    ctx->d_row_src = nullptr; ctx->d_peer_sync = nullptr; ctx->d_flare_params = nullptr;
    ctx->peer_rows_host = nullptr; ctx->peer_host_err = nullptr;
    ctx->peer_world = 0;
    return BHR_OK;
}
