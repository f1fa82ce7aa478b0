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
//   publish   "my rows of frame s are in place" -> every rank
//   rank 0    waits for all tiles, copies the frame to the host, publishes "consumed s"
//   (distributed egress: each rank copies its own rows into a shared, page-locked host frame over
//    its own PCIe link instead of storing them into rank 0's HBM; rank 0 only waits)
// Back-pressure: a rank starts frame s + 1 (overwrites its H-blurred rows) only when every rank has
// finished frame s, and touches rank 0's final buffers / the host frame only when rank 0's caller
// has come back for frame s + 1 (i.e. is done with frame s).
// Flags live in the memory of the rank that waits on them, so spinning is local.
#include "common.cuh"

struct PeerSync {
    unsigned h_ready[16];        // [r] = s: rank r's H pass (and flare sums) of frame s are complete
    unsigned tile_done[16];      // [r] = s: rank r's rows of frame s are stored in rank 0's final buffers
    unsigned consumed;           // = s: rank 0 has copied frame s out of its final buffers
    unsigned pad[31];
    double flare_part[16][3];
};

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// thread r publishes to rank r: field 0 = h_ready (+ flare partial sums), 1 = tile_done, 2 = consumed
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
    } else {
        __threadfence_system();
        st_release_sys(&p->consumed, serial);
    }
}

// thread i waits until flags[i] >= serial (flags are in this GPU's own memory)
__global__ void peer_wait_kernel(const unsigned* __restrict__ flags, int n, unsigned serial) {
    const int i = threadIdx.x;
    if (i < n) {
        while ((int)(ld_acquire_sys(flags + i) - serial) < 0) __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace

static void tile_rows(int H, int world, int rank, int* r0, int* r1) {   // == dist.tile_rows
    const int base = H / world, extra = H % world;
    *r0 = rank * base + (rank < extra ? rank : extra);
    *r1 = *r0 + base + (rank < extra ? 1 : 0);
}

extern "C" int bhr_peer_export(bhr_ctx* ctx, bhr_ipc_handle out[4]) {
    if (!ctx || !out) return BHR_ERR_INVALID;
    BHR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
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
    if (!ctx || !all || world < 1 || world > 16 || rank < 0 || rank >= world) return BHR_ERR_INVALID;
    if (!ctx->peer_sync_own) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_peer_export must run first");
    if (ctx->peer_world) BHR_FAIL(ctx, BHR_ERR_STATE, "peers are already attached");
    BHR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
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
    // which buffer holds row y of the H-blurred layer
    const float** rows = (const float**)malloc(sizeof(float*) * ctx->H);
    if (!rows) BHR_FAIL(ctx, BHR_ERR_NOMEM, "host allocation failed");
    for (int r = 0; r < world; ++r) {
        int r0, r1;
        tile_rows(ctx->H, world, r, &r0, &r1);
        for (int y = r0; y < r1; ++y) rows[y] = ctx->peer_hblur[r];
    }
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_row_src, sizeof(float*) * ctx->H));
    BHR_CUDA(ctx, cudaMemcpy(ctx->d_row_src, rows, sizeof(float*) * ctx->H, cudaMemcpyHostToDevice));
    free(rows);
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_peer_sync, sizeof(PeerSync*) * 16));
    BHR_CUDA(ctx, cudaMemcpy(ctx->d_peer_sync, ctx->peer_sync, sizeof(PeerSync*) * 16, cudaMemcpyHostToDevice));
    BHR_CUDA(ctx, cudaMalloc(&ctx->d_flare_params, bhr_flare_params_size()));
    ctx->peer_rank = rank; ctx->peer_world = world; ctx->peer_serial = 0;
    return BHR_OK;
}

static int wait_consumed(bhr_ctx* ctx) {     // before the composite stores into rank 0's final buffers
    if (ctx->peer_rank != 0) {
        peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(&ctx->peer_sync_own->consumed, 1, ctx->peer_serial - 1);
        BHR_CUDA(ctx, cudaGetLastError());
    }
    return BHR_OK;
}

extern "C" int bhr_render_tiled_peer(bhr_ctx* ctx, const bhr_camera* cam, uint32_t flags, float* out_f32, uint8_t* out_u8) {
    if (!ctx || !cam) return BHR_ERR_INVALID;
    if (!ctx->peer_world) BHR_FAIL(ctx, BHR_ERR_STATE, "bhr_peer_attach has not run");
    BHR_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int rank = ctx->peer_rank, world = ctx->peer_world;
    const unsigned s = ++ctx->peer_serial;
    PeerSync* mine = ctx->peer_sync_own;
    // distributed egress: every rank copies its own rows into the (shared, page-locked) host frame
    const bool own_egress = rank != 0 && (out_f32 || out_u8);
    const bool host_out = out_f32 || out_u8;
    int row0, row1;
    tile_rows(ctx->H, world, rank, &row0, &row1);
    if (rank == 0) {
        // the caller is back for another frame: it is done with frame s - 1 (host frame / final buffers)
        peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, nullptr, s - 1, 2);
    }
    // every rank has finished frame s - 1 (its V pass no longer reads my H-blurred rows)
    peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(mine->tile_done, world, s - 1);
    int rc = bhr_render_rows_stage1(ctx, cam, flags, row0, row1);
    if (rc) return rc;
    const bool flare = ctx->cfg.lens_flare && !(flags & BHR_SKIP_FLARE);
    if (flare) {
        rc = bhr_launch_flare_sums(ctx, row0, row1);
        if (rc) return rc;
    }
    peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, flare ? ctx->d_flare_sums : nullptr, s, 0);
    peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(mine->h_ready, world, s);
    if (flare) {
        rc = bhr_launch_flare_params(ctx, &mine->flare_part[0][0], world, ctx->d_flare_params);
        if (rc) return rc;
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
    const size_t row3 = (size_t)ctx->W * 3;
    if (own_egress) {
        rc = wait_consumed(ctx);                 // the host frame still holds s - 1 until rank 0's caller returns for more
        if (rc) return rc;
        if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32 + row0 * row3, ctx->final_f32 + row0 * row3,
                                                   (row1 - row0) * row3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8 + row0 * row3, ctx->final_u8 + row0 * row3, (row1 - row0) * row3,
                                                  cudaMemcpyDeviceToHost, ctx->stream));
    }
    peer_publish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_peer_sync, rank, world, nullptr, s, 1);
    BHR_CUDA(ctx, cudaGetLastError());
    if (rank == 0) {
        const bool distributed = host_out && ctx->peer_distributed;
        if (distributed) {
            // my own rows go out right away; the other ranks' rows arrive over their own links
            if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32 + row0 * row3, ctx->final_f32 + row0 * row3,
                                                       (row1 - row0) * row3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8 + row0 * row3, ctx->final_u8 + row0 * row3, (row1 - row0) * row3,
                                                      cudaMemcpyDeviceToHost, ctx->stream));
            peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(mine->tile_done, world, s);
        } else {
            peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(mine->tile_done, world, s);
            const size_t n3 = (size_t)ctx->H * row3;
            if (out_f32) BHR_CUDA(ctx, cudaMemcpyAsync(out_f32, ctx->final_f32, n3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            if (out_u8) BHR_CUDA(ctx, cudaMemcpyAsync(out_u8, ctx->final_u8, n3, cudaMemcpyDeviceToHost, ctx->stream));
        }
        BHR_CUDA(ctx, cudaGetLastError());
        if (host_out) BHR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return BHR_OK;
}

extern "C" int bhr_peer_set_distributed_egress(bhr_ctx* ctx, int enabled) {
    if (!ctx) return BHR_ERR_INVALID;
    ctx->peer_distributed = enabled ? 1 : 0;
    return BHR_OK;
}

extern "C" int bhr_peer_detach(bhr_ctx* ctx) {
    if (!ctx) return BHR_ERR_INVALID;
    if (!ctx->peer_world) return BHR_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->peer_world; ++r) {
        if (r == ctx->peer_rank) continue;
        cudaIpcCloseMemHandle(ctx->peer_hblur[r]); cudaIpcCloseMemHandle(ctx->peer_final_f32[r]);
        cudaIpcCloseMemHandle(ctx->peer_final_u8[r]); cudaIpcCloseMemHandle(ctx->peer_sync[r]);
    }
    cudaFree(ctx->d_row_src); cudaFree(ctx->d_peer_sync); cudaFree(ctx->d_flare_params);
    ctx->d_row_src = nullptr; ctx->d_peer_sync = nullptr; ctx->d_flare_params = nullptr;
    ctx->peer_world = 0;
    return BHR_OK;
}
