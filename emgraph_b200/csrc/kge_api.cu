// Context, error reporting, scoring (predict) and IPC helpers of the C ABI (include/kge_b200.h).
#include <stdarg.h>
#include <string.h>

#include "kge_common.cuh"

static thread_local char g_err[1024] = "";

void kge_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* kge_last_error(void) { return g_err; }
extern "C" int kge_abi_version(void) { return KGE_ABI_VERSION; }

extern "C" int kge_ctx_create(int device, kge_ctx** out) {
    KGE_REQUIRE(out != nullptr, "kge_ctx_create: null out");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        kge_set_error("kge_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return -3;
    }
    KGE_REQUIRE(device >= 0 && device < count, "kge_ctx_create: device %d out of range (0..%d)", device, count - 1);
    KGE_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    KGE_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    KGE_REQUIRE(prop.major == 10, "kge_ctx_create: device %d is sm_%d%d; this build targets sm_100a (B200) only", device,
                prop.major, prop.minor);
    kge_ctx* c = new kge_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return 0;
}

extern "C" int kge_ctx_destroy(kge_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    KgeBuf* bufs[] = {&c->sort_tmp, &c->repl, &c->keep,
                      &c->grad_rows, &c->loss_part, &c->loss_scr, &c->pos_off, &c->neg_scores, &c->partial, &c->span_head, &c->ks_in, &c->ks_sel, &c->ks_sorted, &c->sel_flags, &c->sel_count, &c->h_pos, &c->h_loss, &c->h_test, &c->h_counts, &c->h_ranks, &c->q_fold, &c->q_hi, &c->q_lo, &c->e_hi, &c->e_lo,
                      &c->pos_q, &c->excl_lo, &c->excl_hi, &c->f_sp_comp, &c->f_po_comp, &c->f_sp_ent, &c->f_po_ent,
                      &c->f_tmp, &c->f_tmp2, &c->f_count};
    for (KgeBuf* b : bufs) b->release();
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->h_dyn) cudaFreeHost(c->h_dyn);
    if (c->ev_gin) cudaEventDestroy(c->ev_gin);
    for (int i = 0; i < 4; ++i)
        if (c->ev_host[i]) cudaEventDestroy(c->ev_host[i]);
    if (c->gmain) cudaStreamDestroy(c->gmain);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_posfree[i]) cudaEventDestroy(c->ev_posfree[i]);
        c->h_pos2[i].release();
    }
    if (c->cstream) cudaStreamDestroy(c->cstream);
    for (int i = 0; i < 2; ++i)
        for (cudaEvent_t e : {c->ev_set_free[i], c->ev_pro_emit[i], c->ev_pro_sorted[i]})
            if (e) cudaEventDestroy(e);
    for (KgeBuf* b : {&c->alt_repl, &c->alt_keep, &c->alt_ks_in, &c->alt_ks_sorted}) b->release();
    c->d_dyn.release();
    for (KgeGraphEntry& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (c->ev_count) cudaEventDestroy(c->ev_count);
    for (cudaEvent_t e : {c->ev_fork, c->ev_sorted, c->ev_fwd, c->ev_loss})
        if (e) cudaEventDestroy(e);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->lstream) cudaStreamDestroy(c->lstream);
    delete c;
    return 0;
}

extern "C" int64_t kge_ctx_workspace_bytes(kge_ctx* c) {
    if (!c) return 0;
    KgeBuf* bufs[] = {&c->sort_tmp, &c->repl, &c->keep,
                      &c->grad_rows, &c->loss_part, &c->loss_scr, &c->pos_off, &c->neg_scores, &c->partial, &c->span_head, &c->ks_in, &c->ks_sel, &c->ks_sorted, &c->sel_flags, &c->sel_count, &c->h_pos, &c->h_loss, &c->h_test, &c->h_counts, &c->h_ranks, &c->q_fold, &c->q_hi, &c->q_lo, &c->e_hi, &c->e_lo,
                      &c->pos_q, &c->excl_lo, &c->excl_hi, &c->f_sp_comp, &c->f_po_comp, &c->f_sp_ent, &c->f_po_ent,
                      &c->f_tmp, &c->f_tmp2, &c->f_count};
    int64_t tot = (int64_t)c->h_pos2[0].cap + (int64_t)c->h_pos2[1].cap;
    for (KgeBuf* b : {&c->alt_repl, &c->alt_keep, &c->alt_ks_in, &c->alt_ks_sorted}) tot += (int64_t)b->cap;
    for (KgeBuf* b : bufs) tot += (int64_t)b->cap;
    return tot;
}

// ------------------------------------------------------------------------------------------------
// predict: one warp per triple, lanes stride the row (any K)
// reference models/EmbeddingModel.py:2101-2147 (_lookup_embeddings :490-533 + _fn)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float score_triple_warp(int model, int k, int K, const float* __restrict__ s,
                                                   const float* __restrict__ p, const float* __restrict__ o, int lane) {
    float acc = 0.f;
    if (model == KGE_TRANSE_L1) {
        for (int c = lane; c < K; c += 32) acc += fabsf(s[c] + p[c] - o[c]);
        return -warp_sum(acc);
    }
    if (model == KGE_TRANSE_L2) {
        for (int c = lane; c < K; c += 32) {
            float u = s[c] + p[c] - o[c];
            acc = fmaf(u, u, acc);
        }
        return -sqrtf(warp_sum(acc));
    }
    if (model == KGE_DISTMULT) {
        for (int c = lane; c < K; c += 32) acc = fmaf(s[c] * p[c], o[c], acc);
        return warp_sum(acc);
    }
    for (int c = lane; c < k; c += 32) {
        float sr = s[c], si = s[c + k], pr = p[c], pi = p[c + k], orr = o[c], oi = o[c + k];
        acc = fmaf(pr * sr - pi * si, orr, acc);
        acc = fmaf(pr * si + pi * sr, oi, acc);
    }
    acc = warp_sum(acc);
    return model == KGE_HOLE ? (2.0f / (float)k) * acc : acc;
}

__global__ void kge_score_kernel(int model, int k, TableView ent, const float* __restrict__ rel,
                                 const int32_t* __restrict__ triples, int64_t n, float* __restrict__ out, int nl) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n) return;
    const int K = ent.K;
    const float* s = table_row(ent, triples[3 * t + 0]);
    const float* p = rel + (size_t)triples[3 * t + 1] * K;
    const float* o = table_row(ent, triples[3 * t + 2]);
    float f = score_triple_warp(model, k, K, s, p, o, lane);
    if (lane == 0) out[t] = apply_nl(nl, f);
}

extern "C" int kge_score(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                         const int32_t* triples, int64_t n, float* out, void* stream) {
    return kge_predict(ctx, model, k, ent, rel, R, triples, n, KGE_NL_LINEAR, out, stream);
}

extern "C" int kge_predict(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                           const int32_t* triples, int64_t n, int non_linearity, float* out, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_score: null ctx");
    KGE_REQUIRE(non_linearity >= KGE_NL_LINEAR && non_linearity <= KGE_NL_SOFTPLUS, "Invalid non-linearity");
    KGE_REQUIRE(model >= KGE_TRANSE_L1 && model <= KGE_HOLE, "kge_score: unknown model %d", model);
    KGE_REQUIRE(ent != nullptr && rel != nullptr && out != nullptr, "kge_score: null tensor");
    KGE_REQUIRE(ent->K == model_row_width(model, k), "kge_score: table width %d != internal_k %d", ent->K,
                model_row_width(model, k));
    (void)R;
    if (n == 0) return 0;
    KGE_REQUIRE(triples != nullptr, "kge_score: null triples");
    const int warps = 8;
    kge_score_kernel<<<(unsigned)((n + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        model, k, make_view(*ent), rel, triples, n, out, non_linearity);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CUDA IPC: map a peer rank's shard into this process (one process per GPU, NVLink/NVSwitch P2P)
// ------------------------------------------------------------------------------------------------
// device memory that can be exported to peer processes (plain cudaMalloc: torch's caching allocator
// sub-allocates, which cudaIpcGetMemHandle cannot address)
extern "C" int kge_dev_alloc(int64_t bytes, void** out) {
    KGE_REQUIRE(out != nullptr && bytes >= 0, "kge_dev_alloc: bad argument");
    *out = nullptr;
    if (bytes == 0) return 0;
    KGE_CUDA_CHECK(cudaMalloc(out, (size_t)bytes));
    return 0;
}

extern "C" int kge_dev_free(void* p) {
    if (p) KGE_CUDA_CHECK(cudaFree(p));
    return 0;
}

extern "C" int kge_ipc_export(void* dev_ptr, void* handle_out64) {
    KGE_REQUIRE(dev_ptr && handle_out64, "kge_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    KGE_CUDA_CHECK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle_out64, &h, 64);
    return 0;
}

extern "C" int kge_ipc_open(const void* handle64, void** dev_ptr_out) {
    KGE_REQUIRE(handle64 && dev_ptr_out, "kge_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    KGE_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int kge_ipc_close(void* dev_ptr) {
    KGE_REQUIRE(dev_ptr, "kge_ipc_close: null argument");
    KGE_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

extern "C" int kge_enable_peer_access(int device, int peer) {
    int can = 0;
    KGE_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device, peer));
    KGE_REQUIRE(can, "kge_enable_peer_access: device %d cannot access peer %d", device, peer);
    KGE_CUDA_CHECK(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return 0;
    }
    KGE_CUDA_CHECK(e);
    return 0;
}
