// Context, error reporting, scoring (predict) and IPC helpers of the C ABI (include/kge_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

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
                      &c->grad_rows, &c->loss_part, &c->loss_scr, &c->pos_off, &c->p2p_counter, &c->ss_hist, &c->ss_aux, &c->neg_scores, &c->partial, &c->span_head, &c->ks_in, &c->ks_sel, &c->ks_sorted, &c->sel_flags, &c->sel_count, &c->h_pos, &c->h_loss, &c->h_test, &c->h_counts, &c->h_ranks, &c->q_fold, &c->q_hi, &c->q_lo, &c->e_hi, &c->e_lo,
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
    for (cudaEvent_t e : c->tev)
        if (e) cudaEventDestroy(e);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->lstream) cudaStreamDestroy(c->lstream);
    delete c;
    return 0;
}

extern "C" int64_t kge_ctx_workspace_bytes(kge_ctx* c) {
    if (!c) return 0;
    KgeBuf* bufs[] = {&c->sort_tmp, &c->repl, &c->keep,
                      &c->grad_rows, &c->loss_part, &c->loss_scr, &c->pos_off, &c->p2p_counter, &c->ss_hist, &c->ss_aux, &c->neg_scores, &c->partial, &c->span_head, &c->ks_in, &c->ks_sel, &c->ks_sorted, &c->sel_flags, &c->sel_count, &c->h_pos, &c->h_loss, &c->h_test, &c->h_counts, &c->h_ranks, &c->q_fold, &c->q_hi, &c->q_lo, &c->e_hi, &c->e_lo,
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
// All-reduce(sum) of the partial scores of the column-sharded step over peer memory (NVLink / NVSwitch P2P loads and stores,
// no NCCL): every rank owns 1/W of the range, pulls that slice of every rank's partial sums (fixed rank order => the same
// bits on every rank), and pushes the totals into every rank's totals buffer.  Two flag words per (rank, peer) in peer
// memory carry the two hand-shakes -- "my partial sums are written" before the pull, "my slice of your totals is written"
// before the kernel ends -- as monotonic sequence numbers, so nothing is ever reset.  Every wait is preceded on every rank by
// the signal the others wait for, in the same kernel, so the exchange cannot deadlock; a spin that runs out traps (a
// failed launch) instead of hanging the GPU.
// Against NCCL's all-reduce the latency is what matters: the payload is (1+eta) floats per positive -- 1.4 MB per piece for cfg3
// on 8 GPUs, where an NCCL call costs ~100 us and this kernel ~20 us -- and no collective kernel sits on the SMs polling
// while the next phase kernel runs.
// ------------------------------------------------------------------------------------------------
struct P2PReduceArgs {
    const float* sums[KGE_MAX_SHARDS];   // rank p's partial sums (local or peer mapping), same offsets on every rank
    float*       totals[KGE_MAX_SHARDS]; // rank p's totals buffer
    uint32_t*    flags[KGE_MAX_SHARDS];  // rank p's flag words: [0,W) "sums of rank s ready", [8,8+W) "slice of rank s written"
    unsigned int* counter;               // local: CTAs of this launch that finished their slice
    int rank, world;
    int64_t off, len;                    // float range reduced by this launch (multiple of 4, 16-byte aligned)
    uint32_t seq;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t seq) {
    for (uint32_t it = 0; (int32_t)(ld_acquire_sys(p) - seq) < 0; ++it) {
        if (it > (1u << 24)) __trap();  // ~ seconds: a peer never arrived
        __nanosleep(256);
    }
}

__global__ void __launch_bounds__(256) kge_allreduce_p2p_kernel(P2PReduceArgs A) {
    __shared__ bool last;
    const int W = A.world;
    // (1) the partial sums of this rank were written by the kernel before this one on the stream: tell every rank
    if (blockIdx.x == 0 && (int)threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(A.flags[threadIdx.x] + A.rank, A.seq);
    }
    // (2) wait until every rank's partial sums are there
    if ((int)threadIdx.x < W) spin_until(A.flags[A.rank] + threadIdx.x, A.seq);
    __syncthreads();
    // (3) this rank's slice: sum over the ranks in rank order, totals to every rank
    const int64_t nv = A.len / 4;
    const int64_t per = (nv + W - 1) / W;
    const int64_t v0 = (int64_t)A.rank * per, v1 = min(nv, v0 + per);
    for (int64_t v = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = A.off + 4 * v;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < KGE_MAX_SHARDS; ++p) {
            if (p < W) {
                const float4 x = *reinterpret_cast<const float4*>(A.sums[p] + i);
                acc.x += x.x;
                acc.y += x.y;
                acc.z += x.z;
                acc.w += x.w;
            }
        }
#pragma unroll
        for (int p = 0; p < KGE_MAX_SHARDS; ++p)
            if (p < W) *reinterpret_cast<float4*>(A.totals[p] + i) = acc;
    }
    // (4) the last CTA to finish tells every rank that this rank's slice of its totals is written, (5) and waits for the
    // same from every rank: the kernel does not end before this rank's totals are complete
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(A.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        if ((int)threadIdx.x < W) {
            __threadfence_system();
            st_release_sys(A.flags[threadIdx.x] + 8 + A.rank, A.seq);
            spin_until(A.flags[A.rank] + 8 + threadIdx.x, A.seq);
        }
        __syncthreads();
        if (threadIdx.x == 0) *A.counter = 0u;
    }
}

extern "C" int kge_allreduce_p2p(kge_ctx* ctx, const kge_table* sums, const kge_table* totals, const kge_table* flags, int rank,
                                 int64_t off, int64_t len, uint32_t seq, void* stream) {
    KGE_REQUIRE(ctx != nullptr && sums != nullptr && totals != nullptr && flags != nullptr, "kge_allreduce_p2p: null argument");
    const int W = sums->n_shards;
    KGE_REQUIRE(W >= 1 && W <= KGE_MAX_SHARDS && totals->n_shards == W && flags->n_shards == W && rank >= 0 && rank < W,
                "kge_allreduce_p2p: bad rank / world (%d of %d)", rank, W);
    KGE_REQUIRE(off >= 0 && len >= 0 && off % 4 == 0 && len % 4 == 0, "kge_allreduce_p2p: range [%lld,+%lld) must be a multiple of 4 floats",
                (long long)off, (long long)len);
    if (len == 0) return 0;
    if (ctx->p2p_counter.cap == 0) {
        if (ctx->p2p_counter.reserve(256)) return -2;
        KGE_CUDA_CHECK(cudaMemset(ctx->p2p_counter.p, 0, 256));
    }
    P2PReduceArgs A;
    for (int p = 0; p < KGE_MAX_SHARDS; ++p) {
        A.sums[p] = p < W ? sums->shard[p] : nullptr;
        A.totals[p] = p < W ? totals->shard[p] : nullptr;
        A.flags[p] = p < W ? reinterpret_cast<uint32_t*>(flags->shard[p]) : nullptr;
        KGE_REQUIRE(p >= W || (A.sums[p] && A.totals[p] && A.flags[p]), "kge_allreduce_p2p: rank %d's buffers are not mapped", p);
    }
    A.counter = ctx->p2p_counter.as<unsigned int>();
    A.rank = rank;
    A.world = W;
    A.off = off;
    A.len = len;
    A.seq = seq;
    // a SMALL grid: the CTAs poll peer flags while they wait for the slowest rank, and the kernel runs beside the phase kernel
    // of the next piece -- 296 CTAs cost that kernel 50 % (measured on 8 GPUs); 16 CTAs of 256 threads keep ~1 MB of peer
    // loads in flight, enough for the link.  KGE_P2P_CTAS overrides (A/B).
    static int max_ctas = -1;
    if (max_ctas < 0) {
        const char* e = getenv("KGE_P2P_CTAS");
        max_ctas = (e != nullptr && atoi(e) > 0) ? atoi(e) : 16;
    }
    const int64_t per = (len / 4 + W - 1) / W;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((per + 255) / 256, (int64_t)max_ctas));
    kge_allreduce_p2p_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A);
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
