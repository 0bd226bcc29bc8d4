// Shared device/host helpers for the KGE hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/kge_b200.h"

#define KGE_WARP 32

void kge_set_error(const char* fmt, ...);

#define KGE_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            kge_set_error("%s:%d CUDA error %d (%s) in `%s`", __FILE__, __LINE__, (int)_e,    \
                          cudaGetErrorString(_e), #expr);                                     \
            return -2;                                                                        \
        }                                                                                     \
    } while (0)

#define KGE_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            kge_set_error(__VA_ARGS__);                                                       \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

// per-step scalars read from device memory when a step runs from a captured graph
#define KGE_HOST_RING 4
struct KgeStepDyn {
    uint64_t step;
    float    lr_t;
    float    pad;
};

#define KGE_GRAPH_SLOTS 8  // {full batch, last batch} x two batch buffers, with room to spare
struct KgeGraphEntry {
    kge_train_args  key;       // args with step = 0
    cudaStream_t    stream = nullptr;
    cudaGraphExec_t exec = nullptr;
    int             seen = 0;  // calls with this key so far (the first runs eagerly: allocations, attributes)
    uint64_t        ws_epoch = 0;  // g_kge_ws_epoch at capture
    int             variant = 0;   // 0: whole step captured; 1 + set: pipelined main part reading buffer set `set`
    uint64_t        last_use = 0;
};

// bumped whenever a workspace buffer moves: captured graphs hold raw workspace pointers
inline uint64_t g_kge_ws_epoch = 0;

// grow-only device buffer owned by the ctx
struct KgeBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) {
            KGE_CUDA_CHECK(cudaDeviceSynchronize());
            KGE_CUDA_CHECK(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        KGE_CUDA_CHECK(cudaMalloc(&p, want));
        cap = want;
        ++g_kge_ws_epoch;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct kge_ctx {
    int device = 0;
    int sm_count = 148;
    // training workspace
    KgeBuf sort_tmp;
    KgeBuf loss_scr, pos_off, p2p_counter;
    KgeBuf ss_hist, ss_aux;  // kge_sort_small.cu: pass-1 output; grid barrier words + per-tile digit histograms
    KgeBuf repl, keep, grad_rows, loss_part, neg_scores, partial, span_head, reg_partial, touched;
    KgeBuf ks_in, ks_sel, ks_sorted, sel_flags, sel_count;
    // second set of the per-step corruption / sort-key buffers {repl, keep, ks_in, ks_sorted}: a pipelined step
    // (KGE_F_PIPELINE) swaps the sets and runs emit + sort on the side stream, beside the previous step.
    // set_id names the physical set the four members above hold right now; ev_set_free[i] = the last step that
    // read set i has finished (recorded by every single-GPU step), ev_pro[i] = emit + sort into set i done
    KgeBuf alt_repl, alt_keep, alt_ks_in, alt_ks_sorted;
    int    set_id = 0;
    bool   dim_pipelined = false;  // the dimension-sharded step in progress ran its prologue on the side stream
    cudaEvent_t ev_set_free[2] = {nullptr, nullptr}, ev_pro_emit[2] = {nullptr, nullptr}, ev_pro_sorted[2] = {nullptr, nullptr};
    // owner-side slot selection (kge_train_select): count travels to the host behind an event
    // side stream of the single-GPU step (sort + loss reduction beside the forward/backward kernel)
    cudaStream_t  side = nullptr;
    cudaStream_t  lstream = nullptr;  // loss reduction of a PIPELINED step: keeps the side stream free for the next prologue
    cudaEvent_t   ev_fork = nullptr, ev_sorted = nullptr, ev_fwd = nullptr, ev_loss = nullptr;
    // optional per-phase timing of kge_train_step (bench instrumentation): emit | fwd_bwd | reduce | spans
    bool          timing = false, tpending = false;
    cudaEvent_t   tev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double        tacc[6] = {0, 0, 0, 0, 0, 0};
    int           tcount = 0;
    int*          h_count = nullptr;
    cudaEvent_t   ev_count = nullptr;
    bool          sel_valid = false;
    int           rank_nl = 0;  // non-linearity of the kge_rank_counts call in progress
    const int32_t* sel_keys = nullptr;
    int64_t       sel_n = 0, sel_begin = 0, sel_end = 0;
    // staging for the host-buffer entry points
    KgeBuf h_pos, h_loss, h_test, h_counts, h_ranks;
    // whole-step CUDA graphs of the host-buffer training entry (kge_train_step_host): the per-step
    // scalars (step counter, Adam's bias-corrected rate) travel through a small device block that a
    // memcpy node refreshes from pinned host memory, so one instantiated graph serves every step
    cudaStream_t gmain = nullptr;  // capture-able stream the graphed host step runs on
    cudaEvent_t  ev_gin = nullptr;
    // the batch of host step t+1 is copied in on its own stream while step t runs: two device batch buffers,
    // ev_h2d[i] = copy into buffer i done, ev_posfree[i] = the step that read buffer i last has finished
    cudaStream_t cstream = nullptr;
    KgeBuf       h_pos2[2];
    cudaEvent_t  ev_h2d[2] = {nullptr, nullptr}, ev_posfree[2] = {nullptr, nullptr};
    KgeStepDyn* h_dyn = nullptr;  // pinned ring of KGE_HOST_RING blocks (one per in-flight host step)
    cudaEvent_t  ev_host[4] = {nullptr, nullptr, nullptr, nullptr};  // end of the host step that used ring slot i
    uint64_t     host_tick = 0;   // host steps submitted so far
    KgeBuf      d_dyn;
    KgeGraphEntry graphs[KGE_GRAPH_SLOTS];
    // ranking workspace
    KgeBuf q_fold, q_hi, q_lo, e_hi, e_lo, pos_q, excl_lo, excl_hi;
    // filter index (sorted, deduplicated composites and the entity column of each)
    KgeBuf f_sp_comp, f_po_comp, f_sp_ent, f_po_ent, f_tmp, f_tmp2, f_count;
    int64_t f_n_sp = 0, f_n_po = 0;  // capacity (pre-unique) sizes
    int64_t f_E = 0, f_R = 0;
    bool    f_valid = false;
    // cached 3xTF32 split of the local entity shard (ranking)
    const float* split_src = nullptr;
    int64_t split_rows = 0;
    int     split_K = 0;
    uint64_t split_epoch = 0;
};

// ------------------------------------------------------------------------------------------
// table access (row-range sharded over peers)
// ------------------------------------------------------------------------------------------
struct TableView {
    float*  shard[KGE_MAX_SHARDS];
    int64_t rows_per_shard;
    int32_t n_shards;
    int32_t K;
};

static inline TableView make_view(const kge_table& t) {
    TableView v;
    for (int i = 0; i < KGE_MAX_SHARDS; ++i) v.shard[i] = t.shard[i];
    v.rows_per_shard = t.rows_per_shard > 0 ? t.rows_per_shard : t.rows;
    v.n_shards = t.n_shards;
    v.K = t.K;
    return v;
}

static inline bool table_present(const kge_table& t) {
    for (int i = 0; i < KGE_MAX_SHARDS; ++i)
        if (t.shard[i]) return true;
    return false;
}

__device__ __forceinline__ float* table_row(const TableView& t, int64_t row) {
    if (t.n_shards == 1) return t.shard[0] + row * (int64_t)t.K;
    int s = (int)(row / t.rows_per_shard);
    return t.shard[s] + (row - (int64_t)s * t.rows_per_shard) * (int64_t)t.K;
}

// ------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA unit, SASS UBLKCP) wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// spin with a watchdog: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
// 1-D bulk copy global -> shared through the TMA unit; completion is signalled on `bar` as `bytes`
// of transaction count.  dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// tf.cast(score * 1e5, tf.int32): fp32 multiply, truncate toward zero
// (reference models/EmbeddingModel.py:2010-2014; utils/constants.py:87)
__device__ __forceinline__ int quantise_score(float s) {
    return __float2int_rz(__fmul_rn(s, 1e5f));
}

// embedding_model_params['non_linearity'] applied to every score before the loss / the rank comparison
// (models/EmbeddingModel.py:679-689, :801-812, :1868-1881): linear | tanh | sigmoid | softplus, the last one the
// reference's custom_softplus log(1 + 9999*exp(x)) with gradient 1 - 1/(1 + 9999*exp(x)) (:89-96)
__device__ __forceinline__ float apply_nl(int nl, float s) {
    if (nl == KGE_NL_LINEAR) return s;
    if (nl == KGE_NL_TANH) return tanhf(s);
    if (nl == KGE_NL_SIGMOID) return 1.f / (1.f + expf(-s));
    return logf(1.f + 9999.f * expf(s));
}
// d nl(s) / ds from the raw score
__device__ __forceinline__ float apply_nl_grad(int nl, float s) {
    if (nl == KGE_NL_LINEAR) return 1.f;
    if (nl == KGE_NL_TANH) {
        const float t = tanhf(s);
        return 1.f - t * t;
    }
    if (nl == KGE_NL_SIGMOID) {
        const float g = 1.f / (1.f + expf(-s));
        return g * (1.f - g);
    }
    return 1.f - 1.f / (1.f + 9999.f * expf(s));
}

// Philox4x32-10 counter-based RNG (Salmon et al. 2011), one call per negative.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline bool model_is_complex(int model) { return model == KGE_COMPLEX || model == KGE_HOLE; }
static inline int  model_row_width(int model, int k) { return model_is_complex(model) ? 2 * k : k; }
