// KGE training step for sm_100a: corruption generation, sort keys, segmented reduction of duplicate
// rows and the sparse row-wise optimizer.  The fused forward/loss/backward kernel lives in
// kge_train_fwd.cuh (one translation unit per scoring model).
//
//   emit        : counter-based corruption generation (Philox4x32-10) + per-slot sort keys
//   fwd_bwd     : kge_train_fwd.cuh
//   reduce_apply: radix sort of (row id, slot) + atomics-free two-level segmented reduction:
//                 every warp owns KGE_CH consecutive sorted slots; runs that live inside one chunk
//                 are reduced and fed straight into Adam/Adagrad/momentum/SGD for that row; runs that
//                 cross chunk borders (hub entities) leave one partial row per chunk and are finished
//                 by kge_span_apply_kernel in chunk order.  Summation order is fixed by the sort =>
//                 bit-reproducible; only touched rows are read or written.
//
// Replaces reference evaluation/protocol.py:531-659 (generate_corruptions_for_fit) and
// training/{adam,adagrad,momentum,sgd}.py (+ the Keras OptimizerV2 sparse apply with its duplicate
// index summation).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "kge_train_fwd.cuh"
#include "kge_apply.cuh"

// ------------------------------------------------------------------------------------------------
// slot layout for one batch of n positives (S = (3+eta)*n slots)
//   [0,n)            subject row of positive i          key = s_i         gradient: gbuf row i      (gs)
//   [n,2n)           object row of positive i           key = o_i         gradient: gbuf row n+i    (go)
//   [2n,2n+eta*n)    replacement row of negative (j,i)  key = repl[j*n+i] gradient: F(coef, Q row, r)
//   [2n+eta*n,S)     relation row of positive i         key = E + p_i     gradient: gbuf row 2n+i   (gp)
// ------------------------------------------------------------------------------------------------

__global__ void kge_emit_kernel(const int32_t* __restrict__ pos, int64_t n, int eta, int64_t E, int side,
                                const int32_t* __restrict__ repl_in, const uint8_t* __restrict__ keep_in,
                                uint64_t seed, uint64_t step, uint64_t neg_base,
                                const int32_t* __restrict__ neg_list, int64_t neg_n,
                                int32_t* __restrict__ repl_out, uint8_t* __restrict__ keep_out,
                                int32_t* __restrict__ keys, uint64_t* __restrict__ packed, const KgeStepDyn* __restrict__ dyn) {
    int64_t S = (int64_t)(3 + eta) * n;
    if (dyn != nullptr) step = dyn->step;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < S; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t key;
        uint32_t side_bits = 0;  // negatives: the corrupted side rides in the sort entry (kge_apply.cuh: KGE_SLOT_*)
        if (t < n) {
            key = pos[3 * t + 0];
        } else if (t < 2 * n) {
            key = pos[3 * (t - n) + 2];
        } else if (t < 2 * n + (int64_t)eta * n) {
            int64_t q = t - 2 * n;  // j*n + i
            int32_t r;
            uint8_t ks;
            // keep_in[q]: 0 / 1 fix the side of this negative; >= 2 (or no array) leaves it to `side`.  It may
            // come without repl_in: a list-valued corrupt_side stacked into one batch
            const uint8_t kin = keep_in != nullptr ? keep_in[q] : (uint8_t)2;
            if (repl_in != nullptr) {
                r = repl_in[q];
                ks = kin < 2 ? kin : (side == KGE_SIDE_O ? 1 : 0);
            } else {
                uint64_t g = neg_base + (uint64_t)q;
                uint32_t o[4];
                philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32), o);
                // replacement ~ U{0..entities_size-1} (evaluation/protocol.py:616-619), or a uniform pick from
                // entities_list (:620-641); multiply-shift mapping
                r = (int32_t)(((uint64_t)o[0] * (uint64_t)(neg_n > 0 ? neg_n : E)) >> 32);
                if (neg_list != nullptr) r = neg_list[r];
                ks = kin < 2 ? kin : (side == KGE_SIDE_SO ? (uint8_t)(o[1] >> 31) : (side == KGE_SIDE_O ? 1 : 0));
            }
            repl_out[q] = r;
            keep_out[q] = ks;
            key = r;
            side_bits = KGE_SLOT_HAS_SIDE | (ks ? KGE_SLOT_SIDE : 0u);
        } else {
            key = (int32_t)E + pos[3 * (t - 2 * n - (int64_t)eta * n) + 1];
        }
        if (keys != nullptr) keys[t] = key;
        if (packed != nullptr) packed[t] = ((uint64_t)(uint32_t)key << 32) | (uint64_t)((uint32_t)t | side_bits);
    }
}

// deterministic fixed-order reduction of the per-positive loss terms: CTA b sums terms b, b+G, ... in double, the last CTA
// to finish (ticket counter) adds the G partial sums in index order -- one launch, the same bits every time
#define KGE_LOSS_CTAS 64
__global__ void __launch_bounds__(256) kge_loss_reduce_kernel(const float* __restrict__ part, int64_t n, float* __restrict__ out,
                                                              double* __restrict__ scratch, unsigned int* __restrict__ ticket) {
    __shared__ double sm[256];
    __shared__ bool last;
    double acc = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) acc += (double)part[t];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        scratch[blockIdx.x] = sm[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double tot = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) tot += ((volatile double*)scratch)[b];
        out[0] = (float)tot;
        *ticket = 0u;  // ready for the next launch
    }
}

#include "kge_dim.cuh"

int kge_launch_apply_group(const ApplyParams& P, int tmode, cudaStream_t st);  // kge_apply_group.cu
bool kge_apply_group_ok(const ApplyParams& P);
bool kge_small_sort_ok(int64_t n_items, int64_t n_keys);  // kge_sort_small.cu
int kge_small_sort(kge_ctx* ctx, const uint64_t* in, int64_t n_items, int64_t n_keys, uint64_t* out, cudaStream_t st);
int kge_launch_apply_wide(const ApplyParams& P, int tmode, cudaStream_t st);  // kge_apply_wide.cu
bool kge_apply_wide_ok(const ApplyParams& P);


// Level 1: one warp per chunk of KGE_CH sorted slots.  NCA > 0: lanes own ALL their column vectors of
// the row at once (K <= 128*NCA) so that the row's w/m/v and two slots' rows are in flight together;
// NCA == 0: generic column loop (any K, scalar columns).
#define KGE_RA_WARPS 4
// Run ownership: a run is reduced by ONE warp whenever it touches at most two chunks -- the warp that
// holds the run's head also takes the run's slots at the front of the next chunk (lanes 16..31 decode
// them).  Only runs that touch three or more chunks (hub entities) go through per-chunk partial rows
// and the span kernels.  Every chunk evaluates the same predicate from the sorted keys alone:
//   head side : the run is long  <=>  the first key of chunk w+2 still equals the run's key
//   tail side : the run's head lies in chunk w-1  <=>  the last key of chunk w-2 differs
// CS > 1: CS warps share a chunk, each owning a contiguous 1/CS of the columns (same keys, same decisions, half the
// registers): a batch that is only a few waves of warps deep -- cfg3: 6 100 chunks -- is a chain of ~10 dependent memory
// round trips per warp, and twice the resident warps is what shortens it.
template <int V, int TMODE, int NCA, int CS>
__global__ void __launch_bounds__(KGE_RA_WARPS * 32) kge_reduce_apply_kernel(ApplyParams P) {
    __shared__ SlotMeta meta[KGE_RA_WARPS][2 * KGE_CH];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t wg = (int64_t)blockIdx.x * KGE_RA_WARPS + wib;
    const int64_t w = wg / CS;
    const int part_id = (int)(wg - w * CS);  // which 1/CS of the columns this warp owns
    const int64_t b0 = w * KGE_CH;
    if (b0 >= P.n_keys) return;
    const int cnt = (int)min((int64_t)KGE_CH, P.n_keys - b0);
    const int K = P.ent.K;

    // lanes [0,16): own chunk; lanes [16,32): the next chunk (candidates for a spill-over run)
    int32_t key = -2;
    if (b0 + lane < P.n_keys) {
        const uint64_t kv = P.ks[b0 + lane];
        key = (int32_t)(kv >> 32);
        meta[wib][lane] = decode_slot(P.G, (uint32_t)(kv & 0xffffffffu));
    }
    int32_t key_prev = -1, key_prev2 = -1, key_next2 = -1;
    if (lane == 0 && b0 > 0) key_prev = (int32_t)(P.ks[b0 - 1] >> 32);
    if (lane == 1 && b0 > KGE_CH) key_prev2 = (int32_t)(P.ks[b0 - KGE_CH - 1] >> 32);
    if (lane == 2 && b0 + 2 * KGE_CH < P.n_keys) key_next2 = (int32_t)(P.ks[b0 + 2 * KGE_CH] >> 32);
    key_prev = __shfl_sync(0xffffffffu, key_prev, 0);
    key_prev2 = __shfl_sync(0xffffffffu, key_prev2, 1);
    key_next2 = __shfl_sync(0xffffffffu, key_next2, 2);
    const int32_t key_next = __shfl_sync(0xffffffffu, key, KGE_CH);  // -2 when there is no next chunk
    const int32_t key_last = __shfl_sync(0xffffffffu, key, cnt - 1);
    const int32_t left = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane < cnt && (lane == 0 || key != left);
    unsigned heads = __ballot_sync(0xffffffffu, head);
    // slots at the front of the next chunk that continue this chunk's last run
    const int ext = __popc(__ballot_sync(0xffffffffu, lane >= KGE_CH && key == key_last && key >= 0));
    __syncwarp();

    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    const bool pf = prefetch_on(P) && !no_update && P.prefetch_wide;
    while (heads) {
        const int a = __ffs(heads) - 1;
        heads &= heads - 1;
        int b = heads ? (__ffs(heads) - 1) : cnt;
        const int32_t skey = __shfl_sync(0xffffffffu, key, a);
        if (pf) {
            // the runs of a chunk are walked one after the other, each paying a memory round trip for its row's w, m, v:
            // put the optimizer rows of the NEXT run in flight now (L2 prefetch, one 128-byte line per lane and tensor).  One
            // run ahead only: prefetching the whole chunk thrashes L2 (measured, profiles/r02_b_summary.md)
            const int a1 = heads ? (__ffs(heads) - 1) : -1;
            const int32_t k1 = __shfl_sync(0xffffffffu, key, max(a1, 0));
            if (a1 >= 0) {
                const RowPtrs r1 = resolve_row(P, k1);
                if (r1.owned)
                    for (int off = lane * 32; off < K; off += 32 * 32) {
                        prefetch_l2(r1.w + off);
                        if (need_m && r1.m != nullptr) prefetch_l2(r1.m + off);
                        if (need_v && r1.v != nullptr) prefetch_l2(r1.v + off);
                    }
            }
        }
        const bool open_start = (a == 0) && (skey == key_prev);
        bool open_end = (b == cnt) && (skey == key_next);
        const RowPtrs r = resolve_row(P, skey);
        if (!r.owned) continue;
        if (open_start && !open_end && key_prev2 != skey) continue;  // the head's warp (chunk w-1) reduces this run
        if (!open_start && open_end && key_next2 != skey) {
            b = cnt + ext;  // the run ends inside the next chunk: finish it here
            open_end = false;
        }
        if (!open_start && open_end && lane == 0 && part_id == 0) P.span_list[atomicAdd(P.span_count, 1)] = (int32_t)w;
        const bool complete = !open_start && !open_end;
        if (complete && lane == 0 && part_id == 0) mark_touched(P, skey);
        float* part = P.partial + ((size_t)(2 * w + (open_start ? 0 : 1))) * K;
        float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
        if constexpr (NCA > 0) {
            // lanes past the end of the row (of this warp's share of it) read a clamped (valid, duplicate) vector and
            // never store: no divergence inside the load/accumulate loops
            float g[NCA][V], rc[NCA][V], mv[NCA][V], vv[NCA][V];
            int cc[NCA];
            const int nvec_row = K / V;
            const int cps = (nvec_row + CS - 1) / CS;                  // vectors per column share
            const int v_lo = part_id * cps, v_hi = min(v_lo + cps, nvec_row);  // this warp's vectors [v_lo, v_hi)
#pragma unroll
            for (int i = 0; i < NCA; ++i) {
                cc[i] = min(v_lo + lane + 32 * i, v_hi - 1) * V;
#pragma unroll
                for (int x = 0; x < V; ++x) g[i][x] = rc[i][x] = mv[i][x] = vv[i][x] = 0.f;
            }
            if (complete || TMODE != 0) {
#pragma unroll
                for (int i = 0; i < NCA; ++i) ldg_vec<V>(rc[i], r.w + cc[i]);
            }
            if (complete && need_m) {
#pragma unroll
                for (int i = 0; i < NCA; ++i) ldg_vec<V>(mv[i], r.m + cc[i]);
            }
            if (complete && need_v) {
#pragma unroll
                for (int i = 0; i < NCA; ++i) ldg_vec<V>(vv[i], r.v + cc[i]);
            }
            int u = a;
            for (; u + 2 <= b; u += 2) {
                float v2[2][NCA][V];
                const SlotMeta m0 = meta[wib][u], m1 = meta[wib][u + 1];
#pragma unroll
                for (int i = 0; i < NCA; ++i) ldg_vec<V>(v2[0][i], m0.row + cc[i]);
#pragma unroll
                for (int i = 0; i < NCA; ++i) ldg_vec<V>(v2[1][i], m1.row + cc[i]);
#pragma unroll
                for (int i = 0; i < NCA; ++i) add_slot<V, TMODE>(g[i], v2[0][i], m0.c, m0.mode, rc[i]);
#pragma unroll
                for (int i = 0; i < NCA; ++i) add_slot<V, TMODE>(g[i], v2[1][i], m1.c, m1.mode, rc[i]);
            }
            if (u < b) {
                float v1[NCA][V];
                const SlotMeta m0 = meta[wib][u];
#pragma unroll
                for (int i = 0; i < NCA; ++i) ldg_vec<V>(v1[i], m0.row + cc[i]);
#pragma unroll
                for (int i = 0; i < NCA; ++i) add_slot<V, TMODE>(g[i], v1[i], m0.c, m0.mode, rc[i]);
            }
#pragma unroll
            for (int i = 0; i < NCA; ++i) {
                const int vi = v_lo + lane + 32 * i;
                const int c0 = vi * V;
                if (vi >= v_hi) continue;
                if (!complete) {
                    st_vec<V>(part + c0, g[i]);
                    continue;
                }
                reg_add<V>(P, r.is_rel, g[i], rc[i]);
                if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + c0, g[i]);
                if (no_update) continue;
                opt_math<V>(P, reset, g[i], rc[i], mv[i], vv[i]);
                if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + c0, mv[i]);
                if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + c0, vv[i]);
                st_vec<V>(r.w + c0, rc[i]);
            }
        } else {
            static_assert(NCA > 0 || CS == 1, "the generic column loop is not split");
            for (int c0 = lane * V; c0 < K; c0 += 32 * V) {
                float g[V], rc[V], mv[V], vv[V];
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] = rc[x] = mv[x] = vv[x] = 0.f;
                if (complete || TMODE != 0) ld_vec<V>(rc, r.w + c0);
                if (complete && need_m) ld_vec<V>(mv, r.m + c0);
                if (complete && need_v) ld_vec<V>(vv, r.v + c0);
                int u = a;
                for (; u + 4 <= b; u += 4) {
                    float v4[4][V];
#pragma unroll
                    for (int q = 0; q < 4; ++q) ld_vec<V>(v4[q], meta[wib][u + q].row + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) add_slot<V, TMODE>(g, v4[q], meta[wib][u + q].c, meta[wib][u + q].mode, rc);
                }
                for (; u < b; ++u) {
                    float v1[V];
                    ld_vec<V>(v1, meta[wib][u].row + c0);
                    add_slot<V, TMODE>(g, v1, meta[wib][u].c, meta[wib][u].mode, rc);
                }
                if (!complete) {
                    st_vec<V>(part + c0, g);
                    continue;
                }
                reg_add<V>(P, r.is_rel, g, rc);
                if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + c0, g);
                if (no_update) continue;
                opt_math<V>(P, reset, g, rc, mv, vv);
                if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + c0, mv);
                if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + c0, vv);
                st_vec<V>(r.w + c0, rc);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Staged variant (trilinear models, local buffers, K % 4 == 0, K <= 512): every row the warp will
// read -- the gradient/query rows of its slots and, at the end of each complete run, the row's w, m, v
// -- is one entry of a per-warp SEQUENCE that a ring of 2^ns_log2 shared-memory row slots streams in
// by 1-D bulk async copies (TMA unit, SASS UBLKCP).  The lane that owns a slot (or a run head) issues
// the copy as soon as the entry enters the ring window, so a warp keeps up to 2^ns_log2 rows in
// flight instead of two, and the optimizer state is prefetched while the run is still being summed.
// ------------------------------------------------------------------------------------------------
struct RunPlan {
    bool process, complete, open_start, span_head;
    int b;
};

__device__ __forceinline__ RunPlan plan_run(int a, int b_next, int cnt, int ext, int32_t skey, int32_t key_prev, int32_t key_prev2,
                                            int32_t key_next, int32_t key_next2, bool owned) {
    RunPlan pl;
    pl.b = b_next;
    pl.open_start = (a == 0) && (skey == key_prev);
    bool open_end = (b_next == cnt) && (skey == key_next);
    pl.process = owned && !(pl.open_start && !open_end && key_prev2 != skey);
    if (!pl.open_start && open_end && key_next2 != skey) {
        pl.b = cnt + ext;
        open_end = false;
    }
    pl.span_head = !pl.open_start && open_end;
    pl.complete = !pl.open_start && !open_end;
    return pl;
}

__device__ __forceinline__ void lds_vec4(float (&d)[4], uint32_t addr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(addr));
}

template <int NCA>
__global__ void __launch_bounds__(KGE_RA_WARPS * 32) kge_reduce_apply_staged_kernel(ApplyParams P, int ns_log2) {
    constexpr int V = 4;
    __shared__ SlotMeta meta[KGE_RA_WARPS][2 * KGE_CH];
    __shared__ __align__(8) uint64_t bars_all[KGE_RA_WARPS][16];
    extern __shared__ __align__(16) float ring_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * KGE_RA_WARPS + wib;
    const int64_t b0 = w * KGE_CH;
    if (b0 >= P.n_keys) return;
    const int cnt = (int)min((int64_t)KGE_CH, P.n_keys - b0);
    const int K = P.ent.K;
    const int NS = 1 << ns_log2;
    const uint32_t row_bytes = (uint32_t)K * 4u;
    uint64_t* bars = bars_all[wib];
    float* ring = ring_all + (size_t)wib * NS * K;
    if (lane == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }

    int32_t key = -2;
    if (b0 + lane < P.n_keys) {
        const uint64_t kv = P.ks[b0 + lane];
        key = (int32_t)(kv >> 32);
        meta[wib][lane] = decode_slot(P.G, (uint32_t)(kv & 0xffffffffu));
    }
    int32_t key_prev = -1, key_prev2 = -1, key_next2 = -1;
    if (lane == 0 && b0 > 0) key_prev = (int32_t)(P.ks[b0 - 1] >> 32);
    if (lane == 1 && b0 > KGE_CH) key_prev2 = (int32_t)(P.ks[b0 - KGE_CH - 1] >> 32);
    if (lane == 2 && b0 + 2 * KGE_CH < P.n_keys) key_next2 = (int32_t)(P.ks[b0 + 2 * KGE_CH] >> 32);
    key_prev = __shfl_sync(0xffffffffu, key_prev, 0);
    key_prev2 = __shfl_sync(0xffffffffu, key_prev2, 1);
    key_next2 = __shfl_sync(0xffffffffu, key_next2, 2);
    const int32_t key_next = __shfl_sync(0xffffffffu, key, KGE_CH);
    const int32_t key_last = __shfl_sync(0xffffffffu, key, cnt - 1);
    const int32_t left = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane < cnt && (lane == 0 || key != left);
    const unsigned heads0 = __ballot_sync(0xffffffffu, head);
    const int ext = __popc(__ballot_sync(0xffffffffu, lane >= KGE_CH && key == key_last && key >= 0));
    __syncwarp();

    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    const int n_state = no_update ? 0 : 1 + (need_m ? 1 : 0) + (need_v ? 1 : 0);  // w [, m [, v]]

    // ---- plan: sequence position of every entry this warp will read
    const RowPtrs mine = resolve_row(P, max(key, 0));  // the row of this lane's own slot (= its run's row)
    int my_pos = -1, st_pos = -1, st_n = 0;
    {
        int seq = 0;
        unsigned hh = heads0;
        while (hh) {
            const int a = __ffs(hh) - 1;
            hh &= hh - 1;
            const int bn = hh ? (__ffs(hh) - 1) : cnt;
            const int32_t skey = __shfl_sync(0xffffffffu, key, a);
            const bool owned = __shfl_sync(0xffffffffu, (int)mine.owned, a) != 0;
            const RunPlan pl = plan_run(a, bn, cnt, ext, skey, key_prev, key_prev2, key_next, key_next2, owned);
            if (!pl.process) continue;
            const int nst = pl.complete ? n_state : 0;
            if (lane >= a && lane < pl.b) my_pos = seq + (lane - a);
            if (lane == a) {
                st_pos = seq + (pl.b - a);
                st_n = nst;
            }
            seq += (pl.b - a) + nst;
        }
    }
    unsigned issued = 0;  // bit 0: slot row, bits 1..3: w, m, v of the run this lane heads
    int consumed = 0;
    auto issue_window = [&]() {
        const int lim = consumed + NS;
        if (my_pos >= 0 && !(issued & 1u) && my_pos < lim) {
            const uint32_t sl = (uint32_t)my_pos & (uint32_t)(NS - 1);
            mbar_expect_tx(bars + sl, row_bytes);
            bulk_g2s(ring + (size_t)sl * K, meta[wib][lane].row, row_bytes, bars + sl);
            issued |= 1u;
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (t < st_n && !(issued & (2u << t)) && st_pos + t < lim) {
                const uint32_t sl = (uint32_t)(st_pos + t) & (uint32_t)(NS - 1);
                const float* src = t == 0 ? mine.w : (t == 1 ? mine.m : mine.v);
                mbar_expect_tx(bars + sl, row_bytes);
                bulk_g2s(ring + (size_t)sl * K, src, row_bytes, bars + sl);
                issued |= (2u << t);
            }
        }
    };
    issue_window();

    int cc[NCA];
#pragma unroll
    for (int i = 0; i < NCA; ++i) cc[i] = min((lane + 32 * i) * V, K - V);
    auto wait_entry = [&](int q) -> uint32_t {
        const uint32_t sl = (uint32_t)q & (uint32_t)(NS - 1);
        mbar_wait(bars + sl, ((uint32_t)q >> ns_log2) & 1u);
        return smem_u32(ring + (size_t)sl * K);
    };

    unsigned heads = heads0;
    while (heads) {
        const int a = __ffs(heads) - 1;
        heads &= heads - 1;
        const int bn = heads ? (__ffs(heads) - 1) : cnt;
        const int32_t skey = __shfl_sync(0xffffffffu, key, a);
        const RowPtrs r = resolve_row(P, skey);
        const RunPlan pl = plan_run(a, bn, cnt, ext, skey, key_prev, key_prev2, key_next, key_next2, r.owned);
        if (!pl.process) continue;
        if (pl.span_head && lane == 0) P.span_list[atomicAdd(P.span_count, 1)] = (int32_t)w;
        float* part = P.partial + ((size_t)(2 * w + (pl.open_start ? 0 : 1))) * K;
        float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
        float g[NCA][V];
#pragma unroll
        for (int i = 0; i < NCA; ++i)
#pragma unroll
            for (int x = 0; x < V; ++x) g[i][x] = 0.f;
        int u = a;
        for (; u + 2 <= pl.b; u += 2) {
            float v2[2][NCA][V];
            const float c0 = meta[wib][u].c, c1 = meta[wib][u + 1].c;
            const uint32_t s0 = wait_entry(consumed), s1 = wait_entry(consumed + 1);
#pragma unroll
            for (int i = 0; i < NCA; ++i) lds_vec4(v2[0][i], s0 + (uint32_t)cc[i] * 4u);
#pragma unroll
            for (int i = 0; i < NCA; ++i) lds_vec4(v2[1][i], s1 + (uint32_t)cc[i] * 4u);
            __syncwarp();
            consumed += 2;
            issue_window();
#pragma unroll
            for (int i = 0; i < NCA; ++i)
#pragma unroll
                for (int x = 0; x < V; ++x) g[i][x] = fmaf(c0, v2[0][i][x], g[i][x]);
#pragma unroll
            for (int i = 0; i < NCA; ++i)
#pragma unroll
                for (int x = 0; x < V; ++x) g[i][x] = fmaf(c1, v2[1][i][x], g[i][x]);
        }
        if (u < pl.b) {
            float v1[NCA][V];
            const float c0 = meta[wib][u].c;
            const uint32_t s0 = wait_entry(consumed);
#pragma unroll
            for (int i = 0; i < NCA; ++i) lds_vec4(v1[i], s0 + (uint32_t)cc[i] * 4u);
            __syncwarp();
            consumed += 1;
            issue_window();
#pragma unroll
            for (int i = 0; i < NCA; ++i)
#pragma unroll
                for (int x = 0; x < V; ++x) g[i][x] = fmaf(c0, v1[i][x], g[i][x]);
        }
        if (!pl.complete) {
#pragma unroll
            for (int i = 0; i < NCA; ++i) {
                const int c0 = (lane + 32 * i) * V;
                if (c0 < K) st_vec<V>(part + c0, g[i]);
            }
            continue;
        }
        if (dbg != nullptr) {
#pragma unroll
            for (int i = 0; i < NCA; ++i) {
                const int c0 = (lane + 32 * i) * V;
                if (c0 < K) st_vec<V>(dbg + (size_t)r.row * K + c0, g[i]);
            }
        }
        if (no_update) continue;
        const uint32_t sw = wait_entry(consumed);
        const uint32_t sm = need_m ? wait_entry(consumed + 1) : 0u;
        const uint32_t sv = need_v ? wait_entry(consumed + 2) : 0u;
#pragma unroll
        for (int i = 0; i < NCA; ++i) {
            float rc[V], mv[V], vv[V];
#pragma unroll
            for (int x = 0; x < V; ++x) mv[x] = vv[x] = 0.f;
            lds_vec4(rc, sw + (uint32_t)cc[i] * 4u);
            if (need_m) lds_vec4(mv, sm + (uint32_t)cc[i] * 4u);
            if (need_v) lds_vec4(vv, sv + (uint32_t)cc[i] * 4u);
            opt_math<V>(P, reset, g[i], rc, mv, vv);
            const int c0 = (lane + 32 * i) * V;
            if (c0 >= K) continue;
            if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + c0, mv);
            if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + c0, vv);
            st_vec<V>(r.w + c0, rc);
        }
        __syncwarp();
        consumed += n_state;
        issue_window();
    }
}

// last chunk whose first slot still carries `key`.  Round 1 probes the next 32 chunks (nearly every span ends there); a hub
// that goes on is finished by a 32-ary search over the rest of the sorted list -- "chunk c starts with key" is monotone in
// c -- so the top hub of a Zipf graph (600 chunks at cfg3) costs 4 dependent loads instead of 19 (measured: this chain
// was most of kge_span_apply_kernel's duration).
__device__ __forceinline__ int64_t span_last_chunk(const ApplyParams& P, int64_t w, int32_t key, int64_t n_chunks, int lane) {
    {
        const int64_t c = w + 1 + lane;
        const bool same = c < n_chunks && (int32_t)(P.ks[c * KGE_CH] >> 32) == key;
        const unsigned mk = __ballot_sync(0xffffffffu, same);
        if (mk != 0xffffffffu) return w + (__ffs(~mk) - 1);
    }
    int64_t lo = w + 32, hi = n_chunks;  // chunk lo starts with key, chunk hi does not (or does not exist)
    while (hi - lo > 1) {
        const int64_t d = (hi - lo + 31) / 32;
        const int64_t c = lo + (int64_t)(lane + 1) * d;
        const bool same = c < hi && (int32_t)(P.ks[c * KGE_CH] >> 32) == key;
        const int cnt = __popc(__ballot_sync(0xffffffffu, same));  // a prefix of the lanes
        lo += (int64_t)cnt * d;
        hi = min(hi, lo + d);
    }
    return lo;
}

__device__ __forceinline__ const float* span_entry(const ApplyParams& P, int64_t w, int64_t e, int K) {
    // entry 0 is the head's own partial (2w+1); entry e >= 1 is chunk w+e's start-open partial (2(w+e))
    return P.partial + (size_t)(e == 0 ? 2 * w + 1 : 2 * (w + e)) * K;
}

// Level 2: runs that touch three or more chunks (hub entities; everything shorter is finished by the
// run-head warp in level 1).  A fixed grid walks the list of run heads; the KGE_SPAN_WARPS warps of a
// CTA add the per-chunk partial rows of one run (warp j takes partials j, j+KGE_SPAN_WARPS, ...), the
// partial sums are combined through shared memory in warp order (deterministic) and the optimizer is
// applied by the threads that own the columns (their w, m, v loads are issued before the summation).
// KGE_SPAN_WARPS is a template parameter: the longest run (the top hub of a Zipf graph: ~10 % of all slots,
// hundreds of partial rows) is one CTA's serial chain, so the kernel's duration is that run's partial count
// divided by the warps of a CTA.
template <int V, int KGE_SPAN_WARPS>
__global__ void __launch_bounds__(KGE_SPAN_WARPS * 32) kge_span_apply_kernel(ApplyParams P) {
    extern __shared__ __align__(16) float sred[];  // [KGE_SPAN_WARPS][K]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K = P.ent.K;
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    // with the warp-per-run pass in front (dense batches) only the hubs it set aside are left for the CTAs
    // This kernel is the last reader of the two counters: the CTA that draws the last ticket -- every other CTA has read
    // them by then -- zeroes them for the next step's reduction (a memset in front of the level-1 kernel was 2-3 us of
    // every step's critical path).  The ticket's value is only looked at when the CTA is done: nobody waits for it here.
    __shared__ int s_heads;
    int ticket = 0;
    if (threadIdx.x == 0) {
        s_heads = ((volatile int32_t*)P.span_count)[P.span_use_hubs ? 1 : 0];
        ticket = atomicAdd(P.span_ticket, 1);
    }
    __syncthreads();
    const int n_heads = s_heads;
    const int32_t* heads_list = P.span_use_hubs ? P.hub_list : P.span_list;
    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    for (int h = blockIdx.x; h < n_heads; h += gridDim.x) {
        const int64_t w = heads_list[h];
        const int32_t key = (int32_t)(P.ks[w * KGE_CH + KGE_CH - 1] >> 32);
        const RowPtrs r = resolve_row(P, key);
        if (threadIdx.x == 0) mark_touched(P, key);
        // columns owned by this thread in the final combine (at most one vector per thread when K <= 1024)
        const int c_own = threadIdx.x * V;
        float rc0[V], mv0[V], vv0[V];
#pragma unroll
        for (int x = 0; x < V; ++x) rc0[x] = mv0[x] = vv0[x] = 0.f;
        const bool own_fast = c_own < K && K <= KGE_SPAN_WARPS * 32 * V;
        if (own_fast) {
            ld_vec<V>(rc0, r.w + c_own);
            if (need_m) ld_vec<V>(mv0, r.m + c_own);
            if (need_v) ld_vec<V>(vv0, r.v + c_own);
        }
        const int64_t last = span_last_chunk(P, w, key, n_chunks, lane);
        const int64_t n_ent = last - w + 1;
        // The top hub of a Zipf graph owns ~10 % of all slots -- hundreds of partial rows -- and its CTA's chain of loads IS
        // the duration of this kernel.  The warps form a grid of nrg row groups x ncg column groups (a lane owns one vector
        // of its column group), so a warp keeps 8 partial rows in flight instead of walking the row's column groups one
        // after the other with 2: ~10 dependent round trips for a 600-chunk hub at K = 400 instead of ~40.
        const int ncg = (K + 32 * V - 1) / (32 * V);
        int nrg = KGE_SPAN_WARPS;  // rows of sred in use
        if (ncg <= KGE_SPAN_WARPS) {
            nrg = KGE_SPAN_WARPS / ncg;
            const int cg = wib % ncg, rg = wib / ncg;
            const int c0 = (cg * 32 + lane) * V;
            if (rg < nrg && c0 < K) {
                float g[V];
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] = 0.f;
                for (int64_t e = rg; e < n_ent; e += 8 * (int64_t)nrg) {
                    float t[8][V];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
#pragma unroll
                        for (int x = 0; x < V; ++x) t[q][x] = 0.f;
                        if (e + (int64_t)q * nrg < n_ent) ld_vec<V>(t[q], span_entry(P, w, e + (int64_t)q * nrg, K) + c0);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q)
#pragma unroll
                        for (int x = 0; x < V; ++x) g[x] += t[q][x];
                }
                st_vec<V>(sred + (size_t)rg * K + c0, g);
            }
        } else {
            for (int c0 = lane * V; c0 < K; c0 += 32 * V) {
                float g[V];
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] = 0.f;
                int64_t e = wib;
                for (; e + KGE_SPAN_WARPS < n_ent; e += 2 * KGE_SPAN_WARPS) {
                    float t0[V], t1[V];
                    ld_vec<V>(t0, span_entry(P, w, e, K) + c0);
                    ld_vec<V>(t1, span_entry(P, w, e + KGE_SPAN_WARPS, K) + c0);
#pragma unroll
                    for (int x = 0; x < V; ++x) g[x] += t0[x];
#pragma unroll
                    for (int x = 0; x < V; ++x) g[x] += t1[x];
                }
                if (e < n_ent) {
                    float t0[V];
                    ld_vec<V>(t0, span_entry(P, w, e, K) + c0);
#pragma unroll
                    for (int x = 0; x < V; ++x) g[x] += t0[x];
                }
                st_vec<V>(sred + (size_t)wib * K + c0, g);
            }
        }
        __syncthreads();
        float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
        for (int c0 = c_own; c0 < K; c0 += KGE_SPAN_WARPS * 32 * V) {
            float g[V], rc[V], mv[V], vv[V];
#pragma unroll
            for (int x = 0; x < V; ++x) {
                g[x] = 0.f;
                rc[x] = rc0[x];
                mv[x] = mv0[x];
                vv[x] = vv0[x];
            }
            if (!own_fast) {
                ld_vec<V>(rc, r.w + c0);
                if (need_m) ld_vec<V>(mv, r.m + c0);
                if (need_v) ld_vec<V>(vv, r.v + c0);
            }
            for (int j = 0; j < nrg; ++j) {
                float t0[V];
                ld_vec<V>(t0, sred + (size_t)j * K + c0);
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] += t0[x];
            }
            reg_add<V>(P, r.is_rel, g, rc);
            if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + c0, g);
            if (!no_update) {
                opt_math<V>(P, reset, g, rc, mv, vv);
                if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + c0, mv);
                if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + c0, vv);
                st_vec<V>(r.w + c0, rc);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && ticket == (int)gridDim.x - 1) {
        P.span_count[0] = 0;
        P.span_count[1] = 0;
        *P.span_ticket = 0;
    }
}

// Level 2 for DENSE batches (many more slots than rows: every row's run crosses chunk borders, so there are thousands of
// span heads, each with a handful of partial rows): one WARP per run instead of one 1024-thread CTA.  The warp adds the
// run's partial rows in chunk order (fixed order => reproducible) and applies the optimizer; runs with more than
// KGE_SPAN_WARP_MAX partial rows (the hubs) are put on hub_list for kge_span_apply_kernel.
#define KGE_SPAN_WARP_MAX 64
template <int V>
__global__ void __launch_bounds__(256) kge_span_warp_kernel(ApplyParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int K = P.ent.K;
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    const int n_heads = P.span_count[0];
    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    for (int64_t h = warp; h < n_heads; h += nwarps) {
        const int64_t w = P.span_list[h];
        const int32_t key = (int32_t)(P.ks[w * KGE_CH + KGE_CH - 1] >> 32);
        const int64_t last = span_last_chunk(P, w, key, n_chunks, lane);
        const int64_t n_ent = last - w + 1;
        if (n_ent > KGE_SPAN_WARP_MAX) {
            if (lane == 0) P.hub_list[atomicAdd(P.span_count + 1, 1)] = (int32_t)w;
            continue;
        }
        const RowPtrs r = resolve_row(P, key);
        if (lane == 0) mark_touched(P, key);
        float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
        for (int c0 = lane * V; c0 < K; c0 += 32 * V) {
            float g[V], rc[V], mv[V], vv[V];
#pragma unroll
            for (int x = 0; x < V; ++x) g[x] = mv[x] = vv[x] = 0.f;
            ld_vec<V>(rc, r.w + c0);
            if (need_m) ld_vec<V>(mv, r.m + c0);
            if (need_v) ld_vec<V>(vv, r.v + c0);
            int64_t e = 0;
            for (; e + 2 <= n_ent; e += 2) {
                float t0[V], t1[V];
                ld_vec<V>(t0, span_entry(P, w, e, K) + c0);
                ld_vec<V>(t1, span_entry(P, w, e + 1, K) + c0);
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] += t0[x];
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] += t1[x];
            }
            if (e < n_ent) {
                float t0[V];
                ld_vec<V>(t0, span_entry(P, w, e, K) + c0);
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] += t0[x];
            }
            reg_add<V>(P, r.is_rel, g, rc);
            if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + c0, g);
            if (!no_update) {
                opt_math<V>(P, reset, g, rc, mv, vv);
                if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + c0, mv);
                if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + c0, vv);
                st_vec<V>(r.w + c0, rc);
            }
        }
    }
}

// LP regulariser, rows the batch did not touch: g = d(lambda*|w|^p)/dw, fed to the same optimizer math.
// One warp per owned row (entity rows [row_begin,row_end), then every relation row).
template <int V>
__global__ void __launch_bounds__(256) kge_reg_dense_kernel(ApplyParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int K = P.ent.K;
    const int64_t n_own = P.row_end - P.row_begin;
    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    for (int64_t t = warp; t < n_own + P.R; t += nwarps) {
        const int32_t key = (int32_t)(t < n_own ? P.row_begin + t : P.E + (t - n_own));
        if ((P.touched[key >> 5] >> (key & 31)) & 1u) continue;
        const RowPtrs r = resolve_row(P, key);
        float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
        for (int c0 = lane * V; c0 < K; c0 += 32 * V) {
            float g[V], rc[V], mv[V], vv[V];
#pragma unroll
            for (int x = 0; x < V; ++x) g[x] = mv[x] = vv[x] = 0.f;
            ld_vec<V>(rc, r.w + c0);
            if (need_m) ld_vec<V>(mv, r.m + c0);
            if (need_v) ld_vec<V>(vv, r.v + c0);
            reg_add<V>(P, r.is_rel, g, rc);
            if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + c0, g);
            if (no_update) continue;
            opt_math<V>(P, reset, g, rc, mv, vv);
            if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + c0, mv);
            if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + c0, vv);
            st_vec<V>(r.w + c0, rc);
        }
    }
}

// lambda * sum |w|^p over n floats: per-block partial sums (fixed grid, fixed order => reproducible)
__global__ void __launch_bounds__(256) kge_reg_loss_kernel(const float* __restrict__ w, int64_t n, int p, float lam, double* __restrict__ partial) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        acc += (double)reg_term1(w[t], p);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = (double)lam * sm[0];
}
__global__ void kge_reg_loss_add_kernel(const double* __restrict__ partial, int n, float* __restrict__ loss_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += partial[i];
        loss_out[0] = (float)((double)loss_out[0] + acc);
    }
}

__global__ void kge_normalize_rows_kernel(float* emb, int64_t rows, int K) {
    // tf.clip_by_norm(ent_emb, clip_norm=1, axes=1)  (models/EmbeddingModel.py:1434-1439)
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float* p = emb + (size_t)r * K;
    float acc = 0.f;
    for (int c = lane; c < K; c += 32) acc = fmaf(p[c], p[c], acc);
    acc = warp_sum(acc);
    float nrm = sqrtf(acc);
    if (nrm > 1.f) {
        float inv = 1.f / nrm;
        for (int c = lane; c < K; c += 32) p[c] *= inv;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int validate_train(const kge_train_args* a) {
    KGE_REQUIRE(a != nullptr, "kge_train: null args");
    KGE_REQUIRE(a->model >= KGE_TRANSE_L1 && a->model <= KGE_HOLE, "kge_train: unknown model %d", a->model);
    KGE_REQUIRE(a->loss >= KGE_LOSS_PAIRWISE && a->loss <= KGE_LOSS_SELF_ADVERSARIAL, "Unsupported loss function: %d", a->loss);
    KGE_REQUIRE(a->opt >= KGE_OPT_ADAM && a->opt <= KGE_OPT_SGD, "Unsupported optimizer: %d", a->opt);
    KGE_REQUIRE(a->side >= KGE_SIDE_SO && a->side <= KGE_SIDE_O, "Invalid corruption side %d", a->side);
    KGE_REQUIRE(a->k > 0 && a->eta > 0, "kge_train: k and eta must be positive");
    KGE_REQUIRE(a->ent.K == model_row_width(a->model, a->k), "kge_train: table width %d != internal_k %d", a->ent.K,
                model_row_width(a->model, a->k));
    KGE_REQUIRE(a->ent.n_shards >= 1 && a->ent.n_shards <= KGE_MAX_SHARDS, "kge_train: bad shard count");
    KGE_REQUIRE(a->ent.rows + a->R < (int64_t)INT32_MAX, "kge_train: E+R must fit int32 sort keys");
    KGE_REQUIRE(a->n_pos >= 0 && (a->pos != nullptr || a->n_pos == 0), "kge_train: positives missing");
    KGE_REQUIRE(a->non_linearity >= KGE_NL_LINEAR && a->non_linearity <= KGE_NL_SOFTPLUS, "Invalid non-linearity");
    KGE_REQUIRE(a->neg_entities_n >= 0 && a->neg_entities_n <= a->ent.rows && (a->neg_entities == nullptr || a->neg_entities_n > 0),
                "kge_train: bad negative_corruption_entities (n=%lld)", (long long)a->neg_entities_n);
    return 0;
}

static int ensure_train_ws(kge_ctx* ctx, const kge_train_args* a) {
    int64_t n = a->n_pos;
    if (ctx->repl.reserve((size_t)a->eta * n * sizeof(int32_t))) return -2;
    if (ctx->keep.reserve((size_t)a->eta * n)) return -2;
    if (ctx->loss_part.reserve((size_t)(n + 1) * sizeof(float))) return -2;
    if (ctx->loss_scr.cap == 0) {  // partial sums + ticket counter of kge_loss_reduce_kernel; the ticket starts at 0
        if (ctx->loss_scr.reserve((KGE_LOSS_CTAS + 2) * sizeof(double))) return -2;
        KGE_CUDA_CHECK(cudaMemset(ctx->loss_scr.p, 0, (KGE_LOSS_CTAS + 2) * sizeof(double)));
    }
    return 0;
}

extern "C" int64_t kge_train_grad_floats(int eta, int64_t n_pos, int K) { return gbuf_floats(eta, n_pos, K); }
extern "C" int64_t kge_train_grad_head_floats(int eta, int64_t n_pos, int K) { (void)eta; return 3 * n_pos * (int64_t)K; }

static int emit_impl(kge_ctx* ctx, const kge_train_args* a, int32_t* keys_out, uint64_t* packed_out, cudaStream_t st,
                     const KgeStepDyn* dyn = nullptr) {
    if (int rc = ensure_train_ws(ctx, a)) return rc;
    int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    int threads = 256;
    int blocks = (int)std::min<int64_t>((S + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
    kge_emit_kernel<<<blocks, threads, 0, st>>>(a->pos, a->n_pos, a->eta, a->ent.rows, a->side, a->repl, a->keep_subj,
                                                a->seed, a->step, a->neg_index_base, a->neg_entities, a->neg_entities_n,
                                                ctx->repl.as<int32_t>(),
                                                ctx->keep.as<uint8_t>(), keys_out, packed_out, dyn);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_train_emit(kge_ctx* ctx, const kge_train_args* a, int32_t* keys_out, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_emit: null ctx");
    if (int rc = validate_train(a)) return rc;
    if (a->n_pos == 0) return 0;
    KGE_REQUIRE(keys_out != nullptr, "kge_train_emit: keys_out missing");
    ctx->sel_valid = false;
    return emit_impl(ctx, a, keys_out, nullptr, (cudaStream_t)stream);
}

// fork: when `side` is given the loss reduction runs there (behind ev_fwd) and signals ev_loss
static int fwd_bwd_impl(kge_ctx* ctx, const kge_train_args* a, float* grad_buf, cudaStream_t st, cudaStream_t side) {
    KGE_REQUIRE(grad_buf != nullptr, "kge_train_fwd_bwd: grad_buf missing");
    KGE_REQUIRE(a->loss_out != nullptr, "kge_train_fwd_bwd: loss_out missing");
    if (a->n_pos == 0) {
        KGE_CUDA_CHECK(cudaMemsetAsync(a->loss_out, 0, sizeof(float), st));
        return 0;
    }
    FwdBwdParams P;
    P.ent = make_view(a->ent);
    P.rel = a->rel;
    P.pos = a->pos;
    P.repl = ctx->repl.as<int32_t>();
    P.keep = ctx->keep.as<uint8_t>();
    P.n = a->n_pos;
    P.eta = a->eta;
    P.k = a->k;
    P.loss = a->loss;
    P.margin = a->margin;
    P.alpha = a->alpha;
    P.nl = a->non_linearity;
    P.scale = a->model == KGE_HOLE ? 2.0f / (float)(a->k_model > 0 ? a->k_model : a->k) : 1.0f;
    P.gbuf = grad_buf;
    P.loss_part = ctx->loss_part.as<float>();
    P.dbg_scores = a->dbg_scores;
    P.stage = a->stage;
    {   // tables that do not fit L2 make the candidate gather an HBM random-row gather: optional prefetch behind the ring
        // window, KGE_FWD_L2PF=1 (A/B knob)
        static int force = -2;
        if (force == -2) {
            const char* e = getenv("KGE_FWD_L2PF");
            force = e == nullptr ? -1 : (e[0] == '0' ? 0 : 1);
        }
        const bool big = (size_t)a->ent.rows * a->ent.K * sizeof(float) > ((size_t)64 << 20);
        (void)big;
        P.l2_prefetch = force > 0 ? 1 : 0;  // measured slower on cfg5 (fwd_bwd 0.210 -> 0.251 ms): off unless forced
    }
    int rc;
    switch (a->model) {
        case KGE_TRANSE_L1: rc = kge_launch_fwd_bwd_m0(P, ctx->sm_count, st); break;
        case KGE_TRANSE_L2: rc = kge_launch_fwd_bwd_m1(P, ctx->sm_count, st); break;
        case KGE_DISTMULT: rc = kge_launch_fwd_bwd_m2(P, ctx->sm_count, st); break;
        default: rc = kge_launch_fwd_bwd_m3(P, ctx->sm_count, st); break;
    }
    if (rc) return rc;
    cudaStream_t ls = st;
    if (side != nullptr) {
        KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_fwd, st));
        KGE_CUDA_CHECK(cudaStreamWaitEvent(side, ctx->ev_fwd, 0));
        ls = side;
    }
    kge_loss_reduce_kernel<<<KGE_LOSS_CTAS, 256, 0, ls>>>(ctx->loss_part.as<float>(), a->n_pos, a->loss_out, ctx->loss_scr.as<double>(),
                                                         reinterpret_cast<unsigned int*>(ctx->loss_scr.as<double>() + KGE_LOSS_CTAS));
    KGE_CUDA_CHECK(cudaGetLastError());
    if (side != nullptr) KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_loss, side));
    return 0;
}

extern "C" int kge_train_fwd_bwd(kge_ctx* ctx, const kge_train_args* a, float* grad_buf, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_fwd_bwd: null ctx");
    if (int rc = validate_train(a)) return rc;
    return fwd_bwd_impl(ctx, a, grad_buf, (cudaStream_t)stream, nullptr);
}

// Owner-side push (row-sharded multi-GPU): one warp per slot of the all-gathered keys; slots whose row
// this rank owns are copied into the staging buffer of the rank whose batch holds the slot.  The local
// shard is read at random (fast locally), the peer is written with contiguous 1-row stores.
template <int V>
__global__ void __launch_bounds__(256) kge_push_rows_kernel(const int32_t* __restrict__ keys, int64_t n_keys, int64_t S, int64_t ent_slots,
                                                            const float* __restrict__ shard, int64_t row_begin, int64_t row_end, int K,
                                                            TableView stage, int64_t t_start, int64_t group) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nvec = K / V;
    // every lane tests one slot; owned slots are then copied by the whole warp, two rows in flight.
    // Blocks of 32 slots are dealt round-robin over the destination ranks, starting behind this owner
    // (t_start = own + 1): at any moment every owner's stores are spread over all W destinations, so no
    // rank's NVLink ingress is the target of all owners at once (the all-to-all pattern)
    const int W = stage.n_shards;
    const int64_t bpr = (S + 31) / 32;  // blocks per rank
    // `group` consecutive blocks go to the same destination before the deal moves on (group = blocks per rank:
    // one destination region at a time, rotated per owner; group = 1: finest interleave)
    const int64_t gpr = (bpr + group - 1) / group;  // groups per rank
    for (int64_t vb0 = warp; vb0 < gpr * group * W; vb0 += nwarps) {
        const int64_t g = vb0 / group, in_g = vb0 - g * group;
        const int rr = (int)((g + t_start) % W);
        const int64_t lb = (g / W) * group + in_g;  // block inside the destination's slots
        if (lb >= bpr) continue;
        const int64_t vb = lb * W;  // (vb / W) == lb below
        const int64_t tl = lb * 32 + lane;
        int32_t key = -1;
        if (tl < ent_slots) key = keys[(int64_t)rr * S + tl];
        unsigned own = __ballot_sync(0xffffffffu, key >= row_begin && key < row_end);
        while (own) {
            const int l0 = __ffs(own) - 1;
            own &= own - 1;
            const int l1 = own ? __ffs(own) - 1 : -1;
            if (l1 >= 0) own &= own - 1;
            const int32_t k0 = __shfl_sync(0xffffffffu, key, l0);
            const int r0 = rr, r1 = rr;
            const int64_t s0 = (vb / W) * 32 + l0;
            const int32_t k1 = __shfl_sync(0xffffffffu, key, max(l1, 0));
            const int64_t s1 = (vb / W) * 32 + max(l1, 0);
            const float* a = shard + (int64_t)(k0 - row_begin) * K;
            const float* b = shard + (int64_t)(k1 - row_begin) * K;
            float* da = stage.shard[r0] + s0 * K;
            float* db = stage.shard[r1] + s1 * K;
            for (int c = lane; c < nvec; c += 64) {
                float xa[V], xb[V], xc[V], xd[V];
                const bool two = c + 32 < nvec;
                ld_vec<V>(xa, a + (size_t)c * V);
                if (two) ld_vec<V>(xc, a + (size_t)(c + 32) * V);
                if (l1 >= 0) {
                    ld_vec<V>(xb, b + (size_t)c * V);
                    if (two) ld_vec<V>(xd, b + (size_t)(c + 32) * V);
                }
                st_vec<V>(da + (size_t)c * V, xa);
                if (two) st_vec<V>(da + (size_t)(c + 32) * V, xc);
                if (l1 >= 0) {
                    st_vec<V>(db + (size_t)c * V, xb);
                    if (two) st_vec<V>(db + (size_t)(c + 32) * V, xd);
                }
            }
        }
    }
}

extern "C" int kge_train_push_rows(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                                   const kge_table* stage, int64_t row_begin, int64_t row_end, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_push_rows: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(stage != nullptr && keys_all != nullptr, "kge_train_push_rows: missing keys/stage");
    if (n_keys == 0 || row_end <= row_begin) return 0;
    const int K = a->ent.K;
    const int64_t S = (int64_t)(3 + a->eta) * a->n_pos, ent_slots = (int64_t)(2 + a->eta) * a->n_pos;
    KGE_REQUIRE(stage->K == K && stage->rows_per_shard == ent_slots && stage->n_shards >= 1 && n_keys == S * stage->n_shards,
                "kge_train_push_rows: stage must hold (2+eta)*n_pos rows per rank and n_keys = n_ranks*(3+eta)*n_pos");
    const int64_t rps = a->ent.rows_per_shard > 0 ? a->ent.rows_per_shard : a->ent.rows;
    const int own = a->ent.n_shards == 1 ? 0 : (int)(row_begin / rps);
    KGE_REQUIRE(own < a->ent.n_shards && a->ent.shard[own] != nullptr && row_begin == (int64_t)own * rps,
                "kge_train_push_rows: [row_begin,row_end) must be this rank's shard of a->ent");
    // A/B knobs (read per call): KGE_PUSH_CTAS = CTAs per SM, KGE_PUSH_GROUP = blocks of 32 slots that go to
    // one destination before the deal moves to the next (0 = a whole rank's slots).  Measured on 8 x B200
    // (profiles/r01_n_push_sweep_n8.txt): the all-to-all of 1 KiB stores saturates at ~470 GB/s per GPU whatever
    // the grid; few CTAs and one destination region at a time (rotated per owner) give the least skew.
    int ctas = 2;
    int64_t group = 0;
    if (const char* e = getenv("KGE_PUSH_CTAS")) ctas = std::max(1, std::min(16, atoi(e)));
    if (const char* e = getenv("KGE_PUSH_GROUP")) group = atoll(e);
    const int64_t bpr = (S + 31) / 32;
    if (group <= 0 || group > bpr) group = bpr;
    const int blocks = ctx->sm_count * ctas;
    const TableView sv = make_view(*stage);
    const int64_t t_start = (own + 1) % stage->n_shards;
    if (K % 4 == 0)
        kge_push_rows_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(keys_all, n_keys, S, ent_slots, a->ent.shard[own], row_begin, row_end, K, sv, t_start, group);
    else
        kge_push_rows_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(keys_all, n_keys, S, ent_slots, a->ent.shard[own], row_begin, row_end, K, sv, t_start, group);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// owner-side selection of the slots a rank must reduce: keys of its row range + every relation key
__global__ void kge_select_flag_kernel(const int32_t* __restrict__ keys, int64_t n, int64_t E, int64_t row_begin, int64_t row_end,
                                       uint64_t* __restrict__ packed, uint8_t* __restrict__ flags) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t key = keys[t];
        packed[t] = ((uint64_t)(uint32_t)key << 32) | (uint64_t)(uint32_t)t;
        flags[t] = (key >= E || (key >= row_begin && key < row_end)) ? 1 : 0;
    }
}

// KGE_REDUCE_STAGED=1 selects the bulk-copy staged reduction (measured slower than the register path on
// B200 for every benchmark shape -- the reduction is L2-throughput bound, not latency bound -- so it is
// off by default and kept for A/B measurements)
static inline bool reduce_staged_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_REDUCE_STAGED");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

// Warps per CTA of the span/hub reduction.  32 (one 1024-thread CTA per run head) unless the [32][K] staging
// rows would crowd shared memory; KGE_SPAN_WARPS=8|32 overrides (A/B knob, read once).  Measured on B200
// (profiles/r01_x_*): cfg3 span phase 17.2 -> 13.8 us, cfg4 23.3 -> 16.4 us with 32 warps.
static inline int span_warps(int K) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("KGE_SPAN_WARPS");
        forced = (e != nullptr && atoi(e) == 32) ? 32 : (e != nullptr && atoi(e) == 8 ? 8 : 0);
    }
    if (forced > 0) return forced;
    return (size_t)32 * K * sizeof(float) <= 100 * 1024 ? 32 : 8;
}

template <int V, int W>
static int launch_span(const ApplyParams& P, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)W * P.ent.K * sizeof(float);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        KGE_CUDA_CHECK(cudaFuncSetAttribute(kge_span_apply_kernel<V, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    kge_span_apply_kernel<V, W><<<sm_count, W * 32, smem, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int V, int NCA>
static int launch_apply_nca(const ApplyParams& P, int tmode, int sm_count, cudaStream_t st, cudaEvent_t mid, cudaEvent_t pre) {
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    dim3 grid((unsigned)((n_chunks + KGE_RA_WARPS - 1) / KGE_RA_WARPS)), block(KGE_RA_WARPS * 32);
    // P.span_count is zero here: zeroed when the buffer was allocated, and again by every kge_span_apply_kernel
    if (pre != nullptr) KGE_CUDA_CHECK(cudaEventRecord(pre, st));  // bench instrumentation: the level-1 kernel starts here
    bool staged = false;
    if constexpr (V == 4 && NCA > 0) {
        // staged rows need local buffers (bulk copies of peer memory are not used) and 16-byte rows
        staged = tmode == 0 && reduce_staged_enabled() && P.G.n_ranks == 1 && P.ent.n_shards == 1 && P.reg_p <= 0;
        if (staged) {
            const int K = P.ent.K;
            int l = 4;
            while (l > 2 && ((size_t)(1 << l)) * K * 4 > 13 * 1024) --l;
            const size_t smem = (size_t)KGE_RA_WARPS * (1 << l) * K * sizeof(float);
            static size_t staged_set = 0;
            if (smem > 48 * 1024 && smem > staged_set) {
                KGE_CUDA_CHECK(cudaFuncSetAttribute(kge_reduce_apply_staged_kernel<NCA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                staged_set = smem;
            }
            kge_reduce_apply_staged_kernel<NCA><<<grid, block, smem, st>>>(P, l);
        }
    }
    bool grouped = false;
    if constexpr (V == 4 && NCA == 1) {
        // narrow rows: a group of 8 / 16 lanes per chunk instead of a warp (kge_apply_group.cu); KGE_APPLY_GROUP=0: A/B
        static int grp_on = -1;
        if (grp_on < 0) {
            const char* e = getenv("KGE_APPLY_GROUP");
            grp_on = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        grouped = !staged && grp_on != 0 && kge_apply_group_ok(P);
        if (grouped)
            if (int rc = kge_launch_apply_group(P, tmode, st)) return rc;
    }
    // wide rows on one local buffer: the load-list kernel (kge_apply_wide.cu), KGE_APPLY_WIDE=1.  Off by default: measured on
    // B200 (profiles/r02_summary.md, session r2w) it shortens the dependent chain of a chunk from ~10 to ~3 memory round trips
    // and still takes 63 us against 55 us on cfg3 -- both kernels move the same 13 M L2 sectors, ~60 % of the measured L2
    // sector rate, and that, not latency, is what bounds the reduction of a dense batch
    bool wide = false;
    if constexpr (V == 4 && NCA > 0) {
        static int wide_on = -2;
        if (wide_on == -2) {
            const char* e = getenv("KGE_APPLY_WIDE");
            wide_on = e == nullptr ? -1 : atoi(e);
        }
        const bool want = wide_on > 0;
        wide = !staged && !grouped && want && kge_apply_wide_ok(P);
        if (wide)
            if (int rc = kge_launch_apply_wide(P, tmode, st)) return rc;
    }
    // the warp-per-chunk kernel on a shallow grid: two (four) warps per chunk, half (a quarter of) the columns each.
    // KGE_APPLY_SPLIT=0/2/4 forces it (A/B)
    int split = 0;  // warps per chunk: 0/1 = one; 2 or 4 = that many, each with 1/2 or 1/4 of the columns
    if constexpr (V == 4 && NCA >= 2) {
        static int force = -2;
        if (force == -2) {
            const char* e = getenv("KGE_APPLY_SPLIT");
            force = e == nullptr ? -1 : atoi(e);
        }
        // measured (profiles/r02_summary.md, sessions r2z / r2v): cfg3 (6 100 chunks) 87 -> 65 -> 62 us with 1 -> 2 -> 4 warps per
        // chunk, cfg4 (21 000 chunks) 268 -> 255 us with 4; a deep grid of narrower rows (cfg5, 43 000 chunks) already runs at
        // DRAM speed with one warp per chunk
        split = force >= 0 ? force : (n_chunks < (int64_t)sm_count * 256 ? (NCA >= 4 ? 4 : 2) : 0);
        if (split == 4 && NCA < 4) split = 2;
        if (split == 1) split = 0;
    }
    if (staged || grouped || wide) {
    } else if (split == 2) {
        if constexpr (V == 4 && NCA >= 2) {
            dim3 grid2((unsigned)((2 * n_chunks + KGE_RA_WARPS - 1) / KGE_RA_WARPS));
            if (tmode == 0) kge_reduce_apply_kernel<V, 0, NCA / 2, 2><<<grid2, block, 0, st>>>(P);
            else if (tmode == 1) kge_reduce_apply_kernel<V, 1, NCA / 2, 2><<<grid2, block, 0, st>>>(P);
            else kge_reduce_apply_kernel<V, 2, NCA / 2, 2><<<grid2, block, 0, st>>>(P);
        }
    } else if (split == 4) {
        if constexpr (V == 4 && NCA >= 4) {
            dim3 grid4((unsigned)((4 * n_chunks + KGE_RA_WARPS - 1) / KGE_RA_WARPS));
            if (tmode == 0) kge_reduce_apply_kernel<V, 0, NCA / 4, 4><<<grid4, block, 0, st>>>(P);
            else if (tmode == 1) kge_reduce_apply_kernel<V, 1, NCA / 4, 4><<<grid4, block, 0, st>>>(P);
            else kge_reduce_apply_kernel<V, 2, NCA / 4, 4><<<grid4, block, 0, st>>>(P);
        }
    } else if (tmode == 0) kge_reduce_apply_kernel<V, 0, NCA, 1><<<grid, block, 0, st>>>(P);
    else if (tmode == 1) kge_reduce_apply_kernel<V, 1, NCA, 1><<<grid, block, 0, st>>>(P);
    else kge_reduce_apply_kernel<V, 2, NCA, 1><<<grid, block, 0, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    if (mid != nullptr) KGE_CUDA_CHECK(cudaEventRecord(mid, st));
    if (P.span_use_hubs) {  // dense batch: a warp per run first, the CTAs take the hubs it sets aside
        kge_span_warp_kernel<V><<<sm_count * 8, 256, 0, st>>>(P);
        KGE_CUDA_CHECK(cudaGetLastError());
    }
    return span_warps(P.ent.K) == 32 ? launch_span<V, 32>(P, sm_count, st) : launch_span<V, 8>(P, sm_count, st);
}

static int launch_apply(const ApplyParams& P, int tmode, int sm_count, cudaStream_t st, cudaEvent_t mid, cudaEvent_t pre = nullptr) {
    const int K = P.ent.K;
    KGE_REQUIRE((size_t)span_warps(K) * K * sizeof(float) <= 200 * 1024, "kge_train: embedding size %d too large for the span reduction", K);
    if (K % 4 == 0) {
        if (K <= 128) return launch_apply_nca<4, 1>(P, tmode, sm_count, st, mid, pre);
        if (K <= 256) return launch_apply_nca<4, 2>(P, tmode, sm_count, st, mid, pre);
        if (K <= 512) return launch_apply_nca<4, 4>(P, tmode, sm_count, st, mid, pre);
        return launch_apply_nca<4, 0>(P, tmode, sm_count, st, mid, pre);
    }
    return launch_apply_nca<1, 0>(P, tmode, sm_count, st, mid, pre);
}

// packed_in: n_items (key << 32 | global slot) entries, unsorted; sorted by key (stable) into ctx->ks_sorted
static int sort_impl(kge_ctx* ctx, const kge_train_args* a, const uint64_t* packed_in, int64_t n_items, cudaStream_t st) {
    KGE_REQUIRE(n_items <= (int64_t)KGE_SLOT_MASK, "kge_train_apply: too many slots");
    if (n_items == 0) return 0;
    const int K = a->ent.K;
    const int64_t n_chunks = (n_items + KGE_CH - 1) / KGE_CH;
    if (ctx->ks_sorted.reserve((size_t)n_items * 8)) return -2;
    if (ctx->partial.reserve((size_t)2 * n_chunks * K * sizeof(float))) return -2;
    {   // header {span heads, hubs, ticket, pad} + the two lists; the header is zeroed once per allocation (kge_span_apply_kernel
        // leaves it zero) on the stream the sort runs on, which the reduction waits for
        const void* before = ctx->span_head.p;
        if (ctx->span_head.reserve((size_t)(2 * n_chunks + 4) * sizeof(int32_t))) return -2;
        if (ctx->span_head.p != before) KGE_CUDA_CHECK(cudaMemsetAsync(ctx->span_head.p, 0, 4 * sizeof(int32_t), st));
    }
    int64_t E = a->ent.rows;
    // small batch over a small key range: single-launch stable radix sort with CTAs small enough to run beside the
    // forward/backward kernel (kge_sort_small.cu); same output, bit for bit
    if (kge_small_sort_ok(n_items, E + a->R)) return kge_small_sort(ctx, packed_in, n_items, E + a->R, ctx->ks_sorted.as<uint64_t>(), st);
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < E + a->R) ++end_bit;
    size_t tmp_bytes = 0;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, packed_in, ctx->ks_sorted.as<uint64_t>(), (int)n_items, 32,
                                                  32 + end_bit, st));
    if (ctx->sort_tmp.reserve(tmp_bytes)) return -2;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tmp_bytes, packed_in, ctx->ks_sorted.as<uint64_t>(), (int)n_items,
                                                  32, 32 + end_bit, st));
    return 0;
}

static double adam_lr_t(const kge_train_args* a) {
    const bool reset = (a->flags & KGE_F_RESET_STATE) != 0;
    double t = reset ? 1.0 : (double)(a->step < 1 ? 1 : a->step);
    return (double)a->lr * sqrt(1.0 - pow((double)a->beta2, t)) / (1.0 - pow((double)a->beta1, t));
}

static int reduce_impl(kge_ctx* ctx, const kge_train_args* a, int64_t n_items, const kge_table* grads, int64_t row_begin,
                       int64_t row_end, cudaStream_t st, const KgeStepDyn* dyn = nullptr) {
    if (n_items == 0) return 0;
    const int K = a->ent.K;
    const int64_t E = a->ent.rows;
    ApplyParams P;
    P.ks = ctx->ks_sorted.as<uint64_t>();
    P.n_keys = n_items;
    P.G.n_ranks = grads->n_shards;
    P.G.S = grads->rows_per_shard;
    P.G.n = grads->rows_per_shard / (3 + a->eta);
    for (int i = 0; i < KGE_MAX_SHARDS; ++i) {
        P.G.base[i] = grads->shard[i];
        P.G.tail[i] = grads->shard[i];
        if (a->grad_tails != nullptr && i < grads->n_shards)  // never dereferenced below the tail
            P.G.tail[i] = a->grad_tails + (int64_t)i * a->grad_tail_stride - 3 * P.G.n * (int64_t)K;
    }
    P.G.eta = a->eta;
    P.G.K = K;
    P.ent = make_view(a->ent);
    P.ent_m = make_view(a->ent_m);
    P.ent_v = make_view(a->ent_v);
    P.has_m = table_present(a->ent_m);
    P.has_v = table_present(a->ent_v);
    P.rel = a->rel;
    P.rel_m = a->rel_m;
    P.rel_v = a->rel_v;
    P.E = E;
    P.R = a->R;
    P.row_begin = row_begin;
    P.row_end = row_end;
    P.opt = a->opt;
    P.flags = a->flags;
    P.lr = a->lr;
    P.beta1 = a->beta1;
    P.beta2 = a->beta2;
    P.eps = a->eps;
    P.momentum = a->momentum;
    P.partial = ctx->partial.as<float>();
    {
        const int64_t n_chunks = (n_items + KGE_CH - 1) / KGE_CH;
        P.span_count = ctx->span_head.as<int32_t>();
        P.span_ticket = ctx->span_head.as<int32_t>() + 2;
        P.span_list = ctx->span_head.as<int32_t>() + 4;
        P.hub_list = P.span_list + n_chunks;
    }
    // dense batch (on average >= 8 slots per table row): most runs cross chunk borders -> thousands of short spans
    P.span_use_hubs = (n_items >= 8 * (E + a->R)) ? 1 : 0;
    if (const char* e = getenv("KGE_SPAN_WARP")) P.span_use_hubs = (e[0] == '1') ? 1 : 0;
    P.dbg_grad_ent = a->dbg_grad_ent;
    P.dbg_grad_rel = a->dbg_grad_rel;
    P.reg_p = a->reg_p;
    P.reg_lambda_ent = a->reg_lambda_ent;
    P.reg_lambda_rel = a->reg_lambda_rel;
    P.touched = nullptr;
    {   // KGE_APPLY_PREFETCH=1 switches the short-distance L2 prefetch of the narrow-row reduction on (A/B knob, off by default).
        // Measured on B200 (profiles/r02_summary.md): prefetching the optimizer rows of a whole 16-slot chunk thrashes L2 (DRAM
        // reads double, cfg5 reduce_apply 0.70 -> 0.86 ms) and a 2-3 slot look-ahead changes nothing: these kernels are bound by
        // the DRAM request rate of random 128-byte rows and by instruction issue, not by exposed latency
        static int pf = -1;
        if (pf < 0) {
            const char* e = getenv("KGE_APPLY_PREFETCH");
            pf = (e != nullptr && e[0] == '1') ? 1 : 0;
        }
        P.prefetch = pf;
        static int pfw = -1;  // KGE_APPLY_PREFETCH_WIDE=1: one-run-ahead prefetch in the warp-per-chunk kernel too (A/B)
        if (pfw < 0) {
            const char* e = getenv("KGE_APPLY_PREFETCH_WIDE");
            pfw = (e != nullptr && e[0] == '1') ? 1 : 0;
        }
        P.prefetch_wide = pfw;
    }
    const bool reg = a->reg_p > 0 && (a->reg_lambda_ent != 0.f || a->reg_lambda_rel != 0.f);
    if (!reg) P.reg_p = 0;
    if (reg) {
        // penalty of the pre-update parameters of this rank's rows (+ the replicated relation table, counted
        // by the rank that owns row 0), then the touched-row bitmap for the dense pass
        const int nb = ctx->sm_count * 4;
        const size_t words = (size_t)((E + a->R + 31) / 32);
        if (ctx->reg_partial.reserve((size_t)2 * nb * sizeof(double)) || ctx->touched.reserve(words * sizeof(uint32_t))) return -2;
        const int64_t rps = a->ent.rows_per_shard > 0 ? a->ent.rows_per_shard : a->ent.rows;
        const int own = a->ent.n_shards == 1 ? 0 : (int)(row_begin / rps);
        double* part = ctx->reg_partial.as<double>();
        KGE_CUDA_CHECK(cudaMemsetAsync(part, 0, (size_t)2 * nb * sizeof(double), st));
        if (row_end > row_begin)
            kge_reg_loss_kernel<<<nb, 256, 0, st>>>(a->ent.shard[own] + (row_begin - (int64_t)own * rps) * K, (row_end - row_begin) * (int64_t)K,
                                                    a->reg_p, a->reg_lambda_ent, part);
        if (row_begin == 0) kge_reg_loss_kernel<<<nb, 256, 0, st>>>(a->rel, a->R * (int64_t)K, a->reg_p, a->reg_lambda_rel, part + nb);
        KGE_CUDA_CHECK(cudaGetLastError());
        KGE_CUDA_CHECK(cudaMemsetAsync(ctx->touched.p, 0, words * sizeof(uint32_t), st));
        P.touched = ctx->touched.as<uint32_t>();
    }
    const bool reset = (a->flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (a->flags & KGE_F_NO_UPDATE) != 0;
    P.lr_t = (float)adam_lr_t(a);
    P.dyn = dyn;
    if (!no_update && !reset) {
        if (a->opt == KGE_OPT_ADAM)
            KGE_REQUIRE(P.has_m && P.has_v && a->rel_m && a->rel_v, "kge_train: adam state (m,v) missing");
        if (a->opt == KGE_OPT_ADAGRAD || a->opt == KGE_OPT_MOMENTUM)
            KGE_REQUIRE(P.has_m && a->rel_m, "kge_train: optimizer state missing");
    }
    const int tmode = a->model == KGE_TRANSE_L1 ? 1 : (a->model == KGE_TRANSE_L2 ? 2 : 0);
    if (int rc = launch_apply(P, tmode, ctx->sm_count, st, ctx->timing ? ctx->tev[3] : nullptr, ctx->timing ? ctx->tev[6] : nullptr)) return rc;
    if (reg) {
        if (K % 4 == 0) kge_reg_dense_kernel<4><<<ctx->sm_count * 8, 256, 0, st>>>(P);
        else kge_reg_dense_kernel<1><<<ctx->sm_count * 8, 256, 0, st>>>(P);
        KGE_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

// adds the LP penalty computed by reduce_impl to the batch loss; `st` must be ordered behind the loss reduction
static int reg_loss_finish(kge_ctx* ctx, const kge_train_args* a, cudaStream_t st) {
    if (!(a->reg_p > 0 && (a->reg_lambda_ent != 0.f || a->reg_lambda_rel != 0.f))) return 0;
    kge_reg_loss_add_kernel<<<1, 32, 0, st>>>(ctx->reg_partial.as<double>(), 2 * ctx->sm_count * 4, a->loss_out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

static int apply_impl(kge_ctx* ctx, const kge_train_args* a, const uint64_t* packed_in, int64_t n_items, const kge_table* grads,
                      int64_t row_begin, int64_t row_end, cudaStream_t st) {
    if (int rc = sort_impl(ctx, a, packed_in, n_items, st)) return rc;
    return reduce_impl(ctx, a, n_items, grads, row_begin, row_end, st);
}

static int ensure_side_stream(kge_ctx* ctx) {
    if (ctx->side != nullptr) return 0;
    // highest priority: the sort's small kernels must not queue behind the waves of the forward kernel
    int prio_lo = 0, prio_hi = 0;
    KGE_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    KGE_CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_hi));
    KGE_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->lstream, cudaStreamNonBlocking));
    KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_sorted, cudaEventDisableTiming));
    KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fwd, cudaEventDisableTiming));
    KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_loss, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_set_free[i], cudaEventDisableTiming));
        KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_pro_emit[i], cudaEventDisableTiming));
        KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_pro_sorted[i], cudaEventDisableTiming));
    }
    return 0;
}

// Selection of the slots this rank reduces (keys in [row_begin,row_end) or relation keys), started
// early so that the count reaches the host while kge_train_fwd_bwd is still running.
extern "C" int kge_train_select(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys, int64_t row_begin,
                                int64_t row_end, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_select: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(keys_all != nullptr || n_keys == 0, "kge_train_select: keys missing");
    KGE_REQUIRE(n_keys < (int64_t)INT32_MAX, "kge_train_select: too many slots");
    cudaStream_t st = (cudaStream_t)stream;
    ctx->sel_valid = false;
    if (n_keys == 0) return 0;
    if (ctx->ks_in.reserve((size_t)n_keys * 8) || ctx->ks_sel.reserve((size_t)n_keys * 8) || ctx->sel_flags.reserve((size_t)n_keys) ||
        ctx->sel_count.reserve(sizeof(int)))
        return -2;
    if (ctx->h_count == nullptr) {
        KGE_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_count, sizeof(int)));
        KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_count, cudaEventDisableTiming));
    }
    int threads = 256;
    int blocks = (int)std::min<int64_t>((n_keys + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
    kge_select_flag_kernel<<<blocks, threads, 0, st>>>(keys_all, n_keys, a->ent.rows, row_begin, row_end, ctx->ks_in.as<uint64_t>(),
                                                       ctx->sel_flags.as<uint8_t>());
    KGE_CUDA_CHECK(cudaGetLastError());
    size_t tb = 0;
    KGE_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, tb, ctx->ks_in.as<uint64_t>(), ctx->sel_flags.as<uint8_t>(),
                                              ctx->ks_sel.as<uint64_t>(), ctx->sel_count.as<int>(), (int)n_keys, st));
    if (ctx->sort_tmp.reserve(tb)) return -2;
    KGE_CUDA_CHECK(cub::DeviceSelect::Flagged(ctx->sort_tmp.p, tb, ctx->ks_in.as<uint64_t>(), ctx->sel_flags.as<uint8_t>(),
                                              ctx->ks_sel.as<uint64_t>(), ctx->sel_count.as<int>(), (int)n_keys, st));
    KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_count, ctx->sel_count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_count, st));
    ctx->sel_valid = true;
    ctx->sel_keys = keys_all;
    ctx->sel_n = n_keys;
    ctx->sel_begin = row_begin;
    ctx->sel_end = row_end;
    return 0;
}

extern "C" int kge_train_apply(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                               const kge_table* grads, int64_t row_begin, int64_t row_end, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_apply: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(grads != nullptr && keys_all != nullptr, "kge_train_apply: missing keys/grads");
    if (n_keys == 0) return 0;
    KGE_REQUIRE(grads->n_shards >= 1 && grads->rows_per_shard > 0 && grads->rows_per_shard % (3 + a->eta) == 0,
                "kge_train_apply: grads.rows_per_shard must be the slots per rank, (3+eta)*n_pos");
    KGE_REQUIRE(n_keys == grads->rows_per_shard * grads->n_shards, "kge_train_apply: n_keys != n_shards * slots per rank");
    if (!(ctx->sel_valid && ctx->sel_keys == keys_all && ctx->sel_n == n_keys && ctx->sel_begin == row_begin && ctx->sel_end == row_end)) {
        if (int rc = kge_train_select(ctx, a, keys_all, n_keys, row_begin, row_end, stream)) return rc;
    }
    KGE_CUDA_CHECK(cudaEventSynchronize(ctx->ev_count));
    ctx->sel_valid = false;
    const int64_t m = *ctx->h_count;
    if (int rc = apply_impl(ctx, a, ctx->ks_sel.as<uint64_t>(), m, grads, row_begin, row_end, (cudaStream_t)stream)) return rc;
    // phased API: kge_train_fwd_bwd reduced the batch loss on this stream already; the LP penalty of this
    // rank's rows is added to a->loss_out here
    return reg_loss_finish(ctx, a, (cudaStream_t)stream);
}

static void timing_collect(kge_ctx* ctx) {
    if (!ctx->tpending) return;
    cudaEventSynchronize(ctx->tev[4]);
    for (int i = 0; i < 4; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->tev[i], ctx->tev[i + 1]) == cudaSuccess) ctx->tacc[i] += ms;
    }
    {   // side stream: end of the radix sort, measured from the end of emit
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->tev[1], ctx->tev[5]) == cudaSuccess) ctx->tacc[4] += ms;
    }
    {   // end of fwd_bwd -> launch of the level-1 reduction kernel: the wait for the sort (+ two memsets)
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->tev[2], ctx->tev[6]) == cudaSuccess) ctx->tacc[5] += ms;
    }
    ctx->tcount += 1;
    ctx->tpending = false;
}

extern "C" int kge_ctx_set_timing(kge_ctx* ctx, int on) {
    KGE_REQUIRE(ctx != nullptr, "kge_ctx_set_timing: null ctx");
    if (on && ctx->tev[0] == nullptr)
        for (int i = 0; i < 7; ++i) KGE_CUDA_CHECK(cudaEventCreate(&ctx->tev[i]));
    timing_collect(ctx);
    ctx->timing = on != 0;
    for (int i = 0; i < 6; ++i) ctx->tacc[i] = 0;
    ctx->tcount = 0;
    return 0;
}

extern "C" int kge_ctx_get_timing(kge_ctx* ctx, float* ms_out5, int* steps_out) {
    KGE_REQUIRE(ctx != nullptr && ms_out5 != nullptr, "kge_ctx_get_timing: null argument");
    timing_collect(ctx);
    for (int i = 0; i < 5; ++i) ms_out5[i] = ctx->tcount ? (float)(ctx->tacc[i] / ctx->tcount) : 0.f;
    if (steps_out) *steps_out = ctx->tcount;
    return 0;
}

// test hook: the sort of the (key << 32 | slot) entries alone.  algo 0: cub::DeviceRadixSort on the key bits (stable); 1: the
// single-launch sort of kge_sort_small.cu (fails when the size is outside its range)
extern "C" int kge_sort_entries(kge_ctx* ctx, const uint64_t* in, int64_t n, int64_t n_keys, int algo, uint64_t* out, void* stream) {
    KGE_REQUIRE(ctx != nullptr && in != nullptr && out != nullptr && n >= 0 && n_keys > 0, "kge_sort_entries: bad argument");
    KGE_REQUIRE(n <= (int64_t)KGE_SLOT_MASK, "kge_sort_entries: too many entries");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (algo == 1) {
        KGE_REQUIRE(kge_small_sort_ok(n, n_keys), "kge_sort_entries: %lld entries / %lld keys is outside the small sort's range",
                    (long long)n, (long long)n_keys);
        return kge_small_sort(ctx, in, n, n_keys, out, st);
    }
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < n_keys) ++end_bit;
    size_t tmp_bytes = 0;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, (int)n, 32, 32 + end_bit, st));
    if (ctx->sort_tmp.reserve(tmp_bytes)) return -2;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tmp_bytes, in, out, (int)n, 32, 32 + end_bit, st));
    return 0;
}

// the same with the wait for the sort split off: {emit, fwd_bwd, sort wait, level-1 reduction kernel, span/hub reduction,
// end of the sort measured from the end of emit}
extern "C" int kge_ctx_get_timing_ex(kge_ctx* ctx, float* ms_out, int n_out, int* steps_out) {
    KGE_REQUIRE(ctx != nullptr && ms_out != nullptr && n_out >= 6, "kge_ctx_get_timing_ex: need room for 6 phases");
    timing_collect(ctx);
    const double n = ctx->tcount ? (double)ctx->tcount : 1.0;
    ms_out[0] = (float)(ctx->tacc[0] / n);
    ms_out[1] = (float)(ctx->tacc[1] / n);
    ms_out[2] = (float)(ctx->tacc[5] / n);
    ms_out[3] = (float)((ctx->tacc[2] - ctx->tacc[5]) / n);
    ms_out[4] = (float)(ctx->tacc[3] / n);
    ms_out[5] = (float)(ctx->tacc[4] / n);
    if (steps_out) *steps_out = ctx->tcount;
    return 0;
}

#define KGE_TMARK(i)                                                        \
    do {                                                                    \
        if (ctx->timing) KGE_CUDA_CHECK(cudaEventRecord(ctx->tev[i], st));  \
    } while (0)

static int train_step_body(kge_ctx* ctx, const kge_train_args* a, void* stream, const KgeStepDyn* dyn) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_step: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(a->ent.n_shards == 1, "kge_train_step is the single-GPU entry; use the phased calls when sharded");
    if (a->n_pos == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = a->ent.K;
    int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    if (ctx->ks_in.reserve((size_t)S * 8)) return -2;
    if (ctx->grad_rows.reserve((size_t)gbuf_floats(a->eta, a->n_pos, K) * sizeof(float))) return -2;
    if (int rc = ensure_side_stream(ctx)) return rc;
    if (ctx->timing) timing_collect(ctx);
    KGE_TMARK(0);
    if (int rc = emit_impl(ctx, a, nullptr, ctx->ks_in.as<uint64_t>(), st, dyn)) return rc;
    KGE_TMARK(1);
    // fork: the radix sort only needs the keys, so it runs beside the forward/backward kernel
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, st));
    KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    if (int rc = sort_impl(ctx, a, ctx->ks_in.as<uint64_t>(), S, ctx->side)) return rc;
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_sorted, ctx->side));
    if (ctx->timing) KGE_CUDA_CHECK(cudaEventRecord(ctx->tev[5], ctx->side));
    if (int rc = fwd_bwd_impl(ctx, a, ctx->grad_rows.as<float>(), st, ctx->side)) return rc;
    KGE_TMARK(2);
    kge_table g;
    memset(&g, 0, sizeof(g));
    g.shard[0] = ctx->grad_rows.as<float>();
    g.rows = S;
    g.rows_per_shard = S;
    g.n_shards = 1;
    g.K = K;
    KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_sorted, 0));
    if (int rc = reduce_impl(ctx, a, S, &g, 0, a->ent.rows, st, dyn)) return rc;
    KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_loss, 0));  // join
    if (int rc = reg_loss_finish(ctx, a, st)) return rc;
    if (ctx->timing) {
        KGE_CUDA_CHECK(cudaEventRecord(ctx->tev[4], st));
        ctx->tpending = true;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Software-pipelined step (KGE_F_PIPELINE).  emit (corruptions + sort keys) and the radix sort read the batch
// and the (seed, step) counters only, so the step is cut in two:
//   prologue : emit + sort on the side stream into the buffer set {repl, keep, ks_in, ks_sorted} that the step
//              before the previous one used (the two sets alternate); it waits for that step's end
//              (ev_set_free) and, for host batches, for the batch copy -- not for the previous step
//   main     : forward/backward, loss, segmented reduction + optimizer on the caller's stream, behind the
//              previous step as always, and behind the prologue's two events
// Submitting step t+1 while step t runs (any asynchronous caller does) therefore overlaps prologue(t+1) with
// main(t): emit and the sort leave the critical path.  Same kernels, same inputs, same order of every
// floating-point sum => bit-identical to the in-order step.  KGE_PIPELINE=0 forces the in-order step (A/B).
// ------------------------------------------------------------------------------------------------
static inline bool pipeline_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_PIPELINE");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

static inline bool step_is_pipelined(const kge_ctx* ctx, const kge_train_args* a) {
    return (a->flags & KGE_F_PIPELINE) != 0 && !ctx->timing && a->ent.n_shards == 1 && a->n_pos > 0 && pipeline_enabled();
}

// every single-GPU step: the buffer set it read is free again once `st` gets here
static int mark_set_free(kge_ctx* ctx, cudaStream_t st) {
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_set_free[ctx->set_id], st));
    return 0;
}

static int pipeline_prologue(kge_ctx* ctx, const kge_train_args* a, cudaEvent_t batch_ready) {
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(a->ent.n_shards == 1, "kge_train_step is the single-GPU entry; use the phased calls when sharded");
    if (int rc = ensure_side_stream(ctx)) return rc;
    std::swap(ctx->repl, ctx->alt_repl);
    std::swap(ctx->keep, ctx->alt_keep);
    std::swap(ctx->ks_in, ctx->alt_ks_in);
    std::swap(ctx->ks_sorted, ctx->alt_ks_sorted);
    ctx->set_id ^= 1;
    const int sid = ctx->set_id;
    const int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    if (ctx->ks_in.reserve((size_t)S * 8)) return -2;
    if (ctx->grad_rows.reserve((size_t)gbuf_floats(a->eta, a->n_pos, a->ent.K) * sizeof(float))) return -2;
    if (batch_ready != nullptr) KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->side, batch_ready, 0));
    KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->side, ctx->ev_set_free[sid], 0));  // no-op until first recorded
    if (int rc = emit_impl(ctx, a, nullptr, ctx->ks_in.as<uint64_t>(), ctx->side, nullptr)) return rc;
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_pro_emit[sid], ctx->side));
    if (int rc = sort_impl(ctx, a, ctx->ks_in.as<uint64_t>(), S, ctx->side)) return rc;
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_pro_sorted[sid], ctx->side));
    return 0;
}

// wait_inside: enqueue the waits for the prologue here (eager step); false when this is captured into a graph
// (the caller then waits on the launching stream, outside the graph)
static int pipeline_main(kge_ctx* ctx, const kge_train_args* a, cudaStream_t st, const KgeStepDyn* dyn, bool wait_inside) {
    const int sid = ctx->set_id;
    const int K = a->ent.K;
    const int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    if (wait_inside) KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_pro_emit[sid], 0));
    // the loss reduction goes to its own stream: on the side stream it would sit behind ev_fwd(t) and hold back the
    // prologue (emit + sort) of step t+1, which is queued on that in-order stream
    if (int rc = fwd_bwd_impl(ctx, a, ctx->grad_rows.as<float>(), st, ctx->lstream)) return rc;
    kge_table g;
    memset(&g, 0, sizeof(g));
    g.shard[0] = ctx->grad_rows.as<float>();
    g.rows = S;
    g.rows_per_shard = S;
    g.n_shards = 1;
    g.K = K;
    if (wait_inside) KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_pro_sorted[sid], 0));
    if (int rc = reduce_impl(ctx, a, S, &g, 0, a->ent.rows, st, dyn)) return rc;
    KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_loss, 0));  // join
    return reg_loss_finish(ctx, a, st);
}

extern "C" int kge_train_step(kge_ctx* ctx, const kge_train_args* a, void* stream) {
    KGE_REQUIRE(ctx != nullptr && a != nullptr, "kge_train_step: null argument");
    if (step_is_pipelined(ctx, a)) {
        if (int rc = pipeline_prologue(ctx, a, nullptr)) return rc;
        if (int rc = pipeline_main(ctx, a, (cudaStream_t)stream, nullptr, true)) return rc;
        return mark_set_free(ctx, (cudaStream_t)stream);
    }
    if (int rc = train_step_body(ctx, a, stream, nullptr)) return rc;
    return a->n_pos > 0 ? mark_set_free(ctx, (cudaStream_t)stream) : 0;
}

// KGE_GRAPH=0 disables the captured-graph replay of the host-buffer step
static inline bool train_graph_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_GRAPH");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// The step as one graph launch.  Everything in `b` except the step counter is part of the key; the
// first call with a new key runs eagerly (workspace growth, function attributes), the second is
// captured (the side-stream fork/join included) and instantiated, later ones only replay.
// pipelined: only the main part of the step is captured (the caller has issued the prologue and the waits for it)
static int train_step_graphed(kge_ctx* ctx, const kge_train_args* b, cudaStream_t st, const KgeStepDyn* dd, bool pipelined) {
    static uint64_t tick = 0;
    kge_train_args key = *b;
    key.step = 0;
    // a graph holds the raw pointers of the buffer set that was current at capture
    const int variant = (pipelined ? 10 : 0) + ctx->set_id;
    auto body = [&](const KgeStepDyn* dyn) -> int {
        return pipelined ? pipeline_main(ctx, b, st, dyn, false) : train_step_body(ctx, b, st, dyn);
    };
    KgeGraphEntry* e = nullptr;
    for (KgeGraphEntry& g : ctx->graphs)
        if (g.seen > 0 && g.stream == st && g.variant == variant && memcmp(&g.key, &key, sizeof(key)) == 0) e = &g;
    if (e == nullptr) {
        e = &ctx->graphs[0];
        for (KgeGraphEntry& g : ctx->graphs)
            if (g.last_use < e->last_use) e = &g;
        if (e->exec) {
            KGE_CUDA_CHECK(cudaStreamSynchronize(e->stream));
            cudaGraphExecDestroy(e->exec);
        }
        e->exec = nullptr;
        e->key = key;
        e->stream = st;
        e->variant = variant;
        e->seen = 0;
    }
    e->last_use = ++tick;
    e->seen += 1;
    if (e->seen == 1) return body(nullptr);
    // `dd`: this step's {step counter, lr_t} block in device memory, refreshed by the caller on the copy stream
    // (one block per batch buffer, and b->pos -- part of the key -- names the batch buffer: a graph always
    // reads the same block)
    if (e->exec != nullptr && e->ws_epoch != g_kge_ws_epoch) {  // a workspace buffer moved since the capture
        cudaGraphExecDestroy(e->exec);
        e->exec = nullptr;
    }
    if (e->exec == nullptr) {
        cudaGraph_t graph = nullptr;
        e->ws_epoch = g_kge_ws_epoch;
        KGE_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        int rc = body(dd);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            e->seen = 1;  // stay eager for this key
            if (rc == 0) kge_set_error("kge_train_step_host: stream capture failed (%s)", cudaGetErrorString(ce));
            return body(nullptr);
        }
        ce = cudaGraphInstantiate(&e->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            e->exec = nullptr;
            e->seen = 1;
            cudaGetLastError();
            return body(nullptr);
        }
    }
    KGE_CUDA_CHECK(cudaGraphLaunch(e->exec, st));
    return 0;
}

extern "C" int kge_train_step_host_async(kge_ctx* ctx, const kge_train_args* a, const int32_t* pos_host, float* loss_host,
                                         void* stream, int* ticket_out) {
    KGE_REQUIRE(ctx != nullptr && a != nullptr, "kge_train_step_host: null argument");
    KGE_REQUIRE(a->n_pos >= 0 && (a->n_pos == 0 || pos_host != nullptr), "kge_train_step_host: positives missing");
    cudaStream_t st = (cudaStream_t)stream;
    const int slot = (int)(ctx->host_tick % KGE_HOST_RING);
    if (ticket_out) *ticket_out = slot;
    if (ctx->ev_host[slot] == nullptr) KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_host[slot], cudaEventDisableTiming));
    else KGE_CUDA_CHECK(cudaEventSynchronize(ctx->ev_host[slot]));  // the step that used this ring slot last is done
    if (a->n_pos == 0) {
        if (loss_host) *loss_host = 0.f;
        KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_host[slot], st));
        ctx->host_tick += 1;
        return 0;
    }
    if (ctx->h_loss.reserve(sizeof(float))) return -2;
    const bool graphed = train_graph_enabled() && !ctx->timing && a->ent.n_shards == 1;
    const size_t pos_bytes = (size_t)a->n_pos * 3 * sizeof(int32_t);
    kge_train_args b = *a;
    int pb = -1;  // batch buffer of this step (graphed path)
    bool pipelined = false;
    if (graphed) {
        // the caller's stream may be the legacy default stream, which cannot be captured: the whole call
        // runs on a stream of the ctx, ordered behind the caller's stream (kge_train_host_wait joins it)
        if (ctx->gmain == nullptr) {
            KGE_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->gmain, cudaStreamNonBlocking));
            KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_gin, cudaEventDisableTiming));
            KGE_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
                KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_posfree[i], cudaEventDisableTiming));
            }
        }
        KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_gin, st));
        KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->gmain, ctx->ev_gin, 0));
        st = ctx->gmain;
        // H2D of the batch on the copy stream, into the buffer the step before the previous one used: it
        // overlaps the previous step instead of sitting between two graph launches
        pb = (int)(ctx->host_tick & 1);
        if (ctx->h_pos2[pb].reserve(pos_bytes)) return -2;
        KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->cstream, ctx->ev_posfree[pb], 0));  // no-op until first recorded
        KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pos2[pb].p, pos_host, pos_bytes, cudaMemcpyHostToDevice, ctx->cstream));
        // the step counter / Adam's bias-corrected rate of this step travel the same way, from the pinned ring
        // slot of this host step (free: the step that used it last has been waited for above)
        if (ctx->h_dyn == nullptr) {
            KGE_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_dyn, KGE_HOST_RING * sizeof(KgeStepDyn)));
            if (ctx->d_dyn.reserve(2 * sizeof(KgeStepDyn))) return -2;
        }
        KgeStepDyn* hd = ctx->h_dyn + slot;
        hd->step = b.step;
        hd->lr_t = (float)adam_lr_t(&b);
        KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->d_dyn.as<KgeStepDyn>() + pb, hd, sizeof(KgeStepDyn), cudaMemcpyHostToDevice, ctx->cstream));
        KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_h2d[pb], ctx->cstream));
        KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_h2d[pb], 0));
        b.pos = ctx->h_pos2[pb].as<int32_t>();
        if (b.loss_out == nullptr) b.loss_out = ctx->h_loss.as<float>();
        pipelined = step_is_pipelined(ctx, &b);
        if (pipelined) {
            // emit + sort of this step start as soon as its batch is in, beside the previous step
            if (int rc = pipeline_prologue(ctx, &b, ctx->ev_h2d[pb])) return rc;
            KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_pro_emit[ctx->set_id], 0));
            KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_pro_sorted[ctx->set_id], 0));
        }
    } else {
        if (ctx->h_pos.reserve(pos_bytes)) return -2;
        KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pos.p, pos_host, pos_bytes, cudaMemcpyHostToDevice, st));
        b.pos = ctx->h_pos.as<int32_t>();
    }
    if (b.loss_out == nullptr) b.loss_out = ctx->h_loss.as<float>();
    if (int rc = graphed ? train_step_graphed(ctx, &b, st, ctx->d_dyn.as<KgeStepDyn>() + pb, pipelined) : train_step_body(ctx, &b, st, nullptr)) return rc;
    if (pb >= 0) KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_posfree[pb], st));
    if (int rc = mark_set_free(ctx, st)) return rc;
    if (loss_host) KGE_CUDA_CHECK(cudaMemcpyAsync(loss_host, b.loss_out, sizeof(float), cudaMemcpyDeviceToHost, st));
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_host[slot], st));
    ctx->host_tick += 1;
    return 0;
}

extern "C" int kge_train_host_wait(kge_ctx* ctx, int ticket) {
    KGE_REQUIRE(ctx != nullptr && ticket >= 0 && ticket < KGE_HOST_RING, "kge_train_host_wait: bad ticket %d", ticket);
    if (ctx->ev_host[ticket] != nullptr) KGE_CUDA_CHECK(cudaEventSynchronize(ctx->ev_host[ticket]));
    return 0;
}

extern "C" int kge_train_step_host(kge_ctx* ctx, const kge_train_args* a, const int32_t* pos_host, float* loss_host,
                                   void* stream) {
    int ticket = 0;
    if (int rc = kge_train_step_host_async(ctx, a, pos_host, loss_host, stream, &ticket)) return rc;
    return kge_train_host_wait(ctx, ticket);
}

// ------------------------------------------------------------------------------------------------
// Dimension-sharded multi-GPU step (kge_dim.cuh): the local engine of one rank.  a->ent / a->rel are this
// rank's COLUMN slices ([E,Kc], [R,Kc], n_shards == 1, a->k = columns per half of the slice, a->k_model = the
// whole model's k), a->pos the GLOBAL batch.  Per step:
//   kge_train_partial (first chunk: corruptions + sort keys + radix sort, pipelined like the single-GPU step)
//   -> the caller all-reduces the sums over the ranks -> kge_train_backward -> kge_train_reduce.
// The positives may be cut into chunks [i_begin,i_end) so that the all-reduce of one chunk overlaps the
// kernels of its neighbours; the chunks of a step must be submitted in order and cover [0,n_pos).
// ------------------------------------------------------------------------------------------------
static int dim_params(kge_ctx* ctx, const kge_train_args* a, int64_t i0, int64_t i1, float* sums, DimParams& P) {
    KGE_REQUIRE(a->ent.n_shards == 1 && a->ent.shard[0] != nullptr, "kge_train (dimension-sharded): a->ent must be the local column slice");
    KGE_REQUIRE(0 <= i0 && i0 < i1 && i1 <= a->n_pos, "kge_train (dimension-sharded): bad positive range [%lld,%lld) of %lld",
                (long long)i0, (long long)i1, (long long)a->n_pos);
    KGE_REQUIRE(sums != nullptr, "kge_train (dimension-sharded): sums missing");
    KGE_REQUIRE(a->stage == nullptr && a->grad_tails == nullptr, "kge_train (dimension-sharded): the row-sharded exchange buffers do not apply");
    P.ent = a->ent.shard[0];
    P.rel = a->rel;
    P.pos = a->pos;
    P.repl = ctx->repl.as<int32_t>();
    P.keep = ctx->keep.as<uint8_t>();
    P.n = a->n_pos;
    P.i0 = i0;
    P.i1 = i1;
    P.eta = a->eta;
    P.k = a->k;
    P.K = a->ent.K;
    P.loss = a->loss;
    P.nl = a->non_linearity;
    P.margin = a->margin;
    P.alpha = a->alpha;
    P.scale = a->model == KGE_HOLE ? 2.0f / (float)(a->k_model > 0 ? a->k_model : a->k) : 1.0f;
    P.sums = sums;
    P.gbuf = ctx->grad_rows.as<float>();
    P.loss_part = ctx->loss_part.as<float>();
    P.dbg_scores = a->dbg_scores;
    return 0;
}

static int launch_dim(int phase, const kge_train_args* a, const DimParams& P, cudaStream_t st) {
    switch (a->model) {
        case KGE_TRANSE_L1: return kge_launch_dim_m0(phase, P, st);
        case KGE_TRANSE_L2: return kge_launch_dim_m1(phase, P, st);
        case KGE_DISTMULT: return kge_launch_dim_m2(phase, P, st);
        default: return kge_launch_dim_m3(phase, P, st);
    }
}

extern "C" int kge_train_partial(kge_ctx* ctx, const kge_train_args* a, int64_t i_begin, int64_t i_end, float* sums, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_partial: null ctx");
    if (int rc = validate_train(a)) return rc;
    if (a->n_pos == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (i_begin == 0) {
        // a new step: corruptions + sort keys of the whole global batch, sort beside the phase kernels
        const int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
        if (int rc = ensure_side_stream(ctx)) return rc;
        ctx->dim_pipelined = (a->flags & KGE_F_PIPELINE) != 0 && pipeline_enabled();
        if (ctx->dim_pipelined) {
            if (int rc = pipeline_prologue(ctx, a, nullptr)) return rc;
            KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_pro_emit[ctx->set_id], 0));
        } else {
            if (ctx->ks_in.reserve((size_t)S * 8)) return -2;
            if (ctx->grad_rows.reserve((size_t)gbuf_floats(a->eta, a->n_pos, a->ent.K) * sizeof(float))) return -2;
            // a step that was started but never reduced may still be sorting these keys on the side stream
            KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_sorted, 0));
            if (int rc = emit_impl(ctx, a, nullptr, ctx->ks_in.as<uint64_t>(), st, nullptr)) return rc;
            KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, st));
            KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            if (int rc = sort_impl(ctx, a, ctx->ks_in.as<uint64_t>(), S, ctx->side)) return rc;
            KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_sorted, ctx->side));
        }
    }
    DimParams P;
    if (int rc = dim_params(ctx, a, i_begin, i_end, sums, P)) return rc;
    return launch_dim(1, a, P, st);
}

// phase 1 of the whole batch with the entity rows streamed in sorted order (kge_dim.cuh); sums in the chunk-major layout
extern "C" int kge_train_partial_sorted(kge_ctx* ctx, const kge_train_args* a, int n_chunks, float* sums, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_partial_sorted: null ctx");
    if (int rc = validate_train(a)) return rc;
    if (a->n_pos == 0) return 0;
    KGE_REQUIRE(n_chunks >= 1 && n_chunks <= a->n_pos, "kge_train_partial_sorted: bad chunk count %d", n_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    KGE_REQUIRE(S <= (int64_t)KGE_SLOT_MASK && (int64_t)(1 + a->eta) * a->n_pos < ((int64_t)1 << 32), "kge_train_partial_sorted: batch too large");
    if (int rc = ensure_side_stream(ctx)) return rc;
    ctx->dim_pipelined = (a->flags & KGE_F_PIPELINE) != 0 && pipeline_enabled();
    if (ctx->dim_pipelined) {
        if (int rc = pipeline_prologue(ctx, a, nullptr)) return rc;
        KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_pro_emit[ctx->set_id], 0));
    } else {
        if (ctx->ks_in.reserve((size_t)S * 8)) return -2;
        if (ctx->grad_rows.reserve((size_t)gbuf_floats(a->eta, a->n_pos, a->ent.K) * sizeof(float))) return -2;
        KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_sorted, 0));
        if (int rc = emit_impl(ctx, a, nullptr, ctx->ks_in.as<uint64_t>(), st, nullptr)) return rc;
        KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, st));
        KGE_CUDA_CHECK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        if (int rc = sort_impl(ctx, a, ctx->ks_in.as<uint64_t>(), S, ctx->side)) return rc;
        KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_sorted, ctx->side));
    }
    DimParams P;
    if (int rc = dim_params(ctx, a, 0, a->n_pos, sums, P)) return rc;
    DimChunks C;
    C.base = (uint32_t)(a->n_pos / n_chunks);
    C.extra = (uint32_t)(a->n_pos % n_chunks);
    if (ctx->pos_off.reserve((size_t)a->n_pos * sizeof(uint2))) return -2;
    uint2* po = ctx->pos_off.as<uint2>();
    auto go = [&](int phase) -> int {
        switch (a->model) {
            case KGE_TRANSE_L1: return kge_launch_dim_sorted_m0(phase, P, C, po, ctx->ks_sorted.as<uint64_t>(), S, st);
            case KGE_TRANSE_L2: return kge_launch_dim_sorted_m1(phase, P, C, po, ctx->ks_sorted.as<uint64_t>(), S, st);
            case KGE_DISTMULT: return kge_launch_dim_sorted_m2(phase, P, C, po, ctx->ks_sorted.as<uint64_t>(), S, st);
            default: return kge_launch_dim_sorted_m3(phase, P, C, po, ctx->ks_sorted.as<uint64_t>(), S, st);
        }
    };
    if (int rc = go(3)) return rc;  // queries + positives: independent of the sort
    KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->dim_pipelined ? ctx->ev_pro_sorted[ctx->set_id] : ctx->ev_sorted, 0));
    return go(4);
}

extern "C" int kge_train_backward(kge_ctx* ctx, const kge_train_args* a, int64_t i_begin, int64_t i_end, const float* sums, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_backward: null ctx");
    if (int rc = validate_train(a)) return rc;
    if (a->n_pos == 0) return 0;
    DimParams P;
    if (int rc = dim_params(ctx, a, i_begin, i_end, const_cast<float*>(sums), P)) return rc;
    return launch_dim(2, a, P, (cudaStream_t)stream);
}

extern "C" int kge_train_reduce(kge_ctx* ctx, const kge_train_args* a, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_reduce: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(a->loss_out != nullptr, "kge_train_reduce: loss_out missing");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_pos == 0) {
        KGE_CUDA_CHECK(cudaMemsetAsync(a->loss_out, 0, sizeof(float), st));
        return 0;
    }
    KGE_REQUIRE(a->ent.n_shards == 1, "kge_train_reduce: a->ent must be the local column slice");
    const int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    kge_loss_reduce_kernel<<<KGE_LOSS_CTAS, 256, 0, st>>>(ctx->loss_part.as<float>(), a->n_pos, a->loss_out, ctx->loss_scr.as<double>(),
                                                         reinterpret_cast<unsigned int*>(ctx->loss_scr.as<double>() + KGE_LOSS_CTAS));
    KGE_CUDA_CHECK(cudaGetLastError());
    kge_table g;
    memset(&g, 0, sizeof(g));
    g.shard[0] = ctx->grad_rows.as<float>();
    g.rows = S;
    g.rows_per_shard = S;
    g.n_shards = 1;
    g.K = a->ent.K;
    KGE_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->dim_pipelined ? ctx->ev_pro_sorted[ctx->set_id] : ctx->ev_sorted, 0));
    if (int rc = reduce_impl(ctx, a, S, &g, 0, a->ent.rows, st, nullptr)) return rc;
    if (int rc = reg_loss_finish(ctx, a, st)) return rc;
    return mark_set_free(ctx, st);
}

extern "C" int kge_normalize_rows(kge_ctx* ctx, float* emb, int64_t rows, int K, void* stream) {
    KGE_REQUIRE(ctx != nullptr && emb != nullptr, "kge_normalize_rows: null argument");
    if (rows == 0) return 0;
    const int warps = 8;
    kge_normalize_rows_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(emb, rows, K);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}
