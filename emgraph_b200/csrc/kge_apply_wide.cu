// Segmented reduction + sparse row-wise optimizer for WIDE rows (64 < K <= 512 floats) on one local gradient buffer:
// the single-GPU step of every benchmark shape and the wide column slices of the dimension-sharded step.
//
// What the warp-per-chunk kernel of kge_train.cu does with such a chunk (ncu source page, profiles/r02_y_*): it walks the
// chunk run by run, and inside a run two slots at a time.  Every run pays one memory round trip for the row's w, m, v and
// one per pair of slots, all dependent: a chunk of 16 slots in ~2.4 runs is a chain of ~10 round trips per warp, the
// warps sit in long-scoreboard stalls (47 % of the samples) and the kernel moves 2.4 TB/s.  70 % of its instructions are
// control: slot decoding, run bookkeeping, constant-bank reads and a run-time optimizer switch per column vector.
//
// This kernel turns the chunk into a flat LOAD LIST first and then streams it:
//   * one warp of the team that owns the chunk builds, lane-parallel (lane t = sorted slot t of the chunk and of the
//     spill-over part of the next one), one 16-byte item per row the team will read: for every run it reduces, the row's
//     w (and m, v when the optimizer reads them) followed by the run's slots in sorted order.  The run rule -- which
//     chunk finishes a run, which runs are parked as per-chunk partial rows for the span kernels -- is the one of
//     kge_reduce_apply_kernel, evaluated per lane from the slot's key and six chunk-level scalars (as in
//     kge_apply_group.cu); item positions come from one warp prefix sum;
//   * the team (CS warps, each owning 1/CS of the columns, one float4 per lane) walks the list UB items at a time: UB
//     unconditional vector loads in flight, then the items in order -- a state item is kept, a slot item is added to the
//     run's sum, a tail item applies the optimizer (compile-time OPT) or parks the partial sum.  All control is
//     warp-uniform and read from shared memory; a chunk is ~3 round trips instead of ~10.
// Summation order inside a run is the sorted slot order, the optimizer formula is opt_math_t: the results are
// bit-identical to the other two reduction kernels, and chunk ids / partial rows / span_list are theirs, so the span
// kernels that finish long runs do not know which level-1 kernel ran.
#include "kge_apply.cuh"

#define KGE_WD_WARPS 4
#define KGE_WD_MAXITEMS 96  // <= 32 slots + 3 state rows for each of <= 16 runs, padded to a multiple of UB

#define WI_KIND 3u  // 0 slot, 1 w, 2 m, 3 v
#define WI_HEAD 4u
#define WI_TAIL 8u
#define WI_COMPLETE 16u
#define WI_MODE1 32u
#define WI_OPEN_START 64u
#define WI_REL 128u
#define WI_NOP 256u

struct __align__(16) WideItem {
    const float* ptr;  // start of the row (this warp adds its column offset)
    float c;
    uint32_t flags;
};

// slot -> (source row, coefficient, mode) with 32-bit arithmetic (one gradient buffer: n_ranks == 1)
__device__ __forceinline__ void wide_decode(const GradView& G, uint32_t word, uint32_t n, uint32_t& src_off, float& c, uint32_t& mode1) {
    const uint32_t K = (uint32_t)G.K, eta_n = (uint32_t)G.eta * n, t = word & KGE_SLOT_MASK;
    c = 1.f;
    mode1 = 0u;
    if (t < 2u * n) {
        src_off = t * K;
    } else if (t < 2u * n + eta_n) {
        const uint32_t q = t - 2u * n, i = q % n;
        const float* coef = gbuf_coef(G.base[0], n, G.K);
        const uint8_t* keep = gbuf_keep(G.base[0], G.eta, n, G.K);
        c = coef[q];
        const bool kept = (word & KGE_SLOT_HAS_SIDE) ? (word & KGE_SLOT_SIDE) != 0 : keep[q] != 0;
        src_off = ((kept ? 3u : 4u) * n + i) * K;
        mode1 = WI_MODE1;
    } else {
        src_off = (2u * n + (t - 2u * n - eta_n)) * K;
    }
}

// LP regulariser / gradient dump of the parity tests: rare, and reg_grad1's general-p branch is hundreds of instructions
// that the unrolled walk would otherwise carry UB times -- kept out of line (arguments travel in registers)
__device__ __noinline__ float4 wide_extras_impl(int reg_p, float lam, float* dbg, float4 g4, float4 w4) {
    float g[4] = {g4.x, g4.y, g4.z, g4.w};
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
    if (reg_p > 0) {
#pragma unroll
        for (int x = 0; x < 4; ++x) g[x] += reg_grad1(w[x], reg_p, lam);
    }
    if (dbg != nullptr) st_vec<4>(dbg, g);
    return make_float4(g[0], g[1], g[2], g[3]);
}
__device__ __forceinline__ void wide_extras(const ApplyParams& P, bool rel, size_t off, float (&g)[4], const float (&rc)[4]) {
    float* dbg = rel ? P.dbg_grad_rel : P.dbg_grad_ent;
    const float4 r = wide_extras_impl(P.reg_p, rel ? P.reg_lambda_rel : P.reg_lambda_ent, dbg != nullptr ? dbg + off : nullptr,
                                      make_float4(g[0], g[1], g[2], g[3]), make_float4(rc[0], rc[1], rc[2], rc[3]));
    g[0] = r.x; g[1] = r.y; g[2] = r.z; g[3] = r.w;
}

template <int TMODE, int OPT, int CS, int UB>
__global__ void __launch_bounds__(KGE_WD_WARPS * 32, UB >= 8 ? 6 : 8) kge_reduce_apply_wide_kernel(ApplyParams P) {
    constexpr int V = 4;
    constexpr int CPB = KGE_WD_WARPS / CS;  // chunks per CTA
    __shared__ WideItem items[CPB][KGE_WD_MAXITEMS];
    __shared__ int s_n[CPB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int cib = wib / CS, part_id = wib - cib * CS;
    const int64_t w = (int64_t)blockIdx.x * CPB + cib;
    const int64_t b0 = w * KGE_CH;
    const bool live = b0 < P.n_keys;
    const int K = P.ent.K;
    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && OPT != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && OPT == KGE_OPT_ADAM;
    const int32_t E32 = (int32_t)P.E;
    float* const ent_w = P.ent.shard[0];
    float* const ent_m = P.ent_m.shard[0];
    float* const ent_v = P.ent_v.shard[0];
    const float* gbase = P.G.base[0];
    WideItem* const L = items[cib];

    // ---------------- the load list of the chunk: built by the team's first warp, lane t = slot t
    if (part_id == 0) {
        const int cnt = live ? (int)min((int64_t)KGE_CH, P.n_keys - b0) : 0;
        int32_t key = -2;
        uint32_t word = 0;
        if (live && b0 + lane < P.n_keys) {
            const uint64_t kv = P.ks[b0 + lane];
            key = (int32_t)(kv >> 32);
            word = (uint32_t)(kv & 0xffffffffu);
        }
        int32_t key_prev = -1, key_prev2 = -1, key_next2 = -1;
        if (live && lane == 0 && b0 > 0) key_prev = (int32_t)(P.ks[b0 - 1] >> 32);
        if (live && lane == 1 && b0 > KGE_CH) key_prev2 = (int32_t)(P.ks[b0 - KGE_CH - 1] >> 32);
        if (live && lane == 2 && b0 + 2 * KGE_CH < P.n_keys) key_next2 = (int32_t)(P.ks[b0 + 2 * KGE_CH] >> 32);
        key_prev = __shfl_sync(0xffffffffu, key_prev, 0);
        key_prev2 = __shfl_sync(0xffffffffu, key_prev2, 1);
        key_next2 = __shfl_sync(0xffffffffu, key_next2, 2);
        const int32_t key_first = __shfl_sync(0xffffffffu, key, 0);
        const int32_t key_next = __shfl_sync(0xffffffffu, key, KGE_CH);  // -2 when there is no next chunk
        const int32_t key_last = __shfl_sync(0xffffffffu, key, max(cnt - 1, 0));
        const int32_t key_left = __shfl_up_sync(0xffffffffu, key, 1);
        const int32_t key_right = __shfl_down_sync(0xffffffffu, key, 1);
        // slots at the front of the next chunk that continue this chunk's last run
        const int ext = __popc(__ballot_sync(0xffffffffu, lane >= KGE_CH && key == key_last && key >= 0));
        const bool first_open = key_first == key_prev;     // the first run continues a run of chunk w-1
        const bool last_is_first = key_last == key_first;  // one run covers the whole chunk
        const bool last_open_start = last_is_first && first_open;
        const bool last_reaches_next = cnt == KGE_CH && key_last == key_next;
        // a run that starts here and ends inside the next chunk is finished here
        const bool last_spills = last_reaches_next && !last_open_start && key_next2 != key_last;
        const int lim = last_spills ? cnt + ext : cnt;

        const int t = lane;
        const bool in_list = t < lim;
        const bool head = (t == 0) || (key != key_left);
        const bool tail = (t + 1 == lim) || (key_right != key);
        const bool open_start = first_open && key == key_first;
        const bool open_end = last_reaches_next && !last_spills && key == key_last;
        const bool is_rel = key >= E32;
        const bool owned = is_rel || ((int64_t)key >= P.row_begin && (int64_t)key < P.row_end);
        const bool process = in_list && owned && !(open_start && !open_end && key_prev2 != key);  // else chunk w-1 reduces it
        const bool complete = !open_start && !open_end;
        const bool hp = head && process;
        if (hp && !open_start && open_end) P.span_list[atomicAdd(P.span_count, 1)] = (int32_t)w;
        if (hp && complete) mark_touched(P, key);
        int n_state = 0;
        if (hp) n_state = complete ? 1 + (need_m ? 1 : 0) + (need_v ? 1 : 0) : (TMODE != 0 ? 1 : 0);
        const int mine = process ? 1 + n_state : 0;
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (process) {
            int p = incl - mine;
            uint32_t src, mode1;
            float c;
            wide_decode(P.G, word, (uint32_t)P.G.n, src, c, mode1);
            const uint32_t common = (complete ? WI_COMPLETE : 0u) | (open_start ? WI_OPEN_START : 0u) | (is_rel ? WI_REL : 0u);
            if (n_state > 0) {
                const size_t off = (size_t)(uint32_t)(is_rel ? key - E32 : key) * (uint32_t)K;
                WideItem it;
                it.c = 0.f;
                it.ptr = (is_rel ? P.rel : ent_w) + off;
                it.flags = common | 1u | WI_HEAD;
                L[p++] = it;
                if (complete && need_m) {
                    it.ptr = (is_rel ? P.rel_m : ent_m) + off;
                    it.flags = common | 2u;
                    L[p++] = it;
                }
                if (complete && need_v) {
                    it.ptr = (is_rel ? P.rel_v : ent_v) + off;
                    it.flags = common | 3u;
                    L[p++] = it;
                }
            }
            WideItem it;
            it.ptr = gbase + src;
            it.c = c;
            it.flags = common | mode1 | ((head && n_state == 0) ? WI_HEAD : 0u) | (tail ? WI_TAIL : 0u);
            L[p] = it;
        }
        const int padded = (total + UB - 1) / UB * UB;
        if (total + lane < padded) {  // UB <= 32
            WideItem it;
            it.ptr = gbase;
            it.c = 0.f;
            it.flags = WI_NOP;
            L[total + lane] = it;
        }
        if (lane == 0) s_n[cib] = padded;
    }
    if constexpr (CS == 1) __syncwarp();
    else __syncthreads();

    // ---------------- the walk: this warp's share of the columns, one float4 per lane
    const int nvec_row = K / V;
    const int cps = (nvec_row + CS - 1) / CS;  // vectors per column share (<= 32)
    const int v_lo = part_id * cps, v_hi = min(v_lo + cps, nvec_row);
    // lanes past the end of the share read a clamped (valid, duplicate) vector and never store
    const bool col_ok = v_lo + lane < v_hi;
    const int cc = max(min(v_lo + lane, v_hi - 1), 0) * V;
    const float lr_t = P.dyn != nullptr ? P.dyn->lr_t : P.lr_t;
    const bool st_m_ent = OPT != KGE_OPT_SGD && P.has_m, st_m_rel = OPT != KGE_OPT_SGD && P.rel_m != nullptr;
    const bool st_v_ent = OPT == KGE_OPT_ADAM && P.has_v, st_v_rel = OPT == KGE_OPT_ADAM && P.rel_v != nullptr;
    const bool extras = P.reg_p > 0 || P.dbg_grad_ent != nullptr || P.dbg_grad_rel != nullptr;  // LP regulariser / parity tests
    const int n_it = s_n[cib];

    float g[V], rc[V], mv[V], vv[V];
#pragma unroll
    for (int x = 0; x < V; ++x) g[x] = rc[x] = mv[x] = vv[x] = 0.f;
    const float* pw = ent_w;  // row whose w was loaded last (the row the next complete tail updates)
    for (int i0 = 0; i0 < n_it; i0 += UB) {
        float v[UB][V];
#pragma unroll
        for (int q = 0; q < UB; ++q) ldg_vec<V>(v[q], L[i0 + q].ptr + cc);
#pragma unroll
        for (int q = 0; q < UB; ++q) {
            const uint32_t f = L[i0 + q].flags;
            if (f & WI_NOP) continue;
            if (f & WI_HEAD) {
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] = 0.f;
            }
            const uint32_t kind = f & WI_KIND;
            if (kind == 1u) {
                pw = L[i0 + q].ptr;
#pragma unroll
                for (int x = 0; x < V; ++x) rc[x] = v[q][x];
            } else if (kind == 2u) {
#pragma unroll
                for (int x = 0; x < V; ++x) mv[x] = v[q][x];
            } else if (kind == 3u) {
#pragma unroll
                for (int x = 0; x < V; ++x) vv[x] = v[q][x];
            } else {
                add_slot<V, TMODE>(g, v[q], L[i0 + q].c, (f & WI_MODE1) ? 1 : 0, rc);
                if ((f & WI_TAIL) && col_ok) {
                    if (!(f & WI_COMPLETE)) {
                        st_vec<V>(P.partial + ((size_t)(2 * w + ((f & WI_OPEN_START) ? 0 : 1))) * K + cc, g);
                    } else {
                        const bool rel = (f & WI_REL) != 0;
                        const size_t off = (size_t)(pw - (rel ? P.rel : ent_w)) + cc;
                        if (extras) wide_extras(P, rel, off, g, rc);
                        if (!no_update) {
                            opt_math_t<V, OPT>(P, reset, lr_t, g, rc, mv, vv);
                            if (rel ? st_m_rel : st_m_ent) st_vec<V>((rel ? P.rel_m : ent_m) + off, mv);
                            if (rel ? st_v_rel : st_v_ent) st_vec<V>((rel ? P.rel_v : ent_v) + off, vv);
                            st_vec<V>((rel ? P.rel : ent_w) + off, rc);
                        }
                    }
                }
            }
        }
    }
}

template <int TMODE, int OPT, int CS>
static int launch_wide_ub(const ApplyParams& P, cudaStream_t st) {
    constexpr int CPB = KGE_WD_WARPS / CS;
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    dim3 grid((unsigned)((n_chunks + CPB - 1) / CPB)), block(KGE_WD_WARPS * 32);
    // KGE_WIDE_UB=4: four loads in flight per warp and 8 resident CTAs instead of eight loads and 6 (A/B knob, Adam only)
    static int ub = -1;
    if (ub < 0) {
        const char* e = getenv("KGE_WIDE_UB");
        ub = (e != nullptr && atoi(e) == 4) ? 4 : 8;
    }
    if constexpr (OPT == KGE_OPT_ADAM) {
        if (ub == 4) {
            kge_reduce_apply_wide_kernel<TMODE, OPT, CS, 4><<<grid, block, 0, st>>>(P);
            KGE_CUDA_CHECK(cudaGetLastError());
            return 0;
        }
    }
    kge_reduce_apply_wide_kernel<TMODE, OPT, CS, 8><<<grid, block, 0, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int TMODE, int CS>
static int launch_wide_opt(const ApplyParams& P, cudaStream_t st) {
    switch (P.opt) {
        case KGE_OPT_ADAM: return launch_wide_ub<TMODE, KGE_OPT_ADAM, CS>(P, st);
        case KGE_OPT_ADAGRAD: return launch_wide_ub<TMODE, KGE_OPT_ADAGRAD, CS>(P, st);
        case KGE_OPT_MOMENTUM: return launch_wide_ub<TMODE, KGE_OPT_MOMENTUM, CS>(P, st);
        default: return launch_wide_ub<TMODE, KGE_OPT_SGD, CS>(P, st);
    }
}

template <int CS>
static int launch_wide(const ApplyParams& P, int tmode, cudaStream_t st) {
    if (tmode == 0) return launch_wide_opt<0, CS>(P, st);
    if (tmode == 1) return launch_wide_opt<1, CS>(P, st);
    return launch_wide_opt<2, CS>(P, st);
}

// can this launch take the load-list kernel?  16-byte column vectors, one float4 per lane of at most 4 warps, one local
// gradient buffer, one local table shard, 32-bit slot arithmetic
bool kge_apply_wide_ok(const ApplyParams& P) {
    const int K = P.ent.K;
    return K % 4 == 0 && K >= 4 && K <= 512 && P.G.n_ranks == 1 && P.ent.n_shards == 1 && P.G.S < ((int64_t)1 << 31) &&
           5 * P.G.n * (int64_t)K < ((int64_t)1 << 32);
}

// level 1 of the reduction; the caller zeroes span_count first and launches the span kernels after
int kge_launch_apply_wide(const ApplyParams& P, int tmode, cudaStream_t st) {
    KGE_REQUIRE(kge_apply_wide_ok(P), "kge_launch_apply_wide: unsupported shape (K=%d)", P.ent.K);
    const int K = P.ent.K;
    if (K <= 128) return launch_wide<1>(P, tmode, st);
    if (K <= 256) return launch_wide<2>(P, tmode, st);
    return launch_wide<4>(P, tmode, st);
}
