// Segmented reduction + sparse row-wise optimizer for NARROW rows (K <= 64 floats): the column slices of the
// dimension-sharded multi-GPU step (kge_dim.cuh) and small embeddings on one GPU.
//
// kge_reduce_apply_kernel (kge_train.cu) gives a whole warp to a chunk of KGE_CH sorted slots; with a 128-byte
// row only 8 of its 32 lanes would carry data.  Here a GROUP of GS lanes (8 or 16) owns a chunk, so a warp
// walks 32/GS chunks at once and every load instruction fetches 32/GS rows.
//
// Two measured facts shape the kernel (ncu, profiles/r02_c_summary.md): (1) walking a chunk "run by run" makes the
// groups of a warp diverge -- runs differ in length -- and the warp issues every instruction once per group; (2) the
// per-slot control work (slot decoding with 64-bit divisions, run bookkeeping, row addressing), done redundantly by
// every lane, cost 130 warp-instructions per slot against 1 load + 4 FMAs of useful work: the kernel was
// instruction-bound at 42 % of HBM.  So:
//   * a lane-PARALLEL prologue writes one 16-byte descriptor per slot into shared memory: source row, coefficient, key and
//     the flags {head, tail, process, complete, span head, partial index} -- every decision of the run rule is a function
//     of the slot's key and of six chunk-level scalars, so each lane decides its own slots independently;
//   * the walk is SLOT-synchronous: in step u every group reads descriptor u (one 128-bit shared load) and does what its
//     flags say -- fetch the row's w, m, v at a run head, add the slot's row, apply the optimizer (or park the partial
//     sum of a run that crosses chunks) at a tail.  Most runs of a large sparse batch are one or two slots long, so nearly
//     every step does all three in every group and the warp stays converged;
//   * the optimizer rows of the run heads KGE_RAG_PF slots ahead are put in flight with L2 prefetches (short distance:
//     the footprint stays far below L2 -- prefetching a whole chunk thrashes it, measured).
//
// Chunk ids, the rule that decides which chunk finishes a run, the per-chunk partial rows of long runs and the
// span/hub kernel that finishes them are exactly those of the warp-per-chunk kernel, so the two are interchangeable
// per launch and the summation order (sorted slot order) is the same: results are bit-identical between them.
#include "kge_apply.cuh"

#define KGE_RAG_THREADS 128
#define KGE_RAG_PF 3  // prefetch distance in slots

#define RAG_HEAD 1u
#define RAG_TAIL 2u
#define RAG_PROCESS 4u
#define RAG_COMPLETE 8u
#define RAG_SPAN 16u
#define RAG_OPEN_START 32u
#define RAG_MODE1 64u  // replacement row of a negative: F(c, Q, r) for the TransE models

struct __align__(16) RagDesc {
    uint32_t src;   // source row of the slot's contribution: float offset / 4 from the gradient buffer's base
    float c;
    int32_t key;
    uint32_t flags;
};

// slot -> (source row offset in floats, coefficient, mode) with 32-bit arithmetic (one gradient buffer: n_ranks == 1)
__device__ __forceinline__ void rag_decode(const GradView& G, uint32_t word, uint32_t n, uint32_t& src_off, float& c, uint32_t& mode1) {
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
        mode1 = RAG_MODE1;
    } else {
        src_off = (2u * n + (t - 2u * n - eta_n)) * K;
    }
}

// resident CTAs asked of the compiler: the walk is a chain of dependent memory round trips per group, and the kernel runs
// faster the more groups are resident (measured: 5 CTAs/SM 0.87 ms, 8 CTAs/SM 0.67 ms on cfg5 / 8 ranks)
#define KGE_RAG_MINCTAS 10
template <int GS, int TMODE, int OPT>
__global__ void __launch_bounds__(KGE_RAG_THREADS, KGE_RAG_MINCTAS) kge_reduce_apply_group_kernel(ApplyParams P) {
    constexpr int V = 4;
    constexpr int GPB = KGE_RAG_THREADS / GS;  // chunks per CTA
    __shared__ RagDesc desc[GPB][2 * KGE_CH];
    __shared__ int32_t skey[GPB][2 * KGE_CH + 1];
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1), gib = threadIdx.x / GS;
    const int64_t w = (int64_t)blockIdx.x * GPB + gib;
    const int64_t b0 = w * KGE_CH;
    const bool live = b0 < P.n_keys;  // surplus groups of the last CTA idle through the (warp-uniform) loop
    const int cnt = live ? (int)min((int64_t)KGE_CH, P.n_keys - b0) : 0;
    const int K = P.ent.K;
    const uint32_t n32 = (uint32_t)P.G.n;

    // ---- keys of the own chunk [0,16) and of the next one [16,32) (candidates for a spill-over run)
    uint32_t my_slot[2 * KGE_CH / GS];
#pragma unroll
    for (int i = 0; i < 2 * KGE_CH / GS; ++i) {
        const int t = lg + GS * i;
        int32_t key = -2;
        my_slot[i] = 0;
        if (live && b0 + t < P.n_keys) {
            const uint64_t kv = P.ks[b0 + t];
            key = (int32_t)(kv >> 32);
            my_slot[i] = (uint32_t)(kv & 0xffffffffu);
        }
        skey[gib][t] = key;
    }
    if (lg == 0) skey[gib][2 * KGE_CH] = -3;
    const int32_t key_prev = (live && b0 > 0) ? (int32_t)(P.ks[b0 - 1] >> 32) : -1;
    const int32_t key_prev2 = (live && b0 > KGE_CH) ? (int32_t)(P.ks[b0 - KGE_CH - 1] >> 32) : -1;
    const int32_t key_next2 = (live && b0 + 2 * KGE_CH < P.n_keys) ? (int32_t)(P.ks[b0 + 2 * KGE_CH] >> 32) : -1;
    __syncwarp();
    // ---- chunk-level scalars of the run rule
    const int32_t key_first = skey[gib][0];
    const int32_t key_next = skey[gib][KGE_CH];  // -2 when there is no next chunk
    const int32_t key_last = cnt > 0 ? skey[gib][cnt - 1] : -4;
    int ext = 0;  // slots at the front of the next chunk that continue this chunk's last run
    while (ext < KGE_CH && skey[gib][KGE_CH + ext] == key_last && key_last >= 0) ++ext;
    const bool first_open = key_first == key_prev;                       // the first run continues a run of chunk w-1
    const bool last_is_first = key_last == key_first;                    // one run covers the whole chunk
    const bool last_open_start = last_is_first && first_open;
    const bool last_reaches_next = cnt == KGE_CH && key_last == key_next;
    // a run that starts here and ends inside the next chunk is finished here (same rule as the warp-per-chunk kernel)
    const bool last_spills = last_reaches_next && !last_open_start && key_next2 != key_last;
    const int lim = last_spills ? cnt + ext : cnt;

    // ---- lane-parallel descriptors: every decision is a function of the slot's key and the scalars above
#pragma unroll
    for (int i = 0; i < 2 * KGE_CH / GS; ++i) {
        const int t = lg + GS * i;
        if (t >= lim) continue;
        const int32_t key = skey[gib][t];
        uint32_t src, mode1;
        float c;
        rag_decode(P.G, my_slot[i], n32, src, c, mode1);
        const bool head = (t == 0) || (key != skey[gib][t - 1]);
        const bool tail = (t + 1 == lim) || (skey[gib][t + 1] != key);
        const bool open_start = first_open && key == key_first;
        const bool open_end = last_reaches_next && !last_spills && key == key_last;
        const bool is_rel = key >= P.E;
        const bool owned = is_rel || ((int64_t)key >= P.row_begin && (int64_t)key < P.row_end);
        const bool process = owned && !(open_start && !open_end && key_prev2 != key);  // else the head's group (chunk w-1) reduces it
        const bool complete = !open_start && !open_end;
        uint32_t f = mode1;
        f |= head ? RAG_HEAD : 0u;
        f |= tail ? RAG_TAIL : 0u;
        f |= process ? RAG_PROCESS : 0u;
        f |= complete ? RAG_COMPLETE : 0u;
        f |= (head && process && !open_start && open_end) ? RAG_SPAN : 0u;
        f |= open_start ? RAG_OPEN_START : 0u;
        RagDesc d;
        d.src = src >> 2;
        d.c = c;
        d.key = key;
        d.flags = f;
        desc[gib][t] = d;
    }
    __syncwarp();

    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && OPT != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && OPT == KGE_OPT_ADAM;
    const float lr_t = P.dyn != nullptr ? P.dyn->lr_t : P.lr_t;
    const int cc = min(lg * V, K - V);  // lanes past the end of the row read a valid duplicate and never store
    const bool col_ok = lg * V < K;
    const bool pf = prefetch_on(P) && !no_update;  // off unless KGE_APPLY_PREFETCH=1: measured no gain on B200
    const float* gbase = P.G.base[0];
    // table bases of this lane's prefetch duty (lane 0: w, lane 1: m, lane 2: v)
    const int n_lines = (K + 31) / 32;

    float g[V], rc[V], mv[V], vv[V];
#pragma unroll
    for (int x = 0; x < V; ++x) g[x] = rc[x] = mv[x] = vv[x] = 0.f;
    const int32_t E32 = (int32_t)P.E;
    float* const ent_w = P.ent.shard[0];
    float* const ent_m = P.ent_m.shard[0];
    float* const ent_v = P.ent_v.shard[0];
    const bool st_m_ent = OPT != KGE_OPT_SGD && P.has_m, st_m_rel = OPT != KGE_OPT_SGD && P.rel_m != nullptr;
    const bool st_v_ent = OPT == KGE_OPT_ADAM && P.has_v, st_v_rel = OPT == KGE_OPT_ADAM && P.rel_v != nullptr;

    // The body is written branch-free (selects and predicated loads / stores; the optimizer math runs every step on copies):
    // the groups of a warp are at different points of their runs, and every divergent branch would be issued once per group.
    for (int u = 0; __any_sync(0xffffffffu, u < lim); ++u) {
        if (u >= lim) continue;
        const RagDesc d = desc[gib][u];
        const uint32_t f = d.flags;
        const bool head = (f & RAG_HEAD) != 0, proc = (f & RAG_PROCESS) != 0, comp = (f & RAG_COMPLETE) != 0, tail = (f & RAG_TAIL) != 0;
        const bool rel = d.key >= E32;
        const size_t off = (size_t)(uint32_t)(rel ? d.key - E32 : d.key) * (uint32_t)K + cc;
        float* const pw = (rel ? P.rel : ent_w) + off;
        float* const pm = (rel ? P.rel_m : ent_m) + off;
        float* const pv = (rel ? P.rel_v : ent_v) + off;
        if (pf && u + KGE_RAG_PF < lim && lg < 3 * n_lines) {
            // optimizer row of a run head KGE_RAG_PF slots ahead -> L2 (lane = tensor * n_lines + line)
            const RagDesc dn = desc[gib][u + KGE_RAG_PF];
            if ((dn.flags & (RAG_HEAD | RAG_PROCESS | RAG_COMPLETE)) == (RAG_HEAD | RAG_PROCESS | RAG_COMPLETE)) {
                const int tsel = lg / n_lines, line = lg - tsel * n_lines;
                const bool reln = dn.key >= E32;
                const int64_t row = reln ? dn.key - E32 : dn.key;
                const float* base = tsel == 0 ? (reln ? P.rel : ent_w) : (tsel == 1 ? (reln ? P.rel_m : ent_m) : (reln ? P.rel_v : ent_v));
                if (tsel == 0 || (tsel == 1 && need_m) || (tsel == 2 && need_v)) prefetch_l2(base + row * K + line * 32);
            }
        }
        // ---- run head: fresh sum; the row's w, m, v (predicated loads; the registers keep the previous run's values otherwise)
#pragma unroll
        for (int x = 0; x < V; ++x) g[x] = head ? 0.f : g[x];
        const bool hp = head && proc;
        if (hp && (comp || TMODE != 0)) ldg_vec4_cs(rc, pw);  // touched once per step: streaming (evict-first)
        if (hp && comp && need_m) ldg_vec4_cs(mv, pm);
        if (hp && comp && need_v) ldg_vec4_cs(vv, pv);
        if (hp && lg == 0 && ((f & RAG_SPAN) != 0 || (comp && P.touched != nullptr))) {  // rare
            if (f & RAG_SPAN) P.span_list[atomicAdd(P.span_count, 1)] = (int32_t)w;
            if (comp) mark_touched(P, d.key);
        }
        // ---- the slot's contribution
        float v0[V];
#pragma unroll
        for (int x = 0; x < V; ++x) v0[x] = 0.f;
        if (proc) ldg_vec<V>(v0, gbase + ((size_t)d.src << 2) + cc);
        add_slot<V, TMODE>(g, v0, d.c, (f & RAG_MODE1) ? 1 : 0, rc);
        // ---- run tail: optimizer on copies (computed every step, stored at a complete tail), or the partial sum parked
        const bool tp = tail && proc && col_ok;
        if (tp && !comp) st_vec<V>(P.partial + ((size_t)(2 * w + ((f & RAG_OPEN_START) ? 0 : 1))) * K + cc, g);
        float gg[V], w2[V], m2[V], v2[V];
#pragma unroll
        for (int x = 0; x < V; ++x) {
            gg[x] = g[x];
            w2[x] = rc[x];
            m2[x] = mv[x];
            v2[x] = vv[x];
        }
        if (P.reg_p > 0 || P.dbg_grad_ent != nullptr || P.dbg_grad_rel != nullptr) {  // LP regulariser / parity tests: uniform, normally false
            if (tp && comp) {
                float* dbg = rel ? P.dbg_grad_rel : P.dbg_grad_ent;
                reg_add<V>(P, rel, gg, rc);
                if (dbg != nullptr) st_vec<V>(dbg + (size_t)(rel ? d.key - E32 : d.key) * K + cc, gg);
            }
        }
        opt_math_t<V, OPT>(P, reset, lr_t, gg, w2, m2, v2);
        const bool st = tp && comp && !no_update;
        if (st && (rel ? st_m_rel : st_m_ent)) stg_vec4_cs(pm, m2);
        if (st && (rel ? st_v_rel : st_v_ent)) stg_vec4_cs(pv, v2);
        if (st) stg_vec4_cs(pw, w2);
    }
}

template <int GS, int TMODE>
static int launch_group_opt(const ApplyParams& P, cudaStream_t st) {
    constexpr int GPB = KGE_RAG_THREADS / GS;
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    dim3 grid((unsigned)((n_chunks + GPB - 1) / GPB)), block(KGE_RAG_THREADS);
    // KGE_APPLY_MAXCTAS=m (A/B knob): at most m resident CTAs per SM (unused dynamic shared memory does the limiting), so that
    // the sort of the NEXT step, which a pipelined caller runs beside this kernel, finds room on every SM
    static long pad = -1;
    if (pad < 0) {
        const char* e = getenv("KGE_APPLY_MAXCTAS");
        const int m = (e != nullptr && e[0] >= '1' && e[0] <= '9') ? (e[0] - '0') : 0;
        pad = m > 0 ? (long)(233472 / m - 1024 - 12 * 1024) : 0;
    }
    const size_t dsm = pad > 0 ? (size_t)pad : 0;
    auto launch = [&](auto kern) -> int {
        if (dsm > 36 * 1024) KGE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
        kern<<<grid, block, dsm, st>>>(P);
        return 0;
    };
    int rc;
    switch (P.opt) {
        case KGE_OPT_ADAM: rc = launch(kge_reduce_apply_group_kernel<GS, TMODE, KGE_OPT_ADAM>); break;
        case KGE_OPT_ADAGRAD: rc = launch(kge_reduce_apply_group_kernel<GS, TMODE, KGE_OPT_ADAGRAD>); break;
        case KGE_OPT_MOMENTUM: rc = launch(kge_reduce_apply_group_kernel<GS, TMODE, KGE_OPT_MOMENTUM>); break;
        default: rc = launch(kge_reduce_apply_group_kernel<GS, TMODE, KGE_OPT_SGD>); break;
    }
    if (rc) return rc;
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int GS>
static int launch_group(const ApplyParams& P, int tmode, cudaStream_t st) {
    if (tmode == 0) return launch_group_opt<GS, 0>(P, st);
    if (tmode == 1) return launch_group_opt<GS, 1>(P, st);
    return launch_group_opt<GS, 2>(P, st);
}

// can this launch take the narrow-row kernel?  one local gradient buffer, one local table shard, 32-bit slot arithmetic
bool kge_apply_group_ok(const ApplyParams& P) {
    const int K = P.ent.K;
    return K % 4 == 0 && K <= 64 && P.G.n_ranks == 1 && P.ent.n_shards == 1 && P.G.S < ((int64_t)1 << 31) &&
           5 * P.G.n * (int64_t)K < ((int64_t)1 << 32);
}

// level 1 of the reduction; the caller zeroes span_count first and launches the span kernel after
int kge_launch_apply_group(const ApplyParams& P, int tmode, cudaStream_t st) {
    KGE_REQUIRE(kge_apply_group_ok(P), "kge_launch_apply_group: unsupported shape (K=%d)", P.ent.K);
    return P.ent.K <= 32 ? launch_group<8>(P, tmode, st) : launch_group<16>(P, tmode, st);
}
