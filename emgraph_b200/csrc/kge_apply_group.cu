// Segmented reduction + sparse row-wise optimizer for NARROW rows (K <= 64 floats): the column slices of the
// dimension-sharded multi-GPU step (kge_dim.cuh) and small embeddings on one GPU.
//
// kge_reduce_apply_kernel (kge_train.cu) gives a whole warp to a chunk of KGE_CH sorted slots; with a 128-byte
// row only 8 of its 32 lanes would carry data.  Here a GROUP of GS lanes (8 or 16) owns a chunk, so a warp
// walks 32/GS chunks at once and every load instruction fetches 32/GS rows.
//
// The groups of a warp hold runs of different lengths, so walking "run by run" makes them diverge and the warp
// issues every instruction once per group (measured: 78 warp-instructions per slot, IPC-bound at 51 % of HBM).
// The walk is therefore SLOT-synchronous: in step u every group handles slot u of its chunk -- add the slot's row
// to the running sum; if the slot is the head of a run, fetch the row's w, m, v first; if it is the tail, apply the
// optimizer (or park the partial sum of a run that crosses chunks).  Most runs of a large sparse batch are one or two
// slots long, so nearly every step does all three things in every group and the warp stays converged.  The rows of
// the slots two steps ahead are put in flight with L2 prefetches (short distance: the footprint stays far below L2).
//
// Chunk ids, the rule that decides which chunk finishes a run, the per-chunk partial rows of long runs and the
// span/hub kernel that finishes them are exactly those of the warp-per-chunk kernel, so the two are interchangeable
// per launch and the summation order (sorted slot order) is the same: results are bit-identical between them.
#include "kge_apply.cuh"

#define KGE_RAG_THREADS 128
#define KGE_RAG_PF 2  // prefetch distance in slots

template <int GS, int TMODE>
__global__ void __launch_bounds__(KGE_RAG_THREADS) kge_reduce_apply_group_kernel(ApplyParams P) {
    constexpr int V = 4;
    constexpr int GPB = KGE_RAG_THREADS / GS;  // chunks per CTA
    __shared__ SlotMeta meta[GPB][2 * KGE_CH];
    __shared__ int32_t skey[GPB][2 * KGE_CH + 1];
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1), gib = threadIdx.x / GS;
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << (lane & ~(GS - 1)));
    const int64_t w = (int64_t)blockIdx.x * GPB + gib;
    const int64_t b0 = w * KGE_CH;
    const bool live = b0 < P.n_keys;  // surplus groups of the last CTA idle through the (warp-uniform) loop
    const int cnt = live ? (int)min((int64_t)KGE_CH, P.n_keys - b0) : 0;
    const int K = P.ent.K;

    // own chunk in [0,16), the next chunk in [16,32) (candidates for a spill-over run)
    for (int t = lg; t < 2 * KGE_CH; t += GS) {
        int32_t key = -2;
        if (live && b0 + t < P.n_keys) {
            const uint64_t kv = P.ks[b0 + t];
            key = (int32_t)(kv >> 32);
            meta[gib][t] = decode_slot(P.G, (int32_t)(kv & 0xffffffffu));
        }
        skey[gib][t] = key;
    }
    if (lg == 0) skey[gib][2 * KGE_CH] = -3;
    const int32_t key_prev = (live && b0 > 0) ? (int32_t)(P.ks[b0 - 1] >> 32) : -1;
    const int32_t key_prev2 = (live && b0 > KGE_CH) ? (int32_t)(P.ks[b0 - KGE_CH - 1] >> 32) : -1;
    const int32_t key_next2 = (live && b0 + 2 * KGE_CH < P.n_keys) ? (int32_t)(P.ks[b0 + 2 * KGE_CH] >> 32) : -1;
    __syncwarp();
    const int32_t key_next = skey[gib][KGE_CH];  // -2 when there is no next chunk
    const int32_t key_last = cnt > 0 ? skey[gib][cnt - 1] : -4;
    // the last run of the chunk: where it starts, and how far it reaches into the next chunk
    int last_a = cnt - 1;
    while (last_a > 0 && skey[gib][last_a - 1] == key_last) --last_a;
    int ext = 0;
    while (ext < KGE_CH && skey[gib][KGE_CH + ext] == key_last && key_last >= 0) ++ext;
    // a run that starts here and ends inside the next chunk is finished here (same rule as the warp-per-chunk kernel)
    const bool last_open_start = (last_a == 0) && (key_last == key_prev);
    const bool last_spills = cnt == KGE_CH && key_last == key_next && !last_open_start && key_next2 != key_last;
    const int lim = last_spills ? cnt + ext : cnt;

    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    const int cc = min(lg * V, K - V);  // lanes past the end of the row read a valid duplicate and never store
    const bool col_ok = lg * V < K;
    const bool pf = prefetch_on(P) && !no_update;

    // per-run state (group-uniform)
    float g[V], rc[V], mv[V], vv[V];
    bool process = false, complete = false, open_start = false;
    RowPtrs r;
    r.w = r.m = r.v = nullptr;
    r.is_rel = false;
    r.owned = false;
    r.row = 0;
#pragma unroll
    for (int x = 0; x < V; ++x) g[x] = rc[x] = mv[x] = vv[x] = 0.f;

    for (int u = 0; __any_sync(0xffffffffu, u < lim); ++u) {
        if (u >= lim) continue;
        const int32_t key = skey[gib][u];
        const bool head = (u == 0) || (key != skey[gib][u - 1]);
        const bool tail = (u + 1 == lim) || (skey[gib][u + 1] != key);
        if (pf && u + KGE_RAG_PF < lim) {
            // rows needed KGE_RAG_PF slots from now: the slot's gradient row and, at a run head, the row's optimizer state
            const int un = u + KGE_RAG_PF;
            const int32_t kn = skey[gib][un];
            if (lg == 0) prefetch_l2(meta[gib][un].row);
            if (kn != skey[gib][un - 1] && lg == 1) prefetch_row_state(P, kn, need_m, need_v);
        }
        if (head) {
            // the run [u, b): b = next head, or the end of the chunk (+ the spill-over of the last run)
            const bool is_last = (u >= last_a);
            open_start = (u == 0) && (key == key_prev);
            bool open_end = is_last && cnt == KGE_CH && key == key_next && !last_spills;
            r = resolve_row(P, key);
            process = r.owned && !(open_start && !open_end && key_prev2 != key);  // else the head's group (chunk w-1) reduces it
            complete = !open_start && !open_end;
            if (process && !open_start && open_end && lg == 0) P.span_list[atomicAdd(P.span_count, 1)] = (int32_t)w;
            if (process && complete && lg == 0) mark_touched(P, key);
#pragma unroll
            for (int x = 0; x < V; ++x) g[x] = rc[x] = mv[x] = vv[x] = 0.f;
            if (process) {
                if (complete || TMODE != 0) ldg_vec<V>(rc, r.w + cc);
                if (complete && need_m) ldg_vec<V>(mv, r.m + cc);
                if (complete && need_v) ldg_vec<V>(vv, r.v + cc);
            }
        }
        if (process) {
            float v0[V];
            const SlotMeta m0 = meta[gib][u];
            ldg_vec<V>(v0, m0.row + cc);
            add_slot<V, TMODE>(g, v0, m0.c, m0.mode, rc);
        }
        if (tail && process && col_ok) {
            if (!complete) {
                st_vec<V>(P.partial + ((size_t)(2 * w + (open_start ? 0 : 1))) * K + cc, g);
            } else {
                float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
                reg_add<V>(P, r.is_rel, g, rc);
                if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + cc, g);
                if (!no_update) {
                    opt_math<V>(P, reset, g, rc, mv, vv);
                    if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + cc, mv);
                    if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + cc, vv);
                    st_vec<V>(r.w + cc, rc);
                }
            }
        }
    }
    (void)gmask;
}

template <int GS>
static int launch_group(const ApplyParams& P, int tmode, cudaStream_t st) {
    constexpr int GPB = KGE_RAG_THREADS / GS;
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    dim3 grid((unsigned)((n_chunks + GPB - 1) / GPB)), block(KGE_RAG_THREADS);
    if (tmode == 0) kge_reduce_apply_group_kernel<GS, 0><<<grid, block, 0, st>>>(P);
    else if (tmode == 1) kge_reduce_apply_group_kernel<GS, 1><<<grid, block, 0, st>>>(P);
    else kge_reduce_apply_group_kernel<GS, 2><<<grid, block, 0, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// level 1 of the reduction for K % 4 == 0, K <= 64; the caller zeroes span_count first and launches the span kernel after
int kge_launch_apply_group(const ApplyParams& P, int tmode, cudaStream_t st) {
    const int K = P.ent.K;
    KGE_REQUIRE(K % 4 == 0 && K <= 64, "kge_launch_apply_group: K=%d", K);
    return K <= 32 ? launch_group<8>(P, tmode, st) : launch_group<16>(P, tmode, st);
}
