// Segmented reduction + sparse row-wise optimizer for NARROW rows (K <= 64 floats): the column slices of the
// dimension-sharded multi-GPU step (kge_dim.cuh) and small embeddings on one GPU.
//
// kge_reduce_apply_kernel (kge_train.cu) gives a whole warp to a chunk of KGE_CH sorted slots; with a 128-byte
// row only 8 of its 32 lanes would carry data.  Here a GROUP of GS lanes (8 or 16) owns a chunk, so a warp
// walks 32/GS chunks at once and every load instruction fetches 32/GS rows.  Chunk ids, the rule that decides
// which chunk finishes a run, the per-chunk partial rows of long runs and the span/hub kernel that finishes
// them are exactly those of the warp-per-chunk kernel, so the two are interchangeable per launch and the
// summation order (sorted slot order) is the same: results are bit-identical between them.
#include "kge_apply.cuh"

#define KGE_RAG_THREADS 128

template <int GS, int TMODE>
__global__ void __launch_bounds__(KGE_RAG_THREADS) kge_reduce_apply_group_kernel(ApplyParams P) {
    constexpr int V = 4;
    constexpr int GPB = KGE_RAG_THREADS / GS;  // chunks per CTA
    __shared__ SlotMeta meta[GPB][2 * KGE_CH];
    __shared__ int32_t skey[GPB][2 * KGE_CH];
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1), gib = threadIdx.x / GS;
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << (lane & ~(GS - 1)));
    const int64_t w = (int64_t)blockIdx.x * GPB + gib;
    const int64_t b0 = w * KGE_CH;
    if (b0 >= P.n_keys) return;
    const int cnt = (int)min((int64_t)KGE_CH, P.n_keys - b0);
    const int K = P.ent.K;

    // own chunk in [0,16), the next chunk in [16,32) (candidates for a spill-over run)
    for (int t = lg; t < 2 * KGE_CH; t += GS) {
        int32_t key = -2;
        if (b0 + t < P.n_keys) {
            const uint64_t kv = P.ks[b0 + t];
            key = (int32_t)(kv >> 32);
            meta[gib][t] = decode_slot(P.G, (int32_t)(kv & 0xffffffffu));
        }
        skey[gib][t] = key;
    }
    const int32_t key_prev = b0 > 0 ? (int32_t)(P.ks[b0 - 1] >> 32) : -1;
    const int32_t key_prev2 = b0 > KGE_CH ? (int32_t)(P.ks[b0 - KGE_CH - 1] >> 32) : -1;
    const int32_t key_next2 = b0 + 2 * KGE_CH < P.n_keys ? (int32_t)(P.ks[b0 + 2 * KGE_CH] >> 32) : -1;
    __syncwarp(gmask);
    const int32_t key_next = skey[gib][KGE_CH];  // -2 when there is no next chunk
    const int32_t key_last = skey[gib][cnt - 1];
    unsigned heads = 0;
    for (int t = lg; t < cnt; t += GS)
        if (t == 0 || skey[gib][t] != skey[gib][t - 1]) heads |= 1u << t;
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) heads |= __shfl_xor_sync(gmask, heads, o);
    int ext = 0;  // slots at the front of the next chunk that continue this chunk's last run
    while (ext < KGE_CH && skey[gib][KGE_CH + ext] == key_last && key_last >= 0) ++ext;

    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    const bool need_m = !no_update && !reset && P.opt != KGE_OPT_SGD;
    const bool need_v = !no_update && !reset && P.opt == KGE_OPT_ADAM;
    const int cc = min(lg * V, K - V);  // lanes past the end of the row read a valid duplicate and never store
    const bool col_ok = lg * V < K;
    // The runs of a chunk are walked one after the other and each needs its row's w, m, v from HBM: put all of them in
    // flight now (one L2 prefetch per 128-byte line, issued by the lane that holds the run head), so that the walk below
    // finds them in L2 instead of paying one DRAM round trip per run
    if (!no_update && prefetch_on(P))
        for (int t = lg; t < cnt; t += GS)
            if ((heads >> t) & 1u) prefetch_row_state(P, skey[gib][t], need_m, need_v);

    while (heads) {
        const int a = __ffs(heads) - 1;
        heads &= heads - 1;
        int b = heads ? (__ffs(heads) - 1) : cnt;
        const int32_t skey_a = skey[gib][a];
        const bool open_start = (a == 0) && (skey_a == key_prev);
        bool open_end = (b == cnt) && (skey_a == key_next);
        const RowPtrs r = resolve_row(P, skey_a);
        if (!r.owned) continue;
        if (open_start && !open_end && key_prev2 != skey_a) continue;  // the head's group (chunk w-1) reduces this run
        if (!open_start && open_end && key_next2 != skey_a) {
            b = cnt + ext;  // the run ends inside the next chunk: finish it here
            open_end = false;
        }
        if (!open_start && open_end && lg == 0) P.span_list[atomicAdd(P.span_count, 1)] = (int32_t)w;
        const bool complete = !open_start && !open_end;
        if (complete && lg == 0) mark_touched(P, skey_a);
        float* part = P.partial + ((size_t)(2 * w + (open_start ? 0 : 1))) * K;
        float* dbg = r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent;
        float g[V], rc[V], mv[V], vv[V];
#pragma unroll
        for (int x = 0; x < V; ++x) g[x] = rc[x] = mv[x] = vv[x] = 0.f;
        if (complete || TMODE != 0) ldg_vec<V>(rc, r.w + cc);
        if (complete && need_m) ldg_vec<V>(mv, r.m + cc);
        if (complete && need_v) ldg_vec<V>(vv, r.v + cc);
        int u = a;
        for (; u + 2 <= b; u += 2) {
            float v0[V], v1[V];
            const SlotMeta m0 = meta[gib][u], m1 = meta[gib][u + 1];
            ldg_vec<V>(v0, m0.row + cc);
            ldg_vec<V>(v1, m1.row + cc);
            add_slot<V, TMODE>(g, v0, m0.c, m0.mode, rc);
            add_slot<V, TMODE>(g, v1, m1.c, m1.mode, rc);
        }
        if (u < b) {
            float v0[V];
            const SlotMeta m0 = meta[gib][u];
            ldg_vec<V>(v0, m0.row + cc);
            add_slot<V, TMODE>(g, v0, m0.c, m0.mode, rc);
        }
        if (!col_ok) continue;
        if (!complete) {
            st_vec<V>(part + cc, g);
            continue;
        }
        reg_add<V>(P, r.is_rel, g, rc);
        if (dbg != nullptr) st_vec<V>(dbg + (size_t)r.row * K + cc, g);
        if (no_update) continue;
        opt_math<V>(P, reset, g, rc, mv, vv);
        if (r.m && P.opt != KGE_OPT_SGD) st_vec<V>(r.m + cc, mv);
        if (r.v && P.opt == KGE_OPT_ADAM) st_vec<V>(r.v + cc, vv);
        st_vec<V>(r.w + cc, rc);
    }
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
