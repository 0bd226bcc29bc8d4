// Dimension-sharded ("column-parallel") training step for one NVSwitch domain (sm_100a).
//
// Every GPU holds a contiguous COLUMN slice of every embedding row (ent[E,Kc], rel[R,Kc], Kc = K/W, plus the
// optimizer state of that slice) and processes the WHOLE global batch on its slice.  All four scoring functions
// are sums over columns (TransE-L2: the square of the distance is), so the only cross-GPU coupling of a step is
// one sum of (1+eta) floats per positive:
//
//   phase 1  kge_dim_partial_kernel   column-slice partial sums of the positive and its eta negatives
//   ------   all-reduce(sum) of the partial sums over the ranks (4*(1+eta) bytes per positive; the host places it)
//   phase 2  kge_dim_backward_kernel  scores from the totals -> loss, dL/dscore -> gradient rows of the slice
//   then the single-GPU duplicate-row reduction + sparse optimizer runs unchanged on the slice (kge_train.cu).
//
// Against the row-sharded exchange (rows or folded queries travel: >= 2+2W rows of 4K bytes per positive) the
// NVLink volume per positive drops from ~33 KiB to 260 bytes at K = 256, eta = 64, W = 8; every other byte of
// the step stays in local HBM.  Corruptions, sort keys and the loss are replicated: every rank draws the same
// Philox stream for the whole global batch and evaluates the same loss from the same totals.
//
// Rows are narrow here (Kc = 32 floats at W = 8), so one GROUP of GS lanes (8, 16 or 32) owns a positive and
// a warp works on 32/GS positives at once; lanes of a group fetch the replacement ids of U negatives in one
// load, and the loss terms of those U negatives are evaluated one per lane instead of redundantly.
//
// Replaces reference models/EmbeddingModel.py:614-822 (_get_model_loss) for a model whose table is split over
// GPUs; the reference's only answer to large tables is host paging (models/EmbeddingModel.py:645-666, :1251-1281).
#pragma once
#include "kge_train_fwd.cuh"

struct DimParams {
    const float* ent;     // local column slice [E, K]
    const float* rel;     // [R, K]
    const int32_t* pos;   // [n,3] the global batch
    const int32_t* repl;  // [eta*n] replacement of negative (j,i) at j*n+i
    const uint8_t* keep;  // [eta*n] 1 = subject kept
    int64_t n;            // positives of the global batch
    int64_t i0, i1;       // positives handled by this launch
    int eta, k, K, loss, nl;  // k: floats per half (complex) / per row of the local slice
    float margin, scale, alpha;
    // raw sums of the launch's positives, chunk-local layout with nc = i1 - i0:
    //   [0,nc) positives | nc + j*nc + (i - i0) negative (j,i)
    float* sums;
    float* gbuf;        // gradient buffer of the whole batch (layout in kge_train_fwd.cuh)
    float* loss_part;   // [n]
    float* dbg_scores;  // optional [n*(1+eta)]
};

template <int GS>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int GS>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// lane lg of a group owns vectors c = lg + GS*i (i < NCH) of 4 floats; columns past the end of the slice are 0
template <int GS, int NCH, bool CPLX>
__device__ __forceinline__ void grow_load(Row<4, NCH, CPLX>& r, const float* __restrict__ base, int lg, int nvec, int half) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = lg + GS * i;
        if (c < nvec) {
            ld_vec<4>(r.re[i], base + (size_t)c * 4);
            if constexpr (CPLX) ld_vec<4>(r.im[i], base + half + (size_t)c * 4);
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                r.re[i][v] = 0.f;
                if constexpr (CPLX) r.im[i][v] = 0.f;
            }
        }
    }
}
template <int GS, int NCH, bool CPLX>
__device__ __forceinline__ void grow_store(float* __restrict__ base, const Row<4, NCH, CPLX>& r, int lg, int nvec, int half) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = lg + GS * i;
        if (c < nvec) {
            st_vec<4>(base + (size_t)c * 4, r.re[i]);
            if constexpr (CPLX) st_vec<4>(base + half + (size_t)c * 4, r.im[i]);
        }
    }
}
template <int NCH, bool CPLX>
__device__ __forceinline__ void row_select(Row<4, NCH, CPLX>& d, bool first, const Row<4, NCH, CPLX>& a, const Row<4, NCH, CPLX>& b) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            d.re[i][v] = first ? a.re[i][v] : b.re[i][v];
            if constexpr (CPLX) d.im[i][v] = first ? a.im[i][v] : b.im[i][v];
        }
}

#define KGE_DIM_THREADS 256
// The candidate rows of a positive are fetched in rounds of U; a round costs two dependent memory round trips (ids,
// then rows), and ncu showed both phase kernels waiting on exactly that (long-scoreboard stalls, DRAM 40 % busy).  The
// lanes therefore load the ids of the round KGE_DIM_PF rounds ahead at the top of a round and, at its end, put those
// rows in flight with L2 prefetches (no registers held, no dependency): when their round comes the row loads hit L2.
// The footprint -- resident groups x U x KGE_DIM_PF rows -- stays a fraction of L2.
#define KGE_DIM_PF 1000  // rounds of look-ahead; 1000 = off: measured slower on B200 (partial 198 -> 213 us, backward 213 -> 227 us)
__device__ __forceinline__ void dim_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// prefetch the K floats of a row slice: one 128-byte line per 32 floats, lines dealt over the lanes of the group
template <int GS>
__device__ __forceinline__ void dim_prefetch_row(const float* row, int K, int lg_rot) {
    for (int off = lg_rot * 32; off < K; off += GS * 32) dim_prefetch_l2(row + off);
}

// resident CTAs the compiler is asked to allow (register cap): the row registers grow with NCH and the complex halves
template <int MODEL, int NCH>
struct DimOcc {
    static constexpr int regs = NCH * 4 * (MODEL == 3 ? 2 : 1);
    static constexpr int bwd = regs <= 4 ? 4 : (regs <= 8 ? 3 : (regs <= 16 ? 2 : 1));
    static constexpr int fwd = regs <= 4 ? 6 : (regs <= 8 ? 5 : (regs <= 16 ? 3 : 2));
};

// ---------------------------------------------------------------------------------------------- phase 1
template <int MODEL, int GS, int NCH, int U>
__global__ void __launch_bounds__(KGE_DIM_THREADS, (DimOcc<MODEL, NCH>::fwd)) kge_dim_partial_kernel(DimParams P) {
    using A = Algebra<MODEL, 4, NCH>;
    using R = typename A::R;
    static_assert(U <= GS, "the lanes of a group fetch the ids of one round of U negatives");
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1), gbase = lane & ~(GS - 1);
    const int64_t nc = P.i1 - P.i0;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    const bool valid = grp < nc;
    const int64_t il = valid ? grp : nc - 1;  // surplus groups shadow the last positive (warp stays convergent)
    const int64_t i = P.i0 + il, n = P.n;
    const int K = P.K, eta = P.eta;
    const int half = A::CPLX ? P.k : 0;
    const int nvec = (A::CPLX ? P.k : K) / 4;
    float msk[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) msk[c] = 1.f;  // rows are zero-filled past the end: no mask needed

    R Qo, Qs;
    {
        R s, p, o;
        const int32_t si = P.pos[3 * i + 0], pi = P.pos[3 * i + 1], oi = P.pos[3 * i + 2];
        grow_load<GS>(s, P.ent + (size_t)si * K, lg, nvec, half);
        grow_load<GS>(p, P.rel + (size_t)pi * K, lg, nvec, half);
        grow_load<GS>(o, P.ent + (size_t)oi * K, lg, nvec, half);
        A::queries(s, p, o, Qo, Qs);
        const float sp = group_sum<GS>(A::partial(Qo, o, msk));
        if (valid && lg == 0) P.sums[il] = sp;
    }
    float* out = P.sums + nc + il;
    for (int j0 = 0; j0 < eta; j0 += U) {
        // one id / side load per group serves U negatives
        int my_idx = 0, my_keep = 0, pf_idx = -1;
        if (lg < U) {
            const int64_t q = (int64_t)min(j0 + lg, eta - 1) * n + i;
            my_idx = P.repl[q];
            my_keep = P.keep[q];
            if (j0 + KGE_DIM_PF * U + lg < eta) pf_idx = P.repl[(int64_t)(j0 + KGE_DIM_PF * U + lg) * n + i];
        }
        R r[U];
        bool kp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = __shfl_sync(0xffffffffu, my_idx, gbase + u);
            kp[u] = __shfl_sync(0xffffffffu, my_keep, gbase + u) != 0;
            grow_load<GS>(r[u], P.ent + (size_t)idx * K, lg, nvec, half);
        }
        float pv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            R Q;
            row_select(Q, kp[u], Qo, Qs);
            pv[u] = A::partial(Q, r[u], msk);
        }
#pragma unroll
        for (int o2 = GS / 2; o2 > 0; o2 >>= 1)
#pragma unroll
            for (int u = 0; u < U; ++u) pv[u] += __shfl_xor_sync(0xffffffffu, pv[u], o2);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (valid && j0 + u < eta && lg == u) out[(int64_t)(j0 + u) * nc] = pv[u];
        if (pf_idx >= 0) dim_prefetch_row<1>(P.ent + (size_t)pf_idx * K, K, 0);
    }
}

// ---------------------------------------------------------------------------------------------- phase 2
template <int MODEL, int GS, int NCH, int U>
__global__ void __launch_bounds__(KGE_DIM_THREADS, (DimOcc<MODEL, NCH>::bwd)) kge_dim_backward_kernel(DimParams P) {
    using A = Algebra<MODEL, 4, NCH>;
    using R = typename A::R;
    constexpr bool TRANSE = A::TRANSE;
    constexpr int V = 4;  // ROW_FOR
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1), gbase = lane & ~(GS - 1);
    const int64_t nc = P.i1 - P.i0;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    const bool valid = grp < nc;
    const int64_t il = valid ? grp : nc - 1;
    const int64_t i = P.i0 + il, n = P.n;
    const int K = P.K, eta = P.eta, loss = P.loss, nl = P.nl;
    const int half = A::CPLX ? P.k : 0;
    const int nvec = (A::CPLX ? P.k : K) / 4;
    const float margin = P.margin, alpha = P.alpha, scale = P.scale;
    const float* tot = P.sums + nc + il;  // negative (j,i) at tot[j*nc]

    R s, p, o, Qo, Qs, AccO, AccS;
    const int32_t si = P.pos[3 * i + 0], pi = P.pos[3 * i + 1], oi = P.pos[3 * i + 2];
    grow_load<GS>(s, P.ent + (size_t)si * K, lg, nvec, half);
    grow_load<GS>(p, P.rel + (size_t)pi * K, lg, nvec, half);
    grow_load<GS>(o, P.ent + (size_t)oi * K, lg, nvec, half);
    A::queries(s, p, o, Qo, Qs);
    row_zero(AccO);
    row_zero(AccS);

    // the loss sees nl(score); dL/draw = dL/dnl * nl'(raw)  (models/EmbeddingModel.py:679-689, :801-812)
    const float spos_raw = A::finish(P.sums[il], scale);
    const float dpos_nl = apply_nl_grad(nl, spos_raw);
    const float spos = apply_nl(nl, spos_raw);
    const float cpos = clip75(spos);
    const bool pos_in = (spos >= -75.f) && (spos <= 75.f);

    float loss_acc = 0.f, wsum = 0.f;  // lane-partial until the group sums below
    float zinv = 0.f, amax = 0.f, lbar = 0.f;
    const bool two_pass = loss_two_pass(loss);
    if (two_pass) {
        // softmax statistics over the eta totals of this positive (losses/nll_multiclass.py:70-81,
        // losses/self_adversarial.py:97-110); lanes stride over j
        if (loss == KGE_LOSS_MULTICLASS_NLL) {
            float zpart = 0.f;
            for (int j = lg; j < eta; j += GS) zpart += expf(clip75(apply_nl(nl, A::finish(tot[(int64_t)j * nc], scale))));
            const float pe = expf(cpos);
            const float z = group_sum<GS>(zpart) + pe;
            zinv = 1.f / z;
            loss_acc = (lg == 0) ? -logf(pe / z) : 0.f;
        } else {
            float mx = -INFINITY;
            for (int j = lg; j < eta; j += GS) mx = fmaxf(mx, alpha * apply_nl(nl, A::finish(tot[(int64_t)j * nc], scale)));
            amax = group_max<GS>(mx);
            float zpart = 0.f, lpart = 0.f;
            for (int j = lg; j < eta; j += GS) {
                const float st = apply_nl(nl, A::finish(tot[(int64_t)j * nc], scale));
                const float e = expf(alpha * st - amax);
                zpart += e;
                lpart += e * log_sigmoid(-st - margin);
            }
            zinv = 1.f / group_sum<GS>(zpart);
            lbar = group_sum<GS>(lpart) * zinv;
            loss_acc = (lg == 0) ? -log_sigmoid(margin + spos) - lbar : 0.f;
        }
    }

    float* coef = gbuf_coef(P.gbuf, n, K);
    uint8_t* keep_out = gbuf_keep(P.gbuf, eta, n, K);
    for (int j0 = 0; j0 < eta; j0 += U) {
        // lane lg < U owns negative j0+lg of this round: id, side, total -> dL/dscore, coefficient, loss term
        int my_idx = 0, my_keep = 0, pf_idx = -1;
        float my_w = 0.f, my_sn = 0.f;
        if (lg < U) {
            const int j = min(j0 + lg, eta - 1);
            const bool real = j0 + lg < eta;
            const int64_t q = (int64_t)j * n + i;
            my_idx = P.repl[q];
            my_keep = P.keep[q];
            if (j0 + KGE_DIM_PF * U + lg < eta) pf_idx = P.repl[(int64_t)(j0 + KGE_DIM_PF * U + lg) * n + i];
            const float sn = A::finish(tot[(int64_t)j * nc], scale);
            const float st = apply_nl(nl, sn);
            const float dn = apply_nl_grad(nl, sn);
            float w;
            if (loss == KGE_LOSS_MULTICLASS_NLL) {
                const bool in = (st >= -75.f) && (st <= 75.f);
                w = in ? expf(st) * zinv * dn : 0.f;
            } else if (loss == KGE_LOSS_SELF_ADVERSARIAL) {
                const float pj = expf(alpha * st - amax) * zinv;
                w = pj * (sigmoidf(st + margin) - alpha * (log_sigmoid(-st - margin) - lbar)) * dn;
            } else if (loss == KGE_LOSS_PAIRWISE || loss == KGE_LOSS_ABSOLUTE_MARGIN) {
                // losses/pairwise.py:69, absolute_margin.py:69 ; tf.maximum passes the gradient when t >= 0
                const float tt = loss == KGE_LOSS_PAIRWISE ? margin - spos + st : margin + st;
                w = (tt >= 0.f) ? 1.f : 0.f;
                if (real) {
                    loss_acc += fmaxf(tt, 0.f);
                    wsum += w;
                }
                w *= dn;
            } else {
                // losses/nll.py:55-59 : log(1+exp(clip(neg)))
                const float e = expf(clip75(st));
                const bool in = (st >= -75.f) && (st <= 75.f);
                if (real) loss_acc += logf(1.f + e);
                w = in ? e / (1.f + e) * dn : 0.f;
            }
            if (!real) w = 0.f;
            my_w = w;
            my_sn = sn;
            if (valid && real) {
                coef[q] = A::coefficient(w, sn, scale);
                keep_out[q] = (uint8_t)my_keep;
                if (P.dbg_scores != nullptr) P.dbg_scores[n + q] = st;
            }
        }
        R r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = __shfl_sync(0xffffffffu, my_idx, gbase + u);
            grow_load<GS>(r[u], P.ent + (size_t)idx * K, lg, nvec, half);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool kp = __shfl_sync(0xffffffffu, my_keep, gbase + u) != 0;
            const float w = __shfl_sync(0xffffffffu, my_w, gbase + u);
            const float wO = kp ? w : 0.f, wS = kp ? 0.f : w;
            if constexpr (TRANSE) {
                const float sn = __shfl_sync(0xffffffffu, my_sn, gbase + u);
                float inv = 0.f;
                if constexpr (MODEL == 1) inv = sn != 0.f ? 1.f / (-sn) : 0.f;
                const float sg = kp ? 1.f : -1.f;
                ROW_FOR(c, v) {
                    const float q = kp ? Qo.re[c][v] : Qs.re[c][v];
                    const float uu = sg * (q - r[u].re[c][v]);
                    float g;  // d f / d u
                    if constexpr (MODEL == 0) g = (uu > 0.f) ? -1.f : ((uu < 0.f) ? 1.f : 0.f);
                    else g = -uu * inv;
                    AccO.re[c][v] = fmaf(wO, g, AccO.re[c][v]);
                    AccS.re[c][v] = fmaf(wS, g, AccS.re[c][v]);
                }
            } else {
                const float a = wO * scale, b = wS * scale;
                ROW_FOR(c, v) {
                    AccO.re[c][v] = fmaf(a, r[u].re[c][v], AccO.re[c][v]);
                    AccS.re[c][v] = fmaf(b, r[u].re[c][v], AccS.re[c][v]);
                    if constexpr (A::CPLX) {
                        AccO.im[c][v] = fmaf(a, r[u].im[c][v], AccO.im[c][v]);
                        AccS.im[c][v] = fmaf(b, r[u].im[c][v], AccS.im[c][v]);
                    }
                }
            }
        }
        if (pf_idx >= 0) dim_prefetch_row<1>(P.ent + (size_t)pf_idx * K, K, 0);
    }
    loss_acc = group_sum<GS>(loss_acc);
    wsum = group_sum<GS>(wsum);

    float wpos;
    if (loss == KGE_LOSS_PAIRWISE) {
        wpos = -wsum;
    } else if (loss == KGE_LOSS_NLL) {
        // positives are tiled eta times (models/EmbeddingModel.py:724-729)
        const float e = expf(-cpos);
        loss_acc += (float)eta * logf(1.f + e);
        wpos = pos_in ? -(float)eta * (e / (1.f + e)) : 0.f;
    } else if (loss == KGE_LOSS_ABSOLUTE_MARGIN) {
        loss_acc -= (float)eta * spos;
        wpos = -(float)eta;
    } else if (loss == KGE_LOSS_SELF_ADVERSARIAL) {
        wpos = -sigmoidf(-(margin + spos));
    } else {
        wpos = pos_in ? -(1.f - expf(cpos) * zinv) : 0.f;
    }
    wpos *= dpos_nl;
    if (!valid) return;

    float* GB = P.gbuf;
    grow_store<GS>(GB + (size_t)(3 * n + i) * K, Qo, lg, nvec, half);
    grow_store<GS>(GB + (size_t)(4 * n + i) * K, Qs, lg, nvec, half);
    {
        R gs, gp, go;
        A::backward_pos(Qo, o, wpos, spos_raw, scale, go, AccO);
        A::fold(s, p, o, AccO, AccS, gs, gp, go);
        grow_store<GS>(GB + (size_t)i * K, gs, lg, nvec, half);
        grow_store<GS>(GB + (size_t)(n + i) * K, go, lg, nvec, half);
        grow_store<GS>(GB + (size_t)(2 * n + i) * K, gp, lg, nvec, half);
    }
    if (lg == 0) {
        P.loss_part[i] = loss_acc;
        if (P.dbg_scores != nullptr) P.dbg_scores[i] = spos;
    }
}

// ---------------------------------------------------------------------------------------------- phase 1, sorted order
// The candidate gather of kge_dim_partial_kernel reads 128-byte row slices in RANDOM order, and HBM delivers those at ~40 %
// of its copy bandwidth (request-rate bound: ncu, profiles/r02_summary.md).  The sort entries of the step list the same rows
// in ascending row order, duplicates adjacent.  So phase 1 can run over the SORTED entries instead: every entity row is then
// streamed once, in address order, and what is gathered at random is the 2n-row query table, which lives in L2:
//   kge_dim_query_kernel           per positive: Qo, Qs -> the gradient buffer's query rows; the positive's own partial sum
//   kge_dim_sorted_partial_kernel  per sort entry that is a negative (j,i): <Q_side(i), row(key)> -> sums
// Layout of `sums`: chunk-major for n_chunks pieces of the positive range (distributed.py:chunk_bounds), so that the host
// can all-reduce piece c while piece c+1 is still being consumed.
struct DimChunks {
    uint32_t base, extra;  // n = n_chunks*base + extra; the first `extra` chunks hold base+1 positives
};
__device__ __forceinline__ void dim_chunk_of(const DimChunks& C, uint32_t i, uint32_t& lo, uint32_t& nc) {
    const uint32_t split = C.extra * (C.base + 1);
    const uint32_t c = i < split ? i / (C.base + 1) : C.extra + (i - split) / C.base;
    lo = c * C.base + min(c, C.extra);
    nc = C.base + (c < C.extra ? 1u : 0u);
}

template <int MODEL, int GS, int NCH>
__global__ void __launch_bounds__(KGE_DIM_THREADS) kge_dim_query_kernel(DimParams P, DimChunks C, int eta1, uint2* __restrict__ pos_off) {
    using A = Algebra<MODEL, 4, NCH>;
    using R = typename A::R;
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1);
    const int64_t n = P.n;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    const bool valid = grp < n;
    const int64_t i = valid ? grp : n - 1;
    const int K = P.K;
    const int half = A::CPLX ? P.k : 0;
    const int nvec = (A::CPLX ? P.k : K) / 4;
    float msk[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) msk[c] = 1.f;
    R s, p, o, Qo, Qs;
    const int32_t si = P.pos[3 * i + 0], pi = P.pos[3 * i + 1], oi = P.pos[3 * i + 2];
    grow_load<GS>(s, P.ent + (size_t)si * K, lg, nvec, half);
    grow_load<GS>(p, P.rel + (size_t)pi * K, lg, nvec, half);
    grow_load<GS>(o, P.ent + (size_t)oi * K, lg, nvec, half);
    A::queries(s, p, o, Qo, Qs);
    const float sp = group_sum<GS>(A::partial(Qo, o, msk));
    if (!valid) return;
    grow_store<GS>(P.gbuf + (size_t)(3 * n + i) * K, Qo, lg, nvec, half);
    grow_store<GS>(P.gbuf + (size_t)(4 * n + i) * K, Qs, lg, nvec, half);
    if (lg == 0) {
        uint32_t lo, nc;
        dim_chunk_of(C, (uint32_t)i, lo, nc);
        const uint32_t o0 = (uint32_t)eta1 * lo + ((uint32_t)i - lo);
        P.sums[o0] = sp;
        // where the sums of this positive's negatives go: negative j at pos_off[i].x + (1 + j) * pos_off[i].y -- looked up
        // by the sorted kernel (two cached loads instead of two integer divisions per sort entry)
        pos_off[i] = make_uint2(o0, nc);
    }
}

template <int MODEL, int GS, int NCH, int U>
__global__ void __launch_bounds__(KGE_DIM_THREADS) kge_dim_sorted_partial_kernel(DimParams P, const uint2* __restrict__ pos_off,
                                                                                 const uint64_t* __restrict__ ks, int64_t n_keys) {
    using A = Algebra<MODEL, 4, NCH>;
    using R = typename A::R;
    const int lane = threadIdx.x & 31, lg = lane & (GS - 1);
    const uint32_t n = (uint32_t)P.n, eta_n = (uint32_t)P.eta * n;
    const int K = P.K;
    const int half = A::CPLX ? P.k : 0;
    const int nvec = (A::CPLX ? P.k : K) / 4;
    float msk[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) msk[c] = 1.f;
    // a group takes U consecutive sort entries (neighbours often name the same row: the second load hits L1)
    const int64_t x0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS) * U;
    R r[U], Q[U];
    uint32_t off[U];
    bool neg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t x = x0 + u;
        neg[u] = false;
        off[u] = 0;
        uint64_t kv = 0;
        if (x < n_keys) kv = ks[x];
        const uint32_t word = (uint32_t)(kv & 0xffffffffu), t = word & 0x3fffffffu;
        const uint32_t key = (uint32_t)(kv >> 32);
        const bool is_neg = x < n_keys && t >= 2u * n && t < 2u * n + eta_n;
        uint32_t i = 0, j = 0;
        if (is_neg) {
            const uint32_t q = t - 2u * n;
            j = q / n;
            i = q - j * n;
        }
        // the side rides in the sort entry (bit 31 set by kge_emit_kernel); entries without it read the side array
        const bool kept = (word & 0x80000000u) ? (word & 0x40000000u) != 0 : (is_neg ? P.keep[(size_t)j * n + i] != 0 : false);
        neg[u] = is_neg;
        if (is_neg) {
            const uint2 po = pos_off[i];
            off[u] = po.x + (1u + j) * po.y;
            grow_load<GS>(r[u], P.ent + (size_t)key * K, lg, nvec, half);
            grow_load<GS>(Q[u], P.gbuf + (size_t)((kept ? 3u : 4u) * n + i) * K, lg, nvec, half);
        } else {
            row_zero(r[u]);
            row_zero(Q[u]);
        }
    }
    float pv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) pv[u] = A::partial(Q[u], r[u], msk);
#pragma unroll
    for (int o2 = GS / 2; o2 > 0; o2 >>= 1)
#pragma unroll
        for (int u = 0; u < U; ++u) pv[u] += __shfl_xor_sync(0xffffffffu, pv[u], o2);
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (neg[u] && lg == (u & (GS - 1))) P.sums[off[u]] = pv[u];
}

// ---------------------------------------------------------------------------------------------- launch
// KGE_DIM_MAXCTAS=m (A/B knob): at most m resident CTAs of a phase kernel per SM -- an unused dynamic shared-memory
// request does the limiting -- so that the CTAs of the radix sort, which runs beside these kernels on the side stream, find
// registers and shared memory on every SM instead of waiting for whole waves of the phase kernel to drain.  0 = no limit.
static inline size_t dim_pad_smem() {
    static long v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_DIM_MAXCTAS");
        const int m = (e != nullptr && e[0] >= '1' && e[0] <= '8') ? (e[0] - '0') : 0;
        v = m > 0 ? (long)(233472 / m - 1024) : 0;
    }
    return (size_t)v;
}

template <int MODEL, int GS, int NCH, int U>
static int launch_dim_one(int phase, const DimParams& P, cudaStream_t st) {
    const int64_t nc = P.i1 - P.i0;
    const int gpb = KGE_DIM_THREADS / GS;
    dim3 grid((unsigned)((nc + gpb - 1) / gpb)), block(KGE_DIM_THREADS);
    const size_t pad = dim_pad_smem();
    if (pad > 48 * 1024) {
        static bool set = false;
        if (!set) {
            KGE_CUDA_CHECK(cudaFuncSetAttribute(kge_dim_partial_kernel<MODEL, GS, NCH, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
            KGE_CUDA_CHECK(cudaFuncSetAttribute(kge_dim_backward_kernel<MODEL, GS, NCH, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
            set = true;
        }
    }
    if (phase == 1) kge_dim_partial_kernel<MODEL, GS, NCH, U><<<grid, block, pad, st>>>(P);
    else kge_dim_backward_kernel<MODEL, GS, NCH, U><<<grid, block, pad, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// phase 3: queries + positives' sums; phase 4: sorted-order partial sums of the negatives
template <int MODEL, int GS, int NCH, int U>
static int launch_dim_sorted_one(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys,
                                 cudaStream_t st) {
    const int eta1 = P.eta + 1;
    if (phase == 3) {
        const int gpb = KGE_DIM_THREADS / GS;
        kge_dim_query_kernel<MODEL, GS, NCH><<<(unsigned)((P.n + gpb - 1) / gpb), KGE_DIM_THREADS, 0, st>>>(P, C, eta1, pos_off);
    } else {
        const int64_t epb = (int64_t)(KGE_DIM_THREADS / GS) * U;  // entries per CTA
        kge_dim_sorted_partial_kernel<MODEL, GS, NCH, U><<<(unsigned)((n_keys + epb - 1) / epb), KGE_DIM_THREADS, 0, st>>>(P, pos_off, ks, n_keys);
    }
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int MODEL>
static int launch_dim_sorted_model(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys,
                                   cudaStream_t st) {
    constexpr bool Cx = (MODEL == 3);
    const int width = Cx ? P.k : P.K;
    KGE_REQUIRE(width % 4 == 0 && P.K % 4 == 0, "kge_train (dimension-sharded): the local slice needs a multiple of 4 columns per half, got %d", width);
    const int nvec = width / 4;
    if (nvec <= 8) return launch_dim_sorted_one<MODEL, 8, 1, (Cx ? 2 : 4)>(phase, P, C, pos_off, ks, n_keys, st);
    if (nvec <= 16) return launch_dim_sorted_one<MODEL, 16, 1, (Cx ? 2 : 4)>(phase, P, C, pos_off, ks, n_keys, st);
    if (nvec <= 32) return launch_dim_sorted_one<MODEL, 32, 1, (Cx ? 2 : 4)>(phase, P, C, pos_off, ks, n_keys, st);
    if (nvec <= 64) return launch_dim_sorted_one<MODEL, 32, 2, (Cx ? 1 : 2)>(phase, P, C, pos_off, ks, n_keys, st);
    if (nvec <= 128) return launch_dim_sorted_one<MODEL, 32, 4, 1>(phase, P, C, pos_off, ks, n_keys, st);
    kge_set_error("kge_train (dimension-sharded): local slice of %d columns per half is too wide (max 512)", width);
    return -1;
}

// group size / chunks per lane from the vectors per half of the local slice
template <int MODEL>
static int launch_dim_model(int phase, const DimParams& P, cudaStream_t st) {
    constexpr bool C = (MODEL == 3);
    const int width = C ? P.k : P.K;
    KGE_REQUIRE(width % 4 == 0 && P.K % 4 == 0, "kge_train (dimension-sharded): the local slice needs a multiple of 4 columns per half, got %d", width);
    const int nvec = width / 4;
    if (nvec <= 8) return launch_dim_one<MODEL, 8, 1, (C ? 4 : 8)>(phase, P, st);
    if (nvec <= 16) return launch_dim_one<MODEL, 16, 1, (C ? 4 : 8)>(phase, P, st);
    if (nvec <= 32) return launch_dim_one<MODEL, 32, 1, (C ? 4 : 8)>(phase, P, st);
    if (nvec <= 64) return launch_dim_one<MODEL, 32, 2, (C ? 2 : 4)>(phase, P, st);
    if (nvec <= 128) return launch_dim_one<MODEL, 32, 4, (C ? 1 : 2)>(phase, P, st);
    kge_set_error("kge_train (dimension-sharded): local slice of %d columns per half is too wide (max 512)", width);
    return -1;
}

// one translation unit per model (parallel compilation): kge_dim_m{0,1,2,3}.cu
int kge_launch_dim_sorted_m0(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys, cudaStream_t st);
int kge_launch_dim_sorted_m1(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys, cudaStream_t st);
int kge_launch_dim_sorted_m2(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys, cudaStream_t st);
int kge_launch_dim_sorted_m3(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys, cudaStream_t st);
int kge_launch_dim_m0(int phase, const DimParams& P, cudaStream_t st);
int kge_launch_dim_m1(int phase, const DimParams& P, cudaStream_t st);
int kge_launch_dim_m2(int phase, const DimParams& P, cudaStream_t st);
int kge_launch_dim_m3(int phase, const DimParams& P, cudaStream_t st);
