// Shared pieces of the segmented reduction + sparse row-wise optimizer (used by kge_train.cu and by the
// group-per-chunk variant for narrow rows in kge_apply_group.cu).
#pragma once
#include "kge_train_fwd.cuh"

// ------------------------------------------------------------------------------------------------
// segmented reduction of duplicate rows + sparse row-wise optimizer
// ------------------------------------------------------------------------------------------------
#define KGE_CH 16  // sorted slots per warp

struct GradView {
    float*  base[KGE_MAX_SHARDS];  // rank r's gradient buffer (local or peer mapping)
    float*  tail[KGE_MAX_SHARDS];  // base to address rank r's [Qo|Qs|coef|keep] tail with the same offsets:
                                   // == base[r], or a local all-gathered copy minus the head size
    int64_t S;                     // slots per rank
    int64_t n;                     // positives per rank
    int     eta, K, n_ranks;
};

struct ApplyParams {
    const uint64_t* ks;    // sorted (key << 32 | global slot id), global slot id = rank*S + local slot
    int64_t n_keys;
    GradView G;
    TableView ent, ent_m, ent_v;
    float *rel, *rel_m, *rel_v;
    int64_t E, R;
    int64_t row_begin, row_end;  // owned entity rows
    int opt;
    bool has_m, has_v;     // entity optimizer-state tables present
    uint32_t flags;
    float lr, lr_t, beta1, beta2, eps, momentum;
    const KgeStepDyn* dyn;  // when set, lr_t is read from here (graph replay)
    float* partial;        // [2*n_chunks][K]
    int32_t* span_list;    // chunk ids that start a run crossing chunk borders (unordered)
    int32_t* hub_list;     // the subset whose run covers more than KGE_SPAN_WARP_MAX chunks
    int32_t* span_count;   // [2]: {#span heads, #hubs}: zero when the level-1 kernel starts (kge_span_apply_kernel re-zeroes them)
    int32_t* span_ticket;  // CTAs of kge_span_apply_kernel that have read the counters
    int span_use_hubs;     // 1: kge_span_warp_kernel runs first and leaves only hub_list to kge_span_apply_kernel
    float* dbg_grad_ent;
    float* dbg_grad_rel;
    // LP regulariser (regularizers/lp.py:81-113): lambda * sum |w|^p over the WHOLE tables, so every row has
    // a gradient.  Rows touched by the batch get it added in the reduction; `touched` (one bit per sort
    // key, entity ids then E + relation id) tells kge_reg_dense_kernel which rows are left.
    int reg_p;
    float reg_lambda_ent, reg_lambda_rel;
    uint32_t* touched;
    int prefetch;       // 1: short-distance L2 prefetch of the optimizer rows ahead of the walk (narrow-row kernel)
    int prefetch_wide;  // 1: the same, one run ahead, in the warp-per-chunk kernel
};

__device__ __forceinline__ bool prefetch_on(const ApplyParams& P) { return P.prefetch != 0; }

__device__ __forceinline__ float reg_grad1(float w, int p, float lam) {
    if (p == 2) return 2.f * lam * w;
    if (p == 1) return w > 0.f ? lam : (w < 0.f ? -lam : 0.f);
    if (p == 3) return 3.f * lam * w * fabsf(w);
    const float a = fabsf(w);
    return a > 0.f ? lam * (float)p * powf(a, (float)(p - 1)) * (w > 0.f ? 1.f : -1.f) : 0.f;
}
__device__ __forceinline__ float reg_term1(float w, int p) {
    const float a = fabsf(w);
    return p == 2 ? a * a : (p == 1 ? a : (p == 3 ? a * a * a : powf(a, (float)p)));
}

struct SlotMeta {
    const float* row;
    float c;
    int mode;  // 0: add row ; 1: replacement row of a negative, F(c, Q, r)
};

// Low word of a sort entry: slot id in bits [0,30); bit 31 = "the side of this negative travels with the entry", bit 30 =
// that side (1: subject kept -> query Qo).  kge_emit_kernel knows the side when it writes the entry, so the reduction does
// not have to fetch one byte per negative from a 32-byte DRAM sector in sorted (= random) order.  Entries made elsewhere
// (owner-side selection of the row-sharded path) leave both bits clear and the side is read from the gradient buffer.
#define KGE_SLOT_MASK 0x3fffffffu
#define KGE_SLOT_HAS_SIDE 0x80000000u
#define KGE_SLOT_SIDE 0x40000000u

__device__ __forceinline__ SlotMeta decode_slot(const GradView& G, uint32_t word) {
    int rr = 0;
    int64_t t = (int64_t)(word & KGE_SLOT_MASK);
    if (G.n_ranks > 1) {
        rr = (int)(t / G.S);
        t -= (int64_t)rr * G.S;
    }
    float* base = G.base[rr];
    float* tbase = G.tail[rr];
    const int64_t n = G.n;
    SlotMeta m;
    m.c = 1.f;
    m.mode = 0;
    if (t < 2 * n) {
        m.row = base + t * G.K;
    } else if (t < 2 * n + (int64_t)G.eta * n) {
        const int64_t q = t - 2 * n;
        const int64_t i = q % n;
        const float* coef = gbuf_coef(tbase, n, G.K);
        const uint8_t* keep = gbuf_keep(tbase, G.eta, n, G.K);
        m.c = coef[q];
        const bool kept = (word & KGE_SLOT_HAS_SIDE) ? (word & KGE_SLOT_SIDE) != 0 : keep[q] != 0;
        m.row = tbase + ((kept ? 3 : 4) * n + i) * G.K;
        m.mode = 1;
    } else {
        m.row = base + (2 * n + (t - 2 * n - (int64_t)G.eta * n)) * G.K;
    }
    return m;
}

struct RowPtrs {
    float *w, *m, *v;
    bool is_rel, owned;
    int64_t row;
};

__device__ __forceinline__ RowPtrs resolve_row(const ApplyParams& P, int32_t key) {
    RowPtrs r;
    r.is_rel = key >= P.E;
    r.row = r.is_rel ? key - P.E : key;
    r.owned = r.is_rel || (r.row >= P.row_begin && r.row < P.row_end);
    r.m = r.v = nullptr;
    const int K = P.ent.K;
    if (r.is_rel) {
        r.w = P.rel + (size_t)r.row * K;
        if (P.rel_m) r.m = P.rel_m + (size_t)r.row * K;
        if (P.rel_v) r.v = P.rel_v + (size_t)r.row * K;
    } else {
        r.w = table_row(P.ent, r.row);
        if (P.has_m) r.m = table_row(P.ent_m, r.row);
        if (P.has_v) r.v = table_row(P.ent_v, r.row);
    }
    return r;
}

// global (non-generic) vector load: the row pointers come out of shared memory, so the compiler cannot
// prove their address space on its own
template <int V>
__device__ __forceinline__ void ldg_vec(float (&d)[V], const float* p) {
    if constexpr (V == 4) {
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(p));
    } else {
        asm volatile("ld.global.f32 %0, [%1];" : "=f"(d[0]) : "l"(p));
    }
}

// L2 prefetch of one 128-byte line (no register, no dependency): used to put the optimizer rows of a whole chunk of
// sorted slots in flight before the runs are walked one after the other
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// prefetch w (and m, v when the optimizer reads them) of row `key` -- K floats each
__device__ __forceinline__ void prefetch_row_state(const ApplyParams& P, int32_t key, bool need_m, bool need_v);

// streaming variants for data touched once per step (the optimizer rows): evict-first in L1 and L2, so that the rows that
// ARE reused within the step -- queries, coefficients, sort entries -- stay cached
__device__ __forceinline__ void ldg_vec4_cs(float (&d)[4], const float* p) {
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(p));
}
__device__ __forceinline__ void stg_vec4_cs(float* p, const float (&d)[4]) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(d[0]), "f"(d[1]), "f"(d[2]), "f"(d[3]) : "memory");
}

// contribution of one slot to V columns of the gradient; rc = current value of the row being updated.
// Plain gradient rows (mode 0) carry c = 1, so the trilinear models need no branch at all.
template <int V, int TMODE>
__device__ __forceinline__ void add_slot(float (&g)[V], const float (&a)[V], float c, int mode, const float (&rc)[V]) {
#pragma unroll
    for (int x = 0; x < V; ++x) {
        if (TMODE == 0) {
            g[x] = fmaf(c, a[x], g[x]);  // DistMult / ComplEx / HolE: c*Q, or 1*row
        } else if (mode == 0) {
            g[x] += a[x];
        } else if (TMODE == 1) {
            float d = a[x] - rc[x];  // TransE L1: c*sign(Q-r)
            g[x] += d > 0.f ? c : (d < 0.f ? -c : 0.f);
        } else {
            g[x] = fmaf(c, a[x] - rc[x], g[x]);  // TransE L2: c*(Q-r)
        }
    }
}

// square root of the Adam / Adagrad denominators: one MUFU.SQRT (relative error <= 2^-22, far below the 1e-5 the optimizer
// parity is held to) instead of sqrtf's refinement sequence -- the sparse optimizer runs once per touched row and column and
// its instruction count is what bounds the narrow-row reduction
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// pure-register optimizer math on V columns, optimizer OPT fixed at compile time; m/v are the row's state (ignored when
// not needed); lr_t is Adam's bias-corrected rate of this step
template <int V, int OPT>
__device__ __forceinline__ void opt_math_t(const ApplyParams& P, bool reset, float lr_t, const float (&g)[V], float (&wv)[V], float (&mv)[V],
                                           float (&vv)[V]) {
    if constexpr (OPT == KGE_OPT_ADAM) {
        // Keras Adam (beta1 .9, beta2 .999, eps 1e-7): var -= lr_t * m / (sqrt(v) + eps)
        const float b1 = P.beta1, b2 = P.beta2, c1 = 1.f - P.beta1, c2 = 1.f - P.beta2;
#pragma unroll
        for (int x = 0; x < V; ++x) {
            const float m0 = reset ? 0.f : mv[x], v0 = reset ? 0.f : vv[x];
            mv[x] = b1 * m0 + c1 * g[x];
            vv[x] = b2 * v0 + c2 * g[x] * g[x];
            wv[x] = wv[x] - __fdividef(lr_t * mv[x], fast_sqrt(vv[x]) + P.eps);
        }
    } else if constexpr (OPT == KGE_OPT_ADAGRAD) {
        // Keras Adagrad: accumulator starts at 0.1; var -= lr * g / (sqrt(acc) + eps)
#pragma unroll
        for (int x = 0; x < V; ++x) {
            mv[x] = (reset ? 0.1f : mv[x]) + g[x] * g[x];
            wv[x] = wv[x] - __fdividef(P.lr * g[x], fast_sqrt(mv[x]) + P.eps);
        }
    } else if constexpr (OPT == KGE_OPT_MOMENTUM) {
        // Keras SGD momentum: vel = mu*vel - lr*g ; var += vel
#pragma unroll
        for (int x = 0; x < V; ++x) {
            mv[x] = P.momentum * (reset ? 0.f : mv[x]) - P.lr * g[x];
            wv[x] = wv[x] + mv[x];
        }
    } else {
#pragma unroll
        for (int x = 0; x < V; ++x) wv[x] = wv[x] - P.lr * g[x];
    }
}

// the same with the optimizer chosen at run time (one formula: every reduction kernel produces the same bits)
template <int V>
__device__ __forceinline__ void opt_math(const ApplyParams& P, bool reset, const float (&g)[V], float (&wv)[V], float (&mv)[V], float (&vv)[V]) {
    if (P.opt == KGE_OPT_ADAM) opt_math_t<V, KGE_OPT_ADAM>(P, reset, P.dyn != nullptr ? P.dyn->lr_t : P.lr_t, g, wv, mv, vv);
    else if (P.opt == KGE_OPT_ADAGRAD) opt_math_t<V, KGE_OPT_ADAGRAD>(P, reset, 0.f, g, wv, mv, vv);
    else if (P.opt == KGE_OPT_MOMENTUM) opt_math_t<V, KGE_OPT_MOMENTUM>(P, reset, 0.f, g, wv, mv, vv);
    else opt_math_t<V, KGE_OPT_SGD>(P, reset, 0.f, g, wv, mv, vv);
}

// gradient of the LP penalty on V columns of a row whose current values are w
template <int V>
__device__ __forceinline__ void reg_add(const ApplyParams& P, bool is_rel, float (&g)[V], const float (&w)[V]) {
    if (P.reg_p <= 0) return;
    const float lam = is_rel ? P.reg_lambda_rel : P.reg_lambda_ent;
#pragma unroll
    for (int x = 0; x < V; ++x) g[x] += reg_grad1(w[x], P.reg_p, lam);
}
__device__ __forceinline__ void mark_touched(const ApplyParams& P, int32_t key) {
    if (P.touched != nullptr) atomicOr(P.touched + (key >> 5), 1u << (key & 31));
}

__device__ __forceinline__ void prefetch_row_state(const ApplyParams& P, int32_t key, bool need_m, bool need_v) {
    const RowPtrs r = resolve_row(P, key);
    if (!r.owned) return;
    const int K = P.ent.K;
    for (int off = 0; off < K; off += 32) {
        prefetch_l2(r.w + off);
        if (need_m && r.m != nullptr) prefetch_l2(r.m + off);
        if (need_v && r.v != nullptr) prefetch_l2(r.v + off);
    }
}
