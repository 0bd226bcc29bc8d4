// Fused KGE training step for sm_100a.
//
//   emit      : counter-based corruption generation (Philox) + per-slot sort keys
//   fwd_bwd   : ONE warp per positive: gathers s,p,o rows with 128-bit coalesced loads, scores the
//               positive and its eta negatives, evaluates the loss and dL/dscore in registers and
//               writes per-slot gradient rows (no [B*eta,K] gathers are ever materialised)
//   apply     : radix sort of (row id, slot) + atomics-free segmented reduction of duplicate rows
//               + row-wise Adam/Adagrad/momentum/SGD on touched rows only
//
// Replaces reference models/EmbeddingModel.py:614-822 (_get_model_loss),
// evaluation/protocol.py:531-659 (generate_corruptions_for_fit), losses/{pairwise,nll,
// nll_multiclass}.py and training/{adam,adagrad,momentum,sgd}.py (+ Keras OptimizerV2 sparse apply).
#include <cub/device/device_radix_sort.cuh>

#include "kge_common.cuh"

// ------------------------------------------------------------------------------------------------
// slot layout for one batch of n positives (S = (3+eta)*n slots, one gradient row per slot)
//   [0,n)            subject row of positive i          key = s_i
//   [n,2n)           object row of positive i           key = o_i
//   [2n,2n+eta*n)    replacement row of negative (j,i)  key = repl[j*n+i]   (slot 2n + j*n + i)
//   [2n+eta*n,S)     relation row of positive i         key = E + p_i
// ------------------------------------------------------------------------------------------------

__global__ void kge_emit_kernel(const int32_t* __restrict__ pos, int64_t n, int eta, int64_t E, int side,
                                const int32_t* __restrict__ repl_in, const uint8_t* __restrict__ keep_in,
                                uint64_t seed, uint64_t step, uint64_t neg_base,
                                int32_t* __restrict__ repl_out, uint8_t* __restrict__ keep_out,
                                int32_t* __restrict__ keys) {
    int64_t S = (int64_t)(3 + eta) * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < S; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t key;
        if (t < n) {
            key = pos[3 * t + 0];
        } else if (t < 2 * n) {
            key = pos[3 * (t - n) + 2];
        } else if (t < 2 * n + (int64_t)eta * n) {
            int64_t q = t - 2 * n;  // j*n + i
            int32_t r;
            uint8_t ks;
            if (repl_in != nullptr) {
                r = repl_in[q];
                ks = keep_in != nullptr ? keep_in[q] : (side == KGE_SIDE_O ? 1 : 0);
            } else {
                uint64_t g = neg_base + (uint64_t)q;
                uint32_t o[4];
                philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32), o);
                // replacement ~ U{0..E-1} (evaluation/protocol.py:616-619); multiply-shift mapping
                r = (int32_t)(((uint64_t)o[0] * (uint64_t)E) >> 32);
                ks = side == KGE_SIDE_SO ? (uint8_t)(o[1] >> 31) : (side == KGE_SIDE_O ? 1 : 0);
            }
            repl_out[q] = r;
            keep_out[q] = ks;
            key = r;
        } else {
            key = (int32_t)E + pos[3 * (t - 2 * n - (int64_t)eta * n) + 1];
        }
        keys[t] = key;
    }
}

// ------------------------------------------------------------------------------------------------
// register-resident embedding rows: lane l owns vectors c = l + 32*i (i < NCH) of V floats; complex
// rows keep the matching imaginary vector (offset k floats) beside the real one.
// ------------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void ld_vec(float (&d)[V], const float* p) {
    if constexpr (V == 4) {
        float4 t = *reinterpret_cast<const float4*>(p);
        d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
    } else {
        d[0] = *p;
    }
}
template <int V>
__device__ __forceinline__ void st_vec(float* p, const float (&d)[V]) {
    if constexpr (V == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
    } else {
        *p = d[0];
    }
}

template <int V, int NCH, bool CPLX>
struct Row {
    float re[NCH][V];
    float im[CPLX ? NCH : 1][V];
};

template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_zero(Row<V, NCH, CPLX>& r) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            r.re[i][v] = 0.f;
            if constexpr (CPLX) r.im[i][v] = 0.f;
        }
}

// nvec: vectors per half (complex) or per row; half: k floats
template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_load(Row<V, NCH, CPLX>& r, const float* __restrict__ base, int lane, int nvec, int half) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        int c = lane + 32 * i;
        if (c < nvec) {
            ld_vec<V>(r.re[i], base + (size_t)c * V);
            if constexpr (CPLX) ld_vec<V>(r.im[i], base + half + (size_t)c * V);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                r.re[i][v] = 0.f;
                if constexpr (CPLX) r.im[i][v] = 0.f;
            }
        }
    }
}

template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_store(float* __restrict__ base, const Row<V, NCH, CPLX>& r, int lane, int nvec, int half) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        int c = lane + 32 * i;
        if (c < nvec) {
            st_vec<V>(base + (size_t)c * V, r.re[i]);
            if constexpr (CPLX) st_vec<V>(base + half + (size_t)c * V, r.im[i]);
        }
    }
}

#define ROW_FOR(i, v)               \
    _Pragma("unroll") for (int i = 0; i < NCH; ++i) \
    _Pragma("unroll") for (int v = 0; v < V; ++v)

// ------------------------------------------------------------------------------------------------
// model algebra.  MODEL: 0 TransE-L1, 1 TransE-L2, 2 DistMult, 3 ComplEx
// (HolE = ComplEx with score scale 2/k, reference models/HolE.py:189).
//   Qo: query for an object-side replacement (score depends on r through <Qo, r> or |Qo - r|)
//   Qs: query for a subject-side replacement
// ------------------------------------------------------------------------------------------------
template <int MODEL, int V, int NCH>
struct Algebra {
    static constexpr bool CPLX = (MODEL == 3);
    static constexpr bool TRANSE = (MODEL == 0 || MODEL == 1);
    using R = Row<V, NCH, CPLX>;

    __device__ __forceinline__ static void queries(const R& s, const R& p, const R& o, R& Qo, R& Qs) {
        ROW_FOR(i, v) {
            if constexpr (TRANSE) {
                Qo.re[i][v] = s.re[i][v] + p.re[i][v];
                Qs.re[i][v] = o.re[i][v] - p.re[i][v];
            } else if constexpr (MODEL == 2) {
                Qo.re[i][v] = s.re[i][v] * p.re[i][v];
                Qs.re[i][v] = p.re[i][v] * o.re[i][v];
            } else {
                Qo.re[i][v] = p.re[i][v] * s.re[i][v] - p.im[i][v] * s.im[i][v];
                Qo.im[i][v] = p.re[i][v] * s.im[i][v] + p.im[i][v] * s.re[i][v];
                Qs.re[i][v] = p.re[i][v] * o.re[i][v] + p.im[i][v] * o.im[i][v];
                Qs.im[i][v] = p.re[i][v] * o.im[i][v] - p.im[i][v] * o.re[i][v];
            }
        }
    }

    // lane-partial of the reduction that defines the score of (Q, r).  For TransE the two sides
    // differ only in the sign of the difference, which |.| and (.)^2 ignore.
    __device__ __forceinline__ static float partial(const R& Q, const R& r) {
        float acc = 0.f;
        ROW_FOR(i, v) {
            if constexpr (MODEL == 0) {
                acc += fabsf(Q.re[i][v] - r.re[i][v]);
            } else if constexpr (MODEL == 1) {
                float d = Q.re[i][v] - r.re[i][v];
                acc = fmaf(d, d, acc);
            } else if constexpr (MODEL == 2) {
                acc = fmaf(Q.re[i][v], r.re[i][v], acc);
            } else {
                acc = fmaf(Q.re[i][v], r.re[i][v], acc);
                acc = fmaf(Q.im[i][v], r.im[i][v], acc);
            }
        }
        return acc;
    }

    __device__ __forceinline__ static float finish(float sum, float scale) {
        if constexpr (MODEL == 0) return -sum;
        else if constexpr (MODEL == 1) return -sqrtf(sum);
        else return scale * sum;
    }

    // Given w = dL/dscore for a candidate row r on side `obj` (1: r stands for the object):
    //   gr  <- gradient row of r
    //   acc <- running accumulator of that side (trilinear: sum w*r ; TransE: sum w*g(u))
    __device__ __forceinline__ static void backward(const R& Q, const R& r, bool obj, float w, float score, float scale,
                                                    R& gr, R& acc) {
        if constexpr (TRANSE) {
            // u = s+p-r (obj) or r+p-o (subj);  Q - r = u (obj) or -u (subj)
            float inv = 0.f;
            if constexpr (MODEL == 1) inv = score != 0.f ? 1.f / (-score) : 0.f;
            float sg = obj ? 1.f : -1.f;
            ROW_FOR(i, v) {
                float u = sg * (Q.re[i][v] - r.re[i][v]);
                float g;  // d f / d u
                if constexpr (MODEL == 0) g = (u > 0.f) ? -1.f : ((u < 0.f) ? 1.f : 0.f);
                else g = -u * inv;
                float wg = w * g;
                acc.re[i][v] += wg;
                gr.re[i][v] = obj ? -wg : wg;  // du/dr = -1 (obj) / +1 (subj)
            }
        } else {
            float ws = w * scale;
            ROW_FOR(i, v) {
                gr.re[i][v] = ws * Q.re[i][v];
                acc.re[i][v] = fmaf(ws, r.re[i][v], acc.re[i][v]);
                if constexpr (CPLX) {
                    gr.im[i][v] = ws * Q.im[i][v];
                    acc.im[i][v] = fmaf(ws, r.im[i][v], acc.im[i][v]);
                }
            }
        }
    }

    // Fold the side accumulators into the gradients of the positive's own rows.
    //   AccO: accumulator over object-side candidates (incl. the positive's own object)
    //   AccS: accumulator over subject-side candidates
    //   go  : on entry the positive's own-object gradient row (from backward()); on exit complete
    __device__ __forceinline__ static void fold(const R& s, const R& p, const R& o, const R& AccO, const R& AccS,
                                                R& gs, R& gp, R& go) {
        ROW_FOR(i, v) {
            if constexpr (TRANSE) {
                gs.re[i][v] = AccO.re[i][v];
                gp.re[i][v] = AccO.re[i][v] + AccS.re[i][v];
                go.re[i][v] = go.re[i][v] - AccS.re[i][v];
            } else if constexpr (MODEL == 2) {
                gs.re[i][v] = p.re[i][v] * AccO.re[i][v];
                gp.re[i][v] = s.re[i][v] * AccO.re[i][v] + o.re[i][v] * AccS.re[i][v];
                go.re[i][v] = fmaf(p.re[i][v], AccS.re[i][v], go.re[i][v]);
            } else {
                float pr = p.re[i][v], pi = p.im[i][v];
                float ar = AccO.re[i][v], ai = AccO.im[i][v];
                float br = AccS.re[i][v], bi = AccS.im[i][v];
                gs.re[i][v] = pr * ar + pi * ai;
                gs.im[i][v] = pr * ai - pi * ar;
                gp.re[i][v] = s.re[i][v] * ar + s.im[i][v] * ai + br * o.re[i][v] + bi * o.im[i][v];
                gp.im[i][v] = s.re[i][v] * ai - s.im[i][v] * ar + br * o.im[i][v] - bi * o.re[i][v];
                go.re[i][v] += pr * br - pi * bi;
                go.im[i][v] += pr * bi + pi * br;
            }
        }
    }
};

__device__ __forceinline__ float clip75(float x) { return fminf(fmaxf(x, -75.f), 75.f); }

struct FwdBwdParams {
    TableView ent;
    const float* rel;
    const int32_t* pos;
    const int32_t* repl;
    const uint8_t* keep;
    int64_t n;
    int eta, k, loss;
    float margin, scale;
    float* grad_rows;   // [S,K]
    float* loss_part;   // [n]
    float* dbg_scores;  // optional [n*(1+eta)]
};

// One warp per positive.  U candidate rows are kept in flight to cover L2/HBM latency.
template <int MODEL, int V, int NCH, int U>
__global__ void __launch_bounds__(128) kge_fwd_bwd_kernel(FwdBwdParams P) {
    using A = Algebra<MODEL, V, NCH>;
    using R = typename A::R;
    extern __shared__ float s_sc[];  // [warps][eta]: negative scores (multiclass two-pass)

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= P.n) return;
    const int K = P.ent.K;
    const int half = A::CPLX ? P.k : 0;
    const int nvec = (A::CPLX ? P.k : K) / V;
    const int64_t n = P.n;
    const int eta = P.eta;
    float* sc = s_sc + (size_t)wib * eta;

    const int32_t si = P.pos[3 * i + 0], pi = P.pos[3 * i + 1], oi = P.pos[3 * i + 2];
    R s, p, o, Qo, Qs;
    row_load(s, table_row(P.ent, si), lane, nvec, half);
    row_load(p, P.rel + (size_t)pi * K, lane, nvec, half);
    row_load(o, table_row(P.ent, oi), lane, nvec, half);
    A::queries(s, p, o, Qo, Qs);
    const float spos = A::finish(warp_sum(A::partial(Qo, o)), P.scale);

    R AccO, AccS;
    row_zero(AccO);
    row_zero(AccS);

    float* G = P.grad_rows;
    const int64_t slot_neg0 = 2 * n;
    float loss_acc = 0.f;  // identical on all lanes
    float wpos = 0.f;
    const int loss = P.loss;
    const float cpos = clip75(spos);
    const bool pos_in = (spos >= -75.f) && (spos <= 75.f);

    if (loss == KGE_LOSS_MULTICLASS_NLL) {
        // pass 1: scores of all negatives (nll_multiclass.py:70-81 needs the full softmax denominator)
        for (int j0 = 0; j0 < eta; j0 += U) {
            R r[U];
            bool ob[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + u;
                if (j < eta) {
                    int64_t q = (int64_t)j * n + i;
                    ob[u] = P.keep[q] != 0;
                    row_load(r[u], table_row(P.ent, P.repl[q]), lane, nvec, half);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + u;
                if (j < eta) {
                    float sn = A::finish(warp_sum(A::partial(ob[u] ? Qo : Qs, r[u])), P.scale);
                    if (lane == 0) sc[j] = sn;
                }
            }
        }
        __syncwarp();
        float zpart = 0.f;
        for (int j = lane; j < eta; j += 32) zpart += expf(clip75(sc[j]));
        const float pe = expf(cpos);
        const float z = warp_sum(zpart) + pe;
        loss_acc = -logf(pe / z);
        wpos = pos_in ? -(1.f - pe / z) : 0.f;
        const float invz = 1.f / z;
        // pass 2: gradients (rows re-read from L1/L2)
        for (int j0 = 0; j0 < eta; j0 += U) {
            R r[U];
            bool ob[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + u;
                if (j < eta) {
                    int64_t q = (int64_t)j * n + i;
                    ob[u] = P.keep[q] != 0;
                    row_load(r[u], table_row(P.ent, P.repl[q]), lane, nvec, half);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + u;
                if (j < eta) {
                    float sn = sc[j];
                    bool in = (sn >= -75.f) && (sn <= 75.f);
                    float w = in ? expf(sn) * invz : 0.f;
                    R gr;
                    A::backward(ob[u] ? Qo : Qs, r[u], ob[u], w, sn, P.scale, gr, ob[u] ? AccO : AccS);
                    row_store(G + (size_t)(slot_neg0 + (int64_t)j * n + i) * K, gr, lane, nvec, half);
                }
            }
        }
    } else {
        // pairwise / nll: dL/dneg depends on (pos, neg) only -> single pass
        const float margin = P.margin;
        for (int j0 = 0; j0 < eta; j0 += U) {
            R r[U];
            bool ob[U];
            float part[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + u;
                if (j < eta) {
                    int64_t q = (int64_t)j * n + i;
                    ob[u] = P.keep[q] != 0;
                    row_load(r[u], table_row(P.ent, P.repl[q]), lane, nvec, half);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) part[u] = (j0 + u < eta) ? A::partial(ob[u] ? Qo : Qs, r[u]) : 0.f;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1)
#pragma unroll
                for (int u = 0; u < U; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o2);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int j = j0 + u;
                if (j < eta) {
                    float sn = A::finish(part[u], P.scale);
                    float w;
                    if (loss == KGE_LOSS_PAIRWISE) {
                        // losses/pairwise.py:69 ; tf.maximum passes the gradient when t >= 0
                        float t = margin - spos + sn;
                        loss_acc += fmaxf(t, 0.f);
                        w = (t >= 0.f) ? 1.f : 0.f;
                        wpos -= w;
                    } else {
                        // losses/nll.py:55-59 : log(1+exp(clip(neg)))
                        float cn = clip75(sn);
                        float e = expf(cn);
                        loss_acc += logf(1.f + e);
                        bool in = (sn >= -75.f) && (sn <= 75.f);
                        w = in ? e / (1.f + e) : 0.f;
                    }
                    if (P.dbg_scores != nullptr && lane == 0) P.dbg_scores[n + (int64_t)j * n + i] = sn;
                    R gr;
                    A::backward(ob[u] ? Qo : Qs, r[u], ob[u], w, sn, P.scale, gr, ob[u] ? AccO : AccS);
                    row_store(G + (size_t)(slot_neg0 + (int64_t)j * n + i) * K, gr, lane, nvec, half);
                }
            }
        }
        if (loss == KGE_LOSS_NLL) {
            // positives are tiled eta times (models/EmbeddingModel.py:724-729)
            float e = expf(-cpos);
            loss_acc += (float)eta * logf(1.f + e);
            wpos = pos_in ? -(float)eta * (e / (1.f + e)) : 0.f;
        }
    }
    if (loss == KGE_LOSS_MULTICLASS_NLL && P.dbg_scores != nullptr) {
        for (int j = lane; j < eta; j += 32) P.dbg_scores[n + (int64_t)j * n + i] = sc[j];
    }

    // the positive itself: an object-side candidate with r = o and weight dL/dpos
    R gs, gp, go;
    A::backward(Qo, o, true, wpos, spos, P.scale, go, AccO);
    A::fold(s, p, o, AccO, AccS, gs, gp, go);
    row_store(G + (size_t)i * K, gs, lane, nvec, half);
    row_store(G + (size_t)(n + i) * K, go, lane, nvec, half);
    row_store(G + (size_t)(slot_neg0 + (int64_t)eta * n + i) * K, gp, lane, nvec, half);
    if (lane == 0) {
        P.loss_part[i] = loss_acc;
        if (P.dbg_scores != nullptr) P.dbg_scores[i] = spos;
    }
}

// deterministic fixed-order reduction of the per-positive loss terms
__global__ void kge_loss_reduce_kernel(const float* __restrict__ part, int64_t n, float* __restrict__ out) {
    __shared__ double sm[1024];
    double acc = 0.0;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) acc += (double)part[t];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)sm[0];
}

// ------------------------------------------------------------------------------------------------
// segmented reduction of duplicate rows + sparse row-wise optimizer
// ------------------------------------------------------------------------------------------------
struct ApplyParams {
    const int32_t* keys;   // sorted
    const int32_t* slots;  // sorted alongside
    int64_t n_keys;
    TableView grads;       // shard r = rank r's gradient rows, rows_per_shard = slots per rank
    TableView ent, ent_m, ent_v;
    float *rel, *rel_m, *rel_v;
    int64_t E, R;
    int64_t row_begin, row_end;  // owned entity rows
    int opt;
    uint32_t flags;
    float lr, lr_t, beta1, beta2, eps, momentum;
    float* dbg_grad_ent;
    float* dbg_grad_rel;
};

// One warp per run of equal keys: the head warp walks its run (sorted order => fixed summation
// order => bit-reproducible), other warps exit.  Lanes own columns; no atomics, no shuffles in the
// inner loop; 4 gradient rows are kept in flight.
template <int V>
__global__ void __launch_bounds__(256) kge_apply_kernel(ApplyParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= P.n_keys) return;
    const int32_t key = P.keys[t];
    if (t > 0 && P.keys[t - 1] == key) return;  // not a run head
    const bool is_rel = key >= P.E;
    const int64_t row = is_rel ? key - P.E : key;
    if (!is_rel && (row < P.row_begin || row >= P.row_end)) return;
    const int K = P.ent.K;
    float *w, *m = nullptr, *v = nullptr;
    if (is_rel) {
        w = P.rel + (size_t)row * K;
        if (P.rel_m) m = P.rel_m + (size_t)row * K;
        if (P.rel_v) v = P.rel_v + (size_t)row * K;
    } else {
        w = table_row(P.ent, row);
        if (P.ent_m.shard[0]) m = table_row(P.ent_m, row);
        if (P.ent_v.shard[0]) v = table_row(P.ent_v, row);
    }
    // run length (keys are sorted; runs are short except for hub entities)
    int64_t end = t + 1;
    while (end < P.n_keys && P.keys[end] == key) ++end;

    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (P.flags & KGE_F_NO_UPDATE) != 0;
    for (int c0 = lane * V; c0 < K; c0 += 32 * V) {
        float g[V];
#pragma unroll
        for (int x = 0; x < V; ++x) g[x] = 0.f;
        int64_t u = t;
        for (; u + 4 <= end; u += 4) {
            float a[4][V];
#pragma unroll
            for (int q = 0; q < 4; ++q) ld_vec<V>(a[q], table_row(P.grads, P.slots[u + q]) + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] += a[q][x];
        }
        for (; u < end; ++u) {
            float a[V];
            ld_vec<V>(a, table_row(P.grads, P.slots[u]) + c0);
#pragma unroll
            for (int x = 0; x < V; ++x) g[x] += a[x];
        }
        if (is_rel ? (P.dbg_grad_rel != nullptr) : (P.dbg_grad_ent != nullptr)) {
            float* d = (is_rel ? P.dbg_grad_rel : P.dbg_grad_ent) + (size_t)row * K + c0;
            st_vec<V>(d, g);
        }
        if (no_update) continue;
        float wv[V], mv[V], vv[V];
        ld_vec<V>(wv, w + c0);
        if (P.opt == KGE_OPT_ADAM) {
            // Keras Adam (beta1 .9, beta2 .999, eps 1e-7): var -= lr_t * m / (sqrt(v) + eps)
            if (reset) {
#pragma unroll
                for (int x = 0; x < V; ++x) mv[x] = vv[x] = 0.f;
            } else {
                ld_vec<V>(mv, m + c0);
                ld_vec<V>(vv, v + c0);
            }
#pragma unroll
            for (int x = 0; x < V; ++x) {
                mv[x] = P.beta1 * mv[x] + (1.f - P.beta1) * g[x];
                vv[x] = P.beta2 * vv[x] + (1.f - P.beta2) * g[x] * g[x];
                wv[x] = wv[x] - P.lr_t * mv[x] / (sqrtf(vv[x]) + P.eps);
            }
            if (m) st_vec<V>(m + c0, mv);
            if (v) st_vec<V>(v + c0, vv);
        } else if (P.opt == KGE_OPT_ADAGRAD) {
            // Keras Adagrad: accumulator starts at 0.1; var -= lr * g / (sqrt(acc) + eps)
            if (reset) {
#pragma unroll
                for (int x = 0; x < V; ++x) mv[x] = 0.1f;
            } else {
                ld_vec<V>(mv, m + c0);
            }
#pragma unroll
            for (int x = 0; x < V; ++x) {
                mv[x] = mv[x] + g[x] * g[x];
                wv[x] = wv[x] - P.lr * g[x] / (sqrtf(mv[x]) + P.eps);
            }
            if (m) st_vec<V>(m + c0, mv);
        } else if (P.opt == KGE_OPT_MOMENTUM) {
            // Keras SGD momentum: vel = mu*vel - lr*g ; var += vel
            if (reset) {
#pragma unroll
                for (int x = 0; x < V; ++x) mv[x] = 0.f;
            } else {
                ld_vec<V>(mv, m + c0);
            }
#pragma unroll
            for (int x = 0; x < V; ++x) {
                mv[x] = P.momentum * mv[x] - P.lr * g[x];
                wv[x] = wv[x] + mv[x];
            }
            if (m) st_vec<V>(m + c0, mv);
        } else {
#pragma unroll
            for (int x = 0; x < V; ++x) wv[x] = wv[x] - P.lr * g[x];
        }
        st_vec<V>(w + c0, wv);
    }
}

__global__ void kge_iota_kernel(int32_t* v, int64_t n) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        v[t] = (int32_t)t;
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
    KGE_REQUIRE(a->loss >= KGE_LOSS_PAIRWISE && a->loss <= KGE_LOSS_MULTICLASS_NLL, "Unsupported loss function: %d", a->loss);
    KGE_REQUIRE(a->opt >= KGE_OPT_ADAM && a->opt <= KGE_OPT_SGD, "Unsupported optimizer: %d", a->opt);
    KGE_REQUIRE(a->side >= KGE_SIDE_SO && a->side <= KGE_SIDE_O, "Invalid corruption side %d", a->side);
    KGE_REQUIRE(a->k > 0 && a->eta > 0, "kge_train: k and eta must be positive");
    KGE_REQUIRE(a->ent.K == model_row_width(a->model, a->k), "kge_train: table width %d != internal_k %d", a->ent.K,
                model_row_width(a->model, a->k));
    KGE_REQUIRE(a->ent.n_shards >= 1 && a->ent.n_shards <= KGE_MAX_SHARDS, "kge_train: bad shard count");
    KGE_REQUIRE(a->ent.rows + a->R < (int64_t)INT32_MAX, "kge_train: E+R must fit int32 sort keys");
    KGE_REQUIRE(a->n_pos >= 0 && (a->pos != nullptr || a->n_pos == 0), "kge_train: positives missing");
    return 0;
}

template <int MODEL, int V>
static int launch_fwd_bwd_nch(int nch, const FwdBwdParams& P, cudaStream_t st) {
    const int warps = 4;
    dim3 grid((unsigned)((P.n + warps - 1) / warps)), block(warps * 32);
    size_t smem = (size_t)warps * P.eta * sizeof(float);
    switch (nch) {
        case 1: kge_fwd_bwd_kernel<MODEL, V, 1, 4><<<grid, block, smem, st>>>(P); break;
        case 2: kge_fwd_bwd_kernel<MODEL, V, 2, 4><<<grid, block, smem, st>>>(P); break;
        case 3:
        case 4: kge_fwd_bwd_kernel<MODEL, V, 4, 2><<<grid, block, smem, st>>>(P); break;
        case 5:
        case 6:
        case 7:
        case 8: kge_fwd_bwd_kernel<MODEL, V, 8, 1><<<grid, block, smem, st>>>(P); break;
        default: kge_set_error("kge_train: embedding size too large for the fused kernel (chunks/lane=%d)", nch); return -1;
    }
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int MODEL>
static int launch_fwd_bwd_model(const FwdBwdParams& P, cudaStream_t st) {
    const bool cplx = (MODEL == 3);
    const int width = cplx ? P.k : P.ent.K;  // floats per half / per row
    if (width % 4 == 0) {
        int nvec = width / 4;
        return launch_fwd_bwd_nch<MODEL, 4>((nvec + 31) / 32, P, st);
    }
    return launch_fwd_bwd_nch<MODEL, 1>((width + 31) / 32, P, st);
}

static int ensure_train_ws(kge_ctx* ctx, const kge_train_args* a, bool need_grads) {
    int64_t n = a->n_pos, S = (int64_t)(3 + a->eta) * n;
    if (ctx->repl.reserve((size_t)a->eta * n * sizeof(int32_t))) return -2;
    if (ctx->keep.reserve((size_t)a->eta * n)) return -2;
    if (ctx->loss_part.reserve((size_t)(n + 1) * sizeof(float))) return -2;
    if (need_grads && ctx->grad_rows.reserve((size_t)S * a->ent.K * sizeof(float))) return -2;
    return 0;
}

extern "C" int64_t kge_train_grad_rows(int eta, int64_t n_pos) { return (int64_t)(3 + eta) * n_pos; }

extern "C" int kge_train_emit(kge_ctx* ctx, const kge_train_args* a, int32_t* keys_out, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_emit: null ctx");
    if (int rc = validate_train(a)) return rc;
    if (a->n_pos == 0) return 0;
    if (int rc = ensure_train_ws(ctx, a, false)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    int threads = 256;
    int blocks = (int)std::min<int64_t>((S + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
    kge_emit_kernel<<<blocks, threads, 0, st>>>(a->pos, a->n_pos, a->eta, a->ent.rows, a->side, a->repl, a->keep_subj,
                                                a->seed, a->step, a->neg_index_base, ctx->repl.as<int32_t>(),
                                                ctx->keep.as<uint8_t>(), keys_out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_train_fwd_bwd(kge_ctx* ctx, const kge_train_args* a, float* grad_rows, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_fwd_bwd: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(grad_rows != nullptr, "kge_train_fwd_bwd: grad_rows missing");
    KGE_REQUIRE(a->loss_out != nullptr, "kge_train_fwd_bwd: loss_out missing");
    cudaStream_t st = (cudaStream_t)stream;
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
    P.scale = a->model == KGE_HOLE ? 2.0f / (float)a->k : 1.0f;
    P.grad_rows = grad_rows;
    P.loss_part = ctx->loss_part.as<float>();
    P.dbg_scores = a->dbg_scores;
    int rc;
    switch (a->model) {
        case KGE_TRANSE_L1: rc = launch_fwd_bwd_model<0>(P, st); break;
        case KGE_TRANSE_L2: rc = launch_fwd_bwd_model<1>(P, st); break;
        case KGE_DISTMULT: rc = launch_fwd_bwd_model<2>(P, st); break;
        default: rc = launch_fwd_bwd_model<3>(P, st); break;
    }
    if (rc) return rc;
    kge_loss_reduce_kernel<<<1, 1024, 0, st>>>(ctx->loss_part.as<float>(), a->n_pos, a->loss_out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_train_apply(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                               const kge_table* grads, int64_t row_begin, int64_t row_end, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_apply: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(grads != nullptr && keys_all != nullptr, "kge_train_apply: missing keys/grads");
    if (n_keys == 0) return 0;
    KGE_REQUIRE(n_keys < (int64_t)INT32_MAX, "kge_train_apply: too many slots");
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->keys_out.reserve((size_t)n_keys * 4)) return -2;
    if (ctx->vals_in.reserve((size_t)n_keys * 4)) return -2;
    if (ctx->vals_out.reserve((size_t)n_keys * 4)) return -2;
    int64_t E = a->ent.rows;
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < E + a->R) ++end_bit;
    size_t tmp_bytes = 0;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_all, ctx->keys_out.as<int32_t>(),
                                                   ctx->vals_in.as<int32_t>(), ctx->vals_out.as<int32_t>(), (int)n_keys, 0,
                                                   end_bit, st));
    if (ctx->sort_tmp.reserve(tmp_bytes)) return -2;
    {
        int threads = 256;
        int blocks = (int)std::min<int64_t>((n_keys + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
        kge_iota_kernel<<<blocks, threads, 0, st>>>(ctx->vals_in.as<int32_t>(), n_keys);
    }
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, keys_all, ctx->keys_out.as<int32_t>(),
                                                   ctx->vals_in.as<int32_t>(), ctx->vals_out.as<int32_t>(), (int)n_keys, 0,
                                                   end_bit, st));
    ApplyParams P;
    P.keys = ctx->keys_out.as<int32_t>();
    P.slots = ctx->vals_out.as<int32_t>();
    P.n_keys = n_keys;
    P.grads = make_view(*grads);
    P.ent = make_view(a->ent);
    P.ent_m = make_view(a->ent_m);
    P.ent_v = make_view(a->ent_v);
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
    P.dbg_grad_ent = a->dbg_grad_ent;
    P.dbg_grad_rel = a->dbg_grad_rel;
    const bool reset = (a->flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (a->flags & KGE_F_NO_UPDATE) != 0;
    double t = reset ? 1.0 : (double)(a->step < 1 ? 1 : a->step);
    P.lr_t = (float)((double)a->lr * sqrt(1.0 - pow((double)a->beta2, t)) / (1.0 - pow((double)a->beta1, t)));
    if (!no_update && !reset) {
        if (a->opt == KGE_OPT_ADAM)
            KGE_REQUIRE(a->ent_m.shard[0] && a->ent_v.shard[0] && a->rel_m && a->rel_v, "kge_train: adam state (m,v) missing");
        if (a->opt == KGE_OPT_ADAGRAD || a->opt == KGE_OPT_MOMENTUM)
            KGE_REQUIRE(a->ent_m.shard[0] && a->rel_m, "kge_train: optimizer state missing");
    }
    const int warps = 8;
    dim3 grid((unsigned)((n_keys + warps - 1) / warps)), block(warps * 32);
    if (a->ent.K % 4 == 0) kge_apply_kernel<4><<<grid, block, 0, st>>>(P);
    else kge_apply_kernel<1><<<grid, block, 0, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_train_step(kge_ctx* ctx, const kge_train_args* a, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_step: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(a->ent.n_shards == 1, "kge_train_step is the single-GPU entry; use the phased calls when sharded");
    if (a->n_pos == 0) return 0;
    int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    if (int rc = ensure_train_ws(ctx, a, true)) return rc;
    if (ctx->keys_in.reserve((size_t)S * 4)) return -2;
    if (int rc = kge_train_emit(ctx, a, ctx->keys_in.as<int32_t>(), stream)) return rc;
    if (int rc = kge_train_fwd_bwd(ctx, a, ctx->grad_rows.as<float>(), stream)) return rc;
    kge_table g;
    memset(&g, 0, sizeof(g));
    g.shard[0] = ctx->grad_rows.as<float>();
    g.rows = S;
    g.rows_per_shard = S;
    g.n_shards = 1;
    g.K = a->ent.K;
    return kge_train_apply(ctx, a, ctx->keys_in.as<int32_t>(), S, &g, 0, a->ent.rows, stream);
}

extern "C" int kge_train_step_host(kge_ctx* ctx, const kge_train_args* a, const int32_t* pos_host, float* loss_host,
                                   void* stream) {
    KGE_REQUIRE(ctx != nullptr && a != nullptr, "kge_train_step_host: null argument");
    KGE_REQUIRE(a->n_pos >= 0 && (a->n_pos == 0 || pos_host != nullptr), "kge_train_step_host: positives missing");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_pos == 0) {
        if (loss_host) *loss_host = 0.f;
        return 0;
    }
    if (ctx->h_pos.reserve((size_t)a->n_pos * 3 * sizeof(int32_t)) || ctx->h_loss.reserve(sizeof(float))) return -2;
    KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pos.p, pos_host, (size_t)a->n_pos * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    kge_train_args b = *a;
    b.pos = ctx->h_pos.as<int32_t>();
    if (b.loss_out == nullptr) b.loss_out = ctx->h_loss.as<float>();
    if (int rc = kge_train_step(ctx, &b, stream)) return rc;
    if (loss_host) KGE_CUDA_CHECK(cudaMemcpyAsync(loss_host, b.loss_out, sizeof(float), cudaMemcpyDeviceToHost, st));
    KGE_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int kge_normalize_rows(kge_ctx* ctx, float* emb, int64_t rows, int K, void* stream) {
    KGE_REQUIRE(ctx != nullptr && emb != nullptr, "kge_normalize_rows: null argument");
    if (rows == 0) return 0;
    const int warps = 8;
    kge_normalize_rows_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(emb, rows, K);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}
