// Fused forward + loss + backward of one training batch (sm_100a), templated on the scoring model.
//
// One warp (or SPLIT cooperating warps of one CTA) per positive triple:
//   * s, p, o rows are gathered with 128-bit coalesced loads and folded into the two "queries"
//     Qo (object-side candidates) and Qs (subject-side candidates) that stay in registers;
//   * the eta replacement ids of the positive are prefetched by the lanes in one coalesced round and
//     broadcast with shuffles, so U candidate rows are in flight per warp with no index->row
//     dependent-load chain;
//   * score, loss term and dL/dscore are evaluated in registers; the candidate row is folded into a
//     side accumulator (AccO / AccS) from which the gradients of the positive's own rows follow;
//   * NO per-negative gradient row is written.  The gradient of a replacement row r of negative (j,i)
//     is a function of (c_ji, Q_side(i), r):  c*Q (DistMult/ComplEx/HolE), c*sign(Q-r) (TransE L1),
//     c*(Q-r) (TransE L2).  Only the scalar c_ji and the two query rows per positive are stored; the
//     segmented reduction in kge_train.cu re-materialises the row from them.
//
// Per positive the kernel writes 5 rows (gs, go, gp, Qo, Qs) + eta coefficients instead of 3+eta
// rows.  Replaces reference models/EmbeddingModel.py:614-822 (_get_model_loss), losses/*.py and the
// GradientTape backward of training/adam.py:45-46.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "kge_common.cuh"

// ------------------------------------------------------------------------------------------------
// gradient buffer of one rank (caller-owned so that peers can map it), n = positives in the batch:
//   float rows[5n][K] : [0,n) gs | [n,2n) go | [2n,3n) gp | [3n,4n) Qo | [4n,5n) Qs
//   float coef[eta*n] : c_ji at j*n+i
//   uint8 keep[eta*n] : 1 = subject kept (object replaced -> query Qo)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t gbuf_floats(int eta, int64_t n, int K) {
    return 5 * n * (int64_t)K + (int64_t)eta * n + ((int64_t)eta * n + 3) / 4;
}
__host__ __device__ inline float* gbuf_coef(float* base, int64_t n, int K) { return base + 5 * n * (int64_t)K; }
__host__ __device__ inline uint8_t* gbuf_keep(float* base, int eta, int64_t n, int K) {
    return reinterpret_cast<uint8_t*>(gbuf_coef(base, n, K) + (int64_t)eta * n);
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

#define ROW_FOR(i, v)                               \
    _Pragma("unroll") for (int i = 0; i < NCH; ++i) \
    _Pragma("unroll") for (int v = 0; v < V; ++v)

template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_zero(Row<V, NCH, CPLX>& r) {
    ROW_FOR(i, v) {
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

// branch-free variant: lanes past the end of the row read the row's last vector (a valid duplicate)
template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_load_clamped(Row<V, NCH, CPLX>& r, const float* __restrict__ base, int lane, int nvec, int half) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = min(lane + 32 * i, nvec - 1);
        ld_vec<V>(r.re[i], base + (size_t)c * V);
        if constexpr (CPLX) ld_vec<V>(r.im[i], base + half + (size_t)c * V);
    }
}

// zero the vectors of the lanes past the end of the row
template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_mask(Row<V, NCH, CPLX>& r, int lane, int nvec) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        if (lane + 32 * i >= nvec) {
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

template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_add(Row<V, NCH, CPLX>& a, const Row<V, NCH, CPLX>& b) {
    ROW_FOR(i, v) {
        a.re[i][v] += b.re[i][v];
        if constexpr (CPLX) a.im[i][v] += b.im[i][v];
    }
}

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
    // differ only in the sign of the difference, which |.| and (.)^2 ignore.  msk[i] is 1 for the
    // lanes that hold real columns of chunk i and 0 past the end of the row: the trilinear models
    // do not need it (their queries are zero there), the distances do.  V independent accumulators
    // keep the FMA dependency chains short.
    __device__ __forceinline__ static float partial(const R& Q, const R& r, const float (&msk)[NCH]) {
        float a0 = 0.f, a1 = 0.f;  // two independent chains
        ROW_FOR(i, v) {
            float& acc = (v & 1) ? a1 : a0;
            if constexpr (MODEL == 0) {
                acc = fmaf(msk[i], fabsf(Q.re[i][v] - r.re[i][v]), acc);
            } else if constexpr (MODEL == 1) {
                float d = (Q.re[i][v] - r.re[i][v]) * msk[i];
                acc = fmaf(d, d, acc);
            } else if constexpr (MODEL == 2) {
                acc = fmaf(Q.re[i][v], r.re[i][v], acc);
            } else {
                a0 = fmaf(Q.re[i][v], r.re[i][v], a0);
                a1 = fmaf(Q.im[i][v], r.im[i][v], a1);
            }
        }
        return a0 + a1;
    }

    __device__ __forceinline__ static float finish(float sum, float scale) {
        if constexpr (MODEL == 0) return -sum;
        else if constexpr (MODEL == 1) return -sqrtf(sum);
        else return scale * sum;
    }

    // coefficient stored for the replacement row of a negative with dL/dscore = w:
    //   gradient row = c*Q (trilinear) | c*sign(Q-r) (L1) | c*(Q-r) (L2)
    __device__ __forceinline__ static float coefficient(float w, float score, float scale) {
        if constexpr (MODEL == 0) return w;
        else if constexpr (MODEL == 1) return score != 0.f ? w / (-score) : 0.f;
        else return w * scale;
    }

    // acc += w * d score / d(the positive's side of the triple), expressed through the candidate r:
    //   trilinear: acc += w*scale*r ;  TransE: acc += w*g(u), u = s+p-r (obj) or r+p-o (subj)
    __device__ __forceinline__ static void accumulate(const R& Q, const R& r, bool obj, float w, float score, float scale, R& acc) {
        if constexpr (TRANSE) {
            float inv = 0.f;
            if constexpr (MODEL == 1) inv = score != 0.f ? 1.f / (-score) : 0.f;
            const float sg = obj ? 1.f : -1.f;
            ROW_FOR(i, v) {
                float u = sg * (Q.re[i][v] - r.re[i][v]);
                float g;  // d f / d u
                if constexpr (MODEL == 0) g = (u > 0.f) ? -1.f : ((u < 0.f) ? 1.f : 0.f);
                else g = -u * inv;
                acc.re[i][v] = fmaf(w, g, acc.re[i][v]);
            }
        } else {
            const float ws = w * scale;
            ROW_FOR(i, v) {
                acc.re[i][v] = fmaf(ws, r.re[i][v], acc.re[i][v]);
                if constexpr (CPLX) acc.im[i][v] = fmaf(ws, r.im[i][v], acc.im[i][v]);
            }
        }
    }

    // the positive's own object as an object-side candidate with weight w = dL/dpos:
    //   go <- gradient row of o from this term ; AccO updated like accumulate()
    __device__ __forceinline__ static void backward_pos(const R& Qo, const R& o, float w, float score, float scale, R& go, R& AccO) {
        if constexpr (TRANSE) {
            float inv = 0.f;
            if constexpr (MODEL == 1) inv = score != 0.f ? 1.f / (-score) : 0.f;
            ROW_FOR(i, v) {
                float u = Qo.re[i][v] - o.re[i][v];
                float g;
                if constexpr (MODEL == 0) g = (u > 0.f) ? -1.f : ((u < 0.f) ? 1.f : 0.f);
                else g = -u * inv;
                float wg = w * g;
                AccO.re[i][v] += wg;
                go.re[i][v] = -wg;
            }
        } else {
            const float ws = w * scale;
            ROW_FOR(i, v) {
                go.re[i][v] = ws * Qo.re[i][v];
                AccO.re[i][v] = fmaf(ws, o.re[i][v], AccO.re[i][v]);
                if constexpr (CPLX) {
                    go.im[i][v] = ws * Qo.im[i][v];
                    AccO.im[i][v] = fmaf(ws, o.im[i][v], AccO.im[i][v]);
                }
            }
        }
    }

    // Fold the side accumulators into the gradients of the positive's own rows.
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
// log(sigmoid(x)) = -softplus(-x) (tf.math.log_sigmoid) and sigmoid(x)
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }
// losses whose per-negative weight depends on all eta scores of the positive: two passes over the rows
__device__ __forceinline__ bool loss_two_pass(int loss) { return loss == KGE_LOSS_MULTICLASS_NLL || loss == KGE_LOSS_SELF_ADVERSARIAL; }

struct FwdBwdParams {
    TableView ent;
    const float* rel;
    const int32_t* pos;
    const int32_t* repl;
    const uint8_t* keep;
    int64_t n;
    int eta, k, loss;
    float margin, scale, alpha;
    int nl;             // KGE_NL_*: non-linearity on the scores before the loss
    float* gbuf;        // gradient buffer (layout above)
    float* loss_part;   // [n]
    float* dbg_scores;  // optional [n*(1+eta)]
    const float* stage; // optional local [(2+eta)*n][K]: entity row of slot t (subjects, objects, replacements)
    int l2_prefetch;    // 1: L2-prefetch the candidate rows that do not fit the bulk-copy ring window yet
};

// entity row `idx` that sits in entity slot `slot` of this batch: from the staging copy when the
// owners pushed it (row-sharded multi-GPU), else from the table
__device__ __forceinline__ const float* slot_row(const FwdBwdParams& P, int64_t idx, int64_t slot) {
    return P.stage != nullptr ? P.stage + slot * (int64_t)P.ent.K : table_row(P.ent, idx);
}

// shared-memory row loads (the staged rows of the PIPE variant); lanes past the end of the row read
// the row's last vector like row_load_clamped
template <int V, int NCH, bool CPLX>
__device__ __forceinline__ void row_load_smem(Row<V, NCH, CPLX>& r, const float* base, int lane, int nvec, int half) {
    static_assert(V == 4, "staged rows are 128-bit vectors");
    const uint32_t b = smem_u32(base);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = min(lane + 32 * i, nvec - 1);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.re[i][0]), "=f"(r.re[i][1]), "=f"(r.re[i][2]), "=f"(r.re[i][3])
                     : "r"(b + (uint32_t)c * 16u));
        if constexpr (CPLX)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(r.im[i][0]), "=f"(r.im[i][1]), "=f"(r.im[i][2]), "=f"(r.im[i][3])
                         : "r"(b + (uint32_t)(half + c * 4) * 4u));
    }
}

// SPLIT warps cooperate on one positive (its negatives are dealt round-robin); a CTA is 4 warps.
// registers of one row per lane; bounds the resident CTAs the compiler is asked to allow
template <int MODEL, int V, int NCH>
struct RowRegs {
    static constexpr int value = NCH * V * (MODEL == 3 ? 2 : 1);
    static constexpr int min_ctas = value <= 16 ? 3 : (value <= 32 ? 2 : 1);
};

// PIPE: candidate rows are staged in shared memory by 1-D bulk async copies (TMA unit,
// cp.async.bulk -> SASS UBLKCP).  Every warp owns a ring of 2^ns_log2 row slots with one mbarrier
// each: the lane that holds a negative's replacement id issues the copy of that row as soon as its
// position in the processing order enters the ring window.  Candidates are processed in groups of
// KGE_FB_G rows that are all resident at once:
//   pass 1  lane-partial dot products of the G rows, then ONE transposed warp reduction (6 shuffles
//           for 4 candidates instead of 20) that leaves candidate c's score on the lanes with
//           (bit4,bit3) == c
//   loss    exp/log/divide evaluated once per group (each lane for the candidate it holds)
//   pass 2  the rows are read again from shared memory (no registers held across the reduction) and
//           folded into the side accumulator with the weight broadcast from the holder lane
// Local tables only (peer shards keep the register path).
#define KGE_FB_G 4
template <int MODEL, int V, int NCH, bool PIPE>
struct FwdOcc {
    static constexpr int value = PIPE ? (RowRegs<MODEL, V, NCH>::value <= 16 ? 4 : (RowRegs<MODEL, V, NCH>::value <= 32 ? 2 : 1))
                                      : RowRegs<MODEL, V, NCH>::min_ctas;
};

template <int MODEL, int V, int NCH, int U, int SPLIT, bool PIPE>
__global__ void __launch_bounds__(128, (FwdOcc<MODEL, V, NCH, PIPE>::value)) kge_fwd_bwd_kernel(FwdBwdParams P, int ns_log2) {
    using A = Algebra<MODEL, V, NCH>;
    using R = typename A::R;
    constexpr int PP = 4 / SPLIT;  // positives per CTA
    constexpr int G = KGE_FB_G;
    extern __shared__ __align__(16) float smem[];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = wib % SPLIT;
    const int pl = wib / SPLIT;
    const int64_t i = (int64_t)blockIdx.x * PP + pl;
    const bool valid = i < P.n;
    const int K = P.ent.K;
    const int half = A::CPLX ? P.k : 0;
    const int nvec = (A::CPLX ? P.k : K) / V;
    const int64_t n = P.n;
    const int eta = P.eta;
    const int loss = P.loss;

    // shared layout: sc[PP][eta] | red[PP][SPLIT][2] | acc[PP][SPLIT-1][2][K] | PIPE: per warp
    // bars[NS] (8 B each), then ring[NS][K]
    float* sc = smem + (size_t)pl * eta;
    float* red = smem + (size_t)PP * eta + (size_t)pl * SPLIT * 2;
    const size_t acc_off = (((size_t)PP * eta + PP * SPLIT * 2 + 3) & ~(size_t)3);
    float* accs = smem + acc_off + (size_t)pl * (SPLIT - 1) * 2 * K;

    const int NS = 1 << ns_log2;
    uint64_t* bars = nullptr;
    float* ring = nullptr;
    const uint32_t row_bytes = (uint32_t)K * 4u;
    if constexpr (PIPE) {
        const size_t pipe_off = acc_off + (size_t)PP * (SPLIT - 1) * 2 * K;  // multiple of 4 floats
        uint64_t* bars0 = reinterpret_cast<uint64_t*>(smem + pipe_off);
        const size_t rows_off = pipe_off + (size_t)4 * NS * 2;
        bars = bars0 + (size_t)wib * NS;
        ring = smem + rows_off + (size_t)wib * NS * K;
        if (lane == 0) {
            for (int b = 0; b < NS; ++b) mbar_init(bars + b, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }

    float* coef = gbuf_coef(P.gbuf, n, K);
    uint8_t* keep_out = gbuf_keep(P.gbuf, eta, n, K);

    // negatives of this warp: j = sub + SPLIT*m, m in [0,cnt)
    const int cnt = valid ? (eta - sub + SPLIT - 1) / SPLIT : 0;

    // ---- per-batch state of the candidate pipeline (one batch = up to 32 negatives of this warp)
    int lim = 0, my_idx = 0, my_keep = 0, mypos = 0, n_obj = 0;
    int64_t my_q = 0;
    unsigned m_obj = 0, m_sub = 0;
    bool my_issued = true;
    uint32_t seq_base = 0;  // ring sequence number of the batch's first position (PIPE)
    int consumed = 0;       // positions of the batch already consumed (PIPE)

    auto issue_window = [&]() {
        if constexpr (PIPE) {
            if (!my_issued && mypos < consumed + NS) {
                const uint32_t q = seq_base + (uint32_t)mypos;
                const uint32_t sl = q & (uint32_t)(NS - 1);
                mbar_expect_tx(bars + sl, row_bytes);
                bulk_g2s(ring + (size_t)sl * K, slot_row(P, my_idx, 2 * n + my_q), row_bytes, bars + sl);
                my_issued = true;
            }
        }
    };
    // ids of the batch starting at negative m0, processing order (object-side group first), first copies
    auto begin_batch = [&](int m0) {
        lim = min(32, cnt - m0);
        my_idx = 0;
        my_keep = 0;
        my_q = 0;
        if (lane < lim) {
            my_q = (int64_t)(sub + SPLIT * (m0 + lane)) * n + i;
            my_idx = P.repl[my_q];
            my_keep = P.keep[my_q];
        }
        m_obj = __ballot_sync(0xffffffffu, lane < lim && my_keep != 0);
        m_sub = __ballot_sync(0xffffffffu, lane < lim && my_keep == 0);
        if constexpr (PIPE) {
            n_obj = __popc(m_obj);
            const unsigned below = (1u << lane) - 1u;
            mypos = my_keep ? __popc(m_obj & below) : n_obj + __popc(m_sub & below);
            my_issued = !(lane < lim);
            consumed = 0;
            issue_window();
            // rows behind the ring window: the bulk copies of at most 2^ns_log2 rows are in flight per warp, which leaves a
            // random-row gather from HBM latency-bound (ncu r01: 62 % of the copy bandwidth on cfg5).  Their ids are known
            // now, so they are sent towards L2 right away (one prefetch per 128-byte line, no register or smem held);
            // the bulk copy that follows later finds them there.
            if (P.l2_prefetch && lane < lim && !my_issued && P.stage == nullptr) {
                const float* row = table_row(P.ent, my_idx);
                for (int off = 0; off < K; off += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
            }
        }
    };

    R Qo, Qs, AccO, AccS;
    row_zero(AccO);
    row_zero(AccS);
    float spos = 0.f;
    float msk[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) msk[c] = (lane + 32 * c < nvec) ? 1.f : 0.f;
    if (valid) {
        R s, p, o;
        const int32_t si = P.pos[3 * i + 0], pi = P.pos[3 * i + 1], oi = P.pos[3 * i + 2];
        row_load_clamped(s, slot_row(P, si, i), lane, nvec, half);
        row_load_clamped(p, P.rel + (size_t)pi * K, lane, nvec, half);
        row_load_clamped(o, slot_row(P, oi, n + i), lane, nvec, half);
        if constexpr (PIPE) begin_batch(0);  // the first candidate rows travel together with s, p, o
        A::queries(s, p, o, Qo, Qs);
        row_mask(Qo, lane, nvec);  // queries are zero past the end of the row: duplicate columns of
        row_mask(Qs, lane, nvec);  // the clamped candidate loads then contribute nothing
        spos = A::finish(warp_sum(A::partial(Qo, o, msk)), P.scale);
    } else {
        row_zero(Qo);
        row_zero(Qs);
    }
    // the loss sees nl(score) (models/EmbeddingModel.py:679-689, :801-812); the raw score stays around for the
    // backward (TransE L2 divides by the norm) and dL/draw = dL/dnl * nl'(raw)
    const int nl = P.nl;
    const float spos_raw = spos;
    const float dpos_nl = apply_nl_grad(nl, spos_raw);
    spos = apply_nl(nl, spos_raw);
    const float cpos = clip75(spos);
    const bool pos_in = (spos >= -75.f) && (spos <= 75.f);
    const float margin = P.margin;

    float loss_acc = 0.f;  // identical on all lanes
    float wsum = 0.f;      // pairwise: number of active hinges
    float zinv = 0.f;      // multiclass / self-adversarial: 1 / softmax denominator
    float amax = 0.f;      // self-adversarial: max_j alpha*s_j (softmax shift)
    float lbar = 0.f;      // self-adversarial: sum_j p_j * logsigmoid(-s_j - margin)
    const float alpha = P.alpha;
    // pass-2 weight dL/ds_j of the two-pass losses
    auto weight2 = [&](float sn_raw) -> float {
        const float sn = apply_nl(nl, sn_raw);
        const float dn = apply_nl_grad(nl, sn_raw);
        if (loss == KGE_LOSS_MULTICLASS_NLL) {
            // losses/nll_multiclass.py:70-81 : softmax weight, zero outside the clip range
            const bool in = (sn >= -75.f) && (sn <= 75.f);
            return in ? expf(sn) * zinv * dn : 0.f;
        }
        // losses/self_adversarial.py:97-110 : L = -sum_j p_j*l_j, p = softmax(alpha*s), l_j = logsigmoid(-s_j - margin);
        // the gradient flows through p as well: dL/ds_j = p_j*(sigmoid(s_j + margin) - alpha*(l_j - sum_i p_i*l_i))
        const float pj = expf(alpha * sn - amax) * zinv;
        return pj * (sigmoidf(sn + margin) - alpha * (log_sigmoid(-sn - margin) - lbar)) * dn;
    };

    // MODE 0: single pass (pairwise / nll).  MODE 1: scores only (multiclass pass 1).
    // MODE 2: backward with the scores in sc[] (multiclass pass 2).
    // first_ready: the first batch was already begun (PIPE prologue)
    auto sweep = [&](auto mode_tag, bool first_ready) {
        constexpr int MODE = decltype(mode_tag)::value;
        for (int m0 = 0; m0 < cnt; m0 += 32) {
            if (!(first_ready && m0 == 0)) begin_batch(m0);
            float my_c = 0.f, my_sn = 0.f;
            // ---- PIPE: groups of G resident rows, transposed reduction, loss once per group
            [[maybe_unused]] int inv = 0;       // MODE 2: lane x holds the batch lane that owns position x
            [[maybe_unused]] float w_own = 0.f;  // MODE 2: weight of this lane's own negative
            if constexpr (PIPE && MODE == 2) {
                const int x = lane;
                inv = x < n_obj ? (int)__fns(m_obj, 0, x + 1) : (int)__fns(m_sub, 0, x - n_obj + 1);
                if (lane < lim) {
                    const float sn = sc[sub + SPLIT * (m0 + lane)];
                    w_own = weight2(sn);
                    my_c = A::coefficient(w_own, sn, P.scale);
                    my_sn = sn;
                }
            }
            auto group_pipe = [&](auto obj_tag, int x_begin, int x_end) {
                constexpr bool OBJ = decltype(obj_tag)::value;
                const R& Q = OBJ ? Qo : Qs;
                R& Acc = OBJ ? AccO : AccS;
                const int cidx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);  // candidate this lane holds after the reduction
                for (int x0 = x_begin; x0 < x_end; x0 += G) {
                    const int g = min(G, x_end - x0);
                    float tot = 0.f;
                    if constexpr (MODE != 2) {
                        float pv[G];
#pragma unroll
                        for (int c = 0; c < G; ++c) {
                            pv[c] = 0.f;
                            if (c < g) {
                                const uint32_t q = seq_base + (uint32_t)(x0 + c);
                                const uint32_t sl = q & (uint32_t)(NS - 1);
                                mbar_wait(bars + sl, (q >> ns_log2) & 1u);
                                R r;
                                row_load_smem(r, ring + (size_t)sl * K, lane, nvec, half);
                                pv[c] = A::partial(Q, r, msk);
                            }
                        }
                        // transposed reduction: 2 + 1 exchange steps, then a butterfly over the low 3 bits
                        const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
                        float k0 = hi16 ? pv[2] : pv[0], k1 = hi16 ? pv[3] : pv[1];
                        k0 += __shfl_xor_sync(0xffffffffu, hi16 ? pv[0] : pv[2], 16);
                        k1 += __shfl_xor_sync(0xffffffffu, hi16 ? pv[1] : pv[3], 16);
                        tot = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                    }
                    const bool has = cidx < g;
                    float sn = 0.f, w = 0.f;  // sn: RAW score of the candidate this lane holds
                    if constexpr (MODE != 2) sn = A::finish(tot, P.scale);
                    if constexpr (MODE == 0) {
                        float term;
                        const float st = apply_nl(nl, sn);
                        if (loss == KGE_LOSS_PAIRWISE || loss == KGE_LOSS_ABSOLUTE_MARGIN) {
                            // losses/pairwise.py:69, absolute_margin.py:69 ; tf.maximum passes the gradient when t >= 0
                            const float tt = loss == KGE_LOSS_PAIRWISE ? margin - spos + st : margin + st;
                            term = fmaxf(tt, 0.f);
                            w = (tt >= 0.f) ? 1.f : 0.f;
                        } else {
                            // losses/nll.py:55-59 : log(1+exp(clip(neg)))
                            const float e = expf(clip75(st));
                            term = logf(1.f + e);
                            const bool in = (st >= -75.f) && (st <= 75.f);
                            w = in ? e / (1.f + e) : 0.f;
                        }
                        if (!has) {
                            w = 0.f;
                            term = 0.f;
                        }
                        if ((lane & 7) == 0) {  // one lane per candidate feeds the (lane-partial) loss sums
                            loss_acc += term;
                            if (loss == KGE_LOSS_PAIRWISE) wsum += w;
                        }
                        w *= apply_nl_grad(nl, sn);  // dL/d(raw score)
                    }
                    if constexpr (MODE != 2) {
                        // hand score / coefficient to the lane that owns the negative (it stores them)
                        const int rel = mypos - x0;
                        const bool mine_in = lane < lim && rel >= 0 && rel < g;
                        const int holder = (((rel >> 1) & 1) * 16 + (rel & 1) * 8) & 31;
                        const float cv = __shfl_sync(0xffffffffu, A::coefficient(w, sn, P.scale), holder);
                        const float sv = __shfl_sync(0xffffffffu, sn, holder);
                        if (mine_in) {
                            my_c = cv;
                            my_sn = sv;
                        }
                    }
                    if constexpr (MODE != 1) {
#pragma unroll
                        for (int c = 0; c < G; ++c) {
                            if (c < g) {
                                float wc, sc_c = 0.f;
                                if constexpr (MODE == 2) {
                                    const int own = __shfl_sync(0xffffffffu, inv, (x0 + c) & 31);
                                    wc = __shfl_sync(0xffffffffu, w_own, own);
                                    if constexpr (MODEL == 1) sc_c = __shfl_sync(0xffffffffu, my_sn, own);
                                } else {
                                    const int holder = (c >> 1) * 16 + (c & 1) * 8;
                                    wc = __shfl_sync(0xffffffffu, w, holder);
                                    if constexpr (MODEL == 1) sc_c = __shfl_sync(0xffffffffu, sn, holder);
                                }
                                const uint32_t q = seq_base + (uint32_t)(x0 + c);
                                const uint32_t sl = q & (uint32_t)(NS - 1);
                                if constexpr (MODE == 2) mbar_wait(bars + sl, (q >> ns_log2) & 1u);
                                R r;
                                row_load_smem(r, ring + (size_t)sl * K, lane, nvec, half);
                                A::accumulate(Q, r, OBJ, wc, sc_c, P.scale, Acc);
                            }
                        }
                    }
                    // the group's rows are no longer needed: their ring slots take the next positions
                    __syncwarp();
                    consumed += g;
                    issue_window();
                }
            };
            // ---- register path: the negatives of this round, grouped by corrupted side so that the query
            // (Qo / Qs) and the accumulator (AccO / AccS) are compile-time choices inside each group
            auto group = [&](auto obj_tag, unsigned todo) {
                constexpr bool OBJ = decltype(obj_tag)::value;
                const R& Q = OBJ ? Qo : Qs;
                R& Acc = OBJ ? AccO : AccS;
                while (todo) {
                    R r[U];
                    int src[U];
                    float part[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        src[u] = todo ? __ffs(todo) - 1 : -1;
                        todo &= todo - 1;  // 0 stays 0
                        const int idx = __shfl_sync(0xffffffffu, my_idx, max(src[u], 0));
                        if (src[u] >= 0)
                            row_load_clamped(r[u], slot_row(P, idx, 2 * n + (int64_t)(sub + SPLIT * (m0 + src[u])) * n + i), lane, nvec, half);
                    }
                    if constexpr (MODE != 2) {
#pragma unroll
                        for (int u = 0; u < U; ++u) part[u] = (src[u] >= 0) ? A::partial(Q, r[u], msk) : 0.f;
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1)
#pragma unroll
                            for (int u = 0; u < U; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o2);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (src[u] >= 0) {
                            const int j = sub + SPLIT * (m0 + src[u]);
                            float sn, w;
                            if constexpr (MODE == 2) sn = sc[j];
                            else sn = A::finish(part[u], P.scale);
                            if constexpr (MODE == 1) {
                                if (lane == 0) sc[j] = sn;
                            } else {
                                if constexpr (MODE == 2) {
                                    w = weight2(sn);
                                } else if (loss == KGE_LOSS_PAIRWISE || loss == KGE_LOSS_ABSOLUTE_MARGIN) {
                                    // losses/pairwise.py:69, absolute_margin.py:69 ; tf.maximum passes the gradient when t >= 0
                                    const float st = apply_nl(nl, sn);
                                    const float tt = loss == KGE_LOSS_PAIRWISE ? margin - spos + st : margin + st;
                                    loss_acc += fmaxf(tt, 0.f);
                                    w = (tt >= 0.f) ? 1.f : 0.f;
                                    wsum += w;
                                    w *= apply_nl_grad(nl, sn);
                                } else {
                                    // losses/nll.py:55-59 : log(1+exp(clip(neg)))
                                    const float st = apply_nl(nl, sn);
                                    const float e = expf(clip75(st));
                                    loss_acc += logf(1.f + e);
                                    const bool in = (st >= -75.f) && (st <= 75.f);
                                    w = in ? e / (1.f + e) * apply_nl_grad(nl, sn) : 0.f;
                                }
                                const float c = A::coefficient(w, sn, P.scale);
                                if (lane == src[u]) {
                                    my_c = c;
                                    my_sn = sn;
                                }
                                A::accumulate(Q, r[u], OBJ, w, sn, P.scale, Acc);
                            }
                        }
                    }
                }
            };
            if constexpr (PIPE) {
                group_pipe(std::true_type{}, 0, n_obj);
                group_pipe(std::false_type{}, n_obj, lim);
                seq_base += (uint32_t)lim;
                if constexpr (MODE == 1) {
                    if (lane < lim) sc[sub + SPLIT * (m0 + lane)] = my_sn;
                }
            } else {
                group(std::true_type{}, m_obj);
                group(std::false_type{}, m_sub);
            }
            if constexpr (MODE != 1) {
                if (lane < lim) {
                    coef[my_q] = my_c;
                    keep_out[my_q] = (uint8_t)my_keep;
                    if (P.dbg_scores != nullptr) P.dbg_scores[n + my_q] = apply_nl(nl, my_sn);
                }
            }
        }
    };

    if (loss_two_pass(loss)) {
        sweep(std::integral_constant<int, 1>{}, PIPE);
        __syncthreads();
        if (loss == KGE_LOSS_MULTICLASS_NLL) {
            float zpart = 0.f;
            if (valid)
                for (int j = lane; j < eta; j += 32) zpart += expf(clip75(apply_nl(nl, sc[j])));
            const float pe = expf(cpos);
            const float z = warp_sum(zpart) + pe;
            zinv = 1.f / z;
            loss_acc = -logf(pe / z);
        } else {
            // softmax over the eta negatives of this positive (tf.nn.softmax subtracts the max)
            float mx = -INFINITY;
            if (valid)
                for (int j = lane; j < eta; j += 32) mx = fmaxf(mx, alpha * apply_nl(nl, sc[j]));
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
            amax = valid ? mx : 0.f;
            float zpart = 0.f, lpart = 0.f;
            if (valid)
                for (int j = lane; j < eta; j += 32) {
                    const float st = apply_nl(nl, sc[j]);
                    const float e = expf(alpha * st - amax);
                    zpart += e;
                    lpart += e * log_sigmoid(-st - margin);
                }
            const float z = warp_sum(zpart);
            zinv = valid ? 1.f / z : 0.f;
            lbar = warp_sum(lpart) * zinv;
            loss_acc = -log_sigmoid(margin + spos) - lbar;
        }
        sweep(std::integral_constant<int, 2>{}, false);
    } else {
        sweep(std::integral_constant<int, 0>{}, PIPE);
        if constexpr (PIPE) {  // the staged path keeps lane-partial loss sums
            loss_acc = warp_sum(loss_acc);
            wsum = warp_sum(wsum);
        }
    }

    if constexpr (SPLIT > 1) {
        if (sub > 0 && valid) {
            float* a = accs + (size_t)(sub - 1) * 2 * K;
            row_store(a, AccO, lane, nvec, half);
            row_store(a + K, AccS, lane, nvec, half);
            if (lane == 0) {
                red[sub * 2 + 0] = loss_acc;
                red[sub * 2 + 1] = wsum;
            }
        }
        __syncthreads();
        if (sub == 0 && valid) {
#pragma unroll
            for (int s2 = 1; s2 < SPLIT; ++s2) {
                R t0, t1;
                const float* a = accs + (size_t)(s2 - 1) * 2 * K;
                row_load(t0, a, lane, nvec, half);
                row_load(t1, a + K, lane, nvec, half);
                row_add(AccO, t0);
                row_add(AccS, t1);
                if (!loss_two_pass(loss)) loss_acc += red[s2 * 2 + 0];
                wsum += red[s2 * 2 + 1];
            }
        }
    }
    if (sub != 0 || !valid) return;

    float wpos;
    if (loss == KGE_LOSS_PAIRWISE) {
        wpos = -wsum;
    } else if (loss == KGE_LOSS_NLL) {
        // positives are tiled eta times (models/EmbeddingModel.py:724-729)
        const float e = expf(-cpos);
        loss_acc += (float)eta * logf(1.f + e);
        wpos = pos_in ? -(float)eta * (e / (1.f + e)) : 0.f;
    } else if (loss == KGE_LOSS_ABSOLUTE_MARGIN) {
        // positives are tiled eta times: sum(max(margin + neg, 0) - pos_tiled)
        loss_acc -= (float)eta * spos;
        wpos = -(float)eta;
    } else if (loss == KGE_LOSS_SELF_ADVERSARIAL) {
        wpos = -sigmoidf(-(margin + spos));
    } else {
        wpos = pos_in ? -(1.f - expf(cpos) * zinv) : 0.f;
    }
    wpos *= dpos_nl;  // dL/d(raw positive score)

    float* GB = P.gbuf;
    row_store(GB + (size_t)(3 * n + i) * K, Qo, lane, nvec, half);
    row_store(GB + (size_t)(4 * n + i) * K, Qs, lane, nvec, half);
    {
        // the positive itself: an object-side candidate with r = o and weight dL/dpos
        R s, p, o, gs, gp, go;
        const int32_t si = P.pos[3 * i + 0], pi = P.pos[3 * i + 1], oi = P.pos[3 * i + 2];
        row_load_clamped(o, slot_row(P, oi, n + i), lane, nvec, half);
        A::backward_pos(Qo, o, wpos, spos_raw, P.scale, go, AccO);
        row_load_clamped(s, slot_row(P, si, i), lane, nvec, half);
        row_load_clamped(p, P.rel + (size_t)pi * K, lane, nvec, half);
        A::fold(s, p, o, AccO, AccS, gs, gp, go);
        row_store(GB + (size_t)i * K, gs, lane, nvec, half);
        row_store(GB + (size_t)(n + i) * K, go, lane, nvec, half);
        row_store(GB + (size_t)(2 * n + i) * K, gp, lane, nvec, half);
    }
    if (lane == 0) {
        P.loss_part[i] = loss_acc;
        if (P.dbg_scores != nullptr) P.dbg_scores[i] = spos;
    }
}

// resident CTAs of the staged kernel per SM.  3 leaves room on every SM for the CTAs of the radix sort that runs beside it --
// right when the sort is the longer of the two (cfg1-3: ~60 us against 30-53 us); a large batch (cfg5: sort 120 us against
// 210 us) is better served by a fourth CTA (measured: fwd_bwd 0.211 -> 0.191 ms).  KGE_FWD_MAXCTAS overrides (A/B).
static inline int fwd_bwd_max_ctas(int64_t n_slots) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_FWD_MAXCTAS");
        v = (e != nullptr && e[0] >= '1' && e[0] <= '8') ? (e[0] - '0') : 0;
    }
    return v > 0 ? v : (n_slots >= 400000 ? 4 : 3);
}

// ring slots per warp (log2) of the PIPE variant: 16 when four 4-warp CTAs still fit one SM, else 8
// (two groups of KGE_FB_G rows: one being processed, one in flight)
static inline int fwd_bwd_ns_log2(int K) {
    const size_t row = (size_t)K * 4;
    return (16 * row * 4 <= 54 * 1024) ? 4 : 3;
}

static inline size_t fwd_bwd_smem(int split, int eta, int K, bool pipe, int ns_log2) {
    const int pp = 4 / split;
    size_t fl = (((size_t)pp * eta + pp * split * 2 + 3) & ~(size_t)3) + (size_t)pp * (split - 1) * 2 * K;
    if (pipe) {
        const int ns = 1 << ns_log2;
        fl += (size_t)4 * ns * 2 + (size_t)4 * ns * K;
    }
    return fl * sizeof(float);
}

template <int MODEL, int V, int NCH, int U>
static int launch_fwd_bwd_split(int split, bool pipe, const FwdBwdParams& P, cudaStream_t st) {
    const int K = P.ent.K;
    const int ns_log2 = fwd_bwd_ns_log2(K);
    auto go = [&](auto kern, int sp, bool pp_) -> int {
        const int pp = 4 / sp;
        size_t smem = fwd_bwd_smem(sp, P.eta, K, pp_, ns_log2);
        if (pp_) {
            // leave room on every SM for the CTAs of the radix sort that runs beside this kernel on the
            // side stream: at most `m` resident CTAs of this kernel (KGE_FWD_MAXCTAS, default 3)
            const int m = fwd_bwd_max_ctas((int64_t)(3 + P.eta) * P.n);
            const size_t floor_bytes = 233472 / (size_t)(m + 1) - 1024 + 16;
            if (smem < floor_bytes) smem = floor_bytes;
        }
        KGE_REQUIRE(smem <= 220 * 1024, "kge_train: embedding size %d needs %zu bytes of shared memory", K, smem);
        if (smem > 48 * 1024) KGE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)((P.n + pp - 1) / pp)), block(128);
        kern<<<grid, block, smem, st>>>(P, ns_log2);
        KGE_CUDA_CHECK(cudaGetLastError());
        return 0;
    };
    if constexpr (V == 4 && NCH <= 4) {
        if (pipe) {
            if (split >= 4) return go(kge_fwd_bwd_kernel<MODEL, V, NCH, U, 4, true>, 4, true);
            if (split == 2) return go(kge_fwd_bwd_kernel<MODEL, V, NCH, U, 2, true>, 2, true);
            return go(kge_fwd_bwd_kernel<MODEL, V, NCH, U, 1, true>, 1, true);
        }
        if (split >= 4) return go(kge_fwd_bwd_kernel<MODEL, V, NCH, U, 4, false>, 4, false);
        if (split == 2) return go(kge_fwd_bwd_kernel<MODEL, V, NCH, U, 2, false>, 2, false);
    }
    return go(kge_fwd_bwd_kernel<MODEL, V, NCH, U, 1, false>, 1, false);
}

template <int MODEL, int V>
static int launch_fwd_bwd_nch(int nch, int split, bool pipe, const FwdBwdParams& P, cudaStream_t st) {
    constexpr bool C = (MODEL == 3);
    // U candidate rows in flight per warp: ~64 registers of row data
    switch (nch) {
        case 1: return launch_fwd_bwd_split<MODEL, V, 1, (C ? 8 : 8)>(split, pipe, P, st);
        case 2: return launch_fwd_bwd_split<MODEL, V, 2, (C ? 2 : 4)>(split, pipe, P, st);
        case 3:
        case 4: return launch_fwd_bwd_split<MODEL, V, 4, (C ? 2 : 4)>(split, pipe, P, st);
        case 5:
        case 6:
        case 7:
        case 8: return launch_fwd_bwd_split<MODEL, V, 8, (C ? 1 : 2)>(split, false, P, st);
        default: kge_set_error("kge_train: embedding size too large for the fused kernel (chunks/lane=%d)", nch); return -1;
    }
}

// KGE_FWD_PIPE=0 forces the register path (A/B measurements)
static inline bool fwd_bwd_pipe_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_FWD_PIPE");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <int MODEL>
static int launch_fwd_bwd_model(const FwdBwdParams& P, int sm_count, cudaStream_t st) {
    const bool cplx = (MODEL == 3);
    const int width = cplx ? P.k : P.ent.K;  // floats per half / per row
    // cooperate SPLIT warps per positive when the batch alone cannot fill the machine
    int split = 1;
    const int64_t want = (int64_t)sm_count * 24;
    if (P.n < want && P.eta >= 8) split = 2;
    if (P.n * 2 < want && P.eta >= 16) split = 4;
    {
        static int force = -2;  // KGE_FWD_SPLIT=1|2|4 overrides (A/B)
        if (force == -2) {
            const char* e = getenv("KGE_FWD_SPLIT");
            force = e == nullptr ? -1 : atoi(e);
        }
        if (force == 1 || force == 2 || force == 4) split = force;
    }
    if (width % 4 == 0) {
        int nvec = width / 4;
        // staged (bulk-copy) rows need 16-byte row pitch and local memory
        const bool pipe = fwd_bwd_pipe_enabled() && (P.ent.n_shards == 1 || P.stage != nullptr) && (P.ent.K % 4 == 0);
        return launch_fwd_bwd_nch<MODEL, 4>((nvec + 31) / 32, split, pipe, P, st);
    }
    return launch_fwd_bwd_nch<MODEL, 1>((width + 31) / 32, 1, false, P, st);
}

// one translation unit per model (parallel compilation): kge_train_fwd_m{0,1,2,3}.cu
int kge_launch_fwd_bwd_m0(const FwdBwdParams& P, int sm_count, cudaStream_t st);
int kge_launch_fwd_bwd_m1(const FwdBwdParams& P, int sm_count, cudaStream_t st);
int kge_launch_fwd_bwd_m2(const FwdBwdParams& P, int sm_count, cudaStream_t st);
int kge_launch_fwd_bwd_m3(const FwdBwdParams& P, int sm_count, cudaStream_t st);
