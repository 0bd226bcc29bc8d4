// Filtered ranking for sm_100a: device filter index, query folding, all-entity sweep with the
// x1e5 int quantisation, in-kernel filter and fused rank counting.
//
// Replaces reference evaluation/protocol.py:726-979 (evaluate_performance), :448-528
// (generate_corruptions_for_eval), models/EmbeddingModel.py:1845-2033 (eval graph,
// perform_comparision), datasets/sqlite_adapter.py:449-508 (filter queries).
//
// The [T, 2E] score matrix is never materialised: every (query tile x entity tile) is scored in
// registers, quantised, compared with the positive's quantised score and reduced to four counters
// per (test triple, side): gt, eq over all candidates and gt, eq over filtered candidates.  The
// candidate that IS the test triple (e == o on the object sweep, e == s on the subject sweep) is
// skipped by the sweep and added analytically in kge_rank_finalize (its score equals the
// positive's by construction in the reference, models/EmbeddingModel.py:1861-1866).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "kge_common.cuh"

int kge_rank_sweep_tc(kge_ctx* ctx, int model, int K, const float* q, int64_t NQ, int64_t T, const float* ent_local,
                      int64_t row_begin, int64_t row_end, const int32_t* test, const int32_t* pos_q,
                      const int32_t* excl_lo, const int32_t* excl_hi, const int32_t* sp_ent, const int32_t* po_ent,
                      int side_mask, int32_t* counts, const uint32_t* q_absmax, cudaStream_t st);
bool kge_rank_tc_f16();

// ------------------------------------------------------------------------------------------------
// filter index
// ------------------------------------------------------------------------------------------------
__global__ void kge_filter_comp_kernel(const int32_t* __restrict__ tri, int64_t F, int64_t E, int64_t R,
                                       uint64_t* __restrict__ sp, uint64_t* __restrict__ po) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < F; t += (int64_t)gridDim.x * blockDim.x) {
        uint64_t s = (uint64_t)tri[3 * t], p = (uint64_t)tri[3 * t + 1], o = (uint64_t)tri[3 * t + 2];
        sp[t] = (s * (uint64_t)R + p) * (uint64_t)E + o;
        po[t] = (o * (uint64_t)R + p) * (uint64_t)E + s;
    }
}

__global__ void kge_filter_ent_kernel(const uint64_t* __restrict__ comp, const int32_t* __restrict__ cnt, int64_t E,
                                      int32_t* __restrict__ ent) {
    int64_t n = *cnt;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        ent[t] = (int32_t)(comp[t] % (uint64_t)E);
}

extern "C" int kge_filter_clear(kge_ctx* ctx) {
    KGE_REQUIRE(ctx != nullptr, "kge_filter_clear: null ctx");
    ctx->f_valid = false;
    return 0;
}

extern "C" int kge_filter_build(kge_ctx* ctx, const int32_t* triples, int64_t F, int64_t E, int64_t R, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_filter_build: null ctx");
    KGE_REQUIRE(F >= 0 && E > 0 && R > 0, "kge_filter_build: bad sizes");
    KGE_REQUIRE(F < (int64_t)INT32_MAX, "kge_filter_build: too many filter triples");
    KGE_REQUIRE((double)E * (double)R * (double)E < 9.0e18, "kge_filter_build: E*R*E overflows the 64-bit composite key");
    cudaStream_t st = (cudaStream_t)stream;
    ctx->f_E = E;
    ctx->f_R = R;
    if (ctx->f_count.reserve(4 * sizeof(int32_t))) return -2;
    KGE_CUDA_CHECK(cudaMemsetAsync(ctx->f_count.p, 0, 4 * sizeof(int32_t), st));
    ctx->f_valid = true;
    ctx->f_n_sp = ctx->f_n_po = F;
    if (F == 0) return 0;
    KGE_REQUIRE(triples != nullptr, "kge_filter_build: null triples");
    size_t b8 = (size_t)F * sizeof(uint64_t);
    if (ctx->f_sp_comp.reserve(b8) || ctx->f_po_comp.reserve(b8) || ctx->f_tmp.reserve(b8) || ctx->f_tmp2.reserve(b8)) return -2;
    if (ctx->f_sp_ent.reserve((size_t)F * 4) || ctx->f_po_ent.reserve((size_t)F * 4)) return -2;
    int threads = 256;
    int blocks = (int)std::min<int64_t>((F + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
    kge_filter_comp_kernel<<<blocks, threads, 0, st>>>(triples, F, E, R, ctx->f_tmp.as<uint64_t>(), ctx->f_tmp2.as<uint64_t>());
    KGE_CUDA_CHECK(cudaGetLastError());
    int end_bit = 1;
    {
        double top = (double)E * (double)R * (double)E;
        while (end_bit < 64 && ldexp(1.0, end_bit) < top) ++end_bit;
    }
    int32_t* cnt = ctx->f_count.as<int32_t>();
    for (int which = 0; which < 2; ++which) {
        uint64_t* raw = which == 0 ? ctx->f_tmp.as<uint64_t>() : ctx->f_tmp2.as<uint64_t>();
        uint64_t* dst = which == 0 ? ctx->f_sp_comp.as<uint64_t>() : ctx->f_po_comp.as<uint64_t>();
        int32_t* ent = which == 0 ? ctx->f_sp_ent.as<int32_t>() : ctx->f_po_ent.as<int32_t>();
        // sort into dst, then unique back into raw, then copy raw->dst
        size_t tb = 0, tb2 = 0;
        KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tb, raw, dst, (int)F, 0, end_bit, st));
        KGE_CUDA_CHECK(cub::DeviceSelect::Unique(nullptr, tb2, dst, raw, cnt + which, (int)F, st));
        if (ctx->sort_tmp.reserve(std::max(tb, tb2))) return -2;
        KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tb, raw, dst, (int)F, 0, end_bit, st));
        KGE_CUDA_CHECK(cub::DeviceSelect::Unique(ctx->sort_tmp.p, tb2, dst, raw, cnt + which, (int)F, st));
        KGE_CUDA_CHECK(cudaMemcpyAsync(dst, raw, b8, cudaMemcpyDeviceToDevice, st));
        kge_filter_ent_kernel<<<blocks, threads, 0, st>>>(dst, cnt + which, E, ent);
        KGE_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

extern "C" int64_t kge_filter_size_sync(kge_ctx* ctx) {
    if (!ctx || !ctx->f_valid || !ctx->f_count.p) return 0;
    int32_t c[2] = {0, 0};
    if (cudaMemcpy(c, ctx->f_count.p, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return c[0];
}

// ------------------------------------------------------------------------------------------------
// per-test-triple preparation: positive score (normal fp32 _fn), folded queries, exclusion ranges
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t lower_bound_u64(const uint64_t* __restrict__ a, int32_t n, uint64_t key) {
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

struct PrepParams {
    int model, k;
    TableView ent;
    const float* rel;
    const int32_t* test;
    const float* s_rows;  // optional [T,K]: subject / object rows of the test triples, supplied by the caller (the
    const float* o_rows;  // table is sharded and the owners contributed them); NULL: read through `ent`
    int64_t T;
    int filtered;
    int nl;  // KGE_NL_*
    const uint64_t* sp_comp;
    const uint64_t* po_comp;
    const int32_t* f_cnt;
    int64_t E, R;
    float* q;         // [2T,K]: rows [0,T) object-sweep queries, [T,2T) subject-sweep queries
    int32_t* pos_q;   // [T] quantised positive score
    int32_t* excl_lo; // [2T]
    int32_t* excl_hi; // [2T]
    uint32_t* q_absmax;  // optional: bit pattern of the largest |q| over the swept sides (the fp16 split's operand scale)
    int amax_mask;       // bit 0: object-sweep queries count, bit 1: subject-sweep queries
};

__global__ void kge_rank_prepare_kernel(PrepParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= P.T) return;
    const int K = P.ent.K, k = P.k, model = P.model;
    const int32_t si = P.test[3 * t], pi = P.test[3 * t + 1], oi = P.test[3 * t + 2];
    const float* s = P.s_rows != nullptr ? P.s_rows + (size_t)t * K : table_row(P.ent, si);
    const float* p = P.rel + (size_t)pi * K;
    const float* o = P.o_rows != nullptr ? P.o_rows + (size_t)t * K : table_row(P.ent, oi);
    float* qo = P.q + (size_t)t * K;
    float* qs = P.q + (size_t)(P.T + t) * K;
    float acc = 0.f;
    float mo = 0.f, ms = 0.f;  // largest |qo|, |qs| of this lane
    const float hs = model == KGE_HOLE ? 2.0f / (float)k : 1.0f;
    // the loads of four column steps are issued before their stores (the rows are read-only here); the sum keeps its order
    if (model == KGE_TRANSE_L1 || model == KGE_TRANSE_L2) {
#pragma unroll 4
        for (int c = lane; c < K; c += 32) {
            float sv = __ldg(s + c), pv = __ldg(p + c), ov = __ldg(o + c);
            qo[c] = sv + pv;   // S_o[e] = -|| (s+p) - e ||
            qs[c] = ov - pv;   // S_s[e] = -|| e - (o-p) ||
            float u = sv + pv - ov;
            acc = model == KGE_TRANSE_L1 ? acc + fabsf(u) : fmaf(u, u, acc);
        }
        acc = warp_sum(acc);
        acc = model == KGE_TRANSE_L1 ? -acc : -sqrtf(acc);
    } else if (model == KGE_DISTMULT) {
#pragma unroll 4
        for (int c = lane; c < K; c += 32) {
            float sv = __ldg(s + c), pv = __ldg(p + c), ov = __ldg(o + c);
            const float a = sv * pv, b = pv * ov;
            qo[c] = a;
            qs[c] = b;
            mo = fmaxf(mo, fabsf(a));
            ms = fmaxf(ms, fabsf(b));
            acc = fmaf(sv * pv, ov, acc);
        }
        acc = warp_sum(acc);
    } else {
#pragma unroll 2
        for (int c = lane; c < k; c += 32) {
            float sr = __ldg(s + c), si2 = __ldg(s + c + k), pr = __ldg(p + c), pim = __ldg(p + c + k), orr = __ldg(o + c),
                  oim = __ldg(o + c + k);
            float a = pr * sr - pim * si2, b = pr * si2 + pim * sr;
            const float q0 = hs * a, q1 = hs * b, q2 = hs * (pr * orr + pim * oim), q3 = hs * (pr * oim - pim * orr);
            qo[c] = q0;
            qo[c + k] = q1;
            qs[c] = q2;
            qs[c + k] = q3;
            mo = fmaxf(mo, fmaxf(fabsf(q0), fabsf(q1)));
            ms = fmaxf(ms, fmaxf(fabsf(q2), fabsf(q3)));
            acc = fmaf(a, orr, acc);
            acc = fmaf(b, oim, acc);
        }
        acc = hs * warp_sum(acc);
    }
    if (P.q_absmax != nullptr) {
        float m = fmaxf((P.amax_mask & 1) ? mo : 0.f, (P.amax_mask & 2) ? ms : 0.f);
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o2));
        // the running maximum settles after a few warps: most warps see that theirs is not larger and skip the atomic
        // (a stale read only costs an atomic that changes nothing)
        if (lane == 0 && m > 0.f && __float_as_uint(m) > *(volatile uint32_t*)P.q_absmax) atomicMax(P.q_absmax, __float_as_uint(m));
    }
    if (lane == 0) P.pos_q[t] = quantise_score(apply_nl(P.nl, acc));
    if (lane < 2) {
        int32_t lo = 0, hi = 0;
        if (P.filtered) {
            const uint64_t* comp = lane == 0 ? P.sp_comp : P.po_comp;
            int32_t n = P.f_cnt[lane];
            uint64_t a = (uint64_t)(lane == 0 ? si : oi);
            uint64_t prefix = a * (uint64_t)P.R + (uint64_t)pi;
            lo = lower_bound_u64(comp, n, prefix * (uint64_t)P.E);
            hi = lower_bound_u64(comp, n, (prefix + 1) * (uint64_t)P.E);
        }
        P.excl_lo[(int64_t)lane * P.T + t] = lo;
        P.excl_hi[(int64_t)lane * P.T + t] = hi;
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core sweep (all models; the only path for TransE, whose L1/L2 distance has no MMA form)
// ------------------------------------------------------------------------------------------------
struct SweepParams {
    int model;
    int K;
    const float* q;          // [NQ,K]
    int64_t NQ, T;
    const float* ent_local;  // [row_end-row_begin, K]
    int64_t row_begin, row_end;
    const int32_t* test;
    const int32_t* pos_q;
    const int32_t* excl_lo;
    const int32_t* excl_hi;
    const int32_t* sp_ent;
    const int32_t* po_ent;
    int64_t chunk;           // entities per blockIdx.y
    int64_t q_row0;          // first query row handled (side selection)
    int64_t q_rows;          // number of query rows handled
    int32_t* counts;         // [T,2,4]
    int nl;                  // KGE_NL_*
};

#define SW_BM 64
#define SW_BN 64
#define SW_BK 16

template <int MODE>  // 0 dot, 1 L1, 2 L2
__global__ void __launch_bounds__(256) kge_rank_sweep_kernel(SweepParams P) {
    __shared__ float Qs[SW_BK][SW_BM + 4];
    __shared__ float Es[SW_BK][SW_BN + 4];
    __shared__ unsigned long long s_mask[SW_BM];
    __shared__ int32_t s_cur[SW_BM], s_hi[SW_BM], s_self[SW_BM], s_posq[SW_BM];
    __shared__ const int32_t* s_list[SW_BM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = P.q_row0 + (int64_t)blockIdx.x * SW_BM;
    const int64_t m_end = P.q_row0 + P.q_rows;
    const int64_t e0 = P.row_begin + (int64_t)blockIdx.y * P.chunk;
    const int64_t e1 = min(P.row_end, e0 + P.chunk);
    if (e0 >= e1) return;
    const int K = P.K;

    if (tid < SW_BM) {
        int64_t r = m0 + tid;
        int32_t cur = 0, hi = 0, self = -1, pq = 0;
        const int32_t* list = nullptr;
        if (r < m_end) {
            int side = r >= P.T ? 1 : 0;
            int64_t t = r - (int64_t)side * P.T;
            self = side == 0 ? P.test[3 * t + 2] : P.test[3 * t + 0];
            pq = P.pos_q[t];
            list = side == 0 ? P.sp_ent : P.po_ent;
            int32_t lo = P.excl_lo[r];
            hi = P.excl_hi[r];
            // first list entry >= e0
            int32_t a = lo, b = hi;
            while (a < b) {
                int32_t mid = (a + b) >> 1;
                if ((int64_t)list[mid] < e0) a = mid + 1;
                else b = mid;
            }
            cur = a;
        }
        s_cur[tid] = cur;
        s_hi[tid] = hi;
        s_self[tid] = self;
        s_posq[tid] = pq;
        s_list[tid] = list;
    }
    int cnt[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) cnt[i][c] = 0;

    // loader mapping: 256 threads x float4 = 64 rows x 16 floats
    const int lr = tid >> 2, lc = (tid & 3) * 4;

    for (int64_t n0 = e0; n0 < e1; n0 += SW_BN) {
        __syncthreads();
        if (tid < SW_BM) {
            unsigned long long mk = 0ull;
            int32_t cur = s_cur[tid], hi = s_hi[tid];
            const int32_t* list = s_list[tid];
            while (cur < hi) {
                int64_t e = list[cur];
                if (e >= n0 + SW_BN) break;
                mk |= 1ull << (int)(e - n0);
                ++cur;
            }
            s_cur[tid] = cur;
            s_mask[tid] = mk;
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < K; k0 += SW_BK) {
            float qv[4] = {0.f, 0.f, 0.f, 0.f}, ev[4] = {0.f, 0.f, 0.f, 0.f};
            {
                int64_t r = m0 + lr;
                if (r < m_end) {
                    const float* src = P.q + (size_t)r * K + k0 + lc;
                    if ((K & 3) == 0 && k0 + lc + 3 < K) {
                        float4 v = *reinterpret_cast<const float4*>(src);
                        qv[0] = v.x; qv[1] = v.y; qv[2] = v.z; qv[3] = v.w;
                    } else {
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            if (k0 + lc + x < K) qv[x] = src[x];
                    }
                }
                int64_t e = n0 + lr;
                if (e < e1) {
                    const float* src = P.ent_local + (size_t)(e - P.row_begin) * K + k0 + lc;
                    if ((K & 3) == 0 && k0 + lc + 3 < K) {
                        float4 v = *reinterpret_cast<const float4*>(src);
                        ev[0] = v.x; ev[1] = v.y; ev[2] = v.z; ev[3] = v.w;
                    } else {
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            if (k0 + lc + x < K) ev[x] = src[x];
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                Qs[lc + x][lr] = qv[x];
                Es[lc + x][lr] = ev[x];
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < SW_BK; ++kk) {
                float4 a4 = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);
                float4 b4 = *reinterpret_cast<const float4*>(&Es[kk][tx * 4]);
                float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (MODE == 0) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                        else if (MODE == 1) acc[i][j] += fabsf(a[i] - b[j]);
                        else {
                            float d = a[i] - b[j];
                            acc[i][j] = fmaf(d, d, acc[i][j]);
                        }
                    }
            }
        }
        // quantise, compare, count (zero-padded K tail contributes 0 in every mode)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = ty * 4 + i;
            const int32_t pq = s_posq[rl];
            const int32_t self = s_self[rl];
            const unsigned long long mk = s_mask[rl];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cl = tx * 4 + j;
                const int64_t e = n0 + cl;
                float sc = MODE == 0 ? acc[i][j] : (MODE == 1 ? -acc[i][j] : -sqrtf(acc[i][j]));
                int qv2 = quantise_score(apply_nl(P.nl, sc));
                bool valid = (e < e1) && (e != (int64_t)self) && (self >= 0);
                int gt = valid && (qv2 > pq), eq = valid && (qv2 == pq);
                int f = (int)((mk >> cl) & 1ull);
                cnt[i][0] += gt;
                cnt[i][1] += eq;
                cnt[i][2] += gt & f;
                cnt[i][3] += eq & f;
            }
        }
    }
    // reduce over the 16 threads (tx) that share a row, then one atomic per (row, counter)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int v = cnt[i][c];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            cnt[i][c] = v;
        }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t r = m0 + ty * 4 + i;
            if (r < m_end) {
                int side = r >= P.T ? 1 : 0;
                int64_t t = r - (int64_t)side * P.T;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (cnt[i][c]) atomicAdd(&P.counts[(t * 2 + side) * 4 + c], cnt[i][c]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TransE distance sweep, second generation (MODE 1: L1, MODE 2: L2).  The bound is the fp32 ALU pipe --
// two FADDs per (query, entity, column) for L1 (subtract, then add with the |x| source modifier), FADD +
// FFMA for L2 -- so everything else is kept off the issue slots:
//   * 128 x 64 CTA tile, 8 x 4 outputs per thread: 3 LDS.128 feed 64 ALU instructions per column
//     (the 64 x 64 / 4 x 4 kernel above: 2 per 32);
//   * two shared-memory stages, the next k-tile's global loads are issued before the current tile's
//     arithmetic and stored after it: ONE __syncthreads per k-tile instead of two, no exposed load latency;
//   * the last k-tile runs K - k0 columns instead of a zero-padded 16 (K = 100: 100 instead of 112);
//   * the filter mask of an entity tile is double-buffered by tile parity, so its rebuild needs no barrier.
//   * the epilogue compares in the float domain: trunc(y) >= n and trunc(y) > n (y = score * 1e5, n the
//     positive's quantised score) are each ONE float comparison against a per-row threshold computed once per
//     CTA (kge_quant_thresholds), so no F2I and no 64-bit index compares sit between two tiles' FADD streams;
//     the candidate that is the test triple itself and filter hits are patched in on a rare path.
// The columns of a score are accumulated in ascending order by one accumulator, exactly as above: both
// kernels produce bit-identical scores, hence identical counts (NaN scores excepted: F2I maps NaN to 0, a
// float comparison to "not counted").
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float f32_at_least(long long m, bool strictly) {
    // smallest fp32 value >= m (strictly: > m); m is exact in double
    float f = __ll2float_rn(m);
    if (strictly ? (double)f <= (double)m : (double)f < (double)m) f = nextafterf(f, INFINITY);
    return f;
}
// With q(y) = trunc_toward_zero(y) (saturating F2I) and n = the positive's quantised score:
//   q(y) >= n  <=>  y >= t_ge,   q(y) > n  <=>  y >= t_gt      (models/EmbeddingModel.py:2010-2029)
// because trunc(y) >= m <=> y >= m for an integer m > 0 and <=> y > m - 1 for m <= 0.
__device__ __forceinline__ void kge_quant_thresholds(int32_t n, float& t_ge, float& t_gt) {
    const long long m0 = n, m1 = (long long)n + 1;
    t_ge = m0 > 0 ? f32_at_least(m0, false) : f32_at_least(m0 - 1, true);
    // nothing is greater than a saturated positive: NaN compares false against everything
    t_gt = n == INT32_MAX ? __int_as_float(0x7fc00000) : (m1 > 0 ? f32_at_least(m1, false) : f32_at_least(m1 - 1, true));
}

#define S2_BM 128
#define S2_BN 64
#define S2_BK 16

// LINEAR: no score non-linearity (the common case) -- keeps tanhf/expf/logf out of the instruction stream
template <int MODE, bool LINEAR>
__global__ void __launch_bounds__(256, 2) kge_rank_sweep2_kernel(SweepParams P) {
    __shared__ __align__(16) float Qs[2][S2_BK][S2_BM + 4];
    __shared__ __align__(16) float Es[2][S2_BK][S2_BN + 4];
    __shared__ unsigned long long s_mask[2][S2_BM];
    __shared__ int32_t s_cur[S2_BM], s_hi[S2_BM], s_self[S2_BM];
    __shared__ float s_tge[S2_BM], s_tgt[S2_BM];
    __shared__ const int32_t* s_list[S2_BM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = P.q_row0 + (int64_t)blockIdx.x * S2_BM;
    const int64_t m_end = P.q_row0 + P.q_rows;
    const int64_t e0 = P.row_begin + (int64_t)blockIdx.y * P.chunk;
    const int64_t e1 = min(P.row_end, e0 + P.chunk);
    if (e0 >= e1) return;
    const int K = P.K;
    const bool vec_ok = (K & 3) == 0;
    const int ktiles = (K + S2_BK - 1) / S2_BK;

    if (tid < S2_BM) {
        const int64_t r = m0 + tid;
        int32_t cur = 0, hi = 0, self = -1;
        float tge = __int_as_float(0x7fc00000), tgt = tge;  // rows past the end never count (NaN compares false)
        const int32_t* list = nullptr;
        if (r < m_end) {
            const int side = r >= P.T ? 1 : 0;
            const int64_t t = r - (int64_t)side * P.T;
            self = side == 0 ? P.test[3 * t + 2] : P.test[3 * t + 0];
            kge_quant_thresholds(P.pos_q[t], tge, tgt);
            list = side == 0 ? P.sp_ent : P.po_ent;
            int32_t a = P.excl_lo[r], b = P.excl_hi[r];
            hi = b;
            while (a < b) {  // first filter entry >= e0
                const int32_t mid = (a + b) >> 1;
                if ((int64_t)list[mid] < e0) a = mid + 1;
                else b = mid;
            }
            cur = a;
        }
        s_cur[tid] = cur;
        s_hi[tid] = hi;
        s_self[tid] = self;
        s_tge[tid] = tge;
        s_tgt[tid] = tgt;
        s_list[tid] = list;
    }
    int cnt[8][4];  // per query row: {gt, ge, gt & filtered, ge & filtered}; eq = ge - gt at the end
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) cnt[i][c] = 0;

    // loader mapping: 256 threads x float4: rows lr and lr + 64 of the query tile, row lr of the entity tile
    const int lr = tid >> 2, lc = (tid & 3) * 4;
    const float* qrow0 = P.q + (size_t)min(m0 + lr, m_end - 1) * K;       // clamped: rows past the end are
    const float* qrow1 = P.q + (size_t)min(m0 + lr + 64, m_end - 1) * K;  // computed and never counted

    auto load4 = [&](float (&v)[4], const float* row, int k0) {
        const int c = k0 + lc;
        if (vec_ok && c + 3 < K) {
            const float4 t4 = *reinterpret_cast<const float4*>(row + c);
            v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
        } else {
#pragma unroll
            for (int x = 0; x < 4; ++x) v[x] = c + x < K ? row[c + x] : 0.f;
        }
    };

    int par = 0;
    for (int64_t n0 = e0; n0 < e1; n0 += S2_BN, par ^= 1) {
        if (tid < S2_BM) {
            unsigned long long mk = 0ull;
            int32_t cur = s_cur[tid];
            const int32_t hi = s_hi[tid];
            const int32_t* list = s_list[tid];
            while (cur < hi) {
                const int64_t e = list[cur];
                if (e >= n0 + S2_BN) break;
                mk |= 1ull << (int)(e - n0);
                ++cur;
            }
            s_cur[tid] = cur;
            s_mask[par][tid] = mk;
        }
        const float* erow = P.ent_local + (size_t)(min(n0 + lr, e1 - 1) - P.row_begin) * K;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        float q0[4], q1[4], ev[4];
        load4(q0, qrow0, 0);
        load4(q1, qrow1, 0);
        load4(ev, erow, 0);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            Qs[0][lc + x][lr] = q0[x];
            Qs[0][lc + x][lr + 64] = q1[x];
            Es[0][lc + x][lr] = ev[x];
        }
        __syncthreads();
        for (int kt = 0; kt < ktiles; ++kt) {
            const int cur = kt & 1;
            const bool more = kt + 1 < ktiles;
            if (more) {
                load4(q0, qrow0, (kt + 1) * S2_BK);
                load4(q1, qrow1, (kt + 1) * S2_BK);
                load4(ev, erow, (kt + 1) * S2_BK);
            }
            auto column = [&](int kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&Qs[cur][kk][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&Qs[cur][kk][ty * 8 + 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Es[cur][kk][tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d = a[i] - b[j];
                        if (MODE == 1) acc[i][j] += fabsf(d);
                        else acc[i][j] = fmaf(d, d, acc[i][j]);
                    }
            };
            const int kend = min(S2_BK, K - kt * S2_BK);
            if (kend == S2_BK) {
#pragma unroll
                for (int kk = 0; kk < S2_BK; ++kk) column(kk);
            } else {
                for (int kk = 0; kk < kend; ++kk) column(kk);
            }
            if (more) {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    Qs[cur ^ 1][lc + x][lr] = q0[x];
                    Qs[cur ^ 1][lc + x][lr + 64] = q1[x];
                    Es[cur ^ 1][lc + x][lr] = ev[x];
                }
            }
            __syncthreads();
        }
        // compare against the per-row thresholds, count.  lim = this thread's columns inside [n0, e1)
        const int lim = (int)min((int64_t)4, e1 - n0 - tx * 4);
        const int col0 = (int)(n0 - P.row_begin) + tx * 4;  // local index of this thread's first candidate
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = ty * 8 + i;
            const float tge = s_tge[rl], tgt = s_tgt[rl];
            bool g[4], h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float sc = MODE == 1 ? -acc[i][j] : -sqrtf(acc[i][j]);
                const float y = __fmul_rn(LINEAR ? sc : apply_nl(P.nl, sc), 1e5f);
                g[j] = y >= tgt;
                h[j] = y >= tge;
            }
            if (lim < 4) {  // last tile of the range (block-uniform per tx; rare)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = g[j] && j < lim;
                    h[j] = h[j] && j < lim;
                }
            }
            cnt[i][0] += (int)g[0] + (int)g[1] + (int)g[2] + (int)g[3];
            cnt[i][1] += (int)h[0] + (int)h[1] + (int)h[2] + (int)h[3];
            // rare: the test triple's own entity is one of these candidates (never counted), or filter hits
            const int selfrel = s_self[rl] - (int)P.row_begin - col0;
            const unsigned fm = (unsigned)(s_mask[par][rl] >> (tx * 4)) & 0xfu;
            if ((unsigned)selfrel < 4u || fm != 0u) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j == selfrel) {
                        cnt[i][0] -= (int)g[j];
                        cnt[i][1] -= (int)h[j];
                    } else if ((fm >> j) & 1u) {
                        cnt[i][2] += (int)g[j];
                        cnt[i][3] += (int)h[j];
                    }
                }
            }
        }
    }
    // reduce over the 16 threads (tx) that share a query row, then one atomic per (row, counter)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int v = cnt[i][c];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            cnt[i][c] = v;
        }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t r = m0 + ty * 8 + i;
            if (r < m_end) {
                const int side = r >= P.T ? 1 : 0;
                const int64_t t = r - (int64_t)side * P.T;
                // counts layout: {gt, eq, gt_filtered, eq_filtered}
                const int out[4] = {cnt[i][0], cnt[i][1] - cnt[i][0], cnt[i][2], cnt[i][3] - cnt[i][2]};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (out[c]) atomicAdd(&P.counts[(t * 2 + side) * 4 + c], out[c]);
            }
        }
    }
}

// KGE_SWEEP_V1=1 selects the first-generation 64 x 64 kernel for TransE as well (A/B)
static inline bool sweep_v1_forced() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_SWEEP_V1");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

// ------------------------------------------------------------------------------------------------
// rank assembly (models/EmbeddingModel.py:1966-1986 with perform_comparision :1989-2033)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cmp_count(int gt, int eq, int strategy) {
    if (strategy == KGE_STRAT_BEST) return gt;
    if (strategy == KGE_STRAT_MIDDLE) return gt + (eq + 1) / 2;
    return gt + eq;
}

__global__ void kge_rank_finalize_kernel(const int32_t* __restrict__ counts, int64_t T, int side, int strategy,
                                         int filtered, const uint8_t* __restrict__ self_cand, int32_t* __restrict__ ranks) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= T) return;
    // object sweep = side 0, subject sweep = side 1; +1 on eq: the test triple's own candidate (when its
    // entity is among the swept candidates at all)
    const int32_t* co = counts + (t * 2 + 0) * 4;
    const int32_t* cs = counts + (t * 2 + 1) * 4;
    const int so = self_cand != nullptr ? (self_cand[2 * t + 1] ? 1 : 0) : 1;
    const int ss = self_cand != nullptr ? (self_cand[2 * t + 0] ? 1 : 0) : 1;
    int go = co[0], eo = co[1] + so, gfo = filtered ? co[2] : 0, efo = filtered ? co[3] + so : 0;
    int gs = cs[0], es = cs[1] + ss, gfs = filtered ? cs[2] : 0, efs = filtered ? cs[3] + ss : 0;
    int fo = filtered ? cmp_count(gfo, efo, strategy) : 0;
    int fs = filtered ? cmp_count(gfs, efs, strategy) : 0;
    if (side == KGE_RANK_S_O) {
        ranks[2 * t + 0] = cmp_count(gs, es, strategy) + 1 - fs;
        ranks[2 * t + 1] = cmp_count(go, eo, strategy) + 1 - fo;
    } else if (side == KGE_RANK_SPO) {
        ranks[t] = cmp_count(go + gs, eo + es, strategy) + 1 - fs - fo;
    } else if (side == KGE_RANK_S) {
        ranks[t] = cmp_count(gs, es, strategy) + 1 - fs;
    } else {
        ranks[t] = cmp_count(go, eo, strategy) + 1 - fo;
    }
}

static int rank_counts_impl(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                            const float* ent_local, int64_t row_begin, int64_t row_end, const int32_t* test, int64_t T,
                            int side, int filtered, int use_tensor_cores, int non_linearity, int32_t* counts, void* stream,
                            const float* s_rows, const float* o_rows);

extern "C" int kge_rank_counts(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                               const float* ent_local, int64_t row_begin, int64_t row_end, const int32_t* test, int64_t T,
                               int side, int filtered, int use_tensor_cores, int non_linearity, int32_t* counts, void* stream) {
    return rank_counts_impl(ctx, model, k, ent, rel, R, ent_local, row_begin, row_end, test, T, side, filtered, use_tensor_cores,
                            non_linearity, counts, stream, nullptr, nullptr);
}

extern "C" int kge_rank_counts_rows(kge_ctx* ctx, int model, int k, int64_t E, const float* rel, int64_t R,
                                    const float* s_rows, const float* o_rows, const float* ent_local, int64_t row_begin,
                                    int64_t row_end, const int32_t* test, int64_t T, int side, int filtered,
                                    int use_tensor_cores, int non_linearity, int32_t* counts, void* stream) {
    KGE_REQUIRE(T == 0 || (s_rows != nullptr && o_rows != nullptr), "kge_rank_counts_rows: subject / object rows missing");
    kge_table tb;
    memset(&tb, 0, sizeof(tb));
    tb.rows = E;
    tb.rows_per_shard = E;
    tb.n_shards = 1;
    tb.K = model_row_width(model, k);
    return rank_counts_impl(ctx, model, k, &tb, rel, R, ent_local, row_begin, row_end, test, T, side, filtered, use_tensor_cores,
                            non_linearity, counts, stream, s_rows, o_rows);
}

static int rank_counts_impl(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                            const float* ent_local, int64_t row_begin, int64_t row_end, const int32_t* test, int64_t T,
                            int side, int filtered, int use_tensor_cores, int non_linearity, int32_t* counts, void* stream,
                            const float* s_rows, const float* o_rows) {
    KGE_REQUIRE(ctx != nullptr, "kge_rank_counts: null ctx");
    KGE_REQUIRE(non_linearity >= KGE_NL_LINEAR && non_linearity <= KGE_NL_SOFTPLUS, "Invalid non-linearity");
    KGE_REQUIRE(model >= KGE_TRANSE_L1 && model <= KGE_HOLE, "kge_rank_counts: unknown model %d", model);
    KGE_REQUIRE(side >= KGE_RANK_S_O && side <= KGE_RANK_O, "Invalid value for corrupt_side.");
    if (T == 0) return 0;
    KGE_REQUIRE(ent && rel && ent_local && counts, "kge_rank_counts: null tensor");
    KGE_REQUIRE(ent->K == model_row_width(model, k), "kge_rank_counts: table width %d != internal_k %d", ent->K,
                model_row_width(model, k));
    KGE_REQUIRE(row_begin >= 0 && row_end <= ent->rows && row_begin <= row_end, "kge_rank_counts: bad row range");
    KGE_REQUIRE(T < (int64_t)(1 << 30), "kge_rank_counts: too many test triples");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) return 0;
    KGE_REQUIRE(test != nullptr, "kge_rank_counts: null test triples");
    if (filtered) {
        KGE_REQUIRE(ctx->f_valid, "kge_rank_counts: filtered ranking requested but no filter was built");
        KGE_REQUIRE(ctx->f_E == ent->rows && ctx->f_R == R, "kge_rank_counts: filter built for E=%lld R=%lld, model has E=%lld R=%lld",
                    (long long)ctx->f_E, (long long)ctx->f_R, (long long)ent->rows, (long long)R);
    }
    const int K = ent->K;
    KGE_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)T * 8 * sizeof(int32_t), st));
    if (ctx->q_fold.reserve((size_t)2 * T * K * sizeof(float))) return -2;
    if (ctx->pos_q.reserve((size_t)T * 4) || ctx->excl_lo.reserve((size_t)2 * T * 4) || ctx->excl_hi.reserve((size_t)2 * T * 4)) return -2;
    PrepParams pp;
    pp.model = model;
    pp.k = k;
    pp.ent = make_view(*ent);
    pp.rel = rel;
    pp.test = test;
    pp.s_rows = s_rows;
    pp.o_rows = o_rows;
    pp.T = T;
    pp.filtered = filtered && ctx->f_n_sp > 0;
    pp.nl = non_linearity;
    ctx->rank_nl = non_linearity;
    pp.sp_comp = ctx->f_sp_comp.as<uint64_t>();
    pp.po_comp = ctx->f_po_comp.as<uint64_t>();
    pp.f_cnt = ctx->f_count.as<int32_t>();
    pp.E = ent->rows;
    pp.R = R;
    pp.q = ctx->q_fold.as<float>();
    pp.pos_q = ctx->pos_q.as<int32_t>();
    pp.excl_lo = ctx->excl_lo.as<int32_t>();
    pp.excl_hi = ctx->excl_hi.as<int32_t>();
    pp.q_absmax = nullptr;
    pp.amax_mask = side == KGE_RANK_O ? 1 : (side == KGE_RANK_S ? 2 : 3);
    const bool tc_path = use_tensor_cores && !(model == KGE_TRANSE_L1 || model == KGE_TRANSE_L2);
    if (tc_path && kge_rank_tc_f16() && row_end > row_begin) {
        // the fp16 split scales the queries by a power of two taken from their largest magnitude: found here, while they are written
        if (ctx->q_lo.reserve(2 * sizeof(uint32_t))) return -2;
        KGE_CUDA_CHECK(cudaMemsetAsync(ctx->q_lo.p, 0, 2 * sizeof(uint32_t), st));
        pp.q_absmax = ctx->q_lo.as<uint32_t>();
    }
    {
        const int warps = 8;
        kge_rank_prepare_kernel<<<(unsigned)((T + warps - 1) / warps), warps * 32, 0, st>>>(pp);
        KGE_CUDA_CHECK(cudaGetLastError());
    }
    if (row_end == row_begin) return 0;
    // which query rows: [0,T) object sweep, [T,2T) subject sweep
    int64_t q_row0 = 0, q_rows = 2 * T;
    if (side == KGE_RANK_O) q_rows = T;
    if (side == KGE_RANK_S) { q_row0 = T; q_rows = T; }
    const bool trilinear = !(model == KGE_TRANSE_L1 || model == KGE_TRANSE_L2);
    if (use_tensor_cores) {
        KGE_REQUIRE(trilinear, "kge_rank_counts: the tensor-core sweep covers DistMult/ComplEx/HolE; TransE uses the fp32 sweep");
        int side_mask = side == KGE_RANK_O ? 1 : (side == KGE_RANK_S ? 2 : 3);
        return kge_rank_sweep_tc(ctx, model, K, pp.q, 2 * T, T, ent_local, row_begin, row_end, test, pp.pos_q, pp.excl_lo,
                                 pp.excl_hi, ctx->f_sp_ent.as<int32_t>(), ctx->f_po_ent.as<int32_t>(), side_mask, counts, pp.q_absmax, st);
    }
    SweepParams sp;
    sp.model = model;
    sp.K = K;
    sp.q = pp.q;
    sp.NQ = 2 * T;
    sp.T = T;
    sp.ent_local = ent_local;
    sp.row_begin = row_begin;
    sp.row_end = row_end;
    sp.test = test;
    sp.pos_q = pp.pos_q;
    sp.excl_lo = pp.excl_lo;
    sp.excl_hi = pp.excl_hi;
    sp.sp_ent = ctx->f_sp_ent.as<int32_t>();
    sp.po_ent = ctx->f_po_ent.as<int32_t>();
    sp.q_row0 = q_row0;
    sp.q_rows = q_rows;
    sp.counts = counts;
    sp.nl = non_linearity;
    int64_t n_ent = row_end - row_begin;
    if (!trilinear && !sweep_v1_forced()) {
        const int64_t row_tiles2 = (q_rows + S2_BM - 1) / S2_BM;
        // two 256-thread CTAs per SM; ~16 waves of CTAs so that the last, partly filled wave costs a few percent
        // (4 waves left the SMs idle 10 % of the kernel, ncu r01_t); entity chunks a multiple of the tile
        int64_t want = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 32 + row_tiles2 - 1) / row_tiles2);
        int64_t chunk2 = (n_ent + want - 1) / want;
        chunk2 = std::max<int64_t>(S2_BN * 4, ((chunk2 + S2_BN - 1) / S2_BN) * S2_BN);
        const int64_t n_chunks2 = (n_ent + chunk2 - 1) / chunk2;
        KGE_REQUIRE(n_chunks2 <= 65535, "kge_rank_counts: too many entity chunks");
        sp.chunk = chunk2;
        dim3 grid2((unsigned)row_tiles2, (unsigned)n_chunks2), block2(256);
        const bool lin = non_linearity == KGE_NL_LINEAR;
        if (model == KGE_TRANSE_L1) {
            if (lin) kge_rank_sweep2_kernel<1, true><<<grid2, block2, 0, st>>>(sp);
            else kge_rank_sweep2_kernel<1, false><<<grid2, block2, 0, st>>>(sp);
        } else {
            if (lin) kge_rank_sweep2_kernel<2, true><<<grid2, block2, 0, st>>>(sp);
            else kge_rank_sweep2_kernel<2, false><<<grid2, block2, 0, st>>>(sp);
        }
        KGE_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    int64_t row_tiles = (q_rows + SW_BM - 1) / SW_BM;
    // enough CTAs for >= 4 waves, entity chunks a multiple of the tile
    int64_t want_chunks = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 8 + row_tiles - 1) / row_tiles);
    int64_t chunk = (n_ent + want_chunks - 1) / want_chunks;
    chunk = std::max<int64_t>(SW_BN * 4, ((chunk + SW_BN - 1) / SW_BN) * SW_BN);
    int64_t n_chunks = (n_ent + chunk - 1) / chunk;
    KGE_REQUIRE(n_chunks <= 65535, "kge_rank_counts: too many entity chunks");
    sp.chunk = chunk;
    dim3 grid((unsigned)row_tiles, (unsigned)n_chunks), block(256);
    if (trilinear) kge_rank_sweep_kernel<0><<<grid, block, 0, st>>>(sp);
    else if (model == KGE_TRANSE_L1) kge_rank_sweep_kernel<1><<<grid, block, 0, st>>>(sp);
    else kge_rank_sweep_kernel<2><<<grid, block, 0, st>>>(sp);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_rank_finalize(kge_ctx* ctx, const int32_t* counts, int64_t T, int side, int strategy, int filtered,
                                 const uint8_t* self_is_candidate,
                                 int32_t* ranks_out, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_rank_finalize: null ctx");
    KGE_REQUIRE(side >= KGE_RANK_S_O && side <= KGE_RANK_O, "Invalid value for corrupt_side.");
    KGE_REQUIRE(strategy >= KGE_STRAT_WORST && strategy <= KGE_STRAT_MIDDLE, "Invalid ranking_strategy!");
    if (T == 0) return 0;
    KGE_REQUIRE(counts && ranks_out, "kge_rank_finalize: null tensor");
    kge_rank_finalize_kernel<<<(unsigned)((T + 255) / 256), 256, 0, (cudaStream_t)stream>>>(counts, T, side, strategy, filtered, self_is_candidate,
                                                                                        ranks_out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_rank_host(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                             const int32_t* test_host, int64_t T, int side, int strategy, int filtered,
                             int use_tensor_cores, int non_linearity, int32_t* ranks_host, void* stream) {
    KGE_REQUIRE(ctx != nullptr && ent != nullptr, "kge_rank_host: null argument");
    KGE_REQUIRE(ent->n_shards == 1 && ent->shard[0] != nullptr, "kge_rank_host is the single-GPU entry; use kge_rank_counts per shard");
    if (T == 0) return 0;
    KGE_REQUIRE(test_host != nullptr && ranks_host != nullptr, "kge_rank_host: null host buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_out = (size_t)T * (side == KGE_RANK_S_O ? 2 : 1);
    if (ctx->h_test.reserve((size_t)T * 3 * 4) || ctx->h_counts.reserve((size_t)T * 8 * 4) || ctx->h_ranks.reserve(n_out * 4)) return -2;
    KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_test.p, test_host, (size_t)T * 3 * 4, cudaMemcpyHostToDevice, st));
    if (int rc = kge_rank_counts(ctx, model, k, ent, rel, R, ent->shard[0], 0, ent->rows, ctx->h_test.as<int32_t>(), T, side,
                                 filtered, use_tensor_cores, non_linearity, ctx->h_counts.as<int32_t>(), stream))
        return rc;
    if (int rc = kge_rank_finalize(ctx, ctx->h_counts.as<int32_t>(), T, side, strategy, filtered, nullptr, ctx->h_ranks.as<int32_t>(), stream))
        return rc;
    KGE_CUDA_CHECK(cudaMemcpyAsync(ranks_host, ctx->h_ranks.p, n_out * 4, cudaMemcpyDeviceToHost, st));
    KGE_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}
