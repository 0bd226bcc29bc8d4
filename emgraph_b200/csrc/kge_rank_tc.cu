// tcgen05 split-precision ranking sweep for sm_100a (DistMult / ComplEx / HolE).
//
// S[m,e] = sum_k Q[m,k] * Ent[e,k]  with Q the folded queries of the test triples is a
// dense [2T,K] x [K,E] contraction.  fp32 accuracy is kept with a three-term split: every operand x is scaled by a
// power of two (one per operand, from its largest magnitude, so that the low halves stay in the normal fp16 range)
// and written as hi = fp16(x), lo = fp16(x - hi) -- 11 + 11 significant bits, what 3xTF32 carries -- and each
// k-step issues
//     D += Qlo*Ehi ; D += Qhi*Elo ; D += Qhi*Ehi           (kind::f16, fp32 accumulate in TMEM)
// at twice the tensor-pipe rate of kind::tf32 and half the operand bytes; the epilogue multiplies by the exact
// inverse of the two scales.  KGE_RANK_TF32=1 selects the round-1 3xTF32 operands (A/B).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer : cp.async.bulk.tensor (swizzled) of the 4 operand tiles of a k-block (Qhi,Qlo: 128 rows,
//                              Ehi,Elo: 256 rows, one swizzle row wide) into a smem ring of 2 x 96 KB or 4 x 48 KB (TcCfg)
//   warp 1      MMA issuer   : tcgen05.mma.cta_group::1, M=128 N=256, K=16 (f16) / 8 (tf32) per instruction, 3 MMAs per
//                              k-step, accumulator = 128 lanes x 256 columns of TMEM, double buffered (512 cols).  The
//                              k-steps that lie wholly in the zero padding behind K are not issued (K = 400: 25 of 28).
//   warps 2..9  epilogue     : two warps per TMEM lane quarter, each owning half of the tile's columns (with one warp per
//                              scheduler the epilogue of a tile took longer than its MMAs: ~3.5 K dependent instructions
//                              at 0.3 IPC, measured r2ra); tcgen05.ld 32x32b.x32 -> each thread owns ONE query row and 32
//                              candidate scores; x1e5 int truncation (F7), compare with the positive's quantised
//                              score, filter bitmask from the sorted known-triple list, popcount into the
//                              four per-row counters.  The [T,2E] score matrix never leaves the SM.
// The epilogue of tile i overlaps the MMAs of tile i+1.  Work split: the (query tile, entity tile) pairs in query-major
// order, cut into one contiguous range per CTA (ranges differ by at most one tile).
//
// Replaces reference models/EmbeddingModel.py:1856-1866 (score all corruptions), :1942-1986 (filter +
// rank) and :1989-2033 (perform_comparision) for the trilinear models.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "kge_common.cuh"

#define TC_BM 128
#define TC_BN 256
// epilogue warps EW: 8 = two per TMEM lane quarter, each owning half of the tile's columns (default); 4 = one per quarter
// (KGE_RANK_EW=4, the round-2 kernel, kept for A/B)
#define TC_THREADS_MAX (64 + 32 * 8)

// The k-block of a pipeline stage is one swizzle row wide: SWB = 128 bytes (64 fp16 / 32 tf32 columns, 4 MMA k-steps, 96 KB
// per stage, 2 stages) or 64 bytes (32 fp16 columns, 2 k-steps, 48 KB per stage, 4 stages: three loads in flight behind the
// block the tensor pipe is working on, which hides the TMA latency of the short last block of a K that is not a multiple
// of the block width).
#ifndef KGE_RANK_GROUP_DEFAULT
#define KGE_RANK_GROUP_DEFAULT 1
#endif
#ifndef KGE_RANK_SW_DEFAULT
#define KGE_RANK_SW_DEFAULT 64
#endif
template <int SWB>
struct TcCfg {
    static constexpr int STAGES = SWB == 128 ? 2 : 4;
    static constexpr int KSTEPS = SWB / 32;                   // one MMA covers 32 bytes along K: 16 fp16 or 8 tf32
    static constexpr int A_BYTES = TC_BM * SWB;               // 16 KB / 8 KB
    static constexpr int B_BYTES = TC_BN * SWB;               // 32 KB / 16 KB
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool F16>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled shared-memory matrix descriptor (sm_100 UMMA) of a tile whose rows are one swizzle row (SWB bytes) wide:
// start address >> 4, LBO = 1 (unused for swizzled K-major), SBO = 8 rows x SWB bytes >> 4, version 1, layout type 2
// (SWIZZLE_128B) or 4 (SWIZZLE_64B).
template <int SWB>
__device__ __forceinline__ uint64_t make_sw_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * SWB) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(SWB == 128 ? 2 : 4) << 61;
    return d;
}

// instruction descriptor: D=F32 (bits 4-5 = 1), A=B=TF32 (bits 7-9, 10-12 = 2), K-major A and B,
// N>>3 at bits 17-22, M>>4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with A = B = F16 (format code 0), D = F32
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// operand preparation: hi/lo split into zero-padded [2*rows_pad, Kp] (rows [0,rows_pad) hi, then lo)
// ------------------------------------------------------------------------------------------------
__global__ void kge_tf32_split_kernel(const float* __restrict__ src, int64_t rows, int K, int Kp, int64_t rows_pad,
                                      float* __restrict__ dst) {
    const int64_t total = rows_pad * (int64_t)Kp;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / Kp;
        const int c = (int)(t - r * Kp);
        float x = (r < rows && c < K) ? src[r * (int64_t)K + c] : 0.f;
        float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        dst[t] = hi;
        dst[total + t] = x - hi;
    }
}

// largest magnitude of an operand, as the bit pattern of a non-negative float (orders like an unsigned int).
// VEC: 16-byte loads (n a multiple of 4, src 16-byte aligned)
template <bool VEC>
__global__ void kge_absmax_kernel(const float* __restrict__ src, int64_t n, uint32_t* __restrict__ out) {
    float m = 0.f;
    if constexpr (VEC) {
        const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
        const int64_t n4 = n >> 2;
#pragma unroll 4
        for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
            const float4 x = __ldg(s4 + t);
            m = fmaxf(fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        }
    } else {
        for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(src[t]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && __float_as_uint(m) > *(volatile uint32_t*)out) atomicMax(out, __float_as_uint(m));
}

// power of two that brings the operand's largest magnitude into [2^13, 2^14): the low halves of elements down to
// 2^-17 of the maximum are then normal fp16 numbers; smaller elements lose low-half bits that are below 2^-38 of the
// largest product and cannot move a score.  1.0 for an all-zero operand.
__device__ __forceinline__ float split_scale(uint32_t absmax_bits) {
    if (absmax_bits == 0u) return 1.f;
    const int e = (int)((absmax_bits >> 23) & 0xffu) - 127;  // max in [2^e, 2^(e+1))
    return __uint_as_float((uint32_t)(127 + 13 - e) << 23);
}

// fp16 hi/lo split into zero-padded [2*rows_pad, Kp] halves (rows [0,rows_pad) hi, then lo).  One thread per 8 consecutive
// columns of a row (Kp is a multiple of 64): two 16-byte loads when VEC (K a multiple of 4, src 16-byte aligned), one
// 16-byte store into each half.
template <bool VEC>
__global__ void kge_f16_split_kernel(const float* __restrict__ src, int64_t rows, int K, int Kp, int64_t rows_pad,
                                     const uint32_t* __restrict__ absmax, __half* __restrict__ dst) {
    const int c8n = Kp >> 3;
    const int64_t total8 = rows_pad * (int64_t)c8n;
    const float sc = split_scale(*absmax);
    uint4* __restrict__ d_hi = reinterpret_cast<uint4*>(dst);
    uint4* __restrict__ d_lo = reinterpret_cast<uint4*>(dst + rows_pad * (int64_t)Kp);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total8; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / c8n;
        const int c = (int)(t - r * c8n) << 3;
        float x[8];
        if (r < rows && c < K) {
            const float* __restrict__ row = src + r * (int64_t)K + c;
            if (VEC && c + 8 <= K) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(row)), b = __ldg(reinterpret_cast<const float4*>(row) + 1);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
                x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = c + i < K ? row[i] : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.f;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x0 = x[2 * i] * sc, x1 = x[2 * i + 1] * sc;  // exact: power of two
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
            h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        d_hi[t] = make_uint4(h[0], h[1], h[2], h[3]);
        d_lo[t] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

struct TcParams {
    const uint32_t* absmax;  // [2]: bit patterns of max|Q|, max|Ent| (fp16 split: the epilogue undoes the two scales)
    int k_blocks;
    int tail_mmas;    // MMA k-steps of the LAST k-block that hold real columns (the rest of the block is zero padding: skipped)
    int64_t M;        // query rows handled (after side selection)
    int64_t Mp, Np;   // padded row counts of the split operands
    int64_t T;
    int64_t q_row0;   // first query row handled: 0 (object sweep first) or T
    int64_t row_begin, row_end;  // global entity ids of local rows [0, row_end-row_begin)
    const int32_t* test;
    const int32_t* pos_q;
    const int32_t* excl_lo;
    const int32_t* excl_hi;
    const int32_t* sp_ent;
    const int32_t* po_ent;
    int32_t* counts;
    int n_m_tiles, n_n_tiles;
    int group;  // CTAs that share a tile range (1: every CTA has its own)
    int nl;  // KGE_NL_*: non-linearity applied to the scores before the quantisation
};

template <bool F16, int SWB, int TC_EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * TC_EPI_WARPS, 1)
kge_rank_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmE, TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    using C = TcCfg<SWB>;
    constexpr int TC_STAGES = C::STAGES, TC_STAGE_BYTES = C::STAGE_BYTES, TC_A_BYTES = C::A_BYTES, TC_B_BYTES = C::B_BYTES;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full = bars;                    // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;       // [TC_STAGES]
    uint64_t* tfull = bars + 2 * TC_STAGES;   // [2]
    uint64_t* tempty = bars + 2 * TC_STAGES + 2;  // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmE) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&tfull[b], 1);
                mbar_init(&tempty[b], TC_EPI_WARPS);  // one arrive per epilogue warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work split: the (query tile, entity tile) pairs in query-major order, one contiguous range per CTA -- the CTAs differ
    // by at most one tile; a range crosses a query-tile boundary a few times at most, where the epilogue flushes its row
    // P.group = G > 1 (KGE_RANK_GROUP): G neighbouring CTAs share a range and take its tiles in turn (CTA j: first + j, + G, ...),
    // so that at any time they work on the same query tile and on adjacent entity tiles
    const int64_t n_tiles = (int64_t)P.n_m_tiles * P.n_n_tiles;
    const int tstep = P.group;
    const int64_t n_groups = gridDim.x / tstep, gi = blockIdx.x / tstep;
    const int64_t tile0 = n_tiles * gi / n_groups + (blockIdx.x - gi * tstep), tile1 = n_tiles * (gi + 1) / n_groups;
    const int mt_first = (int)(tile0 / P.n_n_tiles), nt_first = (int)(tile0 - (int64_t)mt_first * P.n_n_tiles);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int mt = mt_first, nt = nt_first;
            for (int64_t tile = tile0; tile < tile1; tile += tstep) {
                const int m0 = mt * TC_BM, n0 = nt * TC_BN;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + (size_t)stage * TC_STAGE_BYTES;
                    mbar_expect_tx(&full[stage], TC_STAGE_BYTES);
                    constexpr int KB_ELEMS = SWB / (F16 ? 2 : 4);  // columns per k-block (TMA coordinates count elements)
                    tma_load_2d(st, &tmQ, kb * KB_ELEMS, m0, &full[stage]);
                    tma_load_2d(st + TC_A_BYTES, &tmQ, kb * KB_ELEMS, (int)P.Mp + m0, &full[stage]);
                    tma_load_2d(st + 2 * TC_A_BYTES, &tmE, kb * KB_ELEMS, n0, &full[stage]);
                    tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tmE, kb * KB_ELEMS, (int)P.Np + n0, &full[stage]);
                    if (++stage == TC_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                nt += tstep;
                while (nt >= P.n_n_tiles) {
                    nt -= P.n_n_tiles;
                    ++mt;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = F16 ? make_idesc_f16(TC_BM, TC_BN) : make_idesc_tf32(TC_BM, TC_BN);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t tile_it = 0;
            for (int64_t tile = tile0; tile < tile1; tile += tstep, ++tile_it) {
                const uint32_t as = tile_it & 1, aphase = (tile_it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * TC_BN;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * TC_STAGE_BYTES);
                    const uint64_t a_hi = make_sw_desc<SWB>(sa);
                    const uint64_t a_lo = make_sw_desc<SWB>(sa + TC_A_BYTES);
                    const uint64_t b_hi = make_sw_desc<SWB>(sa + 2 * TC_A_BYTES);
                    const uint64_t b_lo = make_sw_desc<SWB>(sa + 2 * TC_A_BYTES + TC_B_BYTES);
                    // the zero-padded columns behind K add exact zeros: their k-steps are not issued
                    const int nj = kb + 1 == P.k_blocks ? P.tail_mmas : C::KSTEPS;
#pragma unroll
                    for (int j = 0; j < C::KSTEPS; ++j) {
                        if (j < nj) {
                            const uint64_t off = (uint64_t)(j * 32 >> 4);  // one MMA = 32 bytes along K: 8 tf32 or 16 fp16
                            tc_mma<F16>(d_tmem, a_lo + off, b_hi + off, idesc, (kb | j) != 0);
                            tc_mma<F16>(d_tmem, a_hi + off, b_lo + off, idesc, 1);
                            tc_mma<F16>(d_tmem, a_hi + off, b_hi + off, idesc, 1);
                        }
                    }
                    tc_commit(&empty[stage]);  // smem stage reusable once these MMAs retire
                    if (++stage == TC_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&tfull[as]);  // accumulator complete
            }
        }
    } else {
        // epilogue: warp w may only touch TMEM lanes [32*(w%4), +32)
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        // the warps of one quarter split the tile's columns: this warp owns 32-column chunks [c_begin, c_end) of every tile
        constexpr int CHUNKS = TC_BN / 32, PER_WARP = CHUNKS / (TC_EPI_WARPS / 4);
        const int c_begin = ((warp - 2) >> 2) * PER_WARP, c_end = c_begin + PER_WARP;
        // fp16 split: the accumulator holds the score times the two operand scales (powers of two: the product of their
        // inverses is exact, and so is the multiplication unless the score is subnormal)
        const float unscale = F16 ? 1.f / (split_scale(P.absmax[0]) * split_scale(P.absmax[1])) : 1.f;
        uint32_t tile_it = 0;
        int mt = mt_first, nt = nt_first;
        bool fresh = true;  // a new query tile starts: this thread's row state is set up before its first entity tile
        bool row_ok = false;
        int32_t pq = 0, self = -1, cur = 0, hi = 0;
        const int32_t* list = nullptr;
        int64_t tt = 0, next_f = (int64_t)1 << 40;
        int side = 0;
        int c_gt = 0, c_eq = 0, c_gtf = 0, c_eqf = 0;
        for (int64_t tile = tile0; tile < tile1; tile += tstep, ++tile_it) {
            if (fresh) {
                fresh = false;
                const int64_t m = (int64_t)mt * TC_BM + row_in_tile;
                row_ok = m < P.M;
                pq = 0; self = -1; cur = 0; hi = 0;
                c_gt = c_eq = c_gtf = c_eqf = 0;
                if (row_ok) {
                    const int64_t e_first = P.row_begin + (int64_t)nt * TC_BN + c_begin * 32;
                    const int64_t r = P.q_row0 + m;
                    side = r >= P.T ? 1 : 0;
                    tt = r - (int64_t)side * P.T;
                    self = side == 0 ? P.test[3 * tt + 2] : P.test[3 * tt + 0];
                    pq = P.pos_q[tt];
                    list = side == 0 ? P.sp_ent : P.po_ent;
                    int32_t a = P.excl_lo[r], b = P.excl_hi[r];
                    hi = b;
                    while (a < b) {  // first known entity >= first candidate of this range
                        int32_t mid = (a + b) >> 1;
                        if ((int64_t)list[mid] < e_first) a = mid + 1;
                        else b = mid;
                    }
                    cur = a;
                }
                next_f = (cur < hi) ? (int64_t)list[cur] : (int64_t)1 << 40;
            }
            {
                const uint32_t as = tile_it & 1, aphase = (tile_it >> 1) & 1;
                mbar_wait(&tfull[as], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * TC_BN;
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t v[32];
                    tc_ld32(taddr + c * 32, v);
                    tc_wait_ld();
                    const int64_t e_base = P.row_begin + (int64_t)nt * TC_BN + c * 32;
                    uint32_t fmask = 0;
                    while (next_f < e_base + 32) {
                        if (next_f >= e_base) fmask |= 1u << (int)(next_f - e_base);
                        ++cur;
                        next_f = (cur < hi) ? (int64_t)list[cur] : (int64_t)1 << 40;
                    }
                    const int64_t left = P.row_end - e_base;  // valid candidates in this chunk
                    uint32_t valid = left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
                    const int64_t ds = (int64_t)self - e_base;
                    if (ds >= 0 && ds < 32) valid &= ~(1u << (int)ds);
                    uint32_t gtm = 0, eqm = 0;
                    if (P.nl == KGE_NL_LINEAR) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int q = quantise_score(F16 ? __uint_as_float(v[i]) * unscale : __uint_as_float(v[i]));
                            gtm |= (q > pq ? 1u : 0u) << i;
                            eqm |= (q == pq ? 1u : 0u) << i;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int q = quantise_score(apply_nl(P.nl, F16 ? __uint_as_float(v[i]) * unscale : __uint_as_float(v[i])));
                            gtm |= (q > pq ? 1u : 0u) << i;
                            eqm |= (q == pq ? 1u : 0u) << i;
                        }
                    }
                    gtm &= valid;
                    eqm &= valid;
                    c_gt += __popc(gtm);
                    c_eq += __popc(eqm);
                    c_gtf += __popc(gtm & fmask);
                    c_eqf += __popc(eqm & fmask);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
            }
            nt += tstep;
            while (nt >= P.n_n_tiles) {
                nt -= P.n_n_tiles;
                ++mt;
                fresh = true;
            }
            if ((fresh || tile + tstep >= tile1) && row_ok) {  // last entity tile of this row in this CTA's share
                int32_t* dst = P.counts + (tt * 2 + side) * 4;
                if (c_gt) atomicAdd(dst + 0, c_gt);
                if (c_eq) atomicAdd(dst + 1, c_eq);
                if (c_gtf) atomicAdd(dst + 2, c_gtf);
                if (c_eqf) atomicAdd(dst + 3, c_eqf);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_tensormap_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

// 2-D row-major [rows, Kp] (fp32, or fp16 when f16) with a (swb bytes x box_rows) box and the swb-byte swizzle (128 or 64)
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int Kp, int box_rows, bool f16, int swb) {
    auto enc = get_tensormap_encoder();
    KGE_REQUIRE(enc != nullptr, "kge_rank_counts: cuTensorMapEncodeTiled unavailable in this driver");
    const size_t esz = f16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * esz};
    cuuint32_t box[2] = {(cuuint32_t)(swb / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KGE_REQUIRE(r == CUDA_SUCCESS, "kge_rank_counts: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// KGE_RANK_SW=128|64: width in bytes of a pipeline stage's k-block (TcCfg); the 3xTF32 A/B path keeps 128
static int rank_tc_swb() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KGE_RANK_SW");
        v = (e != nullptr && atoi(e) == 64) ? 64 : ((e != nullptr && atoi(e) == 128) ? 128 : KGE_RANK_SW_DEFAULT);
    }
    return v;
}

// KGE_RANK_TF32=1: the round-1 3xTF32 operands (A/B); default: the fp16 split
bool kge_rank_tc_f16() {
    static int use_tf32 = -1;
    if (use_tf32 < 0) {
        const char* e = getenv("KGE_RANK_TF32");
        use_tf32 = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return use_tf32 == 0;
}

int kge_rank_sweep_tc(kge_ctx* ctx, int model, int K, const float* q, int64_t NQ, int64_t T, const float* ent_local,
                      int64_t row_begin, int64_t row_end, const int32_t* test, const int32_t* pos_q,
                      const int32_t* excl_lo, const int32_t* excl_hi, const int32_t* sp_ent, const int32_t* po_ent,
                      int side_mask, int32_t* counts, const uint32_t* q_absmax, cudaStream_t st) {
    (void)model;
    (void)NQ;
    const int64_t n_local = row_end - row_begin;
    if (n_local <= 0 || T <= 0) return 0;
    int64_t q_row0 = 0, M = 2 * T;
    if (side_mask == 1) M = T;
    if (side_mask == 2) {
        q_row0 = T;
        M = T;
    }
    const bool f16 = kge_rank_tc_f16();
    const size_t esz = f16 ? 2 : 4;
    const int swb = f16 ? rank_tc_swb() : 128;
    const int kb_elems = swb / (int)esz;  // columns per k-block: 64 / 32 fp16, 32 tf32
    const int Kp = ((K + kb_elems - 1) / kb_elems) * kb_elems;
    const int64_t Mp = ((M + TC_BM - 1) / TC_BM) * TC_BM;
    const int64_t Np = ((n_local + TC_BN - 1) / TC_BN) * TC_BN;
    KGE_REQUIRE(2 * Mp < (int64_t)INT32_MAX && 2 * Np < (int64_t)INT32_MAX, "kge_rank_counts: operand too large for the TMA coordinates");
    if (ctx->q_hi.reserve((size_t)2 * Mp * Kp * esz)) return -2;
    if (ctx->e_hi.reserve((size_t)2 * Np * Kp * esz)) return -2;
    if (ctx->q_lo.reserve(2 * sizeof(uint32_t))) return -2;  // the two absmax words
    uint32_t* absmax = ctx->q_lo.as<uint32_t>();
    {
        int threads = 256;
        const float* q0 = q + (size_t)q_row0 * K;
        if (f16) {
            // [0]: max|Q| -- already there when the caller folded the queries with q_absmax (same buffer); [1]: max|Ent|
            const bool vq = (K & 3) == 0 && ((uintptr_t)q0 & 15) == 0, ve = (K & 3) == 0 && ((uintptr_t)ent_local & 15) == 0;
            const int64_t nq = M * (int64_t)K, ne = n_local * (int64_t)K;
            const int64_t cap = (int64_t)ctx->sm_count * 8;
            if (q_absmax == nullptr) {
                KGE_CUDA_CHECK(cudaMemsetAsync(absmax, 0, 2 * sizeof(uint32_t), st));
                const int bq = (int)std::max<int64_t>(1, std::min(((vq ? nq >> 2 : nq) + threads * 4 - 1) / (threads * 4), cap));
                if (vq) kge_absmax_kernel<true><<<bq, threads, 0, st>>>(q0, nq, absmax);
                else kge_absmax_kernel<false><<<bq, threads, 0, st>>>(q0, nq, absmax);
            } else {
                KGE_REQUIRE(q_absmax == absmax, "kge_rank_sweep_tc: the query maximum must sit in the context's scale buffer");
            }
            const int be = (int)std::max<int64_t>(1, std::min(((ve ? ne >> 2 : ne) + threads * 4 - 1) / (threads * 4), cap));
            if (ve) kge_absmax_kernel<true><<<be, threads, 0, st>>>(ent_local, ne, absmax + 1);
            else kge_absmax_kernel<false><<<be, threads, 0, st>>>(ent_local, ne, absmax + 1);
            const int64_t tq8 = Mp * (int64_t)(Kp >> 3), te8 = Np * (int64_t)(Kp >> 3);
            const int bsq = (int)std::min<int64_t>((tq8 + threads - 1) / threads, (int64_t)ctx->sm_count * 32);
            const int bse = (int)std::min<int64_t>((te8 + threads - 1) / threads, (int64_t)ctx->sm_count * 32);
            if (vq) kge_f16_split_kernel<true><<<bsq, threads, 0, st>>>(q0, M, K, Kp, Mp, absmax, ctx->q_hi.as<__half>());
            else kge_f16_split_kernel<false><<<bsq, threads, 0, st>>>(q0, M, K, Kp, Mp, absmax, ctx->q_hi.as<__half>());
            if (ve) kge_f16_split_kernel<true><<<bse, threads, 0, st>>>(ent_local, n_local, K, Kp, Np, absmax + 1, ctx->e_hi.as<__half>());
            else kge_f16_split_kernel<false><<<bse, threads, 0, st>>>(ent_local, n_local, K, Kp, Np, absmax + 1, ctx->e_hi.as<__half>());
        } else {
            int64_t tot = Mp * Kp;
            int blocks = (int)std::min<int64_t>((tot + threads - 1) / threads, (int64_t)ctx->sm_count * 32);
            int64_t tot_e = Np * Kp;
            int blocks_e = (int)std::min<int64_t>((tot_e + threads - 1) / threads, (int64_t)ctx->sm_count * 32);
            kge_tf32_split_kernel<<<blocks, threads, 0, st>>>(q0, M, K, Kp, Mp, ctx->q_hi.as<float>());
            kge_tf32_split_kernel<<<blocks_e, threads, 0, st>>>(ent_local, n_local, K, Kp, Np, ctx->e_hi.as<float>());
        }
        KGE_CUDA_CHECK(cudaGetLastError());
    }
    CUtensorMap tmQ, tmE;
    if (int rc = make_tmap(&tmQ, ctx->q_hi.p, 2 * Mp, Kp, TC_BM, f16, swb)) return rc;
    if (int rc = make_tmap(&tmE, ctx->e_hi.p, 2 * Np, Kp, TC_BN, f16, swb)) return rc;

    TcParams P;
    P.k_blocks = Kp / kb_elems;
    {
        const int mma_k = 32 / (int)esz;  // columns per MMA k-step: 16 fp16 or 8 tf32
        P.tail_mmas = (K - (P.k_blocks - 1) * kb_elems + mma_k - 1) / mma_k;
    }
    P.absmax = absmax;
    P.M = M;
    P.Mp = Mp;
    P.Np = Np;
    P.T = T;
    P.q_row0 = q_row0;
    P.row_begin = row_begin;
    P.row_end = row_end;
    P.test = test;
    P.pos_q = pos_q;
    P.excl_lo = excl_lo;
    P.excl_hi = excl_hi;
    P.sp_ent = sp_ent;
    P.po_ent = po_ent;
    P.counts = counts;
    P.n_m_tiles = (int)(Mp / TC_BM);
    P.n_n_tiles = (int)(Np / TC_BN);
    P.nl = ctx->rank_nl;
    const int grid = (int)std::min<int64_t>((int64_t)P.n_m_tiles * P.n_n_tiles, (int64_t)ctx->sm_count);
    static int group = -1;  // KGE_RANK_GROUP=1|2|4
    if (group < 0) {
        const char* e = getenv("KGE_RANK_GROUP");
        const int g = e != nullptr ? atoi(e) : KGE_RANK_GROUP_DEFAULT;
        group = (g == 2 || g == 4) ? g : 1;
    }
    P.group = grid % group == 0 ? group : 1;
    static int ew = -1;  // KGE_RANK_EW=4|8 epilogue warps
    if (ew < 0) {
        const char* e = getenv("KGE_RANK_EW");
        ew = (e != nullptr && atoi(e) == 4) ? 4 : 8;
    }
    auto go = [&](auto kern, int smem, int threads) -> int {
        KGE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, threads, smem, st>>>(tmQ, tmE, P);
        return 0;
    };
    int rc;
    if (!f16) rc = go(kge_rank_tc_kernel<false, 128, 8>, TcCfg<128>::SMEM_BYTES, 64 + 32 * 8);
    else if (swb == 128) rc = ew == 4 ? go(kge_rank_tc_kernel<true, 128, 4>, TcCfg<128>::SMEM_BYTES, 64 + 32 * 4)
                                      : go(kge_rank_tc_kernel<true, 128, 8>, TcCfg<128>::SMEM_BYTES, 64 + 32 * 8);
    else rc = ew == 4 ? go(kge_rank_tc_kernel<true, 64, 4>, TcCfg<64>::SMEM_BYTES, 64 + 32 * 4)
                      : go(kge_rank_tc_kernel<true, 64, 8>, TcCfg<64>::SMEM_BYTES, 64 + 32 * 8);
    if (rc) return rc;
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_has_tensor_core_rank(void) { return 1; }
