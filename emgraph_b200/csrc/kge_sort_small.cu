// Stable counting sort of the (key << 32 | slot) entries of a SMALL batch over a SMALL key range (E + R <= 32 768 keys,
// <= 2^18 entries: FB15k-237 / WN18RR-sized graphs, BASELINE cfg1-3).
//
// Why not cub::DeviceRadixSort here (it stays for everything larger).  The sort runs on the side stream beside
// kge_fwd_bwd_kernel, and the segmented reduction waits for it.  For 98 000 entries CUB launches a histogram, a scan and two
// onesweep passes; alone they take 27 us, beside the forward/backward kernel 61 us (ncu launch list + the step's own events,
// profiles/r02_summary.md): a onesweep CTA is 384 threads x 94 registers = 36 K registers, an SM that holds three
// forward/backward CTAs has 16 K left, so every pass waits until an SM drains, and between two passes the freed room goes
// back to the other kernel.  cfg1/cfg2 (forward/backward 30 us) wait 16 us per step for the sort, cfg3 8 us.
//
// The keys have at most 15 bits, so one counting pass over the whole key does it -- with small CTAs that fit beside
// anything:
//   1. count    a warp owns a SEGMENT of KGE_SS_SEG consecutive entries (input order = slot order); H[segment][key] += 1
//               (fire-and-forget reductions; H is a dense uint32 matrix, <= 16 MB, all-zero between sorts)
//   2. scan     per key, the exclusive prefix of its counts over the segments (written back in place where the count is
//               non-zero) and the key's total; totals are scanned inside the CTA (32 keys) and the CTA sums by the last
//               CTA to finish (ticket), giving every key its first output position
//   3. scatter  each warp walks its segment again, 32 entries at a time: lanes with equal keys find each other with
//               match.any, the lowest one looks up how many entries of the key the segment has shown so far in a per-warp
//               shared-memory table (open addressing; no memory round trip between two steps), lane rank inside the group
//               keeps the input order: position = first(key) + prefix(segment, key) + seen(key) + rank.  Every load of the
//               walk is issued before it starts.  The warp then zeroes the H entries it touched, so H is all-zero again
//               without a memset.
// The three kernels are chains of ~2, ~3 and ~3 dependent memory round trips.
// Equal keys keep their input order (segment order, then order inside the segment): the output is bit-identical to the
// stable radix sort, and everything downstream (summation order of the reduction) is unchanged.
#include "kge_common.cuh"

#define KGE_SS_SEG 512          // entries per warp
#define KGE_SS_MAX_KEYS 32768
#define KGE_SS_MAX_ITEMS (1 << 18)
#define KGE_SS_MAX_H_BYTES ((size_t)16 << 20)
#define KGE_SS_SCAN_KEYS 32     // keys per CTA of the scan kernel
#define KGE_SS_SCAN_GROUPS 8    // segment groups per CTA of the scan kernel

__global__ void __launch_bounds__(256) kge_ss_count_kernel(const uint64_t* __restrict__ in, int n, int n_keys, uint32_t* __restrict__ H) {
    const int lane = threadIdx.x & 31;
    const int seg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int base = seg * KGE_SS_SEG;
    if (base >= n) return;
    uint32_t* row = H + (size_t)seg * n_keys;
#pragma unroll 4
    for (int it = 0; it < KGE_SS_SEG / 32; ++it) {
        const int idx = base + it * 32 + lane;
        if (idx < n) atomicAdd(row + (uint32_t)(in[idx] >> 32), 1u);
    }
}

// grid = ceil(n_keys / 32) CTAs of 256 threads: warp g of a CTA sums segment group g (of 8) for the CTA's 32 keys, lane =
// key.  All of a thread's loads are independent (a few batches of 16 in flight); the second walk over the same addresses,
// which writes the prefixes, is served by L1.
__global__ void __launch_bounds__(KGE_SS_SCAN_KEYS * KGE_SS_SCAN_GROUPS)
kge_ss_scan_kernel(uint32_t* __restrict__ H, int n_segs, int n_keys, uint32_t* __restrict__ key_first, uint32_t* __restrict__ cta_first,
                   uint32_t* __restrict__ cta_total, unsigned int* __restrict__ ticket) {
    __shared__ uint32_t part[KGE_SS_SCAN_GROUPS][KGE_SS_SCAN_KEYS];
    __shared__ bool last;
    const int kk = threadIdx.x % KGE_SS_SCAN_KEYS, g = threadIdx.x / KGE_SS_SCAN_KEYS;
    const int key = blockIdx.x * KGE_SS_SCAN_KEYS + kk;
    const int per = (n_segs + KGE_SS_SCAN_GROUPS - 1) / KGE_SS_SCAN_GROUPS;
    const int s0 = min(g * per, n_segs), s1 = min(s0 + per, n_segs);
    uint32_t sum = 0;
    if (key < n_keys) {
#pragma unroll 16
        for (int s = s0; s < s1; ++s) sum += H[(size_t)s * n_keys + key];
    }
    part[g][kk] = sum;
    __syncthreads();
    // exclusive prefix of this key over the segments, in place (entries the count pass never touched stay zero)
    uint32_t run = 0;
    for (int q = 0; q < g; ++q) run += part[q][kk];
    if (key < n_keys && sum != 0) {
#pragma unroll 16
        for (int s = s0; s < s1; ++s) {
            const uint32_t c = H[(size_t)s * n_keys + key];
            if (c != 0) H[(size_t)s * n_keys + key] = run;
            run += c;
        }
    }
    // first output position of every key: scan of the totals of the CTA's 32 keys (warp 0); the CTA's own offset is added
    // by the scatter
    last = false;
    __syncthreads();
    if (g == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int q = 0; q < KGE_SS_SCAN_GROUPS; ++q) tot += part[q][kk];
        uint32_t incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (kk >= d) incl += up;
        }
        if (key < n_keys) key_first[key] = incl - tot;
        if (kk == KGE_SS_SCAN_KEYS - 1) {
            cta_total[blockIdx.x] = incl;
            __threadfence();
            last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        // the last CTA to finish: exclusive scan of the CTA totals (<= 1024 of them), one warp
        __threadfence();
        const int lane = threadIdx.x;
        uint32_t carry = 0;
        for (int b0 = 0; b0 < (int)gridDim.x; b0 += 32) {
            const int b = b0 + lane;
            const uint32_t v = b < (int)gridDim.x ? ((volatile uint32_t*)cta_total)[b] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            if (b < (int)gridDim.x) cta_first[b] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) *ticket = 0u;  // ready for the next sort
    }
}

// per-warp table of (key + 1) << 16 | count for the keys of one segment: open addressing, linear probing; only the lowest
// lane of every group of equal keys touches it, groups of one step hold different keys
#define KGE_SS_TBL 1024
__device__ __forceinline__ uint32_t ss_take(uint32_t* tbl, uint32_t key, uint32_t cnt) {
    const uint32_t tag = (key + 1u) << 16;
    uint32_t h = (key * 0x9E3779B1u) >> 22;  // 10 bits
    for (;;) {
        uint32_t cur = tbl[h];
        if (cur == 0u) cur = atomicCAS(tbl + h, 0u, tag), cur = cur == 0u ? tag : cur;
        if ((cur & 0xffff0000u) == tag) break;
        h = (h + 1u) & (KGE_SS_TBL - 1u);
    }
    return atomicAdd(tbl + h, cnt) & 0xffffu;  // entries of this key seen in earlier steps of the segment (<= 512)
}

#define KGE_SS_SCATTER_WARPS 4  // 128 threads x 80 registers, 16 KB of tables: fits beside three forward/backward CTAs
__global__ void __launch_bounds__(KGE_SS_SCATTER_WARPS * 32) kge_ss_scatter_kernel(const uint64_t* __restrict__ in, int n, int n_keys, uint32_t* __restrict__ H,
                                                             const uint32_t* __restrict__ key_first, const uint32_t* __restrict__ cta_first,
                                                             uint64_t* __restrict__ out) {
    __shared__ uint32_t tbl_all[KGE_SS_SCATTER_WARPS][KGE_SS_TBL];
    constexpr int NIT = KGE_SS_SEG / 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int seg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int base = seg * KGE_SS_SEG;
    if (base >= n) return;  // whole warps leave together; no CTA-wide barrier below
    uint32_t* tbl = tbl_all[wib];
#pragma unroll
    for (int i = 0; i < KGE_SS_TBL / 32; ++i) tbl[lane + 32 * i] = 0u;
    uint32_t* row = H + (size_t)seg * n_keys;
    const unsigned lt = (1u << lane) - 1u;
    // everything that comes from memory is independent of the ranks: all of it in flight before the walk
    uint64_t e[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int idx = base + it * 32 + lane;
        e[it] = idx < n ? in[idx] : 0ull;
    }
    uint32_t pos[NIT];  // first position of the key + entries of the key in earlier segments
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int idx = base + it * 32 + lane;
        const uint32_t key = (uint32_t)(e[it] >> 32);
        pos[it] = idx < n ? cta_first[key / KGE_SS_SCAN_KEYS] + key_first[key] + row[key] : 0u;
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int idx = base + it * 32 + lane;
        const bool valid = idx < n;
        const uint32_t key = (uint32_t)(e[it] >> 32);
        // lanes past the end get keys of their own, so they never join a group
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? key : (0x80000000u | (uint32_t)lane));
        const int leader = __ffs(peers) - 1;
        uint32_t taken = 0;
        if (valid && lane == leader) taken = ss_take(tbl, key, (uint32_t)__popc(peers));
        taken = __shfl_sync(0xffffffffu, taken, leader);
        if (valid) out[pos[it] + taken + (uint32_t)__popc(peers & lt)] = e[it];
        __syncwarp();  // the table updates of this step are visible to the next one
    }
    // leave H all-zero: the entries this segment touched (every read of them is behind us)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int idx = base + it * 32 + lane;
        if (idx < n) row[(uint32_t)(e[it] >> 32)] = 0u;
    }
}

bool kge_small_sort_ok(int64_t n_items, int64_t n_keys) {
    static int on = -1;  // KGE_SMALL_SORT=0 keeps the radix sort (A/B)
    if (on < 0) {
        const char* e = getenv("KGE_SMALL_SORT");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || n_items <= 0 || n_items > KGE_SS_MAX_ITEMS || n_keys <= 0 || n_keys > KGE_SS_MAX_KEYS) return false;
    const int64_t n_segs = (n_items + KGE_SS_SEG - 1) / KGE_SS_SEG;
    return (size_t)n_segs * (size_t)n_keys * sizeof(uint32_t) <= KGE_SS_MAX_H_BYTES;
}

// in: n_items entries, every key (high word) < n_keys; out: the same entries ordered by key, equal keys in input order
int kge_small_sort(kge_ctx* ctx, const uint64_t* in, int64_t n_items, int64_t n_keys, uint64_t* out, cudaStream_t st) {
    KGE_REQUIRE(kge_small_sort_ok(n_items, n_keys), "kge_small_sort: unsupported size (%lld entries, %lld keys)", (long long)n_items,
                (long long)n_keys);
    const int n = (int)n_items, nk = (int)n_keys;
    const int n_segs = (n + KGE_SS_SEG - 1) / KGE_SS_SEG;
    const int n_ctas_scan = (nk + KGE_SS_SCAN_KEYS - 1) / KGE_SS_SCAN_KEYS;
    const size_t h_bytes = (size_t)n_segs * nk * sizeof(uint32_t);
    const size_t aux_bytes = ((size_t)nk + 2 * (size_t)n_ctas_scan + 64) * sizeof(uint32_t);
    // H is all-zero between sorts (the scatter pass clears what the count pass set); a fresh allocation is zeroed once
    const void* h_before = ctx->ss_hist.p;
    if (ctx->ss_hist.reserve(h_bytes)) return -2;
    if (ctx->ss_hist.p != h_before) KGE_CUDA_CHECK(cudaMemsetAsync(ctx->ss_hist.p, 0, ctx->ss_hist.cap, st));
    const void* a_before = ctx->ss_aux.p;
    if (ctx->ss_aux.reserve(aux_bytes)) return -2;
    if (ctx->ss_aux.p != a_before) KGE_CUDA_CHECK(cudaMemsetAsync(ctx->ss_aux.p, 0, ctx->ss_aux.cap, st));
    uint32_t* H = ctx->ss_hist.as<uint32_t>();
    // the ticket sits at a FIXED place (word 0): every sort leaves it zero, and the arrays behind it move with the key count
    unsigned int* ticket = ctx->ss_aux.as<unsigned int>();
    uint32_t* key_first = ctx->ss_aux.as<uint32_t>() + 32;
    uint32_t* cta_first = key_first + nk;
    uint32_t* cta_total = cta_first + n_ctas_scan;
    const int warps_per_cta = 8;
    const unsigned seg_ctas = (unsigned)((n_segs + warps_per_cta - 1) / warps_per_cta);
    kge_ss_count_kernel<<<seg_ctas, warps_per_cta * 32, 0, st>>>(in, n, nk, H);
    kge_ss_scan_kernel<<<n_ctas_scan, KGE_SS_SCAN_KEYS * KGE_SS_SCAN_GROUPS, 0, st>>>(H, n_segs, nk, key_first, cta_first, cta_total, ticket);
    kge_ss_scatter_kernel<<<(unsigned)((n_segs + KGE_SS_SCATTER_WARPS - 1) / KGE_SS_SCATTER_WARPS), KGE_SS_SCATTER_WARPS * 32, 0, st>>>(
        in, n, nk, H, key_first, cta_first, out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}
