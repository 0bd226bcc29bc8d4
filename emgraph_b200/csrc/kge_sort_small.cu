// Stable sort of the (key << 32 | slot) entries of a SMALL batch over a SMALL key range (<= 2^18 entries, keys < 2^16:
// FB15k-237 / WN18RR-sized graphs, BASELINE cfg1-3) in ONE launch.
//
// Why not cub::DeviceRadixSort here (it stays for everything larger).  The sort runs on the side stream beside
// kge_fwd_bwd_kernel, and the segmented reduction waits for it.  For 98 000 entries CUB launches a histogram, a scan and two
// onesweep passes; alone they take 27 us, beside the forward/backward kernel 61 us (ncu launch list + the step's own events,
// profiles/r02_summary.md): a onesweep CTA is 384 threads x 96 registers = 36 K registers, an SM that holds three
// forward/backward CTAs has 16 K left, so every pass waits until an SM drains, and between two passes the freed room goes
// back to the other kernel.  cfg1/cfg2 (forward/backward 30 us) wait 20 us per step for the sort, cfg3 12 us.
// (A first attempt -- one counting pass over the whole key with a dense segment x key count matrix -- was correct and
// slower still: 98 000 random read-modify-writes into an 11 MB matrix and a scan over all of it, 65 us alone; session r2t.)
//
// This kernel is a two-pass LSD radix sort (8-bit digits) whose CTAs are small -- 256 threads, <= 64 registers, 10 KB of
// shared memory: they fit beside three forward/backward CTAs on every SM -- and stay resident from the first load to the
// last store: a tile of 2 048 entries per CTA, at most 128 CTAs, grid-wide barriers between the phases.  Per pass:
//   rank     warp w walks its 256 consecutive entries 32 at a time; lanes with equal digits find each other with match.any,
//            the lowest one bumps the warp's private counter of that digit, lane rank inside the group keeps the input
//            order.  The counters of the 8 warps are then scanned per digit (warp order = input order) and their sums
//            published as the tile's digit histogram G[tile][digit].
//   barrier  every tile's histogram is visible.
//   place    thread d adds up G[tiles before mine][d] and G[all tiles][d], the totals are scanned over the digits inside the
//            CTA: first position of (digit, tile).  Each entry goes to first(digit, tile) + entries of the digit in earlier
//            warps + its rank in the warp.
// Equal digits keep their input order in both passes, so equal keys keep their input order: the output is bit-identical
// to the stable radix sort it replaces, and the summation order of the reduction is unchanged.
#include "kge_common.cuh"

#define KGE_SS_THREADS 256
#define KGE_SS_WARPS 8
#define KGE_SS_PER_WARP 256                           // entries per warp
#define KGE_SS_IT (KGE_SS_PER_WARP / 32)              // steps of the walk
#define KGE_SS_TILE (KGE_SS_WARPS * KGE_SS_PER_WARP)  // entries per CTA
#define KGE_SS_MAX_TILES 128
#define KGE_SS_MAX_KEYS 65536

// grid-wide barrier for CTAs that are all resident (<= 128 small CTAs on 148 SMs).  bar[0] counts arrivals and never goes
// back; bar[1] holds the value it had when this launch began (the last arrival of a launch's final barrier stores it for
// the next launch).  Barrier k of a launch is complete when the counter reaches base + k * gridDim.x: an arrival is one
// fire-and-forget reduction, a waiter sees the last one a single L2 round trip later.
__device__ __forceinline__ void ss_grid_barrier(unsigned int* bar, unsigned int target, bool final_one) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (final_one) {
            if (atomicAdd(bar, 1u) + 1u == target) bar[1] = target;
        } else {
            atomicAdd(bar, 1u);
        }
        unsigned long long spins = 0;
        while ((int)(((volatile unsigned int*)bar)[0] - target) < 0) {
            if (++spins > (1ull << 28)) __trap();  // seconds: a CTA of this grid never ran -- fail loudly instead of hanging
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ uint64_t ss_ldcg(const uint64_t* p) {
    return (uint64_t)__ldcg(reinterpret_cast<const unsigned long long*>(p));
}

// one pass: entries of this tile from src (stable) to dst by digit (key >> shift) & 255
__device__ __forceinline__ void ss_pass(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n, int shift, bool src_is_input,
                                        uint32_t* __restrict__ G, unsigned int* bar, unsigned int target, bool final_one,
                                        uint32_t (*cntw)[256], uint32_t* first, uint32_t* wsum) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_tiles = gridDim.x, tile = blockIdx.x;
    const int base = tile * KGE_SS_TILE + w * KGE_SS_PER_WARP;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < KGE_SS_WARPS * 256; i += KGE_SS_THREADS) (&cntw[0][0])[i] = 0u;
    uint64_t e[KGE_SS_IT];
#pragma unroll
    for (int it = 0; it < KGE_SS_IT; ++it) {
        const int idx = base + it * 32 + lane;
        // pass 2 reads what other CTAs wrote during this launch: straight from L2
        e[it] = idx < n ? (src_is_input ? src[idx] : ss_ldcg(src + idx)) : 0ull;
    }
    __syncthreads();
    // ---- rank: position of every entry among the entries of its digit inside this warp's 256
    uint32_t rank[KGE_SS_IT];
#pragma unroll
    for (int it = 0; it < KGE_SS_IT; ++it) {
        const bool valid = base + it * 32 + lane < n;
        const uint32_t d = (uint32_t)(e[it] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (256u + (uint32_t)lane));  // idle lanes stay alone
        const int leader = __ffs(peers) - 1;
        uint32_t seen = 0;
        if (valid && lane == leader) {
            seen = cntw[w][d];
            cntw[w][d] = seen + (uint32_t)__popc(peers);
        }
        seen = __shfl_sync(0xffffffffu, seen, leader);
        rank[it] = seen + (uint32_t)__popc(peers & lt);
        __syncwarp();  // the counter updates of this step are visible to the next one
    }
    __syncthreads();
    // ---- per digit: exclusive prefix over the warps (in place), the tile's count published
    {
        const int d = threadIdx.x;  // 256 threads, 256 digits
        uint32_t run = 0;
#pragma unroll
        for (int q = 0; q < KGE_SS_WARPS; ++q) {
            const uint32_t c = cntw[q][d];
            cntw[q][d] = run;
            run += c;
        }
        __stcg(G + (size_t)tile * 256 + d, run);  // G[tile][digit]: written and read 128 bytes per warp
    }
    ss_grid_barrier(bar, target, final_one);
    // ---- place: first output position of (digit, this tile)
    {
        const int d = threadIdx.x;
        uint32_t below = 0, total = 0;
#pragma unroll 16
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t c = __ldcg(G + (size_t)t * 256 + d);
            total += c;
            below += t < tile ? c : 0u;
        }
        // exclusive scan of `total` over the 256 digits
        uint32_t incl = total;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += up;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t off = 0;
        for (int q = 0; q < w; ++q) off += wsum[q];
        first[d] = off + incl - total + below;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < KGE_SS_IT; ++it) {
        if (base + it * 32 + lane < n) {
            const uint32_t d = (uint32_t)(e[it] >> shift) & 255u;
            __stcg(reinterpret_cast<unsigned long long*>(dst) + (first[d] + cntw[w][d] + rank[it]), (unsigned long long)e[it]);
        }
    }
}

// <= 64 registers: 16 K per CTA, the room three forward/backward CTAs leave on an SM
__global__ void __launch_bounds__(KGE_SS_THREADS, 4) kge_small_sort_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ tmp,
                                                                         uint64_t* __restrict__ out, int n, uint32_t* __restrict__ G,
                                                                         unsigned int* bar) {
    __shared__ uint32_t cntw[KGE_SS_WARPS][256];
    __shared__ uint32_t first[256];
    __shared__ uint32_t wsum[KGE_SS_WARPS];
    __shared__ unsigned int base0;
    if (threadIdx.x == 0) base0 = ((volatile unsigned int*)bar)[1];  // stable until this launch's final barrier completes
    __syncthreads();
    const unsigned int base = base0, nt = gridDim.x;
    ss_pass(in, tmp, n, 32, true, G, bar, base + nt, false, cntw, first, wsum);
    ss_grid_barrier(bar, base + 2u * nt, false);  // every tile's entries are in tmp
    ss_pass(tmp, out, n, 40, false, G + (size_t)256 * nt, bar, base + 3u * nt, true, cntw, first, wsum);
}

// The grid barriers need every CTA of the launch resident at the same time: the grid is capped by what the device can hold
// (148 SMs x 4 CTAs on a full B200 -- far above KGE_SS_MAX_TILES -- but a partitioned device may be much smaller)
static int ss_max_resident_ctas() {
    static int v = -1;
    if (v < 0) {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kge_small_sort_kernel, KGE_SS_THREADS, 0) != cudaSuccess)
            v = 0;
        else
            v = sms * per_sm;
    }
    return v;
}

bool kge_small_sort_ok(int64_t n_items, int64_t n_keys) {
    static int on = -1;  // KGE_SMALL_SORT=0 keeps the radix sort (A/B)
    if (on < 0) {
        const char* e = getenv("KGE_SMALL_SORT");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || n_items <= 0 || n_items > (int64_t)KGE_SS_MAX_TILES * KGE_SS_TILE || n_keys <= 0 || n_keys > KGE_SS_MAX_KEYS) return false;
    return (n_items + KGE_SS_TILE - 1) / KGE_SS_TILE <= (int64_t)ss_max_resident_ctas();
}

// in: n_items entries, every key (high word) < n_keys <= 65536; out: the same entries ordered by key, equal keys in input order
int kge_small_sort(kge_ctx* ctx, const uint64_t* in, int64_t n_items, int64_t n_keys, uint64_t* out, cudaStream_t st) {
    KGE_REQUIRE(kge_small_sort_ok(n_items, n_keys), "kge_small_sort: unsupported size (%lld entries, %lld keys)", (long long)n_items,
                (long long)n_keys);
    KGE_REQUIRE(in != out, "kge_small_sort: in-place sort is not supported");
    const int n = (int)n_items;
    const int n_tiles = (n + KGE_SS_TILE - 1) / KGE_SS_TILE;
    // aux: [0,2) barrier {arrivals, value at launch} (zeroed once) | 64 words in: G[2 passes][n_tiles][256]
    const size_t aux_bytes = (64 + (size_t)2 * 256 * KGE_SS_MAX_TILES) * sizeof(uint32_t);
    const void* a_before = ctx->ss_aux.p;
    if (ctx->ss_aux.reserve(aux_bytes)) return -2;
    if (ctx->ss_aux.p != a_before) KGE_CUDA_CHECK(cudaMemsetAsync(ctx->ss_aux.p, 0, ctx->ss_aux.cap, st));
    if (ctx->ss_hist.reserve((size_t)n * sizeof(uint64_t))) return -2;  // the pass-1 output
    unsigned int* bar = ctx->ss_aux.as<unsigned int>();
    uint32_t* G = ctx->ss_aux.as<uint32_t>() + 64;
    kge_small_sort_kernel<<<n_tiles, KGE_SS_THREADS, 0, st>>>(in, ctx->ss_hist.as<uint64_t>(), out, n, G, bar);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}
