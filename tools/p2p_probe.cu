// Peer-memory probe (2 GPUs, one process): random 1 KiB row gathers over NVLink as loads (reader pulls)
// versus stores (owner pushes), for several table sizes.  Build: nvcc -O3 -arch=sm_100a tools/p2p_probe.cu -o gpurun_out/p2p_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// warp per output slot: dst[slot] = src[idx(slot)], rows of K floats (K = 256), U rows in flight per warp
template <int U>
__global__ void gather_rows(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n_slots, uint32_t rows, uint32_t seed) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s0 = warp * U; s0 < n_slots; s0 += nwarps * U) {
        float4 a[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t s = s0 + u;
            const uint32_t r = (uint32_t)(((uint64_t)hash32((uint32_t)s * 2654435761u + seed) * rows) >> 32);
            const float4* p = src + (size_t)r * 64;
            a[u][0] = p[lane];
            a[u][1] = p[lane + 32];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t s = s0 + u;
            if (s < n_slots) {
                float4* q = dst + (size_t)s * 64;
                q[lane] = a[u][0];
                q[lane + 32] = a[u][1];
            }
        }
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
    const int64_t n_slots = 600000;  // 600k rows of 1 KiB = 614 MB moved
    for (double gb : {0.03, 0.25, 2.2}) {
        const uint32_t rows = (uint32_t)(gb * 1e9 / 1024);
        float4 *t0, *t1, *d0, *d1;
        CK(cudaSetDevice(0)); CK(cudaMalloc(&t0, (size_t)rows * 1024)); CK(cudaMalloc(&d0, n_slots * 1024)); CK(cudaMemset(t0, 1, (size_t)rows * 1024));
        CK(cudaSetDevice(1)); CK(cudaMalloc(&t1, (size_t)rows * 1024)); CK(cudaMalloc(&d1, n_slots * 1024)); CK(cudaMemset(t1, 1, (size_t)rows * 1024));
        CK(cudaDeviceSynchronize());
        struct Case { const char* name; int dev; float4* src; float4* dst; };
        Case cases[] = {{"local gather (dev0: t0 -> d0)", 0, t0, d0},
                        {"pull: peer loads (dev0: t1 -> d0)", 0, t1, d0},
                        {"push: peer stores (dev1: t1 -> d0)", 1, t1, d0}};
        for (auto& c : cases) {
            CK(cudaSetDevice(c.dev));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int blocks_per_sm : {4, 8}) {
                for (int U : {2, 4, 8}) {
                    float best = 1e9f;
                    for (int it = 0; it < 4; ++it) {
                        CK(cudaEventRecord(e0));
                        const int grid = 148 * blocks_per_sm;
                        if (U == 2) gather_rows<2><<<grid, 256>>>(c.src, c.dst, n_slots, rows, it);
                        if (U == 4) gather_rows<4><<<grid, 256>>>(c.src, c.dst, n_slots, rows, it);
                        if (U == 8) gather_rows<8><<<grid, 256>>>(c.src, c.dst, n_slots, rows, it);
                        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (it > 0 && ms < best) best = ms;
                    }
                    printf("table %.2f GB  %-36s ctas/sm %d U %d : %.3f ms  %.0f GB/s\n", gb, c.name, blocks_per_sm, U, best, n_slots * 1024.0 / best / 1e6);
                }
            }
        }
        CK(cudaSetDevice(0)); cudaFree(t0); cudaFree(d0);
        CK(cudaSetDevice(1)); cudaFree(t1); cudaFree(d1);
    }
    return 0;
}
