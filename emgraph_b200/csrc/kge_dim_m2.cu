// explicit instantiation unit of the dimension-sharded phase kernels for scoring model 2 (parallel compilation)
#include "kge_dim.cuh"

int kge_launch_dim_m2(int phase, const DimParams& P, cudaStream_t st) { return launch_dim_model<2>(phase, P, st); }
int kge_launch_dim_sorted_m2(int phase, const DimParams& P, const DimChunks& C, uint2* pos_off, const uint64_t* ks, int64_t n_keys, cudaStream_t st) {
    return launch_dim_sorted_model<2>(phase, P, C, pos_off, ks, n_keys, st);
}
