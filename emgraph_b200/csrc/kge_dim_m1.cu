// explicit instantiation unit of the dimension-sharded phase kernels for scoring model 1 (parallel compilation)
#include "kge_dim.cuh"

int kge_launch_dim_m1(int phase, const DimParams& P, cudaStream_t st) { return launch_dim_model<1>(phase, P, st); }
