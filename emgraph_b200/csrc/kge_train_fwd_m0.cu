// Instantiation of the fused forward/backward kernel for scoring model 0 (see kge_train_fwd.cuh).
#include "kge_train_fwd.cuh"

int kge_launch_fwd_bwd_m0(const FwdBwdParams& P, int sm_count, cudaStream_t st) { return launch_fwd_bwd_model<0>(P, sm_count, st); }
