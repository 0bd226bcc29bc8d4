// Instantiation of the fused forward/backward kernel for scoring model 3 (see kge_train_fwd.cuh).
#include "kge_train_fwd.cuh"

int kge_launch_fwd_bwd_m3(const FwdBwdParams& P, int sm_count, cudaStream_t st) { return launch_fwd_bwd_model<3>(P, sm_count, st); }
