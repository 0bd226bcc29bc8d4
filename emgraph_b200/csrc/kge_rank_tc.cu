// tcgen05 3xTF32 ranking sweep (placeholder until the tensor-core path lands).
#include "kge_common.cuh"

int kge_rank_sweep_tc(kge_ctx* ctx, int model, int K, const float* q, int64_t NQ, int64_t T, const float* ent_local,
                      int64_t row_begin, int64_t row_end, const int32_t* test, const int32_t* pos_q,
                      const int32_t* excl_lo, const int32_t* excl_hi, const int32_t* sp_ent, const int32_t* po_ent,
                      int side_mask, int32_t* counts, cudaStream_t st) {
    kge_set_error("kge_rank_counts: tensor-core sweep not built in this version");
    return -4;
}

extern "C" int kge_has_tensor_core_rank(void) { return 0; }
