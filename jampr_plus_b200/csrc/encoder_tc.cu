// bf16 tcgen05 path of the GNN encoders (placeholder until the tensor-core kernels land).
#include "common.cuh"

int jampr_node_encoder_tc(const jampr_dims_t*, const jampr_instances_t*, const jampr_weights_t*, cudaStream_t) {
  return JAMPR_E_UNSUPPORTED;
}
int jampr_tour_encoder_tc(const jampr_dims_t*, const jampr_instances_t*, const jampr_state_t*,
                          const jampr_weights_t*, int, cudaStream_t) {
  return JAMPR_E_UNSUPPORTED;
}
