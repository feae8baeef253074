// Tensor-core (tcgen05 / TMEM / TMA) kNN for feature-space layers -- placeholder entry
// points so the ABI is stable; the kernel lands in a later commit of this round.
#include "common.cuh"

extern "C" int ecb200_split_tf32(const float*, int, int, int, float*, float*, float*, void*) {
  ecb200::set_error("ecb200_split_tf32: tensor-core kNN is not part of this build");
  return ECB200_ERR_ARG;
}
extern "C" int ecb200_knn_tc(const float*, const float*, const float*, int, int, int, int, int32_t*,
                             void*) {
  ecb200::set_error("ecb200_knn_tc: tensor-core kNN is not part of this build");
  return ECB200_ERR_ARG;
}
