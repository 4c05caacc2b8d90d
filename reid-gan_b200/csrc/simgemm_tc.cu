// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
extern "C" {
int reid_knn_candidates_tc(const void*, int64_t, int64_t, int, int64_t, int64_t, int, int32_t*, float*, void*) {
  reid::set_error("reid_knn_candidates_tc: not built yet");
  return REID_ERR_UNSUPPORTED;
}
int reid_features_to_half(const float*, int64_t, int, void*, void*) {
  reid::set_error("reid_features_to_half: not built yet");
  return REID_ERR_UNSUPPORTED;
}
}
