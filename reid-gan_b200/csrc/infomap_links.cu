// f1 -- edge filter of the Infomap clustering variant:
// clustercontrast/utils/infomap_cluster.py get_links :129-144.  For every row i the neighbour list (ascending
// distance 1 - sim) is walked until the first entry with dist > 1 - min_sim; self entries are skipped; every
// visited entry becomes a link (i, nbr) with weight 1 - dist; rows without a link are "single".
// Two passes (count -> caller scans -> fill), one thread per row (k is a few dozen).
#include "common.cuh"

namespace reid {

__device__ __forceinline__ int links_walk(const int32_t* __restrict__ nbr, const float* __restrict__ dist, int k, int64_t i,
                                          double thr, int32_t* __restrict__ dst, double* __restrict__ w) {
  int count = 0;
  for (int j = 0; j < k; ++j) {
    const int32_t n = nbr[j];
    if ((int64_t)n == i) continue;                       // :134-135
    const double d = (double)dist[j];
    if (!(d <= thr)) break;                              // :136 / :140 (sorted: stop at the first failure)
    if (dst) {
      dst[count] = n;
      w[count] = 1.0 - d;                                // :138 float(1 - dists[i][j])
    }
    ++count;
  }
  return count;
}

__global__ void __launch_bounds__(256) links_count_kernel(const int32_t* __restrict__ nbrs, const float* __restrict__ dists,
                                                          int64_t N, int k, double thr, int32_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  cnt[i] = links_walk(nbrs + i * k, dists + i * k, k, i, thr, nullptr, nullptr);
}

__global__ void __launch_bounds__(256) links_fill_kernel(const int32_t* __restrict__ nbrs, const float* __restrict__ dists,
                                                         int64_t N, int k, double thr, const int64_t* __restrict__ ptr,
                                                         int32_t* __restrict__ dst, double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  links_walk(nbrs + i * k, dists + i * k, k, i, thr, dst + ptr[i], w + ptr[i]);
}

}  // namespace reid

extern "C" {

int reid_links_count(const int32_t* nbrs, const float* dists, int64_t N, int k, double min_sim, int32_t* link_cnt,
                     void* stream) {
  using namespace reid;
  REID_CHECK_ARG(nbrs && dists && link_cnt && N >= 0 && k >= 1, "reid_links_count: bad arguments");
  if (N == 0) return REID_OK;
  links_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nbrs, dists, N, k, 1.0 - min_sim, link_cnt);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_links_fill(const int32_t* nbrs, const float* dists, int64_t N, int k, double min_sim, const int64_t* link_ptr,
                    int32_t* link_dst, double* link_weight, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(nbrs && dists && link_ptr && link_dst && link_weight && N >= 0 && k >= 1, "reid_links_fill: bad arguments");
  if (N == 0) return REID_OK;
  links_fill_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nbrs, dists, N, k, 1.0 - min_sim, link_ptr,
                                                                                  link_dst, link_weight);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
