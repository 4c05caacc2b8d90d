// a9 -- centroid initialisation (examples/cluster_contrast_train_usl.py:169-182, 191):
//   centre_k = mean of x[i] over labels[i] == k (members visited in ascending i, like the
//   reference's enumerate(labels) loop feeding torch.stack(...).mean(0)), then F.normalize.
// One CTA per cluster; the label vector (N ints, L2-resident) is streamed once per CTA and the
// member rows are accumulated in registers, so every feature row is read exactly once overall.
#include "common.cuh"

namespace reid {

constexpr int kCenThreads = 256;
constexpr int kCenMaxPerThread = 16;  // D <= 4096

__global__ void __launch_bounds__(kCenThreads) centroids_kernel(const float* __restrict__ x, int64_t N, int64_t D,
                                                                const int64_t* __restrict__ labels, int normalize,
                                                                float* __restrict__ out,
                                                                const int64_t* __restrict__ n_clusters_dev) {
  __shared__ unsigned s_mask[kCenThreads / 32];
  __shared__ float s_red[kCenThreads / 32];
  const int64_t k = blockIdx.x;
  if (n_clusters_dev && k >= *n_clusters_dev) return;      // the grid was sized for the capacity, not the count
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  float acc[kCenMaxPerThread];
#pragma unroll
  for (int u = 0; u < kCenMaxPerThread; ++u) acc[u] = 0.f;
  int count = 0;
  for (int64_t base = 0; base < N; base += kCenThreads) {
    const int64_t i = base + t;
    const bool hit = i < N && labels[i] == k;
    const unsigned b = __ballot_sync(kFull, hit);
    if (lane == 0) s_mask[w] = b;
    __syncthreads();
#pragma unroll 1
    for (int ww = 0; ww < kCenThreads / 32; ++ww) {
      unsigned m = s_mask[ww];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const float* row = x + (base + ww * 32 + bit) * D;
#pragma unroll
        for (int u = 0; u < kCenMaxPerThread; ++u) {
          const int64_t d = t + (int64_t)u * kCenThreads;
          if (d < D) acc[u] = __fadd_rn(acc[u], row[d]);
        }
        ++count;
      }
    }
    __syncthreads();
  }
  const float inv_cnt = (float)count;
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < kCenMaxPerThread; ++u) {
    const int64_t d = t + (int64_t)u * kCenThreads;
    if (d < D) {
      acc[u] = count ? __fdiv_rn(acc[u], inv_cnt) : 0.f;
      ss += acc[u] * acc[u];
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[w] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int ww = 0; ww < kCenThreads / 32; ++ww) tot += s_red[ww];
  const float nrm = normalize ? fmaxf(sqrtf(tot), 1e-12f) : 1.0f;  // F.normalize eps
#pragma unroll
  for (int u = 0; u < kCenMaxPerThread; ++u) {
    const int64_t d = t + (int64_t)u * kCenThreads;
    if (d < D) out[k * D + d] = __fdiv_rn(acc[u], nrm);
  }
}

// dst[r] = src[idx[r]]: the device-side form of `torch.cat([features[f].unsqueeze(0) for f, _, _ in sorted(train)], 0)`
// (examples/cluster_contrast_train_usl.py:153) over a feature store that never left the GPU.
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const int64_t* __restrict__ idx,
                                                          int64_t n, int64_t row_f4, float4* __restrict__ dst) {
  const int64_t r = blockIdx.x;
  if (r >= n) return;
  const float4* s = src + idx[r] * row_f4;
  float4* d = dst + r * row_f4;
  for (int64_t t = threadIdx.x; t < row_f4; t += blockDim.x) d[t] = s[t];
}

}  // namespace reid

extern "C" {

size_t reid_centroids_workspace_bytes(int64_t N, int64_t C) {
  (void)N;
  (void)C;
  return 0;
}

int reid_centroids(const float* x, int64_t N, int64_t D, const int64_t* labels, int64_t C, int normalize, float* out,
                   void* workspace, void* stream) {
  using namespace reid;
  (void)workspace;
  REID_CHECK_ARG(x && labels && (out || C == 0), "reid_centroids: NULL pointer");
  REID_CHECK_ARG(N >= 0 && C >= 0 && D > 0 && D <= (int64_t)kCenThreads * kCenMaxPerThread,
                 "reid_centroids: bad shape N=%lld C=%lld D=%lld (D <= %d)", (long long)N, (long long)C, (long long)D,
                 kCenThreads * kCenMaxPerThread);
  if (C == 0) return REID_OK;
  centroids_kernel<<<(unsigned)C, kCenThreads, 0, (cudaStream_t)stream>>>(x, N, D, labels, normalize, out, nullptr);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_centroids_dev(const float* x, int64_t N, int64_t D, const int64_t* labels, const int64_t* num_clusters_dev,
                       int64_t capacity, int normalize, float* out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && labels && num_clusters_dev && out, "reid_centroids_dev: NULL pointer");
  REID_CHECK_ARG(N >= 0 && capacity >= 1 && capacity < (1ll << 31) && D > 0 && D <= (int64_t)kCenThreads * kCenMaxPerThread,
                 "reid_centroids_dev: bad shape N=%lld capacity=%lld D=%lld (D <= %d)", (long long)N, (long long)capacity,
                 (long long)D, kCenThreads * kCenMaxPerThread);
  centroids_kernel<<<(unsigned)capacity, kCenThreads, 0, (cudaStream_t)stream>>>(x, N, D, labels, normalize, out,
                                                                                num_clusters_dev);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_gather_rows(const float* src, int64_t n_src, const int64_t* idx, int64_t n, int64_t D, float* dst, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(src && idx && dst && n >= 0 && n_src >= 0 && D > 0 && D % 4 == 0, "reid_gather_rows: bad arguments (D %% 4 == 0)");
  REID_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "reid_gather_rows: buffers must be 16-byte aligned");
  if (n == 0) return REID_OK;
  gather_rows_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>((const float4*)src, idx, n, D / 4, (float4*)dst);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
