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

// ---- member lists: one pass over the labels instead of one pass PER CLUSTER -------------------------------------
// centroids_kernel above lets every CTA stream the whole label vector (C x N x 8 bytes of L2 reads: 271 MB at the
// benchmark shape, 16 GB at N = 250k) and adds its members one dependent row at a time.  Here the labels are binned
// once (count -> scan -> scatter), a CTA sorts its own short member list in shared memory (ascending index: the
// reference's enumerate(labels) order, so the fp32 sums are reproducible) and adds the member rows four at a time.
constexpr int kCenListCap = 4096;

__global__ void __launch_bounds__(256) cen_count_kernel(const int64_t* __restrict__ labels, int64_t N, int64_t cap,
                                                        int32_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t l = labels[i];
  if (l >= 0 && l < cap) atomicAdd(&cnt[l], 1);
}

__global__ void __launch_bounds__(256) cen_scatter_kernel(const int64_t* __restrict__ labels, int64_t N, int64_t cap,
                                                          const int64_t* __restrict__ ptr, int32_t* __restrict__ cursor,
                                                          int32_t* __restrict__ members) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t l = labels[i];
  if (l >= 0 && l < cap) members[ptr[l] + atomicAdd(&cursor[l], 1)] = (int32_t)i;
}

__global__ void __launch_bounds__(kCenThreads) centroids_list_kernel(const float* __restrict__ x, int64_t N, int64_t D,
                                                                     const int64_t* __restrict__ labels,
                                                                     const int64_t* __restrict__ ptr,
                                                                     const int32_t* __restrict__ members, int normalize,
                                                                     float* __restrict__ out,
                                                                     const int64_t* __restrict__ n_clusters_dev) {
  __shared__ int32_t s_mem[kCenListCap];
  __shared__ float s_red[kCenThreads / 32];
  const int64_t k = blockIdx.x;
  if (n_clusters_dev && k >= *n_clusters_dev) return;      // the grid was sized for the capacity, not the count
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int64_t a = ptr[k];
  const int n = (int)(ptr[k + 1] - a);
  float acc[kCenMaxPerThread];
#pragma unroll
  for (int u = 0; u < kCenMaxPerThread; ++u) acc[u] = 0.f;
  if (n <= kCenListCap) {
    int n2 = 32;
    while (n2 < n) n2 <<= 1;
    for (int e = t; e < n2; e += kCenThreads) s_mem[e] = e < n ? members[a + e] : 0x7fffffff;
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {                  // bitonic sort, ascending
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int e = t; e < n2; e += kCenThreads) {
          const int p = e ^ j;
          if (p > e) {
            const int32_t v0 = s_mem[e], v1 = s_mem[p];
            if ((v0 > v1) == ((e & kk) == 0)) {
              s_mem[e] = v1;
              s_mem[p] = v0;
            }
          }
        }
        __syncthreads();
      }
    }
    // rows four at a time: the loads of four members are in flight together, the adds stay in member order
    for (int m0 = 0; m0 < n; m0 += 4) {
      float r[4][kCenMaxPerThread];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* row = x + (int64_t)s_mem[m0 + q < n ? m0 + q : m0] * D;
#pragma unroll
        for (int u = 0; u < kCenMaxPerThread; ++u) {
          const int64_t d = t + (int64_t)u * kCenThreads;
          r[q][u] = d < D ? row[d] : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (m0 + q < n) {
#pragma unroll
          for (int u = 0; u < kCenMaxPerThread; ++u) acc[u] = __fadd_rn(acc[u], r[q][u]);
        }
      }
    }
  } else {
    // a cluster too large for the shared-memory list (thousands of members): stream the label vector like
    // centroids_kernel does -- ascending by construction
    for (int64_t i = 0; i < N; ++i) {
      if (labels[i] != k) continue;                          // uniform across the CTA
      const float* row = x + i * D;
#pragma unroll
      for (int u = 0; u < kCenMaxPerThread; ++u) {
        const int64_t d = t + (int64_t)u * kCenThreads;
        if (d < D) acc[u] = __fadd_rn(acc[u], row[d]);
      }
    }
  }
  const float cnt_f = (float)n;
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < kCenMaxPerThread; ++u) {
    const int64_t d = t + (int64_t)u * kCenThreads;
    if (d < D) {
      acc[u] = n ? __fdiv_rn(acc[u], cnt_f) : 0.f;
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

struct CenWs {
  int32_t* cnt;      // cap
  int32_t* cursor;   // cap
  int64_t* ptr;      // cap + 1
  int32_t* members;  // N
};
static size_t cen_carve(void* base, int64_t N, int64_t cap, CenWs* w) {
  unsigned char* p = (unsigned char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += (bytes + 255) / 256 * 256;
    return r;
  };
  CenWs t;
  t.cnt = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
  t.cursor = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
  t.ptr = (int64_t*)take(sizeof(int64_t) * (size_t)(cap + 1));
  t.members = (int32_t*)take(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
  if (w) *w = t;
  return off;
}

static int centroids_lists(const float* x, int64_t N, int64_t D, const int64_t* labels, int64_t cap,
                           const int64_t* n_clusters_dev, int normalize, float* out, void* workspace, cudaStream_t st) {
  CenWs w;
  cen_carve(workspace, N, cap, &w);
  REID_CUDA(cudaMemsetAsync(w.cnt, 0, ((sizeof(int32_t) * (size_t)cap + 255) / 256 * 256) * 2, st));   // cnt + cursor
  if (N > 0) {
    cen_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(labels, N, cap, w.cnt);
    REID_LAUNCH_CHECK();
  }
  int rc = reid_scan_counts(w.cnt, cap, w.ptr, nullptr, (void*)st);
  if (rc != REID_OK) return rc;
  if (N > 0) {
    cen_scatter_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(labels, N, cap, w.ptr, w.cursor, w.members);
    REID_LAUNCH_CHECK();
  }
  centroids_list_kernel<<<(unsigned)cap, kCenThreads, 0, st>>>(x, N, D, labels, w.ptr, w.members, normalize, out,
                                                              n_clusters_dev);
  REID_LAUNCH_CHECK();
  return REID_OK;
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
  if (N < 0 || C < 0) return 0;
  return reid::cen_carve(nullptr, N, C > 0 ? C : 1, nullptr);
}

int reid_centroids(const float* x, int64_t N, int64_t D, const int64_t* labels, int64_t C, int normalize, float* out,
                   void* workspace, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && labels && (out || C == 0), "reid_centroids: NULL pointer");
  REID_CHECK_ARG(N >= 0 && C >= 0 && D > 0 && D <= (int64_t)kCenThreads * kCenMaxPerThread,
                 "reid_centroids: bad shape N=%lld C=%lld D=%lld (D <= %d)", (long long)N, (long long)C, (long long)D,
                 kCenThreads * kCenMaxPerThread);
  if (C == 0) return REID_OK;
  if (workspace)                                      // reid_centroids_workspace_bytes(N, C): member lists, one label pass
    return centroids_lists(x, N, D, labels, C, nullptr, normalize, out, workspace, (cudaStream_t)stream);
  centroids_kernel<<<(unsigned)C, kCenThreads, 0, (cudaStream_t)stream>>>(x, N, D, labels, normalize, out, nullptr);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_centroids_dev(const float* x, int64_t N, int64_t D, const int64_t* labels, const int64_t* num_clusters_dev,
                       int64_t capacity, int normalize, float* out, void* workspace, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && labels && num_clusters_dev && out, "reid_centroids_dev: NULL pointer");
  REID_CHECK_ARG(N >= 0 && capacity >= 1 && capacity < (1ll << 31) && D > 0 && D <= (int64_t)kCenThreads * kCenMaxPerThread,
                 "reid_centroids_dev: bad shape N=%lld capacity=%lld D=%lld (D <= %d)", (long long)N, (long long)capacity,
                 (long long)D, kCenThreads * kCenMaxPerThread);
  if (workspace)
    return centroids_lists(x, N, D, labels, capacity, num_clusters_dev, normalize, out, workspace, (cudaStream_t)stream);
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
