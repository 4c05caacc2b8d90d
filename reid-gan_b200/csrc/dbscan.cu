// a8 -- DBSCAN on precomputed distances, on device
// (examples/cluster_contrast_train_usl.py:160,163; sklearn/cluster/_dbscan.py:397-475,
// _dbscan_inner.pyx).  sklearn's sequential DFS is restated order-free:
//   core(i)  <=> |{j : d_ij <= eps}| >= min_samples          (self included, fp32 compare)
//   clusters  =  connected components of the core-core eps-graph, found with a lock-free
//                union-find that always hooks the larger root under the smaller one, so a
//                component's root is its smallest core index;
//   cluster id = rank of that root among all roots (dbscan_inner opens labels while scanning
//                i ascending); border point -> smallest id among adjacent cores; else -1.
#include "common.cuh"

namespace reid {

// ---- dense input: eps-neighbourhood lists of rows of an N x N matrix -------------------
__global__ void __launch_bounds__(256) dense_count_kernel(const float* __restrict__ dist, int64_t N, int64_t ld,
                                                          float eps, int64_t row_begin, int64_t row_end,
                                                          int32_t* __restrict__ cnt) {
  const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  const float* d = dist + row * ld;
  int c = 0;
  for (int64_t j = lane_id(); j < N; j += 32) c += d[j] <= eps;
  c = warp_sum(c);
  if (lane_id() == 0) cnt[row - row_begin] = c;
}

__global__ void __launch_bounds__(256) dense_fill_kernel(const float* __restrict__ dist, int64_t N, int64_t ld,
                                                         float eps, int64_t row_begin, int64_t row_end,
                                                         const int64_t* __restrict__ ptr, int32_t* __restrict__ idx) {
  const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  const float* d = dist + row * ld;
  int64_t o = ptr[row - row_begin];
  const int lane = lane_id();
  for (int64_t base = 0; base < N; base += 32) {
    const int64_t j = base + lane;
    const bool in = j < N && d[j] <= eps;
    const unsigned b = __ballot_sync(kFull, in);
    if (in) idx[o + __popc(b & ((1u << lane) - 1u))] = (int32_t)j;
    o += __popc(b);
  }
}

// ---- labelling ---------------------------------------------------------------------------
__global__ void init_kernel(int64_t N, const int32_t* __restrict__ cnt, int min_samples, int32_t* __restrict__ parent,
                            uint8_t* __restrict__ core) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  parent[i] = (int32_t)i;
  core[i] = cnt[i] >= min_samples;
}

__device__ __forceinline__ int32_t uf_find(int32_t* parent, int32_t a) {
  const volatile int32_t* vp = parent;  // other warps hook roots concurrently
  int32_t p = vp[a];
  while (p != a) {
    const int32_t g = vp[p];
    if (g != p) atomicMin(&parent[a], g);  // path halving (monotone: parents only decrease)
    a = p;
    p = g;
  }
  return a;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int32_t a, int32_t b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int32_t t = a;
      a = b;
      b = t;
    }
    // a > b: hook root a under b
    const int32_t old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

__global__ void __launch_bounds__(256) union_kernel(int64_t N, const int64_t* __restrict__ ptr,
                                                    const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt,
                                                    const uint8_t* __restrict__ core, int32_t* __restrict__ parent) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N || !core[i]) return;
  const int64_t a = ptr[i];
  const int n = cnt[i];
  for (int e = lane_id(); e < n; e += 32) {
    const int32_t j = idx[a + e];
    if (j != (int32_t)i && core[j]) uf_union(parent, (int32_t)i, j);
  }
}

// ---- owned-pair lists (csrc/jaccard.cu pair_owned): every edge {i, j} is listed by ONE of its rows ----------------
// degree = own entries (the self pair (i, i) is i's own entry when J_ii <= eps) + one per foreign list that names i
__global__ void __launch_bounds__(256) degree_kernel(int64_t N, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                     const int32_t* __restrict__ cnt, int32_t* __restrict__ deg) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int64_t a = ptr[i];
  const int n = cnt[i];
  for (int e = lane_id(); e < n; e += 32) {
    const int32_t j = idx[a + e];
    if (j != (int32_t)i) atomicAdd(&deg[j], 1);
  }
  if (lane_id() == 0 && n) atomicAdd(&deg[i], n);
}

// cores take their component's number, everything else "no label yet" (INT64_MAX)
__global__ void label_core_kernel(int64_t N, const uint8_t* __restrict__ core, const int32_t* __restrict__ parent,
                                  const int64_t* __restrict__ root_rank, int64_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  labels[i] = core[i] ? root_rank[parent[i]] : INT64_MAX;
}

// border points: smallest label among adjacent cores -- an edge is seen once, so it reports in both directions
__global__ void __launch_bounds__(256) label_border_kernel(int64_t N, const int64_t* __restrict__ ptr,
                                                           const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt,
                                                           const uint8_t* __restrict__ core,
                                                           const int32_t* __restrict__ parent,
                                                           const int64_t* __restrict__ root_rank, int64_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int64_t a = ptr[i];
  const int n = cnt[i];
  const bool ci = core[i];
  const long long li = ci ? (long long)root_rank[parent[i]] : 0;
  for (int e = lane_id(); e < n; e += 32) {
    const int32_t j = idx[a + e];
    if (j == (int32_t)i) continue;
    const bool cj = core[j];
    if (cj && !ci) atomicMin(reinterpret_cast<long long*>(&labels[i]), (long long)root_rank[parent[j]]);
    if (ci && !cj) atomicMin(reinterpret_cast<long long*>(&labels[j]), li);
  }
}

__global__ void label_noise_kernel(int64_t N, int64_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && labels[i] == INT64_MAX) labels[i] = -1;
}

__global__ void flatten_kernel(int64_t N, const uint8_t* __restrict__ core, int32_t* __restrict__ parent,
                               int32_t* __restrict__ is_root) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int32_t r = 0;
  if (core[i]) {
    int32_t a = (int32_t)i;
    while (parent[a] != a) a = parent[a];
    parent[i] = a;  // every chain only shrinks towards its (final) root: safe without atomics
    r = (a == (int32_t)i);
  }
  is_root[i] = r;
}

__global__ void __launch_bounds__(256) label_kernel(int64_t N, const int64_t* __restrict__ ptr,
                                                    const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt,
                                                    const uint8_t* __restrict__ core,
                                                    const int32_t* __restrict__ parent,
                                                    const int64_t* __restrict__ root_rank,
                                                    int64_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int lane = lane_id();
  if (core[i]) {
    if (lane == 0) labels[i] = root_rank[parent[i]];
    return;
  }
  const int64_t a = ptr[i];
  const int n = cnt[i];
  int64_t best = INT64_MAX;
  for (int e = lane; e < n; e += 32) {
    const int32_t j = idx[a + e];
    if (core[j]) {
      const int64_t l = root_rank[parent[j]];
      best = l < best ? l : best;
    }
  }
  best = warp_min(best);
  if (lane == 0) labels[i] = best == INT64_MAX ? -1 : best;
}

struct DbscanWs {
  int32_t* parent;
  int32_t* is_root;
  int64_t* root_rank;  // N + 1
  uint8_t* core;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline DbscanWs carve(void* ws, int64_t N) {
  DbscanWs w;
  unsigned char* p = (unsigned char*)ws;
  w.root_rank = (int64_t*)p;
  p += align_up(sizeof(int64_t) * (size_t)(N + 1), 256);
  w.parent = (int32_t*)p;
  p += align_up(sizeof(int32_t) * (size_t)N, 256);
  w.is_root = (int32_t*)p;
  p += align_up(sizeof(int32_t) * (size_t)N, 256);
  w.core = (uint8_t*)p;
  return w;
}

}  // namespace reid

extern "C" {

int reid_dbscan_dense_count(const float* dist, int64_t N, int64_t ld, float eps, int64_t row_begin, int64_t row_end,
                            int32_t* nbr_cnt, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(dist && nbr_cnt, "reid_dbscan_dense_count: NULL pointer");
  REID_CHECK_ARG(N > 0 && ld >= N && 0 <= row_begin && row_begin <= row_end, "reid_dbscan_dense_count: bad shape");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  dense_count_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dist, N, ld, eps, row_begin, row_end,
                                                                               nbr_cnt);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_dbscan_dense_fill(const float* dist, int64_t N, int64_t ld, float eps, int64_t row_begin, int64_t row_end,
                           const int64_t* nbr_ptr, int32_t* nbr_idx, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(dist && nbr_ptr && nbr_idx, "reid_dbscan_dense_fill: NULL pointer");
  REID_CHECK_ARG(N > 0 && ld >= N && 0 <= row_begin && row_begin <= row_end, "reid_dbscan_dense_fill: bad shape");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  dense_fill_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dist, N, ld, eps, row_begin, row_end,
                                                                              nbr_ptr, nbr_idx);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

size_t reid_dbscan_workspace_bytes(int64_t N) {
  using namespace reid;
  if (N < 0) N = 0;
  return align_up(sizeof(int64_t) * (size_t)(N + 1), 256) + 2 * align_up(sizeof(int32_t) * (size_t)N, 256) +
         align_up((size_t)N, 256);
}

int reid_dbscan_labels(int64_t N, const int64_t* nbr_ptr, const int32_t* nbr_idx, const int32_t* nbr_cnt,
                       int min_samples, int64_t* labels, uint8_t* core_mask, int64_t* num_clusters_out,
                       void* workspace, int owned_pairs, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(nbr_ptr && nbr_cnt && labels && workspace, "reid_dbscan_labels: NULL pointer");
  REID_CHECK_ARG(N >= 0 && N < (1ll << 31), "reid_dbscan_labels: bad N");
  if (N == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DbscanWs w = carve(workspace, N);
  const unsigned gt = (unsigned)((N + 255) / 256), gw = (unsigned)((N + 7) / 8);
  const int32_t* degree = nbr_cnt;
  if (owned_pairs) {                                   // every edge is listed once: degrees first (is_root is free until flatten)
    REID_CUDA(cudaMemsetAsync(w.is_root, 0, sizeof(int32_t) * (size_t)N, st));
    degree_kernel<<<gw, 256, 0, st>>>(N, nbr_ptr, nbr_idx, nbr_cnt, w.is_root);
    REID_LAUNCH_CHECK();
    degree = w.is_root;
  }
  init_kernel<<<gt, 256, 0, st>>>(N, degree, min_samples, w.parent, w.core);
  REID_LAUNCH_CHECK();
  union_kernel<<<gw, 256, 0, st>>>(N, nbr_ptr, nbr_idx, nbr_cnt, w.core, w.parent);
  REID_LAUNCH_CHECK();
  flatten_kernel<<<gt, 256, 0, st>>>(N, w.core, w.parent, w.is_root);
  REID_LAUNCH_CHECK();
  int rc = reid_scan_counts(w.is_root, N, w.root_rank, nullptr, stream);
  if (rc != REID_OK) return rc;
  if (owned_pairs) {
    label_core_kernel<<<gt, 256, 0, st>>>(N, w.core, w.parent, w.root_rank, labels);
    REID_LAUNCH_CHECK();
    label_border_kernel<<<gw, 256, 0, st>>>(N, nbr_ptr, nbr_idx, nbr_cnt, w.core, w.parent, w.root_rank, labels);
    REID_LAUNCH_CHECK();
    label_noise_kernel<<<gt, 256, 0, st>>>(N, labels);
    REID_LAUNCH_CHECK();
  } else {
    label_kernel<<<gw, 256, 0, st>>>(N, nbr_ptr, nbr_idx, nbr_cnt, w.core, w.parent, w.root_rank, labels);
    REID_LAUNCH_CHECK();
  }
  if (core_mask) REID_CUDA(cudaMemcpyAsync(core_mask, w.core, (size_t)N, cudaMemcpyDeviceToDevice, st));
  if (num_clusters_out)
    REID_CUDA(cudaMemcpyAsync(num_clusters_out, w.root_rank + N, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return REID_OK;
}
}
