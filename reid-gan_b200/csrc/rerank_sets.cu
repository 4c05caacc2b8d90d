// a2 + a3 -- integer / bitset kernels: k-reciprocal masks and the 2/3-overlap expansion
// (utils/faiss_rerank.py:23-27, 65-69, 72-80).  One warp per query row; neighbour lists are
// staged in shared memory, set membership is decided with warp ballots, the union is
// de-duplicated in a per-warp shared-memory hash set and sorted with a warp bitonic network.
#include "common.cuh"

namespace reid {

// mask bit r <=> row in rank[rank[row, r], :cols]
__global__ void __launch_bounds__(256) reciprocal_kernel(const int32_t* __restrict__ rank, int ncols, int cols,
                                                         int64_t row_begin, int64_t row_end,
                                                         uint64_t* __restrict__ mask_out) {
  const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  const int lane = lane_id();
  uint64_t mask = 0;
  for (int base = 0; base < cols; base += 32) {
    const int r = base + lane;
    bool found = false;
    if (r < cols) {
      const int32_t* nb = rank + (int64_t)rank[row * ncols + r] * ncols;
      for (int q = 0; q < cols; ++q) found |= (nb[q] == (int32_t)row);
    }
    mask |= (uint64_t)__ballot_sync(kFull, found) << base;
  }
  if (lane == 0) mask_out[row - row_begin] = mask;
}

constexpr int kSetSlots = 2048;   // > 2 * (64 + 64*33) is not needed in practice; overflow is checked
constexpr int kListCap = 1024;    // max |E| handled (theoretical max for k1=30 is 30 + 30*16 = 510)
constexpr int kExpandWarps = 4;

__device__ __forceinline__ uint32_t hash32(uint32_t v) {
  v ^= v >> 16;
  v *= 0x7feb352du;
  v ^= v >> 15;
  v *= 0x846ca68bu;
  v ^= v >> 16;
  return v;
}

// ascending bitonic sort of n2 (power of two) ints held in shared memory, by one warp
__device__ __forceinline__ void warp_bitonic_sort(int32_t* a, int n2) {
  const int lane = lane_id();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < n2; t += 32) {
        int p = t ^ j;
        if (p > t) {
          int32_t x = a[t], y = a[p];
          bool up = ((t & k) == 0);
          if ((x > y) == up) {
            a[t] = y;
            a[p] = x;
          }
        }
      }
      __syncwarp();
    }
  }
}

struct ExpandSmem {
  int32_t rlist[64];
  int32_t set[kSetSlots];
  int32_t list[kListCap];
};

__global__ void __launch_bounds__(kExpandWarps * 32) expand_kernel(
    const int32_t* __restrict__ rank, int ncols, const uint64_t* __restrict__ Rmask,
    const uint64_t* __restrict__ Rhmask, int64_t row_begin, int64_t row_end, int stride,
    int32_t* __restrict__ E_pad, int32_t* __restrict__ E_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ExpandSmem& sm = reinterpret_cast<ExpandSmem*>(smem_raw)[threadIdx.x >> 5];
  const int64_t row = row_begin + (int64_t)blockIdx.x * kExpandWarps + (threadIdx.x >> 5);
  if (row >= row_end) return;
  const int lane = lane_id();

  for (int s = lane; s < kSetSlots; s += 32) sm.set[s] = -1;
  // R(row) in rank order
  const uint64_t rm = Rmask[row - row_begin];
  int nR = 0;
  for (int base = 0; base < ncols; base += 32) {
    const int r = base + lane;
    const bool in = r < ncols && ((rm >> r) & 1ull);
    const unsigned b = __ballot_sync(kFull, in);
    if (in) sm.rlist[nR + __popc(b & ((1u << lane) - 1u))] = rank[row * ncols + r];
    nR += __popc(b);
  }
  __syncwarp();

  bool lost = false;  // the hash set filled up: reported through an impossible count
  auto insert = [&](int32_t g) {
    uint32_t h = hash32((uint32_t)g) & (kSetSlots - 1);
    for (int probe = 0; probe < kSetSlots; ++probe) {
      int32_t old = atomicCAS(&sm.set[h], -1, g);
      if (old == -1 || old == g) return;
      h = (h + 1) & (kSetSlots - 1);
    }
    lost = true;
  };

  // lane <-> candidate c = R(row)[lane (+32)]
  for (int ci = lane; ci < nR; ci += 32) {
    const int32_t c = sm.rlist[ci];
    insert(c);
    const uint64_t hm = Rhmask[c];
    const int m = __popcll(hm);
    const int32_t* crow = rank + (int64_t)c * ncols;
    int cnt = 0;
    for (uint64_t bits = hm; bits; bits &= bits - 1) {
      const int32_t g = crow[__ffsll((long long)bits) - 1];
      bool in = false;
      for (int u = 0; u < nR; ++u) in |= (sm.rlist[u] == g);
      cnt += in;
    }
    if (3 * cnt > 2 * m) {  // == len(intersect1d) > 2/3*len  (faiss_rerank.py:77)
      for (uint64_t bits = hm; bits; bits &= bits - 1) insert(crow[__ffsll((long long)bits) - 1]);
    }
  }
  __syncwarp();

  // compact the set into list[], count
  int nE = 0;
  for (int base = 0; base < kSetSlots; base += 32) {
    const int32_t v = sm.set[base + lane];
    const unsigned b = __ballot_sync(kFull, v >= 0);
    if (v >= 0) {
      int p = nE + __popc(b & ((1u << lane) - 1u));
      if (p < kListCap) sm.list[p] = v;
    }
    nE += __popc(b);
  }
  const int cap = stride < kListCap ? stride : kListCap;
  const bool overflow = __any_sync(kFull, lost) || nE > cap;
  if (nE > cap) nE = cap;
  int n2 = 32;
  while (n2 < nE) n2 <<= 1;
  for (int t = nE + lane; t < n2; t += 32) sm.list[t] = 0x7fffffff;
  __syncwarp();
  warp_bitonic_sort(sm.list, n2);
  int32_t* dst = E_pad + (row - row_begin) * (int64_t)stride;
  for (int t = lane; t < nE; t += 32) dst[t] = sm.list[t];
  if (lane == 0) E_cnt[row - row_begin] = overflow ? stride + 1 : nE;  // stride + 1 = "did not fit"
}

}  // namespace reid

extern "C" {

int reid_reciprocal_masks(const int32_t* rank, int64_t N, int ncols, int k, int64_t row_begin, int64_t row_end,
                          uint64_t* mask_out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rank && mask_out, "reid_reciprocal_masks: NULL pointer");
  REID_CHECK_ARG(ncols >= 1 && ncols <= REID_MAX_K1, "reid_reciprocal_masks: ncols=%d not in 1..%d", ncols, REID_MAX_K1);
  REID_CHECK_ARG(k >= 0, "reid_reciprocal_masks: k=%d", k);
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_reciprocal_masks: bad row range");
  const int cols = k + 1 < ncols ? k + 1 : ncols;  // rank[i, :k+1] clamps to the stored columns
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  reciprocal_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(rank, ncols, cols, row_begin, row_end,
                                                                              mask_out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_expand(const int32_t* rank, int64_t N, int ncols, const uint64_t* Rmask, const uint64_t* Rhalf_mask,
                int64_t row_begin, int64_t row_end, int stride, int32_t* E_pad, int32_t* E_cnt, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rank && Rmask && Rhalf_mask && E_pad && E_cnt, "reid_expand: NULL pointer");
  REID_CHECK_ARG(ncols >= 1 && ncols <= REID_MAX_K1, "reid_expand: ncols=%d not in 1..%d", ncols, REID_MAX_K1);
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_expand: bad row range");
  REID_CHECK_ARG(stride >= 1, "reid_expand: stride=%d", stride);
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  const size_t smem = sizeof(ExpandSmem) * kExpandWarps;
  const unsigned grid = (unsigned)((n + kExpandWarps - 1) / kExpandWarps);
  REID_CUDA(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  expand_kernel<<<grid, kExpandWarps * 32, smem, (cudaStream_t)stream>>>(rank, ncols, Rmask, Rhalf_mask, row_begin,
                                                                         row_end, stride, E_pad, E_cnt);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
