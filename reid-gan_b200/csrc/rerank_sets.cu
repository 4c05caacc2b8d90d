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

// Both masks of a row in one pass over its neighbours' lists: R uses the first `cols_full` columns of every list it
// looks into, R_half the first `cols_half` -- the same lists, read once.  Rows outside [row_begin, row_end) only get
// their R_half bit mask (every rank of a row-sharded pass needs R_half of ALL rows, R only of its own).
__global__ void __launch_bounds__(256) reciprocal2_kernel(const int32_t* __restrict__ rank, int64_t N, int ncols, int cols_full,
                                                          int cols_half, int64_t row_begin, int64_t row_end,
                                                          uint64_t* __restrict__ R_out, uint64_t* __restrict__ Rh_out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = lane_id();
  const bool own = row >= row_begin && row < row_end;
  const int cols = own ? cols_full : cols_half;
  uint64_t mf = 0, mh = 0;
  for (int base = 0; base < cols; base += 32) {
    const int r = base + lane;
    bool in_full = false, in_half = false;
    if (r < cols) {
      const int32_t* nb = rank + (int64_t)rank[row * ncols + r] * ncols;
      for (int q = 0; q < cols; ++q) {
        const bool hit = nb[q] == (int32_t)row;
        in_full |= hit;
        in_half |= hit && q < cols_half;
      }
    }
    mf |= (uint64_t)__ballot_sync(kFull, in_full) << base;
    mh |= (uint64_t)__ballot_sync(kFull, in_half && r < cols_half) << base;
  }
  if (lane == 0) {
    if (own) R_out[row - row_begin] = mf;
    Rh_out[row] = mh;
  }
}

constexpr int kListCap = 1024;    // max |E| handled (theoretical max for k1=30 is 30 + 30*16 = 510)
constexpr int kExpandWarps = 8;

__device__ __forceinline__ uint32_t hash32(uint32_t v) {
  v ^= v >> 16;
  v *= 0x7feb352du;
  v ^= v >> 15;
  v *= 0x846ca68bu;
  v ^= v >> 16;
  return v;
}

// position of the (n + 1)-th set bit of m (n < popc(m)): five popc steps instead of the software loop behind __fns
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const int c = __popc((m >> pos) & ((1u << w) - 1u));
    if (n >= c) {
      n -= c;
      pos += w;
    }
  }
  return pos;
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

// Per-warp shared memory, carved by the host for the call's (ncols, half_cols):
//   pairs_cap = ncols * half_cols (rounded up to 32): every (candidate c in R(i), member g of R_half(c)) pair
//   set_slots = power of two >= ncols + pairs_cap: de-duplication table of E(i).  It can never fill up (|E| <= ncols +
//               pairs), and real rows leave it mostly empty (|E| ~ 2 k1 of 512 slots for k1 = 30); the table used to be
//               twice as big, which cost a quarter of the kernel's occupancy for rows that do not exist
struct ExpandLayout {
  int pairs_cap, set_slots, bytes;
};
__host__ __device__ inline ExpandLayout expand_layout(int ncols, int half_cols) {
  ExpandLayout l;
  l.pairs_cap = (ncols * half_cols + 31) / 32 * 32;
  int s = 64;
  while (s < ncols + l.pairs_cap) s <<= 1;
  l.set_slots = s;
  // rlist[64] hr[128] off[65 -> 68] m[64] cnt[64] pass[64] hm[64 x u64] | g[pairs_cap] (later: list) | ci[pairs_cap x u8] | set[]
  const int list_ints = l.pairs_cap + 64;                 // |E| <= ncols + pairs
  l.bytes = (64 + 128 + 68 + 64 + 64 + 64) * 4 + 64 * 8 + list_ints * 4 + l.pairs_cap + l.set_slots * 4;
  l.bytes = (l.bytes + 15) / 16 * 16;
  return l;
}

// One warp per row.  The 2/3-overlap test of faiss_rerank.py:74-78 is evaluated pair-parallel: the (c, g) pairs
// -- c a member of R(i), g a member of R_half(c) -- are spread over the lanes (load-balanced over the scanned
// sizes |R_half(c)|), each lane gathers its g = rank[c, pos] and looks it up in a small hash table of R(i);
// the hits are counted per c with shared-memory atomics.  Candidates that pass contribute their R_half to the
// de-duplication table, which is then compacted and sorted (np.unique order).
__global__ void __launch_bounds__(kExpandWarps * 32) expand_kernel(
    const int32_t* __restrict__ rank, int ncols, int half_cols, const uint64_t* __restrict__ Rmask,
    const uint64_t* __restrict__ Rhmask, int64_t row_begin, int64_t row_end, int stride,
    int32_t* __restrict__ E_pad, int32_t* __restrict__ E_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ExpandLayout L = expand_layout(ncols, half_cols);
  unsigned char* base = smem_raw + (size_t)(threadIdx.x >> 5) * L.bytes;
  uint64_t* s_hm = reinterpret_cast<uint64_t*>(base);                 // [64]
  int32_t* rlist = reinterpret_cast<int32_t*>(base + 64 * 8);         // [64]
  int32_t* hr = rlist + 64;                                           // [128]
  int32_t* s_off = hr + 128;                                          // [68]
  int32_t* s_m = s_off + 68;                                          // [64]
  int32_t* s_cnt = s_m + 64;                                          // [64]
  int32_t* s_pass = s_cnt + 64;                                       // [64]
  int32_t* s_g = s_pass + 64;                                         // [pairs_cap (+64)]  pairs, later the sorted list
  const int list_ints = L.pairs_cap + 64;
  uint8_t* s_ci = reinterpret_cast<uint8_t*>(s_g + list_ints);        // [pairs_cap]
  int32_t* set = reinterpret_cast<int32_t*>(s_ci + L.pairs_cap);      // [set_slots]
  const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;

  for (int s = lane; s < 128; s += 32) hr[s] = -1;
  for (int s = lane; s < 64; s += 32) s_cnt[s] = 0;
  // R(row) in rank order
  const uint64_t rm = Rmask[row - row_begin];
  int nR = 0;
  for (int b0 = 0; b0 < ncols; b0 += 32) {
    const int r = b0 + lane;
    const bool in = r < ncols && ((rm >> r) & 1ull);
    const unsigned b = __ballot_sync(kFull, in);
    if (in) rlist[nR + __popc(b & lt)] = rank[row * ncols + r];
    nR += __popc(b);
  }
  __syncwarp();

  bool lost = false;  // a table filled up / a mask was wider than announced: reported through an impossible count
  // returns true when g was not in the table yet
  auto insert = [&](int32_t* tab, int slots, int32_t g) {
    uint32_t h = hash32((uint32_t)g) & (uint32_t)(slots - 1);
    for (int probe = 0; probe < slots; ++probe) {
      const int32_t old = atomicCAS(&tab[h], -1, g);
      if (old == -1) return true;
      if (old == g) return false;
      h = (h + 1) & (uint32_t)(slots - 1);
    }
    lost = true;
    return false;
  };
  // hash table of R(row), sizes |R_half(c)| and their scan
  int carry = 0;
  for (int b0 = 0; b0 < nR; b0 += 32) {
    const int ci = b0 + lane;
    int m = 0;
    if (ci < nR) {
      const int32_t c = rlist[ci];
      insert(hr, 128, c);
      const uint64_t hm = Rhmask[c];
      s_hm[ci] = hm;
      m = __popcll(hm);
      s_m[ci] = m;
    }
    int inc = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    if (ci < nR) s_off[ci] = carry + inc - m;
    carry += __shfl_sync(kFull, inc, 31);
  }
  const int P = carry;
  if (lane == 0) s_off[nR] = P;
  __syncwarp();
  if (P > L.pairs_cap) {                         // R_half masks wider than half_cols: cannot happen for a consistent call
    if (lane == 0) E_cnt[row - row_begin] = stride + 1;
    return;
  }
  // pairs: gather g, test membership in R(row)
  int ci = 0;                                    // the lane's pairs come in ascending order: the owner only moves forward
  for (int p = lane; p < P; p += 32) {
    while (s_off[ci + 1] <= p) ++ci;             // largest ci with s_off[ci] <= p  (s_off[nR] == P > p)
    const int kth = p - s_off[ci];
    const uint64_t hm = s_hm[ci];
    const uint32_t mlo = (uint32_t)hm, mhi = (uint32_t)(hm >> 32);
    const int nlo = __popc(mlo);
    const int pos = kth < nlo ? nth_set_bit(mlo, kth) : 32 + nth_set_bit(mhi, kth - nlo);
    const int32_t g = rank[(int64_t)rlist[ci] * ncols + pos];
    s_g[p] = g;
    s_ci[p] = (uint8_t)ci;
    uint32_t h = hash32((uint32_t)g) & 127u;
    bool in = false;
    for (int probe = 0; probe < 128; ++probe) {
      const int32_t o = hr[h];
      if (o == g) {
        in = true;
        break;
      }
      if (o == -1) break;
      h = (h + 1) & 127u;
    }
    if (in) atomicAdd(&s_cnt[ci], 1);
  }
  __syncwarp();
  // The de-duplication table is carved for the worst case (every candidate passes: k1 + k1 (h + 1) entries), but a row
  // only uses what its passing candidates can contribute: |R| + sum of the passing |R_half(c)| entries.  Clearing and
  // sweeping the table are a large part of this kernel's instructions, so only a power of two >= twice that is used.
  int contrib = 0;
  for (int ci = lane; ci < nR; ci += 32) {
    const int pass = 3 * s_cnt[ci] > 2 * s_m[ci];    // == len(intersect1d) > 2/3*len  (faiss_rerank.py:77)
    s_pass[ci] = pass;
    contrib += 1 + (pass ? s_m[ci] : 0);
  }
  contrib = warp_sum(contrib);
  int slots_row = 64;
  while (slots_row < 2 * contrib && slots_row < L.set_slots) slots_row <<= 1;
  for (int s = lane; s < slots_row; s += 32) set[s] = -1;
  __syncwarp();
  for (int ci = lane; ci < nR; ci += 32) insert(set, slots_row, rlist[ci]);
  __syncwarp();
  for (int p = lane; p < P; p += 32)
    if (s_pass[s_ci[p]]) insert(set, slots_row, s_g[p]);
  __syncwarp();

  // compact the set into the list (re-using the pair buffer), count
  int32_t* list = s_g;
  const int list_cap = list_ints < kListCap ? list_ints : kListCap;
  int nE = 0;
  for (int b0 = 0; b0 < slots_row; b0 += 32) {
    const int32_t v = set[b0 + lane];
    const unsigned b = __ballot_sync(kFull, v >= 0);
    if (v >= 0) {
      const int p = nE + __popc(b & lt);
      if (p < list_cap) list[p] = v;
    }
    nE += __popc(b);
  }
  const int cap = stride < list_cap ? stride : list_cap;
  const bool overflow = __any_sync(kFull, lost) || nE > cap;
  if (nE > cap) nE = cap;
  int n2 = 32;
  while (n2 < nE) n2 <<= 1;
  for (int t = nE + lane; t < n2; t += 32) list[t] = 0x7fffffff;
  __syncwarp();
  warp_bitonic_sort(list, n2);
  int32_t* dst = E_pad + (row - row_begin) * (int64_t)stride;
  for (int t = lane; t < nE; t += 32) dst[t] = list[t];
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

int reid_reciprocal_masks2(const int32_t* rank, int64_t N, int ncols, int k_full, int k_half, int64_t row_begin,
                           int64_t row_end, uint64_t* R_out, uint64_t* Rhalf_out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rank && R_out && Rhalf_out, "reid_reciprocal_masks2: NULL pointer");
  REID_CHECK_ARG(ncols >= 1 && ncols <= REID_MAX_K1, "reid_reciprocal_masks2: ncols=%d not in 1..%d", ncols, REID_MAX_K1);
  REID_CHECK_ARG(k_full >= 0 && k_half >= 0 && k_half <= k_full, "reid_reciprocal_masks2: k_full=%d k_half=%d", k_full, k_half);
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_reciprocal_masks2: bad row range");
  if (N == 0) return REID_OK;
  const int cf = k_full + 1 < ncols ? k_full + 1 : ncols;    // rank[i, :k+1] clamps to the stored columns
  const int ch = k_half + 1 < ncols ? k_half + 1 : ncols;
  reciprocal2_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(rank, N, ncols, cf, ch, row_begin, row_end,
                                                                              R_out, Rhalf_out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_expand(const int32_t* rank, int64_t N, int ncols, int half_cols, const uint64_t* Rmask, const uint64_t* Rhalf_mask,
                int64_t row_begin, int64_t row_end, int stride, int32_t* E_pad, int32_t* E_cnt, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rank && Rmask && Rhalf_mask && E_pad && E_cnt, "reid_expand: NULL pointer");
  REID_CHECK_ARG(ncols >= 1 && ncols <= REID_MAX_K1, "reid_expand: ncols=%d not in 1..%d", ncols, REID_MAX_K1);
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_expand: bad row range");
  REID_CHECK_ARG(stride >= 1, "reid_expand: stride=%d", stride);
  REID_CHECK_ARG(half_cols >= 1 && half_cols <= ncols, "reid_expand: half_cols=%d not in 1..ncols", half_cols);
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  int warps = kExpandWarps;                    // as many rows per CTA as fit (k1 = 64 needs ~45 KB per row)
  while (warps > 1 && (size_t)expand_layout(ncols, half_cols).bytes * warps > 100 * 1024) warps >>= 1;
  const size_t smem = (size_t)expand_layout(ncols, half_cols).bytes * warps;
  REID_CHECK_ARG(smem <= 220 * 1024, "reid_expand: ncols=%d half_cols=%d need %zu B of shared memory", ncols, half_cols, smem);
  const unsigned grid = (unsigned)((n + warps - 1) / warps);
  REID_CUDA(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  expand_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(rank, ncols, half_cols, Rmask, Rhalf_mask,
                                                                         row_begin, row_end, stride, E_pad, E_cnt);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
