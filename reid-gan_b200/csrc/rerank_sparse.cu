// a4-a6 -- CSR kernels: Gaussian weights V, k2 query expansion V_qe and the inverted index
// (utils/faiss_rerank.py:81-85, 89-94, 98-100).  No dense N x N matrix is ever formed.
#include <cuda_fp16.h>

#include "common.cuh"

namespace reid {

// ---------------------------------------------------------------------------------------
// a4: one warp per row.  s_e = fp32(fp64 dot(x_row, x_e)) -- the same canonical value the
// search produces, so members that are among the row's k1 neighbours reuse the search key
// instead of gathering 4*D bytes.  Then softmax(-(2 - 2 s)) in fp32 exactly as written in
// faiss_rerank.py:81,85 (max-subtracted, like F.softmax).
// ---------------------------------------------------------------------------------------
constexpr int kVWarps = 8;
constexpr int kVMaxRow = 1024;  // max |E| per row (expand_kernel caps at the same value)

// kChunks > 0: D == kChunks * 128, the query row is held in registers (see knn_rescore.cu).
template <int kChunks>
__global__ void __launch_bounds__(kVWarps * 32) v_weights_kernel(
    const float* __restrict__ x, int64_t D, const int32_t* __restrict__ E_pad, int stride,
    const int64_t* __restrict__ E_ptr, int64_t row_begin, int64_t row_end, const int32_t* __restrict__ rank_local,
    const float* __restrict__ key_local, int ncols, const int32_t* __restrict__ perm, int32_t* __restrict__ E_idx,
    float* __restrict__ V_val, int half) {
  __shared__ float s_val[kVWarps][kVMaxRow];
  const int w = threadIdx.x >> 5, lane = lane_id();
  const int64_t slot = (int64_t)blockIdx.x * kVWarps + w;
  if (slot >= row_end - row_begin) return;
  // optional visiting order (the cluster-locality order of the re-score stage): cluster mates gather the same few
  // feature rows, so run back to back they find them in L2
  const int64_t lr = perm ? (int64_t)perm[slot] : slot;
  const int64_t row = row_begin + lr;
  const int64_t p0 = E_ptr[lr];
  const int n = (int)(E_ptr[lr + 1] - p0);
  if (n == 0) return;
  const float* xi = x + row * D;
  float* sv = s_val[w];
  const int32_t* erow = E_pad + lr * (int64_t)stride;   // padded expansion set of this row (sorted)

  // the row's neighbour list and search keys, two entries per lane (ncols <= 64)
  int32_t rk0 = -1, rk1 = -1;
  float kv0 = 0.f, kv1 = 0.f;
  if (rank_local) {
    if (lane < ncols) {
      rk0 = rank_local[lr * ncols + lane];
      kv0 = key_local[lr * ncols + lane];
    }
    if (lane + 32 < ncols) {
      rk1 = rank_local[lr * ncols + lane + 32];
      kv1 = key_local[lr * ncols + lane + 32];
    }
  }
  constexpr int kRegs = kChunks > 0 ? kChunks : 1;
  float4 q[kRegs];
  bool q_loaded = false;

  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    const int32_t mine = e < n ? erow[e] : -1;
    if (e < n) E_idx[p0 + e] = mine;                      // compacted into the CSR on the way
    const int m = min(32, n - e0);
    for (int u = 0; u < m; ++u) {
      const int32_t j = __shfl_sync(kFull, mine, u);
      // reuse the search key when j is one of the row's stored neighbours
      const unsigned b0 = __ballot_sync(kFull, rk0 == j), b1 = __ballot_sync(kFull, rk1 == j);
      float s;
      if (b0) {
        s = __shfl_sync(kFull, kv0, __ffs(b0) - 1);
      } else if (b1) {
        s = __shfl_sync(kFull, kv1, __ffs(b1) - 1);
      } else {
        const float* xj = x + (int64_t)j * D;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        if (kChunks > 0) {
          if (!q_loaded) {
#pragma unroll
            for (int c = 0; c < kRegs; ++c) q[c] = reinterpret_cast<const float4*>(xi)[c * 32 + lane];
            q_loaded = true;
          }
          float4 r[kRegs];
#pragma unroll
          for (int c = 0; c < kRegs; ++c) r[c] = reinterpret_cast<const float4*>(xj)[c * 32 + lane];
#pragma unroll
          for (int c = 0; c < kRegs; ++c) {
            acc0 = fma((double)q[c].x, (double)r[c].x, acc0);
            acc1 = fma((double)q[c].y, (double)r[c].y, acc1);
            acc2 = fma((double)q[c].z, (double)r[c].z, acc2);
            acc3 = fma((double)q[c].w, (double)r[c].w, acc3);
          }
        } else if ((D & 3) == 0) {
          const float4* a4 = reinterpret_cast<const float4*>(xi);
          const float4* b4 = reinterpret_cast<const float4*>(xj);
          for (int64_t d = lane; d < (D >> 2); d += 32) {
            const float4 a = a4[d], r = b4[d];
            acc0 = fma((double)a.x, (double)r.x, acc0);
            acc1 = fma((double)a.y, (double)r.y, acc1);
            acc2 = fma((double)a.z, (double)r.z, acc2);
            acc3 = fma((double)a.w, (double)r.w, acc3);
          }
        } else {
          for (int64_t d = lane; d < D; d += 32) acc0 = fma((double)xi[d], (double)xj[d], acc0);
        }
        s = (float)warp_sum((acc0 + acc1) + (acc2 + acc3));
      }
      if (lane == 0) sv[e0 + u] = s;
    }
  }
  __syncwarp();
  // softmax(-dist), dist = 2 - 2 s
  float mx = -INFINITY;
  for (int e = lane; e < n; e += 32) {
    const float neg = -(2.0f - 2.0f * sv[e]);
    sv[e] = neg;
    mx = fmaxf(mx, neg);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int e = lane; e < n; e += 32) {
    const float ex = expf(sv[e] - mx);
    sv[e] = ex;
    sum += ex;
  }
  sum = warp_sum(sum);
  // use_float16=True (faiss_rerank.py:82-83): the fp32 softmax is stored as float16
  for (int e = lane; e < n; e += 32) {
    const float v = __fdiv_rn(sv[e], sum);
    V_val[p0 + e] = half ? __half2float(__float2half_rn(v)) : v;
  }
}

// ---------------------------------------------------------------------------------------
// a5: one warp per row.  V_qe[i, :] = mean_{r < k2} V[rank[i, r], :]  (np.mean: the k2 rows are added in r
// order in fp32, then divided by k2).  The k2 neighbour rows are folded one after the other (r ascending) into a
// per-warp open-addressing table column -> running sum: inside one V row every column is distinct, so the lanes
// of a step never meet on a slot and each column's additions happen in r order -- the reference's arithmetic
// restricted to the structural non-zeros.  The distinct columns (a few dozen) are then compacted and sorted.
// ---------------------------------------------------------------------------------------
constexpr int kQWarps = 8;

__device__ __forceinline__ void warp_bitonic_sort_kv(uint32_t* key, float* val, int n2) {
  const int lane = lane_id();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < n2; t += 32) {
        const int p = t ^ j;
        if (p > t) {
          const uint32_t a = key[t], b = key[p];
          const bool up = ((t & k) == 0);
          if ((a > b) == up) {
            key[t] = b;
            key[p] = a;
            const float va = val[t];
            val[t] = val[p];
            val[p] = va;
          }
        }
      }
      __syncwarp();
    }
  }
}

// cap = slots of the table = padded output slots per row, a power of two >= k2 * max_row_nnz: the table only has
// to hold the DISTINCT columns (<= k2 * max_row_nnz, so probing always terminates) and the k2 rows overlap heavily,
// which keeps the load factor around 1/8 in practice.
__global__ void __launch_bounds__(kQWarps * 32) query_expand_kernel(
    const int32_t* __restrict__ rank, int ncols, int k2, const int64_t* __restrict__ V_ptr,
    const int32_t* __restrict__ V_idx, const float* __restrict__ V_val, int cap, int64_t row_begin, int64_t row_end,
    int32_t* __restrict__ Q_cnt, int32_t* __restrict__ Q_idx, float* __restrict__ Q_val,
    unsigned long long* __restrict__ overflow_rows, int half) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = lane_id();
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)w * cap;              // table keys, then sort keys
  const int n_warps = blockDim.x >> 5;                  // rows per CTA (fewer than kQWarps when the tables are big)
  float* val = reinterpret_cast<float*>(smem_raw + (size_t)n_warps * cap * sizeof(uint32_t)) + (size_t)w * cap;
  const int64_t row = row_begin + (int64_t)blockIdx.x * n_warps + w;
  if (row >= row_end) return;
  const uint32_t smask = (uint32_t)cap - 1u;
  for (int t = lane; t < cap; t += 32) key[t] = 0xffffffffu;
  __syncwarp();
  // A speculatively sized table (the caller did not know the longest V row) can fill up.  Before a V row is folded
  // in, its length is checked against the free slots (warp-uniform, conservative: every entry could be a new column),
  // so the probing below always finds a free slot; a row that might not fit reports Q_cnt = 0 and is counted in
  // *overflow_rows (the caller redoes the stage with the exact size).
  bool lost = false;
  int used = 0;
  const int limit = cap - (cap >> 2);
  for (int r = 0; r < k2 && !lost; ++r) {
    const int64_t j = rank[row * ncols + r];
    const int64_t a = V_ptr[j];
    const int m = (int)(V_ptr[j + 1] - a);
    if (used + m > limit) {                          // same on every lane
      lost = true;
      break;
    }
    int fresh = 0;
    for (int e = lane; e < m; e += 32) {
      const uint32_t c = (uint32_t)V_idx[a + e];
      const float v = V_val[a + e];
      uint32_t h = (c * 0x9e3779b1u) >> 7 & smask;
      while (true) {
        const uint32_t old = atomicCAS(&key[h], 0xffffffffu, c);
        if (old == 0xffffffffu) {
          val[h] = v;
          ++fresh;
          break;
        }
        if (old == c) {
          val[h] = __fadd_rn(val[h], v);
          break;
        }
        h = (h + 1) & smask;
      }
    }
    used += __reduce_add_sync(kFull, fresh);          // also: row r is folded before row r + 1 starts
  }
  if (lost) {
    if (lane == 0) {
      Q_cnt[row - row_begin] = 0;
      if (overflow_rows) atomicAdd(overflow_rows, 1ull);
    }
    return;
  }
  // compact the occupied slots to the front (in place: the write position never passes the read position); the mean
  // is taken here: sum / k2 in fp32 (np.mean), stored as float16 when use_float16 -- an entry that rounds to zero is
  // no longer a structural non-zero of V_qe (the reference's invIndex is np.where(V != 0), :98-100)
  const float k2f = (float)k2;
  int n = 0;
  for (int base = 0; base < cap; base += 32) {
    const uint32_t c = key[base + lane];
    float v = __fdiv_rn(val[base + lane], k2f);
    if (half) v = __half2float(__float2half_rn(v));
    const bool occ = c != 0xffffffffu && !(half && v == 0.f);
    const unsigned b = __ballot_sync(kFull, occ);
    __syncwarp();
    if (occ) {
      const int p = n + __popc(b & ((1u << lane) - 1u));
      key[p] = c;
      val[p] = v;
    }
    n += __popc(b);
    __syncwarp();
  }
  int n2 = 32;
  while (n2 < n) n2 <<= 1;
  for (int t = n + lane; t < n2; t += 32) {
    key[t] = 0xffffffffu;
    val[t] = 0.f;
  }
  __syncwarp();
  warp_bitonic_sort_kv(key, val, n2);
  const int64_t out = (row - row_begin) * (int64_t)cap;   // padded output: `cap` slots per row
  for (int t = lane; t < n; t += 32) {
    Q_idx[out + t] = (int32_t)key[t];
    Q_val[out + t] = val[t];
  }
  if (lane == 0) Q_cnt[row - row_begin] = n;
}

// padded rows (stride slots, cnt valid) -> CSR at ptr
__global__ void __launch_bounds__(256) csr_compact_kernel(const int32_t* __restrict__ pad_idx,
                                                          const float* __restrict__ pad_val, int64_t stride,
                                                          const int32_t* __restrict__ cnt,
                                                          const int64_t* __restrict__ ptr, int64_t n_rows,
                                                          int32_t* __restrict__ out_idx, float* __restrict__ out_val) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int c = cnt[row];
  const int64_t src = row * stride, dst = ptr[row];
  for (int t = lane_id(); t < c; t += 32) {
    out_idx[dst + t] = pad_idx[src + t];
    out_val[dst + t] = pad_val[src + t];
  }
}

// lists in upper-bound slots (slot_ptr, cnt) -> packed at ptr (indices only)
__global__ void __launch_bounds__(256) lists_compact_kernel(const int64_t* __restrict__ slot_ptr,
                                                            const int32_t* __restrict__ idx,
                                                            const int32_t* __restrict__ cnt,
                                                            const int64_t* __restrict__ ptr, int64_t n_rows,
                                                            int32_t* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int c = cnt[row];
  const int64_t src = slot_ptr[row], dst = ptr[row];
  for (int t = lane_id(); t < c; t += 32) out[dst + t] = idx[src + t];
}

// ---------------------------------------------------------------------------------------
// Exchange format of the row-sharded multi-GPU plan: one fixed-stride record per row,
//   rec[row] = { count, idx[0 .. stride), (val bits[0 .. stride)) }   (int32 words),
// so that ONE all-gather moves a ragged stage output (V rows, V_qe rows, eps-neighbour lists) and every rank
// rebuilds the global CSR from the gathered records itself -- no per-rank length exchange, no padding copies.
// A row longer than `stride` keeps its true count in the record (the receivers see the overflow and the caller
// falls back to the variable-length gather).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_pack_kernel(const int32_t* __restrict__ cnt, const int64_t* __restrict__ ptr,
                                                        const int32_t* __restrict__ idx, const float* __restrict__ val,
                                                        int64_t n_rows, int64_t n_rows_padded, int stride, int words,
                                                        int32_t* __restrict__ rec) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows_padded) return;
  int32_t* r = rec + row * words;
  const int lane = lane_id();
  if (row >= n_rows) {                     // padding rows of the last block
    if (lane == 0) r[0] = 0;
    return;
  }
  const int c = cnt[row];
  if (lane == 0) r[0] = c;
  const int m = c < stride ? c : stride;
  const int64_t a = ptr[row];
  for (int t = lane; t < m; t += 32) {
    r[1 + t] = idx[a + t];
    if (val) r[1 + stride + t] = __float_as_int(val[a + t]);
  }
}

// rank that owns global row i under the partition `bounds` (W + 1 ascending entries)
__device__ __forceinline__ int rows_owner(const int64_t* __restrict__ bounds, int W, int64_t i) {
  int r = 0;
  while (r + 1 < W && i >= bounds[r + 1]) ++r;
  return r;
}

__global__ void __launch_bounds__(256) rows_unpack_counts_kernel(const int32_t* __restrict__ rec, int words, int W,
                                                                 int64_t max_rows, const int64_t* __restrict__ bounds,
                                                                 int64_t N, int32_t* __restrict__ g_cnt, int stride,
                                                                 unsigned long long* __restrict__ overflow_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int r = rows_owner(bounds, W, i);
  int c = rec[((int64_t)r * max_rows + (i - bounds[r])) * words];
  if (overflow_rows && c > stride) {      // a caller that does not read sizes back: keep what travelled, report the row
    atomicAdd(overflow_rows, 1ull);
    c = stride;
  }
  g_cnt[i] = c;
}

__global__ void __launch_bounds__(256) rows_unpack_fill_kernel(const int32_t* __restrict__ rec, int words, int stride,
                                                               int W, int64_t max_rows, const int64_t* __restrict__ bounds,
                                                               int64_t N, const int64_t* __restrict__ g_ptr,
                                                               int32_t* __restrict__ out_idx, float* __restrict__ out_val) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int r = rows_owner(bounds, W, i);
  const int32_t* src = rec + ((int64_t)r * max_rows + (i - bounds[r])) * words;
  const int c = min(src[0], stride);
  const int64_t o = g_ptr[i];
  for (int t = lane_id(); t < c; t += 32) {
    out_idx[o + t] = src[1 + t];
    if (out_val) out_val[o + t] = __int_as_float(src[1 + stride + t]);
  }
}

// ---------------------------------------------------------------------------------------
// a6: CSR -> CSC.  Column histogram, (caller scans), atomic-cursor scatter, then each
// column list is sorted by row so the result does not depend on scheduling.
// ---------------------------------------------------------------------------------------
__global__ void col_count_kernel(const int32_t* __restrict__ idx, int64_t nnz_host, const int64_t* __restrict__ nnz_dev,
                                 int32_t* __restrict__ cnt) {
  const int64_t nnz = nnz_dev ? *nnz_dev : nnz_host;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&cnt[idx[p]], 1);
}

__global__ void __launch_bounds__(256) col_scatter_kernel(const int64_t* __restrict__ ptr,
                                                          const int32_t* __restrict__ idx,
                                                          const float* __restrict__ val, int64_t n_rows,
                                                          const int64_t* __restrict__ C_ptr,
                                                          int32_t* __restrict__ cursor, int32_t* __restrict__ C_idx,
                                                          float* __restrict__ C_val) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  for (int64_t p = ptr[row] + lane_id(); p < ptr[row + 1]; p += 32) {
    const int32_t c = idx[p];
    const int64_t q = C_ptr[c] + atomicAdd(&cursor[c], 1);
    C_idx[q] = (int32_t)row;
    C_val[q] = val[p];
  }
}

constexpr int kColSortCap = 2048;
constexpr int kColSortWarps = 4;

__global__ void __launch_bounds__(kColSortWarps * 32) col_sort_kernel(const int64_t* __restrict__ C_ptr, int64_t n_cols,
                                                                      int32_t* __restrict__ C_idx,
                                                                      float* __restrict__ C_val, int cap, int skip_long) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = lane_id();
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)w * cap;
  float* val = reinterpret_cast<float*>(smem_raw + (size_t)kColSortWarps * cap * 4) + (size_t)w * cap;
  const int64_t c = (int64_t)blockIdx.x * kColSortWarps + w;
  if (c >= n_cols) return;
  const int64_t a = C_ptr[c];
  const int64_t n = C_ptr[c + 1] - a;
  if (n <= 1) return;
  if (n <= cap) {
    int n2 = 32;
    while (n2 < n) n2 <<= 1;
    for (int t = lane; t < n2; t += 32) {
      key[t] = t < n ? (uint32_t)C_idx[a + t] : 0xffffffffu;
      val[t] = t < n ? C_val[a + t] : 0.f;
    }
    __syncwarp();
    warp_bitonic_sort_kv(key, val, n2);
    for (int t = lane; t < n; t += 32) {
      C_idx[a + t] = (int32_t)key[t];
      C_val[a + t] = val[t];
    }
  } else if (!skip_long) {
    // very long column: odd-even transposition sort in place (rare; correctness path)
    for (int64_t round = 0; round < n; ++round) {
      for (int64_t t = (round & 1) + 2 * (int64_t)lane; t + 1 < n; t += 64) {
        const int32_t x0 = C_idx[a + t], x1 = C_idx[a + t + 1];
        if (x0 > x1) {
          C_idx[a + t] = x1;
          C_idx[a + t + 1] = x0;
          const float v = C_val[a + t];
          C_val[a + t] = C_val[a + t + 1];
          C_val[a + t + 1] = v;
        }
      }
      __syncwarp();
    }
  }
}

// Columns longer than col_sort_kernel's shared-memory capacity (hub columns; only met when the caller could not
// size the buffers for the longest column): one CTA per such column, bitonic sort of up to kColLongCap entries in
// shared memory, odd-even transposition in place beyond that.  Every CTA walks the columns with a stride and skips
// the short ones, so no queue and no host knowledge of the lengths is needed.
constexpr int kColLongCap = 8192;
constexpr int kColLongThreads = 512;
__global__ void __launch_bounds__(kColLongThreads) col_sort_long_kernel(const int64_t* __restrict__ C_ptr, int64_t n_cols,
                                                                        int32_t* __restrict__ C_idx, float* __restrict__ C_val,
                                                                        int short_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_raw);
  float* val = reinterpret_cast<float*>(smem_raw + (size_t)kColLongCap * 4);
  const int t = threadIdx.x;
  for (int64_t c0 = (int64_t)blockIdx.x * kColLongThreads; c0 < n_cols; c0 += (int64_t)gridDim.x * kColLongThreads) {
    // which of my 512 columns are long?  (warp-uniform handling: the CTA serves them one after the other)
    const int64_t c = c0 + t;
    const int64_t len = c < n_cols ? C_ptr[c + 1] - C_ptr[c] : 0;
    __shared__ int s_long[kColLongThreads];
    __shared__ int s_n;
    if (t == 0) s_n = 0;
    __syncthreads();
    if (len > short_cap) s_long[atomicAdd(&s_n, 1)] = t;
    __syncthreads();
    const int n_long = s_n;
    for (int q = 0; q < n_long; ++q) {
      const int64_t col = c0 + s_long[q];
      const int64_t a = C_ptr[col];
      const int64_t n = C_ptr[col + 1] - a;
      if (n <= kColLongCap) {
        int n2 = 64;
        while (n2 < n) n2 <<= 1;
        for (int u = t; u < n2; u += kColLongThreads) {
          key[u] = u < n ? (uint32_t)C_idx[a + u] : 0xffffffffu;
          val[u] = u < n ? C_val[a + u] : 0.f;
        }
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1) {
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int u = t; u < n2; u += kColLongThreads) {
              const int p = u ^ j;
              if (p > u) {
                const uint32_t x0 = key[u], x1 = key[p];
                const bool up = ((u & k) == 0);
                if ((x0 > x1) == up) {
                  key[u] = x1;
                  key[p] = x0;
                  const float v = val[u];
                  val[u] = val[p];
                  val[p] = v;
                }
              }
            }
            __syncthreads();
          }
        }
        for (int u = t; u < n; u += kColLongThreads) {
          C_idx[a + u] = (int32_t)key[u];
          C_val[a + u] = val[u];
        }
        __syncthreads();
      } else {
        for (int64_t round = 0; round < n; ++round) {
          for (int64_t u = (round & 1) + 2 * (int64_t)t; u + 1 < n; u += 2 * kColLongThreads) {
            const int32_t x0 = C_idx[a + u], x1 = C_idx[a + u + 1];
            if (x0 > x1) {
              C_idx[a + u] = x1;
              C_idx[a + u + 1] = x0;
              const float v = C_val[a + u];
              C_val[a + u] = C_val[a + u + 1];
              C_val[a + u + 1] = v;
            }
          }
          __syncthreads();
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace reid

extern "C" {

int reid_v_weights(const float* x, int64_t N, int64_t D, const int32_t* E_pad, int stride, const int64_t* E_ptr,
                   int64_t row_begin, int64_t row_end, const int32_t* rank_local, const float* rank_key_local,
                   int ncols, const int32_t* visit_order, int32_t* E_idx, float* V_val, int half_precision, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && E_pad && E_ptr && E_idx && V_val, "reid_v_weights: NULL pointer");
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N && D > 0 && stride >= 1,
                 "reid_v_weights: bad shape");
  REID_CHECK_ARG(stride <= kVMaxRow, "reid_v_weights: stride=%d exceeds %d", stride, kVMaxRow);
  REID_CHECK_ARG((rank_local == nullptr) == (rank_key_local == nullptr), "reid_v_weights: rank and keys go together");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
#define REID_V_LAUNCH(CH)                                                                                          \
  v_weights_kernel<CH><<<(unsigned)((n + kVWarps - 1) / kVWarps), kVWarps * 32, 0, (cudaStream_t)stream>>>(          \
      x, D, E_pad, stride, E_ptr, row_begin, row_end, rank_local, rank_key_local, ncols, visit_order, E_idx, V_val,   \
      half_precision)
  const bool aligned = (((uintptr_t)x) & 15) == 0;
  if (aligned && D == 2048) REID_V_LAUNCH(16);
  else if (aligned && D == 1024) REID_V_LAUNCH(8);
  else if (aligned && D == 512) REID_V_LAUNCH(4);
  else if (aligned && D == 256) REID_V_LAUNCH(2);
  else REID_V_LAUNCH(0);
#undef REID_V_LAUNCH
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_query_expand_stride(int k2, int max_row_nnz) {
  // 3/4 of the slots must hold k2 rows of max_row_nnz entries: with the TRUE maximum the kernel's (conservative)
  // free-slot check can then never reject a row
  int cap = 32;
  while (cap - (cap >> 2) < k2 * max_row_nnz) cap <<= 1;
  return cap;
}

int reid_query_expand(const int32_t* rank, int64_t N, int ncols, int k2, const int64_t* V_ptr, const int32_t* V_idx,
                      const float* V_val, int max_row_nnz, int64_t row_begin, int64_t row_end, int32_t* Q_cnt,
                      int32_t* Q_pad_idx, float* Q_pad_val, uint64_t* overflow_rows, int half_precision, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rank && V_ptr && V_idx && V_val && Q_cnt && Q_pad_idx && Q_pad_val, "reid_query_expand: NULL pointer");
  REID_CHECK_ARG(k2 >= 1 && k2 <= ncols && ncols <= REID_MAX_K1, "reid_query_expand: k2=%d ncols=%d", k2, ncols);
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_query_expand: bad row range");
  REID_CHECK_ARG(max_row_nnz >= 1, "reid_query_expand: max_row_nnz=%d", max_row_nnz);
  REID_CHECK_ARG(N < 0xffffffffll, "reid_query_expand: N exceeds the 32-bit column key");
  const int cap = reid_query_expand_stride(k2, max_row_nnz);
  // rows per CTA: as many warps as fit into shared memory (a large k1 makes the tables big)
  int warps = kQWarps;
  while (warps > 1 && (size_t)warps * cap * 8 > 96 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * cap * 8;
  REID_CHECK_ARG(smem <= 220 * 1024, "reid_query_expand: k2 * max_row_nnz = %d needs %zu B of shared memory per row",
                 k2 * max_row_nnz, smem);
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  const unsigned grid = (unsigned)((n + warps - 1) / warps);
  REID_CUDA(cudaFuncSetAttribute(query_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  query_expand_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(
      rank, ncols, k2, V_ptr, V_idx, V_val, cap, row_begin, row_end, Q_cnt, Q_pad_idx, Q_pad_val,
      (unsigned long long*)overflow_rows, half_precision);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_csr_compact(const int32_t* pad_idx, const float* pad_val, int64_t stride, const int32_t* cnt,
                     const int64_t* ptr, int64_t n_rows, int32_t* out_idx, float* out_val, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(pad_idx && pad_val && cnt && ptr && out_idx && out_val && stride >= 1 && n_rows >= 0,
                 "reid_csr_compact: bad arguments");
  if (n_rows == 0) return REID_OK;
  csr_compact_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(pad_idx, pad_val, stride, cnt, ptr,
                                                                                    n_rows, out_idx, out_val);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_lists_compact(const int64_t* slot_ptr, const int32_t* idx, const int32_t* cnt, const int64_t* ptr,
                       int64_t n_rows, int32_t* out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(slot_ptr && idx && cnt && ptr && out && n_rows >= 0, "reid_lists_compact: bad arguments");
  if (n_rows == 0) return REID_OK;
  lists_compact_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(slot_ptr, idx, cnt, ptr, n_rows,
                                                                                      out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_rows_pack(const int32_t* cnt, const int64_t* ptr, const int32_t* idx, const float* val, int64_t n_rows,
                   int64_t n_rows_padded, int stride, int32_t* rec, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rec && stride >= 1 && n_rows >= 0 && n_rows_padded >= n_rows, "reid_rows_pack: bad arguments");
  REID_CHECK_ARG(n_rows == 0 || (cnt && ptr && idx), "reid_rows_pack: NULL pointer");
  if (n_rows_padded == 0) return REID_OK;
  const int words = 1 + stride * (val ? 2 : 1);
  rows_pack_kernel<<<(unsigned)((n_rows_padded + 7) / 8), 256, 0, (cudaStream_t)stream>>>(cnt, ptr, idx, val, n_rows,
                                                                                       n_rows_padded, stride, words, rec);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_rows_unpack_counts(const int32_t* rec, int stride, int has_val, int world, int64_t max_rows, const int64_t* bounds,
                            int64_t N, int32_t* g_cnt, uint64_t* overflow_rows, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rec && bounds && g_cnt && stride >= 1 && world >= 1 && max_rows >= 0 && N >= 0, "reid_rows_unpack_counts: bad arguments");
  if (N == 0) return REID_OK;
  const int words = 1 + stride * (has_val ? 2 : 1);
  rows_unpack_counts_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rec, words, world, max_rows, bounds,
                                                                                         N, g_cnt, stride,
                                                                                         (unsigned long long*)overflow_rows);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_rows_unpack_fill(const int32_t* rec, int stride, int world, int64_t max_rows, const int64_t* bounds, int64_t N,
                          const int64_t* g_ptr, int32_t* out_idx, float* out_val, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(rec && bounds && g_ptr && out_idx && stride >= 1 && world >= 1 && max_rows >= 0 && N >= 0,
                 "reid_rows_unpack_fill: bad arguments");
  if (N == 0) return REID_OK;
  const int words = 1 + stride * (out_val ? 2 : 1);
  rows_unpack_fill_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(rec, words, stride, world, max_rows,
                                                                                   bounds, N, g_ptr, out_idx, out_val);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_transpose_count(const int32_t* idx, int64_t nnz, const int64_t* nnz_dev, int64_t n_cols, int32_t* col_cnt,
                         void* stream) {
  using namespace reid;
  REID_CHECK_ARG(col_cnt && (nnz == 0 || idx) && n_cols > 0, "reid_transpose_count: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  REID_CUDA(cudaMemsetAsync(col_cnt, 0, sizeof(int32_t) * (size_t)n_cols, st));
  if (nnz == 0) return REID_OK;
  int64_t blocks = (nnz + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  col_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx, nnz, nnz_dev, col_cnt);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_transpose_fill(const int64_t* ptr, const int32_t* idx, const float* val, int64_t n_rows, int64_t n_cols,
                        const int64_t* C_ptr, int32_t* cursor, int32_t* C_idx, float* C_val, int max_col_len,
                        void* stream) {
  using namespace reid;
  REID_CHECK_ARG(ptr && idx && val && C_ptr && cursor && C_idx && C_val, "reid_transpose_fill: NULL pointer");
  REID_CHECK_ARG(n_rows >= 0 && n_cols > 0, "reid_transpose_fill: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  REID_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n_cols, st));
  if (n_rows == 0) return REID_OK;
  col_scatter_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(ptr, idx, val, n_rows, C_ptr, cursor, C_idx, C_val);
  REID_LAUNCH_CHECK();
  // shared-memory sort buffers sized for the longest column when the caller knows it (from the scan of the counts).
  // max_col_len <= 0: unknown (the caller does not read sizes back): warps sort the columns of up to 256 entries,
  // a second launch finds the longer ones itself and gives each a whole CTA.
  const bool unknown = max_col_len <= 0;
  int cap = 32;
  while (cap < (unknown ? 256 : max_col_len) && cap < kColSortCap) cap <<= 1;
  const size_t smem = (size_t)kColSortWarps * cap * 8;
  REID_CUDA(cudaFuncSetAttribute(col_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  col_sort_kernel<<<(unsigned)((n_cols + kColSortWarps - 1) / kColSortWarps), kColSortWarps * 32, smem, st>>>(
      C_ptr, n_cols, C_idx, C_val, cap, unknown ? 1 : 0);
  REID_LAUNCH_CHECK();
  if (unknown) {
    const size_t smem_long = (size_t)kColLongCap * 8;
    REID_CUDA(cudaFuncSetAttribute(col_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_long));
    int64_t g = (n_cols + kColLongThreads - 1) / kColLongThreads;
    if (g > 2 * num_sms()) g = 2 * num_sms();
    col_sort_long_kernel<<<(unsigned)g, kColLongThreads, smem_long, st>>>(C_ptr, n_cols, C_idx, C_val, cap);
    REID_LAUNCH_CHECK();
  }
  return REID_OK;
}
}
