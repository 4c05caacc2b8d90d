// a7 -- Jaccard min-sum on the sparse V_qe (utils/faiss_rerank.py:102-119).
//   t_ij = sum_{c in nz(i) & nz(j), c ascending} min(Vq[i,c], Vq[j,c])   (sequential fp32 adds)
//   J_ij = max(0, 1 - t_ij / (2 - t_ij));  J_ij = 1 exactly when no column is shared.
// The reference walks the non-zero columns of row i in ascending order and, for each, adds
// into temp_min[rows of that column] (:109-110); both kernels below keep exactly that order,
// which is what makes J bit-symmetric and independent of how rows are sharded.
#include <cuda_fp16.h>

#include "common.cuh"

namespace reid {

// T_i = sum_{c in nz(i)} |col(c)|: every (column, partner) pair of the row -- the work of the row and an upper bound of
// its partner count.  S_i (optional) = a much tighter bound of the eps-NEIGHBOUR count: sum_j t_ij <= B_i = sum_c
// V_ic |col(c)| (every min() is at most V_ic), so at most B_i / t_min partners reach t >= t_min (Markov), t_min the
// smallest t with J(t) <= eps.  For a row inside an identity cluster B_i is about the cluster size.  The emitting
// kernels never write past a row's slots (a violated bound is reported, not trusted).
__global__ void __launch_bounds__(256) jaccard_bounds_kernel(const int64_t* __restrict__ Q_ptr,
                                                             const int32_t* __restrict__ Q_idx,
                                                             const float* __restrict__ Q_val,
                                                             const int64_t* __restrict__ C_ptr, int64_t row_begin,
                                                             int64_t row_end, float t_min, int32_t* __restrict__ T_cnt,
                                                             int32_t* __restrict__ S_cnt, int32_t* __restrict__ P_cnt) {
  const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  int64_t s = 0;
  int64_t longest = 0;
  double b = 0.0;
  const int64_t qa = Q_ptr[row], qb = Q_ptr[row + 1];
  for (int64_t p = qa + lane_id(); p < qb; p += 32) {
    const int32_t c = Q_idx[p];
    const int64_t len = C_ptr[c + 1] - C_ptr[c];
    s += len;
    longest = len > longest ? len : longest;
    if (S_cnt) b += (double)Q_val[p] * (double)len;
  }
  s = warp_sum(s);
  longest = warp_max(longest);
  b = warp_sum(b);
  if (lane_id() == 0) {
    const int32_t t = (int32_t)(s > 0x7fffffff ? 0x7fffffff : s);
    T_cnt[row - row_begin] = t;
    if (P_cnt) {
      // GUESS of the number of distinct partners (sizes the hash table of the first attempt; a wrong guess only moves
      // the row to a bigger table).  T counts every (column, partner) pair and over-estimates the partners ~14x on
      // clustered data (a partner shares most of the row's columns).  A row inside an identity cluster of size m has
      // columns of length ~m and ~m..2m partners, so 3 x the longest column + the row's own nnz is a much closer
      // guess; it is never taken above the T-based one.
      // It is the wrong guess on shapes whose neighbourhoods span several identities (Market shape: partners ~ T / 2,
      // 98 % of the rows then waste their first attempt): the eps-graph entry counts the rows it had to move up, and the
      // host layer stops asking for this guess on a data set where more than three quarters of the rows needed it.
      const int64_t by_t = t < 2048 ? t >> 2 : t >> 1;
      const int64_t by_col = 3 * longest + (qb - qa);
      P_cnt[row - row_begin] = (int32_t)(by_col < by_t ? by_col : by_t);
    }
    if (S_cnt) {
      double need = t_min > 0.f ? b / (double)t_min + 2.0 : (double)t;
      if (need > (double)t) need = (double)t;
      S_cnt[row - row_begin] = (int32_t)need;
    }
  }
}

__device__ __forceinline__ uint32_t jhash(uint32_t v) {
  v *= 0x9e3779b1u;
  return v ^ (v >> 15);
}

// use_float16=True (faiss_rerank.py:37): V, V_qe, temp_min and jaccard_dist are float16 arrays, and numpy evaluates
// every float16 operation in float32 and rounds the result to float16.  `half` switches those roundings on: the sum
// is rounded after every column (:110) and each of the three operations of  1 - t / (2 - t)  (:113) separately.
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float acc_add(float a, float b, int half) {
  const float s = __fadd_rn(a, b);
  return half ? round_h(s) : s;
}
// J is symmetric bit for bit (both rows add the same minima over the same shared columns in the same ascending
// order), so an eps-graph consumer that only needs every EDGE once (union-find, degree counting: csrc/dbscan.cu) lets
// each unordered pair {i, j} be accumulated by exactly one of its two rows -- half the table updates.  The owner is
// chosen by the parity of i + j, which splits every row's partners evenly whatever its index (a plain "j > i" rule
// would leave the low rows with all the work, and the first rank of a row-sharded pass with it).  (i, i) is i's own.
__device__ __forceinline__ bool pair_owned(int32_t i, int32_t j) {
  return i == j ? true : ((i < j) == (((i + j) & 1) == 0));
}

__device__ __forceinline__ float jaccard_from_t(float t, int half = 0) {
  float j;
  if (half) j = round_h(__fsub_rn(1.0f, round_h(__fdiv_rn(t, round_h(__fsub_rn(2.0f, t))))));
  else j = __fsub_rn(1.0f, __fdiv_rn(t, __fsub_rn(2.0f, t)));
  return j < 0.f ? 0.f : j;
}

// eps-neighbourhoods, one warp per row, per-warp open-addressing table (j -> running t) in shared memory.
//
// The reference's inner loop is "for each non-zero column c of row i (ascending): temp_min[rows of c] += min(..)"
// (:109-110).  The warp loads the metadata of 32 columns at once (lane-parallel: column id, V_ic, start, length),
// then walks the columns in order, 32 entries of one column per step, with the next step's loads in flight --
// a column no longer costs three dependent round trips.  Every t_ij is the exact sequential fp32 sum of the
// reference, bit-symmetric and independent of sharding.
//
// Rows are dealt to table-size classes on the device (jaccard_classify_kernel); a row that overflows its table
// is pushed to the next class's queue, the last resort being the dense-accumulator kernel further down.  Each
// class is one persistent launch that reads its queue length from device memory: no host round trip.
constexpr int kJWarps = 4;
constexpr int kJClasses = 5;                 // 512, 1024, 2048, 4096, 8192 slots
constexpr int kJE = 1;                       // entries of a column per lane and step, table kernel (registers = occupancy)
constexpr int kJHeavyCtas = 64;              // dense-accumulator rows processed at a time by the last resort

__host__ __device__ constexpr int jclass_slots(int c) { return 512 << c; }

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) jaccard_neighbors_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val,
    int64_t row_begin, int64_t n_rows_host, const int32_t* __restrict__ queue, const int32_t* __restrict__ queue_len,
    int32_t* __restrict__ next_queue, int32_t* __restrict__ next_len, float eps,
    const int64_t* __restrict__ slot_ptr, int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_val,
    int32_t* __restrict__ nbr_cnt, int slots, int64_t nbr_capacity, unsigned long long* __restrict__ slot_overflow,
    int half, const int32_t* __restrict__ T_cnt, int32_t* __restrict__ queues, int32_t* __restrict__ qlen, int64_t q_stride,
    int cur_class, int direct_from, int owned_only, unsigned long long* __restrict__ escalated) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = lane_id();
  int32_t* tkey = reinterpret_cast<int32_t*>(smem_raw) + (size_t)w * slots;
  float* tval = reinterpret_cast<float*>(smem_raw + (size_t)kWarps * slots * 4) + (size_t)w * slots;
  const uint32_t smask = (uint32_t)slots - 1u;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t n_rows = queue_len ? (int64_t)*queue_len : n_rows_host;
  const int limit = (slots >> 1) + (slots >> 2);      // load factor < 3/4 so probing terminates

  for (int64_t li = (int64_t)blockIdx.x * kWarps + w; li < n_rows; li += (int64_t)gridDim.x * kWarps) {
    const int64_t lr = queue ? queue[li] : li;        // local row id
    const int64_t row = row_begin + lr;
    for (int s = lane; s < slots; s += 32) tkey[s] = -1;
    __syncwarp();

    int used = 0;
    bool overflow = false;
    const int64_t qa = Q_ptr[row], qb = Q_ptr[row + 1];
    for (int64_t pc = qa; pc < qb && !overflow; pc += 32) {   // up to 32 columns of the row at a time
      const int64_t p = pc + lane;
      const bool has = p < qb;
      const int32_t c = has ? Q_idx[p] : 0;
      const float vic = has ? Q_val[p] : 0.f;
      const int64_t ca = has ? C_ptr[c] : 0;
      const int len = has ? (int)(C_ptr[c + 1] - ca) : 0;
      const int ncol = (int)min((int64_t)32, qb - pc);
      // Walk the columns in ascending order, 32 entries of ONE column per step: inside a column every j is
      // distinct, so a step needs no ordering between its lanes, and the step sequence is the reference's
      // accumulation order.  The loads of the next step are issued before the current one goes through the table.
      // kJE entries per lane and step (128 entries of ONE column per step): most columns are a single step, and
      // with the next step's loads issued ahead a lane keeps 2 * kJE independent loads in flight
      int k = 0, base = 0;
      int64_t cak = __shfl_sync(kFull, ca, 0);
      int lenk = __shfl_sync(kFull, len, 0);
      float vk = __shfl_sync(kFull, vic, 0);
      int32_t j[kJE];
      float m[kJE];
#pragma unroll
      for (int u = 0; u < kJE; ++u) {
        const int e = u * 32 + lane;
        j[u] = -1;
        m[u] = 0.f;
        if (e < lenk) {
          j[u] = C_idx[cak + e];
          m[u] = fminf(vk, C_val[cak + e]);
        }
      }
      while (k < ncol) {
        int nk = k, nbase = base + 32 * kJE;
        int64_t ncak = cak;
        int nlen = lenk;
        float nv = vk;
        if (nbase >= lenk) {                                   // warp-uniform: next column
          nk = k + 1;
          nbase = 0;
          if (nk < ncol) {
            ncak = __shfl_sync(kFull, ca, nk);
            nlen = __shfl_sync(kFull, len, nk);
            nv = __shfl_sync(kFull, vic, nk);
          }
        }
        int32_t jn[kJE];
        float mn[kJE];
#pragma unroll
        for (int u = 0; u < kJE; ++u) {
          const int e = nbase + u * 32 + lane;
          jn[u] = -1;
          mn[u] = 0.f;
          if (nk < ncol && e < nlen) {
            jn[u] = C_idx[ncak + e];
            mn[u] = fminf(nv, C_val[ncak + e]);
          }
        }
        int fresh_n = 0;
#pragma unroll
        for (int u = 0; u < kJE; ++u) {
          bool fresh = false;
          if (j[u] >= 0 && (!owned_only || pair_owned((int32_t)row, j[u]))) {
            uint32_t h = jhash((uint32_t)j[u]) & smask;
            while (true) {
              const int32_t old = atomicCAS(&tkey[h], -1, j[u]);
              if (old == -1) {
                tval[h] = m[u];                                // 0 + m
                fresh = true;
                break;
              }
              if (old == j[u]) {
                tval[h] = acc_add(tval[h], m[u], half);
                break;
              }
              h = (h + 1) & smask;
            }
          }
          fresh_n += __popc(__ballot_sync(kFull, fresh));
        }
        used += fresh_n;
        __syncwarp();                                          // this step's adds land before the next step's
        if (used > limit) {
          overflow = true;
          break;
        }
        k = nk;
        base = nbase;
        cak = ncak;
        lenk = nlen;
        vk = nv;
#pragma unroll
        for (int u = 0; u < kJE; ++u) {
          j[u] = jn[u];
          m[u] = mn[u];
        }
      }
    }
    if (overflow) {
      if (lane == 0) {
        if (queues) {
          const int t = T_cnt[lr];
          const int need = t < 2048 ? t >> 2 : t >> 1;
          int c2 = 0;
          while (c2 < kJClasses && need > ((jclass_slots(c2) >> 1) + (jclass_slots(c2) >> 2))) ++c2;
          if (c2 <= cur_class) c2 = cur_class + 1;
          if (c2 >= direct_from) c2 = direct_from <= kJClasses ? kJClasses + 1 : (c2 > kJClasses ? kJClasses : c2);
          queues[(int64_t)c2 * q_stride + atomicAdd(&qlen[c2], 1)] = (int32_t)lr;
          if (escalated) atomicAdd(escalated, 1ull);
        } else if (next_queue) {
          next_queue[atomicAdd(next_len, 1)] = (int32_t)lr;
        } else {
          nbr_cnt[lr] = -1;
        }
      }
      __syncwarp();
      continue;
    }
    const int64_t o = slot_ptr[lr];
    const int64_t o_end = min(slot_ptr[lr + 1], nbr_capacity);   // the row's slots (speculative sizes are never trusted)
    int cnt = 0;
    for (int base = 0; base < slots; base += 32) {
      const int32_t j = tkey[base + lane];
      float jd = 2.f;
      if (j >= 0) jd = jaccard_from_t(tval[base + lane], half);
      const bool keep = j >= 0 && jd <= eps;
      const unsigned b = __ballot_sync(kFull, keep);
      if (keep) {
        const int64_t dst = o + cnt + __popc(b & lt);
        if (dst < o_end) {
          nbr_idx[dst] = j;
          if (nbr_val) nbr_val[dst] = jd;
        }
      }
      cnt += __popc(b);
    }
    if (lane == 0) {
      // a row that did not fit keeps only what was stored (consumers never read past the slots) and is reported
      const int64_t room = o_end > o ? o_end - o : 0;
      nbr_cnt[lr] = cnt <= room ? cnt : (int)room;
      if (cnt > room && slot_overflow) atomicAdd(slot_overflow, 1ull);
    }
    __syncwarp();
  }
}

// Rows with thousands of partners (k1 larger than the identity clusters: every row overlaps a sizeable part of the
// set) make the hash tables big and their probing long.  When N floats fit a few times into shared memory the
// accumulator is simply the dense row t[0..N) itself, owned by a whole CTA: the columns are still consumed one after
// the other (the reference's accumulation order), the entries of a column by all threads at once (a column holds
// every j once, so no two threads meet), with the next column's loads in flight.  An update is one load-add-store,
// nothing can overflow, and the ordered emit pass clears the row on the way.
constexpr int kJDThreads = 128;
constexpr int kJDPF = 4;                     // columns whose loads are in flight ahead of the accumulation

__global__ void __launch_bounds__(kJDThreads) jaccard_direct_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val, int64_t N,
    int64_t row_begin, const int32_t* __restrict__ queue, const int32_t* __restrict__ queue_len, float eps,
    const int64_t* __restrict__ slot_ptr, int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_val,
    int32_t* __restrict__ nbr_cnt, int64_t nbr_capacity, unsigned long long* __restrict__ slot_overflow, int half,
    int owned_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* acc = reinterpret_cast<float*>(smem_raw);
  __shared__ int64_t s_ca[32];
  __shared__ int s_len[32];
  __shared__ float s_v[32];
  __shared__ int s_warp[kJDThreads / 32];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int64_t n_pad = (N + kJDThreads - 1) / kJDThreads * kJDThreads;
  const int64_t n_rows = *queue_len;
  for (int64_t s0 = t; s0 < n_pad; s0 += kJDThreads) acc[s0] = 0.f;
  __syncthreads();
  for (int64_t li = blockIdx.x; li < n_rows; li += gridDim.x) {
    const int64_t lr = queue[li];
    const int64_t row = row_begin + lr;
    const int64_t qa = Q_ptr[row], qb = Q_ptr[row + 1];
    for (int64_t pc = qa; pc < qb; pc += 32) {
      const int ncol = (int)min((int64_t)32, qb - pc);
      if (t < ncol) {
        const int32_t c = Q_idx[pc + t];
        const int64_t ca = C_ptr[c];
        s_ca[t] = ca;
        s_len[t] = (int)(C_ptr[c + 1] - ca);
        s_v[t] = Q_val[pc + t];
      }
      __syncthreads();
      // The first kJDThreads entries of the next kJDPF columns are requested ahead of time: a column step is then one
      // shared-memory add and a barrier instead of a round trip to L2 (the kernel was bound by that latency).
      int32_t jq[kJDPF];
      float mq[kJDPF];
#pragma unroll
      for (int u = 0; u < kJDPF; ++u) {
        jq[u] = -1;
        mq[u] = 0.f;
        if (u < ncol && t < s_len[u]) {
          jq[u] = C_idx[s_ca[u] + t];
          mq[u] = fminf(s_v[u], C_val[s_ca[u] + t]);
        }
      }
      for (int k = 0; k < ncol; ++k) {
        const int32_t j = jq[0];
        const float m = mq[0];
#pragma unroll
        for (int u = 0; u + 1 < kJDPF; ++u) {
          jq[u] = jq[u + 1];
          mq[u] = mq[u + 1];
        }
        jq[kJDPF - 1] = -1;
        mq[kJDPF - 1] = 0.f;
        if (k + kJDPF < ncol && t < s_len[k + kJDPF]) {
          jq[kJDPF - 1] = C_idx[s_ca[k + kJDPF] + t];
          mq[kJDPF - 1] = fminf(s_v[k + kJDPF], C_val[s_ca[k + kJDPF] + t]);
        }
        if (j >= 0 && (!owned_only || pair_owned((int32_t)row, j))) acc[j] = acc_add(acc[j], m, half);
        const int len = s_len[k];
        for (int e = t + kJDThreads; e < len; e += kJDThreads) {   // columns longer than the CTA (rare)
          const int32_t j2 = C_idx[s_ca[k] + e];
          if (!owned_only || pair_owned((int32_t)row, j2)) acc[j2] = acc_add(acc[j2], fminf(s_v[k], C_val[s_ca[k] + e]), half);
        }
        __syncthreads();                                       // column k is in before column k + 1 starts
      }
    }
    // emit { j : J <= eps }; the accumulator is cleared on the way.  The order of a row's list does not matter to any
    // consumer (union-find, degree counts, min over adjacent labels; the tests sort), so every thread appends its own
    // finds through a shared-memory counter: no ballots and no barriers inside the sweep over the N slots, which was
    // more than half of this kernel's instructions when it compacted in order.
    const int64_t o = slot_ptr[lr];
    const int64_t o_end = min(slot_ptr[lr + 1], nbr_capacity);
    const int64_t room = o_end > o ? o_end - o : 0;
    if (t == 0) s_base = 0;
    __syncthreads();
    for (int64_t s0 = t; s0 < n_pad; s0 += kJDThreads) {
      const float tv = acc[s0];
      if (tv > 0.f) {
        acc[s0] = 0.f;
        const float jd = jaccard_from_t(tv, half);
        if (jd <= eps) {
          const int pos = atomicAdd(&s_base, 1);
          if (pos < room) {
            nbr_idx[o + pos] = (int32_t)s0;
            if (nbr_val) nbr_val[o + pos] = jd;
          }
        }
      }
    }
    __syncthreads();
    if (t == 0) {
      const int found = s_base;
      nbr_cnt[lr] = found <= room ? found : (int)room;
      if (found > room && slot_overflow) atomicAdd(slot_overflow, 1ull);
    }
  }
}

// T_cnt (upper bound of the partner count) -> table class.  T counts every (column, row) pair; the number of
// DISTINCT partners is a small fraction of it (a partner shares many columns with the row), and the latency-bound
// table kernel lives on occupancy, so the first guess is optimistic -- the class that holds T/4 -- and rows that
// do overflow move up one class at a time.
// direct_from: rows that would need class >= direct_from go to the direct-indexed kernel instead (queue kJClasses + 1);
// kJClasses + 1 when that kernel is not available for this N.
__global__ void __launch_bounds__(256) jaccard_classify_kernel(const int32_t* __restrict__ T_cnt,
                                                               const int32_t* __restrict__ P_cnt, int64_t n_rows,
                                                               int direct_from, int32_t* __restrict__ queues,
                                                               int32_t* __restrict__ qlen) {
  const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int c = -1;
  if (lr < n_rows) {
    const int t = T_cnt[lr];
    // a retry of a long row is expensive: less optimism there; P_cnt (reid_jaccard_bounds) refines the guess
    const int need = P_cnt ? P_cnt[lr] : (t < 2048 ? t >> 2 : t >> 1);
    c = 0;
    while (c < kJClasses && need > ((jclass_slots(c) >> 1) + (jclass_slots(c) >> 2))) ++c;   // c == kJClasses: heavy
    if (c >= direct_from) c = kJClasses + 1;
  }
  // one atomic per (warp, class) instead of one per row
  const int lane = lane_id();
  for (int q = 0; q <= kJClasses + 1; ++q) {
    const unsigned m = __ballot_sync(kFull, c == q);
    if (!m) continue;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(&qlen[q], __popc(m));
    base = __shfl_sync(kFull, base, __ffs(m) - 1);
    if (c == q) queues[(int64_t)q * n_rows + base + __popc(m & ((1u << lane) - 1u))] = (int32_t)lr;
  }
}

// Dense rows for the drop-in return value.  One CTA per row; the accumulator row lives in
// shared memory when N floats fit, else in the output row itself (L2-resident).
template <bool kSmemAcc>
__global__ void __launch_bounds__(256) jaccard_dense_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val, int64_t N,
    int64_t row_begin, float* __restrict__ out, int64_t ld, int half) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t lr = blockIdx.x;
  const int64_t row = row_begin + lr;
  float* orow = out + lr * ld;
  float* acc = kSmemAcc ? reinterpret_cast<float*>(smem_raw) : orow;
  for (int64_t j = threadIdx.x; j < N; j += blockDim.x) acc[j] = 0.f;
  __syncthreads();
  const int64_t qa = Q_ptr[row], qb = Q_ptr[row + 1];
  for (int64_t p = qa; p < qb; ++p) {
    const int32_t c = Q_idx[p];
    const float vic = Q_val[p];
    for (int64_t q = C_ptr[c] + threadIdx.x; q < C_ptr[c + 1]; q += blockDim.x) {
      const int32_t j = C_idx[q];
      const float m = fminf(vic, C_val[q]);
      if (kSmemAcc) {
        acc[j] = acc_add(acc[j], m, half);
      } else {
        __stcg(&acc[j], acc_add(__ldcg(&acc[j]), m, half));
      }
    }
    __syncthreads();
  }
  for (int64_t j = threadIdx.x; j < N; j += blockDim.x) {
    const float t = kSmemAcc ? acc[j] : __ldcg(&acc[j]);
    orow[j] = jaccard_from_t(t, half);
  }
}

// Rows whose partner set does not fit any shared-memory table ("hub" rows): one CTA per row, dense
// accumulator row in global scratch (L2-resident), same ascending-column order, then an ordered
// block-wide compaction of { j : J_ij <= eps }.
__global__ void __launch_bounds__(256) jaccard_neighbors_heavy_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val, int64_t N,
    int64_t row_begin, const int32_t* __restrict__ rows_list, int64_t n_list_host,
    const int32_t* __restrict__ list_len, float eps, const int64_t* __restrict__ slot_ptr,
    int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_val, int32_t* __restrict__ nbr_cnt,
    float* __restrict__ scratch, int64_t nbr_capacity, unsigned long long* __restrict__ slot_overflow, int half,
    int owned_only) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int64_t n_list = list_len ? (int64_t)*list_len : n_list_host;
  float* acc = scratch + (int64_t)blockIdx.x * N;
  for (int64_t li = blockIdx.x; li < n_list; li += gridDim.x) {
  const int64_t lr = rows_list[li];
  const int64_t row = row_begin + lr;
  __syncthreads();
  for (int64_t j = threadIdx.x; j < N; j += blockDim.x) __stcg(&acc[j], 0.f);
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int64_t p = Q_ptr[row]; p < Q_ptr[row + 1]; ++p) {
    const int32_t c = Q_idx[p];
    const float vic = Q_val[p];
    for (int64_t q = C_ptr[c] + threadIdx.x; q < C_ptr[c + 1]; q += blockDim.x) {
      const int32_t j = C_idx[q];
      if (!owned_only || pair_owned((int32_t)row, j)) __stcg(&acc[j], acc_add(__ldcg(&acc[j]), fminf(vic, C_val[q]), half));
    }
    __syncthreads();
  }
  const int64_t o = slot_ptr[lr];
  const int64_t o_end = min(slot_ptr[lr + 1], nbr_capacity);
  const int lane = lane_id(), w = threadIdx.x >> 5;
  for (int64_t base = 0; base < N; base += blockDim.x) {
    const int64_t j = base + threadIdx.x;
    float jd = 2.f;
    if (j < N) jd = jaccard_from_t(__ldcg(&acc[j]), half);
    const bool keep = j < N && jd <= eps;
    const unsigned b = __ballot_sync(kFull, keep);
    if (lane == 0) s_warp[w] = __popc(b);
    __syncthreads();
    int before = s_base;
    for (int ww = 0; ww < w; ++ww) before += s_warp[ww];
    if (keep) {
      const int64_t dst = o + before + __popc(b & ((1u << lane) - 1u));
      if (dst < o_end) {
        nbr_idx[dst] = (int32_t)j;
        if (nbr_val) nbr_val[dst] = jd;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int ww = 0; ww < 8; ++ww) tot += s_warp[ww];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t room = o_end > o ? o_end - o : 0;
    nbr_cnt[lr] = s_base <= room ? s_base : (int)room;
    if (s_base > room && slot_overflow) atomicAdd(slot_overflow, 1ull);
  }
  }
}

}  // namespace reid

namespace reid {
struct JnArgs {
  const int64_t* Q_ptr; const int32_t* Q_idx; const float* Q_val;
  const int64_t* C_ptr; const int32_t* C_idx; const float* C_val;
  int64_t row_begin; float eps; const int64_t* slot_ptr; int32_t* nbr_idx; float* nbr_val; int32_t* nbr_cnt;
  int64_t nbr_capacity; unsigned long long* slot_overflow; int half;
  // escalation of a row that overflowed its table (reid_jaccard_eps_graph only; all NULL / 0 otherwise): the row goes to
  // the class its T-based (pessimistic) size asks for -- at least one class up -- so a wrong optimistic first guess
  // costs one attempt, not a climb through every class
  const int32_t* T_cnt; int32_t* queues; int32_t* qlen; int64_t q_stride; int cur_class; int direct_from;
  int owned_only; unsigned long long* escalated;
};

// one launch of the table kernel: `n_max` bounds the grid, the real row count is *queue_len when given
template <int kWarps>
static int launch_jn(const JnArgs& a, int slots, int64_t n_max, const int32_t* queue, const int32_t* queue_len,
                     int32_t* next_queue, int32_t* next_len, cudaStream_t st) {
  const size_t smem = (size_t)kWarps * slots * 8;
  REID_CUDA(cudaFuncSetAttribute(jaccard_neighbors_kernel<kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;      // persistent grid = exactly the resident CTAs (a later wave would only add a tail)
  REID_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jaccard_neighbors_kernel<kWarps>, kWarps * 32, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (n_max + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (grid > cap) grid = cap;
  jaccard_neighbors_kernel<kWarps><<<(unsigned)grid, kWarps * 32, smem, st>>>(
      a.Q_ptr, a.Q_idx, a.Q_val, a.C_ptr, a.C_idx, a.C_val, a.row_begin, n_max, queue, queue_len, next_queue, next_len,
      a.eps, a.slot_ptr, a.nbr_idx, a.nbr_val, a.nbr_cnt, slots, a.nbr_capacity, a.slot_overflow, a.half, a.T_cnt,
      a.queues, a.qlen, a.q_stride, a.cur_class, a.direct_from, a.owned_only, a.escalated);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

static int launch_jn_slots(const JnArgs& a, int slots, int64_t n_max, const int32_t* queue, const int32_t* queue_len,
                           int32_t* next_queue, int32_t* next_len, cudaStream_t st) {
  if (slots <= 1024) return launch_jn<8>(a, slots, n_max, queue, queue_len, next_queue, next_len, st);
  if (slots <= 2048) return launch_jn<4>(a, slots, n_max, queue, queue_len, next_queue, next_len, st);
  if (slots <= 4096) return launch_jn<6>(a, slots, n_max, queue, queue_len, next_queue, next_len, st);
  if (slots <= 8192) return launch_jn<3>(a, slots, n_max, queue, queue_len, next_queue, next_len, st);
  return launch_jn<1>(a, slots, n_max, queue, queue_len, next_queue, next_len, st);
}

struct JWs {
  int32_t* qlen;     // kJClasses + 1 (+ padding)
  int32_t* queues;   // (kJClasses + 2) x n_rows: table classes, heavy, direct
  float* scratch;    // kJHeavyCtas x N
};
static size_t jws_carve(void* base, int64_t N, int64_t n, JWs* w) {
  unsigned char* p = (unsigned char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += (bytes + 255) / 256 * 256;
    return r;
  };
  JWs t;
  t.qlen = (int32_t*)take(sizeof(int32_t) * 16);
  t.queues = (int32_t*)take(sizeof(int32_t) * (size_t)(kJClasses + 2) * (size_t)(n > 0 ? n : 1));
  t.scratch = (float*)take(sizeof(float) * (size_t)kJHeavyCtas * (size_t)N);
  if (w) *w = t;
  return off;
}
}  // namespace reid

extern "C" {

int reid_jaccard_bounds(const int64_t* Q_ptr, const int32_t* Q_idx, const float* Q_val, const int64_t* C_ptr,
                        int64_t row_begin, int64_t row_end, float eps, int32_t* T_cnt, int32_t* S_cnt, int32_t* P_cnt,
                        void* stream) {
  using namespace reid;
  REID_CHECK_ARG(Q_ptr && Q_idx && C_ptr && T_cnt, "reid_jaccard_bounds: NULL pointer");
  REID_CHECK_ARG(!S_cnt || Q_val, "reid_jaccard_bounds: S_cnt needs Q_val");
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end, "reid_jaccard_bounds: bad row range");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  // J(t) = 1 - t / (2 - t) <= eps  <=>  t >= 2 (1 - eps) / (2 - eps); 0.1 % head-room for the fp32 roundings of J
  float t_min = 0.f;
  if (eps < 1.f) t_min = (float)(2.0 * (1.0 - (double)eps) / (2.0 - (double)eps) * 0.999);
  jaccard_bounds_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(Q_ptr, Q_idx, Q_val, C_ptr, row_begin,
                                                                                  row_end, t_min, T_cnt, S_cnt, P_cnt);
  REID_LAUNCH_CHECK();
  return REID_OK;
}


int reid_jaccard_neighbors(const int64_t* Q_ptr, const int32_t* Q_idx, const float* Q_val, const int64_t* C_ptr,
                           const int32_t* C_idx, const float* C_val, int64_t N, int64_t row_begin, int64_t row_end,
                           const int32_t* rows_list, int64_t n_list, float eps, const int64_t* slot_ptr,
                           int32_t* nbr_idx, float* nbr_val, int32_t* nbr_cnt, int table_slots, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(Q_ptr && Q_idx && Q_val && C_ptr && C_idx && C_val && slot_ptr && nbr_idx && nbr_cnt,
                 "reid_jaccard_neighbors: NULL pointer");
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_jaccard_neighbors: bad row range");
  REID_CHECK_ARG(table_slots >= 64 && (table_slots & (table_slots - 1)) == 0,
                 "reid_jaccard_neighbors: table_slots=%d must be a power of two >= 64", table_slots);
  REID_CHECK_ARG((size_t)table_slots * 8 <= 224 * 1024, "reid_jaccard_neighbors: table_slots=%d does not fit shared memory",
                 table_slots);
  const int64_t n = rows_list ? n_list : row_end - row_begin;
  if (n == 0) return REID_OK;
  const JnArgs a{Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, row_begin, eps, slot_ptr, nbr_idx, nbr_val, nbr_cnt,
                 INT64_MAX, nullptr, 0, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr};
  return launch_jn_slots(a, table_slots, n, rows_list, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

size_t reid_jaccard_eps_graph_workspace_bytes(int64_t N, int64_t n_rows) {
  if (N < 0 || n_rows < 0) return 0;
  return reid::jws_carve(nullptr, N, n_rows, nullptr);
}

int reid_jaccard_eps_graph(const int64_t* Q_ptr, const int32_t* Q_idx, const float* Q_val, const int64_t* C_ptr,
                           const int32_t* C_idx, const float* C_val, int64_t N, int64_t row_begin, int64_t row_end,
                           float eps, const int32_t* T_cnt, const int32_t* P_cnt, const int64_t* slot_ptr, int32_t* nbr_idx, float* nbr_val,
                           int32_t* nbr_cnt, int64_t nbr_capacity, uint64_t* slot_overflow, int half_precision,
                           int owned_pairs_only, uint64_t* escalated_rows, void* workspace, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(Q_ptr && Q_idx && Q_val && C_ptr && C_idx && C_val && T_cnt && slot_ptr && nbr_idx && nbr_cnt && workspace,
                 "reid_jaccard_eps_graph: NULL pointer");
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_jaccard_eps_graph: bad row range");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JWs w;
  jws_carve(workspace, N, n, &w);
  REID_CUDA(cudaMemsetAsync(w.qlen, 0, sizeof(int32_t) * 16, st));
  // direct-indexed accumulator rows (N floats each) when at least two fit into shared memory: they take over from
  // the 4096-slot class up, and the overflow of the last hash class below
  const size_t row_bytes = (size_t)((N + kJDThreads - 1) / kJDThreads * kJDThreads) * sizeof(float);
  const int direct_from = row_bytes * 2 <= 220u * 1024u ? 3 : kJClasses + 1;
  int32_t* direct_q = w.queues + (int64_t)(kJClasses + 1) * n;
  int32_t* direct_len = w.qlen + kJClasses + 1;
  jaccard_classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T_cnt, P_cnt, n, direct_from, w.queues, w.qlen);
  REID_LAUNCH_CHECK();
  if (nbr_capacity <= 0) nbr_capacity = INT64_MAX;           // slots sized by the caller from T_cnt: nothing to guard
  unsigned long long* ovf = (unsigned long long*)slot_overflow;
  JnArgs a{Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, row_begin, eps, slot_ptr, nbr_idx, nbr_val, nbr_cnt,
           nbr_capacity, ovf, half_precision, T_cnt, w.queues, w.qlen, n, 0, direct_from, owned_pairs_only,
           (unsigned long long*)escalated_rows};
  const int n_hash = direct_from < kJClasses ? direct_from : kJClasses;
  for (int c = 0; c < n_hash; ++c) {
    a.cur_class = c;
    int rc = launch_jn_slots(a, jclass_slots(c), n, w.queues + (int64_t)c * n, w.qlen + c, nullptr, nullptr, st);
    if (rc != REID_OK) return rc;
  }
  if (direct_from <= kJClasses) {
    REID_CUDA(cudaFuncSetAttribute(jaccard_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
    int per_sm = 1;
    REID_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jaccard_direct_kernel, kJDThreads, row_bytes));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > n) grid = n;
    jaccard_direct_kernel<<<(unsigned)grid, kJDThreads, row_bytes, st>>>(Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin,
                                                                        direct_q, direct_len, eps, slot_ptr, nbr_idx, nbr_val,
                                                                        nbr_cnt, nbr_capacity, ovf, half_precision,
                                                                        owned_pairs_only);
    REID_LAUNCH_CHECK();
  }
  const int64_t hg = n < kJHeavyCtas ? n : kJHeavyCtas;
  jaccard_neighbors_heavy_kernel<<<(unsigned)hg, 256, 0, st>>>(Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin,
                                                              w.queues + (int64_t)kJClasses * n, 0, w.qlen + kJClasses, eps,
                                                              slot_ptr, nbr_idx, nbr_val, nbr_cnt, w.scratch, nbr_capacity, ovf, half_precision,
                                                              owned_pairs_only);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_jaccard_dense(const int64_t* Q_ptr, const int32_t* Q_idx, const float* Q_val, const int64_t* C_ptr,
                       const int32_t* C_idx, const float* C_val, int64_t N, int64_t row_begin, int64_t row_end,
                       float* out, int64_t ld, int half_precision, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(Q_ptr && Q_idx && Q_val && C_ptr && C_idx && C_val && out, "reid_jaccard_dense: NULL pointer");
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= N && ld >= N, "reid_jaccard_dense: bad shape");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  const size_t smem = (size_t)N * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (smem <= 200 * 1024) {
    REID_CUDA(cudaFuncSetAttribute(jaccard_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jaccard_dense_kernel<true><<<(unsigned)n, 256, smem, st>>>(Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin,
                                                              out, ld, half_precision);
  } else {
    jaccard_dense_kernel<false><<<(unsigned)n, 256, 0, st>>>(Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin,
                                                            out, ld, half_precision);
  }
  REID_LAUNCH_CHECK();
  return REID_OK;
}
int reid_jaccard_neighbors_heavy(const int64_t* Q_ptr, const int32_t* Q_idx, const float* Q_val, const int64_t* C_ptr,
                                 const int32_t* C_idx, const float* C_val, int64_t N, int64_t row_begin,
                                 const int32_t* rows_list, int64_t n_list, float eps, const int64_t* slot_ptr,
                                 int32_t* nbr_idx, float* nbr_val, int32_t* nbr_cnt, float* scratch, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(Q_ptr && Q_idx && Q_val && C_ptr && C_idx && C_val && rows_list && slot_ptr && nbr_idx && nbr_cnt &&
                     scratch,
                 "reid_jaccard_neighbors_heavy: NULL pointer");
  REID_CHECK_ARG(N > 0 && n_list >= 0, "reid_jaccard_neighbors_heavy: bad shape");
  if (n_list == 0) return REID_OK;
  jaccard_neighbors_heavy_kernel<<<(unsigned)n_list, 256, 0, (cudaStream_t)stream>>>(
      Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin, rows_list, n_list, nullptr, eps, slot_ptr, nbr_idx, nbr_val,
      nbr_cnt, scratch, INT64_MAX, nullptr, 0, 0);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
