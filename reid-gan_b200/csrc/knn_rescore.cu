// a1 (second half) -- exact re-score of tensor-core candidates, certificate, final ordering.
//
// The fp16 GEMM (simgemm_tc.cu) gives approximate scores a_j with |a_j - s_j| <= eps for the exact
// canonical key s_j.  Let a_(k) be the k-th largest approximate score of a row.  Every member of the
// exact top-k has s >= (k-th largest s) >= a_(k) - eps, hence a >= a_(k) - 2 eps: the exact top-k lies
// inside the WINDOW { j : a_j >= a_(k) - 2 eps }.  The GEMM epilogue retains the kc best approximate
// scores; if the weakest retained one is already below the window, the window is entirely among the
// candidates and the row is CERTIFIED.  Window members are then re-scored with the canonical key
// fp32(fp64 dot) and ordered by (key desc, index asc) -- bit-identical to reid_knn_exact.  The
// measured |a - s| is audited against eps; any violation un-certifies the row.  Uncertified rows are
// redone by the caller with reid_knn_exact, so the result never depends on eps being right.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace reid {

constexpr int kRsWarps = 4;
constexpr int kRsMaxC = 2 * REID_TC_MAX_SPLITS * REID_TC_KEEP_MAX;  // 512 candidates per row held at most
constexpr int kRsPer = kRsMaxC / 32;
constexpr int kWinMax = 128;                                         // window members per row at most

// |fp16-GEMM score - exact dot| <= 2^-10 ||x_i|| ||x_j||  (both operands rounded to 11 significant bits,
// Cauchy-Schwarz) with 2% head-room, + 2^-14 for the fp32 accumulation.  max_sqnorm = max_i ||x_i||^2.
__host__ __device__ inline float reid_tc_err_bound(float max_sqnorm) {
  return max_sqnorm * (1.02f / 1024.0f) + (1.0f / 16384.0f);
}

// ---- stage 1: selection (no feature traffic) ----------------------------------------------------
// Sweeps the row's candidate lists, finds the k-th best approximate score, checks the certificate and
// writes the window members.  It also emits a locality key for stage 2: the smallest column index among
// the members that score at least half of the best non-self score -- for a row inside an identity cluster
// that is the cluster's smallest index, so cluster mates (who share their candidates) get the same key.
__global__ void __launch_bounds__(kRsWarps * 32) rescore_select_kernel(
    int64_t row_begin, int64_t n_rows, const unsigned long long* __restrict__ cand,
    const int32_t* __restrict__ cand_cnt, const uint32_t* __restrict__ row_tau, int n_lists, int list_cap,
    int64_t list_pitch_rows, int k, float eps_in, const float* __restrict__ max_sqnorm, int32_t* __restrict__ win_cnt, int32_t* __restrict__ win_idx, float* __restrict__ win_a,
    int32_t* __restrict__ uncert, int32_t* __restrict__ key_out, int32_t* __restrict__ hist,
    const int32_t* __restrict__ row_pos, const int32_t* __restrict__ col_orig) {
  // row_pos / col_orig (sample-first symmetric search): the lists live in a permuted row space -- row lr's lists and
  // threshold are those of position row_pos[lr], and a listed column p is the original row col_orig[p].
  __shared__ float s_a[kRsWarps][kRsMaxC];
  __shared__ int32_t s_j[kRsWarps][kRsMaxC];
  const int w = threadIdx.x >> 5, lane = lane_id();
  const int64_t lr = (int64_t)blockIdx.x * kRsWarps + w;
  if (lr >= n_rows) return;
  const int32_t self_orig = (int32_t)(row_begin + lr);
  const float eps = max_sqnorm ? reid_tc_err_bound(*max_sqnorm) : eps_in;
  float* a = s_a[w];
  int32_t* jj = s_j[w];

  // Every column outside the row's lists scored <= bound.  The window [a_(k) - 2 eps, inf) starts at or
  // above bound - 2 eps (a_(k) >= bound because at least `keep` >= k listed columns score >= bound), so a
  // sweep that collects the entries >= bound - 2 eps finds every window member.
  const int64_t pl = row_pos ? (int64_t)row_pos[lr] : lr;
  const int32_t self = row_pos ? (int32_t)pl : self_orig;      // the row's own id in the id space of the lists
  const uint32_t tq = row_tau[pl];
  const float bound = tq ? ord_float(tq) : -INFINITY;

  auto collect = [&](float thr) {
    int cnt = 0;
    for (int q = 0; q < n_lists; ++q) {
      const int64_t li = list_pitch_rows ? q * list_pitch_rows + pl : pl * n_lists + q;   // list-major / row-major
      const int c = min(cand_cnt[li], list_cap);
      const unsigned long long* src = cand + li * (int64_t)list_cap;
      for (int base = 0; base < c; base += 32) {
        const int t = base + lane;
        unsigned long long e = 0;
        float v = -INFINITY;
        if (t < c) {
          e = src[t];
          v = __uint_as_float((uint32_t)(e >> 32));
        }
        const bool in = t < c && v >= thr;
        const unsigned b = __ballot_sync(kFull, in);
        const int pos = cnt + __popc(b & ((1u << lane) - 1u));
        if (in && pos < kRsMaxC) {
          a[pos] = v;
          jj[pos] = (int32_t)(uint32_t)e;                  // a POSITION when col_orig is given: translated on the way out
        }
        cnt += __popc(b);
      }
    }
    __syncwarp();
    return cnt;
  };
  // k-th largest of a[0..m): bitwise binary search on the order-preserving image, every lane holding m / 32 values in
  // registers.  Two savings against the plain 32-step / 16-register form (the kernel is ALU bound): only as many
  // registers as m needs take part (kPer = 4, 8, 12 or 16), and the lowest 8 bits are not resolved -- the result is
  // then at most 2^8 ulp (1.5e-5 on scores below 1) BELOW the true k-th largest, which can only widen the window
  // [a_k - 2 eps, inf) by a hair and make the certificate `bound < a_k - 2 eps` stricter: both safe.
  auto kth_impl = [&](int m, auto per_tag) {
    constexpr int kPer = decltype(per_tag)::value;
    uint32_t o[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int t = u * 32 + lane;
      o[u] = t < m ? float_ord(a[t]) : 0u;
    }
    uint32_t T = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 8; --bit) {
      const uint32_t c2 = T | (1u << bit);
      int c = 0;
#pragma unroll
      for (int u = 0; u < kPer; ++u) c += o[u] >= c2;
      c = __reduce_add_sync(kFull, c);
      if (c >= k) T = c2;
    }
    return m >= k ? ord_float(T) : -INFINITY;
  };
  auto kth_largest = [&](int m) {
    if (m <= 128) return kth_impl(m, std::integral_constant<int, 4>{});
    if (m <= 256) return kth_impl(m, std::integral_constant<int, 8>{});
    if (m <= 384) return kth_impl(m, std::integral_constant<int, 12>{});
    return kth_impl(m, std::integral_constant<int, kRsPer>{});
  };

  bool list_overflow = false;                       // an overflowed list lost columns above the threshold
  for (int q = 0; q < n_lists; ++q)
    list_overflow |= cand_cnt[list_pitch_rows ? q * list_pitch_rows + pl : pl * n_lists + q] > list_cap;
  float thr = bound - 2.0f * eps;
  int n = collect(thr);
  if (n > kRsMaxC) {
    // too many entries above the (loose) rejection threshold: the k-th best of the stored subset is a lower
    // bound of a_(k), so everything in the window scores >= that - 2 eps; sweep again with it
    thr = fmaxf(thr, kth_largest(kRsMaxC) - 2.0f * eps);
    n = collect(thr);
  }
  const bool overflow = n > kRsMaxC;
  if (overflow) n = kRsMaxC;
  const float a_k = kth_largest(n);
  const float lo = a_k - 2.0f * eps;
  bool certified = !overflow && !list_overflow && n >= k && bound < lo;

  // best non-self score -> locality threshold
  float best = -INFINITY;
  for (int t = lane; t < n; t += 32)
    if (jj[t] != self) best = fmaxf(best, a[t]);
  best = warp_max(best);
  const float near = 0.5f * best;
  // window members -> global lists
  int n_w = 0, kmin = 0x7fffffff;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    const int t = base + lane;
    const bool in = t < n && a[t] >= lo;
    const unsigned b = __ballot_sync(kFull, in);
    const int pos = n_w + __popc(b & ((1u << lane) - 1u));
    const int32_t j_orig = in ? (col_orig ? col_orig[jj[t]] : jj[t]) : 0;     // only window members are translated
    if (in && pos < kWinMax) {
      win_idx[lr * kWinMax + pos] = j_orig;
      win_a[lr * kWinMax + pos] = a[t];
    }
    if (in && a[t] >= near) kmin = min(kmin, j_orig);
    n_w += __popc(b);
  }
  kmin = warp_min(kmin);
  if (n_w > kWinMax || n_w < k) certified = false;
  if (lane == 0) {
    win_cnt[lr] = min(n_w, kWinMax);
    uncert[lr] = certified ? 0 : 1;
    if (key_out) {
      if (kmin == 0x7fffffff) kmin = self_orig;
      key_out[lr] = kmin;
      atomicAdd(&hist[kmin], 1);
    }
  }
}

__global__ void order_scatter_kernel(const int32_t* __restrict__ key, int64_t n_rows, const int64_t* __restrict__ ptr,
                                     int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lr >= n_rows) return;
  const int k = key[lr];
  perm[ptr[k] + atomicAdd(&cursor[k], 1)] = (int32_t)lr;
}

// ---- stage 2: exact keys of the window members, audit, final order --------------------------------
// HBM bound: ~36 feature rows of 4*D bytes per query row.  Rows are visited in `perm` order so that
// the rows of one identity cluster, which share most candidates, run back to back and hit in L2.
// kChunks > 0: D == kChunks * 128.  The query row is parked in shared memory (one float4 per lane and chunk,
// conflict-free), candidate rows are streamed half a row at a time through two register buffers so that the
// loads of the next half are in flight while the current one is multiplied and reduced (the kernel is
// latency bound: ~36 dependent gathers per query row).  kChunks == 0: generic D, both rows are streamed.
template <int kChunks>
__global__ void __launch_bounds__(kRsWarps * 32) rescore_exact_kernel(
    const float* __restrict__ x, int64_t D, int64_t row_begin, int64_t n_rows, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ win_cnt, const int32_t* __restrict__ win_idx, const float* __restrict__ win_a, int k,
    float eps_in, const float* __restrict__ max_sqnorm, int32_t* __restrict__ out_idx, float* __restrict__ out_key,
    int32_t* __restrict__ uncert, unsigned* __restrict__ max_err_bits, unsigned long long* __restrict__ uncert_count) {
  constexpr int kQ = kChunks > 0 ? kChunks : 1;
  constexpr int kH = kChunks > 1 ? kChunks / 2 : 1;
  __shared__ uint64_t s_key[kRsWarps][kWinMax];
  __shared__ float4 s_q[kRsWarps][kQ * 32];
  const int w = threadIdx.x >> 5, lane = lane_id();
  const int64_t slot = (int64_t)blockIdx.x * kRsWarps + w;
  if (slot >= n_rows) return;
  const int64_t lr = perm ? perm[slot] : slot;
  const int64_t row = row_begin + lr;
  uint64_t* key = s_key[w];
  const float eps = max_sqnorm ? reid_tc_err_bound(*max_sqnorm) : eps_in;
  const int n_w = win_cnt[lr];
  const float* xi = x + row * D;
  const int32_t* wj = win_idx + lr * kWinMax;
  float worst = 0.f;

  if (kChunks > 1) {
    float4* q = s_q[w];
#pragma unroll
    for (int c = 0; c < kQ; ++c) q[c * 32 + lane] = reinterpret_cast<const float4*>(xi)[c * 32 + lane];
    float4 ra[kH], rb[kH];
    if (n_w > 0) {
      const float4* xj = reinterpret_cast<const float4*>(x + (int64_t)wj[0] * D);
#pragma unroll
      for (int c = 0; c < kH; ++c) ra[c] = xj[c * 32 + lane];
    }
    for (int wi = 0; wi < n_w; ++wi) {
      const int32_t j = wj[wi];
      const float4* xj = reinterpret_cast<const float4*>(x + (int64_t)j * D);
#pragma unroll
      for (int c = 0; c < kH; ++c) rb[c] = xj[(kH + c) * 32 + lane];
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int c = 0; c < kH; ++c) {
        const float4 qq = q[c * 32 + lane];
        acc0 = fma((double)qq.x, (double)ra[c].x, acc0);
        acc1 = fma((double)qq.y, (double)ra[c].y, acc1);
        acc2 = fma((double)qq.z, (double)ra[c].z, acc2);
        acc3 = fma((double)qq.w, (double)ra[c].w, acc3);
      }
      if (wi + 1 < n_w) {
        const float4* xn = reinterpret_cast<const float4*>(x + (int64_t)wj[wi + 1] * D);
#pragma unroll
        for (int c = 0; c < kH; ++c) ra[c] = xn[c * 32 + lane];
      }
#pragma unroll
      for (int c = 0; c < kH; ++c) {
        const float4 qq = q[(kH + c) * 32 + lane];
        acc0 = fma((double)qq.x, (double)rb[c].x, acc0);
        acc1 = fma((double)qq.y, (double)rb[c].y, acc1);
        acc2 = fma((double)qq.z, (double)rb[c].z, acc2);
        acc3 = fma((double)qq.w, (double)rb[c].w, acc3);
      }
      const float s = (float)warp_sum((acc0 + acc1) + (acc2 + acc3));
      worst = fmaxf(worst, fabsf(s - win_a[lr * kWinMax + wi]));
      if (lane == 0) key[wi] = sel_key(s, j);
    }
  } else {
    for (int wi = 0; wi < n_w; ++wi) {
      const int32_t j = wj[wi];
      const float* xj = x + (int64_t)j * D;
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
      if ((D & 3) == 0 && ((((uintptr_t)x) & 15) == 0)) {
        const float4* a4 = reinterpret_cast<const float4*>(xi);
        const float4* b4 = reinterpret_cast<const float4*>(xj);
        for (int64_t d = lane; d < (D >> 2); d += 32) {
          const float4 p = a4[d], r = b4[d];
          acc0 = fma((double)p.x, (double)r.x, acc0);
          acc1 = fma((double)p.y, (double)r.y, acc1);
          acc2 = fma((double)p.z, (double)r.z, acc2);
          acc3 = fma((double)p.w, (double)r.w, acc3);
        }
      } else {
        for (int64_t d = lane; d < D; d += 32) acc0 = fma((double)xi[d], (double)xj[d], acc0);
      }
      const float s = (float)warp_sum((acc0 + acc1) + (acc2 + acc3));
      worst = fmaxf(worst, fabsf(s - win_a[lr * kWinMax + wi]));
      if (lane == 0) key[wi] = sel_key(s, j);
    }
  }
  __syncwarp();
  if (lane == 0) {
    atomicMax(max_err_bits, __float_as_uint(worst));
    if (!(worst <= eps)) uncert[lr] = 1;  // the error model was violated (or NaN): do not trust the window
    if (uncert_count && uncert[lr]) atomicAdd(uncert_count, 1ull);
  }
  // order by (key desc, idx asc); the first k go out
  for (int t = lane; t < n_w; t += 32) {
    const uint64_t me = key[t];
    int rank = 0;
    for (int u = 0; u < n_w; ++u) rank += key[u] > me;
    if (rank < k) {
      out_idx[lr * k + rank] = sel_key_idx(me);
      if (out_key) out_key[lr * k + rank] = sel_key_val(me);
    }
  }
  // a window with fewer than k members (the row is uncertified and will be redone): the unfilled positions still
  // get a VALID index, because a caller that defers the certificate check runs the next stages on these lists
  for (int t = n_w + lane; t < k; t += 32) {
    out_idx[lr * k + t] = (int32_t)row;
    if (out_key) out_key[lr * k + t] = -INFINITY;
  }
}

// ---- stage 2, grouped: a small FP64 GEMM per group of locality-ordered rows -------------------------
// rescore_exact_kernel is bound by the fp32 -> fp64 conversions (two per product: 4.8 G F2F at 16/clk/SM
// = 1.0 ms at N = 32,621).  Rows that are neighbours in the locality order are cluster mates and share
// most of their window, so a CTA takes kGq consecutive rows, builds the UNION of their windows in shared
// memory and computes the whole (kGq x |union|) block of exact dots: every candidate row is fetched and
// converted once per group instead of once per (row, candidate) pair, the query rows are converted once
// into shared memory, and the pairs that nobody asked for are simply not read back.
//   thread (cg, ks): 4 candidates x kGq queries over K-slice ks (float4 of every 64 dims) -> 16 fp64
//   accumulators; the 16 K-slices of a candidate group sit in one half warp and are folded with a
//   transpose-reduce (15 shuffles of a double instead of 64).  The summation order of a dot is fixed by
//   (ks, step) alone, so a key does not depend on which group or slot the pair landed in.
constexpr int kGq = 4;
constexpr int kGThreads = 256;
constexpr int kGChunk = 64;                 // candidates per pass = 16 groups of 4
constexpr int kGSlots = 1024;               // union hash table (|union| <= kGq * kWinMax = 512)

__global__ void __launch_bounds__(kGThreads, 2) rescore_group_kernel(
    const float* __restrict__ x, int64_t D, int64_t row_begin, int64_t n_rows, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ win_cnt, const int32_t* __restrict__ win_idx, const float* __restrict__ win_a, int k,
    float eps_in, const float* __restrict__ max_sqnorm, int32_t* __restrict__ out_idx, float* __restrict__ out_key,
    int32_t* __restrict__ uncert, unsigned* __restrict__ max_err_bits, unsigned long long* __restrict__ uncert_count, int dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sQ = reinterpret_cast<double*>(smem_raw);                                    // [kGq][D]
  uint64_t* s_key = reinterpret_cast<uint64_t*>(sQ + (size_t)kGq * D);                 // [kGq][kWinMax]
  double* s_dot = reinterpret_cast<double*>(s_key + kGq * kWinMax);                    // [kGq][kGChunk]
  int32_t* s_tab = reinterpret_cast<int32_t*>(s_dot + kGq * kGChunk);                  // [kGSlots] column ids
  int32_t* s_ulist = s_tab + kGSlots;                                                  // [kGq * kWinMax]
  uint16_t* s_tid = reinterpret_cast<uint16_t*>(s_ulist + kGq * kWinMax);              // [kGSlots] union position
  uint16_t* s_pid = s_tid + kGSlots;                                                   // [kGq][kWinMax]
  __shared__ int s_U;
  __shared__ int s_nw[kGq];
  __shared__ int64_t s_lr[kGq];
  __shared__ unsigned s_worst[kGq];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float eps = max_sqnorm ? reid_tc_err_bound(*max_sqnorm) : eps_in;

  if (tid < kGq) {
    const int64_t slot = (int64_t)blockIdx.x * kGq + tid;
    const int64_t lr = slot < n_rows ? (perm ? perm[slot] : slot) : -1;
    s_lr[tid] = lr;
    s_nw[tid] = lr >= 0 ? win_cnt[lr] : 0;
    s_worst[tid] = 0u;
  }
  if (tid == 0) s_U = 0;
  for (int h = tid; h < kGSlots; h += kGThreads) s_tab[h] = -1;
  __syncthreads();

  // query rows -> fp64 in shared memory (rows of an incomplete last group are zero)
#pragma unroll
  for (int q = 0; q < kGq; ++q) {
    const int64_t lr = s_lr[q];
    const float4* xi = reinterpret_cast<const float4*>(x + (row_begin + (lr >= 0 ? lr : 0)) * D);
    for (int64_t d4 = tid; d4 < (D >> 2); d4 += kGThreads) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lr >= 0) v = xi[d4];
      // layout [q][step][half][K-slice] in double2 units: the 16 K-slice lanes of a half warp read 256
      // contiguous bytes per (step, half) -- no bank conflicts
      double2* dst = reinterpret_cast<double2*>(sQ + (size_t)q * D) + (d4 >> 4) * 32 + (d4 & 15);
      dst[0] = make_double2((double)v.x, (double)v.y);
      dst[16] = make_double2((double)v.z, (double)v.w);
    }
  }
  // union of the windows
  for (int it = tid; it < kGq * kWinMax; it += kGThreads) {
    const int q = it / kWinMax, t = it % kWinMax;
    if (t < s_nw[q]) {
      const int32_t j = win_idx[s_lr[q] * kWinMax + t];
      uint32_t h = ((uint32_t)j * 0x9e3779b1u) >> 22;          // 10 bits
      while (true) {
        const int32_t old = atomicCAS(&s_tab[h], -1, j);
        if (old == -1) {
          const int id = atomicAdd(&s_U, 1);
          s_ulist[id] = j;
          s_tid[h] = (uint16_t)id;
          break;
        }
        if (old == j) break;
        h = (h + 1) & (kGSlots - 1);
      }
    }
  }
  __syncthreads();
  for (int it = tid; it < kGq * kWinMax; it += kGThreads) {
    const int q = it / kWinMax, t = it % kWinMax;
    if (t < s_nw[q]) {
      const int32_t j = win_idx[s_lr[q] * kWinMax + t];
      uint32_t h = ((uint32_t)j * 0x9e3779b1u) >> 22;
      while (s_tab[h] != j) h = (h + 1) & (kGSlots - 1);
      s_pid[it] = s_tid[h];
    }
  }
  __syncthreads();
  const int U = s_U;

  const int cg = tid >> 4, ks = tid & 15;
  const int steps = (dbg & 1) ? 0 : (int)(D >> 6);
  for (int c0 = 0; c0 < ((dbg & 2) ? 0 : U); c0 += kGChunk) {
    const float4* xc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = c0 + cg * 4 + i;
      xc[i] = reinterpret_cast<const float4*>(x + (int64_t)s_ulist[id < U ? id : 0] * D) + ks;
    }
    double acc[kGq * 4];
#pragma unroll
    for (int v = 0; v < kGq * 4; ++v) acc[v] = 0.0;
    if (c0 + cg * 4 < U) {                                   // candidate groups past the union have nothing to do
#pragma unroll 2
      for (int s = 0; s < steps; ++s) {
        float4 c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = xc[i][s * 16];
        double cd[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cd[i][0] = (double)c[i].x;
          cd[i][1] = (double)c[i].y;
          cd[i][2] = (double)c[i].z;
          cd[i][3] = (double)c[i].w;
        }
#pragma unroll
        for (int q = 0; q < kGq; ++q) {
          const double2* qp = reinterpret_cast<const double2*>(sQ + (size_t)q * D) + s * 32 + ks;
          const double2 q01 = qp[0], q23 = qp[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double a = acc[q * 4 + i];
            a = fma(q01.x, cd[i][0], a);
            a = fma(q01.y, cd[i][1], a);
            a = fma(q23.x, cd[i][2], a);
            a = fma(q23.y, cd[i][3], a);
            acc[q * 4 + i] = a;
          }
        }
      }
    }
    // fold the 16 K-slices: after the rounds lane ks holds the complete dot number ks (= q * 4 + i)
#pragma unroll
    for (int off = 8, n = 8; off >= 1; off >>= 1, n >>= 1) {
      const bool up = (ks & off) != 0;
#pragma unroll
      for (int m = 0; m < n; ++m) {
        const double send = up ? acc[m] : acc[m + n];
        const double keep = up ? acc[m + n] : acc[m];
        acc[m] = keep + __shfl_xor_sync(kFull, send, off);
      }
    }
    s_dot[(ks >> 2) * kGChunk + cg * 4 + (ks & 3)] = acc[0];
    __syncthreads();
    for (int it = tid; it < kGq * kWinMax; it += kGThreads) {
      const int q = it / kWinMax, t = it % kWinMax;
      if (t < s_nw[q]) {
        const int id = (int)s_pid[it] - c0;
        if (id >= 0 && id < kGChunk) {
          const int64_t lr = s_lr[q];
          const float sc = (float)s_dot[q * kGChunk + id];
          const float err = fabsf(sc - win_a[lr * kWinMax + t]);
          atomicMax(&s_worst[q], __float_as_uint(err == err ? err : INFINITY));
          s_key[it] = sel_key(sc, win_idx[lr * kWinMax + t]);
        }
      }
    }
    __syncthreads();
  }

  // per row: audit, then order by (key desc, idx asc); the first k go out
  if (warp < kGq && s_lr[warp] >= 0) {
    const int64_t lr = s_lr[warp];
    const int n_w = s_nw[warp];
    const uint64_t* key = s_key + warp * kWinMax;
    if (lane == 0) {
      const float worst = __uint_as_float(s_worst[warp]);
      atomicMax(max_err_bits, __float_as_uint(worst));
      if (!(worst <= eps)) uncert[lr] = 1;  // the error model was violated (or NaN): do not trust the window
      if (uncert_count && uncert[lr]) atomicAdd(uncert_count, 1ull);
    }
    for (int t = lane; t < n_w; t += 32) {
      const uint64_t me = key[t];
      int rank = 0;
      for (int u = 0; u < n_w; ++u) rank += key[u] > me;
      if (rank < k) {
        out_idx[lr * k + rank] = sel_key_idx(me);
        if (out_key) out_key[lr * k + rank] = sel_key_val(me);
      }
    }
    for (int t = n_w + lane; t < k; t += 32) {       // see rescore_exact_kernel: unfilled positions stay valid indices
      out_idx[lr * k + t] = (int32_t)(row_begin + lr);
      if (out_key) out_key[lr * k + t] = -INFINITY;
    }
  }
}

// ---- stage 2, FP64 tensor-core form: DMMA.8x8x4 over groups of 8 locality-ordered rows ------------------
// The grouped kernel above is bound by two things that have nothing to do with the FP64 pipe (ncu: XU pipe -- the
// fp32 -> fp64 conversions -- saturated, LSU data pipe 63 % from the shared-memory reads of the fp64 query copies,
// FP64 pipe 18-23 % busy).  A group's block of dots (8 queries x |union| candidates x D) is a small GEMM, so it
// is given to the FP64 tensor-core path (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4, full FP64 rate on B200):
//   * M = 8 = the queries of a group, N = 8 candidates per tile, K = 4;
//   * both operands come straight from the fp32 feature rows in global memory (L2-resident thanks to the
//     locality order): lane (g = lane / 4, t = lane % 4) loads ONE float4 of query g and one float4 of candidate g
//     of every tile per 16-wide K block and feeds its four values to four consecutive MMAs (MMA m of the block
//     uses k = 4 t + m on both sides: a permutation of K, which a dot product does not notice).  No shared-memory
//     operand staging at all, 64-byte coalesced segments per (row, block), one conversion per loaded value;
//   * a query value is converted once per 8 candidates x tiles and a candidate value once per 8 queries (the
//     grouped kernel: once per 4), so the XU work per FMA is halved and the shared-memory traffic is gone;
//   * the four warps of a CTA split K into quarters (no imbalance whatever |union| is) and their partial sums are
//     added in warp order: the summation order of a dot is fixed by the kernel alone, not by the group or the slot
//     the pair landed in -- keys are identical for every sharding;
//   * accumulators: 2 doubles per tile and lane -- 16 registers for 64 candidates per pass.
constexpr int kMThreads = 128;              // 4 warps = 4 quarters of K
constexpr int kMTiles = 8;                  // candidate tiles (8 candidates each) per pass
constexpr int kMChunk = kMTiles * 8;        // 64 candidates per pass

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// kMT = M tiles per group: a group is 8 * kMT locality-ordered rows.  With kMT = 2 every candidate value that is
// loaded and converted feeds 16 queries instead of 8 (half the XU work and half the L2 -> SM bytes per FMA) at the
// price of 64 accumulator registers; the union of 16 cluster mates is barely larger than that of 8.
template <int kMT>
struct MmaSmem {
  static constexpr int kQ = 8 * kMT;                       // queries per group
  static constexpr int kSlots = 2 * kQ * kWinMax;          // union hash table (|union| <= kQ * kWinMax)
  static constexpr int kSlotBits = kMT == 1 ? 11 : 12;
  // region 0: the hash table (setup) and, afterwards, the per-quarter partial dots of a pass -- never live together
  static constexpr size_t kTabBytes = (size_t)kSlots * 4 + (size_t)kSlots * 2;
  static constexpr size_t kDotBytes = (size_t)4 * kQ * kMChunk * 8;
  static constexpr size_t kRegion0 = kTabBytes > kDotBytes ? kTabBytes : kDotBytes;
  static constexpr size_t kUlist = kRegion0;                                    // int32 [kQ * kWinMax]
  static constexpr size_t kPid = kUlist + (size_t)kQ * kWinMax * 4;             // uint16 [kQ * kWinMax]
  static constexpr size_t kKey = kPid + (size_t)kQ * kWinMax * 2;               // uint64 [kQ * kWinMax]
  static constexpr size_t kBytes = kKey + (size_t)kQ * kWinMax * 8;
};

template <int kMT, int kUnroll, int kMinBlocks>
__global__ void __launch_bounds__(kMThreads, kMinBlocks) rescore_mma_kernel(
    const float* __restrict__ x, int64_t D, int64_t row_begin, int64_t n_rows, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ win_cnt, const int32_t* __restrict__ win_idx, const float* __restrict__ win_a, int k,
    float eps_in, const float* __restrict__ max_sqnorm, int32_t* __restrict__ out_idx, float* __restrict__ out_key,
    int32_t* __restrict__ uncert, unsigned* __restrict__ max_err_bits, unsigned long long* __restrict__ uncert_count,
    int groups_per_cta) {
  using L = MmaSmem<kMT>;
  constexpr int kQ = L::kQ;
  extern __shared__ __align__(16) unsigned char mma_smem[];
  int32_t* s_tab = reinterpret_cast<int32_t*>(mma_smem);                              // union hash table: column ids
  uint16_t* s_tid = reinterpret_cast<uint16_t*>(mma_smem + (size_t)L::kSlots * 4);    // slot -> position in the union
  double* s_dot = reinterpret_cast<double*>(mma_smem);                                // [4][kQ][kMChunk], after the setup
  int32_t* s_ulist = reinterpret_cast<int32_t*>(mma_smem + L::kUlist);                // the union
  uint16_t* s_pid = reinterpret_cast<uint16_t*>(mma_smem + L::kPid);                  // window entry -> union position
  uint64_t* s_key = reinterpret_cast<uint64_t*>(mma_smem + L::kKey);                  // window entry -> final sort key
  __shared__ int s_U;
  __shared__ int s_nw[kQ];
  __shared__ int64_t s_lr[kQ];
  __shared__ unsigned s_worst[kQ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float eps = max_sqnorm ? reid_tc_err_bound(*max_sqnorm) : eps_in;

  // groups_per_cta > 1: a CTA takes a CONTIGUOUS run of groups (cluster mates re-score against nearly the same
  // candidate rows, so a run re-uses what its first group pulled into L2).  Measured: it does not pay -- 0.62 ms with
  // one group per CTA against 0.74 / 0.87 / 0.89 ms with runs sized for 4 / 2 / 1 waves (the stage is short of
  // independent work per SM, not of L2 hits); the launch below uses groups_per_cta = 1.
  const int64_t n_groups = (n_rows + kQ - 1) / kQ;
  const int64_t grp_end = min(n_groups, ((int64_t)blockIdx.x + 1) * groups_per_cta);
  for (int64_t grp = (int64_t)blockIdx.x * groups_per_cta; grp < grp_end; ++grp) {
  __syncthreads();                                              // the previous group's shared-memory readers are done
  if (tid < kQ) {
    const int64_t slot = grp * kQ + tid;
    const int64_t lr = slot < n_rows ? (perm ? perm[slot] : slot) : -1;
    s_lr[tid] = lr;
    s_nw[tid] = lr >= 0 ? win_cnt[lr] : 0;
    s_worst[tid] = 0u;
  }
  if (tid == 0) s_U = 0;
  for (int h = tid; h < L::kSlots; h += kMThreads) s_tab[h] = -1;
  __syncthreads();
  // union of the windows
  for (int it = tid; it < kQ * kWinMax; it += kMThreads) {
    const int q = it / kWinMax, t = it % kWinMax;
    if (t < s_nw[q]) {
      const int32_t j = win_idx[s_lr[q] * kWinMax + t];
      uint32_t h = ((uint32_t)j * 0x9e3779b1u) >> (32 - L::kSlotBits);
      while (true) {
        const int32_t old = atomicCAS(&s_tab[h], -1, j);
        if (old == -1) {
          const int id = atomicAdd(&s_U, 1);
          s_ulist[id] = j;
          s_tid[h] = (uint16_t)id;
          break;
        }
        if (old == j) break;
        h = (h + 1) & (L::kSlots - 1);
      }
    }
  }
  __syncthreads();
  for (int it = tid; it < kQ * kWinMax; it += kMThreads) {
    const int q = it / kWinMax, t = it % kWinMax;
    if (t < s_nw[q]) {
      const int32_t j = win_idx[s_lr[q] * kWinMax + t];
      uint32_t h = ((uint32_t)j * 0x9e3779b1u) >> (32 - L::kSlotBits);
      while (s_tab[h] != j) h = (h + 1) & (L::kSlots - 1);
      s_pid[it] = s_tid[h];
    }
  }
  __syncthreads();                                              // the table is dead from here on: region 0 becomes s_dot
  const int U = s_U;

  const int g = lane >> 2, tg = lane & 3;
  const int64_t quarter = D >> 2;                               // D % 64 == 0: 16-wide K blocks, four quarters
  // a group that is not full multiplies by row 0's values for the missing queries; those dots are never read
  const float4* qp[kMT];
#pragma unroll
  for (int m = 0; m < kMT; ++m) {
    const int64_t lr_g = s_lr[m * 8 + g];
    qp[m] = reinterpret_cast<const float4*>(x + (row_begin + (lr_g >= 0 ? lr_g : 0)) * D + warp * quarter) + tg;
  }
  const int n_blocks = (int)(quarter >> 4);

  for (int c0 = 0; c0 < U; c0 += kMChunk) {
    const int nt = min(kMTiles, (U - c0 + 7) >> 3);
    const float4* cp[kMTiles];
#pragma unroll
    for (int t = 0; t < kMTiles; ++t) {
      const int id = c0 + t * 8 + g;
      cp[t] = reinterpret_cast<const float4*>(x + (int64_t)s_ulist[id < U ? id : U - 1] * D + warp * quarter) + tg;
    }
    double acc[kMT][kMTiles][2];
#pragma unroll
    for (int m = 0; m < kMT; ++m)
#pragma unroll
      for (int t = 0; t < kMTiles; ++t) acc[m][t][0] = acc[m][t][1] = 0.0;
#pragma unroll kUnroll
    for (int kb = 0; kb < n_blocks; ++kb) {
      float4 a[kMT];
#pragma unroll
      for (int m = 0; m < kMT; ++m) a[m] = qp[m][kb * 4];
      float4 b[kMTiles];
#pragma unroll
      for (int t = 0; t < kMTiles; ++t)
        if (t < nt) b[t] = cp[t][kb * 4];
      double ad[kMT][4];
#pragma unroll
      for (int m = 0; m < kMT; ++m) {
        ad[m][0] = (double)a[m].x;
        ad[m][1] = (double)a[m].y;
        ad[m][2] = (double)a[m].z;
        ad[m][3] = (double)a[m].w;
      }
#pragma unroll
      for (int t = 0; t < kMTiles; ++t) {
        if (t < nt) {
          const double b0 = (double)b[t].x, b1 = (double)b[t].y, b2 = (double)b[t].z, b3 = (double)b[t].w;
#pragma unroll
          for (int m = 0; m < kMT; ++m) {
            dmma_8x8x4(acc[m][t][0], acc[m][t][1], ad[m][0], b0);
            dmma_8x8x4(acc[m][t][0], acc[m][t][1], ad[m][1], b1);
            dmma_8x8x4(acc[m][t][0], acc[m][t][1], ad[m][2], b2);
            dmma_8x8x4(acc[m][t][0], acc[m][t][1], ad[m][3], b3);
          }
        }
      }
    }
    // C fragment: row = g (query within the M tile), columns 2 tg, 2 tg + 1 of the N tile
#pragma unroll
    for (int m = 0; m < kMT; ++m)
#pragma unroll
      for (int t = 0; t < kMTiles; ++t) {
        if (t < nt) {
          double* d = s_dot + ((size_t)warp * kQ + m * 8 + g) * kMChunk + t * 8 + 2 * tg;
          d[0] = acc[m][t][0];
          d[1] = acc[m][t][1];
        }
      }
    __syncthreads();
    for (int it = tid; it < kQ * kWinMax; it += kMThreads) {
      const int q = it / kWinMax, t = it % kWinMax;
      if (t < s_nw[q]) {
        const int id = (int)s_pid[it] - c0;
        if (id >= 0 && id < kMChunk) {
          const int64_t lr = s_lr[q];
          const double* d = s_dot + (size_t)q * kMChunk + id;
          const double dd = ((d[0] + d[(size_t)kQ * kMChunk]) + d[(size_t)2 * kQ * kMChunk]) + d[(size_t)3 * kQ * kMChunk];  // fixed order
          const float sc = (float)dd;
          const float err = fabsf(sc - win_a[lr * kWinMax + t]);
          atomicMax(&s_worst[q], __float_as_uint(err == err ? err : INFINITY));
          s_key[it] = sel_key(sc, win_idx[lr * kWinMax + t]);
        }
      }
    }
    __syncthreads();
  }

  // per row: audit, then order by (key desc, idx asc); the first k go out
  for (int q = warp; q < kQ; q += kMThreads / 32) {
    const int64_t lr = s_lr[q];
    if (lr < 0) continue;
    const int n_w = s_nw[q];
    const uint64_t* key = s_key + q * kWinMax;
    if (lane == 0) {
      const float worst = __uint_as_float(s_worst[q]);
      atomicMax(max_err_bits, __float_as_uint(worst));
      if (!(worst <= eps)) uncert[lr] = 1;  // the error model was violated (or NaN): do not trust the window
      if (uncert_count && uncert[lr]) atomicAdd(uncert_count, 1ull);
    }
    for (int t = lane; t < n_w; t += 32) {
      const uint64_t me = key[t];
      int rank = 0;
      for (int u = 0; u < n_w; ++u) rank += key[u] > me;
      if (rank < k) {
        out_idx[lr * k + rank] = sel_key_idx(me);
        if (out_key) out_key[lr * k + rank] = sel_key_val(me);
      }
    }
    for (int t = n_w + lane; t < k; t += 32) {       // see rescore_exact_kernel: unfilled positions stay valid indices
      out_idx[lr * k + t] = (int32_t)(row_begin + lr);
      if (out_key) out_key[lr * k + t] = -INFINITY;
    }
  }
  }  // groups of this CTA
}

inline size_t rescore_group_smem(int64_t D) {
  return (size_t)kGq * D * 8 + (size_t)kGq * kWinMax * 8 + (size_t)kGq * kGChunk * 8 + (size_t)kGSlots * 4 +
         (size_t)kGq * kWinMax * 4 + (size_t)kGSlots * 2 + (size_t)kGq * kWinMax * 2;
}

struct RescoreWs {
  int64_t* ptr;      // N + 1
  int32_t* hist;     // N
  int32_t* cursor;   // N
  int32_t* key;      // n
  int32_t* perm;     // n
  int32_t* win_cnt;  // n
  int32_t* win_idx;  // n * kWinMax
  float* win_a;      // n * kWinMax
};

inline size_t rs_align(size_t v) { return (v + 255) / 256 * 256; }

inline size_t rescore_carve(void* base, int64_t N, int64_t n, RescoreWs* w) {
  unsigned char* p = (unsigned char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += rs_align(bytes);
    return r;
  };
  RescoreWs t;
  t.ptr = (int64_t*)take(sizeof(int64_t) * (size_t)(N + 1));
  t.hist = (int32_t*)take(sizeof(int32_t) * (size_t)N);
  t.cursor = (int32_t*)take(sizeof(int32_t) * (size_t)N);
  t.key = (int32_t*)take(sizeof(int32_t) * (size_t)n);
  t.perm = (int32_t*)take(sizeof(int32_t) * (size_t)n);
  t.win_cnt = (int32_t*)take(sizeof(int32_t) * (size_t)n);
  t.win_idx = (int32_t*)take(sizeof(int32_t) * (size_t)n * kWinMax);
  t.win_a = (float*)take(sizeof(float) * (size_t)n * kWinMax);
  if (w) *w = t;
  return off;
}

}  // namespace reid

extern "C" {

size_t reid_knn_rescore_window_counts_offset(int64_t N, int64_t n_rows) {
  if (N < 0 || n_rows < 0) return 0;
  reid::RescoreWs w;
  reid::rescore_carve((void*)256, N, n_rows, &w);            // any non-null base: only the offset matters
  return (size_t)((unsigned char*)w.win_cnt - (unsigned char*)256);
}

size_t reid_knn_rescore_order_offset(int64_t N, int64_t n_rows) {
  if (N < 0 || n_rows < 0) return 0;
  reid::RescoreWs w;
  reid::rescore_carve((void*)256, N, n_rows, &w);
  return (size_t)((unsigned char*)w.perm - (unsigned char*)256);
}

size_t reid_knn_rescore_workspace_bytes(int64_t N, int64_t n_rows) {
  if (N < 0 || n_rows < 0) return 0;
  return reid::rescore_carve(nullptr, N, n_rows, nullptr);
}

int reid_knn_rescore(const float* x, int64_t N, int64_t D, int64_t row_begin, int64_t row_end, const uint64_t* cand,
                     const int32_t* cand_cnt, const uint32_t* row_tau, int n_lists, int list_cap, int64_t list_pitch_rows,
                     int k, float err_bound, const float* max_sqnorm, int locality_order, int32_t* out_idx, float* out_key, int32_t* uncertified_flag,
                     float* max_err_out, void* workspace, uint64_t* uncertified_count, void* stream) {
  return reid_knn_rescore_mapped(x, N, D, row_begin, row_end, cand, cand_cnt, row_tau, n_lists, list_cap, list_pitch_rows, k,
                                 err_bound, max_sqnorm, locality_order, nullptr, nullptr, out_idx, out_key, uncertified_flag,
                                 max_err_out, workspace, uncertified_count, stream);
}

int reid_knn_rescore_mapped(const float* x, int64_t N, int64_t D, int64_t row_begin, int64_t row_end, const uint64_t* cand,
                            const int32_t* cand_cnt, const uint32_t* row_tau, int n_lists, int list_cap, int64_t list_pitch_rows,
                            int k, float err_bound, const float* max_sqnorm, int locality_order, const int32_t* row_pos,
                            const int32_t* col_orig, int32_t* out_idx, float* out_key, int32_t* uncertified_flag,
                            float* max_err_out, void* workspace, uint64_t* uncertified_count, void* stream) {
  using namespace reid;
  REID_CHECK_ARG((row_pos == nullptr) == (col_orig == nullptr), "reid_knn_rescore_mapped: row_pos and col_orig come together");
  REID_CHECK_ARG(!row_pos || (row_begin == 0 && row_end == N), "reid_knn_rescore_mapped: the maps cover all N rows");
  REID_CHECK_ARG(x && cand && cand_cnt && row_tau && out_idx && uncertified_flag && max_err_out && workspace,
                 "reid_knn_rescore: NULL pointer");
  REID_CHECK_ARG(N > 0 && D > 0 && 0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_knn_rescore: bad shape");
  REID_CHECK_ARG(n_lists >= 1 && n_lists <= 64, "reid_knn_rescore: n_lists=%d (max 64)", n_lists);
  REID_CHECK_ARG(list_pitch_rows == 0 || list_pitch_rows >= row_end - row_begin, "reid_knn_rescore: list_pitch_rows too small");
  REID_CHECK_ARG(k >= 1 && k <= REID_TC_KEEP_MAX, "reid_knn_rescore: k=%d not in 1..%d", k, REID_TC_KEEP_MAX);
  REID_CHECK_ARG(list_cap >= 1, "reid_knn_rescore: list_cap=%d", list_cap);
  REID_CHECK_ARG(err_bound >= 0.f, "reid_knn_rescore: negative err_bound");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  RescoreWs w;
  rescore_carve(workspace, N, n, &w);
  REID_CUDA(cudaMemsetAsync(max_err_out, 0, sizeof(float), st));
  if (locality_order) REID_CUDA(cudaMemsetAsync(w.hist, 0, rs_align(sizeof(int32_t) * (size_t)N) * 2, st));  // hist + cursor
  const unsigned grid = (unsigned)((n + kRsWarps - 1) / kRsWarps);
  rescore_select_kernel<<<grid, kRsWarps * 32, 0, st>>>(row_begin, n, (const unsigned long long*)cand, cand_cnt, row_tau,
                                                        n_lists, list_cap, list_pitch_rows, k, err_bound, max_sqnorm, w.win_cnt,
                                                        w.win_idx, w.win_a,
                                                        uncertified_flag, locality_order ? w.key : nullptr, w.hist, row_pos, col_orig);
  REID_LAUNCH_CHECK();
  if (locality_order) {
    int rc = reid_scan_counts(w.hist, N, w.ptr, nullptr, stream);
    if (rc != REID_OK) return rc;
    order_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.key, n, w.ptr, w.cursor, w.perm);
    REID_LAUNCH_CHECK();
  }
#define REID_RS_LAUNCH(CH)                                                                                          \
  rescore_exact_kernel<CH><<<grid, kRsWarps * 32, 0, st>>>(x, D, row_begin, n, locality_order ? w.perm : nullptr,     \
                                                           w.win_cnt, w.win_idx, w.win_a, k, err_bound, max_sqnorm,  \
                                                           out_idx, out_key, uncertified_flag, (unsigned*)max_err_out,       \
                                                           (unsigned long long*)uncertified_count)
  const bool aligned = (((uintptr_t)x) & 15) == 0;
  static const bool no_group = dev_env("REID_RESCORE_GROUP", 1) == 0;
  const int rescore_variant = dev_env("REID_RESCORE_VARIANT", 2);   // developer builds: 1 = CUDA-core grouped kernel
  const int mma_waves = dev_env("REID_MMA_WAVES", 0);               // developer builds: > 0 = contiguous runs of groups per CTA
#define REID_MMA_LAUNCH(MT, U, B)                                                                                   \
  do {                                                                                                              \
    const size_t sm_ = MmaSmem<MT>::kBytes;                                                                         \
    REID_CUDA(cudaFuncSetAttribute(rescore_mma_kernel<MT, U, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_)); \
    const int64_t groups_ = (n + 8 * MT - 1) / (8 * MT);                                                            \
    int per_ = mma_waves > 0 ? (int)((groups_ + (int64_t)num_sms() * B * mma_waves - 1) / ((int64_t)num_sms() * B * mma_waves)) : 1; \
    if (per_ < 1) per_ = 1;                                                                                         \
    rescore_mma_kernel<MT, U, B><<<(unsigned)((groups_ + per_ - 1) / per_), kMThreads, sm_, st>>>(                   \
        x, D, row_begin, n, locality_order ? w.perm : nullptr, w.win_cnt, w.win_idx, w.win_a, k, err_bound,         \
        max_sqnorm, out_idx, out_key, uncertified_flag, (unsigned*)max_err_out,                                     \
        (unsigned long long*)uncertified_count, per_);                                                              \
  } while (0)
  if (aligned && D % 64 == 0 && !no_group && rescore_variant >= 2) {
#ifdef REID_DEV
    if (rescore_variant == 3) REID_MMA_LAUNCH(1, 1, 4);
    else if (rescore_variant == 4) REID_MMA_LAUNCH(2, 1, 3);
    else if (rescore_variant == 5) REID_MMA_LAUNCH(2, 2, 2);
    else if (rescore_variant == 6) REID_MMA_LAUNCH(2, 1, 2);
    else if (rescore_variant == 7) REID_MMA_LAUNCH(1, 2, 4);
    else
#endif
      REID_MMA_LAUNCH(1, 2, 4);
  } else if (aligned && D % 64 == 0 && D <= 2048 && !no_group) {
    const size_t smem = rescore_group_smem(D);
    REID_CUDA(cudaFuncSetAttribute(rescore_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rescore_group_kernel<<<(unsigned)((n + kGq - 1) / kGq), kGThreads, smem, st>>>(
        x, D, row_begin, n, locality_order ? w.perm : nullptr, w.win_cnt, w.win_idx, w.win_a, k, err_bound, max_sqnorm,
        out_idx, out_key, uncertified_flag, (unsigned*)max_err_out, (unsigned long long*)uncertified_count,
        dev_env("REID_RG_DEBUG", 0));
  } else if (aligned && D == 2048) REID_RS_LAUNCH(16);
  else if (aligned && D == 1024) REID_RS_LAUNCH(8);
  else if (aligned && D == 512) REID_RS_LAUNCH(4);
  else if (aligned && D == 256) REID_RS_LAUNCH(2);
  else REID_RS_LAUNCH(0);
#undef REID_RS_LAUNCH
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
