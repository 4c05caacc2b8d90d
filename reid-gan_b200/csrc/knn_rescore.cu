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
#include "common.cuh"

namespace reid {

constexpr int kRsWarps = 4;
constexpr int kRsMaxC = 2 * REID_TC_MAX_SPLITS * REID_TC_KEEP_MAX;  // 512 candidates per row at most
constexpr int kRsPer = kRsMaxC / 32;

__global__ void __launch_bounds__(kRsWarps * 32) rescore_kernel(
    const float* __restrict__ x, int64_t N, int64_t D, int64_t row_begin, int64_t row_end,
    const unsigned long long* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
    const uint32_t* __restrict__ row_tau, int n_lists, int k,
    float eps, int32_t* __restrict__ out_idx, float* __restrict__ out_key, int32_t* __restrict__ uncert,
    unsigned* __restrict__ max_err_bits) {
  __shared__ float s_a[kRsWarps][kRsMaxC];
  __shared__ int32_t s_j[kRsWarps][kRsMaxC];
  __shared__ uint64_t s_key[kRsWarps][kRsMaxC];
  const int w = threadIdx.x >> 5, lane = lane_id();
  const int64_t row = row_begin + (int64_t)blockIdx.x * kRsWarps + w;
  if (row >= row_end) return;
  const int64_t lr = row - row_begin;
  float* a = s_a[w];
  int32_t* jj = s_j[w];
  uint64_t* key = s_key[w];

  // Every column outside the row's lists scored <= bound.  The window [a_(k) - 2 eps, inf) starts at or
  // above bound - 2 eps (a_(k) >= bound because at least `keep` >= k listed columns score >= bound), so a
  // sweep that collects the entries >= bound - 2 eps finds every window member.
  const uint32_t tq = row_tau[lr];
  const float bound = tq ? ord_float(tq) : -INFINITY;

  // sweep the row's lists, keeping entries with score >= thr (at most kRsMaxC are stored; all are counted)
  auto collect = [&](float thr) {
    int cnt = 0;
    for (int q = 0; q < n_lists; ++q) {
      const int c = min(cand_cnt[lr * n_lists + q], REID_TC_CAP);
      const unsigned long long* src = cand + (lr * n_lists + q) * (int64_t)REID_TC_CAP;
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
          jj[pos] = (int32_t)(uint32_t)e;
        }
        cnt += __popc(b);
      }
    }
    __syncwarp();
    return cnt;
  };
  // k-th largest of a[0..m): bitwise binary search on the order-preserving integer image
  auto kth_largest = [&](int m) {
    uint32_t o[kRsPer];
#pragma unroll
    for (int u = 0; u < kRsPer; ++u) {
      const int t = u * 32 + lane;
      o[u] = t < m ? float_ord(a[t]) : 0u;
    }
    uint32_t T = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t c2 = T | (1u << bit);
      int c = 0;
#pragma unroll
      for (int u = 0; u < kRsPer; ++u) c += o[u] >= c2;
      c = __reduce_add_sync(kFull, c);
      if (c >= k) T = c2;
    }
    return m >= k ? ord_float(T) : -INFINITY;
  };

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
  bool certified = !overflow && n >= k && bound < lo;

  // window members, compacted to the front of key[] (as indices into a/jj for now)
  int n_w = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    const int t = base + lane;
    const bool in = t < n && a[t] >= lo;
    const unsigned b = __ballot_sync(kFull, in);
    if (in) key[n_w + __popc(b & ((1u << lane) - 1u))] = (uint64_t)t;
    n_w += __popc(b);
  }
  __syncwarp();
  // exact keys of the window members
  float worst = 0.f;
  const float* xi = x + row * D;
  for (int wi = 0; wi < n_w; ++wi) {
    const int t = (int)key[wi];
    const int32_t j = jj[t];
    const float* xj = x + (int64_t)j * D;
    double acc = 0.0;
    if ((D & 3) == 0) {
      const float4* a4 = reinterpret_cast<const float4*>(xi);
      const float4* b4 = reinterpret_cast<const float4*>(xj);
      for (int64_t d = lane; d < (D >> 2); d += 32) {
        const float4 p = a4[d], q = b4[d];
        acc = fma((double)p.x, (double)q.x, acc);
        acc = fma((double)p.y, (double)q.y, acc);
        acc = fma((double)p.z, (double)q.z, acc);
        acc = fma((double)p.w, (double)q.w, acc);
      }
    } else {
      for (int64_t d = lane; d < D; d += 32) acc = fma((double)xi[d], (double)xj[d], acc);
    }
    acc = warp_sum(acc);
    const float s = (float)acc;
    worst = fmaxf(worst, fabsf(s - a[t]));
    __syncwarp();
    if (lane == 0) key[wi] = sel_key(s, j);
  }
  __syncwarp();
  if (!(worst <= eps)) certified = false;  // the error model was violated (or NaN): do not trust the window
  if (n_w < k) certified = false;
  if (lane == 0) {
    atomicMax(max_err_bits, __float_as_uint(worst));
    uncert[lr] = certified ? 0 : 1;
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
}

}  // namespace reid

extern "C" {

int reid_knn_rescore(const float* x, int64_t N, int64_t D, int64_t row_begin, int64_t row_end, const uint64_t* cand,
                     const int32_t* cand_cnt, const uint32_t* row_tau, int n_lists, int k, float err_bound,
                     int32_t* out_idx, float* out_key, int32_t* uncertified_flag, float* max_err_out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && cand && cand_cnt && row_tau && out_idx && uncertified_flag && max_err_out,
                 "reid_knn_rescore: NULL pointer");
  REID_CHECK_ARG(N > 0 && D > 0 && 0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_knn_rescore: bad shape");
  REID_CHECK_ARG(n_lists >= 1 && n_lists <= 2 * REID_TC_MAX_SPLITS, "reid_knn_rescore: n_lists=%d (max %d)", n_lists,
                 2 * REID_TC_MAX_SPLITS);
  REID_CHECK_ARG(k >= 1 && k <= REID_TC_KEEP_MAX, "reid_knn_rescore: k=%d not in 1..%d", k, REID_TC_KEEP_MAX);
  REID_CHECK_ARG(err_bound >= 0.f, "reid_knn_rescore: negative err_bound");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  REID_CUDA(cudaMemsetAsync(max_err_out, 0, sizeof(float), st));
  rescore_kernel<<<(unsigned)((n + kRsWarps - 1) / kRsWarps), kRsWarps * 32, 0, st>>>(
      x, N, D, row_begin, row_end, (const unsigned long long*)cand, cand_cnt, row_tau, n_lists, k, err_bound, out_idx,
      out_key, uncertified_flag, (unsigned*)max_err_out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
