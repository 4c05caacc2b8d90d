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
constexpr int kRsMaxC = 256;  // n_splits * keep

__global__ void __launch_bounds__(kRsWarps * 32) rescore_kernel(
    const float* __restrict__ x, int64_t N, int64_t D, int64_t row_begin, int64_t row_end,
    const unsigned long long* __restrict__ cand, const int32_t* __restrict__ cand_cnt, int n_splits, int keep, int k,
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

  // gather the per-range lists; remember the weakest score of every FULL list
  int n = 0;
  float bound = -INFINITY;  // max over full lists of their weakest retained score
  for (int q = 0; q < n_splits; ++q) {
    const int c = min(cand_cnt[lr * n_splits + q], keep);
    const unsigned long long* src = cand + (lr * n_splits + q) * (int64_t)REID_TC_CAP;
    float mn = INFINITY;
    for (int t = lane; t < c; t += 32) {
      const unsigned long long e = src[t];
      const float v = __uint_as_float((uint32_t)(e >> 32));
      a[n + t] = v;
      jj[n + t] = (int32_t)(uint32_t)e;
      mn = fminf(mn, v);
    }
    mn = warp_min(mn);
    if (c == keep) bound = fmaxf(bound, mn);
    n += c;
  }
  __syncwarp();
  // k-th largest approximate score
  float a_k = -INFINITY;
  for (int t = lane; t < n; t += 32) {
    const float me = a[t];
    int rank = 0;
    for (int u = 0; u < n; ++u) rank += (a[u] > me) || (a[u] == me && u < t);
    if (rank == k - 1) a_k = me;
  }
  a_k = warp_max(a_k);
  const float lo = a_k - 2.0f * eps;
  bool certified = n >= k && bound < lo;

  // exact keys of the window members
  float worst = 0.f;
  const float* xi = x + row * D;
  for (int t = 0; t < n; ++t) {
    const int32_t j = jj[t];
    if (!(a[t] >= lo)) {  // warp-uniform
      if (lane == 0) key[t] = 0;  // below every real key
      continue;
    }
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
    if (lane == 0) key[t] = sel_key(s, j);
  }
  __syncwarp();
  if (!(worst <= eps)) certified = false;  // the error model was violated (or NaN): do not trust the window
  if (lane == 0) {
    atomicMax(max_err_bits, __float_as_uint(worst));
    uncert[lr] = certified ? 0 : 1;
  }
  // order by (key desc, idx asc); the first k go out
  for (int t = lane; t < n; t += 32) {
    const uint64_t me = key[t];
    if (me == 0) continue;
    int rank = 0;
    for (int u = 0; u < n; ++u) rank += key[u] > me;
    if (rank < k) {
      out_idx[lr * k + rank] = sel_key_idx(me);
      if (out_key) out_key[lr * k + rank] = sel_key_val(me);
    }
  }
}

}  // namespace reid

extern "C" {

int reid_knn_rescore(const float* x, int64_t N, int64_t D, int64_t row_begin, int64_t row_end, const uint64_t* cand,
                     const int32_t* cand_cnt, int n_splits, int keep, int k, float err_bound, int32_t* out_idx,
                     float* out_key, int32_t* uncertified_flag, float* max_err_out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && cand && cand_cnt && out_idx && uncertified_flag && max_err_out, "reid_knn_rescore: NULL pointer");
  REID_CHECK_ARG(N > 0 && D > 0 && 0 <= row_begin && row_begin <= row_end && row_end <= N, "reid_knn_rescore: bad shape");
  REID_CHECK_ARG(n_splits >= 1 && n_splits <= REID_TC_MAX_SPLITS && keep >= 1 && keep <= REID_TC_CAP &&
                     n_splits * keep <= kRsMaxC,
                 "reid_knn_rescore: n_splits=%d keep=%d (need n_splits*keep <= %d)", n_splits, keep, kRsMaxC);
  REID_CHECK_ARG(k >= 1 && k <= keep, "reid_knn_rescore: need 1 <= k <= keep (k=%d keep=%d)", k, keep);
  REID_CHECK_ARG(err_bound >= 0.f, "reid_knn_rescore: negative err_bound");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  REID_CUDA(cudaMemsetAsync(max_err_out, 0, sizeof(float), st));
  rescore_kernel<<<(unsigned)((n + kRsWarps - 1) / kRsWarps), kRsWarps * 32, 0, st>>>(
      x, N, D, row_begin, row_end, (const unsigned long long*)cand, cand_cnt, n_splits, keep, k, err_bound, out_idx,
      out_key, uncertified_flag, (unsigned*)max_err_out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
