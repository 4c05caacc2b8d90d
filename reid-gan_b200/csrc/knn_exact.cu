// a1 -- exact kNN with the canonical key (replaces faiss IndexFlatL2.search,
// utils/faiss_rerank.py:58-62).  This is the always-correct path: fp64-accumulated
// dot products on the CUDA cores, rounded once to fp32, then an exact per-row radix
// select on the 64-bit (key, ~index) composite.  The tensor-core path
// (simgemm_tc.cu + knn_rescore.cu) uses it only for rows it cannot certify.
#include "common.cuh"

namespace reid {

constexpr int TM = 64, TN = 64, TK = 32;

// keys[r][j] = fp32( sum_d (double)x[q_r][d] * (double)x[j][d] ), q_r = rows ? rows[r] : row_begin + r
// half_sqnorm (optional, fp64[N]): the squared-L2 form of the key, fp32( dot - ||x_j||^2 / 2 ) -- L2 ascending with the
// per-query constant dropped (rows of different norms; for unit-norm rows both keys order alike).
__global__ void __launch_bounds__(256) dot64_tile_kernel(const float* __restrict__ x, int64_t N, int64_t D,
                                                         const int32_t* __restrict__ rows, int64_t row_begin,
                                                         int64_t n_rows, const double* __restrict__ half_sqnorm,
                                                         float* __restrict__ keys) {
  __shared__ float As[TM][TK + 1];
  __shared__ float Bs[TN][TK + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r0 = (int64_t)blockIdx.y * TM, c0 = (int64_t)blockIdx.x * TN;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int64_t k0 = 0; k0 < D; k0 += TK) {
    // 64 x 32 floats per tile = 2048 elements, 8 per thread
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int lin = threadIdx.x + e * 256;
      int rr = lin >> 5, kk = lin & 31;
      int64_t ar = r0 + rr, bc = c0 + rr, kd = k0 + kk;
      float av = 0.f, bv = 0.f;
      if (kd < D) {
        if (ar < n_rows) {
          int64_t q = rows ? (int64_t)rows[ar] : row_begin + ar;
          av = x[q * D + kd];
        }
        if (bc < N) bv = x[bc * D + kd];
      }
      As[rr][kk] = av;
      Bs[rr][kk] = bv;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (double)As[ty + 16 * i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = (double)Bs[tx + 16 * j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = r0 + ty + 16 * i;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t c = c0 + tx + 16 * j;
      if (c < N) keys[r * N + c] = (float)(half_sqnorm ? acc[i][j] - half_sqnorm[c] : acc[i][j]);
    }
  }
}

// half_sqnorm[j] = 0.5 * sum_d x[j,d]^2 in fp64; one warp per row
__global__ void __launch_bounds__(256) half_sqnorm64_kernel(const float* __restrict__ x, int64_t N, int64_t D,
                                                            double* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  double s = 0.0;
  for (int64_t d = lane_id(); d < D; d += 32) {
    const double v = (double)x[row * D + d];
    s = fma(v, v, s);
  }
  s = warp_sum(s);
  if (lane_id() == 0) out[row] = 0.5 * s;
}

// { max_i ||x_i||^2, min_i ||x_i||^2 } in fp32 (what reid_features_to_half reports on the tensor-core path)
__global__ void __launch_bounds__(256) sqnorm_range_kernel(const float* __restrict__ x, int64_t N, int64_t D,
                                                           float* __restrict__ range) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  float ss = 0.f;
  for (int64_t d = lane_id(); d < D; d += 32) {
    const float v = x[row * D + d];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane_id() == 0) {
    atomicMax((unsigned*)range, __float_as_uint(ss));
    atomicMin((unsigned*)range + 1, __float_as_uint(ss));
  }
}

// One CTA per query row: exact top-k of N keys by (key desc, index asc).
// MSB-first radix select over the 64-bit composite, 8 bits a pass, early exit when the
// chosen bucket is consumed whole; then an O(k^2) rank sort of the k survivors.
constexpr int kSelThreads = 256;
constexpr int kMaxK = 128;

// negate = true selects the k SMALLEST keys (ascending, ties by index) -- the order of np.argsort on distances.
template <bool kNegate>
__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(const float* __restrict__ keys, int64_t N, int k,
                                                                  int32_t* __restrict__ out_idx,
                                                                  float* __restrict__ out_key) {
  __shared__ unsigned hist[256];
  __shared__ uint64_t s_prefix;
  __shared__ int s_need, s_done, s_count;
  __shared__ uint64_t sel[kMaxK];
  const float* kr = keys + (int64_t)blockIdx.x * N;
  const int t = threadIdx.x;
  if (t == 0) {
    s_prefix = 0;
    s_need = k;
    s_done = 0;
    s_count = 0;
  }
  int shift = 56;
  for (int pass = 0; pass < 8; ++pass, shift -= 8) {
    hist[t] = 0;
    __syncthreads();
    if (s_done) break;
    const uint64_t pfx = s_prefix;
    for (int64_t j = t; j < N; j += kSelThreads) {
      uint64_t key = sel_key(kNegate ? -kr[j] : kr[j], (int)j);
      if (pass == 0 || (key >> (shift + 8)) == pfx) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (t == 0) {
      int need = s_need, cum = 0, d = 255;
      for (; d >= 0; --d) {
        if (cum + (int)hist[d] >= need) break;
        cum += hist[d];
      }
      need -= cum;  // still wanted from bucket d
      s_prefix = (pfx << 8) | (unsigned)d;
      s_need = need;
      if ((int)hist[d] == need) s_done = 1;  // bucket taken whole: every key with prefix >= s_prefix is in
    }
    __syncthreads();
  }
  __syncthreads();
  // `shift` now = bit position below the resolved prefix minus 8 -> resolved bits start at shift + 8
  const int low = shift + 8;
  const uint64_t pfx = s_prefix;
  for (int64_t j = t; j < N; j += kSelThreads) {
    uint64_t key = sel_key(kNegate ? -kr[j] : kr[j], (int)j);
    if ((low >= 64 ? 0 : (key >> low)) >= pfx) {
      int p = atomicAdd(&s_count, 1);
      if (p < kMaxK) sel[p] = key;
    }
  }
  __syncthreads();
  const int n = min(s_count, k);
  if (t < n) {
    uint64_t me = sel[t];
    int rank = 0;
    for (int u = 0; u < n; ++u) rank += sel[u] > me;
    out_idx[(int64_t)blockIdx.x * k + rank] = sel_key_idx(me);
    if (out_key) out_key[(int64_t)blockIdx.x * k + rank] = kNegate ? -sel_key_val(me) : sel_key_val(me);
  }
}

}  // namespace reid

extern "C" {

size_t reid_knn_exact_scratch_bytes(int64_t N, int64_t n_rows) {
  if (n_rows < 1) n_rows = 1;
  return (size_t)N * (size_t)n_rows * sizeof(float);
}

static int knn_exact_impl(const float* x, int64_t N, int64_t D, const int32_t* rows_list, int64_t row_begin,
                          int64_t n_rows, int k, int32_t* out_idx, float* out_key, const double* half_sqnorm, float* keys,
                          int64_t chunk, cudaStream_t st) {
  using namespace reid;
  for (int64_t s = 0; s < n_rows; s += chunk) {
    int64_t m = n_rows - s < chunk ? n_rows - s : chunk;
    dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((m + TM - 1) / TM));
    dot64_tile_kernel<<<grid, 256, 0, st>>>(x, N, D, rows_list ? rows_list + s : nullptr, row_begin + s, m, half_sqnorm,
                                           keys);
    REID_LAUNCH_CHECK();
    select_topk_kernel<false><<<(unsigned)m, kSelThreads, 0, st>>>(keys, N, k, out_idx + s * k,
                                                           out_key ? out_key + s * k : nullptr);
    REID_LAUNCH_CHECK();
  }
  return REID_OK;
}

int reid_sqnorm_range(const float* x, int64_t N, int64_t D, float* sqnorm_range, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && sqnorm_range && N >= 0 && D > 0, "reid_sqnorm_range: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  REID_CUDA(cudaMemsetAsync(sqnorm_range, 0, sizeof(float), st));
  REID_CUDA(cudaMemsetAsync(sqnorm_range + 1, 0x7f, sizeof(float), st));
  if (N == 0) return REID_OK;
  sqnorm_range_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(x, N, D, sqnorm_range);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_knn_exact_l2(const float* x, int64_t N, int64_t D, const int32_t* rows_list, int64_t row_begin,
                      int64_t n_rows, int k, int32_t* out_idx, float* out_key, void* scratch, size_t scratch_bytes,
                      void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && out_idx && scratch, "reid_knn_exact_l2: NULL pointer");
  REID_CHECK_ARG(N > 0 && D > 0 && n_rows >= 0, "reid_knn_exact_l2: bad shape N=%lld D=%lld", (long long)N, (long long)D);
  REID_CHECK_ARG(k >= 1 && k <= kMaxK && k <= N, "reid_knn_exact_l2: k=%d out of range (1..min(%d,N))", k, kMaxK);
  REID_CHECK_ARG(N < (1ll << 31), "reid_knn_exact_l2: N too large for int32 indices");
  const size_t head = ((size_t)N * sizeof(double) + 255) / 256 * 256;      // fp64 half norms, then the key rows
  REID_CHECK_ARG(scratch_bytes > head && (scratch_bytes - head) / (sizeof(float) * (size_t)N) >= 1,
                 "reid_knn_exact_l2: scratch too small (%zu bytes, need >= %zu)", scratch_bytes,
                 head + (size_t)N * sizeof(float));
  const int64_t chunk = (int64_t)((scratch_bytes - head) / (sizeof(float) * (size_t)N));
  cudaStream_t st = (cudaStream_t)stream;
  double* hn = (double*)scratch;
  half_sqnorm64_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(x, N, D, hn);
  REID_LAUNCH_CHECK();
  return knn_exact_impl(x, N, D, rows_list, row_begin, n_rows, k, out_idx, out_key, hn, (float*)((unsigned char*)scratch + head),
                        chunk, st);
}

int reid_knn_exact(const float* x, int64_t N, int64_t D, const int32_t* rows_list, int64_t row_begin,
                   int64_t n_rows, int k, int32_t* out_idx, float* out_key, void* scratch, size_t scratch_bytes,
                   void* stream) {
  using namespace reid;
  REID_CHECK_ARG(x && out_idx && scratch, "reid_knn_exact: NULL pointer");
  REID_CHECK_ARG(N > 0 && D > 0 && n_rows >= 0, "reid_knn_exact: bad shape N=%lld D=%lld", (long long)N, (long long)D);
  REID_CHECK_ARG(k >= 1 && k <= kMaxK && k <= N, "reid_knn_exact: k=%d out of range (1..min(%d,N))", k, kMaxK);
  REID_CHECK_ARG(N < (1ll << 31), "reid_knn_exact: N too large for int32 indices");
  int64_t chunk = (int64_t)(scratch_bytes / (sizeof(float) * (size_t)N));
  REID_CHECK_ARG(chunk >= 1, "reid_knn_exact: scratch too small (%zu bytes, need >= %zu)", scratch_bytes,
                 (size_t)N * sizeof(float));
  return knn_exact_impl(x, N, D, rows_list, row_begin, n_rows, k, out_idx, out_key, nullptr, (float*)scratch, chunk,
                        (cudaStream_t)stream);
}

int reid_select_rows(const float* keys, int64_t N, int64_t n_rows, int k, int ascending, int32_t* out_idx, float* out_key,
                     void* stream) {
  using namespace reid;
  REID_CHECK_ARG(keys && out_idx && N > 0 && n_rows >= 0, "reid_select_rows: bad arguments");
  REID_CHECK_ARG(k >= 1 && k <= kMaxK && k <= N && N < (1ll << 31), "reid_select_rows: k=%d out of range", k);
  if (n_rows == 0) return REID_OK;
  if (ascending) select_topk_kernel<true><<<(unsigned)n_rows, kSelThreads, 0, (cudaStream_t)stream>>>(keys, N, k, out_idx, out_key);
  else select_topk_kernel<false><<<(unsigned)n_rows, kSelThreads, 0, (cudaStream_t)stream>>>(keys, N, k, out_idx, out_key);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
