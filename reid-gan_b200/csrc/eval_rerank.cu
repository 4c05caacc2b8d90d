// f2 -- dense front and back ends of the evaluation-time re-ranking,
// clustercontrast/utils/rerank.py re_ranking :31-97.  The k-reciprocal sets, the query expansion, the inverted
// index and the min-sum Jaccard rows are the kernels of the pseudo-label path (rerank_sets.cu, rerank_sparse.cu,
// jaccard.cu); what is specific here is the (query+gallery)^2 matrix of column-max-normalised squared distances
// (:36-41), the exp(-d) weights read from it (:66-67) and the final blend (:95-96).
#include "common.cuh"

namespace reid {

// old[i][j] = block(i, j)^2 over the 2 x 2 block matrix [[q_q, q_g], [q_g^T, g_g]]   (:36-40)
__global__ void __launch_bounds__(256) rr_square_blocks_kernel(const float* __restrict__ qg, const float* __restrict__ qq,
                                                               const float* __restrict__ gg, int64_t Q, int64_t G,
                                                               float* __restrict__ old) {
  const int64_t M = Q + G;
  const int64_t i = blockIdx.y;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if (i < Q) v = j < Q ? qq[i * Q + j] : qg[i * G + (j - Q)];
    else v = j < Q ? qg[j * G + (i - Q)] : gg[(i - Q) * G + (j - Q)];
    old[i * M + j] = __fmul_rn(v, v);
  }
}

// colmax[j] = max_i old[i][j]  (:41 np.max(axis=0)); squares are non-negative, so the float bits order like uints
__global__ void __launch_bounds__(256) rr_colmax_kernel(const float* __restrict__ old, int64_t M, int64_t rows_per_block,
                                                        unsigned* __restrict__ colmax_bits) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float m = 0.f;
  for (int64_t i = r0; i < r1; ++i) m = fmaxf(m, old[i * M + j]);
  atomicMax(&colmax_bits[j], __float_as_uint(m));
}

// dist[i][j] = old[j][i] / colmax[i]   (:41 np.transpose(1. * original_dist / max)); 32 x 32 tiles through smem
__global__ void __launch_bounds__(256) rr_normalise_transpose_kernel(const float* __restrict__ old,
                                                                     const float* __restrict__ colmax, int64_t M,
                                                                     float* __restrict__ dist) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  const int64_t bi = (int64_t)blockIdx.y * 32, bj = (int64_t)blockIdx.x * 32;   // output tile rows bi.., cols bj..
  for (int r = ty; r < 32; r += 8) {                                // read old[bj + r][bi + tx]
    const int64_t a = bj + r, b = bi + tx;
    tile[r][tx] = (a < M && b < M) ? old[a * M + b] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {                                // write dist[bi + r][bj + tx] = old[bj + tx][bi + r] / colmax[bi + r]
    const int64_t i = bi + r, j = bj + tx;
    if (i < M && j < M) dist[i * M + j] = __fdiv_rn(tile[tx][r], colmax[i]);
  }
}

// V[i, e] = exp(-dist[i, e]) / sum_e' exp(-dist[i, e'])   (:66-67), written into the CSR at E_ptr; one warp per row
__global__ void __launch_bounds__(256) rr_weights_kernel(const float* __restrict__ dist, int64_t M,
                                                         const int32_t* __restrict__ E_pad, int stride,
                                                         const int64_t* __restrict__ E_ptr, int64_t n_rows,
                                                         int32_t* __restrict__ E_idx, float* __restrict__ V_val) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = lane_id();
  const int64_t p0 = E_ptr[row];
  const int n = (int)(E_ptr[row + 1] - p0);
  const int32_t* erow = E_pad + row * (int64_t)stride;
  float sum = 0.f;
  for (int e = lane; e < n; e += 32) {
    const int32_t j = erow[e];
    const float w = expf(-dist[row * M + j]);
    E_idx[p0 + e] = j;
    V_val[p0 + e] = w;
    sum += w;
  }
  sum = warp_sum(sum);
  __syncwarp();
  for (int e = lane; e < n; e += 32) V_val[p0 + e] = __fdiv_rn(V_val[p0 + e], sum);
}

// final[i][g] = J[i][Q + g] * (1 - lambda) + dist[i][Q + g] * lambda   (:95-96)
__global__ void __launch_bounds__(256) rr_final_kernel(const float* __restrict__ J, int64_t ldJ, const float* __restrict__ dist,
                                                       int64_t M, int64_t Q, int64_t G, float one_minus_lambda, float lambda,
                                                       float* __restrict__ out) {
  const int64_t i = blockIdx.y;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x)
    out[i * G + g] = __fadd_rn(__fmul_rn(J[i * ldJ + Q + g], one_minus_lambda), __fmul_rn(dist[i * M + Q + g], lambda));
}

}  // namespace reid

extern "C" {

int reid_rr_normalised_distance(const float* q_g, const float* q_q, const float* g_g, int64_t Q, int64_t G, float* scratch_mm,
                                float* colmax, float* dist, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(q_g && q_q && g_g && scratch_mm && colmax && dist && Q >= 1 && G >= 1, "reid_rr_normalised_distance: bad arguments");
  const int64_t M = Q + G;
  REID_CHECK_ARG(M < 65536 * 32, "reid_rr_normalised_distance: matrix too large");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t bx = (M + 255) / 256;
  if (bx > 64) bx = 64;
  rr_square_blocks_kernel<<<dim3((unsigned)bx, (unsigned)M), 256, 0, st>>>(q_g, q_q, g_g, Q, G, scratch_mm);
  REID_LAUNCH_CHECK();
  REID_CUDA(cudaMemsetAsync(colmax, 0, sizeof(float) * (size_t)M, st));
  const int64_t rows_per_block = 256;
  rr_colmax_kernel<<<dim3((unsigned)((M + 255) / 256), (unsigned)((M + rows_per_block - 1) / rows_per_block)), 256, 0, st>>>(
      scratch_mm, M, rows_per_block, (unsigned*)colmax);
  REID_LAUNCH_CHECK();
  rr_normalise_transpose_kernel<<<dim3((unsigned)((M + 31) / 32), (unsigned)((M + 31) / 32)), 256, 0, st>>>(scratch_mm, colmax, M,
                                                                                                             dist);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_rr_weights(const float* dist, int64_t M, const int32_t* E_pad, int stride, const int64_t* E_ptr, int64_t n_rows,
                    int32_t* E_idx, float* V_val, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(dist && E_pad && E_ptr && E_idx && V_val && stride >= 1 && n_rows >= 0, "reid_rr_weights: bad arguments");
  if (n_rows == 0) return REID_OK;
  rr_weights_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dist, M, E_pad, stride, E_ptr, n_rows, E_idx,
                                                                                  V_val);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_rr_final(const float* J, int64_t ldJ, const float* dist, int64_t Q, int64_t G, float one_minus_lambda, float lambda,
                  float* out, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(J && dist && out && Q >= 1 && G >= 1 && ldJ >= Q + G, "reid_rr_final: bad arguments");
  int64_t bx = (G + 255) / 256;
  if (bx > 64) bx = 64;
  rr_final_kernel<<<dim3((unsigned)bx, (unsigned)Q), 256, 0, (cudaStream_t)stream>>>(J, ldJ, dist, Q + G, Q, G, one_minus_lambda,
                                                                                   lambda, out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
