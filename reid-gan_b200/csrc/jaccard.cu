// a7 -- Jaccard min-sum on the sparse V_qe (utils/faiss_rerank.py:102-119).
//   t_ij = sum_{c in nz(i) & nz(j), c ascending} min(Vq[i,c], Vq[j,c])   (sequential fp32 adds)
//   J_ij = max(0, 1 - t_ij / (2 - t_ij));  J_ij = 1 exactly when no column is shared.
// The reference walks the non-zero columns of row i in ascending order and, for each, adds
// into temp_min[rows of that column] (:109-110); both kernels below keep exactly that order,
// which is what makes J bit-symmetric and independent of how rows are sharded.
#include "common.cuh"

namespace reid {

__global__ void __launch_bounds__(256) jaccard_bounds_kernel(const int64_t* __restrict__ Q_ptr,
                                                             const int32_t* __restrict__ Q_idx,
                                                             const int64_t* __restrict__ C_ptr, int64_t row_begin,
                                                             int64_t row_end, int32_t* __restrict__ T_cnt) {
  const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  int64_t s = 0;
  for (int64_t p = Q_ptr[row] + lane_id(); p < Q_ptr[row + 1]; p += 32) {
    const int32_t c = Q_idx[p];
    s += C_ptr[c + 1] - C_ptr[c];
  }
  s = warp_sum(s);
  if (lane_id() == 0) T_cnt[row - row_begin] = (int32_t)(s > 0x7fffffff ? 0x7fffffff : s);
}

__device__ __forceinline__ uint32_t jhash(uint32_t v) {
  v *= 0x9e3779b1u;
  return v ^ (v >> 15);
}

__device__ __forceinline__ float jaccard_from_t(float t) {
  float j = __fsub_rn(1.0f, __fdiv_rn(t, __fsub_rn(2.0f, t)));
  return j < 0.f ? 0.f : j;
}

// One warp per row; per-warp open-addressing table (j -> running t) in shared memory.
// Columns are consumed one after another (ascending), the rows of one column in parallel:
// a column holds each j at most once, so lanes never collide on a slot inside a step.
constexpr int kJWarps = 4;

__global__ void __launch_bounds__(kJWarps * 32) jaccard_neighbors_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val,
    int64_t row_begin, int64_t n_rows, const int32_t* __restrict__ rows_list, float eps,
    const int64_t* __restrict__ slot_ptr, int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_val,
    int32_t* __restrict__ nbr_cnt, int slots) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = lane_id();
  int32_t* tkey = reinterpret_cast<int32_t*>(smem_raw) + (size_t)w * slots;
  float* tval = reinterpret_cast<float*>(smem_raw + (size_t)kJWarps * slots * 4) + (size_t)w * slots;
  const int64_t li = (int64_t)blockIdx.x * kJWarps + w;
  if (li >= n_rows) return;
  const int64_t lr = rows_list ? rows_list[li] : li;  // local row id
  const int64_t row = row_begin + lr;
  const uint32_t smask = (uint32_t)slots - 1u;

  for (int s = lane; s < slots; s += 32) tkey[s] = -1;
  __syncwarp();

  int used = 0;
  bool overflow = false;
  const int64_t qa = Q_ptr[row], qb = Q_ptr[row + 1];
  for (int64_t p = qa; p < qb && !overflow; ++p) {
    const int32_t c = Q_idx[p];
    const float vic = Q_val[p];
    const int64_t ca = C_ptr[c], cb = C_ptr[c + 1];
    for (int64_t q0 = ca; q0 < cb; q0 += 32) {
      const int64_t q = q0 + lane;
      bool fresh = false;
      if (q < cb) {
        const int32_t j = C_idx[q];
        const float m = fminf(vic, C_val[q]);
        uint32_t h = jhash((uint32_t)j) & smask;
        while (true) {
          const int32_t old = atomicCAS(&tkey[h], -1, j);
          if (old == -1) {
            tval[h] = m;  // 0 + m
            fresh = true;
            break;
          }
          if (old == j) {
            tval[h] = __fadd_rn(tval[h], m);
            break;
          }
          h = (h + 1) & smask;
        }
      }
      used += __popc(__ballot_sync(kFull, fresh));
      if (used > (slots >> 1) + (slots >> 2)) {  // keep the load factor under 3/4 so probing terminates
        overflow = true;
        break;
      }
    }
    __syncwarp();
  }
  if (overflow) {
    if (lane == 0) nbr_cnt[lr] = -1;
    return;
  }
  __syncwarp();
  const int64_t o = slot_ptr[lr];
  int cnt = 0;
  for (int base = 0; base < slots; base += 32) {
    const int32_t j = tkey[base + lane];
    float jd = 2.f;
    if (j >= 0) jd = jaccard_from_t(tval[base + lane]);
    const bool keep = j >= 0 && jd <= eps;
    const unsigned b = __ballot_sync(kFull, keep);
    if (keep) {
      const int64_t dst = o + cnt + __popc(b & ((1u << lane) - 1u));
      nbr_idx[dst] = j;
      if (nbr_val) nbr_val[dst] = jd;
    }
    cnt += __popc(b);
  }
  if (lane == 0) nbr_cnt[lr] = cnt;
}

// Dense rows for the drop-in return value.  One CTA per row; the accumulator row lives in
// shared memory when N floats fit, else in the output row itself (L2-resident).
template <bool kSmemAcc>
__global__ void __launch_bounds__(256) jaccard_dense_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val, int64_t N,
    int64_t row_begin, float* __restrict__ out, int64_t ld) {
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
        acc[j] = __fadd_rn(acc[j], m);
      } else {
        __stcg(&acc[j], __fadd_rn(__ldcg(&acc[j]), m));
      }
    }
    __syncthreads();
  }
  for (int64_t j = threadIdx.x; j < N; j += blockDim.x) {
    const float t = kSmemAcc ? acc[j] : __ldcg(&acc[j]);
    orow[j] = jaccard_from_t(t);
  }
}

// Rows whose partner set does not fit any shared-memory table ("hub" rows): one CTA per row, dense
// accumulator row in global scratch (L2-resident), same ascending-column order, then an ordered
// block-wide compaction of { j : J_ij <= eps }.
__global__ void __launch_bounds__(256) jaccard_neighbors_heavy_kernel(
    const int64_t* __restrict__ Q_ptr, const int32_t* __restrict__ Q_idx, const float* __restrict__ Q_val,
    const int64_t* __restrict__ C_ptr, const int32_t* __restrict__ C_idx, const float* __restrict__ C_val, int64_t N,
    int64_t row_begin, const int32_t* __restrict__ rows_list, float eps, const int64_t* __restrict__ slot_ptr,
    int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_val, int32_t* __restrict__ nbr_cnt,
    float* __restrict__ scratch) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int64_t lr = rows_list[blockIdx.x];
  const int64_t row = row_begin + lr;
  float* acc = scratch + (int64_t)blockIdx.x * N;
  for (int64_t j = threadIdx.x; j < N; j += blockDim.x) __stcg(&acc[j], 0.f);
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int64_t p = Q_ptr[row]; p < Q_ptr[row + 1]; ++p) {
    const int32_t c = Q_idx[p];
    const float vic = Q_val[p];
    for (int64_t q = C_ptr[c] + threadIdx.x; q < C_ptr[c + 1]; q += blockDim.x) {
      const int32_t j = C_idx[q];
      __stcg(&acc[j], __fadd_rn(__ldcg(&acc[j]), fminf(vic, C_val[q])));
    }
    __syncthreads();
  }
  const int64_t o = slot_ptr[lr];
  const int lane = lane_id(), w = threadIdx.x >> 5;
  for (int64_t base = 0; base < N; base += blockDim.x) {
    const int64_t j = base + threadIdx.x;
    float jd = 2.f;
    if (j < N) jd = jaccard_from_t(__ldcg(&acc[j]));
    const bool keep = j < N && jd <= eps;
    const unsigned b = __ballot_sync(kFull, keep);
    if (lane == 0) s_warp[w] = __popc(b);
    __syncthreads();
    int before = s_base;
    for (int ww = 0; ww < w; ++ww) before += s_warp[ww];
    if (keep) {
      const int64_t dst = o + before + __popc(b & ((1u << lane) - 1u));
      nbr_idx[dst] = (int32_t)j;
      if (nbr_val) nbr_val[dst] = jd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int ww = 0; ww < 8; ++ww) tot += s_warp[ww];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) nbr_cnt[lr] = s_base;
}

}  // namespace reid

extern "C" {

int reid_jaccard_bounds(const int64_t* Q_ptr, const int32_t* Q_idx, const int64_t* C_ptr, int64_t row_begin,
                        int64_t row_end, int32_t* T_cnt, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(Q_ptr && Q_idx && C_ptr && T_cnt, "reid_jaccard_bounds: NULL pointer");
  REID_CHECK_ARG(0 <= row_begin && row_begin <= row_end, "reid_jaccard_bounds: bad row range");
  const int64_t n = row_end - row_begin;
  if (n == 0) return REID_OK;
  jaccard_bounds_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(Q_ptr, Q_idx, C_ptr, row_begin,
                                                                                  row_end, T_cnt);
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
  const size_t smem = (size_t)kJWarps * table_slots * 8;
  REID_CHECK_ARG(smem <= 224 * 1024, "reid_jaccard_neighbors: table_slots=%d does not fit shared memory", table_slots);
  const int64_t n = rows_list ? n_list : row_end - row_begin;
  if (n == 0) return REID_OK;
  REID_CUDA(cudaFuncSetAttribute(jaccard_neighbors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jaccard_neighbors_kernel<<<(unsigned)((n + kJWarps - 1) / kJWarps), kJWarps * 32, smem, (cudaStream_t)stream>>>(
      Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, row_begin, n, rows_list, eps, slot_ptr, nbr_idx, nbr_val, nbr_cnt,
      table_slots);
  REID_LAUNCH_CHECK();
  return REID_OK;
}

int reid_jaccard_dense(const int64_t* Q_ptr, const int32_t* Q_idx, const float* Q_val, const int64_t* C_ptr,
                       const int32_t* C_idx, const float* C_val, int64_t N, int64_t row_begin, int64_t row_end,
                       float* out, int64_t ld, void* stream) {
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
                                                              out, ld);
  } else {
    jaccard_dense_kernel<false><<<(unsigned)n, 256, 0, st>>>(Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin,
                                                            out, ld);
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
      Q_ptr, Q_idx, Q_val, C_ptr, C_idx, C_val, N, row_begin, rows_list, eps, slot_ptr, nbr_idx, nbr_val, nbr_cnt,
      scratch);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
