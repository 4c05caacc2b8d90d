// Error plumbing + the count->pointer scan used between the two-pass (count / fill) stages.
#include <stdarg.h>

#include "common.cuh"

namespace reid {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// One CTA scans the counts tile by tile: 1024 threads x 16 consecutive elements per tile (four 16-byte loads per
// thread when the array is 16-byte aligned), warp-shuffle scans at two levels, running carry in a register.
// 32k counts = 2 tiles; the stats (total, max, sum of squares) come out of the same pass.
constexpr int kScanPer = 16;
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ cnt, int64_t n,
                                                           int64_t* __restrict__ ptr, int64_t* __restrict__ stats) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t warp_max_s[32];
  __shared__ int64_t warp_sq[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const bool vec = (((uintptr_t)cnt) & 15) == 0;
  int64_t carry = 0, sq = 0;
  int32_t mx = 0;
  for (int64_t base = 0; base < n; base += 1024 * kScanPer) {
    const int64_t i0 = base + (int64_t)t * kScanPer;
    int32_t c[kScanPer];
    if (vec && i0 + kScanPer <= n) {
#pragma unroll
      for (int q = 0; q < kScanPer / 4; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(cnt + i0 + 4 * q);
        c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kScanPer; ++j) c[j] = i0 + j < n ? cnt[i0 + j] : 0;
    }
    int32_t s = 0;                                     // a tile holds < 2^31 in total (counts are small)
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
      mx = max(mx, c[j]);
      sq += (int64_t)c[j] * c[j];
      s += c[j];
    }
    int32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
      const int32_t v = warp_tot[lane];
      int32_t iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(kFull, iv, o);
        if (lane >= o) iv += u;
      }
      warp_tot[lane] = iv - v;                          // exclusive offset of each warp inside the tile
      if (lane == 31) warp_max_s[0] = iv;               // tile total
    }
    __syncthreads();
    int64_t run = carry + warp_tot[w] + (inc - s);
    const int64_t tile_total = warp_max_s[0];
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
      if (i0 + j < n) ptr[i0 + j] = run;
      run += c[j];
    }
    carry += tile_total;
    __syncthreads();                                    // warp_tot / warp_max_s are rewritten by the next tile
  }
  mx = warp_max(mx);
  sq = warp_sum(sq);
  if (lane == 0) {
    warp_max_s[w] = mx;
    warp_sq[w] = sq;
  }
  __syncthreads();
  if (w == 0) {
    const int32_t m = warp_max(warp_max_s[lane]);
    const int64_t q = warp_sum(warp_sq[lane]);
    if (lane == 0) {
      ptr[n] = carry;
      if (stats) {
        stats[0] = carry;
        stats[1] = m;
        stats[2] = q;
      }
    }
  }
}

}  // namespace reid

extern "C" {

int reid_abi_version(void) { return 2; }
const char* reid_last_error(void) { return reid::g_err; }
uint64_t reid_launch_count(void) { return reid::g_launches; }

int reid_scan_counts(const int32_t* cnt, int64_t n, int64_t* ptr_out, int64_t* stats_out, void* stream) {
  REID_CHECK_ARG(n >= 0 && ptr_out, "reid_scan_counts: bad arguments");
  REID_CHECK_ARG(n == 0 || cnt, "reid_scan_counts: cnt is NULL");
  reid::scan_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(cnt, n, ptr_out, stats_out);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
