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

// One CTA scans up to a few hundred thousand counts: every thread owns a contiguous chunk, the 1024
// chunk totals are scanned with warp shuffles (two levels), then the chunk is re-walked.
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ cnt, int64_t n,
                                                           int64_t* __restrict__ ptr, int64_t* __restrict__ stats) {
  __shared__ int64_t warp_tot[32];
  __shared__ int32_t warp_max_s[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int64_t chunk = (n + 1023) / 1024;
  const int64_t a = min(n, (int64_t)t * chunk), b = min(n, a + chunk);
  int64_t s = 0;
  int32_t mx = 0;
  for (int64_t i = a; i < b; ++i) {
    const int32_t c = cnt[i];
    s += c;
    mx = max(mx, c);
  }
  int64_t inc = s;  // inclusive scan inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t v = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += v;
  }
  mx = warp_max(mx);
  if (lane == 31) warp_tot[w] = inc;
  if (lane == 0) warp_max_s[w] = mx;
  __syncthreads();
  if (w == 0) {
    int64_t v = warp_tot[lane];
    int64_t iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(kFull, iv, o);
      if (lane >= o) iv += u;
    }
    warp_tot[lane] = iv - v;  // exclusive offset of each warp
    const int32_t m = warp_max(warp_max_s[lane]);
    if (lane == 31) {
      ptr[n] = iv;
      if (stats) {
        stats[0] = iv;
        stats[1] = m;
      }
    }
  }
  __syncthreads();
  int64_t run = warp_tot[w] + inc - s;
  for (int64_t i = a; i < b; ++i) {
    ptr[i] = run;
    run += cnt[i];
  }
}

}  // namespace reid

extern "C" {

int reid_abi_version(void) { return 1; }
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
