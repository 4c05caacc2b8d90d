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

// One CTA scans up to a few hundred thousand counts: every thread owns a contiguous
// chunk, the 1024 chunk totals are scanned in shared memory, the chunk is re-walked.
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ cnt, int64_t n,
                                                           int64_t* __restrict__ ptr, int64_t* __restrict__ stats) {
  __shared__ int64_t part[1024];
  __shared__ int32_t pmax[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + 1023) / 1024;
  const int64_t a = min(n, (int64_t)t * chunk), b = min(n, a + chunk);
  int64_t s = 0;
  int32_t mx = 0;
  for (int64_t i = a; i < b; ++i) {
    int32_t c = cnt[i];
    s += c;
    mx = max(mx, c);
  }
  part[t] = s;
  pmax[t] = mx;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    int64_t v = t >= o ? part[t - o] : 0;
    int32_t m = t >= o ? pmax[t - o] : 0;
    __syncthreads();
    part[t] += v;
    pmax[t] = max(pmax[t], m);
    __syncthreads();
  }
  int64_t run = part[t] - s;
  for (int64_t i = a; i < b; ++i) {
    ptr[i] = run;
    run += cnt[i];
  }
  if (t == 1023) {
    ptr[n] = part[1023];
    if (stats) {
      stats[0] = part[1023];
      stats[1] = pmax[1023];
    }
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
