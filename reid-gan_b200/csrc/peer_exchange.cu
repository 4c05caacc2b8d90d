// Exchange steps of the row-sharded pass over NVLink peer memory -- plain stores from this GPU's kernels
// into buffers that live on the other GPUs of the box (mapped into this process by the host layer:
// torch.distributed._symmetric_memory hands out one base pointer per rank).  No NCCL call, no padding on the wire:
// only the valid entries of a list travel.
//
// reid_peer_push_lists: the hand-over of the tile-sharded search.  Rank `me` has appended the survivors of ITS tiles
// to partial lists of ALL rows (slot == row, W blocks of B rows, `cap` entries each).  The rows of block w belong
// to rank w, which keeps W partial lists per own row: list q of local row r at (q * B + r).  So block w of this
// rank's lists goes to list `me` of rank w -- the all-to-all of sharded.knn_search_tiles, minus the unused tail
// of every list (the lists are sized for 4x their expected length).
#include "common.cuh"

namespace reid {

__global__ void __launch_bounds__(256) peer_push_lists_kernel(const unsigned long long* __restrict__ part,
                                                              const int32_t* __restrict__ part_cnt, int W, int64_t B,
                                                              int cap, int me, const unsigned long long* __restrict__ peer_base,
                                                              int64_t cnt_offset_bytes) {
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // w * B + r
  if (slot >= (int64_t)W * B) return;
  const int w = (int)(slot / B);
  const int64_t r = slot - (int64_t)w * B;
  const int lane = lane_id();
  const int c_true = part_cnt[slot];
  const int c = c_true < cap ? c_true : cap;
  unsigned char* base = reinterpret_cast<unsigned char*>(peer_base[w]);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(base) + ((int64_t)me * B + r) * cap;
  const unsigned long long* src = part + slot * cap;
  for (int t = lane; t < c; t += 32) dst[t] = src[t];
  // the TRUE count travels (a list that overflowed must be recognised by the owner's certificate)
  if (lane == 0) reinterpret_cast<int32_t*>(base + cnt_offset_bytes)[(int64_t)me * B + r] = c_true;
  __threadfence_system();
}

}  // namespace reid

extern "C" {

int reid_peer_push_lists(const uint64_t* part, const int32_t* part_cnt, int world, int64_t block_rows, int cap, int me,
                         const uint64_t* peer_base, int64_t cnt_offset_bytes, void* stream) {
  using namespace reid;
  REID_CHECK_ARG(part && part_cnt && peer_base, "reid_peer_push_lists: NULL pointer");
  REID_CHECK_ARG(world >= 1 && block_rows >= 0 && cap >= 1 && me >= 0 && me < world && cnt_offset_bytes >= 0 &&
                     cnt_offset_bytes % 4 == 0,
                 "reid_peer_push_lists: bad arguments");
  const int64_t slots = (int64_t)world * block_rows;
  if (slots == 0) return REID_OK;
  peer_push_lists_kernel<<<(unsigned)((slots + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const unsigned long long*)part, part_cnt, world, block_rows, cap, me, (const unsigned long long*)peer_base,
      cnt_offset_bytes);
  REID_LAUNCH_CHECK();
  return REID_OK;
}
}
