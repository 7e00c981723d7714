// Warp-resident top-k selection shared by the IVF select kernels (ivf_scan.cu) and the dense
// score-matrix select of the exact engine (merge.cu: row_select_kernel).
//   WarpTopK     a sorted 128-key list held 4 keys per lane; offer() folds up to 32 new keys in
//   StagedTopK   WarpTopK behind a small shared-memory stage: keys that pass the threshold filter
//                are compacted (ballot + prefix) into the stage and folded in 32 at a time, so the
//                sort + merge network runs once per 32 ACCEPTED keys instead of once per 32 keys
//                looked at (a stream of n keys accepts ~k ln(n/k) of them).
#pragma once
#include "topk.cuh"

namespace b2vs {

constexpr int kListE = 4;  // warp-resident sorted list: 32 * 4 = 128 keys = kMaxFusedK

struct WarpTopK {
  u64 acc[kListE];
  float tau;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < kListE; ++e) acc[e] = kKeyInf;
    tau = __int_as_float(0x7f800000);
  }
  // Each lane offers at most one candidate key (kKeyInf = none). Warp-collective.
  __device__ __forceinline__ void offer(u64 ck, int k, int lane) {
    if (!__any_sync(0xffffffffu, ck != kKeyInf)) return;
    u64 c1[1] = {ck};
    warp_bitonic_sort<1>(c1, lane);
#pragma unroll
    for (int e = 0; e < kListE; ++e) {
      const int i = lane * kListE + e;                       // list element index
      const u64 r = shfl_u64(c1[0], (32 * kListE - 1 - i) & 31);  // reversed candidate run
      if (i >= 32 * kListE - 32) acc[e] = acc[e] < r ? acc[e] : r;
    }
    warp_bitonic_merge<kListE>(acc, lane);
    u64 kth = kKeyInf;
#pragma unroll
    for (int e = 0; e < kListE; ++e)
      if (lane * kListE + e == k - 1) kth = acc[e];
    kth = shfl_u64(kth, (k - 1) / kListE);
    tau = (kth == kKeyInf) ? __int_as_float(0x7f800000) : key_score(kth);
  }
};

constexpr int kStageKeys = 64;   // shared-memory keys per warp behind a StagedTopK

struct StagedTopK {
  WarpTopK tk;
  u64* stage;   // [kStageKeys] shared memory, private to the warp
  int n;        // staged keys (warp-uniform)
  __device__ __forceinline__ void init(u64* stage_mem) {
    tk.init();
    stage = stage_mem;
    n = 0;
  }
  // Every lane passes one key that already passed the caller's filter against tk.tau, or kKeyInf.
  // Warp-collective.  Staged keys that a later, tighter threshold would reject are harmless: the
  // merge drops them.
  __device__ __forceinline__ void push(u64 ck, int k, int lane) {
    const uint32_t m = __ballot_sync(0xffffffffu, ck != kKeyInf);
    if (m == 0) return;
    if (ck != kKeyInf) stage[n + __popc(m & ((1u << lane) - 1u))] = ck;
    n += __popc(m);
    if (n >= 32) {
      __syncwarp();
      const u64 head = stage[lane];
      const u64 tail = (32 + lane < n) ? stage[32 + lane] : kKeyInf;
      __syncwarp();
      if (32 + lane < n) stage[lane] = tail;
      n -= 32;
      tk.offer(head, k, lane);
    }
  }
  __device__ __forceinline__ void flush(int k, int lane) {
    __syncwarp();
    if (n > 0) {
      tk.offer(lane < n ? stage[lane] : kKeyInf, k, lane);
      n = 0;
    }
    __syncwarp();
  }
};

}  // namespace b2vs
