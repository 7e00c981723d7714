// Per-row streaming top-k used by the fused distance kernels.
//
// One thread owns one query row.  It keeps a threshold `tau` (an upper bound on the row's
// current k-th best score) and appends every (score, id) with score < tau to the row's
// candidate buffer (capacity kCap keys, global memory / L2 resident).  When a buffer is about
// to overflow the owning warp sorts it cooperatively (register bitonic network across the 32
// lanes), keeps the k best, and tightens tau.  Keys are 64-bit: order-preserving float bits in
// the high word, the row id inside the shard in the low word, so a plain unsigned compare
// orders by (score, id) and ties resolve to the smaller id.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2vs {

using u64 = unsigned long long;

constexpr int kSortE = 8;            // keys per lane in the warp sort
constexpr int kCap = 32 * kSortE;    // candidate buffer capacity per row (256)
constexpr int kMaxFusedK = 128;      // largest k served by the fused path
constexpr u64 kKeyInf = ~0ull;

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t b = o ^ ((o & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu);
  return __uint_as_float(b);
}
__device__ __forceinline__ u64 pack_key(float score, uint32_t id) {
  return (static_cast<u64>(f2ord(score)) << 32) | id;
}
__device__ __forceinline__ float key_score(u64 key) { return ord2f(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ uint32_t key_id(u64 key) { return static_cast<uint32_t>(key); }

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int mask) {
  uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, mask);
  hi = __shfl_xor_sync(0xffffffffu, hi, mask);
  return (static_cast<u64>(hi) << 32) | lo;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
  uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return (static_cast<u64>(hi) << 32) | lo;
}

// Ascending bitonic sort of 32*E keys held E per lane, blocked layout: lane L holds the
// elements [L*E, L*E+E).  Strides below E are register compare-exchanges, the rest shuffles.
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(u64 (&key)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= E) {
        // partner in another lane: keep own key iff (own < other) == (this position keeps the min)
        const int lm = stride / E;
        const bool lower = (lane & lm) == 0;
        const bool up = size >= N ? true : (((lane * E) & size) == 0);   // size >= E here: lane-only
        const bool keep_min = (lower == up);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const u64 other = shfl_xor_u64(key[e], lm);
          key[e] = ((key[e] < other) == keep_min) ? key[e] : other;
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & stride) == 0) {
            const bool up = (((lane * E + e) & size) == 0);
            const u64 a = key[e], b = key[e + stride];
            const bool swap = (a < b) != up;
            key[e] = swap ? b : a;
            key[e + stride] = swap ? a : b;
          }
        }
      }
    }
  }
}

// Bitonic MERGE of a sequence that is already bitonic (e.g. elementwise min of an ascending and
// a descending run): only the last `size = N` pass of the sort network.
template <int E>
__device__ __forceinline__ void warp_bitonic_merge(u64 (&key)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int stride = N >> 1; stride > 0; stride >>= 1) {
    if (stride >= E) {
      const int lm = stride / E;
      const bool lower = (lane & lm) == 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const u64 other = shfl_xor_u64(key[e], lm);
        key[e] = ((key[e] < other) == lower) ? key[e] : other;
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & stride) == 0) {
          const u64 a = key[e], b = key[e + stride];
          const bool swap = !(a < b);
          key[e] = swap ? b : a;
          key[e + stride] = swap ? a : b;
        }
      }
    }
  }
}

// Warp-cooperative compaction of one row buffer: sort the first n keys, write the best
// min(n, k) back in ascending order (to `dst`, which may be the buffer itself) and return the
// new threshold (+inf while fewer than k candidates exist); min(n, k) keys are kept.  All 32
// lanes must call it.  The sort network is sized to n (64 / 128 / 256 keys): short buffers -
// small k, whose buffers are compacted early, and most final flushes - cost a fraction of the
// full network.  Deliberately not inlined: one copy of each network keeps the fused kernel's
// hot loop inside the instruction cache.
template <int E>
__device__ __forceinline__ float compact_row_impl(const u64* buf, u64* dst, int n, int k, int lane) {
  u64 key[E];
  const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(buf + lane * E);
#pragma unroll
  for (int e = 0; e < E; e += 2) {
    const int idx = lane * E + e;
    ulonglong2 v = make_ulonglong2(kKeyInf, kKeyInf);
    if (idx < n) v = __ldcg(b2 + e / 2);
    key[e] = v.x;
    key[e + 1] = (idx + 1 < n) ? v.y : kKeyInf;
  }
  warp_bitonic_sort<E>(key, lane);
  const int m = n < k ? n : k;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int idx = lane * E + e;
    if (idx < m) __stcg(dst + idx, key[e]);
  }
  u64 kth = kKeyInf;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (lane * E + e == k - 1) kth = key[e];
  kth = shfl_u64(kth, (k - 1) / E);
  return (n >= k) ? key_score(kth) : __int_as_float(0x7f800000);
}
static __device__ __noinline__ float compact_row_256(const u64* buf, u64* dst, int n, int k, int lane) {
  return compact_row_impl<8>(buf, dst, n, k, lane);
}
static __device__ __noinline__ float compact_row_128(const u64* buf, u64* dst, int n, int k, int lane) {
  return compact_row_impl<4>(buf, dst, n, k, lane);
}
static __device__ __noinline__ float compact_row_64(const u64* buf, u64* dst, int n, int k, int lane) {
  return compact_row_impl<2>(buf, dst, n, k, lane);
}
// n and k are warp-uniform.  The chosen network must hold both the n keys and position k - 1.
__device__ __forceinline__ float compact_row(const u64* buf, u64* dst, int n, int k, int lane) {
  const int need = n > k ? n : k;
  if (need <= 64) return compact_row_64(buf, dst, n, k, lane);
  if (need <= 128) return compact_row_128(buf, dst, n, k, lane);
  return compact_row_256(buf, dst, n, k, lane);
}

// Fill level at which a row buffer is compacted.  Small k compacts early: the threshold tightens
// sooner (fewer candidates overall) and the sorts run on the 64- / 128-key networks.
__host__ __device__ inline int compact_trigger(int k) {
  return (k <= 16 ? 64 : (k <= 48 ? 128 : kCap)) - 32;
}

}  // namespace b2vs
