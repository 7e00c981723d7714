// B2VS_METRIC_COSINE support: unit-norm copies of rows / queries and the final 1 - similarity.
// The reference's CPU baseline is scikit-learn NearestNeighbors(metric='cosine', algorithm='brute')
// (Attempt_1/VectorSearch_QuestionRetrieval.ipynb:L878); its GPU embeddings are F.normalize'd
// (generate_embeddings.py:100-105), for which cosine == 1 - inner product.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.h"

namespace b2vs {
namespace {

template <typename T> __device__ __forceinline__ float cvt_in(T v);
template <> __device__ __forceinline__ float cvt_in<float>(float v) { return v; }
template <> __device__ __forceinline__ float cvt_in<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float cvt_in<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// one warp per row: x / max(||x||_2, 1e-12) in fp32, one rounding into T (F.normalize semantics)
template <typename T>
__global__ void unit_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t n, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n; r += nwarps) {
    const T* row = src + r * dim;
    float ss = 0.f;
    for (int j = lane; j < dim; j += 32) { const float v = cvt_in<T>(row[j]); ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    T* out = dst + r * dim;
    for (int j = lane; j < dim; j += 32) out[j] = cvt_out<T>(cvt_in<T>(row[j]) * scale);
  }
}

// similarity (descending, -inf = missing) -> cosine distance (ascending, +inf = missing)
__global__ void cosine_fixup_kernel(float* __restrict__ d, int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float s = d[i];
  d[i] = (s == -INFINITY) ? INFINITY : 1.f - s;
}

}  // namespace

int launch_unit_rows(const void* src, void* dst, int dtype, int64_t n, int dim, cudaStream_t st) {
  if (n <= 0) return B2VS_OK;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 8), 148 * 16));
  switch (dtype) {
    case B2VS_F32:
      unit_rows_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), static_cast<float*>(dst), n, dim);
      break;
    case B2VS_F16:
      unit_rows_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(src), static_cast<__half*>(dst), n, dim);
      break;
    default:
      unit_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src),
                                                             static_cast<__nv_bfloat16*>(dst), n, dim);
      break;
  }
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_cosine_fixup(float* d, int64_t total, cudaStream_t st) {
  if (total <= 0) return B2VS_OK;
  cosine_fixup_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, st>>>(d, total);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

}  // namespace b2vs
