// Encoder hand-off: the bi-encoder's hidden states become query rows on the
// device, with no host hop between the encoder and the search.
//
// Replaces, in the reference's query path (Latest/cuVS-2-gpu/old/generate_embeddings.py):
//   :11-21   last_token_pool(last_hidden_states, attention_mask)
//              left_padding = attention_mask[:, -1].sum() == batch
//              left padded  -> hidden[:, -1]
//              otherwise    -> hidden[b, attention_mask[b].sum() - 1]   (index -1 wraps to T - 1)
//   :100-103 F.normalize(embeddings, p=2, dim=1)          x / max(||x||_2, 1e-12)
//   :105     embeddings.cpu().numpy()  (and the later .to(device) before the search,
//            cuvs-2gpu-main.ipynb cell 16) -- the hop this file removes.
// The 384-d MiniLM embeddings of the Attempt_1 notebooks come from sentence-transformers
// (prepare_dataset.py:149 `model.encode`), whose pooling module is the masked mean
//   sum_t mask[b,t] * h[b,t,:] / max(sum_t mask[b,t], 1e-9)
// (un-vendored dependency, restated from its published definition); pooling = 1 selects it.
//
// Both kernels are HBM-bound and tiny next to the encoder: arithmetic in fp32 from the hidden
// states' own dtype, one rounding into the output dtype.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.h"

namespace b2vs {

namespace {

constexpr int kPoolThreads = 256;

template <typename T> __device__ __forceinline__ float load_f32(const T* p);
template <> __device__ __forceinline__ float load_f32<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_f32<__half>(const __half* p) {
  return __half2float(__ldg(p));
}
template <> __device__ __forceinline__ float load_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}

template <typename T> __device__ __forceinline__ void store_f32(T* p, float v);
template <> __device__ __forceinline__ void store_f32<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_f32<__half>(__half* p, float v) {
  *p = __float2half_rn(v);
}
template <> __device__ __forceinline__ void store_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ long long block_sum_ll(long long v, long long* smem) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // smem may still be read by a previous reduction
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  long long t = 0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += smem[w];
  return t;
}

__device__ __forceinline__ float block_sum_f32(float v, float* smem) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += smem[w];
  return t;
}

// info[0] = 1 iff every sequence's last position is attended (the reference's left-padding test);
// info[1 + b] = number of attended positions of sequence b.  One CTA per sequence + one for the flag.
__global__ void __launch_bounds__(kPoolThreads)
mask_info_kernel(const long long* __restrict__ mask, int batch, int seq_len, int* __restrict__ info) {
  __shared__ long long red[kPoolThreads / 32];
  const int b = blockIdx.x;
  long long s = 0;
  if (b < batch) {
    for (int t = threadIdx.x; t < seq_len; t += blockDim.x)
      s += mask[static_cast<size_t>(b) * seq_len + t];
    s = block_sum_ll(s, red);
    if (threadIdx.x == 0) info[1 + b] = static_cast<int>(s);
  } else {
    for (int r = threadIdx.x; r < batch; r += blockDim.x)
      s += mask[static_cast<size_t>(r) * seq_len + (seq_len - 1)];
    s = block_sum_ll(s, red);
    if (threadIdx.x == 0) info[0] = (s == static_cast<long long>(batch)) ? 1 : 0;
  }
}

// pooled[b, d] in fp32.  grid = (ceil(dim / 256), batch); a thread owns one dimension, a warp
// reads 32 consecutive elements of a token row.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_kernel(const T* __restrict__ hidden, const long long* __restrict__ mask,
            const int* __restrict__ info, int seq_len, int dim, int pooling,
            float* __restrict__ pooled) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  const T* seq = hidden + static_cast<size_t>(b) * seq_len * dim;
  float v;
  if (pooling == 0) {
    int idx = seq_len - 1;
    if (mask != nullptr && info[0] == 0) {
      idx = info[1 + b] - 1;
      if (idx < 0) idx += seq_len;   // python's negative index: an all-zero mask row reads the last token
    }
    v = load_f32(seq + static_cast<size_t>(idx) * dim + d);
  } else {
    float acc = 0.f;
    if (mask == nullptr) {
      for (int t = 0; t < seq_len; ++t) acc += load_f32(seq + static_cast<size_t>(t) * dim + d);
      v = acc / fmaxf(static_cast<float>(seq_len), 1e-9f);
    } else {
      const long long* m = mask + static_cast<size_t>(b) * seq_len;
      for (int t = 0; t < seq_len; ++t) {
        const float w = static_cast<float>(m[t]);   // uniform across the warp: one broadcast load
        if (w != 0.f) acc += w * load_f32(seq + static_cast<size_t>(t) * dim + d);
      }
      v = acc / fmaxf(static_cast<float>(info[1 + b]), 1e-9f);
    }
  }
  pooled[static_cast<size_t>(b) * dim + d] = v;
}

// out[b, :] = pooled[b, :] / max(||pooled[b, :]||_2, 1e-12)  (or a plain cast).  One CTA per row.
template <typename O>
__global__ void __launch_bounds__(kPoolThreads)
normalize_rows_kernel(const float* __restrict__ pooled, int dim, int normalize, O* __restrict__ out) {
  __shared__ float red[kPoolThreads / 32];
  const float* row = pooled + static_cast<size_t>(blockIdx.x) * dim;
  O* dst = out + static_cast<size_t>(blockIdx.x) * dim;
  float scale = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) ss += row[d] * row[d];
    ss = block_sum_f32(ss, red);
    scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  for (int d = threadIdx.x; d < dim; d += blockDim.x) store_f32(dst + d, row[d] * scale);
}

bool valid_dtype3(int d) { return d == B2VS_F32 || d == B2VS_F16 || d == B2VS_BF16; }

}  // namespace

}  // namespace b2vs

using namespace b2vs;

extern "C" int b2vs_pool_normalize(int dev, int dtype, const void* hidden, int batch, int seq_len,
                                   int dim, const int64_t* attention_mask, int pooling,
                                   int normalize, int out_dtype, void* out, void* stream) {
  B2VS_CHECK(hidden != nullptr && out != nullptr, B2VS_EINVAL, "hidden / out pointer is NULL");
  B2VS_CHECK(valid_dtype3(dtype) && valid_dtype3(out_dtype), B2VS_EINVAL, "unknown dtype %d / %d",
             dtype, out_dtype);
  B2VS_CHECK(batch >= 1 && batch <= 65535, B2VS_EINVAL, "batch=%d outside [1, 65535]", batch);
  B2VS_CHECK(seq_len >= 1 && dim >= 1, B2VS_EINVAL, "seq_len and dim must be positive (%d, %d)",
             seq_len, dim);
  B2VS_CHECK(pooling == B2VS_POOL_LAST_TOKEN || pooling == B2VS_POOL_MEAN, B2VS_EINVAL,
             "unknown pooling %d", pooling);
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // scratch: pooled fp32 rows + mask info, per calling thread (grow-only)
  static thread_local DevBuf* ws = nullptr;
  static thread_local int ws_dev = -1;
  if (ws == nullptr || ws_dev != dev) {
    if (ws) { ws->release(); delete ws; }
    ws = new DevBuf();
    ws_dev = dev;
  }
  const size_t pooled_bytes = static_cast<size_t>(batch) * dim * sizeof(float);
  const size_t info_off = static_cast<size_t>(round_up(static_cast<int64_t>(pooled_bytes), 256));
  B2VS_TRY(ws->reserve(info_off + (static_cast<size_t>(batch) + 1) * sizeof(int)));
  float* pooled = ws->as<float>();
  int* info = reinterpret_cast<int*>(ws->as<char>() + info_off);
  const long long* mask = reinterpret_cast<const long long*>(attention_mask);
  if (mask != nullptr) {
    mask_info_kernel<<<batch + 1, kPoolThreads, 0, st>>>(mask, batch, seq_len, info);
    B2VS_CUDA(cudaGetLastError());
  }
  const dim3 grid(static_cast<unsigned>(ceil_div(dim, kPoolThreads)), static_cast<unsigned>(batch));
  switch (dtype) {
    case B2VS_F32:
      pool_kernel<float><<<grid, kPoolThreads, 0, st>>>(static_cast<const float*>(hidden), mask, info,
                                                        seq_len, dim, pooling, pooled);
      break;
    case B2VS_F16:
      pool_kernel<__half><<<grid, kPoolThreads, 0, st>>>(static_cast<const __half*>(hidden), mask,
                                                         info, seq_len, dim, pooling, pooled);
      break;
    default:
      pool_kernel<__nv_bfloat16><<<grid, kPoolThreads, 0, st>>>(
          static_cast<const __nv_bfloat16*>(hidden), mask, info, seq_len, dim, pooling, pooled);
      break;
  }
  B2VS_CUDA(cudaGetLastError());
  switch (out_dtype) {
    case B2VS_F32:
      normalize_rows_kernel<float><<<batch, kPoolThreads, 0, st>>>(pooled, dim, normalize,
                                                                   static_cast<float*>(out));
      break;
    case B2VS_F16:
      normalize_rows_kernel<__half><<<batch, kPoolThreads, 0, st>>>(pooled, dim, normalize,
                                                                    static_cast<__half*>(out));
      break;
    default:
      normalize_rows_kernel<__nv_bfloat16><<<batch, kPoolThreads, 0, st>>>(
          pooled, dim, normalize, static_cast<__nv_bfloat16*>(out));
      break;
  }
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}
