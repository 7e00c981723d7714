// IVF-Flat / IVF-PQ internals (ivf.cu), reached through the C ABI in api.cu.
#pragma once
#include "common.h"

namespace b2vs {
int ivf_search(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
               const b2vs_search_params& sp, float* out_d, int64_t* out_i, cudaStream_t st);
void ivf_fill_info(const b2vs_index* index, b2vs_index_info* info);
void ivf_last_stats(const b2vs_index* index, b2vs_search_stats* stats);
void ivf_destroy(b2vs_index* index);
int cosine_rows(int dtype, int dim, const void* db, int64_t n, cudaStream_t st, DevBuf* buf);
int check_matrix_args(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                      b2vs_index** out);
}  // namespace b2vs
