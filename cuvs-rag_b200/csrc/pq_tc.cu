// Host launcher of the grouped IVF-PQ tensor-core scan (pq_tc.cuh); called from ivf.cu.
#include <algorithm>

#include "common.h"
#include "pq_tc.cuh"

namespace b2vs {

bool pq_grouped_supported(int dim, int dsub) {
  return (dsub == 2 || dsub == 4 || dsub == 8) && dim % kBK == 0 && dim <= 4096;
}

template <int DSUB, bool kCbSmem>
static int launch_one(int grid, const CUtensorMap& tm_q, const PqTcParams& pp, cudaStream_t st) {
  B2VS_CUDA(cudaFuncSetAttribute((pq_tc_kernel<DSUB, kCbSmem>),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kPqTcSmemBytes));
  pq_tc_kernel<DSUB, kCbSmem><<<grid, kPqTcThreads, kPqTcSmemBytes, st>>>(tm_q, pp);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_pq_grouped_scan(int dev, const PqGroupedScanArgs& a, cudaStream_t st) {
  CUtensorMap tm_q;
  B2VS_TRY(encode_tmap_2d(&tm_q, a.q_mat, 1, a.q_rows, a.dim, kPqBoxRows));   // 16-row boxes: loads follow the block's real rows
  PqTcParams pp{};
  BfTcParams& p = pp.tc;
  p.beta = a.beta;
  p.k_blocks = a.dim / kBK;
  p.k = 1;
  p.tile_stride = 1;
  p.alpha = a.alpha;
  p.idesc = ptx::make_idesc_f16(1u, kPqM, kPqN);   // M = list rows of a tile, N = query rows of a block
  p.tau_init = a.tau;
  p.big_cand = a.cand;
  p.big_count = a.count;
  p.big_cap = a.cap;
  p.work = static_cast<const int4*>(a.work);
  p.n_work = a.n_work;
  p.row_query = a.row_query;
  p.row_slot = a.row_slot;
  p.seed_all = a.seed_all;
  pp.codes = static_cast<const uint8_t*>(a.codes);
  pp.cb16 = static_cast<const uint32_t*>(a.cb16t);
  pp.row_bias = a.row_bias;
  pp.mp = a.pq_dim;
  pp.n_groups = a.n_groups;
  pp.cb_words = a.pq_dim * 256 * a.dsub / 2;
  pp.debug = env().pq_debug & (7 | 16);
  pp.qres = (p.k_blocks <= 2 && !(env().pq_debug & 8)) ? 1 : 0;   // B2VS_PQ_DEBUG bit 8: streaming query blocks (A/B)
  const int grid = std::max(1, std::min(a.max_work, sm_count(dev)));
  const bool cb_smem = pp.cb_words * 4 <= kPqTcMaxCbBytes;
  if (a.dsub == 2) return cb_smem ? launch_one<2, true>(grid, tm_q, pp, st) : launch_one<2, false>(grid, tm_q, pp, st);
  if (a.dsub == 4) return cb_smem ? launch_one<4, true>(grid, tm_q, pp, st) : launch_one<4, false>(grid, tm_q, pp, st);
  return cb_smem ? launch_one<8, true>(grid, tm_q, pp, st) : launch_one<8, false>(grid, tm_q, pp, st);
}

}  // namespace b2vs
