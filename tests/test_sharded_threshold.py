"""The threshold exchange of the sharded exact search (csrc/comm.cu: tau_union_cb, csrc/merge.cu:
union_kth_kernel), restated in numpy: every shard samples every (stride * G)-th 256-row tile, hands
over its k best sampled scores, and the k-th best of the UNION is everybody's threshold.  It must
bound the global k-th best score from above (so no shard drops a true neighbour), and it is as tight
as a single GPU's threshold from a stride-`stride` sample of the whole database."""
import numpy as np
import pytest


def shard_sample_topk(scores, stride, k):
    tiles = scores.reshape(-1, 256)
    samp = np.sort(tiles[::stride].ravel())[:k]
    return np.concatenate([samp, np.full(k - samp.shape[0], np.inf, samp.dtype)])


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("k", [10, 100])
def test_union_of_sparser_samples_bounds_the_global_kth_score(world, k):
    g = np.random.default_rng(1000 * world + k)
    rows_per_shard, stride = 256 * 1024, 16
    shards = [g.standard_normal(rows_per_shard).astype(np.float32) + 0.05 * r for r in range(world)]
    union = np.concatenate([shard_sample_topk(s, stride * world, k) for s in shards])
    tau = np.nextafter(np.sort(union)[k - 1], np.float32(np.inf))
    everything = np.concatenate(shards)
    kth = np.sort(everything)[k - 1]
    assert tau > kth                                              # inclusive upper bound
    # every global top-k element is found by its shard's pass against tau
    assert all((s < tau).sum() >= (s <= kth).sum() for s in shards)
    # tightness: candidates per query over all shards ~ k * stride * world (what the buffers are sized for)
    n_cand = sum(int((s < tau).sum()) for s in shards)
    assert n_cand <= 4 * k * stride * world
    # the MIN of per-shard k-th scores at the SAME sparse stride is looser (what the pooling buys)
    tau_min = min(np.sort(s.reshape(-1, 256)[::stride * world].ravel())[k - 1] for s in shards)
    assert tau <= np.nextafter(tau_min, np.float32(np.inf))
