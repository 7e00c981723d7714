"""N>1 host logic on CPU: two gloo ranks shard a corpus by partition_even, each answers for its
shard (oracle exact search standing in for the GPU kernel), ids are made global with the shard
start, lists are all-gathered and merged; the result must equal the single-shard answer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), CUDA_VISIBLE_DEVICES="")
    import torch.distributed as dist
    import cuvs_rag_b200 as b2
    from oracle.exact import exact_knn

    grm = b2.GPUResourceManager()
    comm = grm.get_communicator()            # initialises gloo from the torchrun-style env
    assert comm is not None and grm.get_rank_info() == (rank, world)
    g = torch.Generator().manual_seed(11)
    db = torch.randn(1001, 24, generator=g)  # uneven shards: 501 + 500
    q = torch.randn(9, 24, generator=g)
    start, end = b2.partition_even(db.shape[0], world)[rank]
    d, i = exact_knn(db[start:end], q, 5)
    i = i + start                            # EmbeddingPart.start_index, not rank * len(part)
    from search_result_aggregator import allgather_and_merge
    md, mi = allgather_and_merge(d.unsqueeze(0), i.unsqueeze(0), 5, False, world)
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), mi.numpy())
    np.save(os.path.join(out_dir, f"d_{rank}.npy"), md.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_search_matches_single_shard(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle.exact import exact_knn
    g = torch.Generator().manual_seed(11)
    db = torch.randn(1001, 24, generator=g)
    q = torch.randn(9, 24, generator=g)
    d, i = exact_knn(db, q, 5)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"ids_{r}.npy"), i.numpy())
        np.testing.assert_allclose(np.load(tmp_path / f"d_{r}.npy"), d.numpy(), rtol=1e-5, atol=1e-5)
