"""The cross-shard exchange behind the C ABI (b2vs_comm_*, b2vs_search_sharded): NCCL all-gather of
query slices, threshold all-reduce between the passes of the flat search, all-to-all of per-shard
lists + merge of each rank's slice.  Known answers: queries that ARE rows of the corpus must come
back first with their GLOBAL row id, whichever rank owns the row and whichever rank returns the
query's slice (reference semantics: ids + shard start, cuvs-2gpu-main.ipynb:L1803 with the correct
offset of embedding_distribution_manager.py:25; merge improved_multi_gpu_rag.py:266-275)."""
import os
import subprocess
import sys
import threading

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def make(n, d, nq, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g).to(dtype)
    q = torch.randn(nq, d, generator=g).to(dtype)
    return x, q


def test_single_rank_communicator_is_the_plain_search(b2):
    n = b2._native
    x, q = make(70_000, 128, 300, 3)
    q[:10] = x[1000:1010]
    comm = n.Comm.init_rank("cuda:0", 1, 0, n.Comm.unique_id())
    ix = b2.NativeIndex.flat(x.cuda(), id_offset=500)
    d0, i0 = ix.search(q.cuda(), 20)
    d1, i1 = comm.search_sharded(ix, q.cuda(), 300, 20)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)
    assert (i1[:10, 0].cpu() == torch.arange(1500, 1510)).all()
    # host-buffer variant: H2D + search + D2H inside the call
    d2, i2 = comm.search_sharded(ix, q, 300, 20)
    assert not d2.is_cuda and torch.equal(i2, i0.cpu()) and torch.equal(d2, d0.cpu())
    # the other collectives degenerate to copies / a local merge
    qa = comm.allgather_queries(q.cuda(), 300)
    assert torch.equal(qa, q.cuda())
    md, mi = comm.exchange_merge_topk(d0, i0, 10)
    assert torch.equal(mi, i0[:, :10])
    comm.destroy()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NVLink exchange)")
@pytest.mark.parametrize("kind,k", [("flat", 100), ("flat", 10), ("ivf_flat", 10)])
def test_two_ranks_in_one_process_match_the_unsharded_index(b2, kind, k):
    """Thread-per-GPU mode (b2vs_comm_init_all): two shards, uneven split, ragged query slices."""
    from oracle.exact import topk_parity_report
    n = b2._native
    N, D, Q = 300_001, 128, 1001          # odd sizes: uneven shards, ragged slices
    x, q = make(N, D, Q, 11)
    planted = torch.tensor([0, 7, 150_000, 150_001, 299_999, 300_000])
    q[:6] = x[planted]
    parts = b2.partition_even(N, 2)
    comms = n.Comm.init_all(["cuda:0", "cuda:1"])
    idx = []
    for r, (s, e) in enumerate(parts):
        xs = x[s:e].to(f"cuda:{r}")
        idx.append(b2.NativeIndex.flat(xs, id_offset=s) if kind == "flat"
                   else b2.NativeIndex.ivf_flat(xs, 64, id_offset=s, kmeans_iters=5))
    outs, errs = [None, None], [None, None]

    def worker(r):
        try:
            torch.cuda.set_device(r)
            b, e = comms[r].query_slice(Q)
            kw = {} if kind == "flat" else {"n_probes": 64}
            # twice: the second call runs with warm workspaces and a registered index
            for _ in range(2):
                outs[r] = comms[r].search_sharded(idx[r], q[b:e].contiguous().to(f"cuda:{r}"), Q, k, **kw)
            torch.cuda.synchronize(r)
        except Exception as exc:  # noqa: BLE001
            errs[r] = exc
    th = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert errs == [None, None], errs
    d = torch.cat([outs[0][0].cpu(), outs[1][0].cpu()])
    i = torch.cat([outs[0][1].cpu(), outs[1][1].cpu()])
    assert d.shape == (Q, k)
    assert (i[:6, 0] == planted).all()
    rep = topk_parity_report(d, i, x.float(), q.float(), k)
    assert rep["ok"], rep
    for c in comms:
        c.destroy()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (one process per GPU)")
def test_torchrun_two_ranks_sliced_aggregator_path():
    """One process per GPU under torchrun: SearchResultAggregator(result_layout='sliced') ->
    b2vs_search_sharded_host, communicator bootstrapped through torch.distributed."""
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29641",
         os.path.join(ROOT, "tests", "_sharded_worker.py")],
        capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert r.stdout.count("SHARDED_OK") == 2, r.stdout[-2000:]
