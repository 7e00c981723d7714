"""torchrun worker of tests/test_gpu_comm.py::test_torchrun_two_ranks_sliced_aggregator_path."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuvs_rag_b200 as b2  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    N, D, Q, k = 200_003, 96, 777, 10
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, D, generator=g).to(torch.float16)      # the same corpus on every rank (seeded)
    q = torch.randn(Q, D, generator=g).to(torch.float16)
    planted = torch.tensor([3, 100_001, 100_002, 200_002])
    q[:4] = x[planted]
    s, e = b2.partition_even(N, world)[rank]
    grm = b2.GPUResourceManager(devices=[local])
    ibc = b2.IndexBuildingCoordinator(grm)
    sra = b2.SearchResultAggregator(grm)
    part = b2.EmbeddingPart(local, x[s:e].cuda(), s, e)
    res = ibc._build_single_index(part, b2.IndexBuildConfig("brute_force", {}, parallel_build=False, max_retries=0))
    assert res.success, res.error_message
    qb, qe = b2.partition_even(Q, world)[rank]
    cfg = b2.SearchConfig(k=k, search_params={"result_layout": "sliced", "num_queries_total": Q})
    out = sra.perform_distributed_search(q[qb:qe].contiguous(), {local: res.index}, cfg)   # HOST slice in
    # float64 truth for this rank's slice
    full = ((x.double()[None, :, :] - q[qb:qe].double()[:, None, :]) ** 2).sum(2) if (qe - qb) * N < 5e7 else None
    from oracle.exact import exact_knn
    _, ti = exact_knn(x.float(), q[qb:qe].float(), k)
    got = torch.from_numpy(out.final_indices)
    match = (got == ti).float().mean().item()
    assert match > 0.999, match
    lo, hi = qb, min(qe, 4)
    if hi > lo:
        assert (got[:hi - lo, 0] == planted[lo:hi]).all()
    # replicated layout through the same aggregator (library all-gather + merge on every rank)
    cfg2 = b2.SearchConfig(k=k, search_params={"collect_gpu_results": False})
    out2 = sra.perform_distributed_search(q, {local: res.index}, cfg2)
    _, t_all = exact_knn(x.float(), q.float(), k)
    assert (torch.from_numpy(out2.final_indices) == t_all).float().mean().item() > 0.999
    dist.barrier()
    print("SHARDED_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
