"""Config C4 of BASELINE.json on N GPUs (one process per GPU under torchrun): IVF-PQ, 16384 lists
per shard, M = 64 x 8 bit, 100M x 128 fp16 row-sharded (12.5M rows per GPU at N = 8), 10K-query
batches, n_probes 64 + refine 4, k = 10, NCCL all-gather + GPU merge; recall@10 against the exact
sharded search of the same corpus.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import cuvs_rag_b200 as b2

ROWS_PER_GPU = int(os.environ.get("C4_ROWS_PER_GPU", 12_500_000))
DIM, NLIST, M, NPROBE, REFINE, K, NQ = 128, 16384, 64, 64, 4, 10, 10_000


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # clustered mixture shared by all shards (same component centres), rows differ per shard
    gc = torch.Generator(device=dev).manual_seed(99)
    cent = torch.randn(NLIST, DIM, generator=gc, device=dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.empty((ROWS_PER_GPU, DIM), dtype=torch.float16, device=dev)
    for s in range(0, ROWS_PER_GPU, 1 << 20):
        e = min(ROWS_PER_GPU, s + (1 << 20))
        lab = torch.randint(0, NLIST, (e - s,), generator=g, device=dev)
        x[s:e] = (cent[lab] + 0.42 * torch.randn((e - s, DIM), generator=g, device=dev)).to(torch.float16)
    gq = torch.Generator(device=dev).manual_seed(4321)          # same queries on every rank
    lab = torch.randint(0, NLIST, (NQ,), generator=gq, device=dev)
    q = (cent[lab] + 0.42 * torch.randn((NQ, DIM), generator=gq, device=dev)).to(torch.float16)
    id_offset = rank * ROWS_PER_GPU

    torch.cuda.synchronize(); t0 = time.time()
    ix = b2.NativeIndex.ivf_pq(x, NLIST, M, id_offset=id_offset, kmeans_iters=10)
    torch.cuda.synchronize(); build_s = time.time() - t0

    def gather_merge(d, i):
        if world == 1:
            return d, i
        gd = torch.empty((world,) + tuple(d.shape), dtype=d.dtype, device=dev)
        gi = torch.empty((world,) + tuple(i.shape), dtype=i.dtype, device=dev)
        dist.all_gather_into_tensor(gd, d.contiguous()); dist.all_gather_into_tensor(gi, i.contiguous())
        return b2.merge_topk(gd, gi, K)

    def step():
        d, i = ix.search(q, K, n_probes=NPROBE, refine_ratio=REFINE)
        return gather_merge(d, i)

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record()
    for _ in range(steps):
        md, mi = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    bs = torch.tensor([build_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(bs, op=dist.ReduceOp.MAX)
    # ground truth: exact sharded search of the same corpus
    flat = b2.NativeIndex.flat(x, id_offset=id_offset)
    td, ti = gather_merge(*flat.search(q, K))
    hits = (mi[:, :, None] == ti[:, None, :]).any(2).float().mean().item()
    if rank == 0:
        print(json.dumps({"config": "C4 IVF-PQ nlist=16384 M=64x8b, 128-d fp16, 10K-query batches, k=10",
                          "n_gpus": world, "rows_total": ROWS_PER_GPU * world, "n_probes": NPROBE,
                          "refine_ratio": REFINE, "ms_per_batch": float(ms.item()),
                          "qps": NQ / float(ms.item()) * 1e3, "recall_at_10": hits,
                          "build_s_max_over_ranks": float(bs.item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
