"""torchrun A/B of B2VS_* settings on the sharded exact step (C2: 10M x 768 bf16 rows over the ranks,
10K-query batches, k = 100): device-timed ms per step, max over ranks.
usage: torchrun ... tools/sharded_ab.py "A=1&B=2;C=3;..." [rows_total]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import cuvs_rag_b200 as b2
from cuvs_rag_b200 import _native

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
settings = sys.argv[1].split(";") if len(sys.argv) > 1 else [""]
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
s, e = b2.partition_even(n_total, world)[rank]
g = torch.Generator(device="cuda").manual_seed(100 + rank)
x = torch.randn(e - s, 768, device="cuda", generator=g).to(torch.bfloat16)
gq = torch.Generator(device="cuda").manual_seed(7)
q = torch.randn(10_000, 768, device="cuda", generator=gq).to(torch.bfloat16)      # same on every rank
comm = _native.Comm.from_torch_distributed(torch.device("cuda", local))
ix = b2.NativeIndex.flat(x, id_offset=s)
qb, qe = comm.query_slice(10_000)
q_slice = q[qb:qe].contiguous()
touched = set()
ref = None
for sname in settings:
    for kname in touched:
        os.environ.pop(kname, None)
    touched = set()
    for kv in filter(None, sname.split("&")):
        kname, v = kv.split("=", 1)
        os.environ[kname] = v
        touched.add(kname)
    _native.reload_env()
    for _ in range(3):
        d, i = comm.search_sharded(ix, q_slice, 10_000, 100)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        comm.search_sharded(ix, q_slice, 10_000, 100)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    same = True if ref is None else bool((i == ref).all().item())
    if ref is None:
        ref = i.clone()
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "rows_total": n_total, "setting": sname, "ms_per_step": round(ms.item(), 3),
                          "ids_equal_first_setting": bool(flag.item())}), flush=True)
dist.barrier()
dist.destroy_process_group()
