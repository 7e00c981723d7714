#!/usr/bin/env python
"""bench.py — headline benchmark of the retrieval hot path (contract: see DESIGN.md §Measurement).

Workload (BASELINE.json configs[1], "C2"): exact squared-L2 search, k=100, 10 M x 768 bf16
database sharded by contiguous row ranges over N GPUs (10 M / N rows each: strong scaling),
10 K-query batches; every rank searches its shard with the fused tcgen05 kernel, the per-shard
top-k lists are all-gathered over NCCL/NVLink and merged on the GPU (K8).  A "step" = one
10 K-query batch.  Metric = whole-job queries per second.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm
  python bench.py --impl reference ...                          # CPU arm (FAISS-IndexFlat restatement)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DB = 10_000_000
DIM = 768
N_QUERIES = 10_000
K = 100
METRIC_NAME = "QPS at k=100 exact L2 (10M x 768 bf16, 10K-query batches, global top-k merge)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)   # >= 20 back-to-back batches, >= 2 s (clocks settle)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2vs", choices=["b2vs", "reference"])
    ap.add_argument("--n-db", type=int, default=N_DB)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--queries", type=int, default=N_QUERIES)
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=1000)
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained") or p["bf16_tflops"]),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    except Exception:
        return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(world, args):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r1_c2_bf_tc_fullpass_v2.csv); only valid for the exact configuration it was taken on."""
    if world != 1 or (args.n_db, args.dim, args.queries, args.k) != (N_DB, DIM, N_QUERIES, K):
        return None
    try:
        import csv
        with open(os.path.join(ROOT, "profiles", "r1_c2_bf_tc_fullpass_v2.csv")) as f:
            rows = list(csv.reader(f))
        h, units, v = rows[0], rows[1], rows[2]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(name)
            tot += float(v[i]) * scale.get(units[i], 1.0)
        return tot
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the samples under load (upper half: idle samples at the edges are dropped)
        loaded = sm[len(sm) // 2:] if sm else []
        med = loaded[len(loaded) // 2] if loaded else None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_SAMPLE = {}


def cpu_exact_qps(n_rows, dim, n_queries, k, threads, seconds_budget=30.0):
    """FAISS IndexFlatL2 restatement (oracle.exact.exact_knn) on the host cores: QPS on a bounded
    sample of the workload (n_rows x dim fp32, n_queries queries)."""
    from oracle.exact import exact_knn
    torch.set_num_threads(threads)
    key = (n_rows, dim, n_queries)
    if key not in _CPU_SAMPLE:            # the sample is generated once, outside every timed call
        _CPU_SAMPLE.clear()
        g = torch.Generator().manual_seed(99)
        _CPU_SAMPLE[key] = (torch.randn(n_rows, dim, generator=g), torch.randn(n_queries, dim, generator=g))
    db, q = _CPU_SAMPLE[key]
    exact_knn(db[:50_000], q[:64], k)  # warm the thread pool
    t0 = time.time()
    exact_knn(db, q, k)
    dt = time.time() - t0
    return n_queries / dt, dt


def cpu_sklearn_qps(n_rows, dim, n_queries, k):
    """The reference's own CPU path (Attempt_1/VectorSearch_QuestionRetrieval.ipynb:L878):
    scikit-learn NearestNeighbors(algorithm='brute', n_jobs=-1) on a bounded sample.  Returns
    (QPS on the sample, seconds) or None when scikit-learn is unavailable."""
    try:
        from sklearn.neighbors import NearestNeighbors
    except Exception:
        return None
    g = torch.Generator().manual_seed(98)
    db = torch.randn(n_rows, dim, generator=g).numpy()
    q = torch.randn(n_queries, dim, generator=g).numpy()
    t0 = time.time()
    nn = NearestNeighbors(n_neighbors=min(k, n_rows), algorithm="brute", n_jobs=-1).fit(db)
    nn.kneighbors(q)
    dt = time.time() - t0
    return n_queries / dt, dt


def sklearn_baseline(args, cores):
    """cpu_baseline-shaped entry for the reference's scikit-learn path (None if unavailable)."""
    rows, nq = min(200_000, args.n_db), min(200, args.queries)
    try:
        r = cpu_sklearn_qps(rows, args.dim, nq, args.k)
    except Exception:
        r = None
    if r is None:
        return None
    qps_s, dt = r
    return {"value": qps_s * rows / args.n_db, "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"scikit-learn NearestNeighbors(brute, n_jobs=-1), the reference's CPU baseline "
                      f"(VectorSearch_QuestionRetrieval.ipynb:L878): {rows} x {args.dim} fp32 rows x {nq} "
                      f"queries in {dt:.1f} s, QPS scaled by {rows}/{args.n_db} to the full database"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = min(args.cpu_sample_rows, args.n_db)
    nq = min(args.cpu_sample_queries, args.queries)
    # each "step" = one bounded sample; extrapolate linearly in rows to the full database
    # W warm-up steps on a tenth of the sample (thread pool, page faults), then K timed steps
    # (capped at 50: a step is ~3 s of all host cores)
    n_warm = max(0, min(args.warmup, 5))
    for _ in range(n_warm):
        cpu_exact_qps(min(rows, 100_000), args.dim, min(nq, 128), args.k, cores)
    times = []
    t_begin = time.time()
    for _ in range(max(1, min(args.steps, 50))):
        qps_s, dt = cpu_exact_qps(rows, args.dim, nq, args.k, cores)
        times.append(dt)
        if time.time() - t_begin > 150.0:    # slow hosts: stay within a few minutes, report the steps done
            break
    dt = sum(times) / len(times)
    qps_sample = nq / dt
    qps_full = qps_sample * rows / args.n_db
    sample = (f"{rows} x {args.dim} fp32 rows x {nq} queries per step, blocked sgemm + top-k "
              f"(oracle.exact.exact_knn), QPS scaled by {rows}/{args.n_db} to the full database")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": qps_full, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": n_warm,
        "ms_per_step": dt * 1e3 * (args.n_db / rows) * (args.queries / nq),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C2 exact L2 k=100 10Mx768, CPU host cores", "n_db": args.n_db,
                   "dim": args.dim, "queries_per_batch": args.queries, "k": args.k},
        "cpu_baseline": {"value": qps_full, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps_full, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    sk = sklearn_baseline(args, cores)
    if sk is not None:
        line["cpu_baseline_sklearn"] = sk
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    import cuvs_rag_b200 as b2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    # ---- the reference-facing objects (drop-in surface)
    grm = b2.GPUResourceManager(devices=[local_rank])
    edm = b2.EmbeddingDistributionManager(grm)
    ibc = b2.IndexBuildingCoordinator(grm)
    sra = b2.SearchResultAggregator(grm)

    # ---- synthetic shard, generated on the device (reference style: torch.randn, unseeded there)
    start, end = b2.partition_even(args.n_db, world)[rank]
    n_local = end - start
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = torch.empty((n_local, args.dim), dtype=torch.bfloat16, device=dev)
    chunk = 1 << 20
    for s in range(0, n_local, chunk):
        e = min(s + chunk, n_local)
        shard[s:e] = torch.randn((e - s, args.dim), generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    qgen = torch.Generator(device="cpu").manual_seed(4321)
    q_host = torch.randn((args.queries, args.dim), generator=qgen).to(torch.bfloat16).pin_memory()
    q_dev = q_host.to(dev)

    part = b2.EmbeddingPart(local_rank, shard, start, end)
    cfg = b2.IndexBuildConfig("brute_force", {"metric": "sqeuclidean"}, parallel_build=False, max_retries=0)
    res = ibc._build_single_index(part, cfg)
    if not res.success:
        raise SystemExit(f"index build failed: {res.error_message}")
    index = res.index
    ibc.built_indices[local_rank] = index

    out_d = torch.empty((args.queries, args.k), dtype=torch.float32, device=dev)
    out_i = torch.empty((args.queries, args.k), dtype=torch.int64, device=dev)
    if world > 1:
        g_d = torch.empty((world, args.queries, args.k), dtype=torch.float32, device=dev)
        g_i = torch.empty((world, args.queries, args.k), dtype=torch.int64, device=dev)

    phase_ev = []   # (start, after search, after all-gather, after merge) event tuples, N > 1

    def step_device(time_kernel=False, phases=False):
        """One batch with inputs resident in HBM: local search -> all-gather -> merge."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if phases and world > 1 else None
        if ev:
            ev[0].record()
        index.search(q_dev, args.k, out=(out_d, out_i), time_kernel=time_kernel)
        if world > 1:
            if ev:
                ev[1].record()
            dist.all_gather_into_tensor(g_d, out_d)
            dist.all_gather_into_tensor(g_i, out_i)
            if ev:
                ev[2].record()
            res = b2.merge_topk(g_d, g_i, args.k, descending=False)
            if ev:
                ev[3].record()
                phase_ev.append(ev)
            return res
        return out_d, out_i

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(args.warmup):
        step_device()
    sync_all()

    # ---- timed region (device timed, inputs resident; db shard >> L2 so no flush is needed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    launches = 0
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_device(time_kernel=True, phases=True)
        st = index.last_stats()   # resolves the event pair of this step's fused kernel
        kernel_ms.append(st.kernel_ms)
        launches += st.launches + (1 if world > 1 else 0)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    qps = args.queries * args.steps / (ms_total * 1e-3)
    stats = index.last_stats()

    # ---- e2e: the user-facing call with HOST buffers (H2D + search + merge + D2H every step)
    scfg = b2.SearchConfig(k=args.k, search_params={"collect_gpu_results": False},
                           parallel_search=False, validate_results=False)
    for _ in range(max(1, args.warmup)):
        sra.perform_distributed_search(q_host, {local_rank: index}, scfg)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = sra.perform_distributed_search(q_host, {local_rank: index}, scfg)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_qps = args.queries * args.steps / e2e_s
    h2d = q_host.numel() * q_host.element_size()
    d2h = args.queries * args.k * (4 + 8)

    # ---- parity spot-check against the oracle on a small slice (checker only, untimed)
    parity = None
    if rank == 0 and world == 1:
        try:
            from oracle.exact import topk_parity_report
            nchk = min(n_local, 300000)   # 1172 tiles: exercises the seeded multi-pass path
            sub = b2.NativeIndex.flat(shard[:nchk], metric="sqeuclidean")
            dd, ii = sub.search(q_dev[:64], args.k)
            rep = topk_parity_report(dd.cpu(), ii.cpu(), shard[:nchk].float().cpu(),
                                     q_dev[:64].float().cpu(), args.k)
            parity = {"ok": rep["ok"], "exact_id_match": rep["exact_id_match"]}
            sub.destroy()
        except Exception as exc:  # the bench number stands; parity is reported, not assumed
            parity = {"ok": False, "error": str(exc)[:200]}
    if world > 1:
        # sharded run: size-independent properties of the merged answer (all ranks hold it).
        #  * the best merged hit of every query is the best hit of some shard (all-reduce MIN),
        #  * rows are sorted, ids unique and inside [0, N), every rank computed the same result.
        md, mi = step_device()
        torch.cuda.synchronize()
        best_local = out_d[:, 0].clone()
        dist.all_reduce(best_local, op=dist.ReduceOp.MIN)
        chk = torch.stack([(mi.double().sum()), md.double().sum()])
        chk_max, chk_min = chk.clone(), chk.clone()
        dist.all_reduce(chk_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk_min, op=dist.ReduceOp.MIN)
        srt = bool((md[:, 1:] >= md[:, :-1]).all())
        in_range = bool(((mi >= 0) & (mi < args.n_db)).all())
        uniq = all(len(set(r)) == args.k for r in mi[:64].cpu().tolist())
        parity = {"ok": bool(torch.equal(md[:, 0], best_local)) and srt and in_range and uniq and
                  bool(torch.equal(chk_max, chk_min)),
                  "check": "merged best == min over shards, sorted, unique in-range ids, identical on all ranks"}

    if rank == 0:
        peaks = load_peaks()
        kms = sum(kernel_ms) / max(1, len(kernel_ms))
        flops_per_launch = 2.0 * args.queries * n_local * args.dim
        achieved = flops_per_launch / (kms * 1e-3) / 1e12 if kms > 0 else None
        line = {
            "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "C2: exact L2 k=100, 10Mx768 bf16 row-sharded, 10K-query batches, "
                                   "NCCL all-gather + GPU merge",
                       "n_db": args.n_db, "dim": args.dim, "queries_per_batch": args.queries,
                       "k": args.k, "rows_per_gpu": n_local, "parallelism": f"shard{world}",
                       "l2_policy": "inputs larger than L2 (db shard >= 1.9 GB vs 126 MB L2), no flush",
                       "n_splits": stats.n_splits, "grid": stats.grid},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"],
                         "unit": "TFLOP/s", "frac": (achieved / peaks["tflops"]) if achieved else None,
                         "traffic": ncu_traffic(world, args), "kernel": "bf_tc_kernel", "kernel_ms": kms,
                         "algorithmic_flops_per_launch": flops_per_launch,
                         "peak_source": peaks["source"]},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "api": "SearchResultAggregator.perform_distributed_search (pinned host queries)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": parity,
        }
        if phase_ev:
            # rank 0's device time per phase (mean over the timed steps)
            n = len(phase_ev)
            line["phases_ms"] = {
                "search": sum(e[0].elapsed_time(e[1]) for e in phase_ev) / n,
                "all_gather": sum(e[1].elapsed_time(e[2]) for e in phase_ev) / n,
                "merge": sum(e[2].elapsed_time(e[3]) for e in phase_ev) / n}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            rows = min(args.cpu_sample_rows, args.n_db)
            nq = min(args.cpu_sample_queries, args.queries)
            qps_s, dt = cpu_exact_qps(rows, args.dim, nq, args.k, cores)
            line["cpu_baseline"] = {
                "value": qps_s * rows / args.n_db, "unit": "queries/s", "cores": cores, "kind": "port",
                "sample": f"{rows} x {args.dim} fp32 rows x {nq} queries in {dt:.1f} s "
                          f"(oracle.exact.exact_knn = FAISS IndexFlatL2 restatement), QPS scaled by "
                          f"{rows}/{args.n_db} to the full database"}
            sk = sklearn_baseline(args, cores)
            if sk is not None:
                line["cpu_baseline_sklearn"] = sk
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
