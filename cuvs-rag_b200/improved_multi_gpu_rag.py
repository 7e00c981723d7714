"""Drop-in for the reference's second-generation module ``Latest/cuVS-2-gpu/improved_multi_gpu_rag.py``.

Same public names and call shapes — ``IndexType``, ``SearchConfig`` (``top_k = 2000`` default,
``recall_k_values``), ``GPUConfig``, ``CUDAMemoryManager``, ``ParallelIndexBuilder``
(``build_index_on_gpu`` / ``build_indices_parallel``), ``ParallelSearchEngine``
(``search_on_gpu`` / ``parallel_search`` / ``batch_search``), ``RecallEvaluator``,
``get_memory_stats`` / ``print_memory_status`` — with the cuVS calls behind them
(``improved_multi_gpu_rag.py:126-143`` build, ``:225-233`` search) replaced by ``libb2vs.so``
and the host-side ``np.argsort`` merge (``:266-275``) by the GPU merge kernel.

What changes for a caller, all of it deliberate:
  * indexes are :class:`_native.NativeIndex` objects that already return GLOBAL row ids (each part
    gets ``id_offset`` = rows of the parts before it); the reference's ``parallel_search``
    returns shard-local ids and the notebooks patch them by hand;
  * every shard is asked for ``k`` neighbours, not ``2k`` (``:247``): an exact per-shard top-k
    already contains the shard's part of the global top-k;
  * ``batch_search`` runs the whole list of queries as ONE batch per GPU instead of one thread
    per query (``:279-303``); the return value is still a list of per-query tuples;
  * ``top_k = 2000`` (the default) is served by the large-k paths of the flat, IVF-Flat and IVF-PQ
    indexes (IVF-PQ: sub-vector length 2/4/8 and dim % 64 == 0, else k <= 128);
  * ``FAISS_FLAT`` / ``FAISS_IVF`` map to the exact and IVF-Flat indexes; ``CAGRA`` is out of
    scope and raises; there is no CPU or simulated path here — without CUDA every build raises.
"""
from __future__ import annotations

import gc
import logging
import time
from concurrent.futures import ThreadPoolExecutor
from contextlib import contextmanager
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

import _native
from evaluation import RecallEvaluator  # noqa: F401  (one class, re-exported under the reference's module name)

logger = logging.getLogger(__name__)


class IndexType(Enum):
    """Supported index types (same values as the reference's enum)."""
    IVF_FLAT = "ivf_flat"
    IVF_PQ = "ivf_pq"
    CAGRA = "cagra"
    FAISS_FLAT = "faiss_flat"
    FAISS_IVF = "faiss_ivf"


@dataclass
class SearchConfig:
    """Configuration for search operations (improved_multi_gpu_rag.py:37-48)."""
    top_k: int = 2000
    search_batch_size: int = 100
    num_queries: int = 100
    enable_recall_eval: bool = True
    recall_k_values: Optional[List[int]] = None

    def __post_init__(self):
        if self.recall_k_values is None:
            self.recall_k_values = [1, 5, 10, 50, 100, 500, 1000, 2000]


@dataclass
class GPUConfig:
    """Per-GPU bookkeeping (improved_multi_gpu_rag.py:50-72); the memory limit defaults to a B200."""
    device_id: int
    memory_limit_gb: float = 180.0
    reserved_memory_gb: float = 2.0

    @property
    def device_str(self) -> str:
        return f"cuda:{self.device_id}"

    def get_available_memory(self) -> float:
        if torch.cuda.is_available():
            return torch.cuda.mem_get_info(self.device_id)[0] / 1024 ** 3
        return 0.0

    def can_allocate(self, size_gb: float) -> bool:
        return self.get_available_memory() > (size_gb + self.reserved_memory_gb)


class CUDAMemoryManager:
    """Logs the memory used by an operation and frees the cache after an OOM before re-raising."""

    @staticmethod
    @contextmanager
    def managed_allocation(gpu_config: GPUConfig, operation: str):
        initial = gpu_config.get_available_memory()
        logger.info("[GPU %d] Starting %s with %.2f GB available", gpu_config.device_id, operation, initial)
        try:
            yield
        except torch.cuda.OutOfMemoryError as exc:
            logger.error("[GPU %d] OOM during %s: %s", gpu_config.device_id, operation, exc)
            torch.cuda.empty_cache()
            gc.collect()
            raise
        except Exception as exc:
            logger.error("[GPU %d] Error during %s: %s", gpu_config.device_id, operation, exc)
            raise
        finally:
            used = initial - gpu_config.get_available_memory()
            logger.info("[GPU %d] Completed %s, used %.2f GB", gpu_config.device_id, operation, used)


def _metric_of(params: Dict) -> str:
    return str(params.get("metric", "sqeuclidean"))


class ParallelIndexBuilder:
    """One index per GPU, built from one thread per GPU (ctypes releases the GIL)."""

    def __init__(self, num_gpus: Optional[int] = None):
        self.num_gpus = num_gpus or torch.cuda.device_count()
        self.gpu_configs = [GPUConfig(i) for i in range(self.num_gpus)]
        self.executor = ThreadPoolExecutor(max_workers=max(1, self.num_gpus))
        logger.info("Initialized ParallelIndexBuilder with %d GPUs", self.num_gpus)

    def build_index_on_gpu(self, gpu_config: GPUConfig, embeddings: torch.Tensor,
                           index_type: IndexType, params: Dict) -> Tuple[Any, float]:
        """Build one shard's index on ``gpu_config.device_id``; ``params['id_offset']`` (set by
        ``build_indices_parallel``) is the global row number of the shard's first row."""
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is required: there is no CPU index in this module")
        start = time.time()
        with CUDAMemoryManager.managed_allocation(gpu_config, f"building {index_type.value} index"):
            if not embeddings.is_cuda or embeddings.device.index != gpu_config.device_id:
                embeddings = embeddings.to(gpu_config.device_str)
            embeddings = embeddings.contiguous()
            n = int(embeddings.shape[0])
            metric = _metric_of(params)
            id_offset = int(params.get("id_offset", 0))
            if index_type in (IndexType.IVF_FLAT, IndexType.FAISS_IVF):
                # same default list counts as the reference (:128, :133)
                n_lists = int(params.get("n_lists", min(256, n // 1000 + 1)))
                index = _native.NativeIndex.ivf_flat(embeddings, n_lists, metric=metric,
                                                     id_offset=id_offset,
                                                     kmeans_iters=int(params.get("kmeans_n_iters", 20)))
            elif index_type == IndexType.IVF_PQ:
                if int(params.get("pq_bits", 8)) != 8:
                    raise ValueError("only pq_bits=8 is supported")
                n_lists = int(params.get("n_lists", min(512, n // 500 + 1)))
                index = _native.NativeIndex.ivf_pq(embeddings, n_lists, int(params.get("pq_dim", 96)),
                                                   metric=metric, id_offset=id_offset,
                                                   kmeans_iters=int(params.get("kmeans_n_iters", 20)))
            elif index_type == IndexType.FAISS_FLAT:
                index = _native.NativeIndex.flat(embeddings, metric=metric, id_offset=id_offset)
            elif index_type == IndexType.CAGRA:
                raise ValueError("Unsupported index type: cagra (graph indexes are out of scope)")
            else:
                raise ValueError(f"Unsupported index type: {index_type}")
            torch.cuda.synchronize(gpu_config.device_id)
        build_time = time.time() - start
        logger.info("[GPU %d] Built %s index in %.2fs", gpu_config.device_id, index_type.value, build_time)
        return index, build_time

    def build_indices_parallel(self, embedding_parts: List[torch.Tensor], index_type: IndexType,
                               params: Optional[Dict] = None) -> Dict:
        params = dict(params or {})
        futures = []
        start_row = 0
        for i, embeddings in enumerate(embedding_parts[:self.num_gpus]):
            part_params = dict(params, id_offset=start_row)
            start_row += int(embeddings.shape[0])
            futures.append((i, self.executor.submit(self.build_index_on_gpu, self.gpu_configs[i],
                                                    embeddings, index_type, part_params)))
        gpu_indexes, build_times, failed_gpus = {}, {}, []
        for gpu_id, future in futures:
            try:
                index, build_time = future.result(timeout=300)
                gpu_indexes[gpu_id] = index
                build_times[gpu_id] = build_time
            except Exception as exc:  # same policy as the reference: record the GPU, keep going
                logger.error("Failed to build index on GPU %d: %s", gpu_id, exc)
                failed_gpus.append(gpu_id)
        return {
            "indexes": gpu_indexes,
            "build_times": build_times,
            "total_time": sum(build_times.values()),
            "avg_time": float(np.mean(list(build_times.values()))) if build_times else 0,
            "failed_gpus": failed_gpus,
            "success": len(failed_gpus) == 0,
        }

    def __del__(self):
        if hasattr(self, "executor"):
            self.executor.shutdown(wait=False)


class ParallelSearchEngine:
    """Sharded search: every GPU answers for its rows, one GPU merges (b2vs_merge_topk)."""

    def __init__(self, gpu_indexes: Dict[int, Any], index_type: IndexType, search_config: SearchConfig):
        self.last_missing_shards: List[Tuple[int, str]] = []   # (gpu, error) of the last search
        self.gpu_indexes = gpu_indexes
        self.index_type = index_type
        self.search_config = search_config
        self.num_gpus = len(gpu_indexes)
        self.executor = ThreadPoolExecutor(max_workers=max(1, self.num_gpus))
        logger.info("Initialized ParallelSearchEngine with %d indexes", self.num_gpus)

    # ---- device-side pieces
    def _search_device(self, gpu_id: int, index: Any, query: torch.Tensor, k: int,
                       params: Optional[Dict] = None):
        params = params or {}
        dev = torch.device(f"cuda:{gpu_id}")
        if query.dim() == 1:
            query = query.unsqueeze(0)
        if query.device != dev:
            query = query.to(dev, non_blocking=True)
        return index.search(query.contiguous(), k, n_probes=int(params.get("n_probes", 0) or 0),
                            refine_ratio=int(params.get("refine_ratio", 0) or 0))

    def _search_merged(self, queries: torch.Tensor, k: int, params: Optional[Dict] = None):
        """[Q, D] -> merged (dist [Q, k'], ids [Q, k']) on the first index's GPU."""
        gpus = sorted(self.gpu_indexes)
        # A shard that fails is skipped, as in the reference (:261-263 logs and carries on), but
        # not silently: ``last_missing_shards`` lists (gpu, error) of the call, and a search with
        # no answering shard raises.
        self.last_missing_shards = []
        raw = []
        if len(gpus) > 1:
            futs = [self.executor.submit(self._search_device, g, self.gpu_indexes[g], queries, k, params)
                    for g in gpus]
            for g, f in zip(gpus, futs):
                try:
                    raw.append(f.result(timeout=60))
                except Exception as exc:  # noqa: BLE001 - reported per shard
                    self.last_missing_shards.append((g, f"{type(exc).__name__}: {exc}"))
                    logger.error("search on GPU %d failed, its shard is missing from the result: %s", g, exc)
            if not raw:
                raise RuntimeError(f"search failed on every shard: {self.last_missing_shards}")
        else:
            raw = [self._search_device(gpus[0], self.gpu_indexes[gpus[0]], queries, k, params)]
        primary = raw[0][0].device
        for d, _ in raw:
            if d.device != primary:
                torch.cuda.current_stream(primary).wait_stream(torch.cuda.current_stream(d.device))
        d_all = torch.stack([d.to(primary, non_blocking=True) for d, _ in raw])
        i_all = torch.stack([i.to(primary, non_blocking=True) for _, i in raw])
        if d_all.shape[0] == 1:
            return d_all[0], i_all[0]
        descending = bool(getattr(self.gpu_indexes[gpus[0]], "descending", False))
        return _native.merge_topk(d_all, i_all, min(k, d_all.shape[0] * d_all.shape[2]), descending)

    # ---- the reference's methods
    def search_on_gpu(self, gpu_id: int, index: Any, query: torch.Tensor, k: int
                      ) -> Tuple[np.ndarray, np.ndarray]:
        """One shard's answer as host arrays, like cuVS under the reference's
        ``pylibraft.config.set_output_as(copy_to_host)`` hook (:114)."""
        d, i = self._search_device(gpu_id, index, query, k)
        return d.cpu().numpy(), i.cpu().numpy()

    def parallel_search(self, query: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
        """Global top-k of one query (1-D result arrays, as the reference returns) or of a
        [Q, D] batch ([Q, k] arrays)."""
        single = query.dim() == 1
        d, i = self._search_merged(query.unsqueeze(0) if single else query, self.search_config.top_k)
        d, i = d.cpu().numpy(), i.cpu().numpy()
        return (d[0], i[0]) if single else (d, i)

    def batch_search(self, queries: List[torch.Tensor]) -> List[Tuple[np.ndarray, np.ndarray]]:
        """All queries as one batch per GPU; returns one (distances, ids) tuple per query, in order."""
        if len(queries) == 0:
            return []
        stacked = torch.stack([q.reshape(-1) for q in queries])
        d, i = self._search_merged(stacked, self.search_config.top_k)
        d, i = d.cpu().numpy(), i.cpu().numpy()
        return [(d[j], i[j]) for j in range(len(queries))]

    def __del__(self):
        if hasattr(self, "executor"):
            self.executor.shutdown(wait=False)


def get_memory_stats() -> Dict:
    """Host RSS / CPU load and per-GPU memory, same keys as the reference (:359-386)."""
    stats: Dict[str, Any] = {}
    try:
        import psutil
        stats["ram_gb"] = psutil.Process().memory_info().rss / 1024 ** 3
        stats["cpu_percent"] = psutil.cpu_percent()
    except Exception:  # psutil is optional here
        stats["ram_gb"], stats["cpu_percent"] = 0.0, 0.0
    if torch.cuda.is_available():
        gpu_stats = []
        for i in range(torch.cuda.device_count()):
            free, total = torch.cuda.mem_get_info(i)
            allocated = torch.cuda.memory_allocated(i) / 1024 ** 3
            gpu_stats.append({
                "gpu_id": i, "allocated_gb": allocated,
                "reserved_gb": torch.cuda.memory_reserved(i) / 1024 ** 3,
                "free_gb": free / 1024 ** 3, "total_gb": total / 1024 ** 3,
                "used_percent": allocated / (total / 1024 ** 3) * 100,
            })
        stats["gpu_stats"] = gpu_stats
    return stats


def print_memory_status(label: str = "") -> None:
    stats = get_memory_stats()
    logger.info("%s - RAM: %.2f GB, CPU: %.1f%%", label, stats["ram_gb"], stats["cpu_percent"])
    for gpu in stats.get("gpu_stats", []):
        logger.info("  GPU %d: %.2f/%.2f GB (%.1f%% used)", gpu["gpu_id"], gpu["allocated_gb"],
                    gpu["total_gb"], gpu["used_percent"])
