"""SearchResultAggregator — per-shard search fan-out and the global top-k merge.

The reference ships this module EMPTY (``Attempt_1/search_result_aggregator.py`` is 0 bytes); its
contract is fixed by ``Attempt_1/test_search_result_aggregator.py`` and
``Latest/cuVS-2-gpu/old/DesignDocument.md:119-137``, and its behaviour by
the second-generation ``ParallelSearchEngine`` (``improved_multi_gpu_rag.py:209-277``) and the
notebook hot loop (``cuvs-2gpu-main.ipynb`` cell 16): search every shard, make ids global, keep
the best k.  Here each shard search is one ``b2vs_search`` call (ids come back already global:
row + ``EmbeddingPart.start_index``), and the merge is ``b2vs_merge_topk`` on the GPU after an
NVLink gather (peer copies inside one process, NCCL all-gather across torchrun ranks) instead of
the reference's per-query host ``np.argsort``.
"""
from __future__ import annotations

import logging
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

try:
    from gpu_resource_manager import GPUResourceManager
    import _native
except ImportError:  # pragma: no cover - package-style import
    from .gpu_resource_manager import GPUResourceManager
    from . import _native

logger = logging.getLogger(__name__)

# Same module-flag name as the reference's other modules; True when a CUDA device is present
# (the native library is then required).  Unit tests patch it to False to get simulated searches.
CUVS_AVAILABLE = bool(torch.cuda.is_available())


@dataclass
class SearchResult:
    """One shard's answer: distances [Q, k'] float32, indices [Q, k'] int64 (global ids)."""
    distances: np.ndarray
    indices: np.ndarray
    gpu_id: int
    query_time: float
    k_requested: int
    k_returned: int

    def __post_init__(self):
        if self.gpu_id < 0:
            raise ValueError(f"gpu_id must be non-negative, got {self.gpu_id}")
        if self.query_time < 0:
            raise ValueError(f"query_time must be non-negative, got {self.query_time}")
        if self.k_requested <= 0:
            raise ValueError(f"k_requested must be positive, got {self.k_requested}")
        if self.k_returned > self.k_requested:
            raise ValueError(f"k_returned ({self.k_returned}) cannot exceed k_requested "
                             f"({self.k_requested})")
        if np.ndim(self.distances) != 2:
            raise ValueError(f"distances must be 2D array, got {np.ndim(self.distances)}D")
        if np.shape(self.distances) != np.shape(self.indices):
            raise ValueError(f"distances shape {np.shape(self.distances)} != indices shape "
                             f"{np.shape(self.indices)}")


@dataclass
class AggregatedSearchResult:
    final_distances: np.ndarray
    final_indices: np.ndarray
    total_query_time: float
    gpu_results: List[SearchResult]
    k_requested: int
    k_returned: int
    num_queries: int
    # degraded searches (SearchConfig.search_params["allow_partial"]): the GPUs whose shard did not
    # answer and why - the rows of those shards are absent from final_* (the reference drops
    # failed shards with a log line only, improved_multi_gpu_rag.py:261-263)
    missing_gpus: List[int] = field(default_factory=list)
    shard_errors: Dict[int, str] = field(default_factory=dict)

    @property
    def complete(self) -> bool:
        return not self.missing_gpus

    def __post_init__(self):
        if self.k_requested <= 0:
            raise ValueError(f"k_requested must be positive, got {self.k_requested}")
        if self.num_queries <= 0:
            raise ValueError(f"num_queries must be positive, got {self.num_queries}")
        if self.total_query_time < 0:
            raise ValueError(f"total_query_time must be non-negative, got {self.total_query_time}")


@dataclass
class SearchConfig:
    """``search_params`` keys understood by the native path: ``n_probes`` (alias ``nprobe``),
    ``refine_ratio``, ``k_local`` (per-shard k, default k), ``collect_gpu_results`` (default
    True: per-shard results are copied to the host into ``gpu_results`` as the reference does),
    ``graph`` (replay small IVF batches as a CUDA graph), ``result_layout`` (``"sliced"``: one
    process per GPU, each rank passes ITS slice of the batch and gets that slice's global answer -
    see ``_search_sliced``; default: every rank passes and receives the whole batch),
    ``allow_partial`` (default False: a shard
    that fails fails the search; True: answer from the shards that did respond and list the others
    in ``AggregatedSearchResult.missing_gpus`` / ``shard_errors``)."""
    k: int
    search_params: Optional[Dict[str, Any]] = None
    parallel_search: bool = True
    timeout_seconds: Optional[float] = None
    validate_results: bool = True

    def __post_init__(self):
        if self.k <= 0:
            raise ValueError(f"k must be positive, got {self.k}")
        if self.timeout_seconds is not None and self.timeout_seconds <= 0:
            raise ValueError(f"timeout_seconds must be positive, got {self.timeout_seconds}")


def _host_merge(dist_list: List[np.ndarray], idx_list: List[np.ndarray], k: int,
                descending: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate + stable sort, used ONLY when no CUDA device exists (host-side unit tests)."""
    d = np.concatenate(dist_list, axis=1)
    i = np.concatenate(idx_list, axis=1)
    key = -d if descending else d
    order = np.argsort(key, axis=1, kind="stable")[:, :k]
    return (np.take_along_axis(d, order, 1).astype(np.float32),
            np.take_along_axis(i, order, 1).astype(np.int64))


def _merge_arrays(dist_list: List[np.ndarray], idx_list: List[np.ndarray], k: int,
                  descending: bool = False, device: Optional[torch.device] = None
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Global top-k of per-shard host arrays: runs b2vs_merge_topk on a GPU when one exists.
    The host merge below is reached ONLY on a box without CUDA (host-side unit tests): on a GPU
    box a k beyond the kernel's limit raises, like every other limit of the native path."""
    k_total = sum(d.shape[1] for d in dist_list)
    k = min(k, k_total)
    if torch.cuda.is_available() and k > _native.MAX_K:
        raise RuntimeError(f"b2vs_merge_topk failed (code -4): k_out={k} outside [1, {_native.MAX_K}]")
    if torch.cuda.is_available():
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        k_in = max(d.shape[1] for d in dist_list)
        nq = dist_list[0].shape[0]
        d_all = torch.full((len(dist_list), nq, k_in), float("-inf") if descending else float("inf"),
                           dtype=torch.float32)
        i_all = torch.full((len(dist_list), nq, k_in), -1, dtype=torch.int64)
        for g, (d, i) in enumerate(zip(dist_list, idx_list)):
            d_all[g, :, : d.shape[1]] = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
            i_all[g, :, : i.shape[1]] = torch.from_numpy(np.ascontiguousarray(i, dtype=np.int64))
        out_d, out_i = _native.merge_topk(d_all.to(dev), i_all.to(dev), k, descending)
        return out_d.cpu().numpy(), out_i.cpu().numpy()
    return _host_merge(dist_list, idx_list, k, descending)


def allgather_and_merge(d_all: torch.Tensor, i_all: torch.Tensor, k: int, descending: bool = False,
                        world_size: int = 1, comm: Any = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exchange step of the sharded search: ``d_all`` / ``i_all`` are this process's
    [n_local_shards, Q, k'] results (ids already global).  With ``world_size > 1`` (one process per
    GPU under torchrun) every rank all-gathers the other ranks' lists — NCCL over NVLink for CUDA
    tensors — and then merges; the merge is ``b2vs_merge_topk`` (K8) on the GPU.  CPU tensors take
    the gloo + host-merge route, which exists only so the N>1 host logic is testable without GPUs.
    """
    nq = d_all.shape[1]
    if world_size > 1 and comm is not None and d_all.is_cuda:
        # the exchange lives behind the C ABI: one grouped NCCL all-gather (+ K8b)
        if d_all.shape[0] > 1:   # several local shards: fold them first, then exchange one list
            k_loc = min(d_all.shape[2] * d_all.shape[0], max(k, d_all.shape[2]))
            d_loc, i_loc = _native.merge_topk(d_all, i_all, min(k_loc, _native.MAX_K), descending)
        else:
            d_loc, i_loc = d_all[0].contiguous(), i_all[0].contiguous()
        k_eff = min(k, world_size * d_loc.shape[1])
        if k_eff <= _native.MAX_FUSED_K:
            return comm.allgather_merge_topk(d_loc, i_loc, k_eff, descending)
        g_d, g_i = comm.allgather_topk(d_loc, i_loc)
        return _native.merge_topk(g_d, g_i, k_eff, descending)
    if world_size > 1:
        import torch.distributed as dist
        g_d = torch.empty((world_size,) + tuple(d_all.shape), dtype=d_all.dtype, device=d_all.device)
        g_i = torch.empty((world_size,) + tuple(i_all.shape), dtype=i_all.dtype, device=i_all.device)
        if d_all.is_cuda:
            dist.all_gather_into_tensor(g_d, d_all.contiguous())
            dist.all_gather_into_tensor(g_i, i_all.contiguous())
        else:
            dist.all_gather(list(g_d.unbind(0)), d_all.contiguous())
            dist.all_gather(list(g_i.unbind(0)), i_all.contiguous())
        d_all = g_d.reshape(-1, nq, d_all.shape[-1])
        i_all = g_i.reshape(-1, nq, i_all.shape[-1])
    if d_all.shape[0] == 1 and d_all.shape[2] == k:
        return d_all[0], i_all[0]                    # single shard: already the global top-k
    k_eff = min(k, d_all.shape[0] * d_all.shape[2])
    if d_all.is_cuda:
        return _native.merge_topk(d_all, i_all, k_eff, descending)
    md, mi = _host_merge([x.numpy() for x in d_all.unbind(0)], [x.numpy() for x in i_all.unbind(0)],
                         k_eff, descending)
    return torch.from_numpy(md), torch.from_numpy(mi)


class SearchResultAggregator:
    def __init__(self, gpu_manager: GPUResourceManager):
        if gpu_manager is None or not hasattr(gpu_manager, "validate_gpu_index"):
            raise TypeError("gpu_manager must be a GPUResourceManager instance")
        self.gpu_manager = gpu_manager
        self.search_history: List[AggregatedSearchResult] = []
        self._active_searches: Dict[int, bool] = {}

    # ------------------------------------------------------------------ validation / merge
    def validate_search_results(self, gpu_results: List[SearchResult], expected_queries: int,
                                expected_k: int) -> bool:
        if not gpu_results:
            raise ValueError("gpu_results cannot be empty")
        for r in gpu_results:
            if r.distances.shape[0] != expected_queries:
                raise ValueError(f"GPU {r.gpu_id} result has {r.distances.shape[0]} queries, "
                                 f"expected {expected_queries}")
            if r.distances.shape[1] > max(expected_k, r.k_requested):
                raise ValueError(f"GPU {r.gpu_id} result has {r.distances.shape[1]} neighbours, "
                                 f"expected at most {expected_k}")
            if np.isnan(r.distances).any():
                raise ValueError(f"GPU {r.gpu_id} result contains NaN distances")
        return True

    def merge_search_results(self, results: List[SearchResult], k: int, descending: bool = False
                             ) -> Tuple[np.ndarray, np.ndarray]:
        """Global top-k over per-shard results (ids are already global and pass through)."""
        if not results:
            raise ValueError("Cannot merge empty results list")
        nq = results[0].distances.shape[0]
        for r in results:
            if r.distances.shape[0] != nq:
                raise ValueError(f"GPU {r.gpu_id} result has {r.distances.shape[0]} queries, "
                                 f"expected {nq}")
        return _merge_arrays([r.distances for r in results], [r.indices for r in results], k,
                             descending)

    # ------------------------------------------------------------------ simulated (unit tests)
    def _simulate_search(self, query: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Random, well-formed results for host-side tests (no index, no GPU involved)."""
        nq = query.shape[0]
        distances, _ = torch.sort(torch.rand(nq, k), dim=1)
        indices = torch.randint(0, 1_000_000, (nq, k), dtype=torch.int64)
        return distances, indices

    # ------------------------------------------------------------------ one shard
    def _search_single_gpu(self, gpu_id: int, index: Any, query: torch.Tensor, k_local: int,
                           params: Dict[str, Any]):
        """Returns (dist, ids, seconds, on_device). Native indices answer on their GPU."""
        t0 = time.time()
        self._active_searches[gpu_id] = True
        try:
            if CUVS_AVAILABLE and isinstance(index, _native.NativeIndex):
                dev = index.device
                q = query if query.device == dev else query.to(dev, non_blocking=True)
                if not q.is_contiguous():
                    q = q.contiguous()
                d, i = index.search(q, k_local,
                                    n_probes=int(params.get("n_probes", params.get("nprobe", 0)) or 0),
                                    refine_ratio=int(params.get("refine_ratio", 0) or 0),
                                    graph=bool(params.get("graph", False)))
                return d, i, time.time() - t0, True
            if CUVS_AVAILABLE and torch.cuda.is_available():
                raise TypeError(f"index for GPU {gpu_id} is {type(index).__name__}, not a native "
                                "index; simulated indices cannot be searched on a GPU box")
            self.gpu_manager.get_safe_device_string(gpu_id)  # same validation path as the real one
            d, i = self._simulate_search(query, k_local)
            return d, i, time.time() - t0, False
        finally:
            self._active_searches[gpu_id] = False

    # ------------------------------------------------------------------ the public call
    def perform_distributed_search(self, query: torch.Tensor, indices: Dict[int, Any],
                                   config: SearchConfig) -> AggregatedSearchResult:
        if not isinstance(query, torch.Tensor):
            raise ValueError("query must be a torch.Tensor")
        if query.dim() != 2:
            raise ValueError(f"query must be 2D tensor, got {query.dim()}D")
        if query.shape[0] == 0:
            raise ValueError("query cannot be empty")
        if not indices:
            raise ValueError("indices dictionary cannot be empty")
        for g in indices:
            if not self.gpu_manager.validate_gpu_index(g):
                raise ValueError(f"GPU {g} in indices is not available")
        t0 = time.time()
        params = dict(config.search_params or {})
        if params.get("result_layout") == "sliced":
            return self._search_sliced(query, indices, config, params, t0)
        k = config.k
        k_local = int(params.get("k_local", k))
        collect = bool(params.get("collect_gpu_results", True))
        nq = query.shape[0]
        gpus = sorted(indices)

        allow_partial = bool(params.get("allow_partial", False))
        shard_errors: Dict[int, str] = {}

        def shard_failed(g: int, exc: BaseException) -> None:
            if not allow_partial:
                raise exc          # unchanged type and message: the default is to fail loudly
            shard_errors[g] = f"{type(exc).__name__}: {exc}"
            logger.error("search on GPU %d failed, its shard is missing from the result: %s", g, exc)

        answered: List[int] = []
        raw = []
        if config.parallel_search and len(gpus) > 1:
            with ThreadPoolExecutor(max_workers=len(gpus)) as pool:
                futs = [pool.submit(self._search_single_gpu, g, indices[g], query, k_local, params)
                        for g in gpus]
                for g, f in zip(gpus, futs):
                    try:
                        raw.append(f.result(timeout=config.timeout_seconds))
                        answered.append(g)
                    except Exception as exc:  # noqa: BLE001 - reported per shard
                        shard_failed(g, exc)
        else:
            for g in gpus:
                try:
                    raw.append(self._search_single_gpu(g, indices[g], query, k_local, params))
                    answered.append(g)
                except Exception as exc:  # noqa: BLE001
                    shard_failed(g, exc)
        if not raw:
            raise RuntimeError(f"search failed on every shard: {shard_errors}")
        missing = [g for g in gpus if g not in answered]
        gpus = answered

        on_device = all(r[3] for r in raw)
        descending = any(getattr(indices[g], "descending", False) is True for g in gpus)
        if on_device:
            final_d, final_i, gpu_results = self._merge_on_device(gpus, raw, k, k_local, nq,
                                                                  descending, collect)
        else:
            gpu_results = [
                SearchResult(d.cpu().numpy().astype(np.float32), i.cpu().numpy().astype(np.int64),
                             g, secs, k_local, int(d.shape[1]))
                for g, (d, i, secs, _) in zip(gpus, raw)]
            final_d, final_i = self.merge_search_results(gpu_results, k, descending)
        if config.validate_results:
            if np.isnan(final_d).any():
                raise ValueError("merged result contains NaN distances")
            if final_d.shape != final_i.shape or final_d.shape[0] != nq:
                raise ValueError(f"merged result has shape {final_d.shape}, expected ({nq}, {k})")
        result = AggregatedSearchResult(final_d, final_i, time.time() - t0, gpu_results, k,
                                        int(final_d.shape[1]), nq, missing, shard_errors)
        self.search_history.append(result)
        return result

    def _merge_on_device(self, gpus, raw, k, k_local, nq, descending, collect):
        """Gather every shard's [Q, k'] onto one GPU (NVLink peer copies), add the other ranks'
        results when running one process per GPU (NCCL all-gather), merge with K8, read back."""
        primary = raw[0][0].device
        d_parts = [r[0] if r[0].device == primary else r[0].to(primary, non_blocking=True) for r in raw]
        i_parts = [r[1] if r[1].device == primary else r[1].to(primary, non_blocking=True) for r in raw]
        for r in raw:  # make the primary stream wait for the producers on other devices
            if r[0].device != primary:
                torch.cuda.current_stream(primary).wait_stream(torch.cuda.current_stream(r[0].device))
        d_all = torch.stack(d_parts) if len(d_parts) > 1 else d_parts[0].unsqueeze(0)
        i_all = torch.stack(i_parts) if len(i_parts) > 1 else i_parts[0].unsqueeze(0)
        out_d, out_i = allgather_and_merge(d_all, i_all, k, descending, self._world_size(),
                                           self._exchange_comm(primary))
        final_d = out_d.cpu().numpy()
        final_i = out_i.cpu().numpy()
        gpu_results = []
        for g, (d, i, secs, _) in zip(gpus, raw):
            if collect:
                gpu_results.append(SearchResult(d.cpu().numpy(), i.cpu().numpy(), g, secs, k_local,
                                                int(d.shape[1])))
            else:
                gpu_results.append(SearchResult(np.empty((nq, 0), np.float32),
                                                np.empty((nq, 0), np.int64), g, secs, k_local, 0))
        return final_d, final_i, gpu_results

    def _exchange_comm(self, device):
        try:
            return self.gpu_manager.get_exchange_comm(device.index)
        except Exception as exc:  # noqa: BLE001 - fall back to torch.distributed's collectives
            logger.warning("library-level exchange communicator unavailable (%s)", exc)
            return None

    def _search_sliced(self, query: torch.Tensor, indices: Dict[int, Any], config: SearchConfig,
                       params: Dict[str, Any], t0: float) -> AggregatedSearchResult:
        """``search_params["result_layout"] == "sliced"`` under torchrun (one process per GPU):
        ``query`` is THIS rank's slice of the batch (rows ``partition_even(Q, world)[rank]``, host
        or device).  One ``b2vs_search_sharded`` call does the rest behind the C ABI: the slices
        are all-gathered over NVLink, every rank searches its shard for all Q queries, the lists
        are exchanged all-to-all and rank r merges (and downloads) only its own Q/G answers.
        ``search_params["num_queries_total"]`` gives Q (default: slice rows x world, equal slices)."""
        if len(indices) != 1:
            raise ValueError("sliced results need exactly one shard (one process per GPU)")
        (gpu, index), = indices.items()
        if not (CUVS_AVAILABLE and isinstance(index, _native.NativeIndex)):
            raise TypeError("sliced results need a native index on a CUDA device")
        comm = self.gpu_manager.get_exchange_comm(gpu)
        if comm is None:
            raise RuntimeError("sliced results need a distributed job (torchrun, WORLD_SIZE > 1)")
        nq_total = int(params.get("num_queries_total", query.shape[0] * comm.n_ranks))
        k = config.k
        self._active_searches[gpu] = True
        try:
            q = query if query.is_contiguous() else query.contiguous()
            d, i = comm.search_sharded(index, q, nq_total, k,
                                       n_probes=int(params.get("n_probes", params.get("nprobe", 0)) or 0),
                                       refine_ratio=int(params.get("refine_ratio", 0) or 0))
        finally:
            self._active_searches[gpu] = False
        final_d = d.cpu().numpy() if d.is_cuda else d.numpy()
        final_i = i.cpu().numpy() if i.is_cuda else i.numpy()
        nq = query.shape[0]
        secs = time.time() - t0
        gpu_results = [SearchResult(np.empty((nq, 0), np.float32), np.empty((nq, 0), np.int64), gpu, secs,
                                    k, 0)]
        if config.validate_results and np.isnan(final_d).any():
            raise ValueError("merged result contains NaN distances")
        result = AggregatedSearchResult(final_d, final_i, secs, gpu_results, k, int(final_d.shape[1]), nq)
        self.search_history.append(result)
        return result

    def _world_size(self) -> int:
        try:
            _, world = self.gpu_manager.get_rank_info()
            return int(world) if isinstance(world, int) else 1
        except Exception:
            return 1

    # ------------------------------------------------------------------ bookkeeping
    def get_search_history(self) -> List[AggregatedSearchResult]:
        return list(self.search_history)

    def clear_search_history(self) -> None:
        self.search_history.clear()

    def get_active_searches(self) -> Dict[int, bool]:
        return dict(self._active_searches)

    def __str__(self) -> str:
        return f"SearchResultAggregator(history_size={len(self.search_history)})"

    def __repr__(self) -> str:
        return (f"SearchResultAggregator(gpu_manager={self.gpu_manager!r}, "
                f"history_size={len(self.search_history)}, "
                f"active_searches={sum(1 for a in self._active_searches.values() if a)})")


def combine_search_results(results: List[SearchResult], k: int, descending: bool = False
                           ) -> Tuple[np.ndarray, np.ndarray]:
    """Functional form of ``SearchResultAggregator.merge_search_results``."""
    if not results:
        raise ValueError("Cannot merge empty results list")
    nq = results[0].distances.shape[0]
    for r in results:
        if r.distances.shape[0] != nq:
            raise ValueError(f"GPU {r.gpu_id} result has {r.distances.shape[0]} queries, expected {nq}")
    return _merge_arrays([r.distances for r in results], [r.indices for r in results], k, descending)


def filter_search_results_by_distance(result: SearchResult, max_distance: float) -> SearchResult:
    """Mask every neighbour farther than ``max_distance``: distance -> inf, index -> -1."""
    keep = result.distances <= max_distance
    d = np.where(keep, result.distances, np.float32(np.inf)).astype(np.float32)
    i = np.where(keep, result.indices, -1).astype(np.int64)
    return SearchResult(d, i, result.gpu_id, result.query_time, result.k_requested,
                        int(keep.sum(axis=1).max()) if keep.size else 0)
