"""IndexBuildingCoordinator — builds one index per GPU shard, in parallel, with retries.

Drop-in for the reference's ``Attempt_1/index_building_coordinator.py`` (same dataclasses,
methods, messages; ``test_index_building_coordinator.py`` is the acceptance spec).  The L2 call
site the reference makes into cuVS (``_create_index``, reference :370-420) goes to libb2vs.so:

    'ivf_flat'  -> b2vs_ivfflat_build   (GPU k-means + list fill; was cuvs ivf_flat.build)
    'ivf_pq'    -> b2vs_ivfpq_build     (was cuvs ivf_pq.build)
    'brute_force' / 'flat' -> b2vs_bf_create (exact search; FAISS IndexFlat* equivalent)
    'cagra'     -> accepted by IndexBuildConfig for API compatibility, build raises (out of scope)

``index_params`` keys: ``n_lists`` (default ``max(1, min(256, N // 1000 + 1))`` as in the
reference :394), ``pq_dim``, ``pq_bits``, ``metric`` ('sqeuclidean' | 'inner_product'),
``kmeans_n_iters``, ``kmeans_trainset_fraction``, ``seed``.

``CUVS_AVAILABLE`` keeps the reference's module-flag name: it is True when a CUDA device is
present (the native library is then REQUIRED — a missing libb2vs.so raises).  Only without CUDA
(or with the flag patched off, as the reference unit tests do) does ``_create_index`` return the
reference's simulated ``{"type", "size", "dim"}`` dict; that mode exists for host-side unit tests
and cannot be searched for real results.
"""
from __future__ import annotations

import logging
import time
from concurrent.futures import ThreadPoolExecutor, as_completed
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch

try:
    from gpu_resource_manager import GPUResourceManager
    from embedding_distribution_manager import DistributedEmbeddings, EmbeddingPart
    import _native
except ImportError:  # pragma: no cover - package-style import
    from .gpu_resource_manager import GPUResourceManager
    from .embedding_distribution_manager import DistributedEmbeddings, EmbeddingPart
    from . import _native

logger = logging.getLogger(__name__)

CUVS_AVAILABLE = bool(torch.cuda.is_available())
NATIVE_TYPES = ("ivf_flat", "ivf_pq", "brute_force", "flat")


@dataclass
class IndexBuildResult:
    gpu_id: int
    index: Any
    build_time: float
    success: bool
    error_message: Optional[str] = None
    memory_usage_bytes: int = 0

    def __post_init__(self):
        if self.gpu_id < 0:
            raise ValueError(f"gpu_id must be non-negative, got {self.gpu_id}")
        if self.build_time < 0:
            raise ValueError(f"build_time must be non-negative, got {self.build_time}")
        if self.success and self.index is None:
            raise ValueError("index cannot be None when success is True")
        if not self.success and self.error_message is None:
            raise ValueError("error_message cannot be None when success is False")


@dataclass
class IndexBuildConfig:
    index_type: str
    index_params: Dict[str, Any]
    search_params: Optional[Dict[str, Any]] = None
    parallel_build: bool = True
    max_retries: int = 2
    timeout_seconds: Optional[float] = None

    def __post_init__(self):
        valid_types = ["ivf_flat", "ivf_pq", "cagra", "brute_force", "flat"]
        if self.index_type not in valid_types:
            raise ValueError(f"index_type must be one of {valid_types}, got {self.index_type}")
        if not isinstance(self.index_params, dict):
            raise ValueError("index_params must be a dictionary")
        if self.max_retries < 0:
            raise ValueError(f"max_retries must be non-negative, got {self.max_retries}")
        if self.timeout_seconds is not None and self.timeout_seconds <= 0:
            raise ValueError(f"timeout_seconds must be positive, got {self.timeout_seconds}")


@dataclass
class CoordinatedIndexBuild:
    build_results: List[IndexBuildResult]
    total_build_time: float
    success: bool
    failed_gpus: List[int]
    successful_gpus: List[int]
    config: IndexBuildConfig

    def __post_init__(self):
        if not self.build_results:
            raise ValueError("build_results cannot be empty")
        if self.total_build_time < 0:
            raise ValueError(f"total_build_time must be non-negative, got {self.total_build_time}")
        seen = {r.gpu_id for r in self.build_results}
        if set(self.failed_gpus) | set(self.successful_gpus) != seen:
            raise ValueError("failed_gpus and successful_gpus must match build_results GPU IDs")


class IndexBuildingCoordinator:
    def __init__(self, gpu_manager: GPUResourceManager):
        if gpu_manager is None or not hasattr(gpu_manager, "validate_gpu_index"):
            raise TypeError("gpu_manager must be a GPUResourceManager instance")
        self.gpu_manager = gpu_manager
        self.built_indices: Dict[int, Any] = {}
        self.build_history: List[CoordinatedIndexBuild] = []
        self._active_builds: Dict[int, bool] = {}
        # id offset (EmbeddingPart.start_index) of the shard each built index covers
        self.index_offsets: Dict[int, int] = {}

    # ------------------------------------------------------------------ orchestration
    def build_indices_parallel(self, distributed_embeddings: DistributedEmbeddings,
                               config: IndexBuildConfig) -> CoordinatedIndexBuild:
        if not isinstance(distributed_embeddings, DistributedEmbeddings):
            raise ValueError("distributed_embeddings must be a DistributedEmbeddings instance")
        if not isinstance(config, IndexBuildConfig):
            raise ValueError("config must be an IndexBuildConfig instance")
        t0 = time.time()
        self._cleanup_existing_indices()
        parts = list(distributed_embeddings.parts)
        for p in parts:
            self._active_builds[p.gpu_id] = True
        try:
            if config.parallel_build and len(parts) > 1:
                results = self._build_parallel(parts, config)
            else:
                results = self._build_sequential(parts, config)
        finally:
            for p in parts:
                self._active_builds[p.gpu_id] = False
        ok = sorted(r.gpu_id for r in results if r.success)
        bad = sorted(r.gpu_id for r in results if not r.success)
        for r in results:
            if r.success:
                self.built_indices[r.gpu_id] = r.index
        for p in parts:
            if p.gpu_id in ok:
                self.index_offsets[p.gpu_id] = p.start_index
        build = CoordinatedIndexBuild(results, time.time() - t0, not bad, bad, ok, config)
        self.build_history.append(build)
        if bad:
            logger.warning("index build failed on GPUs %s (ok on %s)", bad, ok)
        return build

    def _build_parallel(self, parts: List[EmbeddingPart], config: IndexBuildConfig
                        ) -> List[IndexBuildResult]:
        # one host thread per GPU: ctypes drops the GIL, so the per-GPU builds overlap
        results: Dict[int, IndexBuildResult] = {}
        with ThreadPoolExecutor(max_workers=len(parts)) as pool:
            futs = {pool.submit(self._build_single_index, p, config): p for p in parts}
            try:
                for fut in as_completed(futs, timeout=config.timeout_seconds):
                    p = futs[fut]
                    try:
                        results[p.gpu_id] = fut.result()
                    except Exception as exc:
                        results[p.gpu_id] = IndexBuildResult(p.gpu_id, None, 0.0, False,
                                                             f"Build failed: {exc}")
            except Exception as exc:  # timeout of as_completed
                for fut, p in futs.items():
                    if p.gpu_id not in results:
                        fut.cancel()
                        results[p.gpu_id] = IndexBuildResult(p.gpu_id, None, 0.0, False,
                                                             f"Build timed out: {exc}")
        return [results[p.gpu_id] for p in parts]

    def _build_sequential(self, parts: List[EmbeddingPart], config: IndexBuildConfig
                          ) -> List[IndexBuildResult]:
        out = []
        for p in parts:
            try:
                out.append(self._build_single_index(p, config))
            except Exception as exc:
                out.append(IndexBuildResult(p.gpu_id, None, 0.0, False, f"Build failed: {exc}"))
        return out

    def _build_single_index(self, part: EmbeddingPart, config: IndexBuildConfig) -> IndexBuildResult:
        """Build on one GPU; up to ``max_retries`` extra attempts with linear back-off."""
        gpu = part.gpu_id
        last_error = "unknown error"
        for attempt in range(config.max_retries + 1):
            t0 = time.time()
            try:
                if not self.gpu_manager.validate_gpu_index(gpu):
                    raise RuntimeError(f"GPU {gpu} is not available")
                index = self._create_index(part.tensor, config, gpu_id=gpu, id_offset=part.start_index)
                elapsed = time.time() - t0
                try:
                    mem = int(self.gpu_manager.get_gpu_memory_info(gpu)["allocated"])
                except Exception:
                    mem = 0
                if not self.validate_index_build(gpu, index, part.tensor):
                    raise RuntimeError("index validation failed")
                return IndexBuildResult(gpu, index, elapsed, True, None, mem)
            except Exception as exc:
                last_error = str(exc)
                logger.warning("GPU %d build attempt %d/%d failed: %s", gpu, attempt + 1,
                               config.max_retries + 1, exc)
                if attempt < config.max_retries:
                    try:
                        self.gpu_manager.cleanup_gpu_resources([gpu])
                    except Exception:
                        pass
                    time.sleep(0.05 * (attempt + 1) if not CUVS_AVAILABLE else 0.5 * (attempt + 1))
        return IndexBuildResult(gpu, None, 0.0, False,
                                f"Failed after {config.max_retries + 1} attempts: {last_error}")

    # ------------------------------------------------------------------ the L2 boundary
    def _create_index(self, embeddings: Any, config: IndexBuildConfig, gpu_id: Optional[int] = None,
                      id_offset: int = 0) -> Any:
        try:
            n, d = int(embeddings.shape[0]), int(embeddings.shape[1])
            if not CUVS_AVAILABLE or not torch.cuda.is_available():
                logger.warning("no CUDA device: returning a SIMULATED %s index (unit-test mode)",
                               config.index_type)
                return {"type": config.index_type, "size": n, "dim": d}
            if config.index_type == "cagra":
                raise NotImplementedError("index_type 'cagra' (graph index) is out of scope; use "
                                          "'ivf_flat', 'ivf_pq' or 'brute_force'")
            if not embeddings.is_cuda:
                raise ValueError(f"shard must be resident on its GPU (got {embeddings.device})")
            p = config.index_params
            metric = p.get("metric", "sqeuclidean")
            stream = None
            if gpu_id is not None and hasattr(self.gpu_manager, "get_stream"):
                try:
                    s = self.gpu_manager.get_stream(gpu_id)
                    stream = s if isinstance(s, torch.cuda.Stream) else None
                except Exception:
                    stream = None
            if stream is not None:
                stream.wait_stream(torch.cuda.current_stream(embeddings.device))
            if config.index_type in ("brute_force", "flat"):
                ix = _native.NativeIndex.flat(embeddings, metric=metric, id_offset=id_offset,
                                              stream=stream)
            else:
                n_lists = int(p.get("n_lists", max(1, min(256, n // 1000 + 1))))
                n_lists = max(1, min(n_lists, n))
                common = dict(metric=metric, id_offset=id_offset,
                              kmeans_iters=int(p.get("kmeans_n_iters", 20)),
                              train_fraction=float(p.get("kmeans_trainset_fraction", 0.5)),
                              seed=int(p.get("seed", 0)), stream=stream)
                if config.index_type == "ivf_flat":
                    ix = _native.NativeIndex.ivf_flat(embeddings, n_lists, **common)
                else:
                    # reference default: pq_dim = min(64, dim // 4) (index_building_coordinator.py:401),
                    # snapped to a divisor of dim the engine supports
                    pq_dim = int(p.get("pq_dim", 0) or _native.NativeIndex.default_pq_dim(d))
                    ix = _native.NativeIndex.ivf_pq(embeddings, n_lists, pq_dim,
                                                    int(p.get("pq_bits", 8)), **common)
            if stream is not None:
                stream.synchronize()
            return ix
        except Exception as exc:
            raise RuntimeError(f"Failed to create {config.index_type} index: {exc}") from exc

    def validate_index_build(self, gpu_id: int, index: Any, embeddings: Any) -> bool:
        try:
            if index is None:
                return False
            n, d = int(embeddings.shape[0]), int(embeddings.shape[1])
            if isinstance(index, dict):  # simulated index
                return index.get("size") == n and index.get("dim") == d
            if isinstance(index, _native.NativeIndex):
                inf = index.info()
                return inf.n_rows == n and inf.dim == d and inf.device == gpu_id
            return True
        except Exception as exc:
            logger.error("validating index on GPU %d raised: %s", gpu_id, exc)
            return False

    # ------------------------------------------------------------------ bookkeeping
    def _drop_index(self, gpu_id: int) -> None:
        ix = self.built_indices.pop(gpu_id, None)
        self.index_offsets.pop(gpu_id, None)
        if isinstance(ix, _native.NativeIndex):
            ix.destroy()

    def _cleanup_existing_indices(self) -> None:
        for g in list(self.built_indices):
            self._drop_index(g)

    def cleanup_failed_builds(self, failed_gpus: List[int]) -> None:
        for g in failed_gpus:
            self._drop_index(g)
            self._active_builds.pop(g, None)
        if failed_gpus:
            try:
                self.gpu_manager.cleanup_gpu_resources(list(failed_gpus))
            except Exception as exc:
                logger.error("cleanup after failed builds raised: %s", exc)

    def get_built_indices(self) -> Dict[int, Any]:
        return dict(self.built_indices)

    def get_index_for_gpu(self, gpu_id: int) -> Optional[Any]:
        return self.built_indices.get(gpu_id)

    def has_active_builds(self) -> bool:
        return any(self._active_builds.values())

    def get_active_build_gpus(self) -> List[int]:
        return [g for g, active in self._active_builds.items() if active]

    def get_build_history(self) -> List[CoordinatedIndexBuild]:
        return list(self.build_history)

    def get_build_summary(self) -> Dict[str, Any]:
        attempts: Dict[int, int] = {}
        wins: Dict[int, int] = {}
        for build in self.build_history:
            for r in build.build_results:
                attempts[r.gpu_id] = attempts.get(r.gpu_id, 0) + 1
                wins[r.gpu_id] = wins.get(r.gpu_id, 0) + (1 if r.success else 0)
        return {
            "total_coordinated_builds": len(self.build_history),
            "successful_coordinated_builds": sum(1 for b in self.build_history if b.success),
            "current_built_indices": len(self.built_indices),
            "active_builds": sum(1 for a in self._active_builds.values() if a),
            "gpu_success_rates": {g: wins[g] / attempts[g] for g in attempts},
        }

    def cleanup_all_indices(self) -> None:
        gpus = sorted(self.built_indices)
        self._cleanup_existing_indices()
        self._active_builds.clear()
        if gpus:
            self.gpu_manager.cleanup_gpu_resources(gpus)

    def __str__(self) -> str:
        return (f"IndexBuildingCoordinator(built_indices={len(self.built_indices)}, "
                f"active_builds={sum(1 for a in self._active_builds.values() if a)})")

    def __repr__(self) -> str:
        return (f"IndexBuildingCoordinator(gpu_manager={self.gpu_manager!r}, "
                f"built_indices={sorted(self.built_indices)}, "
                f"build_history={len(self.build_history)})")
