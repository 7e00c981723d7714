"""EmbeddingDistributionManager — shards an [N, D] embedding matrix across GPUs by row range.

Drop-in for the reference's ``Attempt_1/embedding_distribution_manager.py`` (same names and error
messages; its ``test_embedding_distribution_manager.py`` is the acceptance spec, including the
API the reference implementation itself lacks: ``get_total_memory_usage``,
``cleanup_distribution``, ``get_distribution_summary`` — SURVEY.md Appendix B).

Data layout: shard g holds rows ``[start_g, end_g)`` of the corpus as one contiguous row-major
tensor on ``cuda:g``; ``EmbeddingPart.start_index`` is the id offset every search on that shard
adds to its local row ids (the reference notebooks use ``i * len(part)``, which is wrong for
uneven shards — SURVEY.md §3.6 bug 1).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch

try:  # flat import (reference style: package directory on sys.path) or package import
    from gpu_resource_manager import GPUResourceManager
except ImportError:  # pragma: no cover
    from .gpu_resource_manager import GPUResourceManager

logger = logging.getLogger(__name__)


def _rows_of(t: Any) -> Optional[int]:
    """Row count of a tensor-like (real tensors, or the Mock tensors the reference tests use)."""
    shape = getattr(t, "shape", None)
    if isinstance(shape, (tuple, list, torch.Size)) and len(shape) >= 1 and isinstance(shape[0], int):
        return int(shape[0])
    try:
        n = t.size(0)
        return int(n) if isinstance(n, int) else None
    except Exception:
        return None


def _cols_of(t: Any) -> Optional[int]:
    shape = getattr(t, "shape", None)
    if isinstance(shape, (tuple, list, torch.Size)) and len(shape) >= 2 and isinstance(shape[1], int):
        return int(shape[1])
    try:
        n = t.size(1)
        return int(n) if isinstance(n, int) else None
    except Exception:
        return None


@dataclass
class EmbeddingPart:
    """Rows [start_index, end_index) of the corpus, resident on one GPU."""
    gpu_id: int
    tensor: Any
    start_index: int
    end_index: int

    def __post_init__(self):
        if self.gpu_id < 0:
            raise ValueError(f"gpu_id must be non-negative, got {self.gpu_id}")
        if self.start_index < 0:
            raise ValueError(f"start_index must be non-negative, got {self.start_index}")
        if self.end_index <= self.start_index:
            raise ValueError(f"end_index ({self.end_index}) must be greater than start_index "
                             f"({self.start_index})")
        rows = _rows_of(self.tensor)
        if rows is not None and rows != self.end_index - self.start_index:
            raise ValueError(f"Tensor size ({rows}) doesn't match index range "
                             f"({self.end_index - self.start_index})")

    @property
    def num_rows(self) -> int:
        return self.end_index - self.start_index


@dataclass
class DistributedEmbeddings:
    """All parts of one corpus; parts must tile [0, total_size) without gaps or overlaps."""
    parts: List[EmbeddingPart]
    total_size: int
    embedding_dim: int

    def __post_init__(self):
        if not self.parts:
            raise ValueError("parts list cannot be empty")
        if self.total_size <= 0:
            raise ValueError(f"total_size must be positive, got {self.total_size}")
        if self.embedding_dim <= 0:
            raise ValueError(f"embedding_dim must be positive, got {self.embedding_dim}")
        for pos, part in enumerate(self.parts):
            cols = _cols_of(part.tensor)
            if cols is not None and cols != self.embedding_dim:
                raise ValueError(f"Part {pos} has embedding_dim {cols}, expected {self.embedding_dim}")
        expect = 0
        for part in sorted(self.parts, key=lambda p: p.start_index):
            if part.start_index != expect:
                raise ValueError(f"Gap or overlap detected at index {expect}: next part starts at "
                                 f"{part.start_index}")
            expect = part.end_index
        if expect != self.total_size:
            raise ValueError(f"Parts cover {expect} rows but total_size is {self.total_size}")


STAGE_MIN_BYTES = 64 << 20     # host shards below this go up in one plain copy
STAGE_CHUNK_BYTES = 256 << 20   # size of each of the two pinned staging buffers


def staged_host_to_device(src: torch.Tensor, device: torch.device, dtype: Optional[torch.dtype] = None,
                          chunk_bytes: int = STAGE_CHUNK_BYTES) -> torch.Tensor:
    """Pageable host rows ``[n, D]`` -> one resident ``[n, D]`` device tensor in ``dtype``.

    The reference does ``embeddings[start:end].clone().to('cuda:i')``
    (``embedding_distribution_manager.py:161-166``): an extra host copy, a pageable (slow,
    synchronous) H2D of the whole shard, and — when the shard is to live in 16 bits — the full
    fp32 shard on the device next to its converted copy (C5: 205 GB of fp32 for a 102 GB bf16
    corpus, more than one B200 holds).  Here the shard goes up in chunks through two pinned
    staging buffers on a copy stream: the host memcpy of chunk i+1 overlaps the H2D + on-device
    conversion of chunk i, and the transient device footprint is one chunk.
    """
    n, d = int(src.shape[0]), int(src.shape[1])
    out = torch.empty((n, d), dtype=dtype or src.dtype, device=device)
    rows = max(1, int(chunk_bytes) // max(1, d * src.element_size()))
    copy_stream = torch.cuda.Stream(device)
    copy_stream.wait_stream(torch.cuda.current_stream(device))
    pinned_src = src.is_pinned()
    stage = [] if pinned_src else [torch.empty((min(rows, n), d), dtype=src.dtype, pin_memory=True)
                                   for _ in range(2)]
    done = [None, None]
    for ci, s0 in enumerate(range(0, n, rows)):
        e0 = min(n, s0 + rows)
        if pinned_src:
            host = src[s0:e0]
        else:
            b = ci & 1
            if done[b] is not None:
                done[b].synchronize()          # the H2D that was reading this staging buffer
            host = stage[b][: e0 - s0]
            host.copy_(src[s0:e0])
        with torch.cuda.stream(copy_stream):
            out[s0:e0].copy_(host, non_blocking=True)      # H2D (+ conversion on the device)
            if not pinned_src:
                done[b] = torch.cuda.Event()
                done[b].record(copy_stream)
    torch.cuda.current_stream(device).wait_stream(copy_stream)
    if not pinned_src:
        copy_stream.synchronize()              # the staging buffers are freed on return
    return out


class EmbeddingDistributionManager:
    def __init__(self, gpu_manager: GPUResourceManager):
        # duck-typed on purpose: the reference tests pass Mock(spec=GPUResourceManager)
        if gpu_manager is None or not hasattr(gpu_manager, "distribute_workload"):
            raise TypeError("gpu_manager must be a GPUResourceManager instance")
        self.gpu_manager = gpu_manager
        self.current_distribution: Optional[DistributedEmbeddings] = None

    # ------------------------------------------------------------------ distribute
    def distribute_embeddings(self, embeddings: torch.Tensor,
                              target_gpus: Optional[List[int]] = None,
                              dtype: Optional[torch.dtype] = None,
                              strategy: str = "even") -> DistributedEmbeddings:
        """Shard ``embeddings`` [N, D] over ``target_gpus`` (default: all available GPUs).

        No intermediate host clone; ``dtype`` optionally converts the shard ON the device (fp32
        host data -> bf16/fp16 resident shards, config C2/C3).  Host shards of 64 MB and more go
        up in pinned, double-buffered chunks (``staged_host_to_device``), smaller ones in one copy.
        """
        if not isinstance(embeddings, torch.Tensor):
            raise TypeError("embeddings must be a torch.Tensor")
        if embeddings.dim() != 2:
            raise ValueError(f"embeddings must be 2D tensor, got {embeddings.dim()}D")
        if embeddings.size(0) == 0:
            raise ValueError("embeddings tensor cannot be empty")
        n, d = int(embeddings.size(0)), int(embeddings.size(1))
        available = self.gpu_manager.get_available_gpu_ids()
        if target_gpus is None:
            target_gpus = list(available)
        if not target_gpus:
            raise RuntimeError("No GPUs available for embedding distribution")
        for g in target_gpus:
            if not self.gpu_manager.validate_gpu_index(g):
                raise ValueError(f"Target GPU {g} is not available")
        if list(target_gpus) == list(available):
            ranges = self.gpu_manager.distribute_workload(n, strategy)
        else:
            ranges = self.gpu_manager.distribute_workload(n, strategy, gpu_ids=list(target_gpus))
        parts: List[EmbeddingPart] = []
        try:
            for gpu_id, start, end in ranges:
                if end <= start:
                    continue  # more GPUs than rows
                device_string = self.gpu_manager.get_safe_device_string(gpu_id)
                shard = self._upload(embeddings[start:end], device_string, dtype)
                if isinstance(shard, torch.Tensor) and not shard.is_contiguous():
                    shard = shard.contiguous()
                parts.append(EmbeddingPart(gpu_id, shard, start, end))
            dist = DistributedEmbeddings(parts, n, d)
        except Exception as exc:
            self._release_parts(parts)
            raise RuntimeError(f"Failed to distribute embeddings: {exc}") from exc
        if not self.validate_distribution(dist, _available=available):
            logger.warning("distribution failed device validation (expected only in mocked tests)")
        self.current_distribution = dist
        return dist

    @staticmethod
    def _upload(rows: torch.Tensor, device_string: str, dtype: Optional[torch.dtype]) -> torch.Tensor:
        """One shard onto its device: staged for large host tensors, a plain copy otherwise
        (device-resident input, small shards, and the mocked / CPU-only unit-test mode)."""
        try:
            device = torch.device(device_string)
        except (TypeError, RuntimeError):
            device = None
        if (device is not None and device.type == "cuda" and isinstance(rows, torch.Tensor)
                and not rows.is_cuda and rows.dim() == 2 and torch.cuda.is_available()
                and rows.numel() * rows.element_size() >= STAGE_MIN_BYTES):
            return staged_host_to_device(rows, device, dtype, STAGE_CHUNK_BYTES)
        shard = rows.to(device_string)
        if dtype is not None and shard.dtype != dtype:
            shard = shard.to(dtype)
        return shard

    def load_embedding_parts(self, paths: List[str], dtype: Optional[torch.dtype] = None,
                             target_gpus: Optional[List[int]] = None) -> DistributedEmbeddings:
        """Load the reference's on-disk embedding format (``cuvs-2gpu-main.ipynb`` cells 10/12):
        torch-saved ``[n, D]`` tensors, either one file per GPU (``embeddings_{size}_part{i}.pt``)
        or a single ``embeddings_{size}.pt`` that is split like ``torch.chunk`` over the GPUs.
        Part i goes to the i-th GPU; ``start_index`` is the running row count, so global ids are
        right for uneven parts (380 000 + 370 000 in the reference's 750 k run)."""
        if not paths:
            raise ValueError("paths cannot be empty")
        if len(paths) == 1:
            full = torch.load(paths[0], map_location="cpu")
            return self.distribute_embeddings(full, target_gpus=target_gpus, dtype=dtype)
        gpus = list(target_gpus) if target_gpus is not None else self.gpu_manager.get_available_gpu_ids()
        if len(gpus) < len(paths):
            raise RuntimeError(f"{len(paths)} embedding parts but only {len(gpus)} GPUs available")
        parts: List[EmbeddingPart] = []
        start, dim = 0, None
        for gpu_id, path in zip(gpus, paths):
            if not self.gpu_manager.validate_gpu_index(gpu_id):
                raise ValueError(f"Target GPU {gpu_id} is not available")
            t = torch.load(path, map_location="cpu")
            if not isinstance(t, torch.Tensor) or t.dim() != 2:
                raise ValueError(f"{path} does not hold a 2D embedding tensor")
            if dim is None:
                dim = int(t.shape[1])
            shard = self._upload(t.contiguous(), self.gpu_manager.get_safe_device_string(gpu_id), dtype)
            parts.append(EmbeddingPart(gpu_id, shard, start, start + int(t.shape[0])))
            start += int(t.shape[0])
        dist = DistributedEmbeddings(parts, start, dim)
        self.current_distribution = dist
        return dist

    # ------------------------------------------------------------------ validate
    def validate_distribution(self, distributed_embeddings: DistributedEmbeddings,
                              _available: Optional[List[int]] = None) -> bool:
        try:
            if not isinstance(distributed_embeddings, DistributedEmbeddings):
                return False
            available = _available if _available is not None else self.gpu_manager.get_available_gpu_ids()
            for part in distributed_embeddings.parts:
                if isinstance(available, (list, tuple)) and part.gpu_id not in available:
                    logger.error("part on GPU %d: GPU not in the available list %s", part.gpu_id, available)
                    return False
                if not self.gpu_manager.validate_gpu_index(part.gpu_id):
                    logger.error("part on GPU %d: GPU is not available", part.gpu_id)
                    return False
                dev = getattr(part.tensor, "device", None)
                on_gpu = str(dev) == f"cuda:{part.gpu_id}" or (
                    getattr(dev, "type", None) == "cuda" and getattr(dev, "index", None) == part.gpu_id)
                if not on_gpu:
                    logger.error("part for GPU %d lives on %s", part.gpu_id, dev)
                    return False
                rows = _rows_of(part.tensor)
                if rows is not None and rows != part.num_rows:
                    return False
            return True
        except Exception as exc:
            logger.error("distribution validation raised: %s", exc)
            return False

    # ------------------------------------------------------------------ reshard
    def redistribute_if_needed(self, distributed_embeddings: DistributedEmbeddings
                               ) -> DistributedEmbeddings:
        """Return the distribution unchanged when every shard's GPU is still usable; otherwise
        gather on the host and re-shard over the surviving GPUs."""
        available = self.gpu_manager.get_available_gpu_ids()
        used = [p.gpu_id for p in distributed_embeddings.parts]
        if all(g in available for g in used) and self.validate_distribution(distributed_embeddings):
            return distributed_embeddings
        logger.warning("re-sharding: GPUs %s no longer all usable (available %s)", used, available)
        peer = self._reshard_device_to_device(distributed_embeddings, list(available))
        if peer is not None:
            return peer
        full = self._gather_embeddings_to_cpu(distributed_embeddings)
        self._release_parts(distributed_embeddings.parts)
        return self.distribute_embeddings(full)

    def _reshard_device_to_device(self, dist: DistributedEmbeddings,
                                  target_gpus: List[int]) -> Optional[DistributedEmbeddings]:
        """Elastic re-shard without the host round trip (the reference gathers
        everything on the CPU, ``embedding_distribution_manager.py:274-334``): the new shards are
        allocated on the surviving GPUs and filled by direct device-to-device copies of the
        overlapping row ranges of the old shards (NVLink peer copies between GPUs).  Returns None
        when the shards are not CUDA tensors or a copy fails (source GPU really gone) — the caller
        then falls back to the host route."""
        parts = sorted(dist.parts, key=lambda p: p.start_index)
        if not target_gpus or not parts or not all(
                isinstance(p.tensor, torch.Tensor) and p.tensor.is_cuda for p in parts):
            return None
        try:
            ranges = self.gpu_manager.distribute_workload(dist.total_size, "even", gpu_ids=list(target_gpus))
        except TypeError:          # a manager without the gpu_ids extension
            return None
        new_parts: List[EmbeddingPart] = []
        try:
            for gpu_id, start, end in ranges:
                if end <= start:
                    continue
                keep = next((p for p in parts if p.gpu_id == gpu_id and p.start_index == start
                             and p.end_index == end), None)
                if keep is not None:                       # shard already where it belongs
                    new_parts.append(keep)
                    continue
                dev = torch.device(self.gpu_manager.get_safe_device_string(gpu_id))
                shard = torch.empty((end - start, dist.embedding_dim), dtype=parts[0].tensor.dtype, device=dev)
                for p in parts:
                    lo, hi = max(start, p.start_index), min(end, p.end_index)
                    if lo < hi:
                        shard[lo - start:hi - start].copy_(
                            p.tensor[lo - p.start_index:hi - p.start_index], non_blocking=True)
                new_parts.append(EmbeddingPart(gpu_id, shard, start, end))
            for g in {p.gpu_id for p in new_parts} | {p.gpu_id for p in parts}:
                torch.cuda.synchronize(g)
        except Exception as exc:
            logger.warning("device-to-device re-shard failed (%s): falling back to the host route", exc)
            self._release_parts([p for p in new_parts if p not in parts])
            return None
        kept = {id(p) for p in new_parts}
        self._release_parts([p for p in parts if id(p) not in kept])
        out = DistributedEmbeddings(new_parts, dist.total_size, dist.embedding_dim)
        self.current_distribution = out
        return out

    def _gather_embeddings_to_cpu(self, distributed_embeddings: DistributedEmbeddings) -> torch.Tensor:
        ordered = sorted(distributed_embeddings.parts, key=lambda p: p.start_index)
        return torch.cat([p.tensor.cpu() for p in ordered], dim=0)

    # ------------------------------------------------------------------ bookkeeping
    def get_embedding_part_by_gpu(self, distributed_embeddings: DistributedEmbeddings,
                                  gpu_id: int) -> Optional[EmbeddingPart]:
        for part in distributed_embeddings.parts:
            if part.gpu_id == gpu_id:
                return part
        return None

    def get_total_memory_usage(self, distributed_embeddings: DistributedEmbeddings) -> Dict[int, int]:
        usage: Dict[int, int] = {}
        for part in distributed_embeddings.parts:
            t = part.tensor
            try:
                nbytes = int(t.numel()) * int(t.element_size())
            except Exception:
                nbytes = 0
            usage[part.gpu_id] = usage.get(part.gpu_id, 0) + nbytes
        return usage

    # name used by the reference implementation (embedding_distribution_manager.py:373)
    get_total_gpu_memory_usage = get_total_memory_usage

    def get_distribution_summary(self, distributed_embeddings: DistributedEmbeddings) -> Dict[str, Any]:
        usage = self.get_total_memory_usage(distributed_embeddings)
        return {
            "total_embeddings": distributed_embeddings.total_size,
            "embedding_dimension": distributed_embeddings.embedding_dim,
            "num_gpus": len(distributed_embeddings.parts),
            "gpu_ids": [p.gpu_id for p in distributed_embeddings.parts],
            "part_sizes": [p.num_rows for p in distributed_embeddings.parts],
            "memory_usage_bytes": usage,
            "memory_usage_mb": {g: b / (1024 ** 2) for g, b in usage.items()},
        }

    def _release_parts(self, parts: List[EmbeddingPart]) -> None:
        gpus = sorted({p.gpu_id for p in parts})
        for p in parts:
            p.tensor = None
        if gpus:
            self.gpu_manager.cleanup_gpu_resources(gpus)

    def cleanup_distribution(self, distributed_embeddings: Optional[DistributedEmbeddings] = None) -> None:
        """Drop the shards of ``distributed_embeddings`` (default: the current distribution)."""
        target = distributed_embeddings if distributed_embeddings is not None else self.current_distribution
        if target is None:
            return
        gpus = [p.gpu_id for p in target.parts]
        if target is self.current_distribution:
            self.current_distribution = None
        if distributed_embeddings is None:
            for p in target.parts:
                p.tensor = None
        self.gpu_manager.cleanup_gpu_resources(gpus)

    def cleanup_current_distribution(self) -> None:
        self.cleanup_distribution(None)

    def __str__(self) -> str:
        n = len(self.current_distribution.parts) if self.current_distribution else 0
        return f"EmbeddingDistributionManager(parts={n})"

    def __repr__(self) -> str:
        return (f"EmbeddingDistributionManager(gpu_manager={self.gpu_manager!r}, "
                f"has_current_distribution={self.current_distribution is not None})")
