"""ctypes binding of libb2vs.so (include/b2vs.h) and the build recipe for it.

The library is the product path: every search / build / merge in this package goes through
these entry points.  There is deliberately NO Python or CPU fallback here — if the shared
library is missing or a call fails, a RuntimeError carrying ``b2vs_last_error()`` is raised.
PyTorch only supplies device memory (``tensor.data_ptr()``) and streams.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("B2VS_LIB_PATH") or os.path.join(_HERE, "libb2vs.so")
SOURCES = ["api.cu", "flat.cu", "merge.cu", "kmeans.cu", "ivf_build.cu", "ivf_plan.cu", "ivf_scan.cu",
           "ivf_search.cu", "ivf_graph.cu", "persist.cu", "bigk.cu", "pq_tc.cu", "encode.cu", "cosine.cu",
           "comm.cu"]

METRIC_L2, METRIC_IP, METRIC_COSINE = 0, 1, 2
F32, F16, BF16 = 0, 1, 2
KIND_FLAT, KIND_IVF_FLAT, KIND_IVF_PQ = 0, 1, 2
MAX_FUSED_K = 128
FLAG_TIME_KERNEL, FLAG_TC_SINGLE, FLAG_TC_PAIR, FLAG_GRAPH = 1, 2, 4, 8   # b2vs_search_params.flags
MAX_K = 2048   # flat, IVF-Flat and (grouped-scan shapes) IVF-PQ indexes, merges

_DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}
_METRIC_CODE = {
    "sqeuclidean": METRIC_L2, "l2": METRIC_L2, "euclidean": METRIC_L2, "L2": METRIC_L2,
    "inner_product": METRIC_IP, "ip": METRIC_IP, "IP": METRIC_IP, "dot": METRIC_IP,
    # cosine DISTANCE (1 - cos), ascending: scikit-learn's metric='cosine', the reference's CPU
    # baseline (VectorSearch_QuestionRetrieval.ipynb:L878); the index owns a unit-norm row copy
    "cosine": METRIC_COSINE,
}


class IvfParams(ctypes.Structure):
    _fields_ = [("n_lists", ctypes.c_int32), ("kmeans_iters", ctypes.c_int32),
                ("train_fraction", ctypes.c_float), ("pq_dim", ctypes.c_int32),
                ("pq_bits", ctypes.c_int32), ("seed", ctypes.c_uint64)]


class SearchParams(ctypes.Structure):
    _fields_ = [("n_probes", ctypes.c_int32), ("refine_ratio", ctypes.c_int32),
                ("n_splits", ctypes.c_int32), ("flags", ctypes.c_int32)]


class IndexInfo(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("device", ctypes.c_int32), ("metric", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("dim", ctypes.c_int32), ("n_lists", ctypes.c_int32),
                ("pq_dim", ctypes.c_int32), ("pq_bits", ctypes.c_int32),
                ("n_rows", ctypes.c_int64), ("id_offset", ctypes.c_int64),
                ("device_bytes", ctypes.c_int64)]


class SearchStats(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_int32), ("n_splits", ctypes.c_int32),
                ("grid", ctypes.c_int32), ("mean_candidates", ctypes.c_int32),
                ("algo_flops", ctypes.c_double), ("algo_bytes", ctypes.c_double),
                ("kernel_ms", ctypes.c_double), ("distinct_bytes", ctypes.c_double)]


# Every symbol include/b2vs.h declares; tests assert the built library exports all of them.
EXPORTS = [
    "b2vs_last_error", "b2vs_version", "b2vs_device_count", "b2vs_bf_create",
    "b2vs_ivfflat_build", "b2vs_ivfpq_build", "b2vs_search", "b2vs_search_host",
    "b2vs_merge_topk", "b2vs_kmeans_fit", "b2vs_index_info_get", "b2vs_index_last_stats",
    "b2vs_ivf_list_sizes_host", "b2vs_ivf_centroids_host", "b2vs_index_destroy",
    "b2vs_index_save", "b2vs_index_load", "b2vs_pool_normalize", "b2vs_reload_env",
    "b2vs_comm_unique_id", "b2vs_comm_init_rank", "b2vs_comm_init_all", "b2vs_comm_info",
    "b2vs_comm_destroy", "b2vs_partition_even", "b2vs_allgather_queries", "b2vs_allgather_topk",
    "b2vs_allgather_merge_topk", "b2vs_exchange_merge_topk", "b2vs_allreduce_min_f32",
    "b2vs_search_sharded", "b2vs_search_sharded_host", "b2vs_comm_register_index",
    "b2vs_debug_check_canaries",
]
UNIQUE_ID_BYTES = 128

_lib = None
_lib_lock = threading.Lock()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libb2vs.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a (cross-compiles without a GPU) and link libb2vs.so.  Every
    translation unit is compiled to an object file in parallel; an object is rebuilt when its
    source or any header is newer."""
    from concurrent.futures import ThreadPoolExecutor
    inc = os.path.join(os.path.dirname(_HERE), "include", "b2vs.h")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [inc]
    newest_header = max(os.path.getmtime(h) for h in headers)
    obj_dir = os.path.join(_HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
             "-Xcompiler", "-fPIC"]
    if verbose:
        flags.append("-Xptxas=-v")
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s[:-3] + ".o")
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header))
        jobs.append((src, obj, stale))
    if not any(j[2] for j in jobs) and os.path.exists(LIB_PATH) and \
            os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(j[1]) for j in jobs):
        return LIB_PATH

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return ""
        res = subprocess.run([nvcc_path(), *flags, "-c", "-o", obj + ".tmp", src],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + res.stdout + res.stderr)
        os.replace(obj + ".tmp", obj)
        return res.stderr

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        logs = list(pool.map(compile_one, jobs))
    res = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                          "-o", LIB_PATH + ".tmp", *[j[1] for j in jobs], "-ldl"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print("".join(logs))
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load libb2vs.so (once). Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the CUDA extension is required; there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.b2vs_last_error.restype = ctypes.c_char_p
        L.b2vs_last_error.argtypes = []
        L.b2vs_version.restype = i32
        L.b2vs_device_count.argtypes = [ctypes.POINTER(i32)]
        L.b2vs_bf_create.argtypes = [i32, i32, i32, i32, vp, i64, i64, vp, ctypes.POINTER(vp)]
        L.b2vs_ivfflat_build.argtypes = [i32, i32, i32, i32, vp, i64, i64,
                                         ctypes.POINTER(IvfParams), vp, ctypes.POINTER(vp)]
        L.b2vs_ivfpq_build.argtypes = L.b2vs_ivfflat_build.argtypes
        L.b2vs_search.argtypes = [vp, vp, i32, i32, i32, i32, ctypes.POINTER(SearchParams), vp, vp, vp]
        L.b2vs_search_host.argtypes = L.b2vs_search.argtypes
        L.b2vs_merge_topk.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
        L.b2vs_kmeans_fit.argtypes = [i32, i32, i32, vp, i64, i32, i32, ctypes.c_uint64, vp, vp, vp]
        L.b2vs_index_info_get.argtypes = [vp, ctypes.POINTER(IndexInfo)]
        L.b2vs_index_last_stats.argtypes = [vp, ctypes.POINTER(SearchStats)]
        L.b2vs_ivf_list_sizes_host.argtypes = [vp, vp]
        L.b2vs_ivf_centroids_host.argtypes = [vp, vp]
        L.b2vs_index_destroy.argtypes = [vp]
        L.b2vs_index_save.argtypes = [vp, ctypes.c_char_p]
        L.b2vs_index_load.argtypes = [i32, ctypes.c_char_p, vp, i64, vp, ctypes.POINTER(vp)]
        L.b2vs_pool_normalize.argtypes = [i32, i32, vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]
        L.b2vs_reload_env.argtypes = []
        pi32, pi64, pvp = ctypes.POINTER(i32), ctypes.POINTER(i64), ctypes.POINTER(vp)
        L.b2vs_comm_unique_id.argtypes = [vp]
        L.b2vs_comm_init_rank.argtypes = [i32, i32, i32, vp, pvp]
        L.b2vs_comm_init_all.argtypes = [i32, pi32, pvp]
        L.b2vs_comm_info.argtypes = [vp, pi32, pi32, pi32]
        L.b2vs_comm_destroy.argtypes = [vp]
        L.b2vs_partition_even.argtypes = [i64, i32, i32, pi64, pi64]
        L.b2vs_allgather_queries.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.b2vs_allgather_topk.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
        L.b2vs_allgather_merge_topk.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
        L.b2vs_exchange_merge_topk.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
        L.b2vs_allreduce_min_f32.argtypes = [vp, vp, i64, vp]
        L.b2vs_search_sharded.argtypes = [vp, vp, vp, i32, i32, i32, i32, ctypes.POINTER(SearchParams),
                                          vp, vp, vp]
        L.b2vs_search_sharded_host.argtypes = L.b2vs_search_sharded.argtypes
        L.b2vs_comm_register_index.argtypes = [vp, vp, vp]
        L.b2vs_debug_check_canaries.argtypes = [pi32, pi32]
        for name in EXPORTS:
            if name != "b2vs_last_error":
                getattr(L, name).restype = i32
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b2vs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def metric_code(metric) -> int:
    if isinstance(metric, int):
        return metric
    try:
        return _METRIC_CODE[metric]
    except KeyError:
        raise ValueError(f"unknown metric {metric!r}; expected one of {sorted(set(_METRIC_CODE))}")


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise ValueError(f"unsupported dtype {dtype}; expected float32, float16 or bfloat16")


def _stream_ptr(device: torch.device, stream: Optional[torch.cuda.Stream]) -> int:
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return int(s.cuda_stream)


def _require_cuda_matrix(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (got {t.device})")
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2D [rows, dim] (got shape {tuple(t.shape)})")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous row-major")


class NativeIndex:
    """Owner of one ``b2vs_index*`` living on one GPU."""

    def __init__(self, handle: int, device: torch.device, metric: int, keepalive=None):
        self._h = ctypes.c_void_p(handle)
        self.device = device
        self.metric = metric
        self._keepalive = keepalive  # borrowed database rows must outlive the index
        inf = IndexInfo()
        _check(lib().b2vs_index_info_get(self._h, ctypes.byref(inf)), "b2vs_index_info_get")
        self.dim = int(inf.dim)
        self.kind = int(inf.kind)
        # IVF-PQ refine re-ranks against the borrowed source rows; a cosine index owns a
        # normalised copy of them, so the caller's tensor need not be kept
        self.has_refine_rows = self.kind == KIND_IVF_PQ and keepalive is not None

    def _check_queries(self, queries: torch.Tensor) -> None:
        if queries.shape[1] != self.dim:
            raise ValueError(f"queries have dim {queries.shape[1]}, the index has dim {self.dim}")

    def _check_out(self, out, nq: int, k: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        d, i = out
        for t, dt, name in ((d, torch.float32, "distances"), (i, torch.int64, "ids")):
            if not isinstance(t, torch.Tensor) or tuple(t.shape) != (nq, k) or t.dtype != dt \
                    or not t.is_contiguous() or t.device != device:
                raise ValueError(f"out {name} must be a contiguous [{nq}, {k}] {dt} tensor on {device} "
                                 f"(got {getattr(t, 'shape', None)}, {getattr(t, 'dtype', None)}, "
                                 f"{getattr(t, 'device', None)})")
        return d, i

    def _check_refine(self, refine_ratio: int) -> None:
        if refine_ratio and refine_ratio > 1 and self.kind == KIND_IVF_PQ and not self.has_refine_rows:
            raise ValueError("refine_ratio > 1 needs the shard's source rows: load the index with "
                             "rows=<the shard's [n, dim] tensor>")

    # ------------------------------------------------------------------ constructors
    @classmethod
    def flat(cls, db: torch.Tensor, metric="sqeuclidean", id_offset: int = 0,
             stream: Optional[torch.cuda.Stream] = None) -> "NativeIndex":
        _require_cuda_matrix(db, "db")
        L = lib()
        out = ctypes.c_void_p()
        m = metric_code(metric)
        _check(L.b2vs_bf_create(db.device.index, m, dtype_code(db.dtype), db.shape[1],
                                db.data_ptr(), db.shape[0], int(id_offset),
                                _stream_ptr(db.device, stream), ctypes.byref(out)),
               "b2vs_bf_create")
        return cls(out.value, db.device, m, keepalive=db)

    @classmethod
    def ivf_flat(cls, db: torch.Tensor, n_lists: int, metric="sqeuclidean", id_offset: int = 0,
                 kmeans_iters: int = 20, train_fraction: float = 0.5, seed: int = 0,
                 stream: Optional[torch.cuda.Stream] = None) -> "NativeIndex":
        return cls._ivf(db, "b2vs_ivfflat_build", n_lists, 0, 0, metric, id_offset, kmeans_iters,
                        train_fraction, seed, stream)

    @classmethod
    def ivf_pq(cls, db: torch.Tensor, n_lists: int, pq_dim: int, pq_bits: int = 8,
               metric="sqeuclidean", id_offset: int = 0, kmeans_iters: int = 20,
               train_fraction: float = 0.5, seed: int = 0,
               stream: Optional[torch.cuda.Stream] = None) -> "NativeIndex":
        return cls._ivf(db, "b2vs_ivfpq_build", n_lists, pq_dim, pq_bits, metric, id_offset,
                        kmeans_iters, train_fraction, seed, stream)

    @classmethod
    def _ivf(cls, db, fn, n_lists, pq_dim, pq_bits, metric, id_offset, iters, frac, seed, stream):
        _require_cuda_matrix(db, "db")
        L = lib()
        out = ctypes.c_void_p()
        m = metric_code(metric)
        p = IvfParams(int(n_lists), int(iters), float(frac), int(pq_dim), int(pq_bits), int(seed))
        _check(getattr(L, fn)(db.device.index, m, dtype_code(db.dtype), db.shape[1], db.data_ptr(),
                              db.shape[0], int(id_offset), ctypes.byref(p),
                              _stream_ptr(db.device, stream), ctypes.byref(out)), fn)
        # IVF-PQ borrows the source rows for refine; IVF-Flat owns a copy of everything it needs
        return cls(out.value, db.device, m, keepalive=db if pq_dim else None)

    @classmethod
    def default_pq_dim(cls, dim: int) -> int:
        """The reference's default ``pq_dim = min(64, dim // 4)``
        (Attempt_1/index_building_coordinator.py:401), snapped to the nearest divisor of ``dim``
        that the engine supports (sub-vector length <= 16), preferring sub-vector lengths 2/4/8
        with ``pq_dim % 16 == 0`` - the shapes the grouped tensor-core scan serves."""
        want = max(1, min(64, dim // 4))
        cands = [m for m in range(1, min(dim, 200) + 1) if dim % m == 0 and dim // m <= 16]
        if not cands:
            raise ValueError(f"no supported pq_dim for dim={dim} (needs a divisor m <= 200 with dim/m <= 16)")
        fast = [m for m in cands if dim // m in (2, 4, 8) and m % 16 == 0 and dim % 64 == 0]
        pool = fast or cands
        return min(pool, key=lambda m: (abs(m - want), -m))

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """Write a trained IVF index to ``path`` (flat indexes hold no trained state)."""
        _check(lib().b2vs_index_save(self._h, os.fsencode(path)), "b2vs_index_save")

    @classmethod
    def load(cls, path: str, device, rows: Optional[torch.Tensor] = None, id_offset: int = -1,
             stream: Optional[torch.cuda.Stream] = None) -> "NativeIndex":
        """Load an IVF index saved by ``save`` onto ``device``; ``rows`` (the shard's original
        matrix on that device) is only needed for IVF-PQ searches with refine."""
        device = torch.device(device)
        if rows is not None:
            _require_cuda_matrix(rows, "rows")
        out = ctypes.c_void_p()
        _check(lib().b2vs_index_load(device.index or 0, os.fsencode(path),
                                     rows.data_ptr() if rows is not None else None, int(id_offset),
                                     _stream_ptr(device, stream), ctypes.byref(out)), "b2vs_index_load")
        inf = IndexInfo()
        _check(lib().b2vs_index_info_get(out, ctypes.byref(inf)), "b2vs_index_info_get")
        return cls(out.value, torch.device("cuda", inf.device), inf.metric, keepalive=rows)

    # ------------------------------------------------------------------ search
    def search(self, queries: torch.Tensor, k: int, n_probes: int = 0, refine_ratio: int = 0,
               n_splits: int = 0, stream: Optional[torch.cuda.Stream] = None,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, time_kernel: bool = False,
               graph: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident search: returns (distances f32 [Q,k], ids i64 [Q,k]) on this GPU.
        ``graph=True`` (IVF indexes, Q <= 64, k <= 128) replays the call as one CUDA graph from
        the second call of a (Q, k, dtype, n_probes, refine_ratio) signature on
        (``B2VS_FLAG_GRAPH``): the interactive single-query case is launch-bound."""
        if self._h.value is None:
            raise RuntimeError("index has been destroyed")
        _require_cuda_matrix(queries, "queries")
        if queries.device != self.device:
            raise ValueError(f"queries on {queries.device}, index on {self.device}")
        self._check_queries(queries)
        self._check_refine(refine_ratio)
        nq = queries.shape[0]
        if out is None:
            d = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            i = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        else:
            d, i = self._check_out(out, nq, k, self.device)
        sp = SearchParams(int(n_probes), int(refine_ratio), int(n_splits),
                          (FLAG_TIME_KERNEL if time_kernel else 0) | (FLAG_GRAPH if graph else 0))
        _check(lib().b2vs_search(self._h, queries.data_ptr(), dtype_code(queries.dtype), nq,
                                 queries.shape[1], int(k), ctypes.byref(sp), d.data_ptr(), i.data_ptr(),
                                 _stream_ptr(self.device, stream)), "b2vs_search")
        return d, i

    def search_host(self, queries: torch.Tensor, k: int, n_probes: int = 0, refine_ratio: int = 0,
                    n_splits: int = 0, stream: Optional[torch.cuda.Stream] = None,
                    out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host-buffer search (H2D + search + D2H inside the call); CPU tensors in and out."""
        if self._h.value is None:
            raise RuntimeError("index has been destroyed")
        if queries.is_cuda or queries.dim() != 2 or not queries.is_contiguous():
            raise ValueError("search_host expects a contiguous 2D CPU tensor")
        self._check_queries(queries)
        self._check_refine(refine_ratio)
        nq = queries.shape[0]
        if out is None:
            d = torch.empty((nq, k), dtype=torch.float32)
            i = torch.empty((nq, k), dtype=torch.int64)
        else:
            d, i = self._check_out(out, nq, k, torch.device("cpu"))
        sp = SearchParams(int(n_probes), int(refine_ratio), int(n_splits), 0)
        _check(lib().b2vs_search_host(self._h, queries.data_ptr(), dtype_code(queries.dtype), nq,
                                      queries.shape[1], int(k), ctypes.byref(sp), d.data_ptr(),
                                      i.data_ptr(), _stream_ptr(self.device, stream)), "b2vs_search_host")
        return d, i

    # ------------------------------------------------------------------ introspection
    def info(self) -> IndexInfo:
        inf = IndexInfo()
        _check(lib().b2vs_index_info_get(self._h, ctypes.byref(inf)), "b2vs_index_info_get")
        return inf

    def last_stats(self) -> SearchStats:
        st = SearchStats()
        _check(lib().b2vs_index_last_stats(self._h, ctypes.byref(st)), "b2vs_index_last_stats")
        return st

    def list_sizes(self) -> torch.Tensor:
        inf = self.info()
        out = torch.empty(max(inf.n_lists, 1), dtype=torch.int32)
        _check(lib().b2vs_ivf_list_sizes_host(self._h, out.data_ptr()), "b2vs_ivf_list_sizes_host")
        return out[: inf.n_lists]

    def centroids(self) -> torch.Tensor:
        inf = self.info()
        out = torch.empty((max(inf.n_lists, 1), inf.dim), dtype=torch.float32)
        _check(lib().b2vs_ivf_centroids_host(self._h, out.data_ptr()), "b2vs_ivf_centroids_host")
        return out[: inf.n_lists]

    @property
    def descending(self) -> bool:
        return self.metric == METRIC_IP   # cosine distances (1 - cos) come back ascending

    def destroy(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            lib().b2vs_index_destroy(self._h)
            self._h = ctypes.c_void_p()
            self._keepalive = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def merge_topk(d_all: torch.Tensor, i_all: torch.Tensor, k: int, descending: bool = False,
               stream: Optional[torch.cuda.Stream] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-k over per-shard results stacked as [n_parts, Q, k_in] on one GPU."""
    if d_all.dim() != 3 or d_all.shape != i_all.shape:
        raise ValueError("d_all / i_all must be [n_parts, Q, k_in] with equal shapes")
    if not d_all.is_cuda:
        raise ValueError("merge_topk runs on the GPU; move the stacked results to a CUDA device")
    d_all = d_all.contiguous().to(torch.float32)
    i_all = i_all.contiguous().to(torch.int64)
    g, nq, k_in = d_all.shape
    out_d = torch.empty((nq, k), dtype=torch.float32, device=d_all.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=d_all.device)
    _check(lib().b2vs_merge_topk(d_all.device.index, d_all.data_ptr(), i_all.data_ptr(), g, nq, k_in,
                                 int(k), 1 if descending else 0, out_d.data_ptr(), out_i.data_ptr(),
                                 _stream_ptr(d_all.device, stream)), "b2vs_merge_topk")
    return out_d, out_i


def kmeans_fit(x: torch.Tensor, n_clusters: int, iters: int = 20, seed: int = 0,
               stream: Optional[torch.cuda.Stream] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    _require_cuda_matrix(x, "x")
    cent = torch.empty((n_clusters, x.shape[1]), dtype=torch.float32, device=x.device)
    labels = torch.empty((x.shape[0],), dtype=torch.int32, device=x.device)
    _check(lib().b2vs_kmeans_fit(x.device.index, dtype_code(x.dtype), x.shape[1], x.data_ptr(),
                                 x.shape[0], int(n_clusters), int(iters), int(seed),
                                 cent.data_ptr(), labels.data_ptr(), _stream_ptr(x.device, stream)),
           "b2vs_kmeans_fit")
    return cent, labels


POOL_LAST_TOKEN, POOL_MEAN = 0, 1
_POOLING_CODE = {"last_token": POOL_LAST_TOKEN, "last": POOL_LAST_TOKEN, "mean": POOL_MEAN}


def pool_normalize(hidden: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                   pooling: str = "last_token", normalize: bool = True,
                   out_dtype: Optional[torch.dtype] = None,
                   stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Encoder hand-off (``b2vs_pool_normalize``): hidden states [B, T, D] on a GPU -> pooled,
    L2-normalised query rows [B, D] on the same GPU in ``out_dtype`` (default: the hidden states'
    dtype).  ``pooling="last_token"`` follows the reference's ``last_token_pool``
    (generate_embeddings.py:11-21), ``"mean"`` is the sentence-transformers masked mean."""
    if not isinstance(hidden, torch.Tensor) or hidden.dim() != 3:
        raise ValueError("hidden must be a 3D tensor [batch, seq_len, dim]")
    if not hidden.is_cuda:
        raise ValueError("pool_normalize runs on the GPU; hidden states must be CUDA tensors")
    try:
        mode = _POOLING_CODE[pooling]
    except KeyError:
        raise ValueError(f"unknown pooling {pooling!r}; expected 'last_token' or 'mean'")
    hidden = hidden.contiguous()
    b, t, d = hidden.shape
    if b == 0 or t == 0 or d == 0:
        raise ValueError(f"hidden has an empty dimension: {tuple(hidden.shape)}")
    mask_ptr = None
    if attention_mask is not None:
        if tuple(attention_mask.shape) != (b, t):
            raise ValueError(f"attention_mask shape {tuple(attention_mask.shape)} != ({b}, {t})")
        attention_mask = attention_mask.to(device=hidden.device, dtype=torch.int64).contiguous()
        mask_ptr = attention_mask.data_ptr()
    out = torch.empty((b, d), dtype=out_dtype or hidden.dtype, device=hidden.device)
    _check(lib().b2vs_pool_normalize(hidden.device.index, dtype_code(hidden.dtype), hidden.data_ptr(),
                                     b, t, d, mask_ptr, mode, 1 if normalize else 0,
                                     dtype_code(out.dtype), out.data_ptr(),
                                     _stream_ptr(hidden.device, stream)), "b2vs_pool_normalize")
    return out


def reload_env() -> None:
    """Re-read the B2VS_* switches (the library reads them once; tests flip them in-process)."""
    _check(lib().b2vs_reload_env(), "b2vs_reload_env")


def check_canaries() -> Tuple[int, int]:
    """(live device buffers checked, buffers with an overwritten guard zone); needs B2VS_CANARY=1."""
    n, bad = ctypes.c_int(0), ctypes.c_int(0)
    _check(lib().b2vs_debug_check_canaries(ctypes.byref(n), ctypes.byref(bad)), "b2vs_debug_check_canaries")
    return n.value, bad.value


def partition_even_native(n: int, n_parts: int, rank: int) -> Tuple[int, int]:
    b, e = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(lib().b2vs_partition_even(int(n), int(n_parts), int(rank), ctypes.byref(b), ctypes.byref(e)),
           "b2vs_partition_even")
    return b.value, e.value


class Comm:
    """One rank's ``b2vs_comm*``: the NCCL communicator + exchange collectives of the sharded search
    (include/b2vs.h, "Cross-shard exchange").  Created from a 128-byte id that rank 0 makes with
    ``Comm.unique_id()`` and ships to the other ranks out of band (``from_torch_distributed`` uses
    the already-initialised torch.distributed group for exactly that one broadcast)."""

    def __init__(self, handle: int, device: torch.device, n_ranks: int, rank: int):
        self._h = ctypes.c_void_p(handle)
        self.device = device
        self.n_ranks = n_ranks
        self.rank = rank

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
        _check(lib().b2vs_comm_unique_id(buf), "b2vs_comm_unique_id")
        return buf.raw

    @classmethod
    def init_rank(cls, device, n_ranks: int, rank: int, unique_id: bytes) -> "Comm":
        device = torch.device(device)
        if len(unique_id) != UNIQUE_ID_BYTES:
            raise ValueError(f"unique_id must be {UNIQUE_ID_BYTES} bytes")
        out = ctypes.c_void_p()
        _check(lib().b2vs_comm_init_rank(device.index or 0, int(n_ranks), int(rank),
                                         ctypes.c_char_p(unique_id), ctypes.byref(out)),
               "b2vs_comm_init_rank")
        return cls(out.value, device, int(n_ranks), int(rank))

    @classmethod
    def from_torch_distributed(cls, device) -> "Comm":
        """One process per GPU under torchrun: rank 0's id travels through the existing
        torch.distributed group (a 128-byte broadcast); everything after that is this library."""
        import torch.distributed as dist
        device = torch.device(device)
        rank, world = dist.get_rank(), dist.get_world_size()
        if rank == 0:
            raw = cls.unique_id()
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
        else:
            t = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.to(device)
        dist.broadcast(t, src=0)
        return cls.init_rank(device, world, rank, bytes(t.cpu().numpy().tobytes()))

    @classmethod
    def init_all(cls, devices) -> list:
        """Single process driving several GPUs (the reference's thread-per-GPU mode)."""
        devs = [torch.device(d).index or 0 for d in devices]
        arr = (ctypes.c_int * len(devs))(*devs)
        outs = (ctypes.c_void_p * len(devs))()
        _check(lib().b2vs_comm_init_all(len(devs), arr, outs), "b2vs_comm_init_all")
        return [cls(outs[i], torch.device("cuda", devs[i]), len(devs), i) for i in range(len(devs))]

    def query_slice(self, nq_total: int) -> Tuple[int, int]:
        return partition_even_native(nq_total, self.n_ranks, self.rank)

    def allgather_queries(self, q_local: torch.Tensor, nq_total: int,
                          stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        _require_cuda_matrix(q_local, "q_local")
        b, e = self.query_slice(nq_total)
        if q_local.shape[0] != e - b:
            raise ValueError(f"rank {self.rank} must pass its slice of {e - b} queries (got {q_local.shape[0]})")
        out = torch.empty((nq_total, q_local.shape[1]), dtype=q_local.dtype, device=self.device)
        _check(lib().b2vs_allgather_queries(self._h, q_local.data_ptr(), dtype_code(q_local.dtype),
                                            int(nq_total), q_local.shape[1], out.data_ptr(),
                                            _stream_ptr(self.device, stream)), "b2vs_allgather_queries")
        return out

    def allgather_topk(self, d: torch.Tensor, i: torch.Tensor,
                       stream: Optional[torch.cuda.Stream] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        nq, k = d.shape
        d_all = torch.empty((self.n_ranks, nq, k), dtype=torch.float32, device=self.device)
        i_all = torch.empty((self.n_ranks, nq, k), dtype=torch.int64, device=self.device)
        _check(lib().b2vs_allgather_topk(self._h, d.data_ptr(), i.data_ptr(), nq, k, d_all.data_ptr(),
                                         i_all.data_ptr(), _stream_ptr(self.device, stream)),
               "b2vs_allgather_topk")
        return d_all, i_all

    def allgather_merge_topk(self, d: torch.Tensor, i: torch.Tensor, k_out: int, descending: bool = False,
                             stream: Optional[torch.cuda.Stream] = None,
                             out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        nq, k = d.shape
        if out is None:
            out = (torch.empty((nq, k_out), dtype=torch.float32, device=self.device),
                   torch.empty((nq, k_out), dtype=torch.int64, device=self.device))
        _check(lib().b2vs_allgather_merge_topk(self._h, d.data_ptr(), i.data_ptr(), nq, k, int(k_out),
                                               1 if descending else 0, out[0].data_ptr(),
                                               out[1].data_ptr(), _stream_ptr(self.device, stream)),
               "b2vs_allgather_merge_topk")
        return out

    def exchange_merge_topk(self, d: torch.Tensor, i: torch.Tensor, k_out: int, descending: bool = False,
                            stream: Optional[torch.cuda.Stream] = None,
                            out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """All-to-all + merge: returns the global top-k of THIS rank's query slice."""
        nq, k = d.shape
        b, e = self.query_slice(nq)
        if out is None:
            out = (torch.empty((e - b, k_out), dtype=torch.float32, device=self.device),
                   torch.empty((e - b, k_out), dtype=torch.int64, device=self.device))
        _check(lib().b2vs_exchange_merge_topk(self._h, d.data_ptr(), i.data_ptr(), nq, k, int(k_out),
                                              1 if descending else 0, out[0].data_ptr(),
                                              out[1].data_ptr(), _stream_ptr(self.device, stream)),
               "b2vs_exchange_merge_topk")
        return out

    def allreduce_min(self, values: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        if values.dtype != torch.float32 or not values.is_cuda or not values.is_contiguous():
            raise ValueError("allreduce_min expects a contiguous float32 CUDA tensor")
        _check(lib().b2vs_allreduce_min_f32(self._h, values.data_ptr(), values.numel(),
                                            _stream_ptr(self.device, stream)), "b2vs_allreduce_min_f32")
        return values

    def register_index(self, index: "NativeIndex", stream: Optional[torch.cuda.Stream] = None) -> None:
        """Collective (every rank, same order): agree on the smallest shard so sharded flat searches
        exchange thresholds between their passes.  Done once per index object; ranks create their
        index objects in lockstep (SPMD), so the first sharded search of a new index registers it
        on every rank at the same point."""
        if getattr(index, "_sharded_comm", None) is self:
            return
        _check(lib().b2vs_comm_register_index(self._h, index._h, _stream_ptr(self.device, stream)),
               "b2vs_comm_register_index")
        index._sharded_comm = self

    def search_sharded(self, index: "NativeIndex", q_local: torch.Tensor, nq_total: int, k: int,
                       n_probes: int = 0, refine_ratio: int = 0, time_kernel: bool = False,
                       stream: Optional[torch.cuda.Stream] = None,
                       out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """The whole sharded step (``b2vs_search_sharded`` / ``_host``): ``q_local`` is this rank's
        slice of the batch - a CUDA tensor (device variant) or a CPU tensor (host variant: H2D of
        the slice and D2H of its answer inside the call).  Returns the slice's global top-k."""
        b, e = self.query_slice(nq_total)
        if q_local.dim() != 2 or q_local.shape[0] != e - b or not q_local.is_contiguous():
            raise ValueError(f"rank {self.rank} must pass its contiguous [{e - b}, dim] query slice")
        self.register_index(index, stream)
        index._check_queries(q_local)
        index._check_refine(refine_ratio)
        host = not q_local.is_cuda
        dev = torch.device("cpu") if host else self.device
        if out is None:
            out = (torch.empty((e - b, k), dtype=torch.float32, device=dev),
                   torch.empty((e - b, k), dtype=torch.int64, device=dev))
        else:
            out = index._check_out(out, e - b, k, dev)
        sp = SearchParams(int(n_probes), int(refine_ratio), 0, FLAG_TIME_KERNEL if time_kernel else 0)
        fn = lib().b2vs_search_sharded_host if host else lib().b2vs_search_sharded
        _check(fn(self._h, index._h, q_local.data_ptr(), dtype_code(q_local.dtype), int(nq_total),
                  q_local.shape[1], int(k), ctypes.byref(sp), out[0].data_ptr(), out[1].data_ptr(),
                  _stream_ptr(self.device, stream)), fn.__name__)
        return out

    def destroy(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            lib().b2vs_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def device_count() -> int:
    c = ctypes.c_int(0)
    _check(lib().b2vs_device_count(ctypes.byref(c)), "b2vs_device_count")
    return c.value
