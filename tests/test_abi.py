"""The C-ABI library builds, loads and exports every symbol include/b2vs.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2vs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2vs_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points(b2):
    syms = declared_symbols()
    assert set(syms) == set(b2._native.EXPORTS)


def test_library_builds_and_exports_all_symbols(b2):
    path = b2.build_native()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} missing from {path}"


def test_version_and_error_string(b2):
    lib = b2._native.lib()
    assert lib.b2vs_version() == 200
    assert isinstance(lib.b2vs_last_error(), bytes)


def test_struct_layouts_match_header(b2):
    n = b2._native
    assert ctypes.sizeof(n.IvfParams) == 32
    assert ctypes.sizeof(n.SearchParams) == 16
    assert ctypes.sizeof(n.IndexInfo) == 56
    assert ctypes.sizeof(n.SearchStats) == 48


def test_null_arguments_are_rejected_without_touching_a_gpu(b2):
    lib = b2._native.lib()
    assert lib.b2vs_index_info_get(None, None) == -1
    assert lib.b2vs_search(None, None, 0, 1, 8, 1, None, None, None, None) == -1
    assert b"NULL" in lib.b2vs_last_error()
    assert lib.b2vs_index_destroy(None) == 0
    assert lib.b2vs_merge_topk(0, None, None, 1, 1, 1, 1, 0, None, None, None) == -1


def test_comm_entry_points_reject_bad_arguments_without_a_gpu(b2):
    """The exchange ABI (b2vs_comm_* / b2vs_allgather_topk) validates before touching NCCL."""
    lib = b2._native.lib()
    assert lib.b2vs_comm_init_rank(0, 2, 5, None, None) == -1
    assert lib.b2vs_comm_info(None, None, None, None) == -1
    assert lib.b2vs_comm_destroy(None) == 0
    assert lib.b2vs_allgather_topk(None, None, None, 1, 1, None, None, None) == -1
    assert lib.b2vs_exchange_merge_topk(None, None, None, 1, 1, 1, 0, None, None, None) == -1
    assert lib.b2vs_search_sharded(None, None, None, 0, 1, 8, 1, None, None, None, None) == -1
    assert lib.b2vs_reload_env() == 0


def test_partition_even_matches_the_reference_strategy(b2):
    """b2vs_partition_even == GPUResourceManager.distribute_workload('even')
    (gpu_resource_manager.py:190-202): 300 / 301 items over 3 parts, remainder to the first."""
    pe = b2._native.partition_even_native
    assert [pe(300, 3, r) for r in range(3)] == [(0, 100), (100, 200), (200, 300)]
    assert [pe(301, 3, r) for r in range(3)] == [(0, 101), (101, 201), (201, 301)]
    for n in (1, 7, 10_000, 10_000_019):
        for parts in (1, 2, 3, 8):
            assert [pe(n, parts, r) for r in range(parts)] == list(b2.partition_even(n, parts))


def test_default_pq_dim_follows_the_reference_and_is_buildable(b2):
    """Reference default min(64, dim // 4) (index_building_coordinator.py:401), snapped to a
    supported divisor: every headline dim gives a pq_dim that ivf_build accepts."""
    f = b2.NativeIndex.default_pq_dim
    for dim, want in ((768, 96), (1024, 128), (128, 32), (384, 48), (96, 24), (100, 25), (30, 6)):
        m = f(dim)
        assert dim % m == 0 and dim // m <= 16 and m <= 200, (dim, m)
        assert m == want, (dim, m)


def test_missing_library_fails_loudly(b2, monkeypatch):
    n = b2._native
    monkeypatch.setattr(n, "_lib", None)
    monkeypatch.setattr(n, "LIB_PATH", "/nonexistent/libb2vs.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        n.lib()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cuvs-rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
