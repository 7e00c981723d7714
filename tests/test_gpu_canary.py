"""Own bounds checking (compute-sanitizer is closed on the GPU pool): with B2VS_CANARY=1 every
device buffer the library allocates sits between two 256-byte guard zones; a worker process drives
every search path (exact: fused / large-k / fp32 split; IVF-Flat and IVF-PQ: one-CTA planner, both
seed modes, counting-sort planner, large-k, all three decoder widths; sharded step; cosine; merge)
and b2vs_debug_check_canaries() then reads all guards back."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_no_kernel_writes_outside_the_library_buffers():
    env = dict(os.environ, B2VS_CANARY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_canary_worker.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    print(r.stdout)     # the per-stage guard-zone reports (pytest -s shows them; kept in profiles/)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "CANARY_OK" in r.stdout and "corrupt" in r.stdout
    assert ", 0 corrupt" in r.stdout.splitlines()[-2]


def test_guard_zones_are_off_by_default(b2):
    import ctypes
    n, bad = ctypes.c_int(0), ctypes.c_int(0)
    rc = b2._native.lib().b2vs_debug_check_canaries(ctypes.byref(n), ctypes.byref(bad))
    assert rc == -4 and b"B2VS_CANARY" in b2._native.lib().b2vs_last_error()
