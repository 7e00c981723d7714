import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG_DIR = os.path.join(ROOT, "cuvs-rag_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def b2():
    import cuvs_rag_b200
    return cuvs_rag_b200


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _reload_b2vs_env_switches(request, monkeypatch):
    """The library reads its B2VS_* switches once; tests that flip them through ``setenv`` (below)
    get a re-read, and the defaults are restored when the test ends."""
    yield
    if any(k.startswith("B2VS_") for k in os.environ) or getattr(request.node, "_b2vs_env_touched", False):
        monkeypatch.undo()
        try:
            import cuvs_rag_b200
            cuvs_rag_b200._native.reload_env()
        except Exception:
            pass


@pytest.fixture
def setenv(request, monkeypatch):
    """setenv(name, value | None): set / delete a B2VS_* switch and make the library re-read them."""
    def _set(name, value):
        import cuvs_rag_b200
        request.node._b2vs_env_touched = True
        if value is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, value)
        cuvs_rag_b200._native.reload_env()
    return _set
