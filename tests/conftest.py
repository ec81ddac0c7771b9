import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """libb200olap.so built in-tree (nvcc cross-compiles without a GPU)."""
    from dpu_olap_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctx(built_lib):
    from dpu_olap_b200.ops import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    import json
    g = {}
    for name in ("generator_golden", "arrow_golden"):
        g[name] = json.loads((ROOT / "tests" / "golden" / f"{name}.json").read_text())
    return g
