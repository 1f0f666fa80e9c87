import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


def _ensure_built():
    """The library is a build product (git-ignored): build it in-tree when a fresh checkout runs the tests before
    __graft_entry__.build() (nvcc cross-compiles without a GPU; a no-op when the stamp matches)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "evdr_build", ROOT / "efficient-visual-document-retrieval_b200" / "csrc" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not mod.OUT.exists():
        mod.build()


@pytest.fixture(scope="session")
def evdr():
    _ensure_built()
    import evdr_b200
    return evdr_b200


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture
def plain_bf16(monkeypatch):
    """Pin calls that do not name a precision to the plain bf16 tcgen05 flavours (the default for fp32 page
    tensors is the bf16x3 split, retrieval.default_precision); explicit precision= arguments still win."""
    monkeypatch.setenv("EVDR_PRECISION", "bf16")
