"""CPU: the C-ABI library loads without a GPU, exports every symbol include/evdr.h declares, and the
product path fails loudly (no CPU fallback, never the oracle)."""
import ast
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "efficient-visual-document-retrieval_b200"


def _declared_symbols():
    text = (ROOT / "include" / "evdr.h").read_text()
    return sorted(set(re.findall(r"EVDR_API\s+[\w\s\*]+?\b(evdr_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(evdr):
    names = _declared_symbols()
    assert len(names) >= 20
    lib = C.CDLL(str(evdr._lib.lib_path()))
    for n in names:
        assert hasattr(lib, n), f"libevdr.so does not export {n}"
    assert sorted(evdr._lib.EXPORTED_SYMBOLS) == names, "ctypes table and header disagree"
    assert evdr._lib.load().evdr_version() >= 100


def test_no_link_time_gpu_dependencies(evdr):
    import subprocess
    out = subprocess.run(["ldd", str(evdr._lib.lib_path())], capture_output=True, text=True).stdout
    for forbidden in ("libcuda.so", "libnccl", "libcudart", "libtorch"):
        assert forbidden not in out, out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(evdr):
    lib = evdr._lib.load()
    h = C.c_void_p()
    rc = lib.evdr_create(0, C.byref(h))
    assert rc == -2 and "no CPU fallback" in evdr._lib.last_error()
    Q = torch.randn(2, 4, 128)
    m = torch.ones(2, 4, dtype=torch.bool)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evdr.score_multi_vector_masked(Q, Q, m, m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evdr.PackedPages.from_padded(Q, m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evdr.top_k(torch.randn(2, 5), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evdr.l2_normalize_masked(Q, m)


def test_argument_validation_without_gpu(evdr):
    Q = torch.randn(2, 4, 128)
    with pytest.raises(ValueError):
        evdr.score_multi_vector_masked(Q, torch.randn(3, 5, 64), torch.ones(2, 4), torch.ones(3, 5))
    with pytest.raises(ValueError):
        evdr.score_multi_vector_masked(Q, Q, torch.ones(2, 3), torch.ones(2, 4))
    with pytest.raises(ValueError, match="No queries provided"):
        evdr.BaseVisualRetrieverProcessor.score_multi_vector([], [Q[0]], device="cuda:0")
    with pytest.raises(ValueError, match="No passages provided"):
        evdr.BaseVisualRetrieverProcessor.score_multi_vector([Q[0]], [], device="cuda:0")


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may touch oracle/."""
    for py in PKG.rglob("*.py"):
        tree = ast.parse(py.read_text())
        for node in ast.walk(tree):
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            else:
                continue
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), f"{py} imports the oracle"


def test_kernels_are_native_blackwell(evdr):
    """SASS evidence: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(evdr._lib.lib_path())], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "FMNMX3"):
        assert mnemonic in sass, f"{mnemonic} missing from libevdr.so"
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", str(evdr._lib.lib_path())], capture_output=True, text=True).stdout
