"""CPU: the staged-reference loader (baseline/ref.py + tools/stage_reference.py) and small host-side rules.

The reference only exists in the build container (/root/reference); its files travel to the GPU box as a
byte-identical staged copy under the git-ignored baseline/_ref.  These tests run wherever that copy exists and
skip (with the reason) elsewhere."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from baseline import ref as reference

ROOT = Path(__file__).resolve().parent.parent
needs_ref = pytest.mark.skipif(not reference.available(), reason="baseline/_ref not staged (run tools/stage_reference.py)")


@needs_ref
def test_staged_files_match_manifest_and_source():
    man = reference.verify()
    assert "evaluator/retrieval.py" in man["files"] and "mainv2_iter_liscore.py" in man["files"]
    src = Path(man["source"])
    if src.is_dir():                      # build container: the staged bytes ARE the reference's bytes
        for rel, digest in man["files"].items():
            assert hashlib.sha256((src / rel).read_bytes()).hexdigest() == digest, rel


@needs_ref
def test_tampered_staging_is_refused(tmp_path, monkeypatch):
    import shutil
    fake = tmp_path / "_ref"
    shutil.copytree(reference.REF, fake)
    (fake / "criterion.py").write_text((fake / "criterion.py").read_text() + "\n# edited\n")
    monkeypatch.setattr(reference, "REF", fake)
    monkeypatch.setattr(reference, "_verified", False)
    with pytest.raises(RuntimeError, match="UNMODIFIED"):
        reference.verify()


@needs_ref
def test_reference_scorer_runs_and_matches_the_oracle():
    from oracle import maxsim_oracle as mo
    fn = reference.scorer_module().score_multi_vector_masked
    assert fn.__code__.co_filename.replace("\\", "/").endswith("baseline/_ref/evaluator/retrieval.py")
    g = torch.Generator().manual_seed(0)
    Q, P = torch.randn(3, 5, 16, generator=g), torch.randn(4, 9, 16, generator=g)
    qm, pm = torch.rand(3, 5, generator=g) > 0.2, torch.rand(4, 9, generator=g) > 0.3
    pm[2] = False                                                     # empty page -> 0
    out = fn(Q, P, qm, pm, 3)
    np.testing.assert_allclose(out.numpy(), mo.maxsim_masked(Q.numpy(), P.numpy(), qm.numpy(), pm.numpy()), rtol=1e-5, atol=1e-5)
    assert (out[:, 2] == 0).all()


@needs_ref
def test_script_binds_stock_or_dropin_by_path_order():
    stock = reference.load_script("mainv2_iter_liscore.py", dropin=False)
    ours = reference.load_script("mainv2_iter_liscore.py", dropin=True)
    assert stock.score_multi_vector_masked.__module__ == "evaluator.retrieval"
    assert stock.listwise_distillation_loss.__code__.co_filename.replace("\\", "/").endswith("baseline/_ref/criterion.py")
    assert "efficient-visual-document-retrieval_b200" in ours.score_multi_vector_masked.__module__
    assert "efficient-visual-document-retrieval_b200" in ours.score_preserving_loss.__module__
    # same unmodified script source in both loads
    assert stock.train_one_step.__code__.co_code == ours.train_one_step.__code__.co_code
    import sys
    assert "evaluator.retrieval" not in sys.modules or "baseline/_ref" not in str(getattr(sys.modules["evaluator.retrieval"], "__file__", ""))


def test_default_precision_rule(evdr, monkeypatch):
    monkeypatch.delenv("EVDR_PRECISION", raising=False)
    assert evdr.retrieval.default_precision(torch.zeros(1, 1, 128)) == "bf16x3"
    assert evdr.retrieval.default_precision(torch.zeros(1, 1, 128, dtype=torch.bfloat16)) == "bf16"
    monkeypatch.setenv("EVDR_PRECISION", "fp32")
    assert evdr.retrieval.default_precision(torch.zeros(1, 1, 128, dtype=torch.bfloat16)) == "fp32"
    assert evdr._lib.row_blocks(1) == 1 and evdr._lib.row_blocks(3) == 2


def test_traffic_json_points_at_committed_captures():
    t = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    for key in ("1024x100000", "64x100000", "1x100000"):
        assert t[key]["bytes"] > 26_368_000_000 and (ROOT / t[key]["source"]).is_file()
        assert t[key]["bytes"] < 1.1 * 26_368_000_000 + 1024 * 100_000 * 4 * 1.1, "traffic well above the algorithmic bytes"
