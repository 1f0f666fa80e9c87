"""Generate golden vectors by EXECUTING the unmodified reference (run in the build container only;
/root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

What runs: ``evaluator/retrieval.py::score_multi_vector_masked`` (+ torch autograd through it) and
``BaseVisualRetrieverProcessor.score_multi_vector`` imported from /root/reference, with
* empty ``mteb`` modules pre-registered in sys.modules (the only import that fails, :218), and
* for the list scorer only, ``device='cuda'`` requests redirected to the CPU (``left_padding``
  hard-codes 'cuda', :39,41) by wrapping ``torch.full`` / ``Tensor.to`` -- the reference source is
  not modified.
Inputs come from the package's seeded generator (synth.py); each fixture stores the generator
arguments, a sha256 of the generated inputs (to detect generator drift) and the reference outputs;
small cases store the inputs themselves.
"""
from __future__ import annotations

import hashlib
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")


def load_reference():
    for name in ["mteb", "mteb.evaluation", "mteb.evaluation.evaluators",
                 "mteb.evaluation.evaluators.RetrievalEvaluator"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mteb.evaluation.evaluators.RetrievalEvaluator"].RetrievalEvaluator = object
    sys.path.insert(0, str(REF))
    mod = importlib.import_module("evaluator.retrieval")
    assert Path(mod.__file__).resolve().is_relative_to(REF), mod.__file__
    return mod


def sha(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        a = t.detach().cpu()
        if a.dtype == torch.bfloat16:
            a = a.view(torch.int16)
        h.update(np.ascontiguousarray(a.numpy()).tobytes())
    return h.hexdigest()


def main():
    sys.path.insert(0, str(ROOT))
    synth = importlib.import_module("efficient-visual-document-retrieval_b200.synth")
    ref = load_reference()
    score = ref.score_multi_vector_masked
    torch.set_num_threads(8)

    # ------------------------------------------------------------ micro cases
    g = torch.Generator().manual_seed(7)
    Q = torch.randn(3, 5, 8, generator=g)
    P = torch.randn(4, 7, 8, generator=g)
    qm = torch.rand(3, 5, generator=g) > 0.2
    pm = torch.rand(4, 7, generator=g) > 0.3
    pm[2] = False          # page with every token masked -> 0
    qm[1] = False          # query with every token masked -> 0
    P[3, 4] = P[3, 1]      # two page tokens tie
    pm[3, 1] = pm[3, 4] = True
    micro = {"Q": Q.numpy(), "P": P.numpy(), "qmask": qm.numpy(), "pmask": pm.numpy()}
    micro["out"] = score(Q, P, qm, pm, 2).numpy()
    for cp in (1, 3, 128):
        assert torch.equal(score(Q, P, qm, pm, cp), torch.from_numpy(micro["out"])), "chunk_p must not matter"
    micro["out_bf16_in"] = score(Q.bfloat16(), P.bfloat16(), qm, pm).numpy()
    micro["out_int_masks"] = score(Q, P, qm.long(), pm.int()).numpy()
    Pg = P.clone().requires_grad_(True)
    Qg = Q.clone().requires_grad_(True)
    gout = torch.randn(3, 4, generator=g)
    score(Qg, Pg, qm, pm, 3).backward(gout)
    micro["gout"] = gout.numpy()
    micro["dP"] = Pg.grad.numpy()
    micro["dQ"] = Qg.grad.numpy()
    np.savez_compressed(HERE / "micro.npz", **micro)

    # ------------------------------------------------------------ stored-input medium case (d=128)
    Pm, pmm = synth.make_pages(12, 200, seed=11, n_holes=5, empty_frac=0.1, min_len=60)
    Qm, qmm, relm = synth.make_queries(Pm, pmm, 9, 32, seed=12, sigma=2.0, min_len=8, drop=0.1)
    out = score(Qm, Pm, qmm, pmm, 5)
    Pg = Pm.float().clone().requires_grad_(True)
    Qg = Qm.float().clone().requires_grad_(True)
    gout = torch.randn(9, 12, generator=g)
    score(Qg, Pg, qmm, pmm, 7).backward(gout)
    np.savez_compressed(
        HERE / "medium.npz",
        Q_bf16=Qm.view(torch.int16).numpy(), P_bf16=Pm.view(torch.int16).numpy(),
        qmask=qmm.numpy(), pmask=pmm.numpy(), rel=relm.numpy(), out=out.numpy(),
        gout=gout.numpy(), dP=Pg.grad.numpy(), dQ=Qg.grad.numpy())

    # ------------------------------------------------------------ seeded cases (inputs regenerated)
    seeded = {}
    specs = {
        # name: (n_pages, lp, nq, lq, page kwargs, query kwargs)
        "full1030": (40, 1030, 24, 32, dict(seed=21, n_holes=6, empty_frac=0.05), dict(seed=22, sigma=3.0, min_len=8, drop=0.1)),
        "comp64": (64, 64, 20, 32, dict(seed=31, n_holes=3, empty_frac=0.03, min_len=20), dict(seed=32, sigma=2.5, min_len=8, drop=0.1)),
        "varlen": (50, 333, 10, 27, dict(seed=41, n_holes=4, empty_frac=0.04, min_len=1), dict(seed=42, sigma=2.5, min_len=1, drop=0.2)),
        "lq1": (30, 103, 17, 1, dict(seed=51, n_holes=2, empty_frac=0.0, min_len=40), dict(seed=52, sigma=1.0)),
        "student_mf5": (25, 206, 32, 32, dict(seed=61, n_holes=4, empty_frac=0.04, min_len=100), dict(seed=62, sigma=3.0, min_len=10, drop=0.05)),
    }
    for name, (npg, lp, nq, lq, pk, qk) in specs.items():
        Pp, pmp = synth.make_pages(npg, lp, **pk)
        Qq, qmq, rel = synth.make_queries(Pp, pmp, nq, lq, **qk)
        seeded[f"{name}/out"] = score(Qq, Pp, qmq, pmp, 64).numpy()
        seeded[f"{name}/sha"] = np.array(sha(Qq, Pp, qmq, pmp))
        seeded[f"{name}/rel"] = rel.numpy()
        seeded[f"{name}/spec"] = np.array(repr((npg, lp, nq, lq, pk, qk)))
    # student-size backward (fp32 leaf, like Psb in train_one_step)
    npg, lp, nq, lq, pk, qk = specs["student_mf5"]
    Pp, pmp = synth.make_pages(npg, lp, **pk)
    Qq, qmq, _ = synth.make_queries(Pp, pmp, nq, lq, **qk)
    Pg = Pp.float().clone().requires_grad_(True)
    gout = torch.randn(nq, npg, generator=g)
    score(Qq, Pg, qmq, pmp, 32).backward(gout)
    seeded["student_mf5/gout"] = gout.numpy()
    seeded["student_mf5/dP"] = Pg.grad.numpy()
    np.savez_compressed(HERE / "seeded.npz", **seeded)

    # ------------------------------------------------------------ list scorer (A2), cuda -> cpu shim
    real_full, real_to = torch.full, torch.Tensor.to

    def full_shim(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return real_full(*a, **k)

    def to_shim(self, *a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return real_to(self, *a, **k)

    torch.full, torch.Tensor.to = full_shim, to_shim
    try:
        gl = torch.Generator().manual_seed(77)
        q_lens = [5, 32, 17, 9, 21]
        p_lens = [40, 64, 64, 13, 50, 7, 64]
        qs = [torch.nn.functional.normalize(torch.randn(l, 128, generator=gl), dim=-1).bfloat16().float() for l in q_lens]
        ps = [torch.nn.functional.normalize(torch.randn(l, 128, generator=gl), dim=-1).bfloat16().float() for l in p_lens]
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            out_b3 = ref.BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=3, device="cpu")
            out_b128 = ref.BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=128, device="cpu")
    finally:
        torch.full, torch.Tensor.to = real_full, real_to
    np.savez_compressed(
        HERE / "list_scorer.npz",
        q_lens=np.array(q_lens), p_lens=np.array(p_lens),
        qs=np.concatenate([q.numpy() for q in qs]), ps=np.concatenate([p.numpy() for p in ps]),
        out_b3=out_b3.numpy(), out_b128=out_b128.numpy())
    print("golden written:", sorted(p.name for p in HERE.glob("*.npz")))




if __name__ == "__main__":
    main()
