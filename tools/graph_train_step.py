"""Config-5 training step captured in a CUDA graph (torch.cuda.graph around the drop-in calls): every
launch of libevdr.so goes to the capturing stream, scratch comes from cudaMallocAsync (graph memory nodes),
TMA descriptors are encoded at capture time against static tensors.  Compares losses and the trained student
after N steps with the eager loop and times both.   python tools/graph_train_step.py [Ls] [steps]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

dev = torch.device("cuda", 0)
B, npg = 32, 1000
ls = int(sys.argv[1]) if len(sys.argv) > 1 else 41
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
Pt, pmt = evdr.synth.make_pages(npg, 1030, seed=1, device=dev, empty_frac=0.0)
Qall, qmall, _ = evdr.synth.make_queries(Pt, pmt, 512, 32, seed=2, sigma=3.0)
Qall = Qall.float()
Ptf = Pt.float()
init = torch.randn(npg, ls, 128, device=dev)
pms = torch.ones(npg, ls, dtype=torch.bool, device=dev)


def make():
    p = init.clone().requires_grad_(True)
    return p, torch.optim.AdamW([p], lr=1e-3, capturable=True)


def step_fn(param, opt, Qb, qmb):
    Psb = evdr.l2_normalize_masked(param, pms)
    with torch.no_grad():
        sc_t = evdr.score_multi_vector_masked(Qb, Ptf, qmb, pmt, 64)
    sc_s = evdr.score_multi_vector_masked(Qb, Psb, qmb, pms, 64)
    ll, lm = evdr.liscore_losses(sc_s, sc_t, k=10, temperature=1.0)
    loss = ll + lm
    opt.zero_grad(set_to_none=False)
    loss.backward()
    opt.step()
    return loss.detach()


def batches():
    for i in range(steps):
        j = (i * B) % 512
        yield Qall[j:j + B], qmall[j:j + B]


# ---- eager
param_e, opt_e = make()
for Qb, qmb in list(batches())[:5]:
    step_fn(param_e, opt_e, Qb, qmb)                      # warm-up (allocator, pack cache)
param_e, opt_e = make()
torch.cuda.synchronize()
t0 = time.perf_counter()
losses_e = [step_fn(param_e, opt_e, Qb, qmb) for Qb, qmb in batches()]
torch.cuda.synchronize()
t_eager = (time.perf_counter() - t0) / steps * 1e3

# ---- graph: static input buffers, one capture, replay per step
param_g, opt_g = make()
Qs, qms = Qall[:B].clone(), qmall[:B].clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step_fn(param_g, opt_g, Qs, qms)
torch.cuda.current_stream().wait_stream(s)
param_g.data.copy_(init)
opt_g = torch.optim.AdamW([param_g], lr=1e-3, capturable=True)
with torch.cuda.stream(s):                                  # AdamW state must exist before capture
    step_fn(param_g, opt_g, Qs, qms)
torch.cuda.current_stream().wait_stream(s)
param_g.data.copy_(init)
for st in opt_g.state.values():
    for v in st.values():
        if torch.is_tensor(v):
            v.zero_()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    loss_static = step_fn(param_g, opt_g, Qs, qms)
param_g.data.copy_(init)
for st in opt_g.state.values():
    for v in st.values():
        if torch.is_tensor(v):
            v.zero_()
torch.cuda.synchronize()
t0 = time.perf_counter()
losses_g = []
for Qb, qmb in batches():
    Qs.copy_(Qb)
    qms.copy_(qmb)
    graph.replay()
    losses_g.append(loss_static.clone())
torch.cuda.synchronize()
t_graph = (time.perf_counter() - t0) / steps * 1e3

le = torch.stack(losses_e).cpu()
lg = torch.stack(losses_g).cpu()
print(f"Ls={ls} steps={steps}: eager {t_eager:.3f} ms/step, CUDA graph {t_graph:.3f} ms/step")
print(f"loss first/last eager {le[0]:.6f}/{le[-1]:.6f} graph {lg[0]:.6f}/{lg[-1]:.6f}  max |dloss| {float((le - lg).abs().max()):.3e}")
print(f"student after {steps} steps: max |eager - graph| = {float((param_e - param_g).abs().max()):.3e}")
