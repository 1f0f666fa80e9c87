"""A few drop-in calls on fp32 (not bf16-representable) tensors -> the bf16x3 split flavours, for ncu captures.
    python tools/x3_launch.py [nq] [pages] [tokens]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1663
npg = int(sys.argv[2]) if len(sys.argv) > 2 else 277
lp = int(sys.argv[3]) if len(sys.argv) > 3 else 1030
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
P = torch.nn.functional.normalize(torch.randn(npg, lp, 128, generator=g, device=dev), dim=-1)
Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g, device=dev), dim=-1)
pm = torch.rand(npg, lp, generator=g, device=dev) > 0.02
qm = torch.ones(nq, 32, dtype=torch.bool, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
evdr.score_multi_vector_masked(Q, P, qm, pm, 64)
ev[0].record()
for i in range(4):
    out = evdr.score_multi_vector_masked(Q, P, qm, pm, 64)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
exact = evdr.score_multi_vector_masked(Q[:64], P, qm[:64], pm, precision="fp32")
err = ((out[:64] - exact).abs() / exact.abs().clamp_min(1)).max().item()
print(evdr._lib.last_kernel())
print("ms per call:", [round(m, 3) for m in ms], "TFLOP/s (x3 products):",
      [round(3 * nq * npg * 2 * 128 * 32 * lp / (m * 1e-3) / 1e12, 1) for m in ms], f"max rel err vs fp32 {err:.1e}")
