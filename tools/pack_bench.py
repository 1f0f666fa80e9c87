"""HBM-bound layout producers: packed-stream packing (fp32 / bf16 in, with and without the fused l2-normalise)
and the masked l2-normalise forward / backward, as achieved GB/s of the bytes they must move."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import evdr_b200 as evdr

dev = torch.device("cuda", 0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for npg, lp in ((20000, 1030), (300000, 64), (1000, 41)):
    g = torch.Generator(device=dev).manual_seed(1)
    P = torch.randn(npg, lp, 128, generator=g, device=dev)
    pm = torch.rand(npg, lp, generator=g, device=dev) > 0.05
    slots = ((lp + 15) // 16) * 16
    for name, src, norm, kb in (("fp32 -> bf16", P, False, 1), ("fp32 -> bf16 + l2norm", P, True, 1),
                                ("bf16 -> bf16", P.bfloat16(), False, 1), ("fp32 -> bf16x3", P, False, 3)):
        ms = timed(lambda: evdr.PackedPages.from_padded(src, pm, normalize=norm, varlen=False, k_blocks=kb))
        moved = npg * lp * 128 * src.element_size() + npg * slots * 128 * 2 * kb + npg * lp
        print(f"pack {npg} x {lp:4d}  {name:22s}: {ms:8.3f} ms  {moved / ms / 1e6:8.1f} GB/s")
    Pg = P.clone().requires_grad_(True)
    ms_f = timed(lambda: evdr.l2_normalize_masked(P, pm))
    y = evdr.l2_normalize_masked(Pg, pm)
    gy = torch.randn_like(y)
    ms_b = timed(lambda: torch.autograd.grad(y, Pg, gy, retain_graph=True))
    n = npg * lp * 128 * 4
    print(f"l2norm {npg} x {lp:4d}: fwd {ms_f:8.3f} ms {2 * n / ms_f / 1e6:8.1f} GB/s   bwd {ms_b:8.3f} ms {3 * n / ms_b / 1e6:8.1f} GB/s")
