#!/usr/bin/env python
"""bench.py -- MaxSim query-page pairs/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: 1024 queries (32 tokens, 128-d) scored against
this rank's 100 000-page shard (1030 valid tokens per page, stored as 1040 bf16 slots) with the
tcgen05 MaxSim kernel, followed by the local top-10 and -- for N > 1 -- the exchange + merge of the
candidates (ncclAllGather + merge is the headline; the one-kernel peer-memory exchange over NVLink is timed
beside it, --exchange selects).  BASELINE.json configs[2] per GPU; the corpus grows with N: weak scaling.

value     : whole-job pairs/s with everything resident in HBM (CUDA events, max over ranks)
e2e       : same metric through the reference-facing call score_multi_vector_masked(Q, P, qmask,
            pmask) with the query batch copied from pinned host memory and the (nq, np) score matrix
            read back to the host inside the timed region, every step
roofline  : dominant kernel (maxsim_tc_kernel) timed with CUDA events INSIDE the timed steps: algorithmic
            FLOPs / duration vs the measured bf16 peak of MEASURED_PEAKS.json
cpu_baseline / --impl reference : the reference's OWN score_multi_vector_masked (unmodified file staged in
            baseline/_ref by tools/stage_reference.py; kind "reference") on the box's host cores, all
            threads, on a fixed bounded sample of the workload (config.reference_arm_sample)
gpu_reference : the same unmodified function, torch eager on this B200 (N = 1)
sharded_check (N > 1) : merged top-10 identical on every rank and equal to torch.topk of the all-gathered rows
extras    : single-query streaming scan (HBM roofline), 16 / 32 / 64-query batches (each back to back AND
            duty-cycled, with its own nvidia-smi clock / power record), corpus in pinned host memory,
            top-10 QPS on a 1 M-page corpus sharded over the box (N >= 2; 500 k pages at N = 1)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D = 128
LQ = 32
LP_VALID = 1030          # ColPali tokens per page
LP_SLOTS = 1040          # 65 chunks of 16
TOPK = 10
FLOPS_PER_PAIR = 2 * D * LQ * LP_VALID          # 8 437 760 (padding not counted)
BYTES_PER_PAGE = LP_VALID * D * 2               # 263 680
_JSON_FD = None


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


METRIC = "maxsim_query_page_pairs_per_sec"
UNIT = "pairs/s"


def measured_traffic(nq: int, pages: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this very
    configuration (profiles/traffic.json); None for configurations that were not captured."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text()).get(f"{nq}x{pages}", {}).get("bytes")
    except Exception:
        return None


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    # fallback stated by /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, pw, mx, reasons = [], [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "sm_mhz_min": min(sm) if sm else None, "power_w": statistics.median(pw) if pw else None,
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- baselines
# The reference arm / cpu_baseline score a FIXED bounded sample of the workload (per-pair cost is
# size-independent): REF_Q queries x REF_P pages of 1030 tokens, fp32, chunk_p = 64.
REF_Q, REF_P, REF_CHUNK = 32, 2048, 64
REF_SAMPLE = (f"{REF_Q} queries x {REF_P} pages x {LP_VALID} tokens fp32, chunk_p={REF_CHUNK} "
              f"(bounded sample of the workload; per-pair cost is size-independent)")


def reference_scorer():
    """(fn, kind): the reference's own score_multi_vector_masked, executed UNMODIFIED from baseline/_ref
    (tools/stage_reference.py; sha256-checked) -> kind "reference"; only if that tree is absent, the
    oracle's torch-CPU restatement of the same loop -> kind "port"."""
    from baseline import ref
    if ref.available():
        return ref.scorer_module().score_multi_vector_masked, "reference"
    from oracle.maxsim_oracle import maxsim_masked_torch_cpu
    return maxsim_masked_torch_cpu, "port"


def reference_sample(device="cpu"):
    import torch
    g = torch.Generator().manual_seed(42)
    P = torch.nn.functional.normalize(torch.randn(REF_P, LP_VALID, D, generator=g), dim=-1).to(device)
    Q = torch.nn.functional.normalize(torch.randn(REF_Q, LQ, D, generator=g), dim=-1).to(device)
    return (Q, P, torch.ones(REF_Q, LQ, dtype=torch.bool, device=device),
            torch.ones(REF_P, LP_VALID, dtype=torch.bool, device=device))


def cpu_reference_rate(steps: int = 8, warmup: int = 1):
    """The reference's CPU path on the box's host cores: `steps` timed passes over the fixed sample."""
    import torch
    fn, kind = reference_scorer()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Q, P, qm, pm = reference_sample()
    with torch.no_grad():
        for _ in range(warmup):
            fn(Q, P[:256], qm, pm[:256], REF_CHUNK)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(Q, P, qm, pm, REF_CHUNK)
        dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": REF_Q * REF_P / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} pass(es) of {REF_SAMPLE}, {dt:.2f} s per pass",
            "function": "evaluator/retrieval.py:166-213 score_multi_vector_masked"
                        + (" (unmodified, baseline/_ref)" if kind == "reference" else " (oracle port)")}, dt


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = cpu_reference_rate(steps=max(1, args.steps), warmup=1 if args.warmup else 0)
    value = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def gpu_reference_rate(torch, dev, timed):
    """The SAME unmodified reference function, torch-eager on this B200 (that is the bar a
    CUDA user of the reference has today).  Config 1 in full (500 x 500 x 1030, queries batched by 64 to
    bound the (Q, C, Lq, Lp) intermediate; outputs are batch-invariant) and the reference-arm sample of
    configs[2]; CUDA events, synchronised on both sides."""
    fn, kind = reference_scorer()
    out = {"kind": kind, "function": "evaluator/retrieval.py:166-213 score_multi_vector_masked, torch eager fp32 on the GPU"}
    g = torch.Generator().manual_seed(7)
    P = torch.nn.functional.normalize(torch.randn(500, LP_VALID, D, generator=g), dim=-1).to(dev)
    Q = torch.nn.functional.normalize(torch.randn(500, LQ, D, generator=g), dim=-1).to(dev)
    qm = torch.ones(500, LQ, dtype=torch.bool, device=dev)
    pm = torch.ones(500, LP_VALID, dtype=torch.bool, device=dev)

    def config1():
        with torch.no_grad():
            return torch.cat([fn(Q[i:i + 64], P, qm[i:i + 64], pm, REF_CHUNK) for i in range(0, 500, 64)])

    ms = timed(config1, 3, 1)
    out["config1_500x500"] = {"ms": ms, "pairs_per_s": 500 * 500 / (ms * 1e-3)}
    import evdr_b200 as evdr
    ours = lambda: evdr.score_multi_vector_masked(Q, P, qm, pm, REF_CHUNK)
    ms_o = timed(ours, 10, 3)
    ref_s, our_s = config1(), ours()
    out["config1_500x500"]["ours_ms"] = ms_o
    out["config1_500x500"]["ours_pairs_per_s"] = 500 * 500 / (ms_o * 1e-3)
    out["config1_500x500"]["max_rel_diff"] = float(((ref_s - our_s).abs() / ref_s.abs().clamp_min(1)).max())
    del P, Q, qm, pm, ref_s, our_s
    Qs, Ps, qms, pms = reference_sample(dev)

    def sample():
        with torch.no_grad():
            return fn(Qs, Ps, qms, pms, REF_CHUNK)

    ms = timed(sample, 3, 1)
    out["configs2_sample"] = {"ms": ms, "pairs_per_s": REF_Q * REF_P / (ms * 1e-3), "sample": REF_SAMPLE}
    return out


def workload_config(args, n_gpus):
    return {
        "workload": f"query-batched MaxSim: {args.queries} queries x {args.pages} pages per GPU "
                    f"(BASELINE.json configs[2]), then top-{TOPK}"
                    + (" + candidate exchange over NVLink and merge" if n_gpus > 1 else ""),
        "queries": args.queries, "pages_per_gpu": args.pages, "pages_total": args.pages * n_gpus,
        "tokens_per_page": LP_VALID, "token_slots_per_page": LP_SLOTS, "query_tokens": LQ, "dim": D,
        "top_k": TOPK, "parallelism": f"page-sharded x{n_gpus}, queries replicated",
        "exchange": "none (1 GPU)" if n_gpus == 1 else
                    ("ncclAllGather + merge (headline); one-kernel peer-memory exchange timed beside it"
                     if args.exchange == "nccl" else "one-kernel peer-memory exchange + merge"),
        "l2": f"inputs larger than L2 ({args.pages * LP_SLOTS * D * 2 / 1e9:.1f} GB corpus per GPU vs 126 MB)",
        "reference_arm_sample": REF_SAMPLE,
    }


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages", type=int, default=100_000, help="pages per GPU")
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--no-extras", action="store_true", help="skip the single-query / 64-query side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="candidate exchange of the headline at N > 1: ncclAllGather + merge (north_star) or the "
                         "one-kernel peer-memory exchange; the other mode is timed beside it")
    ap.add_argument("--no-corpus-1m", action="store_true", help="skip the 1M-page (N >= 2) / 500k-page (N = 1) top-10 QPS runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # libraries (NCCL, torch) may print to stdout; keep fd 1 for the ONE JSON line only
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import evdr_b200 as evdr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peer_ok = False
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        evdr.init_nccl_comm(dev)
        # the peer-memory exchange is set up whenever the box allows it, so that both modes are timed;
        # init_peer_exchange either succeeds on every rank or raises on every rank
        try:
            evdr.init_peer_exchange(dev, max_candidates=max(args.queries, 64) * TOPK)
            peer_ok = True
        except Exception as e:
            print(f"[bench] peer exchange unavailable on rank {rank}: {e!r}; using NCCL", file=sys.stderr)
            args.exchange = "nccl"
    n_gpus = world
    xmode = args.exchange if world > 1 else "nccl"
    peaks = load_peaks()

    # ---- synthetic shard, generated on the device (never staged through host memory)
    npg, nq = args.pages, args.queries
    corpus = evdr.synth.make_uniform_corpus(npg, LP_SLOTS, device=dev, seed=42 + 1000 * rank)
    corpus[:, LP_VALID:] = corpus[:, LP_VALID - 1:LP_VALID]       # filler slots duplicate a valid token
    pmask = torch.ones((npg, LP_SLOTS), dtype=torch.bool, device=dev)
    pk = evdr.PackedPages.from_uniform_tokens(corpus, page_offset=rank * npg)
    gq = torch.Generator().manual_seed(43)                           # identical queries on every rank
    Qh = torch.nn.functional.normalize(torch.randn(nq, LQ, D, generator=gq), dim=-1).bfloat16().float().pin_memory()
    qmh = torch.ones((nq, LQ), dtype=torch.bool).pin_memory()
    Q = Qh.to(dev)
    qm = qmh.to(dev)
    scores = torch.empty((nq, npg), dtype=torch.float32, device=dev)
    search = evdr.ShardedSearch((rank * npg, (rank + 1) * npg), mode=xmode)
    kernel_events = []                                                # (start, end) around the dominant kernel

    def step(search=search):
        Qb, qm32 = evdr.pack_queries(Q, qm)                          # 1 launch  (pack_queries_kernel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()                                                  # torch's current stream IS the launch stream
        pk.score_packed(Qb, qm32, out=scores)                        # 1 launch  (maxsim_tc_kernel)
        e1.record()
        kernel_events.append((e0, e1))
        return search.topk(scores, TOPK)                             # 1 launch  (topk_small_kernel); N > 1: + 1
                                                                     # (merge_gathered_kernel after NCCL's own
                                                                     #  all-gather, or exchange_merge_kernel)
    launches_per_step = 3 if n_gpus == 1 else 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    kernel_flavour = evdr._lib.last_kernel()                          # what the dispatch launched (evdr_last_kernel)
    pairs_per_step = nq * npg * n_gpus
    value = pairs_per_step / (ms_step * 1e-3)
    # the dominant kernel's share of those very steps (CUDA events on its launch stream, warm-up excluded)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kernel_events[-args.steps:]) / args.steps
    if world > 1:
        t = torch.tensor([ms_kernel], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_kernel = float(t.item())

    # ---- N > 1: the merged result is checked, and the other exchange mode is timed on the same steps
    sharded_check, exchange_modes = None, None
    if world > 1:
        sharded_check = check_sharded_result(torch, dist, step(), scores, npg, rank, world, dev)
        exchange_modes = {xmode: {"ms_per_step": ms_step, "value": value}}
        other = "peer" if xmode == "nccl" else "nccl"
        if other == "nccl" or peer_ok:
            s2 = evdr.ShardedSearch((rank * npg, (rank + 1) * npg), mode=other)
            ms2 = timed(lambda: step(s2), args.steps, 1)
            exchange_modes[other] = {"ms_per_step": ms2, "value": pairs_per_step / (ms2 * 1e-3)}
            chk2 = check_sharded_result(torch, dist, step(s2), scores, npg, rank, world, dev)
            if chk2 != "ok":
                sharded_check = f"{other}: {chk2}"

    # ---- end to end through the reference-facing call, host buffers in the timed region
    host_scores = torch.empty((nq, npg), dtype=torch.float32).pin_memory()

    def e2e_step():
        Qd = Qh.to(dev, non_blocking=True)
        qd = qmh.to(dev, non_blocking=True)
        s = evdr.score_multi_vector_masked(Qd, corpus, qd, pmask, 64)     # drop-in signature
        host_scores.copy_(s, non_blocking=True)

    ms_e2e = timed(e2e_step, args.steps, args.warmup)
    e2e = {"value": pairs_per_step / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": Qh.numel() * 4 + qmh.numel(), "d2h_bytes_per_step": host_scores.numel() * 4,
           "call": "score_multi_vector_masked(Q, P, qmask, pmask, 64); page tensor resident on the device "
                   "as in the reference (mainv2_iter_listwise.py:96), queries from pinned host memory, "
                   "full (nq, np) score matrix read back; the zero-copy packed view of the frozen page tensor "
                   "comes from the pack cache (built once, on the first warm-up call)"}

    # ---- roofline of the dominant kernel: its CUDA-event time INSIDE the timed steps above
    flops = nq * npg * FLOPS_PER_PAIR
    achieved = flops / (ms_kernel * 1e-3) / 1e12
    # denominator: the measured BURST cuBLAS bf16 peak (the higher of the two measured figures, so the
    # fraction can only be understated); each launch runs ~0.57 s at the power cap, where cuBLAS itself
    # sustains bf16_tflops_sustained -- reported beside it
    roofline = {"kernel": kernel_flavour,
                "bound": "tensor", "achieved": achieved,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"],
                "peak_kind": f"{peaks['source']} burst cuBLAS bf16 (8192^3, best of 10)",
                "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"],
                "sustained_peak": peaks["bf16_tflops_sustained"],
                "ms_per_launch": ms_kernel, "timed": "CUDA events around the kernel inside the timed steps",
                "share_of_step": ms_kernel / ms_step, "algorithmic_flops_per_launch": flops,
                "algorithmic_bytes_per_launch": npg * BYTES_PER_PAGE, "traffic": measured_traffic(nq, npg)}

    extras = {}
    if not args.no_extras:
        try:
            extras = side_measurements(evdr, torch, pk, corpus, dev, peaks, timed, search, n_gpus)
        except Exception as e:  # side numbers must never lose the headline line
            extras = {"error": repr(e)}
        # everything-on-the-host variant of e2e (N = 1): the page corpus itself starts in pinned host
        # memory every step and is streamed through the GPU behind the scoring (HostCorpus)
        if n_gpus == 1:
            try:
                extras["e2e_host_corpus"] = host_corpus_e2e(evdr, torch, corpus, Qh, qmh, host_scores, dev, timed,
                                                            args.steps, pairs_per_step)
            except Exception as e:
                extras["e2e_host_corpus"] = {"error": repr(e)}
        # candidate exchange alone (local top-10 + exchange + merge of resident scores), both modes
        if n_gpus > 1:
            try:
                extras["exchange_us"] = exchange_latency(evdr, torch, dev, rank, npg, timed, peer_ok)
            except Exception as e:
                extras["exchange_us"] = {"error": repr(e)}
        # the unmodified reference function torch-eager on this B200 (N = 1, rank 0)
        if n_gpus == 1:
            try:
                extras["gpu_reference"] = gpu_reference_rate(torch, dev, timed)
            except Exception as e:
                extras["gpu_reference"] = {"error": repr(e)}
        # BASELINE.json configs[3]: top-10 QPS on a 1 M-page corpus sharded over the box (N >= 2: 500 k pages =
        # 133 GB per GPU at N = 2); the 1-GPU point is a 500 k-page corpus
        if not args.no_corpus_1m:
            try:
                del pk, corpus, pmask, scores, host_scores
                torch.cuda.empty_cache()
                extras["corpus_1m" if n_gpus > 1 else "corpus_500k"] = corpus_1m(
                    evdr, torch, dist, dev, rank, n_gpus, peaks, timed, xmode, peer_ok,
                    total=1_000_000 if n_gpus > 1 else 500_000)
            except Exception as e:
                extras["corpus_1m" if n_gpus > 1 else "corpus_500k"] = {"error": repr(e)}

    if isinstance(extras.get("e2e_host_corpus"), dict) and "value" in extras["e2e_host_corpus"]:
        e2e["all_inputs_on_host"] = {k: extras["e2e_host_corpus"][k] for k in
                                     ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and n_gpus == 1:      # reported on rank 0 at N = 1 only
            try:
                cpu, _ = cpu_reference_rate()
            except Exception as e:
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, n_gpus),
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
            "peaks": peaks,
        }
        if n_gpus > 1:
            line["sharded_check"] = sharded_check
            line["exchange_modes"] = exchange_modes
        if isinstance(extras.get("gpu_reference"), dict):
            line["gpu_reference"] = extras.pop("gpu_reference")
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def check_sharded_result(torch, dist, merged, local_scores, npg, rank, world, dev, probes=4):
    """Correctness of the exchange + merge at N > 1: for `probes` queries the merged top-10 must be identical
    on every rank and equal torch.topk over the all-gathered score rows (ties: equal scores suffice)."""
    ts, ti = merged
    nq = local_scores.shape[0]
    rows = sorted({0, nq // 3, (2 * nq) // 3, nq - 1})[:probes]
    idx = torch.tensor(rows, device=dev)
    mine = local_scores.index_select(0, idx).contiguous()
    blocks = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(blocks, mine)
    full = torch.cat(blocks, dim=1)                                  # global page id = rank * npg + local id
    ref_s, ref_i = torch.topk(full, ts.shape[1], dim=1)
    got_s, got_i = ts.index_select(0, idx), ti.index_select(0, idx)
    # every rank takes part in BOTH collectives whatever it found locally (an early return would leave the
    # other ranks waiting in the second all-gather)
    all_i = [torch.empty_like(got_i) for _ in range(world)]
    dist.all_gather(all_i, got_i.contiguous())
    tie_free = (ref_s[:, 1:] != ref_s[:, :-1]).all()
    if not torch.equal(got_s, ref_s):
        status = f"rank {rank}: merged top-k scores differ from torch.topk of the gathered rows"
    elif bool(tie_free) and not torch.equal(got_i, ref_i):
        status = f"rank {rank}: merged top-k ids differ from torch.topk of the gathered rows"
    elif not torch.equal(full.gather(1, got_i.clamp_min(0)), got_s):
        status = f"rank {rank}: merged ids do not address the merged scores"
    elif not all(torch.equal(a, all_i[0]) for a in all_i):
        status = f"rank {rank}: ranks disagree on the merged top-k"
    else:
        status = "ok"
    bad = torch.tensor([0 if status == "ok" else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)                      # rank 0 prints: it must know about every rank
    if status == "ok" and int(bad.item()) != 0:
        status = "a peer rank failed its check"
    return status


def corpus_1m(evdr, torch, dist, dev, rank, n_gpus, peaks, timed, xmode="nccl", peer_ok=False, total=1_000_000):
    """Top-10 QPS on a `total`-page corpus (1 M pages = 266 GB bf16) sharded by page range over the GPUs of the
    box: streaming single-query and 64-query batches; local scan + top-10, candidate exchange, merge -- both
    exchange modes when the peer regions are set up, and the merged result checked (check_sharded_result)."""
    lo, hi = evdr.partition_pages(total, n_gpus)[rank]
    npg = hi - lo
    corpus = evdr.synth.make_uniform_corpus(npg, LP_SLOTS, device=dev, seed=4242 + 1000 * rank)
    pk = evdr.PackedPages.from_uniform_tokens(corpus, page_offset=lo)
    modes = [xmode] + ([m for m in ("nccl", "peer") if m != xmode and (m == "nccl" or peer_ok)] if n_gpus > 1 else [])
    out = {"exchange": xmode, "pages_total": total, "pages_per_gpu": npg, "corpus_gb_per_gpu": npg * LP_SLOTS * D * 2 / 1e9}
    for nq in (1, 64):
        g = torch.Generator().manual_seed(70 + nq)
        Q = torch.nn.functional.normalize(torch.randn(nq, LQ, D, generator=g), dim=-1).to(dev)
        qm = torch.ones((nq, LQ), dtype=torch.bool, device=dev)
        s = torch.empty((nq, npg), dtype=torch.float32, device=dev)
        Qh, qmh = Q.cpu().pin_memory(), qm.cpu().pin_memory()
        hs = torch.empty((nq, TOPK), dtype=torch.float32).pin_memory()
        hi_ = torch.empty((nq, TOPK), dtype=torch.int64).pin_memory()
        rec = {}
        for mode in modes:
            search = evdr.ShardedSearch((lo, hi), mode=mode)

            def one():
                Qb, qm32 = evdr.pack_queries(Q, qm)
                return search.search_packed(pk, Qb, qm32, TOPK)     # scan (+ fused top-k) + exchange + merge

            # end to end: the query batch starts in pinned host memory, the top-10 (scores, ids) end there
            def one_e2e():
                Qb, qm32 = evdr.pack_queries(Qh.to(dev, non_blocking=True), qmh.to(dev, non_blocking=True))
                ts, ti = search.search_packed(pk, Qb, qm32, TOPK)
                hs.copy_(ts, non_blocking=True)
                hi_.copy_(ti, non_blocking=True)

            ms = timed(one, 10, 3)
            ms_e2e = timed(one_e2e, 10, 3)
            r = {"ms_per_batch": ms, "top10_qps": nq / (ms * 1e-3), "pairs_per_s": nq * total / (ms * 1e-3),
                 "ms_per_batch_e2e": ms_e2e, "top10_qps_e2e": nq / (ms_e2e * 1e-3),
                 "call": "pack_queries + evdr_sharded_search (scan" + (" with the top-10 fused into its epilogue"
                         if nq <= 8 else "") + f", {mode} exchange, merge)"}
            if nq == 1:
                # the same sequence captured once as a CUDA graph and replayed (host copies in the timed region)
                try:
                    gs = evdr.GraphedSearch(search, pk, nq, TOPK)

                    def one_graph():
                        ts, ti = gs(Qh, qmh)
                        hs.copy_(ts, non_blocking=True)
                        hi_.copy_(ti, non_blocking=True)

                    ms_g = timed(one_graph, 10, 3)
                    r["ms_per_batch_graph_e2e"] = ms_g
                    r["top10_qps_graph_e2e"] = nq / (ms_g * 1e-3)
                    del gs
                except Exception as e:
                    r["graph_error"] = repr(e)
                # unfused reference point: scores row + two-stage top-k + exchange
                ms_u = timed(lambda: search.topk(pk.score_packed(*evdr.pack_queries(Q, qm), out=s), TOPK), 10, 3)
                r["ms_per_batch_unfused"] = ms_u
            if nq == 1:
                gbs = npg * BYTES_PER_PAGE / (ms * 1e-3) / 1e9
                r["hbm_gbs_per_gpu"] = gbs
                r["frac_of_hbm_peak"] = gbs / peaks["hbm_gbs"]
            else:
                tf = nq * npg * FLOPS_PER_PAIR / (ms * 1e-3) / 1e12
                r["tflops_per_gpu"] = tf
                r["frac_of_bf16_burst_peak"] = tf / peaks["bf16_tflops"]
            if n_gpus > 1:
                pk.score_packed(*evdr.pack_queries(Q, qm), out=s)            # score rows for the check
                r["sharded_check"] = check_sharded_result(torch, dist, one(), s, npg, rank, n_gpus, dev)
            if mode == xmode:
                rec.update(r)
            else:
                rec[f"exchange_{mode}"] = r
        out[f"q{nq}"] = rec
    return out


def exchange_latency(evdr, torch, dev, rank, npg, timed, peer_ok):
    """us per ShardedSearch.topk call on resident (nq, npg) score blocks: ncclAllGather + unpack + merge vs
    the one-kernel peer-memory exchange (only measured when the peer regions were set up)."""
    out = {}
    for nq in (1, 64):
        s = torch.randn((nq, npg), device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))
        for mode in (["nccl", "peer"] if peer_ok else ["nccl"]):
            ss = evdr.ShardedSearch((rank * npg, (rank + 1) * npg), mode=mode)
            out[f"q{nq}_{mode}"] = timed(lambda: ss.topk(s, TOPK), 50, 5) * 1e3
    return out


def host_corpus_e2e(evdr, torch, corpus, Qh, qmh, host_scores, dev, timed, steps, pairs_per_step):
    """The same step with NOTHING resident: queries AND the whole page corpus are read from pinned host
    memory inside the timed region (chunked host->device copies on a copy stream, overlapped with the
    scoring of the previous chunk), the (nq, np) score matrix is read back."""
    need = corpus.numel() * 2
    avail = 0
    with open("/proc/meminfo") as f:
        for ln in f:
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    if avail < 2 * need + (8 << 30):
        return {"skipped": f"host has {avail / 1e9:.0f} GB available, needs {2 * need / 1e9:.0f} GB headroom to pin the corpus"}
    host = torch.empty(corpus.shape, dtype=corpus.dtype, pin_memory=True)
    host.copy_(corpus)
    torch.cuda.synchronize()
    hc = evdr.HostCorpus(host, dev)

    def step():
        Qd = Qh.to(dev, non_blocking=True)
        qd = qmh.to(dev, non_blocking=True)
        host_scores.copy_(hc.score(Qd, qd), non_blocking=True)

    ms = timed(step, max(3, steps), 2)
    # what the link itself does: the same pinned buffer copied to the device with nothing else running
    ms_copy = timed(lambda: hc._stage[0].copy_(host[: hc.ppc], non_blocking=True), 5, 2)
    link_gbs = hc.ppc * corpus.shape[1] * corpus.shape[2] * 2 / (ms_copy * 1e-3) / 1e9
    h2d = need + Qh.numel() * 4 + qmh.numel()
    return {"value": pairs_per_step / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": host_scores.numel() * 4, "h2d_gbs": h2d / (ms * 1e-3) / 1e9, "h2d_copy_only_gbs": link_gbs,
            "pages_per_chunk": hc.ppc,
            "call": "HostCorpus(pinned bf16 corpus).score(Q, qmask): corpus, queries and masks start in pinned "
                    "host memory every step; scores are read back"}


def side_measurements(evdr, torch, pk, corpus, dev, peaks, timed, search, n_gpus):
    """Single-query streaming scan (HBM-bound), 16 / 32-query batches (the mixed tensor + HBM regime) and
    64-query batches (BASELINE.json configs[3]); each run long enough (>= ~1.5 s) for its own nvidia-smi
    clock / power record."""
    out = {}
    npg = pk.n_pages
    for nq in (1, 16, 32, 64):
        g = torch.Generator().manual_seed(50 + nq)
        Q = torch.nn.functional.normalize(torch.randn(nq, LQ, D, generator=g), dim=-1).to(dev)
        qm = torch.ones((nq, LQ), dtype=torch.bool, device=dev)
        Qb, qm32 = evdr.pack_queries(Q, qm)
        s = torch.empty((nq, npg), dtype=torch.float32, device=dev)
        ms0 = timed(lambda: pk.score_packed(Qb, qm32, out=s), 3, 3)
        reps = int(max(10, min(400, 1500.0 / ms0)))
        sampler = ClockSampler(dev.index or 0)
        sampler.start()
        ms_scan = timed(lambda: pk.score_packed(Qb, qm32, out=s), reps, 1)
        clk = sampler.stop()
        ms_search = timed(lambda: search.search_packed(pk, Qb, qm32, TOPK), 10, 3)
        # the same launches duty-cycled (each followed by an idle gap of 4 launch-times): average power stays
        # below the cap, the SM clock stays at its maximum -- what the kernel does when it is NOT power-throttled
        # (the "kernel timed alone" figure; the back-to-back number above is the sustained one)
        iso, t_end = [], time.perf_counter() + 1.2
        while time.perf_counter() < t_end or len(iso) < 5:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pk.score_packed(Qb, qm32, out=s)
            e1.record()
            torch.cuda.synchronize()
            iso.append(e0.elapsed_time(e1))
            time.sleep(4e-3 * iso[-1])
        ms_iso = statistics.median(iso[len(iso) // 3:])
        pages_s = npg / (ms_scan * 1e-3)
        rec = {"ms_scan": ms_scan, "launches_timed": reps, "ms_search_top10": ms_search, "pages_per_s_per_gpu": pages_s,
               "pairs_per_s": nq * npg * n_gpus / (ms_search * 1e-3),
               "queries_per_s_at_this_corpus": nq / (ms_search * 1e-3), "pages_total": npg * n_gpus,
               "clocks": clk, "ms_scan_duty_cycled": ms_iso,
               "duty_cycled_over_back_to_back": ms_scan / ms_iso}
        gbs = pages_s * BYTES_PER_PAGE / 1e9
        tf = nq * npg * FLOPS_PER_PAIR / (ms_scan * 1e-3) / 1e12
        if nq == 1:
            rec["roofline"] = {"kernel": "maxsim_scan_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": measured_traffic(nq, npg),
                               "algorithmic_bytes_per_launch": npg * BYTES_PER_PAGE,
                               "achieved_duty_cycled": npg * BYTES_PER_PAGE / (ms_iso * 1e-3) / 1e9}
        else:
            rec["roofline"] = {"kernel": "maxsim_tc_kernel", "bound": "tensor", "achieved": tf,
                               "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
                               "hbm_gbs_alongside": gbs, "traffic": measured_traffic(nq, npg),
                               "achieved_duty_cycled": nq * npg * FLOPS_PER_PAIR / (ms_iso * 1e-3) / 1e12,
                               "frac_duty_cycled": nq * npg * FLOPS_PER_PAIR / (ms_iso * 1e-3) / 1e12 / peaks["bf16_tflops"]}
        out[f"q{nq}"] = rec
    return out


if __name__ == "__main__":
    main()
