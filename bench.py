#!/usr/bin/env python
"""bench.py -- MaxSim query-page pairs/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: 1024 queries (32 tokens, 128-d) scored against
this rank's 100 000-page shard (1030 valid tokens per page, stored as 1040 bf16 slots) with the
tcgen05 MaxSim kernel, followed by the local top-10 and -- for N > 1 -- the exchange + merge of the
candidates (one peer-memory kernel over NVLink by default, ncclAllGather + merge with --exchange
nccl).  BASELINE.json configs[2] per GPU; the corpus grows with N: weak scaling.

value     : whole-job pairs/s with everything resident in HBM (CUDA events, max over ranks)
e2e       : same metric through the reference-facing call score_multi_vector_masked(Q, P, qmask,
            pmask) with the query batch copied from pinned host memory and the (nq, np) score matrix
            read back to the host inside the timed region, every step
roofline  : dominant kernel (maxsim_tc_kernel) alone: algorithmic FLOPs / CUDA-event duration vs the
            measured bf16 peak of MEASURED_PEAKS.json
cpu_baseline / --impl reference : the reference's algorithm (oracle port, torch CPU, all host
            threads) on a bounded sample of the same workload
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
        sm, mx, reasons = [], None, set()
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
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_rate(seconds_target: float = 12.0, n_queries: int = 32):
    """Reference algorithm (oracle port of evaluator/retrieval.py:166-213, torch CPU, fp32, chunk_p=64)
    on a bounded sample of the workload: n_queries x S pages of 1030 tokens."""
    import torch
    from oracle.maxsim_oracle import maxsim_masked_torch_cpu
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(42)
    def sample(n_pages):
        P = torch.nn.functional.normalize(torch.randn(n_pages, LP_VALID, D, generator=g), dim=-1)
        Q = torch.nn.functional.normalize(torch.randn(n_queries, LQ, D, generator=g), dim=-1)
        return Q, P, torch.ones(n_queries, LQ, dtype=torch.bool), torch.ones(n_pages, LP_VALID, dtype=torch.bool)
    Q, P, qm, pm = sample(64)
    maxsim_masked_torch_cpu(Q, P, qm, pm, 64)                      # warm-up
    t0 = time.perf_counter()
    maxsim_masked_torch_cpu(Q, P, qm, pm, 64)
    rate = n_queries * 64 / (time.perf_counter() - t0)
    n_pages = int(max(64, min(4096, (rate * seconds_target / n_queries) // 64 * 64)))
    Q, P, qm, pm = sample(n_pages)
    # at least `seconds_target` of CPU work: the page sample is capped at 4096 pages (2.2 GB fp32), faster
    # hosts repeat the same pass
    reps, dt = 0, 0.0
    while reps == 0 or (dt < seconds_target and reps < 64):
        t0 = time.perf_counter()
        maxsim_masked_torch_cpu(Q, P, qm, pm, 64)
        dt += time.perf_counter() - t0
        reps += 1
    return {"value": reps * n_queries * n_pages / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} pass(es) of {n_queries} queries x {n_pages} pages x {LP_VALID} tokens fp32, chunk_p=64, "
                      f"{dt:.2f} s in total"}, (Q, P, qm, pm)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port -- the
    reference is pure Python and /root/reference does not exist on the GPU box) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.maxsim_oracle import maxsim_masked_torch_cpu
    base, (Q, P, qm, pm) = cpu_reference_rate(seconds_target=max(2.0, 40.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        maxsim_masked_torch_cpu(Q, P, qm, pm, 64)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        maxsim_masked_torch_cpu(Q, P, qm, pm, 64)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = Q.shape[0] * P.shape[0] / dt
    base["value"] = value
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


def workload_config(args, n_gpus):
    return {
        "workload": f"query-batched MaxSim: {args.queries} queries x {args.pages} pages per GPU "
                    f"(BASELINE.json configs[2]), then top-{TOPK}"
                    + (" + candidate exchange over NVLink and merge" if n_gpus > 1 else ""),
        "queries": args.queries, "pages_per_gpu": args.pages, "pages_total": args.pages * n_gpus,
        "tokens_per_page": LP_VALID, "token_slots_per_page": LP_SLOTS, "query_tokens": LQ, "dim": D,
        "top_k": TOPK, "parallelism": f"page-sharded x{n_gpus}, queries replicated",
        "exchange": "none (1 GPU)" if n_gpus == 1 else
                    ("ncclAllGather + merge" if args.exchange == "nccl" else "one-kernel peer-memory exchange + merge"),
        "l2": f"inputs larger than L2 ({args.pages * LP_SLOTS * D * 2 / 1e9:.1f} GB corpus per GPU vs 126 MB)",
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
    ap.add_argument("--exchange", default="peer", choices=["nccl", "peer"],
                    help="candidate exchange at N > 1: ncclAllGather + merge, or the one-kernel peer-memory exchange")
    ap.add_argument("--corpus-1m", action="store_true", help="also run the 1M-page top-10 QPS measurement at 2 GPUs")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        evdr.init_nccl_comm(dev)
        if args.exchange == "peer":
            # init_peer_exchange either succeeds on every rank or raises on every rank (CUDA IPC peer
            # access unavailable somewhere): then all ranks use the NCCL exchange
            try:
                evdr.init_peer_exchange(dev, max_candidates=max(args.queries, 64) * TOPK)
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

    def step():
        Qb, qm32 = evdr.pack_queries(Q, qm)                          # 1 launch  (pack_queries_kernel)
        pk.score_packed(Qb, qm32, out=scores)                        # 1 launch  (maxsim_tc_kernel)
        return search.topk(scores, TOPK)                             # 1 launch  (topk_select_kernel); N > 1:
                                                                     # + unpack_gathered + merge (+ NCCL's own)
    launches_per_step = 3 if n_gpus == 1 else (4 if xmode == "peer" else 5)

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
    pairs_per_step = nq * npg * n_gpus
    value = pairs_per_step / (ms_step * 1e-3)

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
                   "full (nq, np) score matrix read back"}

    # ---- roofline of the dominant kernel, timed alone on its launch stream
    Qb, qm32 = evdr.pack_queries(Q, qm)
    ms_kernel = timed(lambda: pk.score_packed(Qb, qm32, out=scores), args.steps, 1)
    flops = nq * npg * FLOPS_PER_PAIR
    achieved = flops / (ms_kernel * 1e-3) / 1e12
    # denominator: the measured BURST cuBLAS bf16 peak (the higher of the two measured figures, so the
    # fraction can only be understated); each launch runs ~0.57 s at the power cap, where cuBLAS itself
    # sustains bf16_tflops_sustained -- reported beside it
    roofline = {"kernel": "maxsim_tc_kernel", "bound": "tensor", "achieved": achieved,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"],
                "peak_kind": f"{peaks['source']} burst cuBLAS bf16 (8192^3, best of 10)",
                "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"],
                "sustained_peak": peaks["bf16_tflops_sustained"],
                "ms_per_launch": ms_kernel, "algorithmic_flops_per_launch": flops,
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
                extras["exchange_us"] = exchange_latency(evdr, torch, dev, rank, npg, timed, xmode)
            except Exception as e:
                extras["exchange_us"] = {"error": repr(e)}
        # BASELINE.json configs[3]: top-10 QPS on a 1 M-page corpus sharded over the box (needs the
        # memory of >= 4 GPUs at 100k pages already resident; 2 GPUs only with --corpus-1m)
        if n_gpus >= 4 or (n_gpus >= 2 and args.corpus_1m):
            try:
                del pk, corpus, pmask, scores, host_scores
                torch.cuda.empty_cache()
                extras["corpus_1m"] = corpus_1m(evdr, torch, dev, rank, n_gpus, peaks, timed, xmode)
            except Exception as e:
                extras["corpus_1m"] = {"error": repr(e)}

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
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def corpus_1m(evdr, torch, dev, rank, n_gpus, peaks, timed, xmode="nccl"):
    """Top-10 QPS on a 1,000,000-page corpus (266 GB bf16) sharded by page range over the GPUs of the
    box: streaming single-query and 64-query batches, local scan + top-10, NCCL all-gather, merge."""
    total = 1_000_000
    lo, hi = evdr.partition_pages(total, n_gpus)[rank]
    npg = hi - lo
    corpus = evdr.synth.make_uniform_corpus(npg, LP_SLOTS, device=dev, seed=4242 + 1000 * rank)
    pk = evdr.PackedPages.from_uniform_tokens(corpus, page_offset=lo)
    search = evdr.ShardedSearch((lo, hi), mode=xmode)
    out = {"exchange": xmode, "pages_total": total, "pages_per_gpu": npg, "corpus_gb_per_gpu": npg * LP_SLOTS * D * 2 / 1e9}
    for nq in (1, 64):
        g = torch.Generator().manual_seed(70 + nq)
        Q = torch.nn.functional.normalize(torch.randn(nq, LQ, D, generator=g), dim=-1).to(dev)
        qm = torch.ones((nq, LQ), dtype=torch.bool, device=dev)
        s = torch.empty((nq, npg), dtype=torch.float32, device=dev)

        def one():
            Qb, qm32 = evdr.pack_queries(Q, qm)
            return search.topk(pk.score_packed(Qb, qm32, out=s), TOPK)

        ms = timed(one, 10, 3)
        # end to end: the query batch starts in pinned host memory, the top-10 (scores, ids) end there
        Qh, qmh = Q.cpu().pin_memory(), qm.cpu().pin_memory()
        hs = torch.empty((nq, TOPK), dtype=torch.float32).pin_memory()
        hi_ = torch.empty((nq, TOPK), dtype=torch.int64).pin_memory()

        def one_e2e():
            Qb, qm32 = evdr.pack_queries(Qh.to(dev, non_blocking=True), qmh.to(dev, non_blocking=True))
            ts, ti = search.topk(pk.score_packed(Qb, qm32, out=s), TOPK)
            hs.copy_(ts, non_blocking=True)
            hi_.copy_(ti, non_blocking=True)

        ms_e2e = timed(one_e2e, 10, 3)
        rec = {"ms_per_batch": ms, "top10_qps": nq / (ms * 1e-3), "pairs_per_s": nq * total / (ms * 1e-3),
               "ms_per_batch_e2e": ms_e2e, "top10_qps_e2e": nq / (ms_e2e * 1e-3)}
        if nq == 1:
            gbs = npg * BYTES_PER_PAGE / (ms * 1e-3) / 1e9
            rec["hbm_gbs_per_gpu"] = gbs
            rec["frac_of_hbm_peak"] = gbs / peaks["hbm_gbs"]
        else:
            tf = nq * npg * FLOPS_PER_PAIR / (ms * 1e-3) / 1e12
            rec["tflops_per_gpu"] = tf
            rec["frac_of_bf16_burst_peak"] = tf / peaks["bf16_tflops"]
        out[f"q{nq}"] = rec
    return out


def exchange_latency(evdr, torch, dev, rank, npg, timed, xmode):
    """us per ShardedSearch.topk call on resident (nq, npg) score blocks: ncclAllGather + unpack + merge vs
    the one-kernel peer-memory exchange (only measured when the peer regions were set up)."""
    out = {}
    for nq in (1, 64):
        s = torch.randn((nq, npg), device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))
        for mode in (["nccl", "peer"] if xmode == "peer" else ["nccl"]):
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
    """Single-query streaming scan (HBM-bound) and 64-query batches (BASELINE.json configs[3])."""
    out = {}
    npg = pk.n_pages
    for nq in (1, 64):
        g = torch.Generator().manual_seed(50 + nq)
        Q = torch.nn.functional.normalize(torch.randn(nq, LQ, D, generator=g), dim=-1).to(dev)
        qm = torch.ones((nq, LQ), dtype=torch.bool, device=dev)
        Qb, qm32 = evdr.pack_queries(Q, qm)
        s = torch.empty((nq, npg), dtype=torch.float32, device=dev)
        ms_scan = timed(lambda: pk.score_packed(Qb, qm32, out=s), 10, 3)
        ms_search = timed(lambda: search.topk(pk.score_packed(Qb, qm32, out=s), TOPK), 10, 3)
        pages_s = npg / (ms_scan * 1e-3)
        rec = {"ms_scan": ms_scan, "ms_search_top10": ms_search, "pages_per_s_per_gpu": pages_s,
               "pairs_per_s": nq * npg * n_gpus / (ms_search * 1e-3),
               "queries_per_s_at_this_corpus": nq / (ms_search * 1e-3), "pages_total": npg * n_gpus}
        if nq == 1:
            gbs = pages_s * BYTES_PER_PAGE / 1e9
            rec["roofline"] = {"kernel": "maxsim_tc_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": measured_traffic(nq, npg),
                               "algorithmic_bytes_per_launch": npg * BYTES_PER_PAGE}
        else:
            tf = nq * npg * FLOPS_PER_PAIR / (ms_scan * 1e-3) / 1e12
            rec["roofline"] = {"kernel": "maxsim_tc_kernel", "bound": "tensor", "achieved": tf,
                               "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
                               "traffic": measured_traffic(nq, npg)}
        out[f"q{nq}"] = rec
    return out


if __name__ == "__main__":
    main()
