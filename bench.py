#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: tokens/sec whole box + p50 TTFT, 64 users,
Llama-3-8B, 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W              # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path on the host cores

One "step" = one pass of the hot path over the whole synthetic trace: 64 users x (512-token prompt + 128 greedy
tokens), all arriving at t=0, dispatched by the fair-share scheduler and executed by the on-box GPU workers.

Printed JSON (rank 0, one line):
  value     whole-job tokens/s from DEVICE time: sum of the CUDA-event durations of every prefill pass and decode
            step of the timed steps (token ids are already in HBM when each event pair starts), max over ranks
  e2e       the same metric end to end through the public API (Dispatcher.submit -> C ABI, HOST token buffers,
            H2D of every prompt and D2H of every generated token inside the timed region), barrier-bracketed
            wall clock, max over ranks; also p50/p95 TTFT
  roofline  the decode step (one CUDA-graph launch): algorithmic bytes / CUDA-event time vs the
            measured HBM peak; plus prefill tensor-pipe numbers
  cpu_baseline  the oracle port of the forward pass on this box's host cores, bounded sample (rank 0, N=1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

USERS = 64
PROMPT_LEN = 512
GEN_LEN = 128
METRIC = "tokens/sec whole box + p50 TTFT, 64 users, Llama-3-8B"
UNIT = "tokens/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


def prompts():
    import numpy as np
    from ollamamq_b200.models import LLAMA3_8B
    return [np.random.default_rng(u).integers(0, LLAMA3_8B["vocab"], PROMPT_LEN).astype("int32").tolist()
            for u in range(USERS)]


def shard_users(n_backends: int):
    """Partition of the 64 users over the workers = the reference's own backend pick (least connections,
    round-robin tie-break; dispatcher.rs:247-254) on the t=0 arrival trace.  Every rank computes it alone."""
    import ollamamq_b200 as mq
    s = mq.Scheduler(n_backends, capacity=USERS)
    for u in range(USERS):
        s.enqueue("user%02d" % u)
    out = [[] for _ in range(n_backends)]
    for d in s.drain():
        out[d.backend].append(int(d.user[4:]))
    return out


def _decode_traffic():
    """dram__bytes_read+write of one decode step from the committed ncu capture (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_decode_traffic.json")))["decode_step_dram_bytes"]
    except Exception:
        return None


class ClockSampler:
    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            pass
        sm, pw, mx, reasons = [], [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                c, m, w = float(f[0]), float(f[1]), float(f[2])
            except ValueError:
                continue
            sm.append(c)
            pw.append(w)
            mx = max(mx, m)
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load" = samples drawing at least half of the highest power seen during the timed region
        thr = 0.5 * max(pw) if pw else 0.0
        load = sorted(c for c, w in zip(sm, pw) if w >= thr)
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(load), "sm_mhz_min_under_load": load[0] if load else None,
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ollamamq_b200 as mq
    from ollamamq_b200.models import LLAMA3_8B      # the product arm never touches oracle/

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log("warning: WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    n = max(world, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    my_users = shard_users(n)[rank]
    P = prompts()
    pdl = int(os.environ.get("MQ_PDL", "1"))
    graphs = int(os.environ.get("MQ_GRAPHS", "1"))
    cfg = mq.model_cfg(LLAMA3_8B, max_batch=USERS, max_seq=PROMPT_LEN + GEN_LEN + 16,
                       max_prefill_tokens=int(os.environ.get("MQ_PREFILL_TOKENS", "4736")), use_graphs=graphs,
                       use_pdl=pdl, model_name="llama-3-8b-random-init")
    t0 = time.time()
    wk = mq.Worker(local, cfg)
    wk.init_random(seed=0, std=0.02)
    wk.set_timing(True)
    disp = mq.Dispatcher([wk], capacity=USERS)
    log("[rank %d] worker up in %.1fs, %d users" % (rank, time.time() - t0, len(my_users)))

    def one_step():
        streams = []
        t_start = time.perf_counter()
        for u in my_users:
            streams.append(disp.submit("user%02d" % u, prompt_tokens=P[u], max_new_tokens=GEN_LEN, stream=1))
        for s in streams:
            s.wait(600)
        t_end = time.perf_counter()
        ttft = []
        ntok = 0
        for s in streams:
            if s.rc != 0:
                raise RuntimeError("request failed: rc=%s %s" % (s.rc, s.err))
            ntok += len(s.body) // 4
            ttft.append(s.chunk_times[0] - t_start)
        return t_end - t_start, ntok, ttft

    for _ in range(args.warmup):
        one_step()
    barrier()
    wk.reset_stats()
    sampler = ClockSampler(local) if rank == 0 else None
    wall, toks, ttfts = 0.0, 0, []
    barrier()
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        barrier()
        dt, ntok, tt = one_step()
        barrier()
        wall += allmax(dt)
        toks += ntok
        ttfts += tt
    t_all = time.perf_counter() - t_all0
    clocks = sampler.stop() if sampler else None
    st = wk.stats()
    dev_ms = st["decode_ms"] + st["prefill_ms"]
    dev_ms_max = allmax(dev_ms)
    toks_all = allsum(toks)
    launches = allsum(st["kernel_launches"])
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, ttfts)
        ttfts = [x for g in gathered for x in g]
    ttfts.sort()
    pk = peaks()
    # roofline of the decode step on this rank (rank 0 reports its own)
    dec_steps = max(1, st["decode_steps"])
    dec_gbs = st["decode_bytes"] / (st["decode_ms"] * 1e-3) / 1e9 if st["decode_ms"] > 0 else 0.0
    # prefill flops: 2*P_mm*T + causal attention
    g = LLAMA3_8B
    p_mm = g["n_layers"] * ((g["n_q_heads"] + 2 * g["n_kv_heads"]) * 128 * g["hidden"] + g["hidden"] * g["hidden"]
                            + 3 * g["ffn"] * g["hidden"]) + g["vocab"] * g["hidden"]
    per_prompt = 2.0 * (p_mm - g["vocab"] * g["hidden"]) * PROMPT_LEN + 2.0 * g["vocab"] * g["hidden"] \
        + g["n_layers"] * 4.0 * PROMPT_LEN * PROMPT_LEN * g["n_q_heads"] * 128 / 2
    prefill_tf = per_prompt * len(my_users) * args.steps / (st["prefill_ms"] * 1e-3) / 1e12 if st["prefill_ms"] > 0 else 0.0
    bytes_h2d = sum(len(P[u]) for u in my_users) * 4
    bytes_d2h = len(my_users) * GEN_LEN * 4
    line = {
        "metric": METRIC, "value": toks_all / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights, uniform random token ids)",
        "config": {"workload": "BASELINE configs[1]: Llama-3-8B bf16, 64 concurrent users, 512-token prompt / "
                               "128-token greedy decode, token stream per request",
                   "users": USERS, "prompt_len": PROMPT_LEN, "gen_len": GEN_LEN,
                   "parallelism": "%d independent workers (replicated weights), users sharded by the reference "
                                  "backend pick, no collective on the data path" % n,
                   "l2": "inputs larger than L2: 15 GB of weights + the KV cache stream through HBM every decode step",
                   "pdl": pdl, "cuda_graphs": graphs, "prefill_tokens_per_pass": cfg.max_prefill_tokens},
        "e2e": {"value": toks_all / wall, "unit": UNIT, "h2d_bytes_per_step": bytes_h2d * n,
                "d2h_bytes_per_step": bytes_d2h * n, "ms_per_step": wall / args.steps * 1e3,
                "ttft_p50_ms": ttfts[len(ttfts) // 2] * 1e3, "ttft_p95_ms": ttfts[int(len(ttfts) * 0.95)] * 1e3},
        "ttft_p50_ms": ttfts[len(ttfts) // 2] * 1e3,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "decode step (1 CUDA graph launch = %d kernels; tcgen05 weight-streaming "
                                               "GEMMs + paged-KV attention)" % (1 + 32 * 8 + 3),
                     "achieved": dec_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dec_gbs / pk["hbm_gbs"],
                     "peak_src": pk["src"], "traffic": _decode_traffic(),
                     "decode_ms_per_step": st["decode_ms"] / dec_steps,
                     "decode_bytes_per_step": st["decode_bytes"] / dec_steps,
                     "prefill": {"bound": "tensor", "achieved": prefill_tf, "peak": pk["tf_sustained"],
                                 "unit": "TFLOP/s", "frac": prefill_tf / pk["tf_sustained"],
                                 "prefill_ms_per_step": st["prefill_ms"] / args.steps}},
        "clocks": clocks,
        "wall_s_timed_region": t_all,
    }
    disp.close()
    wk.close()
    if rank == 0:
        line["dispatch"] = dispatch_decisions()
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm
def _fast_cpu_weights(cfg):
    """Weights for a TIMING run of the CPU port: values do not matter, so tile one random block instead of drawing
    8e9 normals.  Stored as fp32: torch's CPU bf16 GEMV path is ~20x slower than its fp32 path on this Xeon
    (measured 2.5 s/token vs 0.2 s/token), and the baseline should be the CPU's best foot forward."""
    import torch
    from oracle.llama_ref import tensor_shapes
    blk = torch.randn(1 << 22) * 0.02
    w = {}
    for name, shape in tensor_shapes(cfg).items():
        n = 1
        for s in shape:
            n *= s
        if name.endswith("norm"):
            w[name] = torch.ones(shape, dtype=torch.float32)
        else:
            reps = (n + blk.numel() - 1) // blk.numel()
            w[name] = blk.repeat(reps)[:n].view(shape).contiguous()
    return w


_CPU_W = None


def cpu_reference_step(gen=4):
    """The reference path on the host cores: the dispatch oracle in front of the CPU port of the forward pass,
    reference semantics (one in-flight request per backend, one backend => requests are served one after the other).
    Bounded sample: ONE request of the trace, the full 512-token prompt + `gen` greedy tokens.
    Returns (t_prefill_s, t_per_decode_token_s, wall_s)."""
    import torch
    from oracle.dispatch_oracle import OraclePy, simulate
    from oracle.llama_ref import LLAMA3_8B, forward
    global _CPU_W
    if _CPU_W is None:
        _CPU_W = _fast_cpu_weights(LLAMA3_8B)
    P = prompts()
    (user, _, _), = simulate(OraclePy(1), [(0, "user00")], lambda u, s, b: 1)
    u = int(user[4:])
    t0 = time.perf_counter()
    kv = []
    logits = forward(_CPU_W, LLAMA3_8B, P[u], torch.float32, 0, kv, last_only=True)
    tok = int(logits[-1].argmax())
    t1 = time.perf_counter()
    for i in range(gen - 1):
        logits = forward(_CPU_W, LLAMA3_8B, [tok], torch.float32, PROMPT_LEN + i, kv, last_only=True)
        tok = int(logits[-1].argmax())
    t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / max(1, gen - 1), t2 - t0


def _cpu_trace_rate(t_prefill, t_tok):
    """tokens/s of the 64 x (512 + 128) trace when requests are served one at a time (capacity 1): every request costs
    t_prefill + 127 * t_tok and yields 128 tokens, so the trace rate equals the per-request rate."""
    return GEN_LEN / (t_prefill + (GEN_LEN - 1) * t_tok)


def _cpu_sample_text(n):
    return ("%d x [1 request of the trace: the 512-token prompt prefilled + 4 greedy tokens decoded, Llama-3-8B geometry, "
            "fp32 math, torch CPU port (oracle/llama_ref.py) behind the dispatch oracle, capacity 1 like the reference]; "
            "value = 128 / (t_prefill + 127 * t_decode_token) from the two MEASURED components, i.e. a linear "
            "extrapolation of the bounded sample to the 128-token requests of the trace; the reference itself "
            "(Rust + Ollama/llama.cpp) cannot be built or installed in this image" % n)


def dispatch_decisions():
    """decisions/s of the C++ scheduler at 64 / 256 users (host CPU, microseconds of work)."""
    import ctypes as C
    import ollamamq_b200 as mq
    out = {"unit": "decisions/s", "reqs_per_user": 64}
    for users in (64, 256):
        nd, sec = C.c_uint64(), C.c_double()
        mq.check(mq.lib.mq_debug_sched_bench(users, 64, 1, 1, C.byref(nd), C.byref(sec)))
        out["users_%d" % users] = nd.value / max(sec.value, 1e-9)
    return out


def cpu_baseline_sample():
    import torch
    tp, tt, _ = cpu_reference_step(4)
    return {"value": _cpu_trace_rate(tp, tt), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": _cpu_sample_text(1), "prefill_s": tp, "decode_s_per_token": tt,
            "ttft_first_request_ms": tp * 1e3}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.warmup > 0:
        cpu_reference_step(2)          # one short warm-up is enough for a CPU loop (weights get paged in)
    tps, tts, wall = [], [], 0.0
    for _ in range(args.steps):
        tp, tt, w = cpu_reference_step(4)
        tps.append(tp)
        tts.append(tt)
        wall += w
    tp, tt = sum(tps) / len(tps), sum(tts) / len(tts)
    v = _cpu_trace_rate(tp, tt)
    # with capacity 1 the k-th of 64 simultaneous users waits for k-1 whole requests: p50 TTFT of the trace
    ttft_p50 = (32 * (tp + (GEN_LEN - 1) * tt) + tp) * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] (bounded sample): Llama-3-8B, 512-token prompt, greedy decode",
                       "users": USERS, "prompt_len": PROMPT_LEN, "gen_len": GEN_LEN},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": _cpu_sample_text(args.steps), "prefill_s": tp, "decode_s_per_token": tt},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "ttft_p50_ms_extrapolated": ttft_p50, "ttft_first_request_ms": tp * 1e3},
            "ttft_p50_ms": ttft_p50, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
