#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): captions/sec, GPT-2 small + MLP mapper, batch 1024 per GPU, greedy 30 tokens, on synthetic
512-d embeddings and random-init weights (configs[1]); decode HBM GB/s / tensor TFLOP/s against the measured peaks.

  python bench.py --gpus N --steps K --warmup W            # this repo (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own algorithm on the host CPU cores

A "step" = one pass of the hot path over one batch: mapper -> prefill -> 29 KV-cached decode steps -> token ids for
1024 images per GPU.  `value` times it with the inputs already in HBM (CUDA events, K steps back to back, max over ranks);
`e2e` times the public `model.generate(image_embeddings=<pinned host tensor>)` call, H2D + D2H inside the timed region.

The headline is quoted in the arithmetic mode that MEETS north_star's caption tolerance (>= 99 % of greedy captions identical to
the fp32 reference): `bf16x2` -- bf16 hi + lo tensor-core operands, fp16 KV cache, exactly re-scored LM head; 99.40 % on the 5 000
rows of configs[1] (profiles/r2_parity.json).  `modes` carries the same measurement for plain `bf16` (faster, 79 % of captions --
below the contract) and `fp32` (CUDA cores, token-exact), each with its parity entry, so every throughput number has its parity
next to it.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = dict(n_embd=768, n_layer=12, n_head=12)  # GPT-2 small (124M)
E, P, V = 512, 10, 50257
W_BODY, W_WTE = 85.1e6, 38.6e6
POOL_ROWS = 5000  # val2017 size
WEIGHT_BYTES = {"bf16": 2, "bf16x2": 4, "fp32": 4}  # per weight element as the kernels read it (bf16x2: hi + lo)
KV_BYTES = {"bf16": 2, "bf16x2": 2, "fp32": 4}      # KV cache / q|k|v element (bf16x2: IEEE half)


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def parity_table() -> dict:
    """Committed parity summary per arithmetic mode (profiles/r2_parity.json, written from the -m gpu tests' report)."""
    path = os.path.join(ROOT, "profiles", "r2_parity.json")
    return json.load(open(path)) if os.path.isfile(path) else {}


def synthetic_pool(n: int, dim: int, seed: int = 1):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, dim, generator=g)
    return x / x.norm(dim=-1, keepdim=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class Watchdog:
    """A bench run must end with ONE JSON line even when a phase hangs (a rank stuck in a collective, a broken interconnect on the
    box): every phase has a time budget, the line assembled so far is handed over after every phase, and when a budget runs out rank 0
    prints that line with the key "incomplete" naming the phase, and every rank leaves with os._exit.  Before the headline numbers
    exist there is nothing to print: the run then ends with exit code 3 and a message on stderr instead of hanging."""

    def __init__(self):
        self.rank, self.emit, self.line = 0, None, None
        self.name, self.budget, self.deadline, self.finished = "start-up", 0.0, None, False
        self.lock = threading.Lock()

    def start(self, rank: int, emit, name: str, budget_s: float) -> None:
        self.rank, self.emit = rank, emit
        self.phase(name, budget_s)
        threading.Thread(target=self._run, daemon=True, name="bench-watchdog").start()

    def phase(self, name: str, budget_s: float) -> None:
        with self.lock:
            self.name, self.budget, self.deadline = name, float(budget_s), time.monotonic() + float(budget_s)

    def have(self, line: dict) -> None:
        with self.lock:
            self.line = dict(line)

    def done(self) -> None:
        with self.lock:
            self.finished = True

    def _run(self) -> None:
        while True:
            time.sleep(1.0)
            with self.lock:
                if self.finished:
                    return
                late = self.deadline is not None and time.monotonic() > self.deadline
                line, name, budget = self.line, self.name, self.budget
            if not late:
                continue
            msg = f"phase '{name}' did not finish within its {budget:.0f} s budget"
            try:
                sys.stderr.write(f"[bench watchdog, rank {self.rank}] {msg}\n")
                sys.stderr.flush()
                if line is not None and self.rank == 0 and self.emit is not None:
                    line["incomplete"] = msg + "; the keys that phase would have added are missing"
                    self.emit(line)
            finally:
                os._exit(0 if line is not None else 3)


WATCHDOG = Watchdog()


# ----------------------------------------------------------------------------------------------------------------------
def build_product_model(dtype: str, device, dims: dict | None = None, embed_dim: int = E, prefix: int = P, init_on_device: bool = False):
    """Random-init weights of the named architecture (seed 0: GPT-2 first, then the mapper -- SURVEY.md 8(d)).  init_on_device draws the
    weights with the CUDA generator instead (same seed on every rank; not the CPU-seeded weights the parity fixtures pin -- used for
    the 774M-parameter job model, whose CPU initialisation alone takes half a minute per rank)."""
    import contextlib
    import torch
    from transformers import GPT2Config, GPT2LMHeadModel
    from gpt2_image_captioning_b200 import ImageCaptioningModel, MLPMappingNetwork

    class Tok:
        eos_token_id = 50256

    dims = dims or MODEL
    torch.manual_seed(0)
    with (torch.device(device) if init_on_device else contextlib.nullcontext()):
        gpt = GPT2LMHeadModel(GPT2Config(**dims))
        mapper = MLPMappingNetwork(prefix_length=prefix, embed_dim=embed_dim, gpt_dim=dims["n_embd"])
    return ImageCaptioningModel(mapper, tokenizer=Tok(), gpt=gpt, engine_dtype=dtype).to(device).eval()


def algorithmic_decode_step(B: int, ctx: float, dtype: str) -> tuple[float, float]:
    """bytes, flops of one decode step, context `ctx` tokens incl. the new one: weights once, KV read + append,
    embedding gather, ids.  Flops are ALGORITHMIC (2 per multiply-add of the model), whatever the tensor cores spend on them."""
    d, L = MODEL["n_embd"], MODEL["n_layer"]
    sw, s = WEIGHT_BYTES[dtype], KV_BYTES[dtype]
    byt = sw * (W_BODY + W_WTE) + B * ctx * 2 * L * d * s + B * 2 * L * d * s + B * d * sw + 8 * B
    flo = 2 * B * (W_BODY + W_WTE) + 4 * B * L * d * ctx
    return byt, flo


GEMM_SHAPES = {"gemm_qkv": (3 * 768, 768), "gemm_proj": (768, 768), "gemm_fc": (4 * 768, 768), "gemm_fc2": (768, 4 * 768), "lm_head": (V, 768)}


def in_graph_timeline(eng, x, N: int, B: int, dtype: str, pk: dict) -> dict | None:
    """Every kernel of every decode step timed INSIDE the CUDA graph of the timed region's own launch chain: each block of each
    GEMM / decode-attention / ln_f / finalize launch stamps %globaltimer after griddepcontrol.wait and at its end, a launch's record keeps
    the earliest begin and the latest end (include/gic_b200.h gic_trace_install).  CUDA events cannot be placed inside the graph
    without breaking its programmatic-dependent-launch chain.  Returns per-family totals over the 29 decode steps of one batch:
    the body GEMMs (qkv / proj / fc / fc2: ONE kernel template), the LM head, decode attention, and the gaps between kernels."""
    import ctypes as C
    import torch
    from gpt2_image_captioning_b200 import _capi
    L = _capi.lib()
    cap = 1 << 14
    buf = torch.zeros(cap, 3, dtype=torch.int64, device=x.device)
    buf[:, 1] = -1  # begin = min over blocks
    try:
        _capi.check(L.gic_trace_install(C.c_void_p(buf.data_ptr()), cap))
        eng.generate_greedy(x, N)
        torch.cuda.synchronize()
    finally:
        _capi.check(L.gic_trace_install(None, 0))
    rec = buf.cpu().tolist()
    rows = sorted(((r[1], r[2], r[0] & 0xFF) for r in rec if r[2] > 0), key=lambda t: t[0])
    fins = [i for i, r in enumerate(rows) if r[2] == 4]
    if len(fins) < N:
        return None
    d, nl = MODEL["n_embd"], MODEL["n_layer"]
    s = KV_BYTES[dtype]
    fam = {k: {"ns": 0.0, "launches": 0} for k in ("gemm_body", "lm_head", "attn_decode", "other")}
    per_gemm = {k: {"ns": 0.0, "launches": 0} for k in ("gemm_qkv", "gemm_proj", "gemm_fc", "gemm_fc2")}
    attn_bytes = gap_ns = span_ns = 0.0
    order = ("gemm_qkv", "gemm_proj", "gemm_fc", "gemm_fc2")
    for t in range(1, N):  # decode step t attends P + t tokens incl. the new one
        sel = rows[fins[t - 1] + 1: fins[t] + 1]
        span_ns += sel[-1][1] - sel[0][0]
        for a, b in zip(sel[:-1], sel[1:]):
            gap_ns += max(0, b[0] - a[1])
        gi = 0
        n_gemm = sum(1 for r in sel if r[2] == 1)
        for b0, e0, k in sel:
            life = e0 - b0
            if k == 1:
                if gi < n_gemm - 1:  # 4 body GEMMs per layer in launch order, the LM head last
                    name = order[gi % 4]
                    per_gemm[name]["ns"] += life; per_gemm[name]["launches"] += 1
                    fam["gemm_body"]["ns"] += life; fam["gemm_body"]["launches"] += 1
                else:
                    fam["lm_head"]["ns"] += life; fam["lm_head"]["launches"] += 1
                gi += 1
            elif k == 2:
                fam["attn_decode"]["ns"] += life; fam["attn_decode"]["launches"] += 1
                attn_bytes += B * (P + t) * 2 * d * s + B * 2 * d * s  # strict SURVEY 8(d): K, V read over the context + the appended pair
            else:
                fam["other"]["ns"] += life; fam["other"]["launches"] += 1
    steps = N - 1
    body_flops = steps * nl * sum(2.0 * B * n * k for n, k in (GEMM_SHAPES[g] for g in order))
    head_flops = steps * 2.0 * B * V * d
    out = {"decode_step_span_us": span_ns / steps / 1e3, "launch_gap_us_per_step": gap_ns / steps / 1e3,
           "method": "%globaltimer: first block past griddepcontrol.wait .. last block's end, every launch inside the CUDA graph, one chain (one batch in flight)",
           "share_of_step": {k: v["ns"] / span_ns for k, v in fam.items()}, "gap_share_of_step": gap_ns / span_ns,
           "per_gemm_avg_us": {k: v["ns"] / max(1, v["launches"]) / 1e3 for k, v in per_gemm.items()}}
    mma_factor = 3 if dtype == "bf16x2" else 1
    tf = body_flops / fam["gemm_body"]["ns"] / 1e3
    out["gemm_body"] = {"kernel": "gemm_bf16_tcgen05_kernel (qkv + proj + fc + fc2 instantiations)", "bound": "tensor", "achieved": tf,
                        "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"], "traffic": None,
                        "avg_us": fam["gemm_body"]["ns"] / fam["gemm_body"]["launches"] / 1e3, "launches": fam["gemm_body"]["launches"],
                        "alg_flops_per_launch": body_flops / fam["gemm_body"]["launches"],
                        "tensor_pipe_TFLOPs": tf * mma_factor, "mmas_per_product": mma_factor,
                        "peak_source": pk["source"] + " (sustained cuBLAS bf16)"}
    tfh = head_flops / max(1.0, fam["lm_head"]["ns"]) / 1e3
    out["lm_head"] = {"kernel": "gemm_bf16_tcgen05_kernel (LM head + fused argmax)", "bound": "tensor", "achieved": tfh, "peak": pk["tf_sustained"],
                      "unit": "TFLOP/s", "frac": tfh / pk["tf_sustained"], "avg_us": fam["lm_head"]["ns"] / max(1, fam["lm_head"]["launches"]) / 1e3}
    gb = attn_bytes / fam["attn_decode"]["ns"]
    out["attn_decode"] = {"kernel": "attn_decode_mma_kernel", "bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": gb / pk["hbm_gbs"], "avg_us": fam["attn_decode"]["ns"] / fam["attn_decode"]["launches"] / 1e3,
                          "launches": fam["attn_decode"]["launches"], "alg_bytes_per_launch": attn_bytes / fam["attn_decode"]["launches"],
                          "bytes": "strict SURVEY 8(d): KV read over the context + the appended K/V (q / o traffic is L2-resident and not counted)",
                          "peak_source": pk["source"]}
    return out


def run_product(args, rank: int, world: int, local_rank: int) -> dict | None:
    import torch
    import torch.distributed as dist

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, K, W = args.batch, args.max_length, args.steps, max(3, args.warmup)
    pool = synthetic_pool(POOL_ROWS, E)
    pool_dev = pool.to(dev)
    pool_pin = pool.pin_memory()

    def batch_idx(i):
        return (torch.arange(B) + (rank * K + i) * B) % POOL_ROWS

    dev_batches = [pool_dev[batch_idx(i).to(dev)].contiguous() for i in range(max(K, W))]
    host_batches = [pool_pin[batch_idx(i)].contiguous().pin_memory() for i in range(max(K, W))]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # F batches in flight per GPU (gpt2_image_captioning_b200/inflight.py: one stream + engine CONTEXT + host thread each, all on one
    # copy of the packed weights; the worker streams wait for this stream and are joined back into it, so the events bracket all of them)
    from gpt2_image_captioning_b200.inflight import map_batches
    F = max(1, args.in_flight)

    def measure(dtype: str, steps: int, clocks: bool = False, e2e_too: bool = True) -> dict:
        model = build_product_model(dtype, dev)

        def dev_step(x):
            return model._get_engine().generate_greedy(x, N)

        def host_step(x):
            return model.generate(image_embeddings=x, max_length=N, temperature=0.0)

        def timed_device(f: int) -> float:
            map_batches(dev_step, [dev_batches[i % len(dev_batches)] for i in range(max(W, f))], f)
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            map_batches(dev_step, [dev_batches[i % len(dev_batches)] for i in range(steps)], f)
            ev1.record()
            barrier()
            return max_over_ranks(ev0.elapsed_time(ev1))

        eng = model._get_engine()
        seq_ms = timed_device(1) if F > 1 else None
        map_batches(dev_step, [dev_batches[i % len(dev_batches)] for i in range(max(W, F))], F)
        barrier()
        launches0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = ClockSampler(local_rank) if clocks else None
        if clk:
            clk.__enter__()
        barrier()
        e0.record()
        map_batches(dev_step, [dev_batches[i % len(dev_batches)] for i in range(steps)], F)
        e1.record()
        barrier()
        if clk:
            clk.__exit__()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        out = {"value": world * B * steps / (ms_total / 1e3), "ms_per_step": ms_total / steps, "steps": steps,
               "gpu_launches": int(eng.launch_count() - launches0), "weight_bytes": sum(e_.weight_bytes() for e_ in model.__dict__["_engines"].values()),
               "engine_contexts": len(model.__dict__["_engines"]),
               "in_flight_1": None if seq_ms is None else {"value": world * B * steps / (seq_ms / 1e3), "ms_per_step": seq_ms / steps,
                                                           "note": "the same batches one after the other on one stream"}}
        if clk:
            out["clocks"] = clk.summary()
        if e2e_too:  # end to end through the public API, pinned host buffers in, ids back on the host
            map_batches(host_step, host_batches[:max(2, F)], F)
            barrier()
            t0 = time.perf_counter()
            res = map_batches(host_step, [host_batches[i % len(host_batches)] for i in range(steps)], F)[-1]
            torch.cuda.synchronize()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
            assert res.device.type == "cpu" and tuple(res.shape) == (B, N)
            out["e2e"] = {"value": world * B * steps / e2e_s, "unit": "captions/s", "h2d_bytes_per_step": B * E * 4, "d2h_bytes_per_step": B * N * 8 + 4}
            barrier()
        out["_model"] = model
        return out

    head = measure(args.dtype, K, clocks=True)
    model = head.pop("_model")
    eng = model._get_engine()
    value, ms_total_per_step = head["value"], head["ms_per_step"]
    pk = peaks()
    par = parity_table()
    # The line is assembled as the phases finish, and handed to the watchdog after every phase: should a later phase hang (a rank stuck
    # in a collective, a box with a broken interconnect), the watchdog prints what has been measured so far with the key "incomplete"
    # instead of leaving the driver without a line.  The headline (value, e2e, clocks) comes first, then rank 0's own measurements (no
    # collectives: roofline, in-graph timeline, sampling), then the phases every rank takes part in (the other modes, the sharded jobs).
    line = {
        "metric": "captions/sec (GPT-2 greedy, 30 tokens/caption)", "value": value, "unit": "captions/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic (seeded L2-normalised 512-d embeddings, random-init weights; no network for COCO / checkpoints)",
        "config": {"workload": "configs[1]: GPT-2 small (124M) + MLP mapping net, prefix_len 10, greedy 30 tokens, batch 1024 per GPU, "
                               f"{F} batches in flight per GPU, 5k-row synthetic embedding pool, image-sharded", "batch_per_gpu": B, "max_length": N,
                   "batches_in_flight_per_gpu": F,
                   "arithmetic": {"bf16x2": "bf16 hi + lo tensor-core operands (3 tcgen05 MMAs per product), fp16 KV cache, fp32 accumulate / residual / LayerNorm / "
                                            "softmax, single-MMA LM head with exact fp32 re-scoring of the near-maximal candidates",
                                  "bf16": "single bf16 operands and KV cache", "fp32": "CUDA-core FFMA GEMMs"}[args.dtype],
                   "l2": "per-step working set (weights 0.25-0.5 GB + up to 1.5 GB KV cache) exceeds the 126 MB L2; no flush needed"},
        "parity": par.get(args.dtype),
        "clocks": head.get("clocks"), "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "in_flight_1": head["in_flight_1"],
        "weight_bytes": head["weight_bytes"], "engine_contexts": head["engine_contexts"],
    }
    WATCHDOG.have(line)

    if rank == 0:
        WATCHDOG.phase("roofline (eager kernel classes + in-graph timeline)", 150)
        # ---- per-class CUDA-event timing of the same step (eager launches: every sample carries an isolated launch's ramp) ------
        eng.profile(True)
        eng.generate_greedy(dev_batches[0], N)
        prof = eng.profile_read()
        eng.profile(False)
        ctx_mean = P + (1 + (N - 1)) / 2.0  # context incl. the new token, mean over decode steps t = 1 .. N-1
        step_total_ms = sum(v["total_ms"] for v in prof.values())
        classes = {name: {"launches": v["launches"], "avg_us": v["total_ms"] / max(1, v["launches"]) * 1e3, "share": v["total_ms"] / step_total_ms}
                   for name, v in prof.items()}
        prefill_ms = sum(v["total_ms"] for n, v in prof.items() if n in ("prefill_gemm", "attn_prefill", "mapper"))

        # ---- roofline: the kernel family with the largest share of the decode step, timed inside the timed region's CUDA graph ----
        tl = in_graph_timeline(eng, dev_batches[0], N, B, args.dtype, pk) if args.dtype != "fp32" else None
        roof = roof_attn = None
        if tl:
            shares = tl["share_of_step"]
            dom = max(("gemm_body", "lm_head", "attn_decode"), key=lambda k: shares[k])
            roof = dict(tl[dom], share_of_decode_step=shares[dom])
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full capture
            if os.path.isfile(tpath):
                roof["traffic"] = json.load(open(tpath)).get(dom + "_" + args.dtype, json.load(open(tpath)).get(dom))
            roof_attn = dict(tl["attn_decode"], share_of_decode_step=shares["attn_decode"])
        # whole decode step against max(t_HBM, t_tensor)  
        byt, flo = algorithmic_decode_step(B, ctx_mean, args.dtype)
        step_ms_eff = (ms_total_per_step - prefill_ms) / max(1, N - 1)
        step_roof = {"alg_bytes": byt, "alg_flops": flo, "t_hbm_us": byt / (pk["hbm_gbs"] * 1e9) * 1e6,
                     "t_tensor_us": flo / (pk["tf_sustained"] * 1e12) * 1e6, "measured_us": step_ms_eff * 1e3,
                     "single_chain_us": tl["decode_step_span_us"] if tl else None,
                     "achieved_GBps": byt / (step_ms_eff * 1e-3) / 1e9, "achieved_TFLOPs": flo / (step_ms_eff * 1e-3) / 1e12,
                     "note": f"measured_us = (batch time - prefill) / {N - 1} with {F} batches in flight: throughput-effective; single_chain_us = one chain's step inside the graph"}
        bound = max(step_roof["t_hbm_us"], step_roof["t_tensor_us"])
        step_roof["frac_of_max_bound"] = bound / step_roof["measured_us"]
        step_roof["frac_of_max_bound_single_chain"] = bound / tl["decode_step_span_us"] if tl else None
        line.update({
            "roofline": roof, "roofline_attention": roof_attn, "frac_of_max_bound": step_roof["frac_of_max_bound"], "decode_step_roofline": step_roof,
            "in_graph": None if not tl else {k: tl[k] for k in ("decode_step_span_us", "launch_gap_us_per_step", "share_of_step", "gap_share_of_step",
                                                                "per_gemm_avg_us", "gemm_body", "lm_head", "attn_decode", "method")},
            "kernel_classes_eager_events": classes})
        WATCHDOG.have(line)
        if not args.no_sampling:
            WATCHDOG.phase("sampling", 90)
            line["sampling"] = sampling_throughput(model, dev_batches[0], N)
            WATCHDOG.have(line)

    # the other arithmetic modes, each next to its parity entry; every rank takes part (barriers, max over ranks)
    if not args.no_modes:
        WATCHDOG.phase("modes (bf16 / fp32 beside the headline)", 200)
        modes = {}
        for dt, st in (("bf16", K), ("bf16x2", K), ("fp32", max(1, min(2, K)))):
            if dt == args.dtype:
                m = {k: v for k, v in head.items() if k in ("value", "ms_per_step", "e2e", "in_flight_1", "weight_bytes", "engine_contexts")}
            else:
                m = measure(dt, st)
                m.pop("_model").invalidate_engine()
                torch.cuda.empty_cache()
            m["parity"] = par.get(dt)
            modes[dt] = m
        line["modes"] = modes
        WATCHDOG.have(line)

    # the literal jobs of configs[1] and configs[3], host to host, sharded over the ranks (strong scaling), with the gather
    if not args.no_job:
        WATCHDOG.phase("jobs (c2 5 000 rows, c4 118 287 rows, host to host)", 300)
        line["job"] = run_jobs(args, model, dev, rank, world)
        WATCHDOG.have(line)

    if rank != 0:
        return None
    if world == 1 and not args.no_cpu_baseline:
        WATCHDOG.phase("cpu baseline", 240)
        line["cpu_baseline"] = cpu_baseline_sample(rows=args.cpu_rows, max_length=N)
    return line


def sampling_throughput(model, x, N: int) -> dict:
    """The reference's DEFAULT generate call (temperature 1.0, top_p 0.9, src/models.py:331-332) on the device."""
    import torch
    model.generate(image_embeddings=x, max_length=N)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        ids = model.generate(image_embeddings=x, max_length=N)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    return {"value": x.shape[0] / ms * 1e3, "unit": "captions/s", "ms_per_batch": ms, "temperature": 1.0, "top_p": 0.9, "tokens": int(ids.shape[1]),
            "note": "one batch at a time; nucleus sampling draws on the device (Philox), full-vocabulary logits per step"}


def run_jobs(args, model, dev, rank: int, world: int) -> dict:
    """The literal jobs, host memory to host memory through `generate_for_embeddings` (sharding.generate_sharded: every rank captions
    its contiguous shard, token ids gathered on rank 0 over the host -- the only communication of the path): configs[1] = 5 000
    embeddings with the headline model, configs[3] = 118 287 x 1024-d embeddings with GPT-2 large.  Strong scaling over the ranks.
    Cross-rank check: every rank also captions the first 128 rows of its RIGHT neighbour's shard; rank 0 compares them with the
    gathered result (with one rank: the job's first rows against a second, single-batch call)."""
    import torch
    import torch.distributed as dist
    from gpt2_image_captioning_b200 import generate_for_embeddings
    from gpt2_image_captioning_b200.sharding import shard_range
    N = args.max_length
    out = {}
    if world > 1:
        # the ids are gathered host to host, not over NCCL.  One node by contract: gloo over the loopback interface (the container
        # hostname may not resolve to a routable address), and a finite timeout so that a broken gather raises instead of hanging
        import datetime
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
        host_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=180))
    else:
        host_group = None

    def one(name, mdl, emb, batch, runs=1):
        lo, hi = shard_range(emb.shape[0], (rank + 1) % world, world)
        probe_rows = slice(lo, min(hi, lo + 128))
        probe = mdl.generate(image_embeddings=emb[probe_rows].to(dev), max_length=N, temperature=0.0).cpu()
        times = []
        for _ in range(runs):  # the first run creates the engine contexts and captures their CUDA graphs; `seconds` is the last run
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ids = generate_for_embeddings(mdl, emb, batch_size=batch, max_length=N, device=dev, in_flight=max(1, args.in_flight), group=host_group)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            times.append(dt)
        dt = times[-1]
        if world > 1:
            probes = [None] * world if rank == 0 else None
            dist.gather_object((probe_rows.start, probe.numpy()), probes, dst=0, group=host_group)
        else:
            probes = [(probe_rows.start, probe.numpy())]
        if rank != 0:
            return
        ok = all(bool((ids[s:s + p.shape[0], : p.shape[1]].numpy() == p).all()) for s, p in probes)
        out[name] = {"rows": int(emb.shape[0]), "seconds": dt, "captions_per_s": emb.shape[0] / dt, "n_gpus": world, "scaling": "strong",
                     "seconds_first_call": times[0], "runs": runs,
                     "ids_shape": list(ids.shape), "batch": batch, "cross_rank_rows_checked": int(sum(p.shape[0] for _, p in probes)),
                     "gathered_ids_equal_independent_generation": ok, "dtype": mdl.engine_dtype}

    one("c2_5000_rows_gpt2_small", model, synthetic_pool(POOL_ROWS, E).pin_memory(), args.batch, runs=2 if world == 1 else 1)
    if not args.no_c4_job:
        model.invalidate_engine()
        torch.cuda.empty_cache()
        large = build_product_model(args.dtype, dev, dict(n_embd=1280, n_layer=36, n_head=20), embed_dim=1024, prefix=10, init_on_device=True)
        one("c4_118287_rows_gpt2_large", large, synthetic_pool(118287, 1024).pin_memory(), args.batch)
        large.invalidate_engine()
        del large
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------------
def reference_generator():
    """The reference's algorithm on the CPU: the oracle port of ImageCaptioningModel.generate around HF GPT2LMHeadModel
    (no KV cache, all-position LM head, per-step host sync) -- /root/reference itself cannot travel to the GPU box."""
    import torch
    from oracle import captioner as oc

    torch.set_num_threads(os.cpu_count() or 1)
    o = oc.CaptionOracle(oc.ModelSpec())
    return o, oc


def cpu_baseline_sample(rows: int, max_length: int) -> dict:
    import torch
    o, oc = reference_generator()
    x = oc.synthetic_embeddings(POOL_ROWS, E, 1)[:rows]
    o.generate(x[:2], 2, backend="hf")  # warm the thread pool / allocator
    t0 = time.perf_counter()
    ids = o.generate(x, max_length, backend="hf")
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "captions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{rows} captions x {ids.shape[1]} tokens = BASELINE.json configs[0] (GPT-2 small + MLP mapper, batch {rows}, fp32, the reference's "
                      f"cache-less generate loop restated around HF GPT2LMHeadModel), {dt:.1f} s on {os.cpu_count()} logical cores"}


def run_reference(args, rank: int, world: int) -> dict | None:
    if rank != 0:
        return None
    import torch
    o, oc = reference_generator()
    K, W, N = args.steps, args.warmup, args.max_length
    pool = oc.synthetic_embeddings(POOL_ROWS, E, 1)
    t0 = time.perf_counter()
    o.generate(pool[:2], N, backend="hf")
    t_row = (time.perf_counter() - t0) / 2
    # the CPU baseline is configs[0], batch 64; fewer rows per step only when K + W steps of 64 would not end within
    # ~5 minutes on this box's cores
    budget_s = 300.0
    rows = int(max(1, min(64, budget_s / ((K + W) * t_row))))
    for i in range(W):
        o.generate(pool[i * rows:(i + 1) * rows], N, backend="hf")
    t0 = time.perf_counter()
    for i in range(K):
        lo = ((W + i) * rows) % (POOL_ROWS - rows)
        ids = o.generate(pool[lo:lo + rows], N, backend="hf")
    dt = time.perf_counter() - t0
    value = rows * K / dt
    sample = (f"each step = one generate() call of {rows} captions x {ids.shape[1]} tokens (configs[0] is batch 64), fp32, "
              f"{torch.get_num_threads()} threads on {os.cpu_count()} logical cores")
    return {
        "impl": "reference", "metric": "captions/sec (GPT-2 greedy, 30 tokens/caption)", "value": value, "unit": "captions/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (same seeded embeddings / random-init weights as the B200 arm)",
        "config": {"workload": "configs[0] / configs[1] model (GPT-2 small + MLP mapping net, prefix_len 10, greedy 30 tokens) on the host CPU, "
                               f"batch {rows} per generate() call", "rows_per_step": rows, "batch": rows},
        "cpu_baseline": {"value": value, "unit": "captions/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference_cuda(args, rank: int, world: int) -> dict | None:
    """Not part of the driver contract (SURVEY.md 8(d) "also worth reporting"): the reference's algorithm through stock
    PyTorch / HF CUDA kernels on the same B200 -- what a user of the reference gets by `model.to("cuda")` -- so the product arm
    has a same-device bar next to the CPU one.  Two variants: the reference's cache-less fp32 loop (src/models.py:327-477
    restated around HF GPT2LMHeadModel, host sync per step) and HF `generate` with its KV cache in bf16 (the strongest
    stock-library configuration of the same model)."""
    if rank != 0:
        return None
    import torch
    o, oc = reference_generator()
    dev = torch.device("cuda:0")
    K, W, N, B = args.steps, args.warmup, args.max_length, args.batch
    gpt, mapper = o.gpt.to(dev).eval(), o.mapper.to(dev).eval()
    pool = oc.synthetic_embeddings(POOL_ROWS, E, 1).to(dev)
    wte = gpt.transformer.wte.weight

    @torch.no_grad()
    def cacheless(x):
        cur = mapper(x)
        finished = torch.zeros(x.shape[0], dtype=torch.bool, device=dev)
        toks = []
        for _ in range(N):
            if bool(finished.all()):
                break
            nxt = torch.argmax(gpt(inputs_embeds=cur).logits[:, -1, :] / 1.0, dim=-1)
            finished = finished | nxt.eq(50256)
            nxt = torch.where(finished, torch.full_like(nxt, 50256), nxt)
            toks.append(nxt.unsqueeze(-1))
            cur = torch.cat((cur, wte[nxt].unsqueeze(1)), dim=1)
        return torch.cat(toks, dim=1)

    def timed(fn):
        for i in range(W):
            fn(pool[(i * B) % (POOL_ROWS - B):][:B])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            ids = fn(pool[((W + i) * B) % (POOL_ROWS - B):][:B])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K, ids

    ms32, ids32 = timed(cacheless)
    import copy
    gpt16, mapper16 = copy.deepcopy(gpt).to(torch.bfloat16), copy.deepcopy(mapper).to(torch.bfloat16)

    @torch.no_grad()
    def hf_cached(x):
        return gpt16.generate(inputs_embeds=mapper16(x.to(torch.bfloat16)), do_sample=False, max_new_tokens=N, min_new_tokens=N,
                              eos_token_id=50256, pad_token_id=50256, use_cache=True)

    ms16, ids16 = timed(hf_cached)
    return {
        "impl": "reference-cuda", "metric": "captions/sec (GPT-2 greedy, 30 tokens/caption)", "value": B / ms32 * 1e3, "unit": "captions/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms32, "higher_is_better": True, "dtype": "f32",
        "data": "synthetic (same seeded embeddings / random-init weights as the B200 arm)",
        "config": {"workload": f"configs[1] model, batch {B}, the reference's cache-less generate loop on cuda:0 through stock PyTorch/HF kernels"},
        "variants": {"fp32_cacheless_reference_loop": {"captions_per_s": B / ms32 * 1e3, "ms_per_step": ms32, "tokens": int(ids32.shape[1])},
                     "bf16_hf_generate_kv_cache": {"captions_per_s": B / ms16 * 1e3, "ms_per_step": ms16, "tokens": int(ids16.shape[1])}},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--dtype", default="bf16x2", choices=["bf16", "bf16x2", "fp32"],
                    help="arithmetic mode of the headline; bf16x2 is the tensor-core mode that meets the north-star caption tolerance")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--max-length", type=int, default=30)
    ap.add_argument("--cpu-rows", type=int, default=64)  # configs[0]: batch 64 (~10 s at ~7 captions/s)
    ap.add_argument("--in-flight", type=int, default=2, help="batches of --batch rows running concurrently per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the other arithmetic modes")
    ap.add_argument("--no-job", action="store_true", help="skip the host-to-host jobs (configs[1] 5 000 rows, configs[3] 118 287 rows)")
    ap.add_argument("--no-c4-job", action="store_true")
    ap.add_argument("--no-sampling", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    # stdout carries exactly one JSON line: route fd 1 to stderr while the run is on (NCCL prints its version banner to stdout at
    # the first collective, warnings of libraries may follow) and write the line to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        line = run_reference(args, rank, world)
    elif args.impl == "reference-cuda":
        line = run_reference_cuda(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            # launched without torchrun: re-exec under torch.distributed.run on this node
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
                   "127.0.0.1", "--master-port", "29577", os.path.abspath(__file__)] + sys.argv[1:]
            os.dup2(json_fd, 1)  # the ranks inherit the real stdout; rank 0 prints the line
            sys.exit(subprocess.call(cmd))
        emitted = []
        emit_lock = threading.Lock()

        def emit(ln: dict) -> None:  # exactly one line, whoever gets there first (the run or its watchdog)
            with emit_lock:
                if not emitted:
                    emitted.append(True)
                    os.write(json_fd, (json.dumps(ln) + "\n").encode())

        WATCHDOG.start(rank, emit, "start-up + headline (process group, weights, CUDA graphs, timed steps)", 420)
        line = run_product(args, rank, world, local_rank)
        if line is not None:
            sys.stdout.flush()
            emit(line)
        # the line is out: tearing the process group down must not be able to hang the run
        WATCHDOG.have({"done": True})
        WATCHDOG.phase("process-group shutdown", 30)
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        WATCHDOG.done()
        return
    if line is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
