#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): captions/sec, GPT-2 small + MLP mapper, bf16, batch 1024 per GPU, greedy 30 tokens,
on synthetic 512-d embeddings and random-init weights (configs[1]); decode HBM GB/s / tensor TFLOP/s against the measured peaks.

  python bench.py --gpus N --steps K --warmup W            # this repo (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own algorithm on the host CPU cores

A "step" = one pass of the hot path over one batch: mapper -> prefill -> 29 KV-cached decode steps -> token ids for
1024 images per GPU.  `value` times it with the inputs already in HBM (CUDA events, K steps back to back, max over ranks);
`e2e` times the public `model.generate(image_embeddings=<pinned host tensor>)` call, H2D + D2H inside the timed region.
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


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


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


# ----------------------------------------------------------------------------------------------------------------------
def build_product_model(dtype: str, device):
    """Random-init weights of the named architecture (seed 0: GPT-2 first, then the mapper -- SURVEY.md 8(d))."""
    import torch
    from transformers import GPT2Config, GPT2LMHeadModel
    from gpt2_image_captioning_b200 import ImageCaptioningModel, MLPMappingNetwork

    class Tok:
        eos_token_id = 50256

    torch.manual_seed(0)
    gpt = GPT2LMHeadModel(GPT2Config(**MODEL))
    mapper = MLPMappingNetwork(prefix_length=P, embed_dim=E, gpt_dim=MODEL["n_embd"])
    return ImageCaptioningModel(mapper, tokenizer=Tok(), gpt=gpt, engine_dtype=dtype).to(device).eval()


def algorithmic_decode_step(B: int, ctx: float, s: int) -> tuple[float, float]:
    """bytes, flops of one decode step (SURVEY.md 8(d) / BASELINE.md section 3), context `ctx` tokens incl. the new one."""
    d, L = MODEL["n_embd"], MODEL["n_layer"]
    byt = s * (W_BODY + W_WTE) + B * ctx * 2 * L * d * s + B * 2 * L * d * s + B * d * s + 8 * B
    flo = 2 * B * (W_BODY + W_WTE) + 4 * B * L * d * ctx
    return byt, flo


def in_graph_timeline(eng, x, N: int, B: int, d: int, s: int, pk: dict) -> dict | None:
    """The dominant HBM-bound kernel timed INSIDE the CUDA graph: every block of every decode-step launch stamps %globaltimer after
    griddepcontrol.wait and at its end, and a launch's record keeps the earliest begin and the latest end (include/gic_b200.h
    gic_trace_install).  CUDA events cannot be placed inside the graph without breaking its programmatic-dependent-launch chain; the
    event-timed `roofline` above therefore carries the ~5 us of an isolated launch in every sample."""
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
    life_ns = byt = gap_ns = span_ns = 0.0
    launches = 0
    for t in range(1, N):  # decode step t attends P + t tokens incl. the new one
        sel = rows[fins[t - 1] + 1: fins[t] + 1]
        span_ns += sel[-1][1] - sel[0][0]
        for a, b in zip(sel[:-1], sel[1:]):
            gap_ns += max(0, b[0] - a[1])
        for b0, e0, k in sel:
            if k == 2:
                life_ns += e0 - b0
                byt += s * B * d * (2 * (P + t) + 2 + 3 + 1)
                launches += 1
    if launches == 0:
        return None
    ach = byt / life_ns  # bytes per ns = GB/s
    return {"kernel": "attn_decode", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "avg_us": life_ns / launches / 1e3, "launches": launches, "decode_step_span_us": span_ns / (N - 1) / 1e3,
            "launch_gap_us_per_step": gap_ns / (N - 1) / 1e3,
            "method": "%globaltimer: first block past griddepcontrol.wait .. last block's end, every launch inside the CUDA graph"}


def attn_back_to_back(dev, B: int, N: int, s: int, pk: dict, traffic) -> dict:
    """The dominant HBM-bound kernel timed with CUDA events the way it runs in a decode step: for every step t = 1 .. N-1 the 12 layers'
    launches back to back (one cache plane per layer, 1.5 GB in all: nothing is found in L2), events around each group of 12
    (include/gic_b200.h gic_bench_attn_decode).  Bytes per launch: K, V [B, ctx, d] read, new K / V appended, q|k|v read, o written."""
    import ctypes as C
    import torch
    from gpt2_image_captioning_b200 import _capi
    lib = _capi.lib()
    d, L, H = MODEL["n_embd"], MODEL["n_layer"], MODEL["n_head"]
    t_max = P + N
    g = torch.Generator(device=dev).manual_seed(3)
    qkv = torch.randn(B, 3 * d, device=dev, generator=g).to(torch.bfloat16)
    kc = torch.randn(L * B * H * t_max * 64, device=dev, generator=g).to(torch.bfloat16)
    vc = torch.randn(L * B * H * t_max * 64, device=dev, generator=g).to(torch.bfloat16)
    out = torch.empty(B, d, device=dev, dtype=torch.bfloat16)
    d_pos = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def group():
        _capi.check(lib.gic_bench_attn_decode(C.c_void_p(qkv.data_ptr()), C.c_void_p(kc.data_ptr()), C.c_void_p(vc.data_ptr()),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(d_pos.data_ptr()), B, H, t_max, L, L, st))

    ms = byt = 0.0
    launches = 0
    for rep in range(4):  # rep 0 = warm-up
        pairs = []
        for t in range(1, N):
            d_pos.fill_(P + t - 1)  # tokens already cached; the launch appends one and attends P + t
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            group()
            e1.record()
            pairs.append((e0, e1, t))
        torch.cuda.synchronize()
        if rep:
            for e0, e1, t in pairs:
                ms += e0.elapsed_time(e1)
                byt += L * s * B * d * (2 * (P + t) + 2 + 3 + 1)
                launches += L
    ach = byt / (ms * 1e-3) / 1e9
    return {"kernel": "attn_decode", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": traffic, "peak_source": pk["source"], "avg_us": ms / launches * 1e3, "launches": launches,
            "method": f"CUDA events around each decode step's {L} attention launches issued back to back (one KV plane per layer, contexts "
                      f"{P + 1}..{P + N - 1}, 3 repetitions); roofline_eager_events times every launch alone, roofline_in_graph inside the CUDA graph"}


def run_product(args, rank: int, world: int, local_rank: int) -> dict | None:
    import torch
    import torch.distributed as dist

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, K, W = args.batch, args.max_length, args.steps, max(3, args.warmup)
    model = build_product_model(args.dtype, dev)
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

    # ---- device-resident throughput -------------------------------------------------------------------------------------
    # F batches in flight per GPU (gpt2_image_captioning_b200/inflight.py: one stream + engine slot + host thread each; the worker
    # streams wait for this stream and are joined back into it, so the events below bracket all of them)
    from gpt2_image_captioning_b200.inflight import map_batches
    F = max(1, args.in_flight)

    def dev_step(x):
        return model._get_engine().generate_greedy(x, N)

    def host_step(x):
        return model.generate(image_embeddings=x, max_length=N, temperature=0.0)

    def timed_device(f: int) -> float:
        map_batches(dev_step, [dev_batches[i % len(dev_batches)] for i in range(max(W, f))], f)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        map_batches(dev_step, dev_batches[:K], f)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1))

    eng = model._get_engine()
    seq_ms = timed_device(1) if F > 1 else None
    map_batches(dev_step, [dev_batches[i % len(dev_batches)] for i in range(max(W, F))], F)
    barrier()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        map_batches(dev_step, dev_batches[:K], F)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - launches0
    value = world * B * K / (ms_total / 1e3)

    # ---- end to end through the public API, host buffers ---------------------------------------------------------------------
    map_batches(host_step, host_batches[:max(2, F)], F)
    barrier()
    t0 = time.perf_counter()
    out = map_batches(host_step, host_batches[:K], F)[-1]
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert out.device.type == "cpu" and tuple(out.shape) == (B, N)
    e2e = {"value": world * B * K / e2e_s, "unit": "captions/s", "h2d_bytes_per_step": B * E * 4, "d2h_bytes_per_step": B * N * 8 + 4}
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    # ---- roofline of the dominant kernel: per-class CUDA-event timing of the same step (eager launches) ---------------------
    pk = peaks()
    s = 2 if args.dtype == "bf16" else 4
    eng.profile(True)
    eng.generate_greedy(dev_batches[0], N)
    prof = eng.profile_read()
    eng.profile(False)
    d, L = MODEL["n_embd"], MODEL["n_layer"]
    ctx_mean = P + (1 + (N - 1)) / 2.0  # context incl. the new token, mean over decode steps t = 1 .. N-1
    step_total_ms = sum(v["total_ms"] for v in prof.values())
    classes = {}
    for name, v in prof.items():
        per = v["total_ms"] / max(1, v["launches"])
        c = {"launches": v["launches"], "avg_us": per * 1e3, "share": v["total_ms"] / step_total_ms}
        if name == "attn_decode":  # per layer launch: read K,V [B, ctx, d] each, append 2 [B, d], read q|k|v, write o
            c["alg_bytes"] = s * B * d * (2 * ctx_mean + 2 + 3 + 1)
            c["GBps"] = c["alg_bytes"] / (per * 1e-3) / 1e9
        gemm_shapes = {"gemm_qkv": (3 * d, d), "gemm_proj": (d, d), "gemm_fc": (4 * d, d), "gemm_fc2": (d, 4 * d), "lm_head": (V, d)}
        if name in gemm_shapes:
            n_, k_ = gemm_shapes[name]
            calls = v["launches"]
            m_mean = B  # lm_head: N calls with M = B (prefill uses the last position only)
            c["alg_flops"] = 2.0 * m_mean * n_ * k_
            c["TFLOPs"] = c["alg_flops"] / (per * 1e-3) / 1e12
            c["alg_bytes"] = s * n_ * k_ + s * m_mean * (n_ + k_)
            c["GBps"] = c["alg_bytes"] / (per * 1e-3) / 1e9
        classes[name] = c
    decode_names = [n for n in classes if n in ("attn_decode", "gemm_qkv", "gemm_proj", "gemm_fc", "gemm_fc2", "lm_head")]
    dom = max(decode_names, key=lambda n: classes[n]["share"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(dom)
    if dom == "attn_decode":
        roof = {"kernel": dom, "bound": "hbm", "achieved": classes[dom]["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"kernel": dom, "bound": "tensor", "achieved": classes[dom]["TFLOPs"], "peak": pk["tf_sustained"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = traffic
    roof["peak_source"] = pk["source"] + (" (sustained)" if roof["bound"] == "tensor" else "")
    roof_eager = roof
    if dom == "attn_decode" and args.dtype == "bf16":
        roof = attn_back_to_back(dev, B, N, s, pk, traffic)
    # whole decode step against max(t_HBM, t_tensor)  
    byt, flo = algorithmic_decode_step(B, ctx_mean, s)
    decode_ms = sum(v["total_ms"] for n, v in prof.items() if n in decode_names + ["layernorm", "finalize", "argmax"]) - \
        sum(v["total_ms"] for n, v in prof.items() if n == "__none__")
    # the eager profile pass serialises launches with events; the graph-replayed step time comes from the main timing:
    prefill_ms = sum(v["total_ms"] for n, v in prof.items() if n in ("prefill_gemm", "attn_prefill", "mapper"))
    step_ms_graph = (ms_total / K - prefill_ms) / max(1, N - 1)
    step_roof = {"alg_bytes": byt, "alg_flops": flo, "t_hbm_us": byt / (pk["hbm_gbs"] * 1e9) * 1e6,
                 "t_tensor_us": flo / (pk["tf_sustained"] * 1e12) * 1e6, "measured_us": step_ms_graph * 1e3,
                 "achieved_GBps": byt / (step_ms_graph * 1e-3) / 1e9, "achieved_TFLOPs": flo / (step_ms_graph * 1e-3) / 1e12,
                 "note": f"measured_us = (batch time - prefill) / {N - 1} with {F} batches in flight: throughput-effective, not one chain's latency"}
    step_roof["frac_of_max_bound"] = max(step_roof["t_hbm_us"], step_roof["t_tensor_us"]) / step_roof["measured_us"]

    in_graph = in_graph_timeline(eng, dev_batches[0], N, B, d, s, pk)

    line = {
        "metric": "captions/sec (GPT-2 greedy, 30 tokens/caption)", "value": value, "unit": "captions/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic (seeded L2-normalised 512-d embeddings, random-init weights; no network for COCO / checkpoints)",
        "config": {"workload": "configs[1]: GPT-2 small (124M) + MLP mapping net, prefix_len 10, greedy 30 tokens, batch 1024 per GPU, "
                               f"{F} batches in flight per GPU, 5k-row synthetic embedding pool, image-sharded", "batch_per_gpu": B, "max_length": N, "batches_in_flight_per_gpu": F,
                   "l2": "per-step working set (0.25 GB bf16 weights + up to 1.5 GB KV cache) exceeds the 126 MB L2; no flush needed"},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "in_flight_1": None if seq_ms is None else {"value": world * B * K / (seq_ms / 1e3), "ms_per_step": seq_ms / K,
                                                    "note": "the same K batches one after the other on one stream"},
        "roofline": roof, "roofline_eager_events": roof_eager, "roofline_in_graph": in_graph, "decode_step_roofline": step_roof, "kernel_classes": classes,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(rows=args.cpu_rows, max_length=N)
    if world > 1:
        dist.destroy_process_group()
    return line


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
            "sample": f"{rows} captions x {ids.shape[1]} tokens of the same workload (GPT-2 small + MLP mapper, fp32, the reference's "
                      f"cache-less generate loop restated around HF GPT2LMHeadModel), {dt:.1f} s on {os.cpu_count()} logical cores"}


def run_reference(args, rank: int, world: int) -> dict | None:
    if rank != 0:
        return None
    import torch
    o, oc = reference_generator()
    K, W, N = args.steps, args.warmup, args.max_length
    pool = oc.synthetic_embeddings(POOL_ROWS, E, 1)
    t0 = time.perf_counter()
    o.generate(pool[:1], N, backend="hf")
    t_row = time.perf_counter() - t0
    budget_s = 150.0
    rows = int(max(1, min(64, budget_s / ((K + W) * t_row))))
    for i in range(W):
        o.generate(pool[i * rows:(i + 1) * rows], N, backend="hf")
    t0 = time.perf_counter()
    for i in range(K):
        lo = ((W + i) * rows) % (POOL_ROWS - rows)
        ids = o.generate(pool[lo:lo + rows], N, backend="hf")
    dt = time.perf_counter() - t0
    value = rows * K / dt
    sample = (f"each step = {rows} captions x {ids.shape[1]} tokens (bounded sample of the 1024-row batch), fp32, "
              f"{torch.get_num_threads()} threads on {os.cpu_count()} logical cores")
    return {
        "impl": "reference", "metric": "captions/sec (GPT-2 greedy, 30 tokens/caption)", "value": value, "unit": "captions/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (same seeded embeddings / random-init weights as the B200 arm)",
        "config": {"workload": "configs[1] model (GPT-2 small + MLP mapping net, prefix_len 10, greedy 30 tokens) on the host CPU",
                   "rows_per_step": rows},
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
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "bf16x2", "fp32"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--max-length", type=int, default=30)
    ap.add_argument("--cpu-rows", type=int, default=128)  # ~17 s of host work at ~7 captions/s
    ap.add_argument("--in-flight", type=int, default=2, help="batches of --batch rows running concurrently per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        line = run_product(args, rank, world, local_rank)
    if line is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
