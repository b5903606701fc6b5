#!/usr/bin/env python
"""Secondary BASELINE.json configs at full size on one B200 (the headline, configs[1], is bench.py):

  c3  GPT-2 medium + 8-layer transformer mapper, prefix_len 40, beam width 5, 30 tokens
  c4  GPT-2 large + MLP mapper on 1024-d embeddings, greedy 30 tokens
  c5  RAT: cosine top-k over the 118 287-row image matrix -> caption rows of the 591 753-row caption matrix -> mean-add,
      then GPT-2 small greedy; also the bare top-5 scan over the 591 753 caption rows

Synthetic seeded inputs, random-init weights (no network).  Each config prints one JSON line: captions/s with the inputs
resident in HBM (CUDA events, warm-up 2, 3 timed repetitions back to back) through the public generate() API.

  python tools/bench_configs.py [--configs c3,c4,c5] [--batch 1024] [--dtype bf16]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


class Tok:
    eos_token_id = 50256


IN_FLIGHT = 1


def timed(fn, warmup=2, reps=3):
    """ms per call of fn; with --in-flight F every repetition runs F calls concurrently (inflight.map_batches) and the time is per call."""
    from gpt2_image_captioning_b200.inflight import map_batches
    F = IN_FLIGHT

    def rep():
        return fn() if F == 1 else map_batches(lambda _: fn(), list(range(F)), F)[-1]

    for _ in range(warmup):
        out = rep()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = rep()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * F), out


def build(dims, mapper_kind, E, P, dtype, dev, beams=1, rat=False):
    from transformers import GPT2Config, GPT2LMHeadModel
    from gpt2_image_captioning_b200 import (ImageCaptioningModel, MLPMappingNetwork, RetrievalAugmentedTransformer,
                                            TransformerMappingNetwork)
    torch.manual_seed(0)
    gpt = GPT2LMHeadModel(GPT2Config(**dims))
    d = dims["n_embd"]
    if mapper_kind == "mlp":
        mapper = MLPMappingNetwork(prefix_length=P, embed_dim=E, gpt_dim=d)
    else:
        mapper = TransformerMappingNetwork(embed_dim=E, gpt_dim=d, prefix_length=P, hidden_length=10, num_layers=8)
    if rat:
        return RetrievalAugmentedTransformer(E, 4, "mean", mapper, tokenizer=Tok(), gpt=gpt, engine_dtype=dtype).to(dev).eval()
    return ImageCaptioningModel(mapper, tokenizer=Tok(), gpt=gpt, engine_dtype=dtype, num_beams=beams).to(dev).eval()


def main():
    global IN_FLIGHT
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c4,c5")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--dtypes", default="bf16x2,bf16", help="arithmetic modes to time; every line carries the mode's parity entry (profiles/r2_parity.json)")
    ap.add_argument("--max-length", type=int, default=30)
    ap.add_argument("--in-flight", type=int, default=2, help="batches running concurrently (inflight.py); 1 = one at a time")
    a = ap.parse_args()
    IN_FLIGHT = max(1, a.in_flight)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B, N = a.batch, a.max_length
    want = a.configs.split(",")
    par = bench.parity_table()
    for dtype in a.dtypes.split(","):
        run_dtype(a, dev, B, N, want, dtype, par.get(dtype, {}))


def run_dtype(a, dev, B, N, want, dtype, par):
    global IN_FLIGHT
    a.dtype = dtype
    if "c3" in want:
        model = build(dict(n_embd=1024, n_layer=24, n_head=16), "tfm", 512, 40, a.dtype, dev, beams=5)
        x = bench.synthetic_pool(B, 512).to(dev)
        ms, ids = timed(lambda: model.generate(image_embeddings=x, max_length=N, temperature=0.0))
        print(json.dumps({"config": "c3: GPT-2 medium + 8-layer transformer mapper, prefix_len 40, beam 5 (ancestry-table KV)", "batch": B, "dtype": a.dtype, "in_flight": IN_FLIGHT,
                          "ms_per_batch": ms, "captions_per_s": B / ms * 1e3, "ids_shape": list(ids.shape),
                          "parity": {k: par.get(k) for k in ("c3_beam5_32_rows_vs_hf", "c3_medium_tfm_full32", "c3_medium_tfm_full256")}}), flush=True)
        del model
        torch.cuda.empty_cache()
    if "c4" in want:
        model = build(dict(n_embd=1280, n_layer=36, n_head=20), "mlp", 1024, 10, a.dtype, dev)
        x = bench.synthetic_pool(B, 1024).to(dev)
        ms, ids = timed(lambda: model.generate(image_embeddings=x, max_length=N, temperature=0.0))
        print(json.dumps({"config": "c4: GPT-2 large + MLP mapper, 1024-d embeddings, greedy", "batch": B, "dtype": a.dtype, "in_flight": IN_FLIGHT, "ms_per_batch": ms,
                          "captions_per_s": B / ms * 1e3, "ids_shape": list(ids.shape), "parity": {k: par.get(k) for k in ("c4_large_mlp_full16", "c4_large_mlp_full256")}}), flush=True)
        del model
        torch.cuda.empty_cache()
    if "c5" in want:
        from gpt2_image_captioning_b200.database import GpuFlatStore
        n_img, n_cap = 118287, 591753
        g = torch.Generator().manual_seed(2)
        img = torch.randn(n_img, 512, generator=g)
        img /= img.norm(dim=-1, keepdim=True)
        cap = torch.randn(n_cap, 512, generator=g)
        cap /= cap.norm(dim=-1, keepdim=True)
        names = [f"{i:012d}.jpg" for i in range(n_img)]
        store = GpuFlatStore(img, cap, names, [{"filename": names[min(j // 5, n_img - 1)]} for j in range(n_cap)], device=dev)
        model = build(bench.MODEL, "mlp", 512, 10, a.dtype, dev, rat=True)
        x = bench.synthetic_pool(B, 512).to(dev)
        keep, IN_FLIGHT = IN_FLIGHT, 1  # the two retrieval-only timings are single calls
        ms_r, _ = timed(lambda: store.retrieve_and_aggregate(x, top_i=5, top_k=5))
        ms_s, _ = timed(lambda: store.caption_index.search_device(x, 5))
        IN_FLIGHT = keep
        ms, ids = timed(lambda: model.generate(store, 5, 5, image_embeddings=x, max_length=N, temperature=0.0))
        print(json.dumps({"config": "c5: RAT, top-5 over 118 287 images -> caption rows of 591 753 -> mean-add, GPT-2 small greedy", "batch": B,
                          "dtype": a.dtype, "in_flight": IN_FLIGHT, "ms_per_batch": ms, "captions_per_s": B / ms * 1e3, "retrieve_and_aggregate_ms": ms_r,
                          "top5_over_591753_rows_ms": ms_s, "top5_scan_TFLOPs": 2.0 * B * n_cap * 512 / ms_s * 1e-9,
                          "ids_shape": list(ids.shape), "parity": {"c5_rat_rows_vs_fp32_engine": par.get("c5_rat_rows_vs_fp32_engine"),
                                                                   "retrieval": bench.parity_table().get("retrieval")}}), flush=True)
        del model, store
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
