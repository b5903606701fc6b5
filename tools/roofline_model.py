#!/usr/bin/env python
"""Roofline bounds of one whole batch (prefill + decode steps) for configs 2-5 of BASELINE.json, from the shapes alone (SURVEY.md 8(d)
formulas), next to the measured time per batch from the committed profiles.  CPU only.

    python tools/roofline_model.py > profiles/r2ah_roofline_model.txt

Per decode step with R rows in flight (R = images x beams), context c (incl. the new token), L layers, width d, weight element size
s_w, KV element size s:

    bytes = s_w (W_body + W_wte) + KV_read(c) + R 2 L d s (append) + R d s_w (next-token gather) + 8 R
    flops = 2 R (W_body + W_wte) + 4 R L d c                      (ALGORITHMIC: whatever the tensor cores spend on a product)
    KV_read(c) = R c 2 L d s            greedy
               = (R / beams) P 2 L d s + R (c - P) 2 L d s        beam search: the image prefix is read once per image
    t_step >= max(bytes / HBM peak, flops / tensor peak)

Prefill (R_img rows x P positions): flops = 2 R_img P W_body + 2 R_img W_wte (last position only) + attention; its bytes are the
weights once more.  The peaks are the measured ones of MEASURED_PEAKS.json (copy bandwidth, sustained cuBLAS bf16).
"""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = 50257
MODELS = {"small": (768, 12), "medium": (1024, 24), "large": (1280, 36)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")  # driver-written per pod; B200_PROFILING.md's fallback figures otherwise
    if not os.path.isfile(path):
        return 6650.0e9, 1400.0e12
    p = json.load(open(path))
    return p["hbm_gbs"] * 1e9, p["bf16_tflops_sustained"] * 1e12


def batch_bound(model: str, images: int, P: int, n_tokens: int, beams: int, s_w: int, s_kv: int, mma_per_product: int = 1):
    d, L = MODELS[model]
    w_body, w_wte = 12 * d * d * L, V * d
    hbm, tf = peaks()
    R = images * beams
    # prefill over the images' prefixes
    pf_flops = 2.0 * images * P * w_body + 2.0 * images * w_wte + 4.0 * images * L * d * P * (P + 1) / 2
    pf_bytes = s_w * (w_body + w_wte) + images * P * 2 * L * d * s_kv
    t_alg = max(pf_bytes / hbm, pf_flops / tf)
    t_exec = max(pf_bytes / hbm, (mma_per_product * 2.0 * images * P * w_body + 2.0 * images * w_wte) / tf)
    tot_bytes, tot_flops = pf_bytes, pf_flops
    for t in range(1, n_tokens):  # decode step t attends P + t tokens incl. the new one
        c = P + t
        kv_read = (images * P + R * (c - P)) * 2 * L * d * s_kv if beams > 1 else R * c * 2 * L * d * s_kv
        byt = s_w * (w_body + w_wte) + kv_read + R * 2 * L * d * s_kv + R * d * s_w + 8 * R
        flo = 2.0 * R * (w_body + w_wte) + 4.0 * R * L * d * c
        t_alg += max(byt / hbm, flo / tf)
        t_exec += max(byt / hbm, (mma_per_product * 2.0 * R * w_body + 2.0 * R * w_wte + 4.0 * R * L * d * c) / tf)
        tot_bytes += byt
        tot_flops += flo
    return {"t_alg_ms": t_alg * 1e3, "t_exec_ms": t_exec * 1e3, "GB": tot_bytes / 1e9, "TFLOP": tot_flops / 1e12}


def measured():
    out = {}
    b = json.load(open(os.path.join(ROOT, "profiles", "r2ac_bench.json")))
    out[("c2", "bf16x2")] = b["modes"]["bf16x2"]["ms_per_step"]
    out[("c2", "bf16")] = b["modes"]["bf16"]["ms_per_step"]
    for line in open(os.path.join(ROOT, "profiles", "r2y_configs.jsonl")):
        r = json.loads(line)
        out[(r["config"][:2], r["dtype"])] = r["ms_per_batch"]
    return out


def main():
    hbm, tf = peaks()
    print(f"# tools/roofline_model.py: bounds of one 1024-image batch (30 tokens) against the measured time per batch; peaks {hbm / 1e9:.0f} GB/s, {tf / 1e12:.0f} TFLOP/s")
    print("# t_alg: algorithmic flops (1 per multiply-add pair x2) / sustained tensor peak vs bytes / HBM peak, per step, summed;")
    print("# t_exec: the same with the tensor flops the mode EXECUTES in the body GEMMs (bf16x2: three MMAs per product)")
    print(f"{'config':44s} {'mode':7s} {'GB':>7s} {'TFLOP':>7s} {'t_alg ms':>9s} {'t_exec ms':>9s} {'measured ms':>11s} {'frac alg':>8s} {'frac exec':>9s}")
    m = measured()
    rows = [("c2", "c2 GPT-2 small, MLP mapper, P 10, greedy", "small", 10, 1), ("c3", "c3 GPT-2 medium, tfm mapper, P 40, beam 5", "medium", 40, 5),
            ("c4", "c4 GPT-2 large, MLP mapper, P 10, greedy", "large", 10, 1), ("c5", "c5 RAT: c2 + retrieval (0.61 ms, not in the bound)", "small", 10, 1)]
    for key, name, model, P, beams in rows:
        for mode, s_w, s_kv, mma in (("bf16x2", 4, 2, 3), ("bf16", 2, 2, 1)):
            r = batch_bound(model, 1024, P, 30, beams, s_w, s_kv, mma)
            ms = m.get((key, mode))
            print(f"{name:44s} {mode:7s} {r['GB']:7.1f} {r['TFLOP']:7.1f} {r['t_alg_ms']:9.2f} {r['t_exec_ms']:9.2f} "
                  f"{ms if ms is None else round(ms, 1):>11} {'' if ms is None else format(r['t_alg_ms'] / ms, '8.2f')} {'' if ms is None else format(r['t_exec_ms'] / ms, '9.2f')}")


if __name__ == "__main__":
    main()
