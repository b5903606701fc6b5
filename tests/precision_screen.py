"""CPU screen of tensor-core operand schemes against the north-star caption tolerance.  TEST INFRASTRUCTURE.

    python tests/precision_screen.py [--rows 512] [--schemes fp16,bf16,...] [--out profiles/r2_precision_screen.jsonl]

north_star: "in bf16 mode ... at least 99% of greedy captions must match exactly" (against the fp32 reference,
random-init weights).  Round 1 measured 79.7 % for plain bf16 operands, and 100 % for bf16 hi+lo
(3 MMAs per product).  This script emulates cheaper candidates with the oracle's rounding hooks
(oracle/captioner.py gpt2_forward rnd / rnd_w / rnd_kv: fp32 accumulation, fp32 residual stream, LN and
softmax, exactly what the engine keeps in fp32) on the first rows of BASELINE.json configs[1] and reports the
fraction of 30-token captions identical to the fp32 KV-cached oracle.  Negative results are recorded too.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import captioner as oc  # noqa: E402


def r_id(t):
    return t


def r_bf16(t):
    return t.bfloat16().float()


def r_fp16(t):
    return t.half().float()


def r_bits(bits):
    """Round to `bits` significand bits (incl. the implicit one), unbounded exponent."""
    def f(t):
        m, e = torch.frexp(t)
        return torch.ldexp(torch.round(m * (1 << bits)) / (1 << bits), e)
    return f


def r_split(hi, lo):
    """hi + lo two-term split: hi = hi(t), lo = lo(t - hi)."""
    def f(t):
        h = hi(t)
        return h + lo(t - h)
    return f


def r_tf32_trunc(t):
    """tf32 operand as the tensor core reads an fp32 register: low 13 mantissa bits dropped (truncation)."""
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


SCHEMES = {
    # name: (activation, weight, kv, MMAs per product, operand bytes relative to bf16)
    "bf16": (r_bf16, r_bf16, r_bf16, 1, 1.0),
    "fp16": (r_fp16, r_fp16, r_fp16, 1, 1.0),
    "fp16_kv32": (r_fp16, r_fp16, r_id, 1, 1.0),
    "fp16_Wx2": (r_fp16, r_split(r_fp16, r_fp16), r_fp16, 2, 1.5),
    "fp16_Ax2": (r_split(r_fp16, r_fp16), r_fp16, r_fp16, 2, 1.5),
    "fp16_Ax2_kv32": (r_split(r_fp16, r_fp16), r_fp16, r_id, 2, 1.5),
    "fp16_Wx2_kv32": (r_fp16, r_split(r_fp16, r_fp16), r_id, 2, 1.5),
    "tf32_trunc": (r_tf32_trunc, r_tf32_trunc, r_id, 2, 2.0),
    "bf16x2": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), r_id, 3, 2.0),
    "fp16x2": (r_split(r_fp16, r_fp16), r_split(r_fp16, r_fp16), r_id, 3, 2.0),
    # LM head made exact (1 bf16 MMA proposes candidates, the near-maximal ones are re-scored in fp32): "hx_" prefix
    "hx_fp16": (r_fp16, r_fp16, r_fp16, 1, 1.0, "exact_head"),
    "hx_bf16": (r_bf16, r_bf16, r_bf16, 1, 1.0, "exact_head"),
    "hx_fp16_Wx2": (r_fp16, r_split(r_fp16, r_fp16), r_fp16, 2, 1.5, "exact_head"),
    "hx_fp16_Wx2_kv32": (r_fp16, r_split(r_fp16, r_fp16), r_id, 2, 1.5, "exact_head"),
    "hx_fp16_Ax2": (r_split(r_fp16, r_fp16), r_fp16, r_fp16, 2, 1.5, "exact_head"),
    "hx_bf16x2_kv16": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), r_fp16, 3, 2.0, "exact_head"),
    "hx_bf16x2_kvbf16": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), r_bf16, 3, 2.0, "exact_head"),
    # which of q / k / v pays for the fp16 stores of the split mode (everything else hi + lo, exact head)
    "hx_bf16x2_q16": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), (r_fp16, r_id, r_id), 3, 2.0, "exact_head"),
    "hx_bf16x2_k16": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), (r_id, r_fp16, r_id), 3, 2.0, "exact_head"),
    "hx_bf16x2_v16": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), (r_id, r_id, r_fp16), 3, 2.0, "exact_head"),
    "hx_bf16x2_kv16_q32": (r_split(r_bf16, r_bf16), r_split(r_bf16, r_bf16), (r_id, r_fp16, r_fp16), 3, 2.0, "exact_head"),
    "head_only_fp16": (r_id, r_id, r_id, 0, 0, "head_fp16"),
    "head_only_bf16": (r_id, r_id, r_id, 0, 0, "head_bf16"),
    "Aonly_fp16": (r_fp16, r_id, r_id, 0, 0),
    "Wonly_fp16": (r_id, r_fp16, r_id, 0, 0),
    "KVonly_fp16": (r_id, r_id, r_fp16, 0, 0),
    "KVonly_bf16": (r_id, r_id, r_bf16, 0, 0),
}


def rounded_weights(w, rw):
    out = dict(w)
    out["wte"] = rw(w["wte"])
    out["layers"] = [{k: (rw(v) if k.endswith("_w") and v.dim() == 2 else v) for k, v in lw.items()} for lw in w["layers"]]
    return out


@torch.no_grad()
def generate(o, x, n_tokens, ra, rw, rkv, head=None, chunk=256):
    w = rounded_weights(o.w, rw)
    wte_in = w["wte"]  # next-token embeddings come from the rounded table
    rnd_head = None
    if head == "exact_head":
        w["wte"] = o.w["wte"]
        rnd_head = (r_id, r_id)
    elif head == "head_fp16":
        rnd_head = (r_fp16, r_fp16)
    elif head == "head_bf16":
        rnd_head = (r_bf16, r_bf16)
    mw = {k: (rw(v) if v.dim() == 2 else v) for k, v in o.mw.items()}
    outs = []
    for s in range(0, x.shape[0], chunk):
        xb = x[s:s + chunk]
        if o.spec.mapper == "mlp":
            # mapper: activations rounded by `ra` (input and tanh output), weights already rounded
            h = ra(torch.tanh(ra(xb) @ mw["model.0.weight"].t() + mw["model.0.bias"]))
            cur = (h @ mw["model.2.weight"].t() + mw["model.2.bias"]).view(xb.shape[0], o.spec.prefix_length, -1)
        else:
            cur = o.prefix(xb)  # transformer mapper (configs[2]): kept exact here -- the screen is about the 24 GPT-2 layers behind it
        kv = [None] * w["n_layer"]
        toks = []
        step_in = cur
        for _ in range(n_tokens):
            logits = oc.gpt2_forward(w, step_in, kv=kv, last_only=True, rnd=ra, rnd_w=r_id, rnd_kv=rkv, rnd_head=rnd_head or (ra, r_id))[:, -1, :]
            nxt = torch.argmax(logits, dim=-1)
            toks.append(nxt.unsqueeze(-1))
            step_in = wte_in[nxt].unsqueeze(1)
        outs.append(torch.cat(toks, dim=1))
    return torch.cat(outs, dim=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=512)
    ap.add_argument("--tokens", type=int, default=30)
    ap.add_argument("--schemes", default=",".join(SCHEMES))
    ap.add_argument("--out", default="")
    ap.add_argument("--spec", default="c2", choices=["c2", "c3", "c4"],
                    help="c2: GPT-2 small, 512-d embeddings (configs[1]); c3: GPT-2 medium + transformer mapper, prefix 40 (configs[2]); c4: GPT-2 large, 1024-d (configs[3])")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    if a.spec == "c4":  # the rows of tests/golden/c4_large_mlp_full256.npz
        o = oc.CaptionOracle(oc.ModelSpec(gpt="large", embed_dim=1024, prefix_length=10))
        x = oc.synthetic_embeddings(256, 1024, 1)[: a.rows]
    elif a.spec == "c3":  # the rows of tests/golden/c3_medium_tfm_full256.npz
        o = oc.CaptionOracle(oc.ModelSpec(gpt="medium", mapper="transformer", embed_dim=512, prefix_length=40, hidden_length=10, mapper_layers=8))
        x = oc.synthetic_embeddings(256, 512, 1)[: a.rows]
    else:
        o = oc.CaptionOracle(oc.ModelSpec())
        x = oc.synthetic_embeddings(5000)[: a.rows]
    t0 = time.time()
    ref = generate(o, x, a.tokens, r_id, r_id, r_id)
    print(f"fp32 reference: {time.time() - t0:.1f}s", flush=True)
    for name in a.schemes.split(","):
        ra, rw, rkv, mmas, rel_bytes = SCHEMES[name][:5]
        head = SCHEMES[name][5] if len(SCHEMES[name]) > 5 else None
        t0 = time.time()
        ids = generate(o, x, a.tokens, ra, rw, rkv, head)
        same = (ids == ref).all(dim=1)
        first = torch.where((ids != ref).any(dim=1), (ids != ref).float().argmax(dim=1), torch.full((ids.shape[0],), -1))
        rec = {"scheme": name, "spec": a.spec, "rows": int(x.shape[0]), "tokens": a.tokens, "captions_identical": int(same.sum()),
               "match": round(float(same.float().mean()), 4), "mmas_per_product": mmas, "operand_bytes_rel": rel_bytes,
               "median_first_flip_step": int(first[first >= 0].median()) if (first >= 0).any() else None,
               "seconds": round(time.time() - t0, 1)}
        print(json.dumps(rec), flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
