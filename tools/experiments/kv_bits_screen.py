"""How many significand bits the K and V planes need in the split mode (CPU emulation, 512 rows of configs[1]) -> profiles/r2ag_precision_screen_kv_bits.jsonl."""
import sys, os, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import precision_screen as ps
from oracle import captioner as oc
torch.set_num_threads(os.cpu_count())
o = oc.CaptionOracle(oc.ModelSpec())
x = oc.synthetic_embeddings(5000)[:512]
ref = ps.generate(o, x, 30, ps.r_id, ps.r_id, ps.r_id)
x2 = ps.r_split(ps.r_bf16, ps.r_bf16)
for name, rq, rk, rv in (("K bf16 (8 bits), q/V fp32", ps.r_id, ps.r_bf16, ps.r_id), ("K 5 bits", ps.r_id, ps.r_bits(5), ps.r_id), ("K 4 bits (e4m3-like)", ps.r_id, ps.r_bits(4), ps.r_id),
                         ("q,K 4 bits", ps.r_bits(4), ps.r_bits(4), ps.r_id), ("V 13 bits", ps.r_id, ps.r_id, ps.r_bits(13)), ("V 15 bits", ps.r_id, ps.r_id, ps.r_bits(15)),
                         ("K 4 bits + V fp16", ps.r_id, ps.r_bits(4), ps.r_fp16)):
    ids = ps.generate(o, x, 30, x2, x2, (rq, rk, rv), "exact_head")
    same = (ids == ref).all(dim=1)
    print(json.dumps({"scheme": "split operands + exact head; " + name, "rows": 512, "captions_identical": int(same.sum()), "match": round(float(same.float().mean()), 4)}), flush=True)
