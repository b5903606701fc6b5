"""CPU emulation of the engine roundings on all 5 000 rows of configs[1] -> profiles/r2ag_precision_screen_c2_full.jsonl (predicted mismatching rows next to the engine report)."""
import sys, os, json, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT); os.chdir(ROOT)
import precision_screen as ps, golden_util as gu
from oracle import captioner as oc
torch.set_num_threads(os.cpu_count())
o = oc.CaptionOracle(oc.ModelSpec())
x = oc.synthetic_embeddings(5000)
ref = torch.from_numpy(gu.load("c2_small_mlp_full5000")["ids"].astype(np.int64))
out = []
for name in ("hx_bf16x2_kv16", "hx_bf16"):
    ra, rw, rkv, mm, rb, head = ps.SCHEMES[name]
    ids = ps.generate(o, x, 30, ra, rw, rkv, head)
    same = (ids == ref).all(dim=1)
    bad = torch.nonzero(~same).flatten().tolist()
    rec = {"scheme": name, "spec": "c2, all 5000 rows of tests/golden/c2_small_mlp_full5000.npz", "rows": 5000, "captions_identical": int(same.sum()),
           "match": round(float(same.float().mean()), 4), "mismatched_rows": len(bad), "ragged_last_batch_match": round(float(same[4096:].float().mean()), 4),
           "first_mismatched_rows": bad[:40]}
    print(json.dumps(rec), flush=True)
    out.append(rec)
open("profiles/r2ag_precision_screen_c2_full.jsonl", "w").write("\n".join(json.dumps(r) for r in out) + "\n")
