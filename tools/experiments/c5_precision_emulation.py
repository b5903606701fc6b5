"""CPU emulation of the engine roundings on configs[4] (RAT): the configs[1] model on the 1024 retrieval-augmented embeddings of the full-size
retrieval fixture -> profiles/r2ag_precision_screen_c5.jsonl.   python tools/experiments/c5_precision_emulation.py"""
import sys, os, json, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT); os.chdir(ROOT)
import precision_screen as ps, golden_util as gu
from oracle import captioner as oc
torch.set_num_threads(os.cpu_count())
g = gu.load("c5_retrieval_full")
gen = torch.Generator().manual_seed(int(g["db_seed"]))
img = torch.randn(int(g["n_img"]), 512, generator=gen); img /= img.norm(dim=-1, keepdim=True)
cap = torch.randn(int(g["n_cap"]), 512, generator=gen); cap /= cap.norm(dim=-1, keepdim=True)
q = oc.synthetic_embeddings(1024, 512, 1)
q[:4] = img[torch.from_numpy(g["exact_rows"].astype(np.int64))]
rows = torch.from_numpy(g["rat_rows"].astype(np.int64))
ret = torch.zeros(1024, 5, 512)
m = rows >= 0
ret[m] = cap[rows[m]]
aug = q + ret.mean(dim=1)
assert np.allclose(aug[:64].numpy(), g["aug64"], atol=1e-6)
del img, cap
o = oc.CaptionOracle(oc.ModelSpec())
t0 = time.time()
ref = ps.generate(o, aug, 30, ps.r_id, ps.r_id, ps.r_id)
out = []
for name in ("hx_bf16x2_kv16", "hx_bf16"):
    ra, rw, rkv, mm, rb, head = ps.SCHEMES[name]
    ids = ps.generate(o, aug, 30, ra, rw, rkv, head)
    same = (ids == ref).all(dim=1)
    rec = {"scheme": name, "spec": "c5 (configs[1] model on the 1024 retrieval-augmented embeddings of tests/golden/c5_retrieval_full.npz: query + mean of its 5 retrieved caption rows)",
           "rows": 1024, "tokens": 30, "captions_identical": int(same.sum()), "match": round(float(same.float().mean()), 4),
           "first_256_rows_identical": int(same[:256].sum())}
    print(json.dumps(rec), flush=True)
    out.append(rec)
open("profiles/r2ag_precision_screen_c5.jsonl", "w").write("\n".join(json.dumps(r) for r in out) + "\n")
