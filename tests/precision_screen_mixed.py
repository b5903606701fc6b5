"""CPU screen of MIXED operand schemes: one GEMM family of the GPT-2 layer at a time downgraded from bf16 hi+lo (3 MMAs per product)
to a cheaper scheme, everything else as the bf16x2 engine runs it (exact LM head, fp16 KV cache).  TEST INFRASTRUCTURE (uses oracle/).

    python tests/precision_screen_mixed.py [rows] [scheme,scheme,...]      results: profiles/r2p_precision_screen_mixed.jsonl

Question it answers: can any of qkv / proj / fc / fc2 drop to 1 or 2 MMAs and keep north_star's ">= 99 % of greedy captions identical"
on random-init weights?  Measured: no -- every single-family downgrade to one fp16 MMA lands at 98.3-98.7 %, fp16 A x split W on qkv
alone at 99.2 % (no margin), against 99.6 % for the all-bf16x2 engine emulation.
"""
import sys, os, json, time, math, torch
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
from oracle import captioner as oc
import precision_screen as ps
r_id, r_fp16, r_bf16 = ps.r_id, ps.r_fp16, ps.r_bf16
x2 = ps.r_split(r_bf16, r_bf16)
f2 = ps.r_split(r_fp16, r_fp16)

@torch.no_grad()
def forward(w, x, kv, ops, rkv):
    B, T, d = x.shape; H = w["n_head"]; hd = d // H
    past = 0 if kv[0] is None else kv[0][0].shape[2]
    pos = torch.arange(past, past + T)
    h = x + w["wpe"][pos]
    for li, lw in enumerate(w["layers"]):
        ra, rw_ = ops["qkv"]
        a = ra(oc.layer_norm(h, lw["ln1_w"], lw["ln1_b"]))
        qkv = rkv(a @ rw_(lw["attn_w"]) + lw["attn_b"])
        q, k, v = qkv.split(d, dim=-1)
        q = q.view(B, T, H, hd).transpose(1, 2); k = k.view(B, T, H, hd).transpose(1, 2); v = v.view(B, T, H, hd).transpose(1, 2)
        if kv[li] is not None:
            k = torch.cat((kv[li][0], k), dim=2); v = torch.cat((kv[li][1], v), dim=2)
        kv[li] = [k, v]
        S = k.shape[2]
        att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        causal = torch.ones(S, S, dtype=torch.bool).tril()[S - T:, :]
        att = att.masked_fill(~causal, float("-inf")).softmax(-1)
        ra, rw_ = ops["proj"]
        o = ra((att @ v).transpose(1, 2).reshape(B, T, d))
        h = h + (o @ rw_(lw["proj_w"]) + lw["proj_b"])
        ra, rw_ = ops["fc"]
        m = ra(oc.layer_norm(h, lw["ln2_w"], lw["ln2_b"]))
        g = oc.gelu_new(m @ rw_(lw["fc_w"]) + lw["fc_b"])
        ra, rw_ = ops["fc2"]
        h = h + (ra(g) @ rw_(lw["fc2_w"]) + lw["fc2_b"])
    h = oc.layer_norm(h[:, -1:, :], w["lnf_w"], w["lnf_b"])
    return h @ w["wte"].t()

@torch.no_grad()
def generate(o, x, n_tokens, ops, rkv, chunk=256):
    w = o.w
    # cache rounded weights per op
    cache = {}
    def mk(rw):
        def f(t):
            key = (id(rw), t.data_ptr())
            if key not in cache: cache[key] = rw(t)
            return cache[key]
        return f
    ops = {k: (ra, mk(rw)) for k, (ra, rw) in ops.items()}
    outs = []
    for s in range(0, x.shape[0], chunk):
        xb = x[s:s + chunk]
        hh = x2(torch.tanh(x2(xb) @ x2(o.mw["model.0.weight"]).t() + o.mw["model.0.bias"]))
        cur = (hh @ x2(o.mw["model.2.weight"]).t() + o.mw["model.2.bias"]).view(xb.shape[0], o.spec.prefix_length, -1)
        kv = [None] * w["n_layer"]; toks = []; step_in = cur
        for _ in range(n_tokens):
            logits = forward(w, step_in, kv, ops, rkv)[:, -1, :]
            nxt = torch.argmax(logits, dim=-1); toks.append(nxt.unsqueeze(-1))
            step_in = w["wte"][nxt].unsqueeze(1)
        outs.append(torch.cat(toks, dim=1))
    return torch.cat(outs, dim=0)

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.set_num_threads(os.cpu_count())
o = oc.CaptionOracle(oc.ModelSpec())
x = oc.synthetic_embeddings(5000)[:rows]
X2 = (x2, x2); F1 = (r_fp16, r_fp16); FA1 = (r_fp16, f2); FW1 = (f2, r_fp16); EX=(r_id, r_id)
schemes = {
  "ref": (dict(qkv=EX, proj=EX, fc=EX, fc2=EX), r_id),
  "all_x2_kv16": (dict(qkv=X2, proj=X2, fc=X2, fc2=X2), r_fp16),
  "qkv_fp16": (dict(qkv=F1, proj=X2, fc=X2, fc2=X2), r_fp16),
  "qkv_fc_fp16": (dict(qkv=F1, proj=X2, fc=F1, fc2=X2), r_fp16),
  "qkv_fp16A_Wx2": (dict(qkv=FA1, proj=X2, fc=X2, fc2=X2), r_fp16),
  "fc2_fp16": (dict(qkv=X2, proj=X2, fc=X2, fc2=F1), r_fp16),
  "proj_fp16": (dict(qkv=X2, proj=F1, fc=X2, fc2=X2), r_fp16),
  "fc_fp16": (dict(qkv=X2, proj=X2, fc=F1, fc2=X2), r_fp16),
}
want = sys.argv[2].split(",") if len(sys.argv) > 2 else list(schemes)
ref = None
for name in ["ref"] + [n for n in want if n != "ref"]:
    ops, rkv = schemes[name]
    t0 = time.time()
    ids = generate(o, x, 30, ops, rkv)
    if name == "ref": ref = ids; print("ref", round(time.time()-t0,1), flush=True); continue
    same = (ids == ref).all(dim=1)
    print(json.dumps({"scheme": name, "rows": rows, "captions_identical": int(same.sum()), "match": round(float(same.float().mean()), 4), "seconds": round(time.time()-t0, 1)}), flush=True)
