"""Per-step logits of the bf16 engine (eager, logits tap) against the fp32 oracle: fused-LayerNorm vs unfused."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import golden_util as gu, gpu_util

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_mlp_eos"
g = gu.load(name)
model, oracle, x = gpu_util.product_model(g, "bf16")
n = min(8, x.shape[0]); steps = 6
ids_ref, logs_ref = oracle.generate(x[:n], steps, kv_cache=True, return_logits=True)
eng = model._get_engine()
ids, _, logits = eng.generate_greedy(x[:n].to("cuda:0"), steps, return_logits=True)
logits = logits.cpu()
print("fuse disabled" if os.environ.get("GIC_NO_LNFUSE") == "1" else "fused", name, "ids match", float((ids.cpu() == ids_ref).float().mean()))
for s in range(min(steps, len(logs_ref))):
    ref = logs_ref[s]; got = logits[s][:n]
    err = (got - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
    print(f"step {s}: max rel err per row:", " ".join(f"{v:.4f}" for v in err.tolist()))
