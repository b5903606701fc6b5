"""Per-kernel-class device time of one full-size batch of a secondary config (c3 | c4 | c5) via the engine's event profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools.bench_configs import build
dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"  # c3 | c4 | c5 | c2 (the headline model) ; a 4th argument "sample" profiles temperature 1.0 / top_p 0.9
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
DT = sys.argv[3] if len(sys.argv) > 3 else "bf16"
if cfg == "c3":
    model = build(dict(n_embd=1024, n_layer=24, n_head=16), "tfm", 512, 40, DT, dev, beams=5); E = 512
elif cfg == "c4":
    model = build(dict(n_embd=1280, n_layer=36, n_head=20), "mlp", 1024, 10, DT, dev); E = 1024
else:
    model = build(bench.MODEL, "mlp", 512, 10, DT, dev); E = 512
x = bench.synthetic_pool(B, E).to(dev)
kw = dict(temperature=1.0, top_p=0.9) if (len(sys.argv) > 4 and sys.argv[4] == "sample") else dict(temperature=0.0)
model.generate(image_embeddings=x, max_length=30, **kw)
eng = model._get_engine()
eng.profile(True)
model.generate(image_embeddings=x, max_length=30, **kw)
torch.cuda.synchronize()
prof = eng.profile_read()
eng.profile(False)
tot = sum(v["total_ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"{k:16s} launches {v['launches']:5d}  avg {1e3 * v['total_ms'] / max(1, v['launches']):9.1f} us  total {v['total_ms']:8.2f} ms  {100 * v['total_ms'] / tot:5.1f} %")
print("sum", tot, "ms")
