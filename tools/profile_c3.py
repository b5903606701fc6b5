"""Per-kernel-class device time of one full-size config-3 batch (GPT-2 medium, transformer mapper, P=40, beam 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools.bench_configs import build
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model = build(dict(n_embd=1024, n_layer=24, n_head=16), "tfm", 512, 40, "bf16", dev, beams=5)
x = bench.synthetic_pool(B, 512).to(dev)
model.generate(image_embeddings=x, max_length=30, temperature=0.0)
eng = model._get_engine()
eng.profile(True)
model.generate(image_embeddings=x, max_length=30, temperature=0.0)
torch.cuda.synchronize()
prof = eng.profile_read()
eng.profile(False)
tot = sum(v["total_ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"{k:16s} launches {v['launches']:5d}  avg {1e3 * v['total_ms'] / max(1, v['launches']):9.1f} us  total {v['total_ms']:8.2f} ms  {100 * v['total_ms'] / tot:5.1f} %")
print("sum", tot, "ms")
