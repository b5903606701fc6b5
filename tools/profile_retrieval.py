#!/usr/bin/env python
"""The configs[4] retrieval scan (1024 queries, top-5 over 591 753 x 512) once between cudaProfilerStart/Stop, for an ncu launch list:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/<tag>_retrieval_launches.csv \
      python tools/profile_retrieval.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from gpt2_image_captioning_b200.database import _GpuFlatIndex  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(2)
cap = torch.randn(591753, 512, generator=g)
cap /= cap.norm(dim=-1, keepdim=True)
index = _GpuFlatIndex(cap.to(dev))
x = bench.synthetic_pool(1024, 512).to(dev)
for _ in range(2):
    s, i = index.search_device(x, 5)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    s, i = index.search_device(x, 5)
b.record()
torch.cuda.synchronize()
print("top-5 over 591753 x 512 for 1024 queries: %.3f ms per call" % (a.elapsed_time(b) / 5))
torch.cuda.profiler.start()
s, i = index.search_device(x, 5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(s.shape), tuple(i.shape))
