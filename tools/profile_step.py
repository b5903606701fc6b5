#!/usr/bin/env python
"""One generate() of the headline workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py [--batch 1024] [--max-length 30] [--dtype bf16]
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> -c 3 -o gpurun_out/prof \
      python tools/profile_step.py

GIC_NO_GRAPH=1 is set so every kernel is an ordinary launch (the same kernels the CUDA graph replays).
"""
import argparse
import os
import sys

os.environ.setdefault("GIC_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--max-length", type=int, default=30)
ap.add_argument("--dtype", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
model = bench.build_product_model(a.dtype, dev)
eng = model._get_engine()
x = bench.synthetic_pool(bench.POOL_ROWS, bench.E)[: a.batch].to(dev)
eng.generate_greedy(x, a.max_length)  # warm-up (allocations, module load)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ids, n = eng.generate_greedy(x, a.max_length)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("generated", tuple(ids.shape), "gen_len", int(n.item()))
