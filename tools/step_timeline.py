#!/usr/bin/env python
"""In-situ timeline of one decode step inside the CUDA graph (no profiler: every block of every GEMM / decode-attention /
ln_f / finalize launch stamps %globaltimer after griddepcontrol.wait and at its end; a launch's record keeps the earliest begin and the
latest end, include/gic_b200.h gic_trace_install).

  python tools/step_timeline.py [--batch 1024] [--max-length 30] [--step 15] > profiles/<tag>_step_timeline.txt

For the chosen decode step prints, per kernel in launch order: begin offset, duration (first block past its wait .. last block's
end) and the GAP between the previous kernel's end and this kernel's begin -- the launch / dependency-resolution cost the
persistent-kernel plan of DESIGN.md section 8 is after.
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from gpt2_image_captioning_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--max-length", type=int, default=30)
ap.add_argument("--step", type=int, default=15)
ap.add_argument("--dtype", default="bf16x2")
a = ap.parse_args()
dev = torch.device("cuda:0")
model = bench.build_product_model(a.dtype, dev)
eng = model._get_engine()
x = bench.synthetic_pool(bench.POOL_ROWS, bench.E)[: a.batch].to(dev)
eng.generate_greedy(x, a.max_length)
torch.cuda.synchronize()
cap = 1 << 14
buf = torch.zeros(cap, 3, dtype=torch.int64, device=dev)
buf[:, 1] = -1  # begin = min over blocks: start from all ones
L = _capi.lib()
_capi.check(L.gic_trace_install(C.c_void_p(buf.data_ptr()), cap))
eng.generate_greedy(x, a.max_length)
torch.cuda.synchronize()
_capi.check(L.gic_trace_install(None, 0))
rec = buf.cpu()
rec = rec[rec[:, 2] > 0]
kinds = {1: "gemm", 2: "attn_decode", 3: "layernorm", 4: "finalize", 5: "attn_prefill", 6: "lm_head_rescore"}
rows = sorted(((int(r[1]), int(r[2]), int(r[0]) & 0xFF, int(r[0]) >> 8) for r in rec if int(r[2]) > 0), key=lambda t: t[0])
# decode steps are delimited by finalize records; step 0 = prefill + first token
fins = [i for i, r in enumerate(rows) if r[2] == 4]
if a.step >= len(fins):
    raise SystemExit(f"only {len(fins)} steps traced")
lo, hi = fins[a.step - 1] + 1, fins[a.step] + 1
sel = rows[lo:hi]
t0 = sel[0][0]
print(f"# decode step {a.step} of {len(fins) - 1} (B = {a.batch}, GPT-2 small {a.dtype}): {len(sel)} launches, "
      f"{(sel[-1][1] - t0) / 1e3:.1f} us from the first kernel's begin to the last kernel's end")
print(f"# {'kernel':28s} {'begin_us':>9s} {'kernel_us':>9s} {'gap_us':>7s}")
tot_gap = tot_life = 0.0
prev_end = None
for b, e_, k, d in sel:
    name = kinds.get(k, str(k))
    if k == 1:
        name = f"gemm bn={d >> 4} epi={(d >> 1) & 7}{' pair' if d & 1 else ''}"
    gap = (b - prev_end) / 1e3 if prev_end is not None else 0.0
    life = (e_ - b) / 1e3
    tot_gap += max(gap, 0.0)
    tot_life += life
    print(f"  {name:28s} {(b - t0) / 1e3:9.2f} {life:9.2f} {gap:7.2f}")
    prev_end = e_
print(f"# sum of kernel durations {tot_life:.1f} us, sum of gaps {tot_gap:.1f} us")
