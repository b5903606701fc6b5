#!/usr/bin/env python
"""EOS-realistic throughput: GPT-2 small + MLP mapper, batch 1024, max_length 50, with the tied EOS embedding row
scaled (the `eos_row_scale` trick of tests/golden/make_golden.py at full model size) so that rows stop at different lengths -- mean ~12
tokens, a tail to 50 -- the way a trained captioner does, instead of the random-init never-EOS behaviour of the headline benchmark.

Times `model.generate` (device-resident inputs, CUDA events, 3 repetitions) three ways: every row decoded to max_length
(GIC_NO_EARLY_EXIT=1), whole-batch early exit only (GIC_NO_COMPACT=1), and early exit + finished-row compaction (default); checks that
the three produce identical ids.  One JSON line per arithmetic mode.

  python tools/bench_eos.py [--scale 3.5] [--batch 1024] [--max-length 50] [--dtypes bf16x2,bf16]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=3.5)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--max-length", type=int, default=50)
    ap.add_argument("--dtypes", default="bf16x2,bf16")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    from gpt2_image_captioning_b200 import CaptionEngine
    x = bench.synthetic_pool(bench.POOL_ROWS, bench.E)[: a.batch].to(dev)
    for dtype in a.dtypes.split(","):
        res = {}
        ids_ref = None
        for name, env in (("decode_every_row_to_max_length", {"GIC_NO_EARLY_EXIT": "1"}), ("whole_batch_early_exit", {"GIC_NO_COMPACT": "1"}),
                          ("early_exit_and_compaction", {})):
            for k in ("GIC_NO_EARLY_EXIT", "GIC_NO_COMPACT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            model = bench.build_product_model(dtype, dev)
            with torch.no_grad():
                model.gpt.transformer.wte.weight[50256] *= a.scale
            model.invalidate_engine()
            for _ in range(2):
                ids = model.generate(image_embeddings=x, max_length=a.max_length, temperature=0.0)
            torch.cuda.synchronize()
            c0 = CaptionEngine.compaction_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ids = model.generate(image_embeddings=x, max_length=a.max_length, temperature=0.0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            if ids_ref is None:
                ids_ref = ids
            same = ids.shape[1] <= ids_ref.shape[1] and bool((ids == ids_ref[:, : ids.shape[1]]).all()) and bool((ids_ref[:, ids.shape[1]:] == 50256).all())
            res[name] = {"ms_per_batch": ms, "captions_per_s": a.batch / ms * 1e3, "L_gen": int(ids.shape[1]), "ids_equal_full_decode": same,
                         "compactions_per_batch": (CaptionEngine.compaction_count() - c0) / 3}
            model.invalidate_engine()
            del model
            torch.cuda.empty_cache()
        eos = ids_ref == 50256
        first = torch.where(eos.any(dim=1), eos.float().argmax(dim=1), torch.full((ids_ref.shape[0],), ids_ref.shape[1], device=ids_ref.device))
        print(json.dumps({"workload": f"GPT-2 small + MLP mapper, batch {a.batch}, max_length {a.max_length}, EOS row x{a.scale}", "dtype": dtype,
                          "first_eos_mean": float(first.float().mean()), "first_eos_p10_p50_p90": [float(v) for v in torch.quantile(first.float(), torch.tensor([0.1, 0.5, 0.9], device=first.device))],
                          "rows_never_finishing": int((first >= ids_ref.shape[1]).sum()), **res}), flush=True)


if __name__ == "__main__":
    main()
