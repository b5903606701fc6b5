"""Experiment: n batches of 1024 in flight per GPU (one engine + stream + host thread each) against the sequential loop bench.py
times.  Same model / batches as bench.py's device-resident leg.   python tools/two_in_flight.py [--engines 2] [--steps 10]"""
import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from gpt2_image_captioning_b200 import CaptionEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engines", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B, N, K = args.batch, 30, args.steps
    model = bench.build_product_model("bf16", dev)
    engines = [model._get_engine()]
    for _ in range(args.engines - 1):
        engines.append(CaptionEngine(model.gpt, model.mapping_network, int(model.tokenizer.eos_token_id), dtype="bf16"))
    pool = bench.synthetic_pool(bench.POOL_ROWS, bench.E).to(dev)
    batches = [pool[(torch.arange(B, device=dev) + i * B) % bench.POOL_ROWS].contiguous() for i in range(K)]
    want = [engines[0].generate_greedy(b, N)[0].clone() for b in batches]
    torch.cuda.synchronize()

    def sequential():
        t0 = time.perf_counter()
        for b in batches:
            engines[0].generate_greedy(b, N)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    outs = [None] * K

    def worker(j, stream):
        torch.cuda.set_device(dev)
        with torch.cuda.stream(stream):
            for i in range(j, K, len(engines)):
                outs[i] = engines[j].generate_greedy(batches[i], N)[0]
        stream.synchronize()

    def concurrent():
        streams = [torch.cuda.Stream(dev) for _ in engines]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts = [threading.Thread(target=worker, args=(j, s)) for j, s in enumerate(streams)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    for name, fn in (("sequential", sequential), (f"{len(engines)} in flight", concurrent)):
        fn()
        best = min(fn() for _ in range(3))
        print(f"{name:>14}: {best * 1e3 / K:7.2f} ms per batch of {B}   {B * K / best:9.0f} captions/s", flush=True)
    same = all(torch.equal(o, w) for o, w in zip(outs, want))
    print("concurrent ids identical to sequential:", same)


if __name__ == "__main__":
    main()
