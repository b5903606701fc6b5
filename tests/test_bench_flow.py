"""bench.py's control flow on the CPU with a FAKE engine: the phases, the phase-by-phase assembly of the JSON line and the watchdog.

No number produced here means anything (the fake engine returns zeros and canned timings) -- the point is that the driver's
`python bench.py` cannot die of a KeyError / NameError in code paths that only run on a GPU box, that exactly one JSON line comes out
with every key of the contract, and that a phase which hangs costs its keys, not the line.  The real thing is `-m gpu` + `bench.py`
on a B200.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = textwrap.dedent('''
    import json, os, sys, time, types
    sys.path.insert(0, %(root)r)
    import torch
    import bench

    HANG = os.environ.get("FAKE_HANG", "")

    # ---- CUDA plumbing that bench.py touches, as CPU no-ops ----
    class FakeEvent:
        def __init__(self, enable_timing=False): self.t = 0.0
        def record(self): self.t = time.perf_counter()
        def elapsed_time(self, other): return max(1e-3, (other.t - self.t) * 1e3)
    torch.cuda.Event = FakeEvent
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    _to = torch.Tensor.to
    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda") and not (isinstance(x, str) and x.startswith("cuda")))
        k = {n: v for n, v in k.items() if not (n == "device" and str(v).startswith("cuda"))}
        return _to(self, *a, **k) if (a or k) else self
    torch.Tensor.to = to

    # ---- the product model without the 124M-parameter GPT-2 and without the CUDA engine ----
    class FakeEngine:
        launches = 0
        def generate_greedy(self, x, n):
            FakeEngine.launches += 65 * n
            time.sleep(0.002)
            return torch.zeros(x.shape[0], n, dtype=torch.int64), torch.tensor([n], dtype=torch.int32)
        def launch_count(self): return FakeEngine.launches
        def weight_bytes(self): return 123
        def profile(self, on): pass
        def profile_read(self):
            return {k: {"launches": 3, "total_ms": 1.5} for k in ("mapper", "prefill_gemm", "attn_prefill", "gemm_qkv", "attn_decode", "lm_head", "finalize")}
    class FakeModel:
        engine_dtype = "bf16x2"
        class tokenizer: eos_token_id = 50256
        def __init__(self, dtype): self.engine_dtype = dtype; self._engines = {0: FakeEngine(), 1: FakeEngine()}
        def _get_engine(self): return self._engines[0]
        def invalidate_engine(self): pass
        def to(self, d): return self
        def eval(self): return self
        def generate(self, image_embeddings=None, max_length=30, temperature=1.0, top_p=0.9, **kw):
            if HANG == "jobs" and image_embeddings.shape[0] != 1024 and temperature == 0.0:
                time.sleep(3600)
            return torch.zeros(image_embeddings.shape[0], max_length, dtype=torch.int64)
    bench.build_product_model = lambda dtype, device, *a, **k: FakeModel(dtype)
    canned = json.load(open(os.path.join(%(root)r, "profiles", "r2ac_bench.json")))["in_graph"]
    def fake_timeline(eng, x, N, B, dtype, pk):
        if HANG == "timeline":
            time.sleep(3600)
        return dict(canned)
    bench.in_graph_timeline = fake_timeline
    bench.cpu_baseline_sample = lambda rows, max_length: {"value": 1.0, "unit": "captions/s", "cores": 1, "kind": "port", "sample": "fake"}
    if HANG:  # shrink the budgets so the test takes seconds
        real_phase = bench.Watchdog.phase
        bench.Watchdog.phase = lambda self, name, budget_s: real_phase(self, name, 4.0)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # the ranks' plumbing over gloo instead of NCCL
        import torch.distributed as dist
        real_init = dist.init_process_group
        dist.init_process_group = lambda backend, device_id=None, **k: real_init("gloo", **k)
        real_tensor = torch.tensor
        torch.tensor = lambda *a, **k: real_tensor(*a, **{n: v for n, v in k.items() if n != "device"})
    sys.argv = ["bench.py", "--gpus", os.environ.get("WORLD_SIZE", "1"), "--steps", "2", "--warmup", "1", "--no-c4-job"]
    bench.main()
''')


def _run(hang: str = ""):
    env = dict(os.environ, FAKE_HANG=hang, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, "-c", HARNESS % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    return r


def test_bench_prints_one_complete_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "roofline_attention", "decode_step_roofline", "frac_of_max_bound", "in_graph",
                "modes", "job", "sampling", "cpu_baseline", "parity"):
        assert key in line, key
    assert "incomplete" not in line
    assert set(line["modes"]) == {"bf16", "bf16x2", "fp32"} and all("parity" in m for m in line["modes"].values())
    assert line["e2e"]["h2d_bytes_per_step"] == 1024 * 512 * 4 and line["gpu_launches"] > 0
    assert line["roofline"]["kernel"].startswith("gemm_bf16_tcgen05_kernel") and line["roofline"]["traffic"] is not None
    assert line["job"]["c2_5000_rows_gpt2_small"]["rows"] == 5000


def test_a_hanging_phase_costs_its_keys_not_the_line():
    for hang, missing, kept in (("timeline", ("roofline", "modes", "job"), ("value", "e2e", "clocks")),
                                ("jobs", ("job",), ("value", "e2e", "roofline", "modes", "sampling"))):
        r = _run(hang)
        assert r.returncode == 0, (hang, r.stderr[-3000:])
        lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1, (hang, r.stdout[-2000:])
        line = json.loads(lines[0])
        assert "incomplete" in line and "did not finish" in line["incomplete"], hang
        for k in missing:
            assert k not in line, (hang, k)
        for k in kept:
            assert k in line, (hang, k)


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 8])
def test_ranks_one_line_and_a_clean_exit(world):
    """The multi-rank flow (barriers, max over ranks, the sharded job with its gloo gather over the loopback interface, rank 0 prints,
    the other rank prints nothing, the process group is torn down) with gloo standing in for NCCL."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, FAKE_HANG="", RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", HARNESS % {"root": ROOT}], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                      env=env, cwd=ROOT))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
    assert all(o[0].strip() == "" for o in outs[1:])
    lines = [ln for ln in outs[0][0].splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == world and "incomplete" not in line and "cpu_baseline" not in line
    job = line["job"]["c2_5000_rows_gpt2_small"]
    assert job["n_gpus"] == world and job["ids_shape"] == [5000, 30] and job["gathered_ids_equal_independent_generation"] is True
