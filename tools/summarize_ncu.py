#!/usr/bin/env python
"""Turn the ncu outputs of tools/gpu_round.sh into the committed summaries under profiles/.

  python tools/summarize_ncu.py <tag>        # reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_step.ncu-rep

writes profiles/<tag>_launches_summary.txt  (per-kernel launch count, total / mean device time, SHARE of the pass)
       profiles/<tag>_step_metrics.csv      (one row per profiled launch of one decode step: duration, DRAM bytes,
                                             DRAM / L2 / tensor-pipe utilisation, registers, grid)
       profiles/<tag>_step_metrics.txt      (the same grouped by kernel class)
       profiles/ncu_traffic.json            (dram bytes per launch per kernel class -- bench.py's roofline.traffic)
ncu's per-launch times are cold-cache and serialised: compare shares, not absolutes (B200_PROFILING.md).
"""
from __future__ import annotations

import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    n = name.replace("void ", "").replace("gic::", "")
    return n.split("(")[0]


def classify(name: str, grid: str, counters: dict) -> str:
    """decode-step kernel class (matches bench.py kernel_classes): the 4 body GEMMs are told apart by launch order."""
    s = short(name)
    if s.startswith("gemm_bf16"):
        i = counters["gemm"] % 4 if counters["in_layers"] else -1
        counters["gemm"] += 1
        return ("gemm_qkv", "gemm_proj", "gemm_fc", "gemm_fc2")[i] if i >= 0 else "lm_head"
    for key in ("attn_decode", "attn_seq", "layernorm", "finalize", "kv_reorder", "lm_head_candidates", "lm_head_rescore_pairs"):
        if s.startswith(key):
            return key
    return s


def launches_summary(tag: str) -> None:
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if not os.path.isfile(path):
        print("no", path)
        return
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    out = [f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none, one generate() of the headline workload",
           f"# (tools/profile_step.py, GIC_NO_GRAPH=1: the same kernels the CUDA graph replays).  {len(rows)} launches, {total / 1e3:.1f} us total.",
           "# per-launch times are cold-cache and serialised: the SHARE column is what the bench's kernel classes must agree with.",
           f"{'kernel':58s} {'grid':>14s} {'block':>12s} {'launches':>8s} {'total_us':>10s} {'mean_us':>8s} {'share':>6s}"]
    for (k, g, b), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k[:58]:58s} {g:>14s} {b:>12s} {n:8d} {ns / 1e3:10.1f} {ns / 1e3 / n:8.2f} {100 * ns / total:5.1f}%")
    dst = os.path.join(ROOT, "profiles", f"{tag}_launches_summary.txt")
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


WANT = collections.OrderedDict([
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("dram__bytes.sum.per_second", "dram_GBps"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])
SCALE = {"Gbyte/s": 1.0, "Tbyte/s": 1e3, "Mbyte/s": 1e-3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def step_metrics(tag: str) -> None:
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_step.ncu-rep")
    if not os.path.isfile(rep):
        print("no", rep)
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    counters = {"gemm": 0, "in_layers": True}
    recs = []
    body = rows[2:]
    n_gemm = sum(1 for r in body if short(r[col["Kernel Name"]]).startswith("gemm_bf16"))
    for r in body:
        name = r[col["Kernel Name"]]
        counters["in_layers"] = counters["gemm"] < n_gemm - (n_gemm % 4)  # the trailing GEMM of a step is the LM head
        rec = {"class": classify(name, r[col.get("Grid Size", 0)], counters), "kernel": short(name)}
        for m, nice in WANT.items():
            if m not in col:
                rec[nice] = ""
                continue
            v = r[col[m]].replace(",", "")
            try:
                f = float(v)
            except ValueError:
                rec[nice] = v
                continue
            rec[nice] = f * SCALE.get(units[col[m]], 1.0)
        recs.append(rec)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    dst = os.path.join(ROOT, "profiles", f"{tag}_step_metrics.csv")
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["class", "kernel"] + list(WANT.values()))
        w.writeheader()
        for rec in recs:
            w.writerow({k: (f"{v:.3f}" if isinstance(v, float) else v) for k, v in rec.items()})
    print("wrote", dst)
    # grouped
    groups = collections.OrderedDict()
    for rec in recs:
        groups.setdefault(rec["class"], []).append(rec)
    tot = sum(r["dur_us"] for r in recs if isinstance(r["dur_us"], float))
    out = [f"# {tag}: ncu --set full --clock-control none over the last layer + head of one decode step (B=1024, GPT-2 small, {os.environ.get('DTYPE', 'bf16x2')}); means per kernel class.",
           "# dur = gpu__time_duration (cold-cache, serialised under the profiler: compare shares); dram = dram__bytes_read+write per launch;",
           "# l2_to_sm = l1tex__m_xbar2l1tex_read_bytes; pct columns are % of peak sustained over the kernel's elapsed time.",
           f"{'class':12s} {'n':>3s} {'dur_us':>8s} {'share':>6s} {'dram_MB':>9s} {'l2_to_sm_MB':>11s} {'dramGB/s':>8s} {'l2%':>6s} {'tensor%':>7s} {'sm%':>6s} {'regs':>5s} {'grid':>6s}"]
    traffic = {}

    def mean(rs, k):
        v = [r[k] for r in rs if isinstance(r[k], float)]
        return sum(v) / len(v) if v else float("nan")

    for cls, rs in groups.items():
        dram = mean(rs, "dram_rd_MB") + mean(rs, "dram_wr_MB")
        traffic[cls] = dram * 1e6
        share = sum(r["dur_us"] for r in rs if isinstance(r["dur_us"], float)) / tot
        out.append(f"{cls:12s} {len(rs):3d} {mean(rs, 'dur_us'):8.2f} {100 * share:5.1f}% {dram:9.2f} {mean(rs, 'l2_to_sm_MB'):11.2f} "
                   f"{mean(rs, 'dram_GBps'):8.0f} {mean(rs, 'l2_pct'):6.1f} {mean(rs, 'tensor_pct'):7.1f} {mean(rs, 'sm_pct'):6.1f} "
                   f"{mean(rs, 'regs'):5.0f} {mean(rs, 'grid'):6.0f}")
    out.append(f"# sum of profiled launch durations: {tot:.1f} us over {len(recs)} launches")
    dst = os.path.join(ROOT, "profiles", f"{tag}_step_metrics.txt")
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)
    # bench.py's roofline.traffic: dram bytes per launch per class; "gemm_body" = mean over the four body GEMMs of a layer
    body = [traffic[k] for k in ("gemm_qkv", "gemm_proj", "gemm_fc", "gemm_fc2") if k in traffic]
    if body:
        traffic["gemm_body"] = sum(body) / len(body)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    old = json.load(open(tpath)) if os.path.isfile(tpath) else {}
    dt = os.environ.get("DTYPE", "bf16x2")
    old.update({f"{k}_{dt}": v for k, v in traffic.items()})
    json.dump(old, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(out))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches_summary(tag)
    step_metrics(tag)
