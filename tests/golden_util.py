"""Helpers shared by the parity tests: load a golden fixture and rebuild its pinned model/inputs."""
from __future__ import annotations

import functools
import os

import numpy as np
import torch

from oracle import captioner as oc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def have(name: str) -> bool:
    return os.path.isfile(os.path.join(GOLDEN_DIR, name + ".npz"))


def load(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def spec_of(g: dict) -> oc.ModelSpec:
    return oc.ModelSpec(
        gpt=str(g["spec_gpt"]), mapper=str(g["spec_mapper"]), embed_dim=int(g["spec_embed_dim"]),
        prefix_length=int(g["spec_prefix_length"]), hidden_length=int(g["spec_hidden_length"]),
        mapper_layers=int(g["spec_mapper_layers"]), seed=int(g["spec_seed"]),
    )


@functools.lru_cache(maxsize=4)
def _modules(spec: oc.ModelSpec, eos: int, eos_row_scale: float):
    gpt, mapper = oc.build_modules(spec)
    if eos_row_scale != 1.0:
        with torch.no_grad():
            gpt.transformer.wte.weight[eos] *= eos_row_scale
    return gpt, mapper


def rebuild(g: dict):
    """(spec, gpt, mapper, task_prefix|None, embeddings[n_rows]) exactly as make_golden.py built them."""
    spec = spec_of(g)
    eos = int(g.get("eos", oc.EOS_TOKEN_ID))
    scale = float(g.get("eos_row_scale", 1.0))
    gpt, mapper = _modules(spec, eos, scale)
    task = torch.from_numpy(g["task_prefix"]) if "task_prefix" in g else None
    n_rows = int(g["n_rows"])
    x = oc.synthetic_embeddings(int(g.get("emb_total", n_rows)), spec.embed_dim, int(g.get("emb_seed", 1)))[:n_rows]
    return spec, gpt, mapper, task, x


def check_fingerprint(g: dict, gpt, mapper):
    fp = oc.weight_fingerprint(gpt, mapper)
    for k, v in fp.items():
        ref = float(g["fp_" + k])
        assert abs(v - ref) <= 1e-9 * max(1.0, abs(ref)), (
            f"weight fingerprint {k} differs ({v} vs {ref}): this box's torch RNG stream does not reproduce the pinned weights")


def golden_batches(g: dict):
    """Yield (row_slice, L_gen, ids[rows, L_gen]) per reference generate() call."""
    ids, lens, batch = g["ids"], g["batch_lens"], int(g["batch"])
    for bi, start in enumerate(range(0, int(g["n_rows"]), batch)):
        sl = slice(start, min(start + batch, int(g["n_rows"])))
        L = int(lens[bi])
        yield sl, L, ids[sl, :L].astype(np.int64)
