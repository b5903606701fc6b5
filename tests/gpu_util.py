"""Shared helpers for the -m gpu parity tests: build the product model from a golden fixture's pinned weights."""
from __future__ import annotations

import numpy as np
import torch

import golden_util as gu
from oracle import captioner as oc
from oracle.ref_harness import StubTokenizer


def product_model(g: dict, dtype: str, device: str = "cuda:0"):
    """(model, oracle, x) -- `model` is the product ImageCaptioningModel carrying the fixture's exact weights."""
    from gpt2_image_captioning_b200 import ImageCaptioningModel, MLPMappingNetwork, TransformerMappingNetwork

    spec, gpt, mapper_ref, task, x = gu.rebuild(g)
    if "fp_abs_sum" in g:  # the beam fixtures carry no fingerprint of their own (same weights as the greedy ones)
        gu.check_fingerprint(g, gpt, mapper_ref)
    d = spec.dims["n_embd"]
    if spec.mapper == "mlp":
        mapper = MLPMappingNetwork(prefix_length=spec.prefix_length, embed_dim=spec.embed_dim, gpt_dim=d)
    else:
        mapper = TransformerMappingNetwork(embed_dim=spec.embed_dim, gpt_dim=d, prefix_length=spec.prefix_length,
                                           hidden_length=spec.hidden_length, num_layers=spec.mapper_layers)
    mapper.load_state_dict(mapper_ref.state_dict())
    tok = StubTokenizer()
    tok.eos_token_id = int(g.get("eos", oc.EOS_TOKEN_ID))
    import copy
    model = ImageCaptioningModel(mapper, tokenizer=tok, gpt=copy.deepcopy(gpt), engine_dtype=dtype)
    if task is not None:
        model.task_prefix_embeds = torch.nn.Parameter(task.clone())
    model = model.to(device).eval()
    oracle = oc.CaptionOracle(spec, gpt, mapper_ref, task_prefix_embeds=task)
    return model, oracle, x


def first_mismatch_audit(oracle, x_row: torch.Tensor, got: np.ndarray, want: np.ndarray, max_length: int):
    """For one row whose tokens differ from the reference: the reference's own top-2 logit gap at the first
    differing step (a near-tie there means the two fp32 summation orders legitimately disagree)."""
    n = min(len(got), len(want))
    diff = np.nonzero(got[:n] != want[:n])[0]
    if len(diff) == 0:
        return None
    step = int(diff[0])
    _, logs = oracle.generate(x_row.unsqueeze(0), step + 1, kv_cache=True, return_logits=True)
    top2 = torch.topk(logs[step][0], 2)
    return {"step": step, "gap": float(top2.values[0] - top2.values[1]), "ref_top2": top2.indices.tolist(), "got": int(got[step])}
