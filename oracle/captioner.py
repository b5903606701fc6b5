"""CPU restatement of the reference caption-generation path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference (or third-party) lines it restates.  Paths are
relative to /root/reference unless prefixed `HF:` (transformers, pinned 4.57.3 in
the reference's uv.lock:3069-3070; 5.5.0 in this image -- same GPT-2 math) or
`torch:` (pinned 2.9.1, uv.lock:2903-2904; 2.11.0 here).

Two flavours of the same algorithm live here:
  * `backend="hf"`      -- the reference's generate loop restated around the very
                           third-party module the reference calls
                           (`GPT2LMHeadModel.forward(inputs_embeds=)`); this is what
                           bench.py times as the CPU baseline ("port").
  * `backend="restated"`-- GPT-2 / mapper arithmetic written out with plain tensor
                           ops (no nn.Module, no HF), with an optional KV-cached
                           variant used only to extend parity sets to sizes the
                           quadratic loop cannot reach in seconds.
Pinned by tests/test_oracle.py against tests/golden/*.npz, which were produced by
the UNMODIFIED reference (tests/golden/make_golden.py via oracle/ref_harness.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn as nn

EOS_TOKEN_ID = 50256  # GPT-2 <|endoftext|>; tokenizer.eos_token_id, src/models.py:348
VOCAB = 50257


# ----------------------------------------------------------------------------------------
# Model specification and the pinned synthetic inputs
# ----------------------------------------------------------------------------------------
GPT2_SIZES = {
    "small": dict(n_embd=768, n_layer=12, n_head=12),
    "medium": dict(n_embd=1024, n_layer=24, n_head=16),
    "large": dict(n_embd=1280, n_layer=36, n_head=20),
    # tiny shapes for fast unit tests only (not a reference configuration)
    "tiny": dict(n_embd=128, n_layer=2, n_head=2),
}


@dataclass(frozen=True)
class ModelSpec:
    gpt: str = "small"
    mapper: str = "mlp"  # "mlp" | "transformer"
    embed_dim: int = 512
    prefix_length: int = 10
    hidden_length: int = 10  # transformer mapper only (config.yml:19)
    mapper_layers: int = 8  # transformer mapper only (src/models.py:102)
    seed: int = 0

    @property
    def dims(self) -> dict:
        return GPT2_SIZES[self.gpt]


class MLPMapperModule(nn.Module):
    """Parameter container with the same construction order / state_dict keys as
    MLPMappingNetwork (src/models.py:23-56): Linear(E, P*d//2) -> Tanh -> Linear(P*d//2, P*d)."""

    def __init__(self, prefix_length: int, embed_dim: int, gpt_dim: int):
        super().__init__()
        self.prefix_length, self.embed_dim, self.gpt_dim = prefix_length, embed_dim, gpt_dim
        out = prefix_length * gpt_dim
        self.model = nn.Sequential(nn.Linear(embed_dim, out // 2), nn.Tanh(), nn.Linear(out // 2, out))

    def forward(self, x):  # src/models.py:58-74
        return self.model(x).view(x.shape[0], self.prefix_length, self.gpt_dim)


class TransformerMapperModule(nn.Module):
    """Same construction order / keys as TransformerMappingNetwork (src/models.py:96-139)."""

    def __init__(self, embed_dim: int, gpt_dim: int, prefix_length: int, hidden_length: int, num_layers: int = 8):
        super().__init__()
        self.embed_dim, self.gpt_dim = embed_dim, gpt_dim
        self.hidden_length, self.prefix_length = hidden_length, prefix_length
        self.linear = nn.Linear(embed_dim, hidden_length * gpt_dim)
        self.prefix_const = nn.Parameter(torch.randn(prefix_length, gpt_dim))
        layer = nn.TransformerEncoderLayer(
            d_model=gpt_dim, nhead=8, dim_feedforward=int(gpt_dim * 4), batch_first=True,
            activation="relu", norm_first=True,
        )
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)

    def forward(self, x):  # src/models.py:141-174
        b = x.shape[0]
        t = self.linear(x).view(b, self.hidden_length, self.gpt_dim)
        seq = torch.cat((t, self.prefix_const.unsqueeze(0).expand(b, -1, -1)), dim=1)
        return self.transformer(seq)[:, self.hidden_length:, :]


def build_modules(spec: ModelSpec):
    """Pinned weights: torch.manual_seed(seed); GPT2LMHeadModel(GPT2Config(...)) THEN the mapper,
    fp32 on CPU.  Returns (gpt, mapper) in eval mode."""
    from transformers import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(spec.seed)
    gpt = GPT2LMHeadModel(GPT2Config(**spec.dims))
    d = spec.dims["n_embd"]
    if spec.mapper == "mlp":
        mapper = MLPMapperModule(spec.prefix_length, spec.embed_dim, d)
    elif spec.mapper == "transformer":
        mapper = TransformerMapperModule(spec.embed_dim, d, spec.prefix_length, spec.hidden_length, spec.mapper_layers)
    else:
        raise ValueError(spec.mapper)
    return gpt.eval(), mapper.eval()


def synthetic_embeddings(n: int, embed_dim: int = 512, seed: int = 1) -> torch.Tensor:
    """randn rows, L2-normalised like the extractors (src/embeddings/clip.py:135-137)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, embed_dim, generator=g)
    return x / x.norm(dim=-1, keepdim=True)


def weight_fingerprint(gpt, mapper) -> dict:
    """Cheap float64 checksums so a box whose RNG stream differs fails loudly, not subtly."""
    sd = {**{"gpt." + k: v for k, v in gpt.state_dict().items()},
          **{"mapping_network." + k: v for k, v in mapper.state_dict().items()}}
    tot = 0.0
    for k in sorted(sd):
        if sd[k].dtype.is_floating_point:
            tot += float(sd[k].double().abs().sum())
    wte = gpt.transformer.wte.weight.detach()
    return {"abs_sum": tot, "wte_0_0": float(wte[0, 0]), "wte_last": float(wte[-1, -1])}


# ----------------------------------------------------------------------------------------
# Restated arithmetic (plain tensor ops)
# ----------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm, eps 1e-5 (HF:models/gpt2/configuration_gpt2.py layer_norm_epsilon; biased variance)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """NewGELUActivation (HF:activations.py:66): tanh form."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def gpt2_weights(gpt) -> dict:
    """Flat dict of the HF GPT-2 parameters.  Conv1D weights are [in, out] (HF:pytorch_utils.py:97-123)."""
    sd = gpt.state_dict()
    L = gpt.config.n_layer
    w = {"wte": sd["transformer.wte.weight"], "wpe": sd["transformer.wpe.weight"],
         "lnf_w": sd["transformer.ln_f.weight"], "lnf_b": sd["transformer.ln_f.bias"],
         "n_layer": L, "n_head": gpt.config.n_head, "layers": []}
    for i in range(L):
        p = f"transformer.h.{i}."
        w["layers"].append({
            "ln1_w": sd[p + "ln_1.weight"], "ln1_b": sd[p + "ln_1.bias"],
            "attn_w": sd[p + "attn.c_attn.weight"], "attn_b": sd[p + "attn.c_attn.bias"],
            "proj_w": sd[p + "attn.c_proj.weight"], "proj_b": sd[p + "attn.c_proj.bias"],
            "ln2_w": sd[p + "ln_2.weight"], "ln2_b": sd[p + "ln_2.bias"],
            "fc_w": sd[p + "mlp.c_fc.weight"], "fc_b": sd[p + "mlp.c_fc.bias"],
            "fc2_w": sd[p + "mlp.c_proj.weight"], "fc2_b": sd[p + "mlp.c_proj.bias"],
        })
    return w


def _identity(t: torch.Tensor) -> torch.Tensor:
    return t


def round_bf16(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bfloat16 and back: the storage rounding of GIC_DTYPE_BF16 operands."""
    return t.bfloat16().float()


def gpt2_forward(w: dict, inputs_embeds: torch.Tensor, kv: list | None = None, last_only: bool = False, rnd=_identity,
                 rnd_w=None, rnd_kv=None, rnd_head=None):
    """GPT2LMHeadModel.forward(inputs_embeds=) restated (HF:models/gpt2/modeling_gpt2.py:658-726,
    GPT2Model.forward :522-636, GPT2Block :262-309, GPT2Attention :144-226, GPT2MLP :238-243).

    h = x + wpe[pos]; per layer: a = LN1(h); qkv = a.Wqkv + b; causal softmax(q.k^T/sqrt(64)).v per head;
    h += attn.Wo + b; m = LN2(h); h += gelu_new(m.Wfc + b).Wproj + b; logits = LNf(h).wte^T (tied, no bias).
    `kv` (list of per-layer [k, v]) turns on the incremental form: positions continue after the cache
    (the reference itself never passes past_key_values; used only to extend parity sets).
    `rnd` (identity by default) is applied wherever the CUDA engine's bf16 mode stores a GEMM operand or the KV
    cache in bfloat16 (LayerNorm outputs, q/k/v, attention output, GELU output, every weight matrix): with
    rnd=round_bf16 this is a CPU emulation of that mode (fp32 accumulation, residual stream and logits), used to
    separate "bf16 rounding" from "kernel bug" in the parity tests.  `rnd_w` / `rnd_kv` (default: same as
    `rnd`) give the weight matrices and the q/k/v (KV-cache) stores their own rounding, so that mixed
    schemes (fp16 activations with split weights, fp32 KV ...) can be screened on the CPU (`rnd_kv` may also be a
    (q, k, v) tuple of roundings); `rnd_head` = (activation,
    weight) roundings of the LM head alone (default: those of the body).
    Returns fp32 logits [B, T, V] (or [B, 1, V] if last_only)."""
    B, T, d = inputs_embeds.shape
    H = w["n_head"]
    hd = d // H
    rw = rnd if rnd_w is None else rnd_w
    rkv = rnd if rnd_kv is None else rnd_kv
    past = 0 if not kv or kv[0] is None else kv[0][0].shape[2]
    pos = torch.arange(past, past + T)
    h = inputs_embeds + w["wpe"][pos]  # HF :579-585 -- prefix tokens also get wpe
    for li, lw in enumerate(w["layers"]):
        a = rnd(layer_norm(h, lw["ln1_w"], lw["ln1_b"]))
        qkv = a @ rw(lw["attn_w"]) + lw["attn_b"]  # Conv1D = addmm(bias, x, W[in,out])
        if isinstance(rkv, tuple):  # (q, k, v) roundings screened one at a time
            q, k, v = (r(t) for r, t in zip(rkv, qkv.split(d, dim=-1)))
        else:
            q, k, v = rkv(qkv).split(d, dim=-1)
        q = q.view(B, T, H, hd).transpose(1, 2)
        k = k.view(B, T, H, hd).transpose(1, 2)
        v = v.view(B, T, H, hd).transpose(1, 2)
        if kv is not None:
            if kv[li] is not None:
                k = torch.cat((kv[li][0], k), dim=2)
                v = torch.cat((kv[li][1], v), dim=2)
            kv[li] = [k, v]
        S = k.shape[2]
        att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        causal = torch.ones(S, S, dtype=torch.bool).tril()[S - T:, :]
        att = att.masked_fill(~causal, float("-inf")).softmax(-1)
        o = rnd((att @ v).transpose(1, 2).reshape(B, T, d))
        h = h + (o @ rw(lw["proj_w"]) + lw["proj_b"])
        m = rnd(layer_norm(h, lw["ln2_w"], lw["ln2_b"]))
        h = h + (rnd(gelu_new(m @ rw(lw["fc_w"]) + lw["fc_b"])) @ rw(lw["fc2_w"]) + lw["fc2_b"])
    if last_only:
        h = h[:, -1:, :]
    rha, rhw = (rnd, rw) if rnd_head is None else rnd_head
    h = rha(layer_norm(h, w["lnf_w"], w["lnf_b"]))
    return h @ rhw(w["wte"]).t()


def mlp_mapper(mw: dict, x: torch.Tensor, prefix_length: int, rnd=_identity) -> torch.Tensor:
    """MLPMappingNetwork.forward (src/models.py:58-74): view(tanh(x W1^T + b1) W2^T + b2, [B,P,d])."""
    h = rnd(torch.tanh(rnd(x) @ rnd(mw["model.0.weight"]).t() + mw["model.0.bias"]))
    y = h @ rnd(mw["model.2.weight"]).t() + mw["model.2.bias"]
    return y.view(x.shape[0], prefix_length, -1)


def transformer_mapper(mw: dict, x: torch.Tensor, prefix_length: int, hidden_length: int,
                       num_layers: int, nhead: int = 8) -> torch.Tensor:
    """TransformerMappingNetwork.forward (src/models.py:141-174).  Layer math is the pre-LN encoder layer
    (torch:nn/modules/transformer.py:946-950): x += MHA(LN1(x)); x += W2.relu(W1.LN2(x)); MHA uses the
    packed in_proj_weight [3d,d] in q,k,v order, scale 1/sqrt(d/8), no mask; no final norm."""
    B = x.shape[0]
    d = mw["prefix_const"].shape[1]
    hd = d // nhead
    t = (x @ mw["linear.weight"].t() + mw["linear.bias"]).view(B, hidden_length, d)
    h = torch.cat((t, mw["prefix_const"].unsqueeze(0).expand(B, -1, -1)), dim=1)
    S = h.shape[1]
    for i in range(num_layers):
        p = f"transformer.layers.{i}."
        a = layer_norm(h, mw[p + "norm1.weight"], mw[p + "norm1.bias"])
        qkv = a @ mw[p + "self_attn.in_proj_weight"].t() + mw[p + "self_attn.in_proj_bias"]
        q, k, v = qkv.split(d, dim=-1)
        q = q.view(B, S, nhead, hd).transpose(1, 2)
        k = k.view(B, S, nhead, hd).transpose(1, 2)
        v = v.view(B, S, nhead, hd).transpose(1, 2)
        att = ((q @ k.transpose(-1, -2)) / math.sqrt(hd)).softmax(-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, d)
        h = h + (o @ mw[p + "self_attn.out_proj.weight"].t() + mw[p + "self_attn.out_proj.bias"])
        m = layer_norm(h, mw[p + "norm2.weight"], mw[p + "norm2.bias"])
        f = torch.relu(m @ mw[p + "linear1.weight"].t() + mw[p + "linear1.bias"])
        h = h + (f @ mw[p + "linear2.weight"].t() + mw[p + "linear2.bias"])
    return h[:, hidden_length:, :]


def trim_length(ids: torch.Tensor, max_length: int, eos: int = EOS_TOKEN_ID) -> int:
    """L_gen of src/models.py:389-391,453-463: the loop stops BEFORE a step once every row has emitted
    EOS, so L_gen = N if some row never emits EOS within N steps, else max_b(first EOS index)+1."""
    if max_length == 0 or ids.numel() == 0:
        return 0
    is_eos = ids == eos
    if not bool(is_eos.any(dim=1).all()):
        return max_length
    first = is_eos.float().argmax(dim=1)
    return min(max_length, int(first.max()) + 1)


# ----------------------------------------------------------------------------------------
# The oracle object
# ----------------------------------------------------------------------------------------
@dataclass
class CaptionOracle:
    spec: ModelSpec
    gpt: nn.Module = field(default=None, repr=False)
    mapper: nn.Module = field(default=None, repr=False)
    task_prefix_embeds: torch.Tensor | None = None  # [Tp, d], appended AFTER the image prefix (src/models.py:364-375)

    def __post_init__(self):
        if self.gpt is None:
            self.gpt, self.mapper = build_modules(self.spec)
        self.w = gpt2_weights(self.gpt)
        self.mw = {k: v for k, v in self.mapper.state_dict().items()}

    # -- mapper ---------------------------------------------------------------------------
    @torch.no_grad()
    def prefix(self, x: torch.Tensor, backend: str = "restated", rnd=_identity) -> torch.Tensor:
        if backend == "hf":
            p = self.mapper(x)
        elif self.spec.mapper == "mlp":
            p = mlp_mapper(self.mw, x, self.spec.prefix_length, rnd)
        else:
            p = transformer_mapper(self.mw, x, self.spec.prefix_length, self.spec.hidden_length, self.spec.mapper_layers)
        if self.task_prefix_embeds is not None:
            p = torch.cat((p, self.task_prefix_embeds.unsqueeze(0).expand(x.shape[0], -1, -1)), dim=1)
        return p

    # -- greedy generate (src/models.py:327-477, temperature == 0 branch) -----------------------
    @torch.no_grad()
    def generate(self, x: torch.Tensor, max_length: int = 30, backend: str = "restated", kv_cache: bool = False,
                 return_logits: bool = False, emulate_bf16: bool = False):
        """Token ids int64 [B, L_gen].  `kv_cache=False` is the reference's algorithm verbatim: re-forward the
        whole growing sequence every step (src/models.py:395,466-469), logits of ALL positions computed
        (HF :705-706) and only the last used (:398); argmax ties -> lowest index; rows that have emitted EOS
        keep emitting EOS (:453-460); stop before a step once all rows are finished (:390-391)."""
        B = x.shape[0]
        rnd = round_bf16 if emulate_bf16 else _identity
        if emulate_bf16 and (backend != "restated" or self.spec.mapper != "mlp"):
            raise ValueError("bf16 emulation is implemented for the restated backend with the MLP mapper")
        wte_in = rnd(self.w["wte"])  # the bf16 engine gathers next-token embeddings from its bf16 table
        cur = self.prefix(x, backend, rnd)
        finished = torch.zeros(B, dtype=torch.bool)
        toks, logits_log = [], []
        kv = [None] * self.w["n_layer"] if kv_cache else None
        step_in = cur
        for _ in range(max_length):
            if bool(finished.all()):
                break
            if backend == "hf":
                if kv_cache:
                    raise ValueError("the reference never uses past_key_values; hf backend is cache-less")
                logits = self.gpt(inputs_embeds=cur).logits[:, -1, :]
            elif kv_cache:
                logits = gpt2_forward(self.w, step_in, kv=kv, last_only=True, rnd=rnd)[:, -1, :]
            else:
                logits = gpt2_forward(self.w, cur, rnd=rnd)[:, -1, :]
            if return_logits:
                logits_log.append(logits.clone())
            nxt = torch.argmax(logits / 1.0, dim=-1)  # :401-403 divides by 1.0 when temperature == 0
            finished = finished | nxt.eq(EOS_TOKEN_ID)
            nxt = torch.where(finished, torch.full_like(nxt, EOS_TOKEN_ID), nxt)
            toks.append(nxt.unsqueeze(-1))
            step_in = wte_in[nxt].unsqueeze(1)  # :466 wte lookup of the new token
            cur = torch.cat((cur, step_in), dim=1)
        ids = torch.cat(toks, dim=1) if toks else torch.empty((B, 0), dtype=torch.long)
        return (ids, logits_log) if return_logits else ids

    # -- beam search: NOT in the reference; oracle = HF GenerationMixin ---------
    @torch.no_grad()
    def generate_beam(self, x: torch.Tensor, max_length: int = 30, num_beams: int = 5) -> torch.Tensor:
        p = self.prefix(x)
        out = self.gpt.generate(
            inputs_embeds=p, num_beams=num_beams, do_sample=False, max_new_tokens=max_length,
            early_stopping=False, length_penalty=1.0, num_return_sequences=1,
            eos_token_id=EOS_TOKEN_ID, pad_token_id=EOS_TOKEN_ID,
        )
        return out


# ----------------------------------------------------------------------------------------
# Retrieval (src/database/faiss_store.py:132-251 + src/models.py:589-625,655-695)
# ----------------------------------------------------------------------------------------
def flat_ip_search(db: np.ndarray, q: np.ndarray, k: int):
    """faiss.IndexFlatIP.search semantics (faiss-cpu 1.13.1): exact q @ db^T, k best per row in
    descending score order, ids int64, -1 / -inf padding when k > n.  Ties: LOWEST index first
    (faiss's own tie order is unspecified; this is the rule the CUDA kernel implements)."""
    s = q.astype(np.float32) @ db.astype(np.float32).T
    n = db.shape[0]
    kk = min(k, n)
    order = np.lexsort((np.broadcast_to(np.arange(n), s.shape), -s), axis=1)[:, :kk]
    scores = np.take_along_axis(s, order, axis=1)
    if kk < k:
        order = np.concatenate([order, -np.ones((q.shape[0], k - kk), np.int64)], axis=1)
        scores = np.concatenate([scores, np.full((q.shape[0], k - kk), -np.inf, np.float32)], axis=1)
    return scores.astype(np.float32), order.astype(np.int64)


def retrieve_caption_rows(scores: np.ndarray, idx: np.ndarray, caption_rows_of_image, top_i: int, top_k: int):
    """The Python filter of faiss_store.py:160-183 and the row selection of :208-229, on integer ids:
    skip idx == -1 and score > 0.9999; keep the first top_i image hits; concatenate each hit's caption rows in
    hit order, stopping once >= top_k; keep the first top_k.  Returns int64 [B, top_k], -1 = zero padding row."""
    B = scores.shape[0]
    out = -np.ones((B, top_k), np.int64)
    for b in range(B):
        hits = []
        for s, i in zip(scores[b], idx[b]):
            if i == -1 or float(s) > 0.9999:
                continue
            hits.append(int(i))
            if len(hits) >= top_i:
                break
        rows: list[int] = []
        for i in hits:
            rows.extend(caption_rows_of_image(i))
            if len(rows) >= top_k:
                break
        rows = rows[:top_k]
        out[b, : len(rows)] = rows
    return out


def retrieve_and_aggregate(img_db: np.ndarray, cap_db: np.ndarray, caption_rows_of_image, q: np.ndarray,
                           top_i: int, top_k: int, aggregation: str = "mean") -> np.ndarray:
    """_retrieve_batch + RetrievalAggregator (src/models.py:655-695,589-625): search top_i+10 over the IMAGE
    matrix (faiss_store.py:153-155), gather caption rows (zero rows for padding, :222-226,241-244) and
    combine: mean over all top_k rows INCLUDING padding (:591), then query + aggregated (:623)."""
    scores, idx = flat_ip_search(img_db, q, top_i + 10)
    rows = retrieve_caption_rows(scores, idx, caption_rows_of_image, top_i, top_k)
    ret = np.zeros((q.shape[0], top_k, cap_db.shape[1]), np.float32)
    m = rows >= 0
    ret[m] = cap_db[rows[m]]
    if aggregation == "mean":
        agg = ret.mean(axis=1)
    elif aggregation == "max":
        agg = ret.max(axis=1)
    elif aggregation == "sum_norm":
        n = np.maximum(np.linalg.norm(ret, axis=2, keepdims=True), 1e-12)
        s = (ret / n).sum(axis=1)
        agg = s / np.maximum(np.linalg.norm(s, axis=1, keepdims=True), 1e-12)
    else:
        raise ValueError(aggregation)
    return (q + agg).astype(np.float32), rows
