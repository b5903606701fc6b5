"""Host-side mirror of the reference's model API (src/models.py) with `generate` running on the B200 engine.

Same class names, constructor arguments, method signatures, return types and state_dict keys as the reference
(`mapping_network.*`, `gpt.*`, `task_prefix_embeds`, `aggregator.*`), so `src/eval.py`-style loops
(`model.to(device); model.eval(); model.generate(image_embeddings=..., max_length=..., temperature=..., top_p=...)`,
src/eval.py:189-210,281-289) run unchanged.  Parameters stay ordinary `nn.Parameter`s owned by PyTorch; the engine's
packed copies are derived from them and re-derived when they change (load_saved_parameters, .to(), training steps).

Only `generate` (greedy / sampling / beam) is on the CUDA path.  `forward` (teacher-forced training, src/models.py:237-325,
717-746) is plain differentiable PyTorch as in the reference and is outside the accelerated path (the RAT forward uses the
device kNN only to FETCH the retrieved rows; pooling goes through the `RetrievalAggregator` module so that its parameters
receive gradients).  There is no CPU fallback for `generate`.

The engine's packed weights are keyed on every parameter's `(data_ptr, _version)`; code that edits weights through
`p.data` (which does not bump `_version`) must call `model.invalidate_engine()` -- `load_saved_parameters`, `load_state_dict`
and `train()` do.
"""
from __future__ import annotations

import threading
from typing import Literal

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import CaptionEngine
from .inflight import current_slot

_ENGINE_LOCK = threading.Lock()

__all__ = ["MLPMappingNetwork", "TransformerMappingNetwork", "ImageCaptioningModel", "RetrievalAggregator",
           "RetrievalAugmentedTransformer", "accelerate"]


class MLPMappingNetwork(nn.Module):
    """image embedding [B,E] -> prefix tokens [B,P,d] through Linear -> activation -> Linear (src/models.py:14-74)."""

    def __init__(self, prefix_length: int = 10, embed_dim: int = 512, gpt_dim: int = 768, bias: bool = True,
                 activation: nn.Module = nn.Tanh()) -> None:
        super().__init__()
        self.prefix_length, self.embed_dim, self.gpt_dim = prefix_length, embed_dim, gpt_dim
        out_features = prefix_length * gpt_dim
        self.model = nn.Sequential(
            nn.Linear(embed_dim, out_features // 2, bias=bias),  # bottleneck = half of the output (src/models.py:50)
            activation,
            nn.Linear(out_features // 2, out_features, bias=bias),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x).view(x.shape[0], self.prefix_length, self.gpt_dim)


class TransformerMappingNetwork(nn.Module):
    """Linear to `hidden_length` image tokens + learned `prefix_const`, 8 pre-LN encoder layers, last P tokens out
    (src/models.py:77-174)."""

    def __init__(self, embed_dim: int, gpt_dim: int, prefix_length: int, hidden_length: int, num_layers: int = 8) -> None:
        super().__init__()
        self.embed_dim, self.gpt_dim = embed_dim, gpt_dim
        self.hidden_length, self.prefix_length = hidden_length, prefix_length
        self.linear = nn.Linear(embed_dim, hidden_length * gpt_dim)
        self.prefix_const = nn.Parameter(torch.randn(prefix_length, gpt_dim), requires_grad=True)
        layer = nn.TransformerEncoderLayer(d_model=gpt_dim, nhead=8, dim_feedforward=int(gpt_dim * 4), batch_first=True,
                                           activation="relu", norm_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        tokens = self.linear(x).view(x.shape[0], self.hidden_length, self.gpt_dim)
        learned = self.prefix_const.unsqueeze(0).expand(x.shape[0], -1, -1)
        return self.transformer(torch.cat((tokens, learned), dim=1))[:, self.hidden_length:, :]


class _EngineMixin:
    """Lazily (re)builds the CaptionEngine from the module's current parameters."""

    engine_dtype: str = "bf16"

    def _engine_key(self):
        params = list(self.gpt.parameters()) + list(self.mapping_network.parameters())
        if getattr(self, "task_prefix_embeds", None) is not None:
            params.append(self.task_prefix_embeds)
        return (self.engine_dtype, str(params[0].device), int(self.tokenizer.eos_token_id),
                tuple((p.data_ptr(), p._version) for p in params))

    def invalidate_engine(self) -> None:
        """Drop the packed weight copies; the next `generate` re-derives them from the module's current parameters.  Needed only
        after in-place edits through `p.data` (EMA, clipping, HF-style `_init_weights`), which the `(data_ptr, _version)` key
        cannot see."""
        with _ENGINE_LOCK:
            for old in self.__dict__.get("_engines", {}).values():
                old.close()
            self.__dict__.get("_engines", {}).clear()
            self.__dict__["_engine_key_cached"] = None

    def _get_engine(self) -> CaptionEngine:
        """The engine context of the calling thread's slot (inflight.current_slot(); slot 0 outside inflight.map_batches).  Slot 0
        owns the packed weights; every other slot is a context cloned from it (own stream, workspace, KV cache, CUDA graph, NO second
        weight copy), so batches in different slots can run concurrently."""
        key = self._engine_key()
        with _ENGINE_LOCK:
            engines = self.__dict__.setdefault("_engines", {})
            if self.__dict__.get("_engine_key_cached") != key:
                for old in engines.values():
                    old.close()
                engines.clear()
                self.__dict__["_engine_key_cached"] = key
            slot = current_slot()
            eng = engines.get(slot)
            if eng is None and slot != 0 and 0 in engines:
                eng = engines[slot] = engines[0].clone_context()
            if eng is None:
                mapper = self.mapping_network
                if isinstance(mapper, MLPMappingNetwork) or hasattr(mapper, "model"):
                    act = mapper.model[1]
                    if not isinstance(act, nn.Tanh):
                        raise NotImplementedError(f"the engine implements the reference's default Tanh mapper activation, not {type(act).__name__}")
                task = getattr(self, "task_prefix_embeds", None)
                eng = CaptionEngine(self.gpt, mapper, int(self.tokenizer.eos_token_id), task_prefix_embeds=task, dtype=self.engine_dtype)
                engines[0] = eng  # the weight owner
                if slot != 0:
                    eng = engines[slot] = eng.clone_context()
        return eng

    def _generate_on_engine(self, image_embeddings: torch.Tensor, max_length: int, temperature: float, top_p: float) -> torch.Tensor:
        self.eval()  # src/models.py:351
        device = image_embeddings.device
        batch = image_embeddings.shape[0]
        if temperature < 0:
            raise ValueError(f"temperature must be >= 0, got {temperature}")
        if max_length <= 0 or batch == 0:
            # max_length == 0 -> [B, 0] (src/models.py:471-473); an empty batch stops at once (`is_finished.all()` of an
            # empty tensor is True, :390) -> [0, 0]
            return torch.empty((batch, 0), dtype=torch.long, device=device)
        eng = self._get_engine()
        beams = int(getattr(self, "num_beams", 1) or 1)
        with torch.cuda.device(eng.device):
            if beams > 1 and temperature > 0:
                raise NotImplementedError("beam search is deterministic: use temperature=0.0 with num_beams > 1")
            if beams > 1:
                ids, _, gen_len = eng.generate_beam(image_embeddings, max_length, beams, float(getattr(self, "length_penalty", 1.0)))
                return ids[:, : int(gen_len.item())].to(device)  # HF crops to the longest selected hypothesis
            if temperature > 0:
                # temperature / top-p sampling (src/models.py:400-449); top_p >= 1 samples the full softmax (:407)
                ids, gen_len = eng.generate_sample(image_embeddings, max_length, float(temperature), float(top_p))
            else:
                ids, gen_len = eng.generate_greedy(image_embeddings, max_length)
            # one D2H read at the end of the loop (the reference syncs every step, src/models.py:390)
            n = int(gen_len.item())
        return ids[:, :n].to(device)


class ImageCaptioningModel(_EngineMixin, nn.Module):
    """Mapping network + GPT-2; same constructor as the reference (src/models.py:183-235) plus keyword-only engine knobs."""

    def __init__(self, mapping_network: nn.Module, image_prefix_length: int | None = None, prefix_task_prompt: str | None = None,
                 tokenizer=None, gpt=None, freeze_gpt_weights: bool = True, *, engine_dtype: str = "bf16", num_beams: int = 1,
                 length_penalty: float = 1.0) -> None:
        super().__init__()
        self.image_prefix_length = image_prefix_length or mapping_network.prefix_length
        self.mapping_network = mapping_network
        if gpt is None:  # same default as the reference (needs the HF hub / a local cache)
            from transformers import GPT2LMHeadModel
            gpt = GPT2LMHeadModel.from_pretrained("gpt2")
        self.gpt = gpt
        self.gpt_embedding_size = self.gpt.transformer.wte.weight.shape[1]
        if tokenizer is None:
            from transformers import GPT2Tokenizer
            tokenizer = GPT2Tokenizer.from_pretrained("gpt2")
            tokenizer.pad_token = tokenizer.eos_token
        self.tokenizer = tokenizer
        for p in self.gpt.parameters():
            p.requires_grad = not freeze_gpt_weights
        self.task_prefix_embeds: nn.Parameter | None = None
        if prefix_task_prompt:
            with torch.no_grad():
                ids = self.tokenizer.encode(prefix_task_prompt, return_tensors="pt")
                init = self.gpt.transformer.wte(ids.to(self.gpt.transformer.wte.weight.device)).squeeze(0)
            self.task_prefix_embeds = nn.Parameter(init.clone(), requires_grad=True)
        self.engine_dtype, self.num_beams, self.length_penalty = engine_dtype, num_beams, length_penalty

    # -- training forward: plain PyTorch, as in the reference (src/models.py:237-325) -----------------------------------
    def _prefix_tokens(self, image_embeddings: torch.Tensor) -> torch.Tensor:
        prefix = self.mapping_network(image_embeddings)
        if self.task_prefix_embeds is not None:  # task tokens go AFTER the image prefix (src/models.py:277-280)
            prefix = torch.cat((prefix, self.task_prefix_embeds.unsqueeze(0).expand(prefix.shape[0], -1, -1)), dim=1)
        return prefix

    def forward(self, caption_token_ids: torch.Tensor, image_embeddings: torch.Tensor, attention_mask: torch.Tensor | None = None,
                labels: torch.Tensor | None = None):
        prefix = self._prefix_tokens(image_embeddings)
        n_prefix = prefix.shape[1]
        inputs = torch.cat((prefix, self.gpt.transformer.wte(caption_token_ids)), dim=1)
        if labels is not None:  # no loss on the prefix positions
            labels = torch.cat((labels.new_full((labels.shape[0], n_prefix), -100), labels), dim=1)
        if attention_mask is not None:
            attention_mask = torch.cat((attention_mask.new_ones((attention_mask.shape[0], n_prefix)), attention_mask), dim=1)
        return self.gpt.forward(inputs_embeds=inputs, labels=labels, attention_mask=attention_mask)

    # -- the accelerated path ------------------------------------------------------------------------------------------------
    def generate(self, image_embeddings: torch.Tensor, max_length: int = 50, temperature: float = 1.0, top_p: float = 0.9) -> torch.Tensor:
        """int64 [B, L_gen] on `image_embeddings.device`; exact contract of src/models.py:327-477 for temperature == 0."""
        return self._generate_on_engine(image_embeddings, max_length, temperature, top_p)

    def generate_captions(self, image_embeddings: torch.Tensor, **kwargs) -> list[str]:
        return self.tokenizer.batch_decode(self.generate(image_embeddings, **kwargs), skip_special_tokens=True)

    # -- checkpoints (src/models.py:489-547) ---------------------------------------------------------------------------------
    def save_parameters(self, output_path: str) -> None:
        trainable = {n for n, p in self.named_parameters() if p.requires_grad}
        keep = {n: t for n, t in self.state_dict().items() if n in trainable or not n.startswith("gpt.")}
        print(f"Saving {len(keep)} trainable parameters and buffers to {output_path}.")
        torch.save(keep, output_path)

    def load_saved_parameters(self, checkpoint_path: str, device: torch.device | None = None) -> None:
        result = self.load_state_dict(torch.load(checkpoint_path, map_location=device), strict=False)
        if result.unexpected_keys:
            raise ValueError(f"Unexpected keys found in the checkpoint: {result.unexpected_keys}")
        missing = [k for k in result.missing_keys if not k.startswith("gpt.")]
        if missing:
            raise ValueError(f"Missing keys found in the checkpoint that are not from frozen GPT weights: {missing}")
        self.invalidate_engine()

    def train(self, mode: bool = True):
        if mode:  # weights are about to change; `generate` re-packs them on its next call
            self.invalidate_engine()
        return super().train(mode)


class RetrievalAggregator(nn.Module):
    """Pools the retrieved caption embeddings and adds them to the query (src/models.py:550-625)."""

    def __init__(self, embed_dim: int, aggregation_type: Literal["mean", "max", "sum_norm", "attention"] = "mean"):
        super().__init__()
        self.aggregation_type, self.embed_dim = aggregation_type, embed_dim
        if aggregation_type == "attention":
            self.attention_proj = nn.Linear(embed_dim, 1)

    def forward(self, query_embedding: torch.Tensor, retrieved_embeddings: torch.Tensor) -> torch.Tensor:
        kind = self.aggregation_type
        if kind == "mean":
            pooled = retrieved_embeddings.mean(dim=1)
        elif kind == "max":
            pooled = retrieved_embeddings.max(dim=1)[0]
        elif kind == "sum_norm":
            pooled = F.normalize(F.normalize(retrieved_embeddings, p=2, dim=2).sum(dim=1), p=2, dim=1)
        elif kind == "attention":
            weights = F.softmax(self.attention_proj(retrieved_embeddings), dim=1)
            pooled = (retrieved_embeddings * weights).sum(dim=1)
        else:
            raise ValueError(f"Unknown aggregation_type: {kind}")
        return query_embedding + pooled


class RetrievalAugmentedTransformer(ImageCaptioningModel):
    """RAT: kNN over the image matrix -> caption rows -> aggregate + add -> generate (src/models.py:628-785).
    `db_store` is a `gpt2_image_captioning_b200.database.GpuFlatStore` or any FAISS-flavoured store of the reference
    (duck type of src/models.py:673; uploaded to HBM once and cached, see `_device_store`)."""

    def __init__(self, embed_dim: int, max_workers: int = 4,
                 aggregation_type: Literal["mean", "max", "sum_norm", "attention"] = "mean", *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.max_workers = max_workers
        self.aggregator = RetrievalAggregator(embed_dim, aggregation_type)

    def _device_store(self, db_store):
        """The HBM-resident store behind `db_store`.  A `GpuFlatStore` is used as it is; any other FAISS-flavoured store --
        the reference dispatches on `hasattr(db_store, "image_index")` (src/models.py:673) and its callers pass
        `create_faiss_store(...)` straight in (src/eval.py:281-289) -- is uploaded ONCE (raw vectors of both indices + metadata,
        `GpuFlatStore.from_faiss_store`) and cached per store object."""
        if hasattr(db_store, "retrieve_and_aggregate"):
            return db_store
        if not hasattr(db_store, "image_index"):
            raise NotImplementedError("only FAISS-flavoured stores (objects with .image_index / .caption_index, src/models.py:673) run on "
                                      "the B200 path; the ObjectBox backend of the reference is out of scope (SURVEY.md section 2, row 6)")
        from .database import GpuFlatStore

        cache = self.__dict__.setdefault("_store_cache", {})
        hit = cache.get(id(db_store))
        if hit is not None and hit[0]() is db_store:
            return hit[1]
        import weakref

        dev = next(self.mapping_network.parameters()).device
        gpu_store = GpuFlatStore.from_faiss_store(db_store, device=dev)
        try:
            ref = weakref.ref(db_store, lambda _r, k=id(db_store): cache.pop(k, None))
        except TypeError:  # not weak-referenceable: keep it alive with the cache entry
            ref = (lambda obj: (lambda: obj))(db_store)
        cache[id(db_store)] = (ref, gpu_store)
        return gpu_store

    def _augment(self, db_store, image_embeddings: torch.Tensor, top_i: int, top_k: int) -> torch.Tensor:
        """Fused, non-differentiable retrieve + pool + add for `generate` (runs under no_grad on detached tensors)."""
        store = self._device_store(db_store)
        kind = self.aggregator.aggregation_type
        if kind == "attention":  # learned pooling (src/models.py:606-616), fused with the gather like the other modes
            proj = self.aggregator.attention_proj
            return store.retrieve_and_aggregate(image_embeddings, top_i=top_i, top_k=top_k, aggregation=kind,
                                                attention_weight=proj.weight, attention_bias=proj.bias)
        return store.retrieve_and_aggregate(image_embeddings, top_i=top_i, top_k=top_k, aggregation=kind)

    def _retrieve_batch(self, db_store, image_embeddings: torch.Tensor, top_i: int, top_k: int) -> torch.Tensor:
        """float32 [B, top_k, E] retrieved caption embeddings on the input's device (src/models.py:655-695), a constant w.r.t. autograd
        exactly as in the reference (which goes through numpy)."""
        store = self._device_store(db_store)
        with torch.no_grad():
            out = store.retrieve_caption_embeddings(image_embeddings.detach(), top_i=top_i, top_k=top_k)
        return out.to(image_embeddings.device)

    def forward(self, db_store, top_i: int, top_k: int, caption_token_ids: torch.Tensor, image_embeddings: torch.Tensor,
                attention_mask: torch.Tensor | None = None, labels: torch.Tensor | None = None):
        # training path (src/models.py:717-746): retrieval is a constant, pooling + add go through the differentiable module so
        # that `aggregator.attention_proj` (and anything upstream of image_embeddings) receives gradients
        retrieved = self._retrieve_batch(db_store, image_embeddings, top_i, top_k)
        augmented = self.aggregator(image_embeddings, retrieved)
        return super().forward(caption_token_ids=caption_token_ids, image_embeddings=augmented, attention_mask=attention_mask,
                               labels=labels)

    def generate(self, db_store, top_k: int, top_i: int, image_embeddings: torch.Tensor, max_length: int = 50,
                 temperature: float = 1.0, top_p: float = 0.9) -> torch.Tensor:
        augmented = self._augment(db_store, image_embeddings, top_i, top_k)
        return super().generate(augmented, max_length, temperature, top_p)

    def generate_captions(self, db_store, top_k: int, top_i: int, image_embeddings: torch.Tensor, **kwargs) -> list[str]:
        return self.tokenizer.batch_decode(self.generate(db_store, top_k, top_i, image_embeddings, **kwargs), skip_special_tokens=True)


def accelerate(model, engine_dtype: str = "bf16", num_beams: int = 1):
    """Drop-in for an EXISTING reference model instance (`src.models.ImageCaptioningModel`): replaces its bound
    `generate` with the engine-backed one, leaving parameters, `forward`, checkpoints and every caller untouched."""
    import types

    for attr in ("mapping_network", "gpt", "tokenizer"):
        if not hasattr(model, attr):
            raise TypeError(f"accelerate() expects an ImageCaptioningModel-like object with .{attr}")
    model.engine_dtype, model.num_beams = engine_dtype, num_beams
    for name in ("_engine_key", "_get_engine", "_generate_on_engine", "invalidate_engine"):
        setattr(model, name, types.MethodType(getattr(_EngineMixin, name), model))

    def generate(self, image_embeddings, max_length: int = 50, temperature: float = 1.0, top_p: float = 0.9):
        return self._generate_on_engine(image_embeddings, max_length, temperature, top_p)

    model.generate = types.MethodType(generate, model)
    return model
