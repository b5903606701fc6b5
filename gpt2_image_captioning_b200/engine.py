"""CaptionEngine: Python owner of one `gic_engine` handle (packed weights) and its workspace.

PyTorch owns the parameters (`model.gpt`, `model.mapping_network`); the engine holds packed copies derived from them
and is rebuilt when they change (checkpoints are loaded into ordinary nn.Parameters).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi, ops


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _check_supported(gcfg, mapping_network) -> None:
    """The kernels implement exactly the arithmetic of the reference's models (HF GPT-2 defaults, HF:models/gpt2/configuration_gpt2.py;
    nn.TransformerEncoderLayer as src/models.py:129-139 builds it).  Anything else would produce wrong tokens silently, so it is refused."""
    bad = []
    if getattr(gcfg, "activation_function", "gelu_new") != "gelu_new":
        bad.append(f"activation_function={gcfg.activation_function!r} (kernels: gelu_new)")
    if abs(float(getattr(gcfg, "layer_norm_epsilon", 1e-5)) - 1e-5) > 1e-12:
        bad.append(f"layer_norm_epsilon={gcfg.layer_norm_epsilon} (kernels: 1e-5)")
    if not getattr(gcfg, "scale_attn_weights", True):
        bad.append("scale_attn_weights=False (kernels scale by 1/sqrt(head_dim))")
    if getattr(gcfg, "scale_attn_by_inverse_layer_idx", False):
        bad.append("scale_attn_by_inverse_layer_idx=True")
    if getattr(gcfg, "reorder_and_upcast_attn", False):
        bad.append("reorder_and_upcast_attn=True")
    if getattr(gcfg, "add_cross_attention", False):
        bad.append("add_cross_attention=True")
    layers = getattr(getattr(mapping_network, "transformer", None), "layers", None)
    if layers is not None:
        for i, lyr in enumerate(layers):
            if not getattr(lyr, "norm_first", False):
                bad.append(f"mapper layer {i}: norm_first=False (kernels: pre-LN, src/models.py:135)")
            act = getattr(lyr, "activation", None)
            if act is not torch.nn.functional.relu and getattr(act, "__name__", "") != "relu" and not isinstance(act, torch.nn.ReLU):
                bad.append(f"mapper layer {i}: activation {act!r} (kernels: relu)")
            for nm in ("norm1", "norm2"):
                if abs(float(getattr(lyr, nm).eps) - 1e-5) > 1e-12:
                    bad.append(f"mapper layer {i}: {nm}.eps={getattr(lyr, nm).eps} (kernels: 1e-5)")
            if not getattr(lyr.self_attn, "batch_first", True):
                bad.append(f"mapper layer {i}: batch_first=False")
        if getattr(mapping_network.transformer, "norm", None) is not None:
            bad.append("mapper: final encoder norm (the reference has none)")
    if bad:
        raise NotImplementedError("the B200 engine does not implement this model variant: " + "; ".join(bad))


class CaptionEngine:
    def __init__(self, gpt, mapping_network, eos_token_id: int, task_prefix_embeds: torch.Tensor | None = None,
                 dtype: str = "bf16", device: torch.device | str | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("CaptionEngine needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _capi.lib()
        self.device = torch.device(device) if device is not None else next(gpt.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError(f"model parameters live on {self.device}; move the model to a CUDA device first (model.to('cuda'))")
        if dtype not in _capi.DTYPES:
            raise ValueError(f"unknown engine dtype {dtype!r}; choose from {sorted(_capi.DTYPES)}")
        self.dtype = dtype
        self._handle = C.c_void_p()
        self._ws: torch.Tensor | None = None
        gcfg = gpt.config
        _check_supported(gcfg, mapping_network)
        sd = {k: v for k, v in gpt.state_dict().items()}
        msd = {k: v for k, v in mapping_network.state_dict().items()}
        is_tfm = "prefix_const" in msd
        d = int(gcfg.n_embd)
        cfg = _capi.Config()
        cfg.abi_version = _capi.ABI_VERSION
        cfg.dtype = _capi.DTYPES[dtype]
        cfg.n_embd, cfg.n_layer, cfg.n_head = d, int(gcfg.n_layer), int(gcfg.n_head)
        cfg.vocab_size, cfg.n_positions = int(sd["transformer.wte.weight"].shape[0]), int(sd["transformer.wpe.weight"].shape[0])
        cfg.mapper_kind = _capi.MAPPER_TRANSFORMER if is_tfm else _capi.MAPPER_MLP
        cfg.prefix_length = int(mapping_network.prefix_length)
        cfg.task_prefix_length = 0 if task_prefix_embeds is None else int(task_prefix_embeds.shape[0])
        cfg.eos_token_id = int(eos_token_id)
        if is_tfm:
            cfg.embed_dim = int(msd["linear.weight"].shape[1])
            cfg.hidden_length = int(mapping_network.hidden_length)
            cfg.mapper_layers = len(mapping_network.transformer.layers)
            cfg.mapper_heads = int(mapping_network.transformer.layers[0].self_attn.num_heads)
        else:
            cfg.embed_dim = int(msd["model.0.weight"].shape[1])
            if "model.0.bias" not in msd:
                raise NotImplementedError("MLP mapper without bias is not supported by the engine")
        self.cfg = cfg
        self.embed_dim, self.gpt_dim, self.vocab = cfg.embed_dim, d, cfg.vocab_size
        self.prefix_total = cfg.prefix_length + cfg.task_prefix_length
        with torch.cuda.device(self.device):
            _capi.check(self.lib.gic_engine_create(C.byref(cfg), C.byref(self._handle)))
            keep = []  # fp32 staging tensors must outlive the asynchronous pack kernels
            stream = torch.cuda.current_stream().cuda_stream

            def p(t):
                t = _f32(t, self.device)
                keep.append(t)
                return t.data_ptr()

            layers = (_capi.Gpt2LayerWeights * cfg.n_layer)()
            for i in range(cfg.n_layer):
                pre = f"transformer.h.{i}."
                lw = layers[i]
                lw.ln1_w, lw.ln1_b = p(sd[pre + "ln_1.weight"]), p(sd[pre + "ln_1.bias"])
                lw.attn_w, lw.attn_b = p(sd[pre + "attn.c_attn.weight"]), p(sd[pre + "attn.c_attn.bias"])
                lw.proj_w, lw.proj_b = p(sd[pre + "attn.c_proj.weight"]), p(sd[pre + "attn.c_proj.bias"])
                lw.ln2_w, lw.ln2_b = p(sd[pre + "ln_2.weight"]), p(sd[pre + "ln_2.bias"])
                lw.fc_w, lw.fc_b = p(sd[pre + "mlp.c_fc.weight"]), p(sd[pre + "mlp.c_fc.bias"])
                lw.fc2_w, lw.fc2_b = p(sd[pre + "mlp.c_proj.weight"]), p(sd[pre + "mlp.c_proj.bias"])
            gw = _capi.Gpt2Weights()
            gw.wte, gw.wpe = p(sd["transformer.wte.weight"]), p(sd["transformer.wpe.weight"])
            gw.lnf_w, gw.lnf_b = p(sd["transformer.ln_f.weight"]), p(sd["transformer.ln_f.bias"])
            gw.layers = layers
            _capi.check(self.lib.gic_engine_load_gpt2(self._handle, C.byref(gw), stream))
            if is_tfm:
                tl = (_capi.TfmLayerWeights * cfg.mapper_layers)()
                for i in range(cfg.mapper_layers):
                    pre = f"transformer.layers.{i}."
                    t = tl[i]
                    t.norm1_w, t.norm1_b = p(msd[pre + "norm1.weight"]), p(msd[pre + "norm1.bias"])
                    t.norm2_w, t.norm2_b = p(msd[pre + "norm2.weight"]), p(msd[pre + "norm2.bias"])
                    t.in_proj_w, t.in_proj_b = p(msd[pre + "self_attn.in_proj_weight"]), p(msd[pre + "self_attn.in_proj_bias"])
                    t.out_proj_w, t.out_proj_b = p(msd[pre + "self_attn.out_proj.weight"]), p(msd[pre + "self_attn.out_proj.bias"])
                    t.lin1_w, t.lin1_b = p(msd[pre + "linear1.weight"]), p(msd[pre + "linear1.bias"])
                    t.lin2_w, t.lin2_b = p(msd[pre + "linear2.weight"]), p(msd[pre + "linear2.bias"])
                tw = _capi.TfmMapperWeights()
                tw.linear_w, tw.linear_b = p(msd["linear.weight"]), p(msd["linear.bias"])
                tw.prefix_const = p(msd["prefix_const"])
                tw.layers = tl
                _capi.check(self.lib.gic_engine_load_tfm_mapper(self._handle, C.byref(tw), stream))
            else:
                mw = _capi.MlpMapperWeights()
                mw.w1, mw.b1 = p(msd["model.0.weight"]), p(msd["model.0.bias"])
                mw.w2, mw.b2 = p(msd["model.2.weight"]), p(msd["model.2.bias"])
                _capi.check(self.lib.gic_engine_load_mlp_mapper(self._handle, C.byref(mw), stream))
            if task_prefix_embeds is not None:
                _capi.check(self.lib.gic_engine_load_task_prefix(self._handle, p(task_prefix_embeds), stream))
            torch.cuda.current_stream().synchronize()
            del keep

    def clone_context(self) -> "CaptionEngine":
        """A second context on THIS engine's packed weights (own stream, workspace, KV cache, CUDA graph): what an extra in-flight
        slot uses (inflight.py).  Holds a reference to its parent so the weights outlive it."""
        c = object.__new__(CaptionEngine)
        c.lib, c.device, c.dtype, c.cfg = self.lib, self.device, self.dtype, self.cfg
        c.embed_dim, c.gpt_dim, c.vocab, c.prefix_total = self.embed_dim, self.gpt_dim, self.vocab, self.prefix_total
        c._ws = None
        c._parent = self
        c._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.gic_engine_clone(self._handle, C.byref(c._handle)))
        self.__dict__.setdefault("_children", []).append(c)
        return c

    # ------------------------------------------------------------------------------------------------------------
    @property
    def handle(self) -> int:
        return self._handle.value

    def weight_bytes(self) -> int:
        return int(self.lib.gic_engine_weight_bytes(self._handle))

    def workspace(self, batch: int, max_new: int, beams: int = 1) -> torch.Tensor:
        need = int(self.lib.gic_workspace_bytes(self._handle, batch, max_new, beams))
        if need == 0:
            raise ValueError(f"bad workspace request batch={batch} max_new={max_new} beams={beams}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- measurement hooks (bench.py) ----------------------------------------------------------------------------------
    def profile(self, on: bool) -> None:
        """While on, generate runs without the CUDA graph with CUDA events around every kernel class."""
        _capi.check(self.lib.gic_profile_enable(self._handle, 1 if on else 0))

    def profile_read(self) -> dict[str, dict]:
        buf = (_capi.ProfileEntry * 64)()
        n = C.c_int(0)
        _capi.check(self.lib.gic_profile_read(self._handle, buf, 64, C.byref(n)))
        return {buf[i].name.decode(): {"launches": int(buf[i].launches), "total_ms": float(buf[i].total_ms)} for i in range(n.value)}

    @staticmethod
    def launch_count() -> int:
        return int(_capi.lib().gic_launch_count())

    @staticmethod
    def compaction_count() -> int:
        return int(_capi.lib().gic_compaction_count())

    def close(self) -> None:
        for child in self.__dict__.pop("_children", []):  # contexts borrowing these weights go first
            child.close()
        if self._handle:
            self.lib.gic_engine_destroy(self._handle)
            self._handle = C.c_void_p()
        self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------
    def _check_x(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2 or x.shape[1] != self.embed_dim:
            raise ValueError(f"image_embeddings must be [batch, {self.embed_dim}], got {tuple(x.shape)}")
        return x.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()

    def mapper_forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._check_x(x)
        B = x.shape[0]
        out = torch.empty(B, self.prefix_total, self.gpt_dim, dtype=torch.float32, device=self.device)
        if B == 0:
            return out
        ops.mapper_forward(self.handle, x, out, self.workspace(B, 0))
        return out

    def generate_greedy(self, x: torch.Tensor, max_new_tokens: int, return_logits: bool = False):
        """ids int64 [B, max_new_tokens] (full length), gen_len int32 [1] (device) [, logits fp32 [max_new, B, V]]."""
        x = self._check_x(x)
        B = x.shape[0]
        ids = torch.empty(B, max_new_tokens, dtype=torch.int64, device=self.device)
        gen_len = torch.zeros(1, dtype=torch.int32, device=self.device)
        logits = torch.empty(max_new_tokens, B, self.vocab, dtype=torch.float32, device=self.device) if return_logits else None
        if B > 0 and max_new_tokens > 0:
            ops.generate_greedy(self.handle, x, max_new_tokens, ids, gen_len, logits, self.workspace(B, max_new_tokens))
        return (ids, gen_len, logits) if return_logits else (ids, gen_len)

    def generate_sample(self, x: torch.Tensor, max_new_tokens: int, temperature: float = 1.0, top_p: float = 0.9, seed: int | None = None):
        """Temperature / nucleus sampling (src/models.py:400-449): ids int64 [B, max_new_tokens], gen_len int32 [1].  `seed` keys the
        device Philox stream; by default it is drawn from torch's global CPU generator, so torch.manual_seed makes runs repeatable."""
        x = self._check_x(x)
        B = x.shape[0]
        ids = torch.empty(B, max_new_tokens, dtype=torch.int64, device=self.device)
        gen_len = torch.zeros(1, dtype=torch.int32, device=self.device)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if B > 0 and max_new_tokens > 0:
            # one step's logits; kept with the engine: the captured decode graphs hold its address
            scratch = getattr(self, "_sample_scratch", None)
            ld = (self.vocab + 31) // 32 * 32  # rows padded for aligned vector stores / bulk copies (include/gic_b200.h)
            if scratch is None or scratch.numel() < B * ld:
                self._sample_scratch = scratch = torch.empty(B * ld, dtype=torch.float32, device=self.device)
            ops.generate_sample(self.handle, x, max_new_tokens, float(temperature), float(top_p), int(seed), ids, gen_len, scratch,
                                self.workspace(B, max_new_tokens))
        return ids, gen_len

    def generate_beam(self, x: torch.Tensor, max_new_tokens: int, num_beams: int = 5, length_penalty: float = 1.0):
        """ids int64 [B, max_new_tokens] (eos padded), scores fp32 [B], gen_len int32 [1] = longest selected hypothesis."""
        x = self._check_x(x)
        B = x.shape[0]
        ids = torch.empty(B, max_new_tokens, dtype=torch.int64, device=self.device)
        scores = torch.empty(B, dtype=torch.float32, device=self.device)
        gen_len = torch.zeros(1, dtype=torch.int32, device=self.device)
        if B > 0 and max_new_tokens > 0:
            ops.generate_beam(self.handle, x, max_new_tokens, num_beams, length_penalty, ids, scores, gen_len,
                              self.workspace(B, max_new_tokens, num_beams))
        return ids, scores, gen_len
