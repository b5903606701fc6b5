"""ctypes binding of libgic_b200.so (include/gic_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` / `python -m gpt2_image_captioning_b200.build`
(nvcc, sm_100a only).  There is NO fallback: if the library is missing or fails to load, importing
`lib()` raises, and every compute entry point fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgic_b200.so")

ABI_VERSION = 1
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, -1, -2, -3, -4
DTYPE_F32, DTYPE_BF16, DTYPE_BF16X2 = 0, 1, 2
MAPPER_MLP, MAPPER_TRANSFORMER = 0, 1
AGG = {"mean": 0, "max": 1, "sum_norm": 2}
DTYPES = {"fp32": DTYPE_F32, "float32": DTYPE_F32, "bf16": DTYPE_BF16, "bfloat16": DTYPE_BF16, "bf16x2": DTYPE_BF16X2}

_fp = C.c_void_p  # device pointers travel as integers


class GicError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgic_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "dtype", "n_embd", "n_layer", "n_head", "vocab_size", "n_positions", "mapper_kind", "embed_dim",
        "prefix_length", "hidden_length", "mapper_layers", "mapper_heads", "task_prefix_length", "eos_token_id")]


class Gpt2LayerWeights(C.Structure):
    _fields_ = [(n, _fp) for n in ("ln1_w", "ln1_b", "attn_w", "attn_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc_w", "fc_b",
                                   "fc2_w", "fc2_b")]


class Gpt2Weights(C.Structure):
    _fields_ = [("wte", _fp), ("wpe", _fp), ("lnf_w", _fp), ("lnf_b", _fp), ("layers", C.POINTER(Gpt2LayerWeights))]


class MlpMapperWeights(C.Structure):
    _fields_ = [(n, _fp) for n in ("w1", "b1", "w2", "b2")]


class TfmLayerWeights(C.Structure):
    _fields_ = [(n, _fp) for n in ("norm1_w", "norm1_b", "norm2_w", "norm2_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
                                   "lin1_w", "lin1_b", "lin2_w", "lin2_b")]


class TfmMapperWeights(C.Structure):
    _fields_ = [("linear_w", _fp), ("linear_b", _fp), ("prefix_const", _fp), ("layers", C.POINTER(TfmLayerWeights))]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int32), ("total_ms", C.c_float)]


# name -> (restype, argtypes); every symbol include/gic_b200.h declares
SIGNATURES = {
    "gic_last_error": (C.c_char_p, []),
    "gic_abi_version": (C.c_int, []),
    "gic_device_check": (C.c_int, []),
    "gic_engine_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "gic_engine_destroy": (C.c_int, [C.c_void_p]),
    "gic_engine_clone": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gic_engine_load_gpt2": (C.c_int, [C.c_void_p, C.POINTER(Gpt2Weights), C.c_void_p]),
    "gic_engine_load_mlp_mapper": (C.c_int, [C.c_void_p, C.POINTER(MlpMapperWeights), C.c_void_p]),
    "gic_engine_load_tfm_mapper": (C.c_int, [C.c_void_p, C.POINTER(TfmMapperWeights), C.c_void_p]),
    "gic_engine_load_task_prefix": (C.c_int, [C.c_void_p, _fp, C.c_void_p]),
    "gic_engine_weight_bytes": (C.c_size_t, [C.c_void_p]),
    "gic_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "gic_mapper_forward": (C.c_int, [C.c_void_p, _fp, C.c_int, _fp, _fp, C.c_size_t, C.c_void_p]),
    "gic_generate_greedy": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_size_t, C.c_void_p]),
    "gic_generate_sample": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_ulonglong, _fp, _fp, _fp, _fp, C.c_size_t, C.c_void_p]),
    "gic_generate_beam": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp, _fp, _fp, _fp, C.c_size_t, C.c_void_p]),
    "gic_kv_reorder": (C.c_int, [C.c_void_p, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gic_topk_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gic_topk_tc_supported": (C.c_int, [C.c_int, C.c_int]),
    "gic_topk_tc_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gic_pack_bf16x2": (C.c_int, [_fp, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gic_topk_ip_tc": (C.c_int, [_fp, _fp, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, C.c_void_p]),
    "gic_topk_ip": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, C.c_void_p]),
    "gic_select_caption_rows": (C.c_int, [_fp, _fp, C.c_int, C.c_int, _fp, _fp, C.c_int, C.c_int, _fp, C.c_void_p]),
    "gic_gather_caption_rows": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, C.c_void_p]),
    "gic_gather_attention_add": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_void_p]),
    "gic_gather_aggregate_add": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_void_p]),
    "gic_launch_count": (C.c_ulonglong, []),
    "gic_compaction_count": (C.c_ulonglong, []),
    "gic_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "gic_profile_read": (C.c_int, [C.c_void_p, C.POINTER(ProfileEntry), C.c_int, C.POINTER(C.c_int)]),
    "gic_test_gemm": (C.c_int, [C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gic_test_layernorm": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_void_p]),
    "gic_test_attn_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gic_test_attn_decode_beam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p]),
    "gic_test_attn_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gic_test_sample_top_p": (C.c_int, [_fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_ulonglong, C.c_int, _fp, C.c_void_p]),
    "gic_trace_install": (C.c_int, [C.c_void_p, C.c_uint]),
    "gic_bench_attn_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gic_test_ln_mlp": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libgic_b200.so (once).  Fails loudly -- there is no pure-PyTorch or CPU path behind it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  gpt2_image_captioning_b200 has no fallback implementation.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    got = L.gic_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libgic_b200.so ABI version {got} != binding {ABI_VERSION}; rebuild the library")
    _lib = L
    return L


def check(code: int) -> None:
    if code != OK:
        raise GicError(code, lib().gic_last_error().decode("utf-8", "replace"))
