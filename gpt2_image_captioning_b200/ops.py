"""torch custom ops (namespace `gic::`) over the C ABI: each op passes `tensor.data_ptr()` + the current CUDA stream
to libgic_b200.so.  PyTorch is plumbing here (device memory, streams); the arithmetic is in the CUDA library."""
from __future__ import annotations

import torch

from . import _capi


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gic ops need CUDA tensors: libgic_b200 has no CPU path")


@torch.library.custom_op("gic::mapper_forward", mutates_args=("prefix_out", "workspace"))
def mapper_forward(engine: int, image_embeddings: torch.Tensor, prefix_out: torch.Tensor, workspace: torch.Tensor) -> None:
    _need_cuda(image_embeddings, prefix_out, workspace)
    L = _capi.lib()
    _capi.check(L.gic_mapper_forward(engine, _ptr(image_embeddings), image_embeddings.shape[0], _ptr(prefix_out), _ptr(workspace),
                                     workspace.numel(), _stream()))


@torch.library.custom_op("gic::generate_greedy", mutates_args=("ids_out", "gen_len_out", "logits_out", "workspace"))
def generate_greedy(engine: int, image_embeddings: torch.Tensor, max_new_tokens: int, ids_out: torch.Tensor,
                    gen_len_out: torch.Tensor, logits_out: torch.Tensor | None, workspace: torch.Tensor) -> None:
    _need_cuda(image_embeddings, ids_out, gen_len_out, logits_out, workspace)
    L = _capi.lib()
    _capi.check(L.gic_generate_greedy(engine, _ptr(image_embeddings), image_embeddings.shape[0], max_new_tokens, _ptr(ids_out),
                                      _ptr(gen_len_out), _ptr(logits_out), _ptr(workspace), workspace.numel(), _stream()))


@torch.library.custom_op("gic::generate_sample", mutates_args=("ids_out", "gen_len_out", "logits_scratch", "workspace"))
def generate_sample(engine: int, image_embeddings: torch.Tensor, max_new_tokens: int, temperature: float, top_p: float, seed: int,
                    ids_out: torch.Tensor, gen_len_out: torch.Tensor, logits_scratch: torch.Tensor, workspace: torch.Tensor) -> None:
    _need_cuda(image_embeddings, ids_out, gen_len_out, logits_scratch, workspace)
    L = _capi.lib()
    _capi.check(L.gic_generate_sample(engine, _ptr(image_embeddings), image_embeddings.shape[0], max_new_tokens, float(temperature), float(top_p),
                                      int(seed), _ptr(ids_out), _ptr(gen_len_out), _ptr(logits_scratch), _ptr(workspace), workspace.numel(),
                                      _stream()))


@torch.library.custom_op("gic::test_sample_top_p", mutates_args=("tokens_out",))
def test_sample_top_p(logits: torch.Tensor, temperature: float, top_p: float, seed: int, step: int, tokens_out: torch.Tensor) -> None:
    _need_cuda(logits, tokens_out)
    L = _capi.lib()
    _capi.check(L.gic_test_sample_top_p(_ptr(logits), logits.shape[0], logits.shape[1], float(temperature), float(top_p), int(seed), int(step),
                                        _ptr(tokens_out), _stream()))


@torch.library.custom_op("gic::generate_beam", mutates_args=("ids_out", "scores_out", "gen_len_out", "workspace"))
def generate_beam(engine: int, image_embeddings: torch.Tensor, max_new_tokens: int, num_beams: int, length_penalty: float,
                  ids_out: torch.Tensor, scores_out: torch.Tensor, gen_len_out: torch.Tensor, workspace: torch.Tensor) -> None:
    _need_cuda(image_embeddings, ids_out, scores_out, gen_len_out, workspace)
    L = _capi.lib()
    _capi.check(L.gic_generate_beam(engine, _ptr(image_embeddings), image_embeddings.shape[0], max_new_tokens, num_beams,
                                    float(length_penalty), _ptr(ids_out), _ptr(scores_out), _ptr(gen_len_out), _ptr(workspace),
                                    workspace.numel(), _stream()))


@torch.library.custom_op("gic::kv_reorder", mutates_args=("kv_dst",))
def kv_reorder(engine: int, kv_src: torch.Tensor, kv_dst: torch.Tensor, beam_idx: torch.Tensor, ctx_len: int, t_max: int) -> None:
    _need_cuda(kv_src, kv_dst, beam_idx)
    L = _capi.lib()
    _capi.check(L.gic_kv_reorder(engine, _ptr(kv_src), _ptr(kv_dst), _ptr(beam_idx), beam_idx.numel(), ctx_len, t_max, _stream()))


@torch.library.custom_op("gic::topk_ip", mutates_args=("scores_out", "idx_out", "workspace"))
def topk_ip(queries: torch.Tensor, db: torch.Tensor, k: int, scores_out: torch.Tensor, idx_out: torch.Tensor,
            workspace: torch.Tensor) -> None:
    _need_cuda(queries, db, scores_out, idx_out, workspace)
    L = _capi.lib()
    _capi.check(L.gic_topk_ip(_ptr(queries), _ptr(db), queries.shape[0], db.shape[0], db.shape[1], k, _ptr(scores_out), _ptr(idx_out),
                              _ptr(workspace), workspace.numel(), _stream()))


@torch.library.custom_op("gic::pack_bf16x2", mutates_args=("hi", "lo"))
def pack_bf16x2(src: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor) -> None:
    """src fp32 -> hi = bf16(src), lo = bf16(src - hi) (same shape, bf16): the operands of the bf16x2 tensor-core mode."""
    _need_cuda(src, hi, lo)
    L = _capi.lib()
    _capi.check(L.gic_pack_bf16x2(_ptr(src), hi.data_ptr(), lo.data_ptr(), src.numel(), _stream()))


@torch.library.custom_op("gic::topk_ip_tc", mutates_args=("scores_out", "idx_out", "workspace"))
def topk_ip_tc(queries: torch.Tensor, db: torch.Tensor, db_hi: torch.Tensor, db_lo: torch.Tensor, db_norm_max: float, k: int,
               scores_out: torch.Tensor, idx_out: torch.Tensor, workspace: torch.Tensor) -> None:
    _need_cuda(queries, db, db_hi, db_lo, scores_out, idx_out, workspace)
    L = _capi.lib()
    _capi.check(L.gic_topk_ip_tc(_ptr(queries), _ptr(db), db_hi.data_ptr(), db_lo.data_ptr(), float(db_norm_max), queries.shape[0], db.shape[0],
                                 db.shape[1], k, _ptr(scores_out), _ptr(idx_out), _ptr(workspace), workspace.numel(), _stream()))


@torch.library.custom_op("gic::select_caption_rows", mutates_args=("rows_out",))
def select_caption_rows(scores: torch.Tensor, idx: torch.Tensor, cap_row_start: torch.Tensor, cap_row_ids: torch.Tensor | None,
                        top_i: int, top_k: int, rows_out: torch.Tensor) -> None:
    _need_cuda(scores, idx, cap_row_start, cap_row_ids, rows_out)
    L = _capi.lib()
    _capi.check(L.gic_select_caption_rows(_ptr(scores), _ptr(idx), scores.shape[0], scores.shape[1], _ptr(cap_row_start),
                                          _ptr(cap_row_ids), top_i, top_k, _ptr(rows_out), _stream()))


@torch.library.custom_op("gic::gather_caption_rows", mutates_args=("out",))
def gather_caption_rows(cap_db: torch.Tensor, rows: torch.Tensor, out: torch.Tensor) -> None:
    """out[b, j, :] = cap_db[rows[b, j]] (zero row for -1): fp32 [B, top_k, D]."""
    _need_cuda(cap_db, rows, out)
    L = _capi.lib()
    _capi.check(L.gic_gather_caption_rows(_ptr(cap_db), _ptr(rows), rows.shape[0], rows.shape[1], cap_db.shape[1], _ptr(out), _stream()))


@torch.library.custom_op("gic::gather_attention_add", mutates_args=("out",))
def gather_attention_add(queries: torch.Tensor, cap_db: torch.Tensor, rows: torch.Tensor, attn_w: torch.Tensor, attn_b: torch.Tensor,
                         out: torch.Tensor) -> None:
    _need_cuda(queries, cap_db, rows, attn_w, attn_b, out)
    L = _capi.lib()
    _capi.check(L.gic_gather_attention_add(_ptr(queries), _ptr(cap_db), _ptr(rows), queries.shape[0], rows.shape[1], queries.shape[1],
                                           _ptr(attn_w), _ptr(attn_b), _ptr(out), _stream()))


@torch.library.custom_op("gic::gather_aggregate_add", mutates_args=("out",))
def gather_aggregate_add(queries: torch.Tensor, cap_db: torch.Tensor, rows: torch.Tensor, aggregation: int, out: torch.Tensor) -> None:
    _need_cuda(queries, cap_db, rows, out)
    L = _capi.lib()
    _capi.check(L.gic_gather_aggregate_add(_ptr(queries), _ptr(cap_db), _ptr(rows), queries.shape[0], rows.shape[1], queries.shape[1],
                                           aggregation, _ptr(out), _stream()))


@torch.library.custom_op("gic::test_gemm", mutates_args=("C",))
def test_gemm(dtype: int, A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None, C: torch.Tensor, epilogue: int) -> None:
    _need_cuda(A, W, bias, C)
    L = _capi.lib()
    _capi.check(L.gic_test_gemm(dtype, _ptr(A), _ptr(W), _ptr(bias), _ptr(C), A.shape[0], W.shape[0], A.shape[1], epilogue, _stream()))


@torch.library.custom_op("gic::test_ln_mlp", mutates_args=("h", "hb_out", "stats_out"))
def test_ln_mlp(h: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, wfc: torch.Tensor, bfc: torch.Tensor, wfc2: torch.Tensor,
                bfc2: torch.Tensor, hb_out: torch.Tensor, stats_out: torch.Tensor, split_k: int) -> None:
    _need_cuda(h, gamma, beta, wfc, bfc, wfc2, bfc2, hb_out, stats_out)
    L = _capi.lib()
    _capi.check(L.gic_test_ln_mlp(_ptr(h), _ptr(gamma), _ptr(beta), _ptr(wfc), _ptr(bfc), _ptr(wfc2), _ptr(bfc2), _ptr(hb_out),
                                  _ptr(stats_out), h.shape[0], h.shape[1], split_k, _stream()))


@torch.library.custom_op("gic::test_attn_decode", mutates_args=("kcache", "vcache", "out"))
def test_attn_decode(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, out: torch.Tensor, pos: int, variant: int) -> None:
    """qkv [rows, 3*H*64] bf16; kcache / vcache [rows, H, t_max, 64] bf16; out [rows, H*64] bf16."""
    _need_cuda(qkv, kcache, vcache, out)
    L = _capi.lib()
    rows, H, t_max = kcache.shape[0], kcache.shape[1], kcache.shape[2]
    _capi.check(L.gic_test_attn_decode(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), out.data_ptr(), pos, rows, H, t_max, variant, _stream()))


@torch.library.custom_op("gic::test_attn_decode_beam", mutates_args=("kcache", "vcache", "out", "out_lo"))
def test_attn_decode_beam(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, out: torch.Tensor, out_lo: torch.Tensor, anc: torch.Tensor, pos: int,
                          n_prefix: int, beams: int, f16: bool, shared: bool) -> None:
    """Beam-search decode attention through an ancestry table.  qkv [rows, 3*H*64], kcache / vcache [rows, H, t_max, 64], out (and, f16, out_lo)
    [rows, H*64], all 2-byte elements (bf16, or IEEE half viewed as bf16 when f16); anc int32 [rows, anc_ld]; out_lo is ignored unless f16."""
    _need_cuda(qkv, kcache, vcache, out, anc)
    L = _capi.lib()
    rows, H, t_max = kcache.shape[0], kcache.shape[1], kcache.shape[2]
    _capi.check(L.gic_test_attn_decode_beam(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), out.data_ptr(), out_lo.data_ptr() if f16 else None, anc.data_ptr(),
                                            anc.shape[1], pos, rows, H, t_max, n_prefix, beams, int(f16), int(shared), _stream()))


@torch.library.custom_op("gic::test_attn_prefill", mutates_args=("kcache", "vcache", "out"))
def test_attn_prefill(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, out: torch.Tensor, S: int) -> None:
    """qkv [rows*S, 3*H*64] bf16; kcache / vcache [rows, H, t_max, 64] bf16; out [rows*S, H*64] bf16."""
    _need_cuda(qkv, kcache, vcache, out)
    L = _capi.lib()
    rows, H, t_max = kcache.shape[0], kcache.shape[1], kcache.shape[2]
    _capi.check(L.gic_test_attn_prefill(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), out.data_ptr(), rows, S, H, t_max, _stream()))


@torch.library.custom_op("gic::test_layernorm", mutates_args=("y",))
def test_layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, y: torch.Tensor) -> None:
    _need_cuda(x, w, b, y)
    L = _capi.lib()
    _capi.check(L.gic_test_layernorm(_ptr(x), _ptr(w), _ptr(b), _ptr(y), x.shape[0], x.shape[1], _stream()))
