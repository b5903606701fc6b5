/* gic_b200.h -- C ABI of libgic_b200.so: the B200 (sm_100a) caption-generation hot path.
 *
 * The reference (thenoobychocobo/gpt2-image-captioning) has no FFI of its own: the path is a Python
 * class API (`ImageCaptioningModel.generate`, src/models.py:327-477; `RetrievalAugmentedTransformer
 * .generate`, src/models.py:748-771).  This header is the boundary a maintainer binds with ctypes
 * (INTEGRATION.md shows the stub); each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer marked `dev` is a DEVICE pointer owned by the caller
 *   - no ownership transfer: the caller allocates outputs and one workspace (size from
 *     gic_workspace_bytes); the opaque engine handle owns only its packed weight copies
 *   - every function returns GIC_OK (0) or a negative error code and never throws;
 *     gic_last_error() returns the message of the last failure on the calling thread
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous
 *     on that stream unless stated otherwise; one host thread per engine handle at a time, different handles may be
 *     driven concurrently from different host threads on different streams (batches in flight: inflight.py)
 *   - there is no CPU fallback: without a CUDA device every compute call fails with GIC_ERR_CUDA
 */
#ifndef GIC_B200_H_
#define GIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIC_ABI_VERSION 1

#if defined(__GNUC__)
#define GIC_API __attribute__((visibility("default")))
#else
#define GIC_API
#endif

enum { GIC_OK = 0, GIC_ERR_INVALID = -1, GIC_ERR_CUDA = -2, GIC_ERR_UNSUPPORTED = -3, GIC_ERR_WORKSPACE = -4 };

/* arithmetic modes.  F32: CUDA-core FFMA GEMMs, fp32 everywhere (token-exact parity mode).
 * BF16: bf16 weights/activations/KV cache, tcgen05 MMA with fp32 TMEM accumulators, fp32 residual stream,
 * LayerNorm statistics, softmax and logits.  BF16X2: every GEMM operand split hi+lo (two bf16), three
 * tcgen05 MMAs per product -- ~16 mantissa bits on the tensor cores -- with q | k | v and the KV cache in IEEE half
 * (11 significant bits in the bf16 cache's bytes); the tensor-core mode that meets the north-star tolerance
 * (>= 99 % of greedy captions identical to the fp32 reference) and the one the headline benchmark is quoted in. */
enum { GIC_DTYPE_F32 = 0, GIC_DTYPE_BF16 = 1, GIC_DTYPE_BF16X2 = 2 };
enum { GIC_MAPPER_MLP = 0, GIC_MAPPER_TRANSFORMER = 1 };
enum { GIC_AGG_MEAN = 0, GIC_AGG_MAX = 1, GIC_AGG_SUM_NORM = 2 };

typedef struct gic_engine gic_engine; /* opaque */

typedef struct gic_config {
  int32_t abi_version;     /* GIC_ABI_VERSION */
  int32_t dtype;           /* GIC_DTYPE_* */
  /* GPT-2 (HF GPT2Config): small 768/12/12, medium 1024/24/16, large 1280/36/20; head_dim must be 64 */
  int32_t n_embd, n_layer, n_head, vocab_size, n_positions;
  /* mapping network (src/models.py:14-174) */
  int32_t mapper_kind;     /* GIC_MAPPER_* */
  int32_t embed_dim;       /* E: 512 CLIP ViT-B/32, 1024 DINOv3 / ViT-L */
  int32_t prefix_length;   /* image prefix tokens P */
  int32_t hidden_length;   /* transformer mapper: image tokens Hl (src/models.py:119) */
  int32_t mapper_layers;   /* transformer mapper: encoder layers (8) */
  int32_t mapper_heads;    /* transformer mapper: nhead (8, src/models.py:131) */
  int32_t task_prefix_length; /* rows of task_prefix_embeds appended AFTER the image prefix (src/models.py:364-375); 0 = none */
  int32_t eos_token_id;    /* tokenizer.eos_token_id (src/models.py:348); 50256 for GPT-2 */
} gic_config;

/* fp32 parameter tables in the reference's NATIVE layouts (HF Conv1D weights are [in,out]; nn.Linear and
 * wte are [out,in]).  The engine makes its own packed copies; the caller's tensors are not referenced
 * after the load call's stream work completes. */
typedef struct gic_gpt2_layer_weights {
  const float *ln1_w, *ln1_b;     /* [d] */
  const float *attn_w, *attn_b;   /* c_attn  [d,3d], [3d]   (HF modeling_gpt2.py:185) */
  const float *proj_w, *proj_b;   /* c_proj  [d,d],  [d]    (:223) */
  const float *ln2_w, *ln2_b;     /* [d] */
  const float *fc_w, *fc_b;       /* mlp.c_fc   [d,4d], [4d] (:238-243) */
  const float *fc2_w, *fc2_b;     /* mlp.c_proj [4d,d], [d] */
} gic_gpt2_layer_weights;

typedef struct gic_gpt2_weights {
  const float *wte;               /* [V,d], tied LM head (HF :646,651) */
  const float *wpe;               /* [n_positions,d] */
  const float *lnf_w, *lnf_b;     /* [d] */
  const gic_gpt2_layer_weights* layers; /* host array of n_layer entries (device pointers inside) */
} gic_gpt2_weights;

typedef struct gic_mlp_mapper_weights { /* MLPMappingNetwork.model.{0,2} (src/models.py:52-56) */
  const float *w1, *b1;           /* [P*d/2, E], [P*d/2] */
  const float *w2, *b2;           /* [P*d, P*d/2], [P*d] */
} gic_mlp_mapper_weights;

typedef struct gic_tfm_layer_weights { /* nn.TransformerEncoderLayer, norm_first, relu (src/models.py:129-136) */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b;   /* [d] */
  const float *in_proj_w, *in_proj_b;                   /* [3d,d], [3d]  (q,k,v order) */
  const float *out_proj_w, *out_proj_b;                 /* [d,d], [d] */
  const float *lin1_w, *lin1_b;                         /* [4d,d], [4d] */
  const float *lin2_w, *lin2_b;                         /* [d,4d], [d] */
} gic_tfm_layer_weights;

typedef struct gic_tfm_mapper_weights { /* TransformerMappingNetwork (src/models.py:119-139) */
  const float *linear_w, *linear_b;                     /* [Hl*d, E], [Hl*d] */
  const float *prefix_const;                            /* [P, d] */
  const gic_tfm_layer_weights* layers;                  /* host array of mapper_layers entries */
} gic_tfm_mapper_weights;

/* ---- lifecycle ------------------------------------------------------------------------------------ */
GIC_API const char* gic_last_error(void);
GIC_API int gic_abi_version(void);
GIC_API int gic_device_check(void); /* GIC_OK iff device 0.. current device is sm_100 (compute capability 10.x) */

/* replaces: model construction + `.to(device)` (src/eval.py:189-190) for the engine's packed weights */
GIC_API int gic_engine_create(const gic_config* cfg, gic_engine** out);
GIC_API int gic_engine_destroy(gic_engine* e);
/* another CONTEXT on the packed weights of `src` (own stream, events, CUDA-graph cache): lets a second batch be in flight on the same
 * GPU without a second weight copy.  The clone owns no weights (gic_engine_weight_bytes == 0) and must be destroyed before `src`;
 * one host thread per context at a time, different contexts may be driven concurrently. */
GIC_API int gic_engine_clone(const gic_engine* src, gic_engine** out);
GIC_API int gic_engine_load_gpt2(gic_engine* e, const gic_gpt2_weights* w, void* stream);
GIC_API int gic_engine_load_mlp_mapper(gic_engine* e, const gic_mlp_mapper_weights* w, void* stream);
GIC_API int gic_engine_load_tfm_mapper(gic_engine* e, const gic_tfm_mapper_weights* w, void* stream);
GIC_API int gic_engine_load_task_prefix(gic_engine* e, const float* task_prefix_embeds /* dev [Tp,d] */, void* stream);
GIC_API size_t gic_engine_weight_bytes(const gic_engine* e); /* bytes of packed weights held by the handle */

/* workspace for one generate call: activations + KV cache [L][2][rows][H][T_max][64] + LM-head partials.
 * rows = batch * max(1, num_beams); T_max = prefix_length + task_prefix_length + max_new_tokens. */
GIC_API size_t gic_workspace_bytes(const gic_engine* e, int batch, int max_new_tokens, int num_beams);

/* ---- the hot path ---------------------------------------------------------------------------------- */
/* replaces MLPMappingNetwork.forward / TransformerMappingNetwork.forward (src/models.py:58-74,141-174)
 * plus the task-prefix concat (:364-375).  prefix_out: dev fp32 [B, P_total, d]. */
GIC_API int gic_mapper_forward(gic_engine* e, const float* image_embeddings /* dev [B,E] */, int batch,
                       float* prefix_out, void* workspace, size_t workspace_bytes, void* stream);

/* replaces the greedy branch of ImageCaptioningModel.generate (src/models.py:327-477, temperature == 0):
 * mapper -> prefill -> (max_new_tokens-1) KV-cached decode steps with the LM head fused with argmax
 * (ties -> lowest index), EOS rows keep emitting EOS (:453-460).  Always writes the full
 * ids_out dev int64 [B, max_new_tokens]; *gen_len_out (dev int32, may be NULL) receives L_gen of
 * :390-391 (the caller slices [:, :L_gen]).  logits_out (dev fp32 [max_new_tokens, B, V], may be NULL)
 * receives every step's last-position logits -- a parity/debug tap, not used by the product path.
 * Rows finish individually (:453-460): between chunks of 4 decode steps the call stops once every row has emitted EOS (:390-391) and
 * shrinks the batch to its unfinished rows when at least a quarter of the slots can be dropped (GIC_NO_COMPACT=1 / GIC_NO_EARLY_EXIT=1
 * turn these off; the tokens do not depend on them). */
GIC_API int gic_generate_greedy(gic_engine* e, const float* image_embeddings /* dev [B,E] */, int batch, int max_new_tokens,
                        int64_t* ids_out, int32_t* gen_len_out, float* logits_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* replaces the sampling branch of ImageCaptioningModel.generate (src/models.py:400-449, temperature > 0): logits / temperature,
 * nucleus filter (keep the sorted tokens up to and including the first whose cumulative probability exceeds top_p; top_p >= 1 keeps
 * all), one multinomial draw per row and step from a Philox4x32-10 stream keyed by (seed, row, step).  Same outputs and EOS rules as
 * gic_generate_greedy; logits_scratch: dev fp32 [B, V rounded up to a multiple of 32] (one step's logits, 16-byte aligned; the tensor-core
 * engines pad the rows so that the LM-head epilogue and the sampler move whole aligned vectors).  The decode steps replay the same CUDA
 * graphs as the greedy call (temperature / top_p / seed are read from the workspace), keyed by the scratch address: pass the same buffer
 * on every call.  Parity with torch.multinomial is distributional. */
GIC_API int gic_generate_sample(gic_engine* e, const float* image_embeddings /* dev [B,E] */, int batch, int max_new_tokens, float temperature,
                                float top_p, unsigned long long seed, int64_t* ids_out, int32_t* gen_len_out, float* logits_scratch,
                                void* workspace, size_t workspace_bytes, void* stream);

/* Beam search (not in the reference; semantics of HF GenerationMixin._beam_search, generation/utils.py:3076-3385,
 * do_sample=False, early_stopping=False, length_penalty given, num_return_sequences=1, eos = pad): ids_out dev int64
 * [B, max_new_tokens] = best finished hypothesis per image padded with eos; scores_out dev fp32 [B] (may be NULL) its
 * length-normalised score; *gen_len_out dev int32 (may be NULL) the longest selected hypothesis (HF crops to it).
 * 2 <= num_beams <= 8.  The tensor-core engines never reorder the KV cache: decode attention follows a per-hypothesis ancestry table
 * (cache row of every generated position); the fp32 engine (and GIC_BEAM_REORDER=1) reorders with the gic_kv_reorder gather. */
GIC_API int gic_generate_beam(gic_engine* e, const float* image_embeddings, int batch, int max_new_tokens, int num_beams,
                      float length_penalty, int64_t* ids_out, float* scores_out, int32_t* gen_len_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* replaces DynamicCache.reorder_cache / index_select(0, beam_idx) per layer (HF cache_utils.py:81-85):
 * dst[l][kv][r] = src[l][kv][beam_idx[r]] for the first `ctx_len` positions.  Element type follows the engine dtype
 * (F32: float; BF16 / BF16X2: 2-byte elements). */
GIC_API int gic_kv_reorder(gic_engine* e, const void* kv_src, void* kv_dst, const int32_t* beam_idx /* dev [rows] */,
                   int rows, int ctx_len, int t_max, void* stream);

/* replaces faiss IndexFlatIP.search as called by retrieve_images_by_vector_similarity
 * (src/database/faiss_store.py:153-155): exact inner product of q [B,D] against db [N,D] (fp32, row-major),
 * k best per row, descending, ties -> lowest index; scores_out dev fp32 [B,k], idx_out dev int64 [B,k]
 * (-1 / -inf when k > N).  workspace from gic_topk_workspace_bytes. */
GIC_API size_t gic_topk_workspace_bytes(int batch, int n_rows, int dim, int k);
GIC_API int gic_topk_ip(const float* queries, const float* db, int batch, int n_rows, int dim, int k,
                float* scores_out, int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream);
/* The same search with the scan on the tensor cores: a bf16x2 (hi + lo operands) tcgen05 GEMM finds 32 candidates per query, which are
 * re-scored exactly in fp32 in the exact path's summation order; a per-query certificate (worst candidate's approximate score + error
 * bound < k-th exact score) proves no other row can enter the top k, and an uncertified query is rescanned exactly on the device.
 * Scores and indices are therefore those of gic_topk_ip.  Needs dim % 64 == 0 and k <= 16 (gic_topk_tc_supported); db_hi / db_lo =
 * gic_pack_bf16x2(db) [n_rows, dim] bf16 each, db_norm_max = the largest row norm of db. */
GIC_API int gic_topk_tc_supported(int dim, int k);
GIC_API size_t gic_topk_tc_workspace_bytes(int batch, int n_rows, int dim, int k);
GIC_API int gic_pack_bf16x2(const float* src, void* hi, void* lo, size_t n, void* stream);
GIC_API int gic_topk_ip_tc(const float* queries, const float* db, const void* db_hi, const void* db_lo, float db_norm_max, int batch, int n_rows,
                           int dim, int k, float* scores_out, int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream);

/* replaces the Python hit filter + caption-row selection (faiss_store.py:160-183,208-229) on integer ids:
 * image hits -> first top_i with idx != -1 and score <= 0.9999 -> their caption rows via a CSR table
 * (cap_row_start dev int64 [N_img+1]; the caption rows of image i are cap_row_ids[start[i] .. start[i+1]), or the
 * range itself when cap_row_ids is NULL) in hit order, first top_k; rows_out dev int64 [B, top_k], -1 = zero padding row. */
GIC_API int gic_select_caption_rows(const float* scores, const int64_t* idx, int batch, int k_searched,
                            const int64_t* cap_row_start, const int64_t* cap_row_ids, int top_i, int top_k,
                            int64_t* rows_out, void* stream);

/* replaces get_caption_embeddings' reconstruct loop and zero padding alone (faiss_store.py:229-251), for callers that pool with the
 * differentiable RetrievalAggregator module (the training forward, src/models.py:733-735): out dev fp32 [B, top_k, dim],
 * out[b, j] = cap_db[rows[b, j]] or a zero row for -1. */
GIC_API int gic_gather_caption_rows(const float* cap_db, const int64_t* rows, int batch, int top_k, int dim, float* out, void* stream);

/* replaces get_caption_embeddings' reconstruct loop + RetrievalAggregator.forward (faiss_store.py:229-251,
 * src/models.py:589-625): out[b] = q[b] + agg_k(cap_db[rows[b,k]]) with zero rows for -1 (padding counts in the mean). */
GIC_API int gic_gather_aggregate_add(const float* queries, const float* cap_db, const int64_t* rows, int batch, int top_k, int dim,
                             int aggregation, float* out, void* stream);
/* the "attention" pooling of RetrievalAggregator (src/models.py:606-616): weights = softmax_k(attention_proj(r_k)), out[b] = q[b] + sum_k
 * weights_k r_k over the gathered rows (zero rows for -1: their score is the bias).  attn_w dev fp32 [dim], attn_b dev fp32 [1]. */
GIC_API int gic_gather_attention_add(const float* queries, const float* cap_db, const int64_t* rows, int batch, int top_k, int dim,
                                     const float* attn_w, const float* attn_b, float* out, void* stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------------ */
/* kernels launched by this library in this process so far (CUDA-graph replays count their kernel nodes) */
GIC_API unsigned long long gic_launch_count(void);
/* finished-row compactions performed by gic_generate_greedy in this process so far (a batch whose rows emit EOS at different steps is
 * shrunk to its live rows between chunks of decode steps; the KV cache stays in place behind a slot -> row map) */
GIC_API unsigned long long gic_compaction_count(void);
/* while enabled, generate calls run without the CUDA graph and bracket every kernel class with CUDA events on the
 * launch stream; gic_profile_read synchronises and returns per-class launch counts and summed device time.
 * A "class" spans one logical op (e.g. "attn_decode", "gemm_qkv", "lm_head", "layernorm", "prefill_gemm"). */
typedef struct gic_profile_entry { char name[32]; int32_t launches; float total_ms; } gic_profile_entry;
GIC_API int gic_profile_enable(gic_engine* e, int on);
GIC_API int gic_profile_read(gic_engine* e, gic_profile_entry* out, int max_entries, int* n_out);

/* ---- kernel-level entry points (unit tests / profiling; same kernels the path uses) --------------------- */
/* C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias[N]) ; epilogue: 0 none, 1 tanh, 2 gelu_tanh, 3 relu, 4 += residual(fp32 C in place).
 * dtype F32: A,W,C fp32 (CUDA cores).  BF16 / BF16X2: A,W fp32 inputs are packed internally, C fp32. */
GIC_API int gic_test_gemm(int dtype, const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int epilogue, void* stream);
/* fused bf16 MLP sub-block (LayerNorm folded into c_fc, GELU, c_proj + residual with bf16 copy + row statistics, optional
 * split-K) on caller data -- the kernel-level parity hook for HF GPT2Block's ln_2 + GPT2MLP (HF:models/gpt2/modeling_gpt2.py:238-243,304-307) */
GIC_API int gic_test_ln_mlp(float* h, const float* gamma, const float* beta, const float* wfc, const float* bfc, const float* wfc2, const float* bfc2,
                            void* hb_out, void* stats_out, int M, int d, int split_k, void* stream);
/* one decode-attention launch on caller data: qkv [rows, 3*H*64] bf16, K / V caches [rows][H][t_max][64] bf16 with `pos` cached tokens;
 * appends the new K / V at `pos`, writes out [rows, H*64] bf16 (HF:models/gpt2/modeling_gpt2.py:185-220 for one query).  variant < 0: product kernel */
GIC_API int gic_test_attn_decode(const void* qkv, void* kcache, void* vcache, void* out, int pos, int rows, int H, int t_max, int variant, void* stream);
/* one beam-search decode-attention launch on caller data (no cache reorder: HF DynamicCache.reorder_cache, HF:cache_utils.py:81-85, replaced by an
 * ancestry table): rows = images * beams hypotheses; row r reads positions [0, n_prefix) from the prefill row of its image ((r / beams) * beams),
 * position n_prefix + g from cache row anc[r * anc_ld + g] (g < pos - n_prefix), plus its own new token, which is appended to row r at `pos`.
 * f16 != 0: q / k / v / cache are IEEE half and out_lo receives the remainder of the bf16 output (bf16x2 engine).  shared: 1 = the image's
 * beams share one read of the prefix (attn_decode_beam_kernel), 0 = one walk per hypothesis. */
GIC_API int gic_test_attn_decode_beam(const void* qkv, void* kcache, void* vcache, void* out, void* out_lo, const int32_t* anc, int anc_ld, int pos, int rows, int H,
                                      int t_max, int n_prefix, int beams, int f16, int shared, void* stream);
/* measurement hook (bench.py `roofline`): `launches` back-to-back product decode-attention launches, ASYNCHRONOUS on `stream` (the caller
 * brackets them with CUDA events); launch i works on cache plane i % planes (kcache / vcache + plane * rows*H*t_max*64 elements), the way
 * the 12 layers of a decode step do, so consecutive launches do not find each other's lines in L2.  d_pos: dev int, tokens already cached. */
GIC_API int gic_bench_attn_decode(const void* qkv, void* kcache, void* vcache, void* out, const int* d_pos, int rows, int H, int t_max, int planes,
                                  int launches, void* stream);
/* one causal prefill-attention launch on caller data: qkv [rows*S, 3*H*64] bf16 -> out [rows*S, H*64] bf16; K / V land in the caches
 * [rows][H][t_max][64] at positions 0..S-1 (HF:models/gpt2/modeling_gpt2.py:185-220, HF:cache_utils.py:102-121) */
GIC_API int gic_test_attn_prefill(const void* qkv, void* kcache, void* vcache, void* out, int rows, int S, int H, int t_max, void* stream);
/* one sampling launch on caller data: logits dev fp32 [B, V] -> tokens dev int32 [B] (src/models.py:400-449) */
GIC_API int gic_test_sample_top_p(const float* logits, int B, int V, float temperature, float top_p, unsigned long long seed, int step,
                                  int32_t* tokens_out, void* stream);
/* in-situ timeline: every GEMM / decode-attention / ln_f / finalize launch takes the next record of buf (dev uint64 [3 * cap], pre-filled
 * with (0, ~0, 0)): kind | detail << 8, begin_ns = min over its blocks of %globaltimer after griddepcontrol.wait, end_ns = max over its
 * blocks at their end.  While installed, generate captures all decode steps into one graph.  buf = NULL: off. */
GIC_API int gic_trace_install(void* buf, unsigned int cap);
GIC_API int gic_test_layernorm(const float* x, const float* w, const float* b, float* y, int rows, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIC_B200_H_ */
