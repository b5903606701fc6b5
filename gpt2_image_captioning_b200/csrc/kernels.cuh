// Launch-function declarations for every kernel of the caption path (definitions in the .cu files).
#pragma once
#include "common.cuh"

namespace gic {

enum Epilogue { EPI_NONE = 0, EPI_TANH = 1, EPI_GELU = 2, EPI_RELU = 3, EPI_RESIDUAL = 4, EPI_ARGMAX = 5 /* internal: set by part_val */,
                EPI_ARGMAX2 = 6 /* internal: set by part_val2 (best + runner-up value per slot) */,
                EPI_TOPK = 7 /* internal: set by topk_v (per-row running top-8 + drop bound, no score matrix) */,
                EPI_BEAM = 8 /* internal: set by topk_v + beam_m (per-row running top-16 + online log-sum-exp: the beam-search LM head) */ };

// Where a producer kernel writes an activation: fp32 and/or bf16 (hi) and/or the bf16 remainder (lo = bf16(v - hi),
// the second half of a BF16X2 GEMM operand).  Any pointer may be null.
struct ActOut {
  float* f32 = nullptr;
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  __device__ __forceinline__ void write(size_t i, float v) const {
    if (f32) f32[i] = v;
    if (hi) {
      const bf16 h = __float2bfloat16_rn(v);
      hi[i] = h;
      if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
};

// ---- sgemm_fp32.cu : CUDA-core fp32 GEMM, C[M,N] = epi(A[M,K] . W[N,K]^T + bias) ----------------------
// A rows have stride lda (elements); W is [N,K] K-major; C row stride ldc.  EPI_RESIDUAL: C += result (in place).
int launch_sgemm_nt(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N, int K, int epilogue,
                    cudaStream_t st);

// ---- gemm_tcgen05.cu : bf16 tcgen05/TMEM GEMM fed by TMA ----------------------------------------------------
struct alignas(64) TmaDesc { unsigned char bytes[128]; };  // CUtensorMap
int tma_init();  // resolves cuTensorMapEncodeTiled through the runtime
// bf16 row-major [rows, cols] operand map: one box = [box_rows][64 cols], 128B swizzle, out-of-bounds -> zero
int make_tma_2d_bf16(TmaDesc* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows);

struct GemmBf16Args {
  TmaDesc a_hi, w_hi, a_lo, w_lo;  // lo maps used only when split (BF16X2); A box rows = 128, W box rows = block_n
  int M = 0, N = 0, K = 0;
  int block_n = 128;               // 32 | 64 | 128 | 192 | 256 (split: <= 128)
  int split = 0;                   // 0: one MMA per k-step; 1: hi.hi + hi.lo + lo.hi
  int epilogue = EPI_NONE;
  int mma_repeat = 1;              // microbenchmark probe only (re-issues every MMA this many times)
  const float* bias = nullptr;     // [N] or null
  ActOut out;                      // fp32 (EPI_RESIDUAL: in-place +=) and/or bf16 hi/lo, row stride ld_out
  int out_f16 = 0;                 // out.hi receives IEEE half instead of bf16 (q | k | v of the bf16x2 mode: fp16 KV cache)
  int ld_out = 0;
  float* part_val = nullptr;       // fused LM-head argmax: [M][part_ld] best value per (tile, column-parity warp) slot ...
  int* part_idx = nullptr;         // ... and its lowest column index (cols >= N masked)
  int part_ld = 0;                 // slots per row (>= 2 * n_tiles)
  float* part_val2 = nullptr;      // optional [M][part_ld]: second-best value of each slot (lm_head_rescore_kernel's candidate test)
  // fused top-k scan (retrieval): per row and stream the GEMM_TOPK_KEEP (8) best (score, column) + the largest dropped score; stream count from
  // gemm_topk_streams(M, N, block_n, pair).  topk_v / topk_i [M][streams][8], topk_u [M][streams]
  float* topk_v = nullptr; int* topk_i = nullptr; float* topk_u = nullptr; int topk_streams = 0;
  // beam-search LM head (HF:generation/utils.py:3252-3256 log_softmax, :2981-2987 top 2 * beams): with beam_m / beam_s set the streams keep
  // GEMM_BEAM_KEEP (16) candidates each (topk_v / topk_i [M][streams][16], no drop bound) and every stream's running maximum and
  // sum of exp(logit - maximum) over the columns it saw ([M][streams] each): log-sum-exp and top-2K without a logits matrix
  float* beam_m = nullptr; float* beam_s = nullptr;
  // LayerNorm folded into the GEMM (A holds the RAW rows x, W was packed as gamma_k * W[n,k], bias as b_n + sum_k beta_k W[n,k]):
  //   out[r,n] = rstd_r * (acc[r,n] - mean_r * ln_colsum[n]) + bias[n],   mean / rstd from sum_parts ln_stats[part][r * mul + off]
  const float2* ln_stats = nullptr;  // [ln_parts][ln_stats_ld] (sum x, sum x^2) partials written by the producer of A
  int ln_parts = 0; long ln_stats_ld = 0; int ln_row_mul = 1, ln_row_off = 0;
  const float* ln_colsum = nullptr;  // [N] sum_k of the packed (bf16) gamma-folded weights
  float2* stats_out = nullptr;       // [ceil(N / 32)][ln_stats_ld]: per-row (sum, sum of squares) of the values this GEMM writes, per 32-column chunk
  int w_static = 0;  // the weights were written long before this launch (engine weights): their first tiles may be fetched before
                     // griddepcontrol.wait, i.e. while the previous kernel is still finishing
  int no_pdl = 0;  // launch in plain stream order (no programmatic dependent launch): for callers outside the engine's launch chain, whose
                   // neighbours are ordinary <<<>>> launches (the retrieval scan)
  int pair = 0;  // CTA pairs: 256 x block_n tiles by two CTAs (cta_group::2); w_hi's box holds block_n / 2 rows
  // split-K for short-and-wide problems (few output tiles, long K): split_k consecutive CTAs (or CTA pairs) per tile park fp32 partials in
  // splitk_ws; summed in slice order (deterministic).  When every work item has its own resident CTA the slices of a tile wait for each
  // other and each finishes its share of the tile (cooperative reduction), otherwise the CTA that arrives last does
  int split_k = 1; float* splitk_ws = nullptr;  /* [split_k][M][N] fp32 */  int* splitk_counters = nullptr;  /* [2 * tiles], zero on entry and exit */
  long long* trace = nullptr;      // microbenchmark only: device buffer of >= 640 int64 for CTA 0's clock64 timeline
};
int launch_gemm_bf16(const GemmBf16Args& a, cudaStream_t st);
int gemm_topk_streams(int M, int N, int block_n, int pair);
constexpr int GEMM_TOPK_KEEP_PUBLIC = 8;  // == GEMM_TOPK_KEEP (gemm_tcgen05.cu)
constexpr int GEMM_BEAM_KEEP = 16;        // candidates per stream of the beam-search LM head (>= 2 * beams for beams <= 8)
int gemm_bf16_pick_block_n(int M, int N, int split);
// tile width for an M x N x K problem cut into split_k K slices (split: bf16x2 operands)
// pair (may be null): out, 1 = run as CTA pairs (cta_group::2); the W tensor map's box is then block_n / 2 rows and GemmBf16Args::pair is set
void gemm_bf16_pick(int M, int N, int K, int split, int split_k, int* block_n, int* pair = nullptr);
int gemm_bf16_split_k_for(int N, int K);  // K split of a decode-size residual GEMM: depends on the shape only

int gemm_bf16_configure();  // cudaFuncSetAttribute for every instantiation (call once, outside stream capture)

// ---- elementwise.cu -----------------------------------------------------------------------------------------------
int launch_layernorm(const float* x, long x_row_stride, const float* w, const float* b, ActOut y, int rows, int d, cudaStream_t st);
// scale_k (optional, [K]): multiply every weight by scale_k[k] -- LayerNorm gamma folded into the weight (K = in-features)
int launch_pack_weight(const float* in, int R, int C, bool transpose, ActOut out, cudaStream_t st, const float* scale_k = nullptr);
// LayerNorm folding, load time: colsum[n] = sum_k float(w_packed[n,k]);  bias_out[n] = bias[n] + sum_k beta[k] * w(n,k)
// (w_src: the original fp32 weight, [K,N] when `transposed` else [N,K])
// (w_packed_lo: the bf16x2 remainder of the packed weight; colsum then sums hi + lo)
int launch_fold_ln(const bf16* w_packed, const float* w_src, bool transposed, const float* beta, const float* bias, float* colsum,
                   float* bias_out, int N, int K, cudaStream_t st, const bf16* w_packed_lo = nullptr);
// raw rows for a GEMM with folded LayerNorm: xb = bf16(x), stats[row] = (sum xb, sum xb^2)
// (xb_lo: bf16x2 remainder; the statistics are then those of hi + lo)
int launch_row_stats(const float* x, long x_row_stride, bf16* xb, float2* stats, int rows, int d, cudaStream_t st, bf16* xb_lo = nullptr);
int launch_convert(const float* in, ActOut out, size_t n, cudaStream_t st);
int launch_embed_prefix(const float* prefix, int P_img, const float* task, int P_task, const float* wpe, float* h, float* prefix_out,
                        int B, int d, cudaStream_t st);
int launch_slice_tokens(const float* in, int S, int tok0, int n_tok, float* out, int B, int d, cudaStream_t st);
int launch_build_mapper_seq(const float* lin, const float* prefix_const, float* seq, int B, int Hl, int P, int d, cudaStream_t st);

// ---- attention.cu -------------------------------------------------------------------------------------------------
// KV cache layout: [L][2][rows][H][T_max][64], element T.  qkv rows are [.., 3d] (q | k | v), head h at h*64.
// cache_row_mult: sequence b writes its K/V into cache row b * cache_row_mult (beam search prefills B rows into a
// cache laid out for B * beams rows)
template <typename T>
int launch_attn_prefill(const T* qkv, T* kcache, T* vcache, ActOut out, int B, int P, int H, int t_max, int cache_row_mult,
                        cudaStream_t st);
// row_map (optional, dev int [rows]): compacted batch -- activation slot r attends / appends to cache row row_map[r]; -1 = padding slot
template <typename T>
int launch_attn_decode(const T* qkv, T* kcache, T* vcache, ActOut out, const int* d_pos, int rows, int H, int t_max, cudaStream_t st,
                       const int* row_map = nullptr);
// bf16 decode attention through a beam-ancestry table instead of a reordered cache (beam search); see attention.cu
// (out_lo non-null: the fp16-cache / hi + lo output flavour of the bf16x2 engine)
int launch_attn_decode_indirect(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, const int* d_pos, int rows, int H, int t_max, const int* anc,
                                int anc_ld, int n_prefix, int beams, cudaStream_t st, bf16* out_lo = nullptr);
void attn_decode_set_beam_shared(int v);  // test knob: 1 = attn_decode_beam_kernel (prefix shared by an image's beams), 0 = one walk per hypothesis, < 0 = default
// bf16x2 engine: q | k | v and the KV cache are IEEE half (2-byte elements, typed bf16* for the shared plumbing), the output a bf16 hi + lo pair
int launch_attn_decode_f16(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out_hi, bf16* out_lo, const int* d_pos, int rows, int H, int t_max,
                           cudaStream_t st, const int* row_map = nullptr);
int launch_attn_prefill_f16(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out_hi, bf16* out_lo, int B, int P, int H, int t_max, int cache_row_mult,
                            cudaStream_t st);
bool attn_decode_indirect_available();
void attn_decode_set_variant(int v);  // microbenchmark: ring geometry of the bulk-copy decode kernel (0 = product)
int attn_decode_configure();  // cudaFuncSetAttribute for the bulk-copy decode kernel (call once, outside stream capture)
template <typename T>
int launch_attn_encoder(const T* qkv, ActOut out, int B, int S, int H, int hd, cudaStream_t st);
template <typename T>
int launch_kv_reorder(const T* src, T* dst, const int* beam_idx, int L, int rows, int H, int ctx_len, int t_max, cudaStream_t st);

// ---- lmhead.cu ----------------------------------------------------------------------------------------------------
constexpr int LMHEAD_F32_PARTS = 8;
int launch_argmax_partials(const float* logits, int B, int V, float* part_val, int* part_idx, int part_ld, cudaStream_t st);
struct FinalizeArgs {
  const float* part_val; const int* part_idx; int n_parts, part_ld;  // [B][part_ld], the first n_parts slots of a row are valid
  int B, d, eos, max_new, P, n_pos;
  int* d_step;              // device scalar: index of the token being produced; advanced by the kernel
  int* d_pos;               // device scalar: KV position of the token fed to the next decode step
  int* done_counter;        // device scalar used to elect the last block
  StepTrace step_trace = {nullptr, 0, 0};  // filled by launch_finalize_token
  int* fin_counter = nullptr; int* all_done = nullptr;  // optional: rows finished this step / set to 1 once every row has emitted EOS
  unsigned char* finished;  // [B]
  int* first_eos;           // [B]; max_new = never
  int64_t* ids_out;         // [B, max_new]
  const float* wte_f32; const bf16* wte_bf16;  // exactly one: embedding table [V,d]
  const float* wpe;         // [n_pos, d]
  float* h_next;            // [B, d] fp32: next-step input = wte[tok] + wpe[P + step]
  bf16* hb_next;            // optional [B, d]: its bf16 copy (A operand of the first GEMM with folded LayerNorm) ...
  bf16* hb_next_lo = nullptr;  // ... bf16x2: the remainder bf16(x - hi) (fp32 embedding table only) ...
  float2* stats_next;       // ... and [B] (sum, sum of squares) of that copy
  const int* row_map = nullptr;  // compacted batch: caption row of activation slot b (-1 = padding slot); null = identity
  int* live_rows = nullptr;      // optional device scalar: unfinished rows after this step
  const unsigned long long* packed_best = nullptr;  // exactly re-scored head: the row's token as a packed (value, ~column) key instead of partials
  int* rescore_counters = nullptr;  // ... and its two list counters, reset for the next step
};
int launch_finalize_token(const FinalizeArgs& a, cudaStream_t st);
// exact greedy token from the single-MMA bf16 head's per-slot (best, column, runner-up) partials (lmhead.cu): candidates within the
// rigorous rounding bound of the approximate maximum are listed as (row, column) pairs and re-scored in fp32 over the whole GPU
struct RescoreArgs {
  const float* h; long h_row_stride; const float* lnw; const float* lnb;  // residual rows and ln_f
  const float* wte; const float* wte_norm; const float* slot_norm_max;     // fp32 table [V,d], its row norms [V], per-slot largest norm for this tiling
  const float* part_val; const int* part_idx; const float* part_val2; int n_parts, part_ld, block_n;
  int rows, V, d;
  float* a_f32;                  // [rows, d] scratch: ln_f(h) in fp32
  int2* pairs; int pair_cap; int row_budget;
  int* pair_count; int* flag_count;  // device counters (adjacent ints: rescore_counters[0], [1]); zero before the first step, reset by finalize
  int* row_flag; int* flag_rows;     // [rows] each; row_flag is cleared by ... the candidates kernel of the next step (see below)
  unsigned long long* best;      // [rows] packed (value, ~column)
  const int* row_map;            // compacted batches: padding slots are skipped
  StepTrace step_trace;
};
int launch_lm_head_rescore(const RescoreArgs& r, cudaStream_t st);
int launch_row_norms(const float* w, int N, int K, float* out, cudaStream_t st);
int launch_slot_norm_max(const float* norms, int V, int block_n, float* out, cudaStream_t st);
// finished-row compaction between decode chunks (lmhead.cu): packs the live rows' next-step state to the first m_new slots and writes
// the new slot -> caption-row map; the KV cache stays where it is
struct CompactArgs {
  const unsigned char* finished; const int* row_map_old; int m_old, m_new, d;
  int* row_map_new; int* src_slot;
  float* h; float* h_tmp; bf16* a_hi; bf16* a_hi_tmp; bf16* a_lo; bf16* a_lo_tmp; float2* stats; float2* stats_tmp;
};
int launch_compact_rows(const CompactArgs& c, cudaStream_t st);
// temperature / top-p sampling of one token per row from fp32 logits [B, V] (src/models.py:400-449); the token goes to slot 0 of the
// row's (value, index) partials.  step: *d_step unless step_override >= 0 (the Philox counter is (row, step)).
// dev_params (optional): the kernel reads temperature / top_p / seed from device memory instead of its arguments, so that a CUDA graph
// holding the launch can be replayed by calls with other values (launch_set_sample_params writes them, outside the graph).
struct SampleParams { float inv_temperature, top_p; unsigned long long seed; };
int launch_set_sample_params(SampleParams* dst, float temperature, float top_p, unsigned long long seed, cudaStream_t st);
int launch_sample_top_p(const float* logits, int B, int V, float temperature, float top_p, unsigned long long seed, const int* d_step,
                        int step_override, float* part_val, int* part_idx, int part_ld, cudaStream_t st, const SampleParams* dev_params = nullptr,
                        long ld = 0 /* row stride of `logits` in floats; 0 = V */);
int launch_init_decode_state(unsigned char* finished, int* first_eos, int B, int max_new, int* d_step, int* d_pos, int* done_counter,
                             int P, int* fin_counter, int* all_done, int64_t* ids, int eos, cudaStream_t st);
int launch_gen_len(const int* first_eos, int B, int max_new, int* gen_len_out, cudaStream_t st);
int launch_spin(long long cycles, cudaStream_t st);  // profiling aid (see lmhead.cu)

// ---- beam.cu ------------------------------------------------------------------------------------------------------
struct BeamState {
  int B, beams, max_new, V, eos;
  int* run_seq[2]; int* fin_seq[2];       // [B, beams, max_new] token ids, double-buffered across steps
  float* run_score; float* fin_score;     // [B, beams]
  unsigned char* fin_flag; int* fin_len;  // [B, beams]
  unsigned char* unsat;                   // [B] early-stop heuristic still unsatisfied
  int* beam_idx; int* next_tok;           // [B * beams] parent cache row / token of every surviving beam
  float* cand_score; int* cand_idx;       // [B, 2 * beams]
  int* anc[2];                            // [B * beams, max_new] ancestry tables (double-buffered): cache row of every generated position
  float* lse; float* row_val; int* row_idx;  // [B * beams], [B * beams, 2 * beams]: per-row log-sum-exp and top-2*beams continuations
  // fused LM head (EPI_BEAM): per row and stream the kept candidates and the log-sum-exp partials; null = logits are materialised instead
  float* tk_v = nullptr; int* tk_i = nullptr; float* tk_m = nullptr; float* tk_s = nullptr; size_t tk_slots = 0;  // tk_slots: (row, stream) pairs allocated
};
int launch_beam_init(const BeamState& s, cudaStream_t st);
int launch_beam_topk(const float* logits, int B, int rows_per_image, int n_live, const float* run_score, int beams, int V, int K,
                     float* lse /* [B * rows_per_image] */, float* row_val, int* row_idx /* [B * rows_per_image, K] */, float* cand_score,
                     int* cand_idx, cudaStream_t st);
// the same from the EPI_BEAM streams of the LM-head GEMM (tk_v / tk_i [rows, streams, GEMM_BEAM_KEEP], bm / bs [rows, streams]): no logits matrix
int launch_beam_topk_streams(const float* tk_v, const int* tk_i, const float* bm, const float* bs, int streams, int B, int rows_per_image, int n_live,
                             const float* run_score, int beams, int V, int K, float* lse, float* row_val, int* row_idx, float* cand_score, int* cand_idx,
                             cudaStream_t st);
int launch_beam_update(const BeamState& s, int step, float len_denom, cudaStream_t st);
// anc_new[r][g] = g == t - 1 ? beam_idx[r] : anc_old[beam_idx[r]][g] for g < t (t = tokens generated so far)
int launch_beam_ancestry(const int* anc_old, int* anc_new, const int* beam_idx, int rows, int ld, int t, cudaStream_t st);
int launch_beam_embed(const int* next_tok, const float* wte_f32, const bf16* wte_bf16, const float* wpe, int pos, int d, float* h, int rows,
                      cudaStream_t st);
int launch_set_int(int* p, int v, cudaStream_t st);
int launch_beam_finalize(const BeamState& s, int final_buf, int64_t* ids_out, float* scores_out, int* gen_len_out, cudaStream_t st);

// ---- retrieval.cu -------------------------------------------------------------------------------------------------
size_t topk_workspace_bytes(int B, int N, int D, int k);
// tensor-core candidate scan (bf16x2 tcgen05) + exact re-scoring + certificate / exact fix-up: the exact path's scores and indices
bool topk_tc_supported(int D, int k);
size_t topk_tc_workspace_bytes(int B, int N, int D, int k);
int launch_topk_ip_tc(const float* q, const float* db, const bf16* db_hi, const bf16* db_lo, float db_norm_max, int B, int N, int D, int k,
                      float* scores, int64_t* idx, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_topk_ip(const float* q, const float* db, int B, int N, int D, int k, float* scores, int64_t* idx, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int launch_select_caption_rows(const float* scores, const int64_t* idx, int B, int k_searched, const int64_t* cap_row_start,
                               const int64_t* cap_row_ids, int top_i, int top_k, int64_t* rows_out, cudaStream_t st);
int launch_gather_caption_rows(const float* cap_db, const int64_t* rows, int n_rows, int D, float* out, cudaStream_t st);
int launch_gather_attention_add(const float* q, const float* cap_db, const int64_t* rows, int B, int top_k, int D, const float* attn_w,
                                const float* attn_b, float* out, cudaStream_t st);
int launch_gather_aggregate_add(const float* q, const float* cap_db, const int64_t* rows, int B, int top_k, int D, int aggregation,
                                float* out, cudaStream_t st);

}  // namespace gic
