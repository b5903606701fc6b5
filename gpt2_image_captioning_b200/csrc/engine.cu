// Host side of libgic_b200.so: the opaque engine (packed weights), workspace carving, the generate drivers
// (mapper -> prefill -> KV-cached decode loop, captured in a CUDA graph) and the extern "C" entry points.
// See include/gic_b200.h for the contract and the reference lines each entry point replaces.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "kernels.cuh"

namespace gic {

// ---------------------------------------------------------------------------------------------------------------
// error state
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ---------------------------------------------------------------------------------------------------------------
// packed weights
// ---------------------------------------------------------------------------------------------------------------
// W tensor maps are keyed by their box height = block_n
static const int kNumBoxes = 6;
static const int kBoxRows[kNumBoxes] = {32, 64, 96, 128, 192, 256};  // (96: half of a CTA pair's 192-wide tile)
static int box_rows_index(int rows) { for (int i = 0; i < kNumBoxes; ++i) if (kBoxRows[i] == rows) return i; return -1; }

struct Linear {  // y = x . W^T + b with W packed as [N,K] K-major
  int N = 0, K = 0;
  float* w_f32 = nullptr;
  bf16* w_hi = nullptr;
  bf16* w_lo = nullptr;
  float* bias = nullptr;
  float* colsum = nullptr;     // non-null: a LayerNorm is folded into this layer (w = gamma * W, bias = b + W . beta), see launch_fold_ln
  TmaDesc tm_hi[kNumBoxes], tm_lo[kNumBoxes];  // per box height in kBoxRows
};
struct Norm { float* w = nullptr; float* b = nullptr; };
struct GptLayer { Norm ln1, ln2; Linear attn, proj, fc, fc2; };
struct TfmLayer { Norm n1, n2; Linear in_proj, out_proj, lin1, lin2; };

// activation buffer in the engine's arithmetic mode (F32: f32; BF16: hi; BF16X2: hi + lo)
struct Act {
  float* f32 = nullptr;
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  ActOut out() const { ActOut o; o.f32 = f32; o.hi = hi; o.lo = lo; return o; }
};

struct Workspace {
  int B = 0, rows = 0, max_new = 0, beams = 1, P = 0, t_max = 0, m_max = 0;
  Act x_in, map_hidden, a, o, f;   // GEMM A operands
  float* map_lin = nullptr;        // transformer mapper: Linear output [B, Hl*d] fp32
  float* prefix = nullptr;         // mapper output [B, P_img, d] fp32
  float* h = nullptr;              // residual stream [m_max, d] fp32
  float* h_dec = nullptr;          // decode residual [rows, d] fp32
  float* qkv_f32 = nullptr;        // [m_max, 3d]  (F32; BF16X2: the transformer mapper's encoder layers only)
  bf16* qkv_bf16 = nullptr;        //              (BF16: bf16; BF16X2: IEEE half behind the same 2-byte type -- GPT-2 layers)
  void* kv = nullptr;              // [L][2][rows][H][t_max][64]  (F32: float; BF16: bf16; BF16X2: IEEE half)
  size_t kv_layer_elems = 0;       // elements of one K (or V) plane of one layer
  float* logits = nullptr;         // [rows, V] fp32 (F32 mode)
  float* part_val = nullptr; int* part_idx = nullptr; int n_parts_max = 0;  // [rows][n_parts_max]
  float* part_val2 = nullptr;      // [rows][n_parts_max] runner-up value per slot (bf16x2 engine: exactly re-scored head)
  // ... and the re-scoring state: ln_f(h) in fp32, the (row, column) candidate list, its counters {pairs, flagged rows}, the rows that did
  // not fit the list, and the packed (value, ~column) result per row
  float* rs_a = nullptr; int2* rs_pairs = nullptr; int rs_pair_cap = 0; int* rs_counters = nullptr; int* rs_row_flag = nullptr; int* rs_flag_rows = nullptr;
  unsigned long long* rs_best = nullptr;
  float* splitk_ws = nullptr; size_t splitk_ws_floats = 0; int* splitk_counters = nullptr;  // split-K partials / per-tile arrival counters
  float2* ln_stats = nullptr; int ln_parts_max = 0;  // [ln_parts_max][m_max] row (sum, sum of squares) partials (folded LayerNorm)
  int64_t* ids = nullptr;          // [rows, max_new]
  unsigned char* finished = nullptr; int* first_eos = nullptr;
  int *d_step = nullptr, *d_pos = nullptr, *done_counter = nullptr;
  SampleParams* sample_params = nullptr;  // temperature / top_p / seed of the running gic_generate_sample call (read by the captured sampling launches)
  int *fin_counter = nullptr, *all_done = nullptr;  // [sub-batch] rows finished in the current step / every row has emitted EOS
  int* live_rows = nullptr;        // [sub-batch] unfinished rows after the last step (the host sizes the compacted batch from it)
  // finished-row compaction (greedy): slot -> caption row map of the compacted batch (null while the batch is whole), scratch for the move
  const int* row_map = nullptr;
  int* row_map_buf = nullptr; int* src_slot = nullptr;
  float* h_tmp = nullptr; bf16* a_tmp_hi = nullptr; bf16* a_tmp_lo = nullptr; float2* stats_tmp = nullptr;
  // beam search only
  void* kv2 = nullptr;             // second KV cache (reorder target; the two swap every step); null with the ancestry table
  const int* anc = nullptr; int anc_ld = 0;  // beam search without reordering: ancestry table of the current step (see attention.cu)
  BeamState beam;
  size_t bytes = 0;
};

}  // namespace gic

using namespace gic;

static const int kNumHeadTiles = 5;
static const int kHeadTiles[kNumHeadTiles] = {32, 64, 128, 192, 256};

struct gic_engine {
  gic_config cfg;
  int d = 0, L = 0, H = 0, V = 0, P_img = 0, P_task = 0, E = 0;
  bool split = false;  // BF16X2
  bool beam_fused_head = false;  // beam search: log-sum-exp + top-2K candidates from the LM-head epilogue (no [rows, V] logits); GIC_BEAM_LOGITS=1: off
  bool beam_indirect = false;  // BF16 beam search: no KV reorder, decode attention reads through an ancestry table (GIC_BEAM_REORDER=1: gather)
  bool fuse_ln = false;  // BF16 / BF16X2: ln_1 / ln_2 folded into the GEMM that follows them (no LayerNorm launches inside the GPT-2 blocks)
  bool use_splitk = false;  // GIC_SPLITK=1 turns the K split of the decode-size residual GEMMs on.  Off by default: measured round 1
                            // (profiles/r1w_microbench.txt), fc2 with a 3-way K split + last-CTA reduction takes 22 us against 15 us unsplit
  bool fuse_lnf = false; // ... and ln_f into the LM head (GIC_LNF_FUSE=1).  Off by default: measured round 1, the folded head re-reads the
                         // row statistics and column sums for each of its ~10 tiles per CTA and costs 79 us against 3.8 + 55 us
  bool rescore_head = false;  // BF16X2 greedy: the LM head runs with single bf16 operands (1 MMA per product) and the candidates within the rounding
                              // margin of its maximum are re-scored exactly in fp32 (lm_head_rescore_kernel).  GIC_X2_HEAD_FULL=1: the 3-MMA head
  float* wte_norm = nullptr;  // [V] |wte[n]|_2: the per-column error bound of the single-MMA head
  float* slot_norm_max[kNumHeadTiles] = {};  // per LM-head tile width (kHeadTiles): the largest norm inside each (tile, column-parity) slot
  bf16* wte_gather = nullptr;  // BF16: unfolded bf16 embedding table for the next-token gather
  bool tc = false;     // tensor-core modes (BF16 / BF16X2)
  std::vector<void*> allocs;
  size_t weight_bytes = 0;
  bool gpt_loaded = false, mapper_loaded = false;
  // GPT-2
  Linear lm_head;       // wte as [V,d]
  float* wte_f32 = nullptr;  // embedding table (F32 / BF16X2); BF16 gathers from lm_head.w_hi
  float* wpe = nullptr;
  Norm lnf;
  std::vector<GptLayer> layers;
  float* task_prefix = nullptr;  // [P_task, d]
  // mappers
  Linear map1, map2;             // MLP
  Linear tfm_linear; float* tfm_prefix_const = nullptr; std::vector<TfmLayer> tfm_layers;
  // decode-step CUDA graph cache (one entry)
  struct GraphEntry { int M; cudaGraphExec_t exec; int nodes; };  // one chunk of decode steps over M activation slots
  std::vector<GraphEntry> graphs;  // M = the whole batch, plus one entry per compacted size used so far
  void* graph_ws = nullptr; int graph_B = 0, graph_max_new = 0;
  const float* graph_logits = nullptr;  // logits tap baked into the captured launches (sampling; null = the greedy graphs)
  bool use_graph = true;
  int graph_steps = 1;  // decode steps held by each graph
  unsigned int graph_trace_gen = 0;        // trace_generation() the graph was captured under
  // early exit (src/models.py:390-391): the host looks at the `all rows finished` flag of chunk c - 1 while chunk c runs
  int* h_done = nullptr;  // pinned [2]
  cudaEvent_t ev_done[2] = {nullptr, nullptr};
  // all generate work runs on this private stream (the caller's stream may be the legacy default stream, which cannot
  // be captured into a graph); it is forked from / joined to the caller's stream with events
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // decode runs as two half-batches on two streams (captured as two branches of one graph): the HBM-bound attention of
  // one half overlaps the operand-delivery-bound GEMMs of the other and each fills the other's launch gaps
  static constexpr int MAX_SUB = 8;
  cudaStream_t sub_stream[MAX_SUB] = {};  // [0] unused (= stream)
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_SUB] = {};
  // temperature / top-p sampling (gic_generate_sample): set for the duration of one call
  struct SampleCfg { bool on = false; float temperature = 1.f, top_p = 1.f; unsigned long long seed = 0; float* logits = nullptr; int ld = 0; } sample;
  int sub_batches = 1;  // GIC_SUBBATCH=n: decode as n row groups (whole 128-row GEMM tiles) on n streams, each kernel limited to 1/n of the SMs
  // per-kernel-class CUDA-event profiling (bench.py roofline leg); generate runs eagerly while enabled
  bool profiling = false;
  struct ProfRec { const char* cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
};

namespace gic {

static StepTrace g_step_trace = {nullptr, 0, 0};
static unsigned int g_trace_gen = 0;
StepTrace trace_desc() {
  StepTrace t = g_step_trace;
  if (t.buf) ++g_step_trace.slot;  // one slot per launch
  return t;
}
bool trace_on() { return g_step_trace.buf != nullptr; }
unsigned int trace_generation() { return g_trace_gen; }

static thread_local int g_cta_limit = 0;
int cta_limit() { return g_cta_limit; }
void set_cta_limit(int ctas) { g_cta_limit = ctas < 0 ? 0 : ctas; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("GIC_NO_PDL"); on = (v && v[0] == '1') ? 0 : 1; }
  return on == 1;
}

// kernels launched by this library (graph replays count their kernel nodes) -- bench.py's `gpu_launches`
// (atomic + a per-thread count: engines may be driven from several host threads, one stream each)
static std::atomic<unsigned long long> g_compactions{0};  // finished-row compactions performed (gic_compaction_count)
static std::atomic<unsigned long long> g_launches{0};
static thread_local unsigned long long tl_launches = 0;
void note_launch() { ++tl_launches; g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfScope {
  gic_engine* e; cudaStream_t st; bool on;
  ProfScope(const gic_engine* ce, const char* cat, cudaStream_t s) : e(const_cast<gic_engine*>(ce)), st(s), on(ce->profiling) {
    if (!on) return;
    gic_engine::ProfRec r; r.cat = cat;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    e->prof.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(e->prof.back().b, st); }
};

static int dev_alloc(gic_engine* e, void** p, size_t bytes) {
  GIC_CHECK_CUDA(cudaMalloc(p, bytes));
  e->allocs.push_back(*p);
  e->weight_bytes += bytes;
  return GIC_OK;
}

static int copy_vec(gic_engine* e, float** dst, const float* src, size_t n, cudaStream_t st) {
  GIC_REQUIRE(src != nullptr, "null weight pointer");
  GIC_TRY(dev_alloc(e, (void**)dst, n * sizeof(float)));
  GIC_CHECK_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return GIC_OK;
}

// in: fp32 [N,K] (transpose = false, nn.Linear / wte) or [K,N] (transpose = true, HF Conv1D)
static int pack_linear(gic_engine* e, Linear* lin, const float* w, const float* bias, int N, int K, bool transpose, cudaStream_t st,
                       const float* ln_gamma = nullptr, const float* ln_beta = nullptr) {
  GIC_REQUIRE(w != nullptr, "null weight pointer");
  lin->N = N;
  lin->K = K;
  const size_t n = (size_t)N * K;
  ActOut out;
  if (!e->tc) {
    GIC_TRY(dev_alloc(e, (void**)&lin->w_f32, n * sizeof(float)));
    out.f32 = lin->w_f32;
  } else {
    GIC_TRY(dev_alloc(e, (void**)&lin->w_hi, n * sizeof(bf16)));
    out.hi = lin->w_hi;
    if (e->split) {
      GIC_TRY(dev_alloc(e, (void**)&lin->w_lo, n * sizeof(bf16)));
      out.lo = lin->w_lo;
    }
  }
  // source is [R,C]: transpose -> [C,R] = [N,K]
  const int R = transpose ? K : N, C = transpose ? N : K;
  GIC_TRY(launch_pack_weight(w, R, C, transpose, out, st, ln_gamma));
  if (ln_gamma) {
    GIC_REQUIRE(e->tc && ln_beta, "LayerNorm folding is a tensor-core-engine feature");
    GIC_TRY(dev_alloc(e, (void**)&lin->colsum, (size_t)N * sizeof(float)));
    GIC_TRY(dev_alloc(e, (void**)&lin->bias, (size_t)N * sizeof(float)));
    GIC_TRY(launch_fold_ln(lin->w_hi, w, transpose, ln_beta, bias, lin->colsum, lin->bias, N, K, st, lin->w_lo));
  } else if (bias) GIC_TRY(copy_vec(e, &lin->bias, bias, N, st));
  if (e->tc) {
    GIC_REQUIRE(K % 64 == 0, "tensor-core modes need K (%d) to be a multiple of 64", K);
    for (int i = 0; i < kNumBoxes; ++i) {
      GIC_TRY(make_tma_2d_bf16(&lin->tm_hi[i], lin->w_hi, N, K, K, kBoxRows[i]));
      if (e->split) GIC_TRY(make_tma_2d_bf16(&lin->tm_lo[i], lin->w_lo, N, K, K, kBoxRows[i]));
    }
  }
  return GIC_OK;
}

static int copy_norm(gic_engine* e, Norm* n, const float* w, const float* b, int d, cudaStream_t st) {
  GIC_TRY(copy_vec(e, &n->w, w, d, st));
  GIC_TRY(copy_vec(e, &n->b, b, d, st));
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------------------
struct Carver {
  unsigned char* base; size_t off = 0;
  explicit Carver(void* b) : base((unsigned char*)b) {}
  template <typename T> T* take(size_t n) {
    off = align_up(off, 1024);  // TMA-friendly alignment for every buffer
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

static void carve_act(const gic_engine* e, Carver& c, Act* a, size_t n) {
  if (!e->tc) a->f32 = c.take<float>(n);
  else {
    a->hi = c.take<bf16>(n);
    if (e->split) a->lo = c.take<bf16>(n);
  }
}

// the KV cache (and q | k | v) hold 2-byte elements: bf16 in the BF16 engine, IEEE half in the fused BF16X2 engine
static bool kv16(const gic_engine* e) { return e->cfg.dtype == GIC_DTYPE_BF16 || (e->split && e->fuse_ln); }

// tile shape of the beam-search LM head (EPI_BEAM): CTA pairs on 256-wide tiles when the row tiles pair up, else single CTAs
static void beam_head_shape(const gic_engine* e, int M, int* bn, int* pair) {
  *pair = (ceil_div(M, 128) % 2 == 0 && cta_limit() == 0) ? 1 : 0;
  *bn = (*pair || !e->split) ? 256 : 128;
}

static void carve(const gic_engine* e, void* base, int B, int max_new, int beams, Workspace* w) {
  const int d = e->d;
  w->B = B; w->beams = beams < 1 ? 1 : beams; w->rows = B * w->beams; w->max_new = max_new;
  w->P = e->P_img + e->P_task;
  w->t_max = w->P + max_new;
  const int S_map = e->cfg.mapper_kind == GIC_MAPPER_TRANSFORMER ? e->cfg.hidden_length + e->P_img : 0;
  int m = B * w->P;
  if (B * S_map > m) m = B * S_map;
  if (w->rows > m) m = w->rows;
  w->m_max = m;
  Carver c(base);
  carve_act(e, c, &w->x_in, (size_t)B * e->E);
  if (e->cfg.mapper_kind == GIC_MAPPER_MLP) carve_act(e, c, &w->map_hidden, (size_t)B * (e->P_img * d / 2));
  else w->map_lin = c.take<float>((size_t)B * e->cfg.hidden_length * d);
  w->prefix = c.take<float>((size_t)B * e->P_img * d);
  w->h = c.take<float>((size_t)m * d);
  w->h_dec = c.take<float>((size_t)w->rows * d);
  carve_act(e, c, &w->a, (size_t)m * d);
  carve_act(e, c, &w->o, (size_t)m * d);
  carve_act(e, c, &w->f, (size_t)m * 4 * d);
  if (kv16(e)) w->qkv_bf16 = c.take<bf16>((size_t)m * 3 * d);
  if (!kv16(e)) w->qkv_f32 = c.take<float>((size_t)m * 3 * d);
  else if (e->split && S_map > 0) w->qkv_f32 = c.take<float>((size_t)B * S_map * 3 * d);  // the transformer mapper's layers stay unfused
  w->kv_layer_elems = (size_t)w->rows * e->H * w->t_max * 64;
  const size_t kv_elems = (size_t)e->L * 2 * w->kv_layer_elems;
  if (kv16(e)) w->kv = c.take<bf16>(kv_elems);
  else w->kv = c.take<float>(kv_elems);
  if (!e->tc) {
    w->logits = c.take<float>((size_t)w->rows * e->V);
    w->n_parts_max = LMHEAD_F32_PARTS;
  } else {
    w->n_parts_max = 2 * ceil_div(e->V, 32);
    // beam search: log-softmax + top-2K come out of the LM-head epilogue (EPI_BEAM streams, below); full fp32 rows only with GIC_BEAM_LOGITS=1
    if (w->beams > 1 && !e->beam_fused_head) w->logits = c.take<float>((size_t)w->rows * e->V);
  }
  if (w->beams > 1) {
    BeamState& bs = w->beam;
    if (e->beam_indirect) {  // no second cache: attention follows the beams' ancestry instead
      bs.anc[0] = c.take<int>((size_t)w->rows * (max_new > 0 ? max_new : 1));
      bs.anc[1] = c.take<int>((size_t)w->rows * (max_new > 0 ? max_new : 1));
    } else if (kv16(e)) w->kv2 = c.take<bf16>(kv_elems);
    else w->kv2 = c.take<float>(kv_elems);
    const size_t nseq = (size_t)w->rows * (max_new > 0 ? max_new : 1);
    bs.B = B; bs.beams = w->beams; bs.max_new = max_new; bs.V = e->V; bs.eos = e->cfg.eos_token_id;
    for (int i = 0; i < 2; ++i) { bs.run_seq[i] = c.take<int>(nseq); bs.fin_seq[i] = c.take<int>(nseq); }
    bs.run_score = c.take<float>(w->rows); bs.fin_score = c.take<float>(w->rows);
    bs.fin_flag = c.take<unsigned char>(w->rows); bs.fin_len = c.take<int>(w->rows);
    bs.unsat = c.take<unsigned char>(B);
    bs.beam_idx = c.take<int>(w->rows); bs.next_tok = c.take<int>(w->rows);
    bs.cand_score = c.take<float>((size_t)2 * w->rows); bs.cand_idx = c.take<int>((size_t)2 * w->rows);
    bs.lse = c.take<float>(w->rows); bs.row_val = c.take<float>((size_t)2 * w->beams * w->rows); bs.row_idx = c.take<int>((size_t)2 * w->beams * w->rows);
    if (e->beam_fused_head) {
      // (row, stream) slots of the two launch shapes: B rows after the prefill, B * beams rows per decode step
      int bn = 0, pair = 0;
      beam_head_shape(e, B, &bn, &pair);
      const size_t s0 = (size_t)B * gemm_topk_streams(B, e->V, bn, pair);
      beam_head_shape(e, w->rows, &bn, &pair);
      const size_t s1 = (size_t)w->rows * gemm_topk_streams(w->rows, e->V, bn, pair);
      bs.tk_slots = s0 > s1 ? s0 : s1;
      bs.tk_v = c.take<float>(bs.tk_slots * GEMM_BEAM_KEEP); bs.tk_i = c.take<int>(bs.tk_slots * GEMM_BEAM_KEEP);
      bs.tk_m = c.take<float>(bs.tk_slots); bs.tk_s = c.take<float>(bs.tk_slots);
    }
  }
  if (e->fuse_ln) {
    w->ln_parts_max = ceil_div(d, 32);
    w->ln_stats = c.take<float2>((size_t)w->ln_parts_max * m);
  }
  if (e->tc && e->fuse_ln) {  // split-K may be used by the decode-size residual GEMMs (few tiles, K up to 4 d)
    w->splitk_ws_floats = (size_t)4 * w->rows * d;
    w->splitk_ws = c.take<float>(w->splitk_ws_floats);
    w->splitk_counters = c.take<int>(4096);
  }
  w->part_val = c.take<float>((size_t)w->n_parts_max * w->rows);
  w->part_idx = c.take<int>((size_t)w->n_parts_max * w->rows);
  if (e->rescore_head && w->beams == 1) {
    w->part_val2 = c.take<float>((size_t)w->n_parts_max * w->rows);
    w->rs_a = c.take<float>((size_t)w->rows * d);
    w->rs_pair_cap = w->rows * 64;
    w->rs_pairs = c.take<int2>((size_t)w->rs_pair_cap);
    w->rs_counters = c.take<int>(2);
    w->rs_row_flag = c.take<int>(w->rows); w->rs_flag_rows = c.take<int>(w->rows);
    w->rs_best = c.take<unsigned long long>(w->rows);
  }
  w->ids = c.take<int64_t>((size_t)w->rows * (max_new > 0 ? max_new : 1));
  w->finished = c.take<unsigned char>(w->rows);
  w->first_eos = c.take<int>(w->rows);
  w->d_step = c.take<int>(gic_engine::MAX_SUB);  // [sub-batch]
  w->d_pos = c.take<int>(gic_engine::MAX_SUB);
  w->done_counter = c.take<int>(gic_engine::MAX_SUB);
  w->fin_counter = c.take<int>(gic_engine::MAX_SUB);
  w->all_done = c.take<int>(gic_engine::MAX_SUB);
  w->live_rows = c.take<int>(gic_engine::MAX_SUB);
  w->sample_params = c.take<SampleParams>(1);
  if (w->beams == 1) {
    w->row_map_buf = c.take<int>(w->rows); w->src_slot = c.take<int>(w->rows);
    w->h_tmp = c.take<float>((size_t)w->rows * d);
    if (e->fuse_ln) {
      w->a_tmp_hi = c.take<bf16>((size_t)w->rows * d);
      if (e->split) w->a_tmp_lo = c.take<bf16>((size_t)w->rows * d);
      w->stats_tmp = c.take<float2>(w->rows);
    }
  }
  w->bytes = align_up(c.off, 1024) + 1024;
}

// ---------------------------------------------------------------------------------------------------------------
// building blocks
// ---------------------------------------------------------------------------------------------------------------
// out = epi(A . W^T + b).  A: Act with M rows of K (dense).  Outputs follow `out` (row stride ld_out).
// folded-LayerNorm plumbing of one GEMM call: where the A rows' statistics come from / where this GEMM's output statistics go
struct LnIo {
  const float2* stats_in = nullptr; int parts_in = 0; long stats_ld = 0; int row_mul = 1, row_off = 0;
  float2* stats_out = nullptr;
  long a_row_stride = 0;  // A rows are `a_row_stride` elements apart (0: dense)
  float* splitk_ws = nullptr; size_t splitk_ws_floats = 0; int* splitk_counters = nullptr;  // set: this GEMM may split K
};
// number of statistics parts a residual GEMM leaves behind: one per 32 output columns, whatever the tile width or M
// (so that a row's LayerNorm statistics are summed in the same order in any batch)
static int residual_stats_parts(const gic_engine* e, int M) { (void)M; return ceil_div(e->d, 32); }

static int linear(const gic_engine* e, const Linear& lin, const Act& A, int M, int epilogue, ActOut out, int ld_out, cudaStream_t st,
                  float* part_val = nullptr, int* part_idx = nullptr, int* n_parts = nullptr, int part_ld = 0, const LnIo* ln = nullptr,
                  bool out_f16 = false) {
  if (!e->tc) {
    GIC_REQUIRE(out.f32 != nullptr, "fp32 linear needs an fp32 output");
    return launch_sgemm_nt(A.f32, lin.K, lin.w_f32, lin.bias, out.f32, ld_out, M, lin.N, lin.K, epilogue, st);
  }
  GemmBf16Args g;
  int bn = 0, sk = 1;
  if (ln && ln->splitk_ws && !part_val && e->use_splitk) {
    sk = gemm_bf16_split_k_for(lin.N, lin.K);
    if ((size_t)sk * M * lin.N > ln->splitk_ws_floats) sk = 1;
  }
  // CTA pairs (cta_group::2) exist for the fused-LayerNorm engine's GEMMs: folded qkv / fc, residual + statistics, LM-head argmax
  const bool folded_in = ln && ln->stats_in, stats_out = ln && ln->stats_out;
  // (bf16x2: folded qkv -> fp16, folded fc + GELU -> hi + lo, residual + hi / lo copies + statistics, LM-head argmax)
  const bool pair_kernel = e->fuse_ln && sk == 1 &&
                           ((part_val && !out.f32 && !folded_in) ||
                            (!part_val && folded_in && out.hi && !out.f32 && (e->split ? ((epilogue == EPI_NONE && out_f16) || (epilogue == EPI_GELU && out.lo)) : (epilogue == EPI_NONE || epilogue == EPI_GELU))) ||
                            (!part_val && stats_out && epilogue == EPI_RESIDUAL));
  int pair = 0;
  gemm_bf16_pick(M, lin.N, lin.K, e->split ? 1 : 0, sk, &bn, pair_kernel ? &pair : nullptr);
  if (sk > 1) { g.split_k = sk; g.splitk_ws = ln->splitk_ws; g.splitk_counters = ln->splitk_counters; }
  g.pair = pair;
  // the wide bf16x2 tiles exist for the fused GPT-2 layer's GEMMs, the LM head and plain fp32 outputs only (gemm_tcgen05.cu
  // GIC_GEMM_VARIANTS_SPLIT_WIDE); everything else in that mode (mapper GEMMs) stays on the narrow tiles
  if (e->split && !pair && bn > 64) {
    const bool wide_ok = (folded_in && ((epilogue == EPI_NONE && out_f16) || (epilogue == EPI_GELU && out.lo))) || (stats_out && epilogue == EPI_RESIDUAL) ||
                         (part_val && !folded_in) || (!part_val && !folded_in && !stats_out && epilogue == EPI_NONE && out.f32 && !out.hi);
    if (!wide_ok) bn = 64;
  }
  const int bi = box_rows_index(pair ? bn / 2 : bn);
  GIC_REQUIRE(bi >= 0, "no W tensor map for box height %d", pair ? bn / 2 : bn);
  GIC_TRY(make_tma_2d_bf16(&g.a_hi, A.hi, M, lin.K, (ln && ln->a_row_stride) ? ln->a_row_stride : lin.K, 128));
  g.w_hi = lin.tm_hi[bi];
  if (e->split) {
    GIC_TRY(make_tma_2d_bf16(&g.a_lo, A.lo, M, lin.K, (ln && ln->a_row_stride) ? ln->a_row_stride : lin.K, 128));
    g.w_lo = lin.tm_lo[bi];
  }
  g.M = M; g.N = lin.N; g.K = lin.K; g.block_n = bn; g.split = e->split ? 1 : 0; g.epilogue = epilogue; g.bias = lin.bias;
  g.w_static = 1;  // packed at load time
  g.out = out; g.out_f16 = out_f16 ? 1 : 0; g.ld_out = ld_out; g.part_val = part_val; g.part_idx = part_idx; g.part_ld = part_ld;
  GIC_REQUIRE((lin.colsum != nullptr) == (ln != nullptr && ln->stats_in != nullptr), "linear: folded-LayerNorm weights and row statistics must come together");
  if (ln) {
    g.ln_stats = ln->stats_in; g.ln_parts = ln->parts_in; g.ln_stats_ld = ln->stats_ld; g.ln_row_mul = ln->row_mul; g.ln_row_off = ln->row_off;
    g.ln_colsum = lin.colsum; g.stats_out = ln->stats_out;
  }
  if (n_parts) *n_parts = 2 * ceil_div(lin.N, bn);  // one (value, index) slot per (tile, column-parity epilogue warp)
  return launch_gemm_bf16(g, st);
}

// q | k | v of the UNFUSED layers (fp32 / unfused engines, and the transformer mapper's encoder layers in every mode)
static ActOut qkv_out(const gic_engine* e, const Workspace& w) {
  ActOut o;
  if (e->cfg.dtype == GIC_DTYPE_BF16) o.hi = w.qkv_bf16;
  else o.f32 = w.qkv_f32;
  return o;
}

// attention sub-step shared by both layer variants
static int attention(const gic_engine* e, const Workspace& w, int l, int M, bool prefill, cudaStream_t st) {
  ProfScope ps(e, prefill ? "attn_prefill" : "attn_decode", st);
  if (e->split && e->fuse_ln) {  // fp16 q | k | v and cache, bf16 hi + lo output
    bf16* kc = (bf16*)w.kv + (size_t)(2 * l) * w.kv_layer_elems;
    bf16* vc = kc + w.kv_layer_elems;
    if (prefill) return launch_attn_prefill_f16(w.qkv_bf16, kc, vc, w.o.hi, w.o.lo, w.B, w.P, e->H, w.t_max, w.beams, st);
    if (w.anc) return launch_attn_decode_indirect(w.qkv_bf16, kc, vc, w.o.hi, w.d_pos, M, e->H, w.t_max, w.anc, w.anc_ld, w.P, w.beams, st, w.o.lo);
    return launch_attn_decode_f16(w.qkv_bf16, kc, vc, w.o.hi, w.o.lo, w.d_pos, M, e->H, w.t_max, st, w.row_map);
  }
  if (e->cfg.dtype == GIC_DTYPE_BF16) {
    bf16* kc = (bf16*)w.kv + (size_t)(2 * l) * w.kv_layer_elems;
    bf16* vc = kc + w.kv_layer_elems;
    if (prefill) return launch_attn_prefill<bf16>(w.qkv_bf16, kc, vc, w.o.out(), w.B, w.P, e->H, w.t_max, w.beams, st);
    if (w.anc) return launch_attn_decode_indirect(w.qkv_bf16, kc, vc, w.o.hi, w.d_pos, M, e->H, w.t_max, w.anc, w.anc_ld, w.P, w.beams, st);
    return launch_attn_decode<bf16>(w.qkv_bf16, kc, vc, w.o.out(), w.d_pos, M, e->H, w.t_max, st, w.row_map);
  }
  float* kc = (float*)w.kv + (size_t)(2 * l) * w.kv_layer_elems;
  float* vc = kc + w.kv_layer_elems;
  if (prefill) return launch_attn_prefill<float>(w.qkv_f32, kc, vc, w.o.out(), w.B, w.P, e->H, w.t_max, w.beams, st);
  return launch_attn_decode<float>(w.qkv_f32, kc, vc, w.o.out(), w.d_pos, M, e->H, w.t_max, st, w.row_map);
}

// one GPT-2 block over M rows of the residual stream `h` (HF GPT2Block.forward :262-309).
// bf16 engine (fuse_ln): w.a holds bf16(h) and w.ln_stats its per-row (sum, sum of squares) in `parts_in` parts on entry;
// both are refreshed by the residual GEMMs, so the block is 5 launches: qkv, attention, proj, fc, fc2.
static int gpt_layer(const gic_engine* e, const Workspace& w, int l, float* h, int M, bool prefill, cudaStream_t st, int parts_in = 0) {
  const GptLayer& Lw = e->layers[l];
  const int d = e->d;
  ActOut hres; hres.f32 = h;
  if (e->fuse_ln) {
    const int parts_res = residual_stats_parts(e, M);
    LnIo in; in.stats_in = w.ln_stats; in.parts_in = parts_in; in.stats_ld = w.m_max;
    LnIo res; res.stats_out = w.ln_stats; res.stats_ld = w.m_max;
    if (!prefill) { res.splitk_ws = w.splitk_ws; res.splitk_ws_floats = w.splitk_ws_floats; res.splitk_counters = w.splitk_counters; }
    ActOut hres2 = hres; hres2.hi = w.a.hi; hres2.lo = w.a.lo;  // (lo: bf16x2 only)
    ActOut qo; qo.hi = w.qkv_bf16;
    { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_qkv", st);
      GIC_TRY(linear(e, Lw.attn, w.a, M, EPI_NONE, qo, 3 * d, st, nullptr, nullptr, nullptr, 0, &in, e->split)); }
    GIC_TRY(attention(e, w, l, M, prefill, st));
    { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_proj", st);
      GIC_TRY(linear(e, Lw.proj, w.o, M, EPI_RESIDUAL, hres2, d, st, nullptr, nullptr, nullptr, 0, &res)); }
    in.parts_in = parts_res;
    { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_fc", st);
      GIC_TRY(linear(e, Lw.fc, w.a, M, EPI_GELU, w.f.out(), 4 * d, st, nullptr, nullptr, nullptr, 0, &in)); }
    { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_fc2", st);
      GIC_TRY(linear(e, Lw.fc2, w.f, M, EPI_RESIDUAL, hres2, d, st, nullptr, nullptr, nullptr, 0, &res)); }
    return GIC_OK;
  }
  { ProfScope ps(e, "layernorm", st); GIC_TRY(launch_layernorm(h, d, Lw.ln1.w, Lw.ln1.b, w.a.out(), M, d, st)); }
  { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_qkv", st); GIC_TRY(linear(e, Lw.attn, w.a, M, EPI_NONE, qkv_out(e, w), 3 * d, st)); }
  GIC_TRY(attention(e, w, l, M, prefill, st));
  { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_proj", st); GIC_TRY(linear(e, Lw.proj, w.o, M, EPI_RESIDUAL, hres, d, st)); }
  { ProfScope ps(e, "layernorm", st); GIC_TRY(launch_layernorm(h, d, Lw.ln2.w, Lw.ln2.b, w.a.out(), M, d, st)); }
  { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_fc", st); GIC_TRY(linear(e, Lw.fc, w.a, M, EPI_GELU, w.f.out(), 4 * d, st)); }
  { ProfScope ps(e, prefill ? "prefill_gemm" : "gemm_fc2", st); GIC_TRY(linear(e, Lw.fc2, w.f, M, EPI_RESIDUAL, hres, d, st)); }
  return GIC_OK;
}

// every layer over M rows; the statistics parts of layer 0's input come from the caller (1 after row_stats / finalize)
static int gpt_layers(const gic_engine* e, const Workspace& w, float* h, int M, bool prefill, cudaStream_t st) {
  for (int l = 0; l < e->L; ++l) GIC_TRY(gpt_layer(e, w, l, h, M, prefill, st, l == 0 ? 1 : residual_stats_parts(e, M)));
  return GIC_OK;
}

// ln_f on `rows` rows (row r at h + r*stride) -> LM head -> (val, idx) partials -> finalize (token, EOS rules, next input).
// body_rows: the M the layers ran over (prefill: B*P, rows = B picks position P-1 of every sequence).
static int lm_head_and_token(const gic_engine* e, const Workspace& w, const float* h_buf, long first_off, long row_stride, int rows, int body_rows,
                             float* logits_tap, cudaStream_t st) {
  const int d = e->d;
  const float* h = h_buf + first_off;
  int n_parts = 0;
  bool packed = false;  // the token arrives as w.rs_best (exactly re-scored head) instead of argmax partials
  // row stride of the logits tap: V for the per-step tap of the tests; the sampling scratch of the tensor-core engines is padded to a
  // multiple of 32 floats so that the GEMM epilogue stores whole aligned 16-byte vectors and the sampler's bulk copies start aligned
  const int ld_tap = (e->sample.on && e->tc) ? e->sample.ld : e->V;
  if (e->fuse_ln && e->fuse_lnf) {
    // rows of bf16(h) in w.a at the same (stride, offset) as h in its buffer; statistics indexed by the body row
    const long off = first_off;
    Act a = w.a; a.hi += off;
    LnIo in; in.stats_in = w.ln_stats; in.parts_in = e->L > 0 ? residual_stats_parts(e, body_rows) : 1; in.stats_ld = w.m_max;
    in.row_mul = (int)(row_stride / d); in.row_off = (int)(off / d); in.a_row_stride = row_stride;
    ActOut o; o.f32 = logits_tap;  // null on the product path: logits never reach HBM
    ProfScope ps(e, "lm_head", st);
    GIC_TRY(linear(e, e->lm_head, a, rows, EPI_NONE, o, ld_tap, st, w.part_val, w.part_idx, &n_parts, w.n_parts_max, &in));
  } else {
    { ProfScope ps(e, "layernorm", st); GIC_TRY(launch_layernorm(h, row_stride, e->lnf.w, e->lnf.b, w.a.out(), rows, d, st)); }
    if (e->rescore_head && !logits_tap && w.rs_best) {
      // single-MMA head on the hi halves + exact re-scoring of the near-maximal candidates (lmhead.cu)
      GemmBf16Args g;
      int bn = 0, pair = 0;
      gemm_bf16_pick(rows, e->V, d, 0, 1, &bn, &pair);
      const int bi = box_rows_index(pair ? bn / 2 : bn);
      GIC_REQUIRE(bi >= 0, "no W tensor map for box height %d", pair ? bn / 2 : bn);
      GIC_TRY(make_tma_2d_bf16(&g.a_hi, w.a.hi, rows, d, d, 128));
      g.w_hi = e->lm_head.tm_hi[bi];
      g.M = rows; g.N = e->V; g.K = d; g.block_n = bn; g.pair = pair; g.w_static = 1;
      g.part_val = w.part_val; g.part_idx = w.part_idx; g.part_val2 = w.part_val2; g.part_ld = w.n_parts_max;
      n_parts = 2 * ceil_div(e->V, bn);
      { ProfScope ps(e, "lm_head", st); GIC_TRY(launch_gemm_bf16(g, st)); }
      int ti = -1;
      for (int i = 0; i < kNumHeadTiles; ++i) if (kHeadTiles[i] == bn) ti = i;
      GIC_REQUIRE(ti >= 0, "no slot norms for LM-head tile width %d", bn);
      RescoreArgs ra;
      ra.h = h; ra.h_row_stride = row_stride; ra.lnw = e->lnf.w; ra.lnb = e->lnf.b;
      ra.wte = e->wte_f32; ra.wte_norm = e->wte_norm; ra.slot_norm_max = e->slot_norm_max[ti];
      ra.part_val = w.part_val; ra.part_idx = w.part_idx; ra.part_val2 = w.part_val2; ra.n_parts = n_parts; ra.part_ld = w.n_parts_max; ra.block_n = bn;
      ra.rows = rows; ra.V = e->V; ra.d = d;
      ra.a_f32 = w.rs_a; ra.pairs = w.rs_pairs; ra.pair_cap = w.rs_pair_cap; ra.row_budget = 1024;
      ra.pair_count = w.rs_counters; ra.flag_count = w.rs_counters + 1; ra.row_flag = w.rs_row_flag; ra.flag_rows = w.rs_flag_rows;
      ra.best = w.rs_best; ra.row_map = w.row_map; ra.step_trace = StepTrace{nullptr, 0, 0};
      { ProfScope ps(e, "lm_head_rescore", st); GIC_TRY(launch_lm_head_rescore(ra, st)); }
      packed = true;
      n_parts = 1;
    } else if (!e->tc) {
      float* lg = logits_tap ? logits_tap : w.logits;
      ActOut o; o.f32 = lg;
      { ProfScope ps(e, "lm_head", st); GIC_TRY(linear(e, e->lm_head, w.a, rows, EPI_NONE, o, e->V, st)); }
      { ProfScope ps(e, "argmax", st); GIC_TRY(launch_argmax_partials(lg, rows, e->V, w.part_val, w.part_idx, w.n_parts_max, st)); }
      n_parts = LMHEAD_F32_PARTS;
    } else {
      ActOut o; o.f32 = logits_tap;  // null on the product path: logits never reach HBM
      ProfScope ps(e, "lm_head", st);
      GIC_TRY(linear(e, e->lm_head, w.a, rows, EPI_NONE, o, ld_tap, st, w.part_val, w.part_idx, &n_parts, w.n_parts_max));
    }
  }
  if (e->sample.on) {
    // sampling: the step's logits were tapped into e->sample.logits; the drawn token replaces the argmax partials
    GIC_REQUIRE(logits_tap != nullptr, "sampling needs the logits tap");
    ProfScope pss(e, "sample", st);
    GIC_TRY(launch_sample_top_p(logits_tap, rows, e->V, e->sample.temperature, e->sample.top_p, e->sample.seed, w.d_step, -1, w.part_val, w.part_idx,
                                w.n_parts_max, st, w.sample_params, ld_tap));
    n_parts = 1;
  }
  ProfScope psf(e, "finalize", st);
  FinalizeArgs fa;
  fa.part_val = w.part_val; fa.part_idx = w.part_idx; fa.n_parts = n_parts; fa.part_ld = w.n_parts_max;
  fa.B = rows; fa.d = d; fa.eos = e->cfg.eos_token_id; fa.max_new = w.max_new; fa.P = w.P; fa.n_pos = e->cfg.n_positions;
  fa.d_step = w.d_step; fa.d_pos = w.d_pos; fa.done_counter = w.done_counter; fa.fin_counter = w.fin_counter; fa.all_done = w.all_done;
  fa.finished = w.finished; fa.first_eos = w.first_eos; fa.ids_out = w.ids; fa.row_map = w.row_map; fa.live_rows = w.live_rows;
  fa.wte_f32 = e->wte_f32; fa.wte_bf16 = e->wte_f32 ? nullptr : e->wte_gather;
  fa.wpe = e->wpe; fa.h_next = w.h_dec;
  if (packed) { fa.packed_best = w.rs_best; fa.rescore_counters = w.rs_counters; fa.n_parts = 0; }
  fa.hb_next = e->fuse_ln ? w.a.hi : nullptr; fa.hb_next_lo = e->fuse_ln ? w.a.lo : nullptr; fa.stats_next = e->fuse_ln ? w.ln_stats : nullptr;
  return launch_finalize_token(fa, st);
}

static int decode_step(const gic_engine* e, const Workspace& w, float* logits_tap, cudaStream_t st) {
  GIC_TRY(gpt_layers(e, w, w.h_dec, w.rows, false, st));
  return lm_head_and_token(e, w, w.h_dec, 0, e->d, w.rows, w.rows, logits_tap, st);
}

// view of rows [row0, row0 + nrows) of a workspace, with its own device-side step / position counters (`sub`)
static Workspace slice_rows(const gic_engine* e, const Workspace& w, int row0, int nrows, int sub) {
  Workspace s = w;
  const size_t d = e->d;
  auto shift = [&](Act& a, size_t width) {
    if (a.f32) a.f32 += (size_t)row0 * width;
    if (a.hi) a.hi += (size_t)row0 * width;
    if (a.lo) a.lo += (size_t)row0 * width;
  };
  s.rows = nrows; s.B = nrows;
  shift(s.a, d); shift(s.o, d); shift(s.f, 4 * d);
  if (s.qkv_f32) s.qkv_f32 += (size_t)row0 * 3 * d;
  if (s.qkv_bf16) s.qkv_bf16 += (size_t)row0 * 3 * d;
  s.h_dec += (size_t)row0 * d;
  const size_t kv_row = (size_t)e->H * w.t_max * 64;  // the per-layer plane stride (kv_layer_elems) keeps the full row count
  if (kv16(e)) s.kv = (bf16*)w.kv + (size_t)row0 * kv_row;
  else s.kv = (float*)w.kv + (size_t)row0 * kv_row;
  if (s.logits) s.logits += (size_t)row0 * e->V;
  if (s.ln_stats) s.ln_stats += row0;
  s.part_val += (size_t)w.n_parts_max * row0;
  s.part_idx += (size_t)w.n_parts_max * row0;
  if (s.part_val2) s.part_val2 += (size_t)w.n_parts_max * row0;
  s.rs_best = nullptr;  // row groups share one candidate list: they take the 3-MMA head instead of the re-scored one
  s.ids += (size_t)row0 * w.max_new;
  s.finished += row0; s.first_eos += row0;
  s.d_step += sub; s.d_pos += sub; s.done_counter += sub; s.fin_counter += sub; s.all_done += sub; s.live_rows += sub;
  return s;
}

// one decode step for every row: as `sub_batches` row groups of whole 128-row GEMM tiles, each on its own stream and with
// every persistent kernel limited to its share of the SMs (see gic_engine), when the batch is large enough
static int decode_step_all(gic_engine* e, const Workspace& w, float* logits_tap, cudaStream_t st) {
  int S = e->sub_batches;
  const int tiles = (w.rows + 127) / 128;
  if (S > tiles) S = tiles;
  if (S < 2 || logits_tap) {
    // GIC_CTA_LIMIT=n (experiment knob): every persistent kernel of the decode chain takes at most n CTAs, so that the chains of
    // several batches in flight (inflight.py) run side by side instead of taking turns on the whole GPU
    static const int lim = [] { const char* v = getenv("GIC_CTA_LIMIT"); return v ? atoi(v) : 0; }();
    if (lim <= 0) return decode_step(e, w, logits_tap, st);
    set_cta_limit(lim);
    const int r = decode_step(e, w, logits_tap, st);
    set_cta_limit(0);
    return r;
  }
  const int rows_per = ((tiles + S - 1) / S) * 128;
  S = (w.rows + rows_per - 1) / rows_per;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  GIC_CHECK_CUDA(cudaEventRecord(e->ev_fork, st));
  set_cta_limit(sms / S);
  int r = GIC_OK;
  for (int i = 0; i < S && r == GIC_OK; ++i) {
    const int row0 = i * rows_per, n = (w.rows - row0 < rows_per) ? w.rows - row0 : rows_per;
    const Workspace sw = slice_rows(e, w, row0, n, i);
    cudaStream_t si = i == 0 ? st : e->sub_stream[i];
    if (i > 0 && cudaStreamWaitEvent(si, e->ev_fork, 0) != cudaSuccess) { r = GIC_ERR_CUDA; break; }
    r = decode_step(e, sw, nullptr, si);
    if (r == GIC_OK && i > 0 && (cudaEventRecord(e->ev_join[i], si) != cudaSuccess || cudaStreamWaitEvent(st, e->ev_join[i], 0) != cudaSuccess)) r = GIC_ERR_CUDA;
  }
  set_cta_limit(0);
  if (r == GIC_ERR_CUDA) set_error("decode_step_all: stream fork / join failed: %s", cudaGetErrorString(cudaGetLastError()));
  return r;
}

// mapping network: image embeddings [B,E] -> prefix tokens fp32 [B,P_img,d] in w.prefix
static int mapper_forward(const gic_engine* e, const Workspace& w, const float* x, cudaStream_t st) {
  const int B = w.B, d = e->d;
  Act xin;
  if (!e->tc) xin.f32 = const_cast<float*>(x);
  else {
    GIC_TRY(launch_convert(x, w.x_in.out(), (size_t)B * e->E, st));
    xin = w.x_in;
  }
  if (e->cfg.mapper_kind == GIC_MAPPER_MLP) {
    // Linear -> Tanh -> Linear -> view [B,P,d]  (src/models.py:52-56,71-74)
    GIC_TRY(linear(e, e->map1, xin, B, EPI_TANH, w.map_hidden.out(), e->map1.N, st));
    ActOut o; o.f32 = w.prefix;
    GIC_TRY(linear(e, e->map2, w.map_hidden, B, EPI_NONE, o, e->map2.N, st));
    return GIC_OK;
  }
  // transformer mapper (src/models.py:141-174)
  const int Hl = e->cfg.hidden_length, P = e->P_img, S = Hl + P, M = B * S, heads = e->cfg.mapper_heads;
  ActOut lo; lo.f32 = w.map_lin;
  GIC_TRY(linear(e, e->tfm_linear, xin, B, EPI_NONE, lo, e->tfm_linear.N, st));
  GIC_TRY(launch_build_mapper_seq(w.map_lin, e->tfm_prefix_const, w.h, B, Hl, P, d, st));
  ActOut hres; hres.f32 = w.h;
  for (size_t l = 0; l < e->tfm_layers.size(); ++l) {
    const TfmLayer& T = e->tfm_layers[l];
    GIC_TRY(launch_layernorm(w.h, d, T.n1.w, T.n1.b, w.a.out(), M, d, st));
    GIC_TRY(linear(e, T.in_proj, w.a, M, EPI_NONE, qkv_out(e, w), 3 * d, st));
    if (e->cfg.dtype == GIC_DTYPE_BF16) GIC_TRY(launch_attn_encoder<bf16>(w.qkv_bf16, w.o.out(), B, S, heads, d / heads, st));
    else GIC_TRY(launch_attn_encoder<float>(w.qkv_f32, w.o.out(), B, S, heads, d / heads, st));
    GIC_TRY(linear(e, T.out_proj, w.o, M, EPI_RESIDUAL, hres, d, st));
    GIC_TRY(launch_layernorm(w.h, d, T.n2.w, T.n2.b, w.a.out(), M, d, st));
    GIC_TRY(linear(e, T.lin1, w.a, M, EPI_RELU, w.f.out(), 4 * d, st));
    GIC_TRY(linear(e, T.lin2, w.f, M, EPI_RESIDUAL, hres, d, st));
  }
  return launch_slice_tokens(w.h, S, Hl, P, w.prefix, B, d, st);
}

// run on the engine's private stream, ordered after everything already queued on the caller's stream ...
static int fork_stream(gic_engine* e, cudaStream_t user) {
  GIC_CHECK_CUDA(cudaEventRecord(e->ev_in, user));
  GIC_CHECK_CUDA(cudaStreamWaitEvent(e->stream, e->ev_in, 0));
  return GIC_OK;
}
// ... and make the caller's stream wait for it
static int join_stream(gic_engine* e, cudaStream_t user) {
  GIC_CHECK_CUDA(cudaEventRecord(e->ev_out, e->stream));
  GIC_CHECK_CUDA(cudaStreamWaitEvent(user, e->ev_out, 0));
  return GIC_OK;
}

static int check_ready(const gic_engine* e) {
  GIC_REQUIRE(e != nullptr, "null engine");
  GIC_REQUIRE(e->gpt_loaded, "GPT-2 weights not loaded (gic_engine_load_gpt2)");
  GIC_REQUIRE(e->mapper_loaded, "mapping-network weights not loaded");
  GIC_REQUIRE(e->P_task == 0 || e->task_prefix != nullptr, "task prefix declared in the config but not loaded");
  return GIC_OK;
}

// per-context resources of an engine handle: the private stream generate runs on, its fork / join events, the sub-batch streams and
// the pinned early-exit flag
static int create_context_resources(gic_engine* e) {
  bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaHostAlloc((void**)&e->h_done, 4 * sizeof(int), cudaHostAllocDefault) == cudaSuccess &&
       cudaEventCreateWithFlags(&e->ev_done[0], cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&e->ev_done[1], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 1; i < gic_engine::MAX_SUB && ok; ++i)
    ok = cudaStreamCreateWithFlags(&e->sub_stream[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("could not create the engine stream / events: %s", cudaGetErrorString(cudaGetLastError()));
    return GIC_ERR_CUDA;
  }
  return GIC_OK;
}

}  // namespace gic

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" {

int gic_engine_destroy(gic_engine* e);

const char* gic_last_error(void) { return gic::get_error(); }
int gic_abi_version(void) { return GIC_ABI_VERSION; }

int gic_device_check(void) {
  int dev = 0;
  GIC_CHECK_CUDA(cudaGetDevice(&dev));
  // (cudaGetDeviceProperties takes milliseconds and is called from per-request entry points: ask once per device)
  static unsigned char checked[64] = {0};
  if (dev >= 0 && dev < 64 && checked[dev]) return GIC_OK;
  cudaDeviceProp prop;
  GIC_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major == 10 && dev >= 0 && dev < 64) checked[dev] = 1;
  if (prop.major != 10) {
    gic::set_error("device %d (%s) is sm_%d%d; this library contains sm_100a code only", dev, prop.name, prop.major, prop.minor);
    return GIC_ERR_UNSUPPORTED;
  }
  return GIC_OK;
}

int gic_engine_create(const gic_config* cfg, gic_engine** out) {
  GIC_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  GIC_REQUIRE(cfg->abi_version == GIC_ABI_VERSION, "ABI version mismatch: caller %d, library %d", cfg->abi_version, GIC_ABI_VERSION);
  GIC_REQUIRE(cfg->dtype >= GIC_DTYPE_F32 && cfg->dtype <= GIC_DTYPE_BF16X2, "unknown dtype %d", cfg->dtype);
  GIC_REQUIRE(cfg->n_embd > 0 && cfg->n_head > 0 && cfg->n_embd == cfg->n_head * 64,
              "GPT-2 head_dim must be 64 (n_embd %d, n_head %d)", cfg->n_embd, cfg->n_head);
  GIC_REQUIRE(cfg->n_embd % 64 == 0 && cfg->n_embd <= 1280, "n_embd %d unsupported (multiple of 64, <= 1280)", cfg->n_embd);
  GIC_REQUIRE(cfg->n_layer > 0 && cfg->vocab_size > 0 && cfg->n_positions > 0, "bad GPT-2 dimensions");
  GIC_REQUIRE(cfg->embed_dim > 0 && cfg->embed_dim % 8 == 0, "embed_dim %d must be a positive multiple of 8", cfg->embed_dim);
  GIC_REQUIRE(cfg->prefix_length > 0 && cfg->task_prefix_length >= 0, "bad prefix lengths");
  GIC_REQUIRE(cfg->mapper_kind == GIC_MAPPER_MLP || cfg->mapper_kind == GIC_MAPPER_TRANSFORMER, "unknown mapper kind %d", cfg->mapper_kind);
  if (cfg->mapper_kind == GIC_MAPPER_TRANSFORMER) {
    GIC_REQUIRE(cfg->hidden_length > 0 && cfg->mapper_layers > 0 && cfg->mapper_heads > 0, "bad transformer-mapper dimensions");
    GIC_REQUIRE(cfg->n_embd % cfg->mapper_heads == 0, "transformer mapper: n_embd %d not divisible by %d heads", cfg->n_embd,
                cfg->mapper_heads);
  } else {
    GIC_REQUIRE((cfg->prefix_length * cfg->n_embd) % 16 == 0, "MLP mapper hidden size must be a multiple of 8");
  }
  GIC_TRY(gic_device_check());
  gic_engine* e = new gic_engine();
  e->cfg = *cfg;
  e->d = cfg->n_embd; e->L = cfg->n_layer; e->H = cfg->n_head; e->V = cfg->vocab_size;
  e->P_img = cfg->prefix_length; e->P_task = cfg->task_prefix_length; e->E = cfg->embed_dim;
  e->tc = cfg->dtype != GIC_DTYPE_F32;
  e->split = cfg->dtype == GIC_DTYPE_BF16X2;
  {
    const char* nf = getenv("GIC_NO_LNFUSE");
    e->fuse_ln = e->tc && !(nf && nf[0] == '1');  // GIC_NO_LNFUSE=1: LayerNorm as its own kernel; BF16X2 then also keeps fp32 q | k | v and cache
    e->beam_indirect = e->fuse_ln && gic::attn_decode_indirect_available();
    const char* bl = getenv("GIC_BEAM_LOGITS");
    e->beam_fused_head = e->tc && !(bl && bl[0] == '1');
    const char* sk = getenv("GIC_SPLITK");
    // off unless GIC_SPLITK=1 in both tensor-core modes: measured again in round 2 for bf16x2, whose fc2 main loop is 2.5x longer
    // (profiles/r2f_splitk.txt): 29.2 us with the 3-way K split against 22.6 us unsplit -- the parked-partials reduction costs more than
    // the shorter main loop saves.  The cooperative reduction (every slice finishes a share of the tile; CTA pairs allowed) brings the
    // split to 23.5 us -- still not below the unsplit pair kernel: 144 CTAs x 64 KB of partials make ~9 MB of L2 writes + reads per launch
    // (profiles/r2n_splitk_coop_microbench.txt)
    e->use_splitk = sk && sk[0] == '1';
    const char* hf = getenv("GIC_LNF_FUSE");
    e->fuse_lnf = e->fuse_ln && !e->split && hf && hf[0] == '1';
    const char* xh = getenv("GIC_X2_HEAD_FULL");
    e->rescore_head = e->split && e->fuse_ln && !(xh && xh[0] == '1');
  }
  const char* ng = getenv("GIC_NO_GRAPH");
  e->use_graph = !(ng && ng[0] == '1');
  if (e->tc) {
    int r = gic::tma_init();
    if (r == GIC_OK) r = gic::gemm_bf16_configure();
    if (r == GIC_OK) r = gic::attn_decode_configure();
    if (r != GIC_OK) { delete e; return r; }
  }
  const char* sb = getenv("GIC_SUBBATCH");
  if (sb && sb[0] >= '1' && sb[0] <= '8') e->sub_batches = sb[0] - '0';
  if (create_context_resources(e) != GIC_OK) { delete e; return GIC_ERR_CUDA; }
  *out = e;
  return GIC_OK;
}

// A second CONTEXT on the same packed weights: its own stream, events, CUDA-graph cache and early-exit state, so that another batch can
// be in flight (inflight.py) without a second copy of the weights (every in-flight slot used to be a full engine).
// The clone borrows every weight pointer and tensor map of `src`, owns none of them (gic_engine_weight_bytes == 0), and must be
// destroyed before `src`.
int gic_engine_clone(const gic_engine* src, gic_engine** out) {
  GIC_REQUIRE(src != nullptr && out != nullptr, "null argument");
  GIC_REQUIRE(src->gpt_loaded && src->mapper_loaded, "clone an engine after its weights are loaded");
  gic_engine* e = new gic_engine(*src);
  e->allocs.clear();
  e->weight_bytes = 0;
  e->graphs.clear(); e->graph_ws = nullptr; e->graph_B = 0; e->graph_max_new = 0; e->graph_steps = 1;
  e->prof.clear(); e->profiling = false;
  e->sample = gic_engine::SampleCfg();
  e->stream = nullptr; e->ev_in = e->ev_out = e->ev_fork = nullptr; e->h_done = nullptr;
  e->ev_done[0] = e->ev_done[1] = nullptr;
  for (int i = 0; i < gic_engine::MAX_SUB; ++i) { e->sub_stream[i] = nullptr; e->ev_join[i] = nullptr; }
  if (create_context_resources(e) != GIC_OK) { gic_engine_destroy(e); return GIC_ERR_CUDA; }
  *out = e;
  return GIC_OK;
}

int gic_engine_destroy(gic_engine* e) {
  if (!e) return GIC_OK;
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  for (int i = 0; i < 2; ++i) if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
  if (e->h_done) cudaFreeHost(e->h_done);
  for (int i = 1; i < gic_engine::MAX_SUB; ++i) {
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    if (e->sub_stream[i]) cudaStreamDestroy(e->sub_stream[i]);
  }
  if (e->stream) cudaStreamDestroy(e->stream);
  for (void* p : e->allocs) cudaFree(p);
  delete e;
  return GIC_OK;
}

size_t gic_engine_weight_bytes(const gic_engine* e) { return e ? e->weight_bytes : 0; }

int gic_engine_load_gpt2(gic_engine* e, const gic_gpt2_weights* w, void* stream) {
  GIC_REQUIRE(e && w && w->layers, "null argument");
  GIC_REQUIRE(!e->gpt_loaded, "GPT-2 weights already loaded; create a new engine to reload");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = e->d;
  // bf16 engine: ln_f is folded into the LM head (the tied table then differs from the embedding table, kept separately)
  const bool fold_head = e->fuse_ln && e->fuse_lnf;
  GIC_TRY(pack_linear(e, &e->lm_head, w->wte, nullptr, e->V, d, false, st, fold_head ? w->lnf_w : nullptr, fold_head ? w->lnf_b : nullptr));
  if (fold_head) {
    GIC_TRY(dev_alloc(e, (void**)&e->wte_gather, (size_t)e->V * d * sizeof(bf16)));
    ActOut go; go.hi = e->wte_gather;
    GIC_TRY(launch_convert(w->wte, go, (size_t)e->V * d, st));
  } else {
    e->wte_gather = e->lm_head.w_hi;
  }
  if (e->cfg.dtype == GIC_DTYPE_F32) e->wte_f32 = e->lm_head.w_f32;
  else if (e->cfg.dtype == GIC_DTYPE_BF16X2) GIC_TRY(copy_vec(e, &e->wte_f32, w->wte, (size_t)e->V * d, st));
  if (e->rescore_head) {
    GIC_TRY(dev_alloc(e, (void**)&e->wte_norm, (size_t)e->V * sizeof(float)));
    GIC_TRY(launch_row_norms(e->wte_f32, e->V, d, e->wte_norm, st));
    for (int i = 0; i < kNumHeadTiles; ++i) {
      GIC_TRY(dev_alloc(e, (void**)&e->slot_norm_max[i], (size_t)2 * ceil_div(e->V, kHeadTiles[i]) * sizeof(float)));
      GIC_TRY(launch_slot_norm_max(e->wte_norm, e->V, kHeadTiles[i], e->slot_norm_max[i], st));
    }
  }
  GIC_TRY(copy_vec(e, &e->wpe, w->wpe, (size_t)e->cfg.n_positions * d, st));
  GIC_TRY(copy_norm(e, &e->lnf, w->lnf_w, w->lnf_b, d, st));
  e->layers.resize(e->L);
  for (int l = 0; l < e->L; ++l) {
    const gic_gpt2_layer_weights& s = w->layers[l];
    GptLayer& D = e->layers[l];
    GIC_TRY(copy_norm(e, &D.ln1, s.ln1_w, s.ln1_b, d, st));
    GIC_TRY(copy_norm(e, &D.ln2, s.ln2_w, s.ln2_b, d, st));
    // Conv1D [d,3d] -> [3d,d]; bf16 engine: ln_1 folded into c_attn, ln_2 into c_fc
    GIC_TRY(pack_linear(e, &D.attn, s.attn_w, s.attn_b, 3 * d, d, true, st, e->fuse_ln ? s.ln1_w : nullptr, e->fuse_ln ? s.ln1_b : nullptr));
    GIC_TRY(pack_linear(e, &D.proj, s.proj_w, s.proj_b, d, d, true, st));
    GIC_TRY(pack_linear(e, &D.fc, s.fc_w, s.fc_b, 4 * d, d, true, st, e->fuse_ln ? s.ln2_w : nullptr, e->fuse_ln ? s.ln2_b : nullptr));
    GIC_TRY(pack_linear(e, &D.fc2, s.fc2_w, s.fc2_b, d, 4 * d, true, st));
  }
  e->gpt_loaded = true;
  return GIC_OK;
}

int gic_engine_load_mlp_mapper(gic_engine* e, const gic_mlp_mapper_weights* w, void* stream) {
  GIC_REQUIRE(e && w, "null argument");
  GIC_REQUIRE(e->cfg.mapper_kind == GIC_MAPPER_MLP, "engine was not configured for the MLP mapper");
  GIC_REQUIRE(!e->mapper_loaded, "mapper weights already loaded");
  cudaStream_t st = (cudaStream_t)stream;
  const int out = e->P_img * e->d, hid = out / 2;
  GIC_TRY(pack_linear(e, &e->map1, w->w1, w->b1, hid, e->E, false, st));
  GIC_TRY(pack_linear(e, &e->map2, w->w2, w->b2, out, hid, false, st));
  e->mapper_loaded = true;
  return GIC_OK;
}

int gic_engine_load_tfm_mapper(gic_engine* e, const gic_tfm_mapper_weights* w, void* stream) {
  GIC_REQUIRE(e && w && w->layers, "null argument");
  GIC_REQUIRE(e->cfg.mapper_kind == GIC_MAPPER_TRANSFORMER, "engine was not configured for the transformer mapper");
  GIC_REQUIRE(!e->mapper_loaded, "mapper weights already loaded");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = e->d;
  GIC_TRY(pack_linear(e, &e->tfm_linear, w->linear_w, w->linear_b, e->cfg.hidden_length * d, e->E, false, st));
  GIC_TRY(copy_vec(e, &e->tfm_prefix_const, w->prefix_const, (size_t)e->P_img * d, st));
  e->tfm_layers.resize(e->cfg.mapper_layers);
  for (int l = 0; l < e->cfg.mapper_layers; ++l) {
    const gic_tfm_layer_weights& s = w->layers[l];
    TfmLayer& D = e->tfm_layers[l];
    GIC_TRY(copy_norm(e, &D.n1, s.norm1_w, s.norm1_b, d, st));
    GIC_TRY(copy_norm(e, &D.n2, s.norm2_w, s.norm2_b, d, st));
    GIC_TRY(pack_linear(e, &D.in_proj, s.in_proj_w, s.in_proj_b, 3 * d, d, false, st));
    GIC_TRY(pack_linear(e, &D.out_proj, s.out_proj_w, s.out_proj_b, d, d, false, st));
    GIC_TRY(pack_linear(e, &D.lin1, s.lin1_w, s.lin1_b, 4 * d, d, false, st));
    GIC_TRY(pack_linear(e, &D.lin2, s.lin2_w, s.lin2_b, d, 4 * d, false, st));
  }
  e->mapper_loaded = true;
  return GIC_OK;
}

int gic_engine_load_task_prefix(gic_engine* e, const float* task, void* stream) {
  GIC_REQUIRE(e && task, "null argument");
  GIC_REQUIRE(e->P_task > 0, "engine was configured without a task prefix");
  GIC_REQUIRE(e->task_prefix == nullptr, "task prefix already loaded");
  return copy_vec(e, &e->task_prefix, task, (size_t)e->P_task * e->d, (cudaStream_t)stream);
}

size_t gic_workspace_bytes(const gic_engine* e, int batch, int max_new_tokens, int num_beams) {
  if (!e || batch <= 0 || max_new_tokens < 0) return 0;
  Workspace w;
  carve(e, nullptr, batch, max_new_tokens, num_beams, &w);
  return w.bytes;
}

static int prepare_ws(const gic_engine* e, void* workspace, size_t workspace_bytes, int batch, int max_new, int beams, Workspace* w) {
  GIC_REQUIRE(batch > 0, "batch must be positive (got %d)", batch);
  GIC_REQUIRE(max_new >= 0, "max_new_tokens must be >= 0");
  GIC_REQUIRE(workspace != nullptr, "null workspace");
  // the last generated token is never fed back (src/models.py:466-469 runs after the final step but its result is unused), so the
  // largest position embedded is P + max_new - 2; the reference accepts max_length up to n_positions - P + 1
  GIC_REQUIRE(e->P_img + e->P_task + max_new - 1 <= e->cfg.n_positions, "prefix + max_new_tokens - 1 (%d) exceeds n_positions (%d)",
              e->P_img + e->P_task + max_new - 1, e->cfg.n_positions);
  void* aligned = (void*)align_up((size_t)workspace, 1024);
  const size_t lost = (size_t)((unsigned char*)aligned - (unsigned char*)workspace);
  carve(e, aligned, batch, max_new, beams, w);
  if (w->bytes - 1024 + lost > workspace_bytes) {  // gic_workspace_bytes includes 1024 bytes of alignment slack
    set_error("workspace too small: need %zu bytes, got %zu", w->bytes, workspace_bytes);
    return GIC_ERR_WORKSPACE;
  }
  return GIC_OK;
}

int gic_mapper_forward(gic_engine* e, const float* x, int batch, float* prefix_out, void* workspace, size_t workspace_bytes, void* stream) {
  GIC_TRY(check_ready(e));
  GIC_REQUIRE(x && prefix_out, "null argument");
  Workspace w;
  GIC_TRY(prepare_ws(e, workspace, workspace_bytes, batch, 0, 1, &w));
  cudaStream_t user = (cudaStream_t)stream, st = e->stream;
  GIC_TRY(fork_stream(e, user));
  GIC_TRY(mapper_forward(e, w, x, st));
  // [B, P_img + P_task, d]: image prefix then the task rows (src/models.py:364-375)
  GIC_TRY(launch_embed_prefix(w.prefix, e->P_img, e->task_prefix, e->P_task, nullptr, nullptr, prefix_out, batch, e->d, st));
  return join_stream(e, user);
}

// the body of gic_generate_greedy between fork_stream and join_stream (the caller joins on every path, errors included)
static int generate_greedy_on_stream(gic_engine* e, const float* x, int max_new, int64_t* ids_out, int32_t* gen_len_out, float* logits_out,
                                     void* workspace, Workspace& w, cudaStream_t st) {
  const int d = e->d, B = w.B, P = w.P;
  GIC_TRY(launch_init_decode_state(w.finished, w.first_eos, B, max_new, w.d_step, w.d_pos, w.done_counter, P, w.fin_counter, w.all_done, w.ids,
                                   e->cfg.eos_token_id, st));
  if (w.splitk_counters) GIC_CHECK_CUDA(cudaMemsetAsync(w.splitk_counters, 0, 4096 * sizeof(int), st));
  if (w.rs_counters) GIC_CHECK_CUDA(cudaMemsetAsync(w.rs_counters, 0, 2 * sizeof(int), st));
  if (e->sample.on) GIC_TRY(launch_set_sample_params(w.sample_params, e->sample.temperature, e->sample.top_p, e->sample.seed, st));
  if (e->profiling) GIC_TRY(launch_spin(150000000LL, st));  // ~75 ms: lets the host queue ahead so events time the device only
  { ProfScope ps(e, "mapper", st); GIC_TRY(mapper_forward(e, w, x, st)); }
  GIC_TRY(launch_embed_prefix(w.prefix, e->P_img, e->task_prefix, e->P_task, e->wpe, w.h, nullptr, B, d, st));
  // ---- prefill over the P prefix tokens of every row ----
  if (e->fuse_ln) GIC_TRY(launch_row_stats(w.h, d, w.a.hi, w.ln_stats, B * P, d, st, w.a.lo));
  GIC_TRY(gpt_layers(e, w, w.h, B * P, true, st));
  // only the last position feeds the LM head (the reference computes all positions and keeps [:, -1, :], src/models.py:398)
  GIC_TRY(lm_head_and_token(e, w, w.h, (long)(P - 1) * d, (long)P * d, B, B * P, logits_out, st));

  // ---- decode: max_new-1 identical steps; positions / step index live on the device.  They run in chunks of DECODE_CHUNK steps
  // (one CUDA graph per chunk); while chunk c runs the host reads chunk c-1's "every row has emitted EOS" flag and live-row count:
  // it stops launching once the flag is set -- the reference's `if is_finished.all(): break` (src/models.py:390-391) without a
  // per-step sync -- and SHRINKS the batch to the live rows when enough have finished (compact_rows_kernel: the reference finishes
  // rows one by one, :453-460, and trained models stop at 10-20 of 50 tokens).  Random-init weights never emit EOS, so the headline
  // benchmark runs every chunk at full size. ----
  const int steps = max_new - 1;
  constexpr int DECODE_CHUNK = 4;
  // (sampling taps every step's logits into the SAME scratch rows and reads its parameters from device memory: capturable; the
  // per-step logits tap of the tests is not)
  const bool graph_ok = e->use_graph && !e->profiling && (logits_out == nullptr || e->sample.on) && steps >= 2;
  static const bool per_step = [] { const char* v = getenv("GIC_GRAPH_PER_STEP"); return v && v[0] == '1'; }();
  const char* ne = getenv("GIC_NO_EARLY_EXIT");  // (read per call: a test flips it inside one process)
  const bool no_early = ne && ne[0] == '1';
  const char* nc = getenv("GIC_NO_COMPACT");
  // (a traced run keeps one record per launch: one graph holding every step, no replays)
  const bool early = !no_early && !gic::trace_on() && e->sub_batches < 2 && !(logits_out && !e->sample.on) && e->h_done != nullptr;
  const bool compact = early && !(nc && nc[0] == '1') && !logits_out && w.row_map_buf != nullptr && B >= 256;
  const int chunk = (graph_ok && per_step) ? 1 : (early ? DECODE_CHUNK : (graph_ok ? steps : DECODE_CHUNK));
  Workspace wc = w;  // the batch as the decode steps see it: all B rows, or the first M slots of the compacted state
  auto eager_step = [&](int s) {
    return decode_step_all(e, wc, logits_out ? (e->sample.on ? logits_out : logits_out + (size_t)s * B * e->V) : nullptr, st);
  };
  int s_next = 1;  // next decode step (1-based, as the logits tap is indexed)
  // the steps that do not fill a whole chunk go first, as ordinary launches (the host is far ahead of the device after prefill)
  const int lead = graph_ok ? steps % chunk : 0;
  for (; s_next <= lead; ++s_next) GIC_TRY(eager_step(s_next));
  if (!(e->graph_ws == workspace && e->graph_B == B && e->graph_max_new == max_new && e->graph_steps == chunk && e->graph_trace_gen == gic::trace_generation() &&
        e->graph_logits == logits_out)) {
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
    e->graphs.clear();
    e->graph_ws = workspace; e->graph_B = B; e->graph_max_new = max_new; e->graph_steps = chunk; e->graph_trace_gen = gic::trace_generation();
    e->graph_logits = logits_out;
  }
  // the chunk graph over the current batch view (captured on first use, one per batch size)
  auto chunk_graph = [&](const gic_engine::GraphEntry** out) -> int {
    for (auto& g : e->graphs) if (g.M == wc.rows) { *out = &g; return GIC_OK; }
    cudaGraph_t graph = nullptr;
    const unsigned long long before = tl_launches;
    GIC_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int r = GIC_OK;
    for (int s = 0; s < chunk && r == GIC_OK; ++s) r = decode_step_all(e, wc, e->sample.on ? logits_out : nullptr, st);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    gic_engine::GraphEntry ge;
    ge.M = wc.rows; ge.exec = nullptr;
    ge.nodes = (int)(tl_launches - before);  // captured, not executed
    g_launches.fetch_sub(tl_launches - before, std::memory_order_relaxed);
    tl_launches = before;
    if (r != GIC_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    GIC_CHECK_CUDA(ce);
    ce = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    GIC_CHECK_CUDA(ce);
    e->graphs.push_back(ge);
    *out = &e->graphs.back();
    return GIC_OK;
  };
  int m_pending = 0;  // > 0: shrink the batch to this many slots before the next chunk
  for (int c = 0; s_next <= steps; ++c) {
    if (m_pending > 0) {
      CompactArgs ca;
      ca.finished = w.finished; ca.row_map_old = wc.row_map; ca.m_old = wc.rows; ca.m_new = m_pending; ca.d = d;
      ca.row_map_new = w.row_map_buf; ca.src_slot = w.src_slot;
      ca.h = w.h_dec; ca.h_tmp = w.h_tmp; ca.a_hi = e->fuse_ln ? w.a.hi : nullptr; ca.a_hi_tmp = w.a_tmp_hi;
      ca.a_lo = e->fuse_ln ? w.a.lo : nullptr; ca.a_lo_tmp = w.a_tmp_lo; ca.stats = e->fuse_ln ? w.ln_stats : nullptr; ca.stats_tmp = w.stats_tmp;
      { ProfScope ps(e, "compact_rows", st); GIC_TRY(launch_compact_rows(ca, st)); }
      wc.rows = m_pending; wc.row_map = w.row_map_buf;
      m_pending = 0;
      g_compactions.fetch_add(1, std::memory_order_relaxed);
    }
    const int n = steps - s_next + 1 < chunk ? steps - s_next + 1 : chunk;
    if (graph_ok) {
      const gic_engine::GraphEntry* ge = nullptr;
      GIC_TRY(chunk_graph(&ge));
      GIC_CHECK_CUDA(cudaGraphLaunch(ge->exec, st));
      g_launches.fetch_add((unsigned long long)ge->nodes, std::memory_order_relaxed);
    } else {
      for (int i = 0; i < n; ++i) GIC_TRY(eager_step(s_next + i));
    }
    s_next += n;
    if (early && s_next <= steps) {
      int* hd = e->h_done + 2 * (c & 1);  // {every row finished, live rows} after chunk c
      GIC_CHECK_CUDA(cudaMemcpyAsync(hd, w.all_done, sizeof(int), cudaMemcpyDeviceToHost, st));
      GIC_CHECK_CUDA(cudaMemcpyAsync(hd + 1, w.live_rows, sizeof(int), cudaMemcpyDeviceToHost, st));
      GIC_CHECK_CUDA(cudaEventRecord(e->ev_done[c & 1], st));
      if (c >= 1) {
        GIC_CHECK_CUDA(cudaEventSynchronize(e->ev_done[(c - 1) & 1]));
        const int* hp = e->h_done + 2 * ((c - 1) & 1);
        if (hp[0]) break;  // (chunk c is already queued: at most one chunk runs past the end)
        if (compact) {
          // rows only ever finish, so the count after chunk c-1 bounds the live rows at any later time.  Sizes are multiples of 256 rows
          // (whole CTA pairs of 128-row tiles), and a shrink must drop at least a quarter of the current slots to pay for its three launches
          const int live = hp[1] < 1 ? 1 : hp[1];
          const int m_new = live <= 128 ? 128 : ((live + 255) / 256) * 256;
          if (m_new <= wc.rows - wc.rows / 4 && m_new < wc.rows) m_pending = m_new;
        }
      }
    }
  }
  GIC_CHECK_CUDA(cudaMemcpyAsync(ids_out, w.ids, (size_t)B * max_new * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (gen_len_out) GIC_TRY(launch_gen_len(w.first_eos, B, max_new, gen_len_out, st));
  return GIC_OK;
}

int gic_generate_greedy(gic_engine* e, const float* x, int batch, int max_new, int64_t* ids_out, int32_t* gen_len_out, float* logits_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
  GIC_TRY(check_ready(e));
  GIC_REQUIRE(x && ids_out, "null argument");
  GIC_REQUIRE(max_new >= 1, "max_new_tokens must be >= 1 (the caller handles 0, src/models.py:471-473)");
  Workspace w;
  GIC_TRY(prepare_ws(e, workspace, workspace_bytes, batch, max_new, 1, &w));
  cudaStream_t user = (cudaStream_t)stream;
  GIC_TRY(fork_stream(e, user));
  const int r = generate_greedy_on_stream(e, x, max_new, ids_out, gen_len_out, logits_out, workspace, w, e->stream);
  if (r != GIC_OK) {
    // keep the caller's stream ordered behind whatever was queued before the failure (the error text of `r` stands)
    if (cudaEventRecord(e->ev_out, e->stream) == cudaSuccess) cudaStreamWaitEvent(user, e->ev_out, 0);
    cudaGetLastError();
    return r;
  }
  return join_stream(e, user);
}

// the `temperature > 0` branch of ImageCaptioningModel.generate (src/models.py:400-449): the greedy driver with every step's
// logits tapped into logits_scratch and the token drawn by sample_top_p_kernel; the decode steps replay the same chunked CUDA graphs as
// the greedy path (temperature / top_p / seed are read from the workspace, so one capture serves every call)
int gic_generate_sample(gic_engine* e, const float* x, int batch, int max_new, float temperature, float top_p, unsigned long long seed,
                        int64_t* ids_out, int32_t* gen_len_out, float* logits_scratch, void* workspace, size_t workspace_bytes, void* stream) {
  GIC_TRY(check_ready(e));
  GIC_REQUIRE(logits_scratch != nullptr, "null argument");
  GIC_REQUIRE(temperature > 0.f && top_p > 0.f, "sampling needs temperature > 0 and top_p > 0 (temperature %g, top_p %g)", temperature, top_p);
  e->sample.on = true; e->sample.temperature = temperature; e->sample.top_p = top_p; e->sample.seed = seed; e->sample.logits = logits_scratch;
  e->sample.ld = (e->V + 31) / 32 * 32;
  const int r = gic_generate_greedy(e, x, batch, max_new, ids_out, gen_len_out, logits_scratch, workspace, workspace_bytes, stream);
  e->sample.on = false;
  return r;
}

// ln_f -> LM head with the full fp32 logit rows materialised (beam search needs log-softmax + top-2K over beams x V)
static int lm_head_logits(const gic_engine* e, const Workspace& w, const float* h_buf, long first_off, long row_stride, int rows, int body_rows,
                          cudaStream_t st) {
  ActOut o; o.f32 = w.logits;
  if (e->fuse_ln && e->fuse_lnf) {
    Act a = w.a; a.hi += first_off;
    LnIo in; in.stats_in = w.ln_stats; in.parts_in = e->L > 0 ? residual_stats_parts(e, body_rows) : 1; in.stats_ld = w.m_max;
    in.row_mul = (int)(row_stride / e->d); in.row_off = (int)(first_off / e->d); in.a_row_stride = row_stride;
    ProfScope ps(e, "lm_head", st);
    return linear(e, e->lm_head, a, rows, EPI_NONE, o, e->V, st, nullptr, nullptr, nullptr, 0, &in);
  }
  { ProfScope ps(e, "layernorm", st); GIC_TRY(launch_layernorm(h_buf + first_off, row_stride, e->lnf.w, e->lnf.b, w.a.out(), rows, e->d, st)); }
  ProfScope ps(e, "lm_head", st);
  return linear(e, e->lm_head, w.a, rows, EPI_NONE, o, e->V, st);
}

// ln_f -> LM head whose epilogue keeps, per row and stream, the 16 best logits and the log-sum-exp partials (EPI_BEAM): what HF's
// log_softmax + topk(2 * beams) (HF:generation/utils.py:3252-3256,2981-2987) need, without the [rows, V] fp32 matrix (1 GB per step at config 3)
static int lm_head_beam(const gic_engine* e, const Workspace& w, const float* h_buf, long first_off, long row_stride, int rows, int* streams_out, cudaStream_t st) {
  { ProfScope ps(e, "layernorm", st); GIC_TRY(launch_layernorm(h_buf + first_off, row_stride, e->lnf.w, e->lnf.b, w.a.out(), rows, e->d, st)); }
  ProfScope ps(e, "lm_head", st);
  const Linear& lin = e->lm_head;
  GemmBf16Args g;
  int bn = 0, pair = 0;
  beam_head_shape(e, rows, &bn, &pair);
  const int bi = box_rows_index(pair ? bn / 2 : bn);
  GIC_REQUIRE(bi >= 0, "no W tensor map for box height %d", pair ? bn / 2 : bn);
  GIC_TRY(make_tma_2d_bf16(&g.a_hi, w.a.hi, rows, lin.K, lin.K, 128));
  g.w_hi = lin.tm_hi[bi];
  if (e->split) {
    GIC_TRY(make_tma_2d_bf16(&g.a_lo, w.a.lo, rows, lin.K, lin.K, 128));
    g.w_lo = lin.tm_lo[bi];
  }
  const int streams = gemm_topk_streams(rows, lin.N, bn, pair);
  GIC_REQUIRE((size_t)rows * streams <= w.beam.tk_slots, "beam LM head: %d rows x %d streams exceed the workspace", rows, streams);
  g.M = rows; g.N = lin.N; g.K = lin.K; g.block_n = bn; g.pair = pair; g.split = e->split ? 1 : 0; g.epilogue = EPI_NONE; g.w_static = 1;
  g.topk_v = w.beam.tk_v; g.topk_i = w.beam.tk_i; g.beam_m = w.beam.tk_m; g.beam_s = w.beam.tk_s; g.topk_streams = streams;
  *streams_out = streams;
  return launch_gemm_bf16(g, st);
}

int gic_generate_beam(gic_engine* e, const float* x, int batch, int max_new, int num_beams, float length_penalty, int64_t* ids_out,
                      float* scores_out, int32_t* gen_len_out, void* workspace, size_t workspace_bytes, void* stream) {
  GIC_TRY(check_ready(e));
  GIC_REQUIRE(x && ids_out, "null argument");
  GIC_REQUIRE(max_new >= 1, "max_new_tokens must be >= 1");
  GIC_REQUIRE(num_beams >= 2 && num_beams <= 8, "num_beams must be in [2, 8] (got %d); use gic_generate_greedy for 1", num_beams);
  Workspace w;
  GIC_TRY(prepare_ws(e, workspace, workspace_bytes, batch, max_new, num_beams, &w));
  cudaStream_t user = (cudaStream_t)stream, st = e->stream;
  const int d = e->d, B = batch, P = w.P, nb = num_beams, rows = w.rows, K = 2 * nb;
  GIC_TRY(fork_stream(e, user));
  GIC_TRY(launch_beam_init(w.beam, st));
  if (w.splitk_counters) GIC_CHECK_CUDA(cudaMemsetAsync(w.splitk_counters, 0, 4096 * sizeof(int), st));
  { ProfScope ps(e, "mapper", st); GIC_TRY(mapper_forward(e, w, x, st)); }
  GIC_TRY(launch_embed_prefix(w.prefix, e->P_img, e->task_prefix, e->P_task, e->wpe, w.h, nullptr, B, d, st));
  // prefill once per image; its K/V land in cache row b*beams and the first reorder fans them out to every beam
  if (e->fuse_ln) GIC_TRY(launch_row_stats(w.h, d, w.a.hi, w.ln_stats, B * P, d, st, w.a.lo));
  GIC_TRY(gpt_layers(e, w, w.h, B * P, true, st));
  int streams = 0;
  if (e->beam_fused_head) GIC_TRY(lm_head_beam(e, w, w.h, (long)(P - 1) * d, (long)P * d, B, &streams, st));
  else GIC_TRY(lm_head_logits(e, w, w.h, (long)(P - 1) * d, (long)P * d, B, B * P, st));
  for (int t = 0; t < max_new; ++t) {
    const int live = t == 0 ? 1 : nb;  // only beam 0 is live at the first step (running scores 0, -1e9, ...)
    { ProfScope ps(e, "beam_topk", st);
      if (e->beam_fused_head)
        GIC_TRY(launch_beam_topk_streams(w.beam.tk_v, w.beam.tk_i, w.beam.tk_m, w.beam.tk_s, streams, B, live, live, w.beam.run_score, nb, e->V, K, w.beam.lse,
                                         w.beam.row_val, w.beam.row_idx, w.beam.cand_score, w.beam.cand_idx, st));
      else
        GIC_TRY(launch_beam_topk(w.logits, B, live, live, w.beam.run_score, nb, e->V, K, w.beam.lse, w.beam.row_val, w.beam.row_idx, w.beam.cand_score,
                                 w.beam.cand_idx, st)); }
    const float denom = (float)pow((double)(t + 1), (double)length_penalty);
    { ProfScope ps(e, "beam_update", st); GIC_TRY(launch_beam_update(w.beam, t, denom, st)); }
    if (t + 1 == max_new) break;
    // reorder_cache: dst[row] = src[beam_idx[row]] over the P + t cached positions, then swap the two caches
    if (e->beam_indirect) {
      // no gather: record where each surviving hypothesis' generated positions live (its parents' cache rows)
      ProfScope ps(e, "kv_reorder", st);
      GIC_TRY(launch_beam_ancestry(w.beam.anc[t & 1], w.beam.anc[(t + 1) & 1], w.beam.beam_idx, rows, max_new, t, st));
      w.anc = w.beam.anc[(t + 1) & 1]; w.anc_ld = max_new;
    } else {
      ProfScope ps(e, "kv_reorder", st);
      if (kv16(e))
        GIC_TRY(launch_kv_reorder<bf16>((const bf16*)w.kv, (bf16*)w.kv2, w.beam.beam_idx, e->L, rows, e->H, P + t, w.t_max, st));
      else
        GIC_TRY(launch_kv_reorder<float>((const float*)w.kv, (float*)w.kv2, w.beam.beam_idx, e->L, rows, e->H, P + t, w.t_max, st));
      void* tmp = w.kv; w.kv = w.kv2; w.kv2 = tmp;
    }
    GIC_TRY(launch_beam_embed(w.beam.next_tok, e->wte_f32, e->wte_f32 ? nullptr : e->wte_gather, e->wpe, P + t, d, w.h_dec, rows, st));
    GIC_TRY(launch_set_int(w.d_pos, P + t, st));
    if (e->fuse_ln) GIC_TRY(launch_row_stats(w.h_dec, d, w.a.hi, w.ln_stats, rows, d, st, w.a.lo));
    GIC_TRY(gpt_layers(e, w, w.h_dec, rows, false, st));
    if (e->beam_fused_head) GIC_TRY(lm_head_beam(e, w, w.h_dec, 0, d, rows, &streams, st));
    else GIC_TRY(lm_head_logits(e, w, w.h_dec, 0, d, rows, rows, st));
  }
  GIC_TRY(launch_beam_finalize(w.beam, max_new & 1, ids_out, scores_out, gen_len_out, st));
  return join_stream(e, user);
}

int gic_kv_reorder(gic_engine* e, const void* kv_src, void* kv_dst, const int32_t* beam_idx, int rows, int ctx_len, int t_max, void* stream) {
  GIC_REQUIRE(e && kv_src && kv_dst && beam_idx, "null argument");
  GIC_REQUIRE(rows > 0 && ctx_len >= 0 && ctx_len <= t_max, "bad sizes rows=%d ctx_len=%d t_max=%d", rows, ctx_len, t_max);
  if (kv16(e))
    return launch_kv_reorder<bf16>((const bf16*)kv_src, (bf16*)kv_dst, beam_idx, e->L, rows, e->H, ctx_len, t_max, (cudaStream_t)stream);
  return launch_kv_reorder<float>((const float*)kv_src, (float*)kv_dst, beam_idx, e->L, rows, e->H, ctx_len, t_max, (cudaStream_t)stream);
}

size_t gic_topk_workspace_bytes(int batch, int n_rows, int dim, int k) { return gic::topk_workspace_bytes(batch, n_rows, dim, k); }

int gic_topk_tc_supported(int dim, int k) { return gic::topk_tc_supported(dim, k) ? 1 : 0; }
size_t gic_topk_tc_workspace_bytes(int batch, int n_rows, int dim, int k) { return gic::topk_tc_workspace_bytes(batch, n_rows, dim, k); }

int gic_pack_bf16x2(const float* src, void* hi, void* lo, size_t n, void* stream) {
  GIC_REQUIRE(src && hi && lo, "null argument");
  ActOut o; o.hi = (bf16*)hi; o.lo = (bf16*)lo;
  return launch_convert(src, o, n, (cudaStream_t)stream);
}

int gic_topk_ip_tc(const float* q, const float* db, const void* db_hi, const void* db_lo, float db_norm_max, int batch, int n_rows, int dim, int k,
                   float* scores_out, int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream) {
  GIC_REQUIRE(q && db && db_hi && db_lo && scores_out && idx_out && workspace, "null argument");
  GIC_TRY(gic_device_check());
  return launch_topk_ip_tc(q, db, (const bf16*)db_hi, (const bf16*)db_lo, db_norm_max, batch, n_rows, dim, k, scores_out, idx_out, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

int gic_topk_ip(const float* q, const float* db, int batch, int n_rows, int dim, int k, float* scores_out, int64_t* idx_out, void* workspace,
                size_t workspace_bytes, void* stream) {
  GIC_REQUIRE(q && db && scores_out && idx_out && workspace, "null argument");
  return launch_topk_ip(q, db, batch, n_rows, dim, k, scores_out, idx_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int gic_select_caption_rows(const float* scores, const int64_t* idx, int batch, int k_searched, const int64_t* cap_row_start,
                            const int64_t* cap_row_ids, int top_i, int top_k, int64_t* rows_out, void* stream) {
  GIC_REQUIRE(scores && idx && cap_row_start && rows_out, "null argument");
  return launch_select_caption_rows(scores, idx, batch, k_searched, cap_row_start, cap_row_ids, top_i, top_k, rows_out,
                                    (cudaStream_t)stream);
}

int gic_gather_caption_rows(const float* cap_db, const int64_t* rows, int batch, int top_k, int dim, float* out, void* stream) {
  GIC_REQUIRE(cap_db && rows && out, "null argument");
  GIC_REQUIRE(batch >= 0 && top_k > 0 && dim > 0, "bad sizes batch=%d top_k=%d dim=%d", batch, top_k, dim);
  return launch_gather_caption_rows(cap_db, rows, batch * top_k, dim, out, (cudaStream_t)stream);
}

int gic_gather_attention_add(const float* q, const float* cap_db, const int64_t* rows, int batch, int top_k, int dim, const float* attn_w,
                             const float* attn_b, float* out, void* stream) {
  GIC_REQUIRE(q && cap_db && rows && attn_w && attn_b && out, "null argument");
  return launch_gather_attention_add(q, cap_db, rows, batch, top_k, dim, attn_w, attn_b, out, (cudaStream_t)stream);
}

int gic_gather_aggregate_add(const float* q, const float* cap_db, const int64_t* rows, int batch, int top_k, int dim, int aggregation,
                             float* out, void* stream) {
  GIC_REQUIRE(q && cap_db && rows && out, "null argument");
  GIC_REQUIRE(batch > 0, "batch must be positive");
  return launch_gather_aggregate_add(q, cap_db, rows, batch, top_k, dim, aggregation, out, (cudaStream_t)stream);
}

unsigned long long gic_launch_count(void) { return gic::g_launches.load(std::memory_order_relaxed); }
unsigned long long gic_compaction_count(void) { return gic::g_compactions.load(std::memory_order_relaxed); }

int gic_profile_enable(gic_engine* e, int on) {
  GIC_REQUIRE(e != nullptr, "null engine");
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof.clear();
  e->profiling = on != 0;
  return GIC_OK;
}

int gic_profile_read(gic_engine* e, gic_profile_entry* out, int max_entries, int* n_out) {
  GIC_REQUIRE(e && out && n_out && max_entries > 0, "bad argument");
  GIC_CHECK_CUDA(cudaDeviceSynchronize());
  int n = 0;
  for (auto& r : e->prof) {
    float ms = 0.f;
    GIC_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    int j = 0;
    for (; j < n; ++j) if (strncmp(out[j].name, r.cat, sizeof(out[j].name)) == 0) break;
    if (j == n) {
      if (n == max_entries) continue;
      memset(&out[n], 0, sizeof(out[n]));
      strncpy(out[n].name, r.cat, sizeof(out[n].name) - 1);
      ++n;
    }
    out[j].launches += 1;
    out[j].total_ms += ms;
  }
  *n_out = n;
  return GIC_OK;
}

// ---- kernel-level test entry points ---------------------------------------------------------------------------------
int gic_test_gemm(int dtype, const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int epilogue, void* stream) {
  GIC_REQUIRE(A && W && C, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GIC_DTYPE_F32) return launch_sgemm_nt(A, K, W, bias, C, N, M, N, K, epilogue, st);
  GIC_REQUIRE(dtype == GIC_DTYPE_BF16 || dtype == GIC_DTYPE_BF16X2, "unknown dtype %d", dtype);
  GIC_TRY(gic_device_check());
  GIC_TRY(gic::tma_init());
  GIC_TRY(gic::gemm_bf16_configure());
  const bool split = dtype == GIC_DTYPE_BF16X2;
  bf16 *a_hi = nullptr, *a_lo = nullptr, *w_hi = nullptr, *w_lo = nullptr;
  const size_t na = (size_t)M * K, nw = (size_t)N * K;
  GIC_CHECK_CUDA(cudaMalloc(&a_hi, na * 2));
  GIC_CHECK_CUDA(cudaMalloc(&w_hi, nw * 2));
  if (split) {
    GIC_CHECK_CUDA(cudaMalloc(&a_lo, na * 2));
    GIC_CHECK_CUDA(cudaMalloc(&w_lo, nw * 2));
  }
  ActOut ao; ao.hi = a_hi; ao.lo = a_lo;
  ActOut wo; wo.hi = w_hi; wo.lo = w_lo;
  int r = launch_convert(A, ao, na, st);
  if (r == GIC_OK) r = launch_convert(W, wo, nw, st);
  GemmBf16Args g;
  int bn = 0, pair = 0;
  const bool pair_kernel = !split && epilogue == EPI_NONE && N % 32 == 0;  // the one fp32-output CTA-pair instantiation
  gemm_bf16_pick(M, N, K, split ? 1 : 0, 1, &bn, pair_kernel ? &pair : nullptr);
  if (split && bn > 64 && epilogue != EPI_NONE) bn = 64;  // the wide bf16x2 tiles carry the engine's epilogues only
  g.pair = pair;
  if (r == GIC_OK) r = make_tma_2d_bf16(&g.a_hi, a_hi, M, K, K, 128);
  if (r == GIC_OK) r = make_tma_2d_bf16(&g.w_hi, w_hi, N, K, K, pair ? bn / 2 : bn);
  if (r == GIC_OK && split) r = make_tma_2d_bf16(&g.a_lo, a_lo, M, K, K, 128);
  if (r == GIC_OK && split) r = make_tma_2d_bf16(&g.w_lo, w_lo, N, K, K, bn);
  g.M = M; g.N = N; g.K = K; g.block_n = bn; g.split = split; g.epilogue = epilogue; g.bias = bias;
  g.out.f32 = C; g.ld_out = N;
  if (r == GIC_OK) r = launch_gemm_bf16(g, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(a_hi); cudaFree(w_hi); cudaFree(a_lo); cudaFree(w_lo);
  if (r != GIC_OK) return r;
  GIC_CHECK_CUDA(ce);
  return GIC_OK;
}

// The fused bf16 MLP sub-block exactly as a decode / prefill layer runs it, on caller data (kernel-level parity test):
//   xb, stats = row_stats(h);  f = gelu(LN_folded(xb) . Wfc^T + bfc)  [bf16];  h += f . Wfc2^T + bfc2  (+ bf16 copy + statistics),
// fc2 optionally K-split.  h [M,d] fp32 in/out; hb_out [M,d] bf16; stats_out [d/32][M] float2 (sum, sum of squares of hb_out).
int gic_test_ln_mlp(float* h, const float* gamma, const float* beta, const float* wfc, const float* bfc, const float* wfc2, const float* bfc2,
                    void* hb_out, void* stats_out, int M, int d, int split_k, void* stream) {
  GIC_REQUIRE(h && gamma && beta && wfc && bfc && wfc2 && bfc2 && hb_out && stats_out, "null argument");
  GIC_REQUIRE(d % 64 == 0 && split_k >= 1 && split_k <= 4, "d must be a multiple of 64 and split_k in [1, 4]");
  cudaStream_t st = (cudaStream_t)stream;
  GIC_TRY(gic_device_check());
  GIC_TRY(gic::tma_init());
  GIC_TRY(gic::gemm_bf16_configure());
  const int d4 = 4 * d;
  bf16 *xb = nullptr, *wfc_p = nullptr, *wfc2_p = nullptr, *f = nullptr;
  float *colsum = nullptr, *bias_f = nullptr, *ws = nullptr;
  float2* stats = nullptr;
  int* counters = nullptr;
  std::vector<void*> owned;
  auto take = [&](void** p, size_t bytes) { cudaError_t ce = cudaMalloc(p, bytes); if (ce == cudaSuccess) owned.push_back(*p); return ce; };
  int r = GIC_OK;
  do {
    if (take((void**)&xb, (size_t)M * d * 2) || take((void**)&wfc_p, (size_t)d4 * d * 2) || take((void**)&wfc2_p, (size_t)d * d4 * 2) ||
        take((void**)&f, (size_t)M * d4 * 2) || take((void**)&colsum, (size_t)d4 * 4) || take((void**)&bias_f, (size_t)d4 * 4) ||
        take((void**)&stats, (size_t)M * 8) || take((void**)&ws, (size_t)split_k * M * d * 4) || take((void**)&counters, 4096 * 4)) {
      set_error("gic_test_ln_mlp: cudaMalloc failed"); r = GIC_ERR_CUDA; break;
    }
    if (cudaMemsetAsync(counters, 0, 4096 * 4, st) != cudaSuccess) { set_error("memset failed"); r = GIC_ERR_CUDA; break; }
    ActOut o1; o1.hi = wfc_p;
    ActOut o2; o2.hi = wfc2_p;
    if ((r = launch_pack_weight(wfc, d4, d, false, o1, st, gamma)) != GIC_OK) break;      // nn.Linear layout [N,K], gamma folded
    if ((r = launch_fold_ln(wfc_p, wfc, false, beta, bfc, colsum, bias_f, d4, d, st)) != GIC_OK) break;
    if ((r = launch_pack_weight(wfc2, d, d4, false, o2, st)) != GIC_OK) break;
    if ((r = launch_row_stats(h, d, xb, stats, M, d, st)) != GIC_OK) break;
    {
      GemmBf16Args g;
      int bn = 0, pair = 0;
      gemm_bf16_pick(M, d4, d, 0, 1, &bn, &pair);
      g.pair = pair;
      if ((r = make_tma_2d_bf16(&g.a_hi, xb, M, d, d, 128)) != GIC_OK) break;
      if ((r = make_tma_2d_bf16(&g.w_hi, wfc_p, d4, d, d, pair ? bn / 2 : bn)) != GIC_OK) break;
      g.M = M; g.N = d4; g.K = d; g.block_n = bn; g.epilogue = EPI_GELU; g.bias = bias_f; g.out.hi = f; g.ld_out = d4;
      g.ln_stats = stats; g.ln_parts = 1; g.ln_stats_ld = M; g.ln_colsum = colsum;
      if ((r = launch_gemm_bf16(g, st)) != GIC_OK) break;
    }
    {
      GemmBf16Args g;
      int bn = 0, pair = 0;
      gemm_bf16_pick(M, d, d4, 0, split_k, &bn, &pair);
      g.pair = pair;
      if ((r = make_tma_2d_bf16(&g.a_hi, f, M, d4, d4, 128)) != GIC_OK) break;
      if ((r = make_tma_2d_bf16(&g.w_hi, wfc2_p, d, d4, d4, pair ? bn / 2 : bn)) != GIC_OK) break;
      g.M = M; g.N = d; g.K = d4; g.block_n = bn; g.epilogue = EPI_RESIDUAL; g.bias = bfc2; g.out.f32 = h; g.out.hi = (bf16*)hb_out; g.ld_out = d;
      g.stats_out = (float2*)stats_out; g.ln_stats_ld = M;
      if (split_k > 1) { g.split_k = split_k; g.splitk_ws = ws; g.splitk_counters = counters; }
      if ((r = launch_gemm_bf16(g, st)) != GIC_OK) break;
    }
  } while (0);
  cudaError_t ce = cudaStreamSynchronize(st);
  for (void* p : owned) cudaFree(p);
  if (r != GIC_OK) return r;
  GIC_CHECK_CUDA(ce);
  return GIC_OK;
}

int gic_test_layernorm(const float* x, const float* w, const float* b, float* y, int rows, int d, void* stream) {
  GIC_REQUIRE(x && w && b && y, "null argument");
  ActOut o; o.f32 = y;
  return launch_layernorm(x, d, w, b, o, rows, d, (cudaStream_t)stream);
}

// One decode-attention launch on caller data (kernel-level parity hook for GPT2Attention with T_q = 1 over a KV cache,
// HF:models/gpt2/modeling_gpt2.py:185-220): qkv [rows, 3 H 64] bf16, K / V caches [rows][H][t_max][64] bf16 holding `pos` tokens;
// appends the new token's K / V at position pos and writes out [rows, H 64] bf16.  variant: ring / kernel shape (0 = product).
int gic_test_attn_decode(const void* qkv, void* kcache, void* vcache, void* out, int pos, int rows, int H, int t_max, int variant, void* stream) {
  GIC_REQUIRE(qkv && kcache && vcache && out, "null argument");
  GIC_REQUIRE(pos >= 0 && pos < t_max && rows > 0 && H > 0, "bad shape: pos %d t_max %d rows %d H %d", pos, t_max, rows, H);
  cudaStream_t st = (cudaStream_t)stream;
  GIC_TRY(gic_device_check());
  GIC_TRY(gic::attn_decode_configure());
  int* d_pos = nullptr;
  GIC_CHECK_CUDA(cudaMalloc((void**)&d_pos, sizeof(int)));
  cudaError_t ce = cudaMemcpyAsync(d_pos, &pos, sizeof(int), cudaMemcpyHostToDevice, st);
  int r = GIC_OK;
  if (ce == cudaSuccess) {
    ActOut o; o.hi = (bf16*)out;
    gic::attn_decode_set_variant(variant);
    r = launch_attn_decode<bf16>((const bf16*)qkv, (bf16*)kcache, (bf16*)vcache, o, d_pos, rows, H, t_max, st);
    gic::attn_decode_set_variant(-1);
    ce = cudaStreamSynchronize(st);
  }
  cudaFree(d_pos);
  if (r != GIC_OK) return r;
  GIC_CHECK_CUDA(ce);
  return GIC_OK;
}

// Kernel-level hook of the beam-search decode attention (ancestry table instead of a reordered cache; see the header)
int gic_test_attn_decode_beam(const void* qkv, void* kcache, void* vcache, void* out, void* out_lo, const int32_t* anc, int anc_ld, int pos, int rows, int H,
                              int t_max, int n_prefix, int beams, int f16, int shared, void* stream) {
  GIC_REQUIRE(qkv && kcache && vcache && out && anc, "null argument");
  GIC_REQUIRE((f16 != 0) == (out_lo != nullptr), "the fp16 flavour writes hi + lo outputs, the bf16 one a single output");
  GIC_REQUIRE(pos >= n_prefix && pos < t_max && rows > 0 && H > 0 && beams >= 1 && rows % beams == 0 && anc_ld >= pos - n_prefix,
              "bad shape: pos %d n_prefix %d t_max %d rows %d H %d beams %d anc_ld %d", pos, n_prefix, t_max, rows, H, beams, anc_ld);
  cudaStream_t st = (cudaStream_t)stream;
  GIC_TRY(gic_device_check());
  GIC_TRY(gic::attn_decode_configure());
  int* d_pos = nullptr;
  GIC_CHECK_CUDA(cudaMalloc((void**)&d_pos, sizeof(int)));
  cudaError_t ce = cudaMemcpyAsync(d_pos, &pos, sizeof(int), cudaMemcpyHostToDevice, st);
  int r = GIC_OK;
  if (ce == cudaSuccess) {
    gic::attn_decode_set_beam_shared(shared ? 1 : 0);
    r = launch_attn_decode_indirect((const bf16*)qkv, (bf16*)kcache, (bf16*)vcache, (bf16*)out, d_pos, rows, H, t_max, anc, anc_ld, n_prefix, beams, st, (bf16*)out_lo);
    gic::attn_decode_set_beam_shared(-1);
    ce = cudaStreamSynchronize(st);
  }
  cudaFree(d_pos);
  if (r != GIC_OK) return r;
  GIC_CHECK_CUDA(ce);
  return GIC_OK;
}

// Measurement hook: back-to-back product decode-attention launches over `planes` cache planes, asynchronous on `stream` (see the header).
int gic_bench_attn_decode(const void* qkv, void* kcache, void* vcache, void* out, const int* d_pos, int rows, int H, int t_max, int planes,
                          int launches, void* stream) {
  GIC_REQUIRE(qkv && kcache && vcache && out && d_pos, "null argument");
  GIC_REQUIRE(rows > 0 && H > 0 && t_max > 0 && planes > 0 && launches > 0, "bad shape: rows %d H %d t_max %d planes %d launches %d", rows, H, t_max, planes, launches);
  cudaStream_t st = (cudaStream_t)stream;
  GIC_TRY(gic_device_check());
  GIC_TRY(gic::attn_decode_configure());
  const size_t plane = (size_t)rows * H * t_max * 64;
  ActOut o; o.hi = (bf16*)out;
  for (int i = 0; i < launches; ++i) {
    const size_t off = (size_t)(i % planes) * plane;
    GIC_TRY(launch_attn_decode<bf16>((const bf16*)qkv, (bf16*)kcache + off, (bf16*)vcache + off, o, d_pos, rows, H, t_max, st));
  }
  return GIC_OK;
}

// One causal prefill-attention launch on caller data (GPT2Attention over the S prefix tokens of every row, HF:models/gpt2/modeling_gpt2.py:185-220):
// qkv [rows * S, 3 H 64] bf16 -> out [rows * S, H 64] bf16, K / V written to caches [rows][H][t_max][64] at positions 0..S-1.
int gic_test_attn_prefill(const void* qkv, void* kcache, void* vcache, void* out, int rows, int S, int H, int t_max, void* stream) {
  GIC_REQUIRE(qkv && kcache && vcache && out, "null argument");
  GIC_REQUIRE(S >= 1 && S <= t_max && rows > 0 && H > 0, "bad shape: S %d t_max %d rows %d H %d", S, t_max, rows, H);
  cudaStream_t st = (cudaStream_t)stream;
  GIC_TRY(gic_device_check());
  ActOut o; o.hi = (bf16*)out;
  GIC_TRY(launch_attn_prefill<bf16>((const bf16*)qkv, (bf16*)kcache, (bf16*)vcache, o, rows, S, H, t_max, 1, st));
  GIC_CHECK_CUDA(cudaStreamSynchronize(st));
  return GIC_OK;
}

// one sampling launch on caller data: logits dev fp32 [B, V] -> tokens dev int32 [B] (kernel-level hook for src/models.py:400-449)
int gic_test_sample_top_p(const float* logits, int B, int V, float temperature, float top_p, unsigned long long seed, int step, int32_t* tokens_out,
                          void* stream) {
  GIC_REQUIRE(logits && tokens_out && B > 0 && V > 0 && step >= 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  GIC_TRY(gic_device_check());
  float* pv = nullptr;
  GIC_CHECK_CUDA(cudaMalloc((void**)&pv, (size_t)B * sizeof(float)));
  int r = launch_sample_top_p(logits, B, V, temperature, top_p, seed, nullptr, step, pv, tokens_out, 1, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(pv);
  if (r != GIC_OK) return r;
  GIC_CHECK_CUDA(ce);
  return GIC_OK;
}

// in-situ timeline of the decode-step kernels (tools/step_timeline.py): install a device buffer of `cap` (kind, begin_ns, end_ns) uint64
// records, pre-filled with (0, ~0, 0); every GEMM / decode attention / ln_f / finalize launch takes the next record.  buf = NULL uninstalls.
int gic_trace_install(void* buf, unsigned int cap) {
  gic::g_step_trace.buf = (unsigned long long*)buf;
  gic::g_step_trace.slot = 0;
  gic::g_step_trace.cap = buf ? cap : 0;
  ++gic::g_trace_gen;  // graphs captured under the previous descriptor are re-captured
  return GIC_OK;
}

}  // extern "C"
