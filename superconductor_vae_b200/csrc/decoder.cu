// EnhancedTransformerDecoder engine: weight layout, memory builder, KV-cache decode loop.
// Reference: src/superconductor/models/autoregressive_decoder.py:544-899 (module + memory),
// :1175-1557 (KV-cache decode).  See DESIGN.md for the data layout and the kernel list.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/scvae_b200.h"
#include "decode_kernels.cuh"
#include "weights.cuh"

using namespace scv;

namespace {

struct DecLayer {
  LNp n1, n2, n3;
  __nv_bfloat16* sa_in_w = nullptr; float* sa_in_b = nullptr; int sa_in_ld = 0;   // [3d, d]
  __nv_bfloat16* sa_in_wt = nullptr;                      // tcgen05 tiles of the whole [3d, d]
  __nv_bfloat16* ca_q_wt = nullptr; __nv_bfloat16* ca_kv_wt = nullptr;   // tiles of rows [0,d) and [d,3d)
  Lin sa_out;
  __nv_bfloat16* ca_in_w = nullptr; float* ca_in_b = nullptr; int ca_in_ld = 0;   // [3d, d] rows q;k;v
  Lin ca_out;
  Lin ff1, ff2;
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    SCV_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

constexpr int kMaxSub = 4;

// Everything that is baked into the kernel arguments of one decode step.
struct GraphKey {
  int B, M, steps_max, n_sub, top_k, has_masks, want_lp, want_ent, has_forced, prof, mem_rows;
  unsigned flags;
  float temperature, top_p, stop_boost, hard_stop, site_dup;
  bool operator==(const GraphKey& o) const {
    return B == o.B && M == o.M && steps_max == o.steps_max && n_sub == o.n_sub && top_k == o.top_k &&
           has_masks == o.has_masks && want_lp == o.want_lp && want_ent == o.want_ent && has_forced == o.has_forced && prof == o.prof && mem_rows == o.mem_rows &&
           flags == o.flags && temperature == o.temperature && top_p == o.top_p && stop_boost == o.stop_boost &&
           hard_stop == o.hard_stop && site_dup == o.site_dup;
  }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; };

}  // namespace

struct scv_decoder {
  scv_decoder_config cfg{};
  WeightStore ws;
  __nv_bfloat16* emb = nullptr; int ld_emb = 0;
  float* pe = nullptr;
  Lin l2m_a, l2m_b; LNp l2m_ln;
  Lin s2m_a, s2m_b; LNp s2m_ln;
  Lin h2m_a, h2m_b, h2m_c; LNp h2m_ln;
  Lin skip_a, skip_b;
  std::vector<DecLayer> layers;
  LNp out_ln; Lin out_a, out_b;
  Lin stop_a, stop_b, dup_a, dup_b;
  LNp tt_ln; Lin tt_a, tt_b, tt_c;
  // workspaces (engine-owned, grown on demand)
  DevBuf x, xn, qkv, attn, q2, ff, h1, h2, t3, logits, tlog, slog, ckv, kvpool, cur, fin, ptab, state, mtmp;
  DevBuf xn_s, attn_s, ff_s, h2_s;   // SplitTile (bf16 hi/lo) activations feeding the tcgen05 projections
  DevBuf msplit;                     // SplitTile scratch of the memory builder / memory-token projection
  DevBuf seen, dlog;                 // site-dup gating: [B, V] seen-element bitmap, [B] site_dup_head logit
  DevBuf row_map;                    // compaction of finished rows: slot -> row of the call's batch
  DevBuf o_tok, o_lp, o_ent, masks_buf, forced_buf;   // engine-owned I/O so that a captured step never bakes caller pointers
  std::vector<GraphEntry> graphs;    // one instantiated CUDA graph of a decode step per call configuration
  size_t ws_signature = 0;
  unsigned tune_epoch_seen = 0;
  int* pinned = nullptr;              // host-pinned: [0..1] done polls, [2..9] StepState copy
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t main = nullptr;             // the decode loop's own stream
  cudaStream_t sub[kMaxSub] = {};          // sub-batch streams (created on first use)
  cudaEvent_t ev_sub[kMaxSub] = {};
  cudaEvent_t ev_fork = nullptr;
  int last_B = 0;
  int launches_per_step = 0;               // kernels in one step (what a replayed graph launches)
  // small-batch persistent step (decode_small.cu)
  DevBuf fw_skip;                          // teacher-forced forward: [B, L] key padding bytes
  DevBuf sm_phases, sm_bar, sm_h2b, sm_t3s, sm_t3d;
  std::vector<SmallPhase> sm_host;
  DevBuf sm_part;                          // per-CTA partial sums of the fused feed-forward block
  int sm_n_phases = 0, sm_grid = 0;
  bool small_active = false;               // this call decodes through the persistent small-batch kernel
  // cluster-parallel small-batch decode (decode_cluster.cu)
  DevBuf cl_instr, cl_dbg;
  std::vector<int> cl_kinds;
  ClProgram cl_prog{};
  bool cl_active = false;

  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }

  int ensure_streams() {
    if (ev_fork != nullptr) return 0;
    for (int i = 0; i < kMaxSub; ++i) {
      SCV_CUDA(cudaStreamCreateWithFlags(&sub[i], cudaStreamNonBlocking));
      SCV_CUDA(cudaEventCreateWithFlags(&ev_sub[i], cudaEventDisableTiming));
    }
    SCV_CUDA(cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking));
    SCV_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    return 0;
  }

  ~scv_decoder() {
    for (DevBuf* b : {&x, &xn, &qkv, &attn, &q2, &ff, &h1, &h2, &t3, &logits, &tlog, &slog, &ckv, &kvpool, &cur,
                      &fin, &ptab, &state, &mtmp, &xn_s, &attn_s, &ff_s, &h2_s, &o_tok, &o_lp, &o_ent, &masks_buf,
                      &forced_buf, &seen, &dlog, &msplit, &sm_phases, &sm_bar, &sm_h2b, &sm_t3s, &sm_t3d, &fw_skip, &cl_instr, &cl_dbg, &sm_part, &row_map})
      b->release();
    drop_graphs();
    if (pinned) cudaFreeHost(pinned);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : ev_sub) if (e) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (auto& st_ : sub) if (st_) cudaStreamDestroy(st_);
    if (main) cudaStreamDestroy(main);
  }
};

static int dec_register(scv_decoder* D) {
  const scv_decoder_config& c = D->cfg;
  const int d = c.d_model;
  WeightStore& W = D->ws;
  D->emb = W.add_matrix("token_embedding.weight", c.vocab_size, d, &D->ld_emb);
  D->pe = W.add_vector("pos_encoding.pe", (int64_t)c.pe_len * d);
  if (!D->emb || !D->pe) return 2;
  const int n_lat = d * c.n_memory_tokens;
  if (c.memory_bottleneck_dim > 0) {
    SCV_TRY(W.add_linear("latent_to_memory.0", c.memory_bottleneck_dim, c.latent_dim, &D->l2m_a, true, true));
    SCV_TRY(W.add_layernorm("latent_to_memory.1", c.memory_bottleneck_dim, &D->l2m_ln));
    SCV_TRY(W.add_linear("latent_to_memory.3", n_lat, c.memory_bottleneck_dim, &D->l2m_b, true, true));
  } else {
    SCV_TRY(W.add_linear("latent_to_memory.0", n_lat / 2, c.latent_dim, &D->l2m_a, true, true));
    SCV_TRY(W.add_linear("latent_to_memory.2", n_lat, n_lat / 2, &D->l2m_b, true, true));
  }
  if (c.skip_n_tokens > 0) {
    const int n_skip = d * c.skip_n_tokens;
    SCV_TRY(W.add_linear("skip_to_memory.0", n_skip / 2, c.encoder_skip_dim, &D->skip_a));
    SCV_TRY(W.add_linear("skip_to_memory.2", n_skip, n_skip / 2, &D->skip_b, true, true));
  }
  if (c.n_stoich_tokens > 0) {
    SCV_TRY(W.add_linear("stoich_to_memory.0", d, c.stoich_input_dim, &D->s2m_a));
    SCV_TRY(W.add_layernorm("stoich_to_memory.1", d, &D->s2m_ln));
    SCV_TRY(W.add_linear("stoich_to_memory.3", d * c.n_stoich_tokens, d, &D->s2m_b, true, true));
  }
  if (c.heads_n_tokens > 0) {
    SCV_TRY(W.add_linear("heads_to_memory.0", d / 2, c.heads_input_dim, &D->h2m_a));
    SCV_TRY(W.add_layernorm("heads_to_memory.1", d / 2, &D->h2m_ln));
    SCV_TRY(W.add_linear("heads_to_memory.3", d, d / 2, &D->h2m_b, true, true));
    SCV_TRY(W.add_linear("heads_to_memory.5", d * c.heads_n_tokens, d, &D->h2m_c, true, true));
  }
  D->layers.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    DecLayer& L = D->layers[i];
    const std::string p = "transformer_decoder.layers." + std::to_string(i) + ".";
    L.sa_in_w = W.add_matrix(p + "self_attn.in_proj_weight", 3 * d, d, &L.sa_in_ld);
    L.sa_in_b = W.add_vector(p + "self_attn.in_proj_bias", 3 * d);
    L.ca_in_w = W.add_matrix(p + "multihead_attn.in_proj_weight", 3 * d, d, &L.ca_in_ld);
    L.ca_in_b = W.add_vector(p + "multihead_attn.in_proj_bias", 3 * d);
    if (!L.sa_in_w || !L.sa_in_b || !L.ca_in_w || !L.ca_in_b) return 2;
    L.sa_in_wt = W.add_tiled_view(p + "self_attn.in_proj_weight", 0, 3 * d);
    L.ca_q_wt = W.add_tiled_view(p + "multihead_attn.in_proj_weight", 0, d);
    L.ca_kv_wt = W.add_tiled_view(p + "multihead_attn.in_proj_weight", d, 2 * d);
    if (!L.sa_in_wt || !L.ca_q_wt || !L.ca_kv_wt) return 2;
    SCV_TRY(W.add_linear(p + "self_attn.out_proj", d, d, &L.sa_out, true, true));
    SCV_TRY(W.add_linear(p + "multihead_attn.out_proj", d, d, &L.ca_out, true, true));
    SCV_TRY(W.add_linear(p + "linear1", c.dim_feedforward, d, &L.ff1, true, true));
    SCV_TRY(W.add_linear(p + "linear2", d, c.dim_feedforward, &L.ff2, true, true));
    SCV_TRY(W.add_layernorm(p + "norm1", d, &L.n1));
    SCV_TRY(W.add_layernorm(p + "norm2", d, &L.n2));
    SCV_TRY(W.add_layernorm(p + "norm3", d, &L.n3));
  }
  SCV_TRY(W.add_layernorm("output_proj.0", d, &D->out_ln));
  SCV_TRY(W.add_linear("output_proj.1", d, d, &D->out_a, true, true));
  SCV_TRY(W.add_linear("output_proj.4", c.vocab_size, d, &D->out_b, true, true));
  SCV_TRY(W.add_linear("stop_head.0", d / 4, d, &D->stop_a, true, true));
  SCV_TRY(W.add_linear("stop_head.2", 1, d / 4, &D->stop_b));
  SCV_TRY(W.add_linear("site_dup_head.0", d / 4, d, &D->dup_a, true, true));
  SCV_TRY(W.add_linear("site_dup_head.2", 1, d / 4, &D->dup_b));
  for (const char* n : {"site_dup_head.0.weight", "site_dup_head.0.bias", "site_dup_head.2.weight", "site_dup_head.2.bias"})
    W.mark_optional(n);   // older checkpoints lack it; only used when site_dup_threshold > 0
  SCV_TRY(W.add_layernorm("token_type_head.0", d, &D->tt_ln));
  SCV_TRY(W.add_linear("token_type_head.1", d, d, &D->tt_a, true, true));
  SCV_TRY(W.add_linear("token_type_head.4", d / 4, d, &D->tt_b, true, true));
  SCV_TRY(W.add_linear("token_type_head.7", 5, d / 4, &D->tt_c));
  return 0;
}

static LinearArgs lin_args(const float* x, int ldx, const Lin& L, float* y, int ldy, int M, int act,
                           const int* done = nullptr) {
  LinearArgs a;
  a.x = x; a.ldx = ldx; a.w = L.w; a.ldw = L.ldw; a.wt = L.wt; a.bias = L.b; a.y = y; a.ldy = ldy;
  a.M = M; a.N = L.N; a.K = L.K; a.act = act; a.done_flag = done;
  return a;
}

extern "C" {

int scv_decoder_create(const scv_decoder_config* cfg, scv_decoder** out) {
  SCV_REQUIRE(cfg && out, "null argument");
  SCV_REQUIRE(cfg->d_model > 0 && cfg->nhead > 0 && cfg->d_model % cfg->nhead == 0,
              "d_model %d must be a positive multiple of nhead %d", cfg->d_model, cfg->nhead);
  SCV_REQUIRE(cfg->d_model / cfg->nhead <= 128, "head_dim > 128 is not supported");
  SCV_REQUIRE(cfg->num_layers > 0 && cfg->vocab_size > kEndIdx && cfg->pe_len > 1 && cfg->dim_feedforward > 0,
              "bad decoder shape");
  SCV_REQUIRE(cfg->latent_dim > 0 && cfg->n_memory_tokens > 0, "bad memory shape");
  int dev = 0;
  SCV_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SCV_CUDA(cudaGetDeviceProperties(&prop, dev));
  SCV_REQUIRE(prop.major == 10, "scvae_b200 is built for sm_100a; device %s is sm_%d%d", prop.name, prop.major,
              prop.minor);
  scv_decoder* D = new scv_decoder();
  D->cfg = *cfg;
  int rc = dec_register(D);
  if (rc == 0 && cudaMallocHost(reinterpret_cast<void**>(&D->pinned), 32 * sizeof(int)) != cudaSuccess) {
    set_error("cudaMallocHost failed");
    rc = 2;
  }
  if (rc == 0) {
    for (auto& e : D->ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { set_error("event create failed"); rc = 2; }
  }
  if (rc != 0) { delete D; return rc; }
  *out = D;
  return 0;
}

void scv_decoder_destroy(scv_decoder* dec) { delete dec; }

int scv_decoder_load_weight(scv_decoder* dec, const char* name, const float* src, int64_t numel, void* stream) {
  SCV_REQUIRE(dec && name && src, "null argument");
  return dec->ws.load(name, src, numel, static_cast<cudaStream_t>(stream));
}

int scv_decoder_missing_weights(scv_decoder* dec) {
  std::string first;
  const int n = dec->ws.missing(&first);
  if (n > 0) set_error("%d state_dict entries not loaded, first: %s", n, first.c_str());
  return n;
}

int scv_decoder_build_memory(scv_decoder* D, int32_t B, const float* z, const float* skip, const float* stoich,
                             const float* heads_in, float* memory_out, int32_t* n_tokens_out, void* stream) {
  SCV_REQUIRE(D && z && memory_out && B > 0, "build_memory: bad arguments");
  SCV_REQUIRE(scv_decoder_missing_weights(D) == 0, "build_memory: weights missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const scv_decoder_config& c = D->cfg;
  const int d = c.d_model;
  const bool use_skip = skip != nullptr && c.skip_n_tokens > 0;          // (:806)
  const bool use_stoich = stoich != nullptr && c.n_stoich_tokens > 0;    // (:813)
  const bool use_heads = heads_in != nullptr && c.heads_n_tokens > 0;    // (:821)
  const int M = c.n_memory_tokens + (use_skip ? c.skip_n_tokens : 0) + (use_stoich ? c.n_stoich_tokens : 0) +
                (use_heads ? c.heads_n_tokens : 0);
  if (n_tokens_out) *n_tokens_out = M;
  const int ldm = M * d;
  const int hid = std::max(std::max(D->l2m_a.N, d), c.skip_n_tokens > 0 ? D->skip_a.N : 0);
  SCV_TRY(D->mtmp.ensure((size_t)B * (hid + d) * sizeof(float)));
  float* t0 = D->mtmp.as<float>();
  float* t1 = t0 + (size_t)B * hid;
  int col = 0;
  // latent tokens (:800-801).  These two projections carry 85 % of the memory builder's flops: hand the activations
  // over as SplitTiles (z split once, the hidden layer written split by the first GEMM's epilogue) instead of
  // re-splitting fp32 rows inside every column tile of the GEMM.
  const int Hd = D->l2m_a.N;
  static const int forced_simt = [] { const char* e = getenv("SCV_LINEAR_IMPL"); return e ? atoi(e) : 0; }();
  const bool split_path = forced_simt != 1 && c.latent_dim % 64 == 0 && Hd % 64 == 0 && (d * c.n_memory_tokens) % 4 == 0 &&
                          (c.memory_bottleneck_dim == 0 || Hd <= 1024) && ldm % 4 == 0;
  if (split_path) {
    const size_t zs_bytes = split_tile_bytes(B, c.latent_dim), hs_bytes = split_tile_bytes(B, Hd);
    SCV_TRY(D->msplit.ensure(zs_bytes + hs_bytes));
    unsigned char* zs = D->msplit.as<unsigned char>();
    unsigned char* hs = zs + zs_bytes;
    SCV_TRY(launch_layernorm_split(z, c.latent_dim, nullptr, nullptr, zs, B, c.latent_dim, 0, nullptr, s));
    LinearArgs g1 = lin_args(z, c.latent_dim, D->l2m_a, t0, Hd, B, c.memory_bottleneck_dim > 0 ? ACT_NONE : ACT_GELU);
    g1.a_split = zs;
    if (c.memory_bottleneck_dim > 0) {
      SCV_TRY(launch_linear(g1, 0, s));                                             // fp32 rows for the LayerNorm
      SCV_TRY(launch_layernorm_split(t0, Hd, D->l2m_ln.g, D->l2m_ln.b, hs, B, Hd, 1, nullptr, s, ACT_GELU));
    } else {
      g1.y_split = hs;
      SCV_TRY(launch_linear(g1, 0, s));
    }
    LinearArgs g2 = lin_args(t0, Hd, D->l2m_b, memory_out + col, ldm, B, ACT_NONE);
    g2.a_split = hs;
    SCV_TRY(launch_linear(g2, 0, s));
  } else {
    if (c.memory_bottleneck_dim > 0) {
      SCV_TRY(launch_linear(lin_args(z, c.latent_dim, D->l2m_a, t0, Hd, B, ACT_NONE), 0, s));
      SCV_TRY(launch_layernorm(t0, Hd, D->l2m_ln.g, D->l2m_ln.b, t0, Hd, B, Hd, ACT_GELU, nullptr, s));
    } else {
      SCV_TRY(launch_linear(lin_args(z, c.latent_dim, D->l2m_a, t0, Hd, B, ACT_GELU), 0, s));
    }
    SCV_TRY(launch_linear(lin_args(t0, Hd, D->l2m_b, memory_out + col, ldm, B, ACT_NONE), 0, s));
  }
  col += c.n_memory_tokens * d;
  // The second layers of the small conditioning branches ([B, d] x [d, 4 d] each) are 110 GFLOP apiece at 52.8 K rows
  // (3 ms each on the fp32 CUDA-core kernel): tensor cores for calls of >= cond_tc_min_rows rows (default 16384: the
  // encoder -> memory-token pipeline of BASELINE config 5).  Decode-sized calls keep the fp32 kernel (every product exact):
  // with the two-term activation split of the tensor path 1 of the 4096 greedy rows of config 2 left the oracle at a near-tie.
  auto lin_impl = [&](LinearArgs a) { a.wt = B >= tun().cond_tc_min_rows ? a.wt : nullptr; return launch_linear(a, 0, s); };
  if (use_skip) {                                                          // (:806-809)
    SCV_TRY(launch_linear(lin_args(skip, c.encoder_skip_dim, D->skip_a, t0, D->skip_a.N, B, ACT_GELU), 0, s));
    SCV_TRY(lin_impl(lin_args(t0, D->skip_a.N, D->skip_b, memory_out + col, ldm, B, ACT_NONE)));
    col += c.skip_n_tokens * d;
  }
  if (use_stoich) {                                                        // (:813-816)
    SCV_TRY(launch_linear(lin_args(stoich, c.stoich_input_dim, D->s2m_a, t0, d, B, ACT_NONE), 0, s));
    SCV_TRY(launch_layernorm(t0, d, D->s2m_ln.g, D->s2m_ln.b, t0, d, B, d, ACT_GELU, nullptr, s));
    SCV_TRY(lin_impl(lin_args(t0, d, D->s2m_b, memory_out + col, ldm, B, ACT_NONE)));
    col += c.n_stoich_tokens * d;
  }
  if (use_heads) {                                                         // (:858-868)
    SCV_TRY(launch_linear(lin_args(heads_in, c.heads_input_dim, D->h2m_a, t0, d / 2, B, ACT_NONE), 0, s));
    SCV_TRY(launch_layernorm(t0, d / 2, D->h2m_ln.g, D->h2m_ln.b, t0, d / 2, B, d / 2, ACT_GELU, nullptr, s));
    SCV_TRY(lin_impl(lin_args(t0, d / 2, D->h2m_b, t1, d, B, ACT_GELU)));
    SCV_TRY(lin_impl(lin_args(t1, d, D->h2m_c, memory_out + col, ldm, B, ACT_NONE)));
    col += c.heads_n_tokens * d;
  }
  return 0;
}

// The tensor-core step (activations handed from kernel to kernel as bf16 hi/lo SplitTiles) needs a batch that
// fills UMMA tiles and feature dims that are whole 64-wide k-blocks; smaller shapes use the CUDA-core kernels.
static bool use_tensor_cores(const scv_decoder_config& c, int B) {
  static const int forced = [] { const char* e = getenv("SCV_LINEAR_IMPL"); return e ? atoi(e) : 0; }();
  static const int min_rows = [] { const char* e = getenv("SCV_TC_MIN_ROWS"); return e ? atoi(e) : 1; }();
  return forced != 1 && B >= min_rows && c.d_model % 64 == 0 && c.dim_feedforward % 64 == 0 && (c.d_model / c.nhead) % 8 == 0 &&
         c.vocab_size % 4 == 0 && c.d_model <= 1024;
}

static int ensure_workspace(scv_decoder* D, int B, int M, bool need_site_dup = false, int Bm = 0) {
  if (Bm <= 0) Bm = B;                               // rows of the memory (fewer than B when RLOO samples share it)
  const scv_decoder_config& c = D->cfg;
  const size_t d = c.d_model, f = sizeof(float);
  const int pps = ceil_div(c.pe_len, kPagePos);
  SCV_TRY(D->x.ensure(B * d * f));
  SCV_TRY(D->xn.ensure(B * d * f));
  SCV_TRY(D->qkv.ensure(B * 3 * d * f));
  SCV_TRY(D->attn.ensure(B * d * f));
  SCV_TRY(D->q2.ensure(B * d * f));
  SCV_TRY(D->ff.ensure((size_t)B * c.dim_feedforward * f));
  SCV_TRY(D->h1.ensure(B * d * f));
  SCV_TRY(D->h2.ensure(B * d * f));
  SCV_TRY(D->t3.ensure(B * d * f));
  SCV_TRY(D->logits.ensure((size_t)B * c.vocab_size * f));
  SCV_TRY(D->tlog.ensure((size_t)B * 8 * f));
  SCV_TRY(D->slog.ensure((size_t)B * f));
  SCV_TRY(D->ckv.ensure((size_t)c.num_layers * Bm * M * 2 * d * f));
  SCV_TRY(D->kvpool.ensure((size_t)B * pps * c.num_layers * 2 * kPagePos * d * f));
  SCV_TRY(D->cur.ensure((size_t)B * sizeof(int)));
  SCV_TRY(D->fin.ensure((size_t)B));
  SCV_TRY(D->ptab.ensure((size_t)B * pps * sizeof(int)));
  SCV_TRY(D->state.ensure(sizeof(StepState)));
  if (use_tensor_cores(c, B)) {
    // zero-filled once: k-block padding columns (d/4 = 144 -> 192 for C576) must read as zeros forever
    for (DevBuf* b : {&D->xn_s, &D->attn_s, &D->h2_s}) {
      const size_t need = split_tile_bytes(B, c.d_model);
      if (need > b->cap) { SCV_TRY(b->ensure(need)); SCV_CUDA(cudaMemset(b->p, 0, need)); }
    }
    const size_t need = split_tile_bytes(B, c.dim_feedforward);
    if (need > D->ff_s.cap) { SCV_TRY(D->ff_s.ensure(need)); SCV_CUDA(cudaMemset(D->ff_s.p, 0, need)); }
  }
  const size_t io = (size_t)B * (c.pe_len - 1);
  SCV_TRY(D->o_tok.ensure(io * sizeof(long long)));
  SCV_TRY(D->o_lp.ensure(io * f));
  SCV_TRY(D->o_ent.ensure(io * f));
  SCV_TRY(D->forced_buf.ensure(io * sizeof(long long)));
  SCV_TRY(D->masks_buf.ensure((size_t)5 * c.vocab_size));
  SCV_TRY(D->row_map.ensure((size_t)B * sizeof(int)));
  if (need_site_dup) {      // captured steps bake these pointers in as well, so they belong to the signature below
    SCV_TRY(D->seen.ensure((size_t)B * c.vocab_size));
    SCV_TRY(D->dlog.ensure((size_t)B * sizeof(float)));
  }
  // captured steps hold raw workspace pointers: drop them whenever any buffer was reallocated
  size_t sig = 0;
  for (const DevBuf* b : {&D->x, &D->xn, &D->qkv, &D->attn, &D->q2, &D->ff, &D->h1, &D->h2, &D->t3, &D->logits, &D->tlog,
                          &D->slog, &D->ckv, &D->kvpool, &D->cur, &D->fin, &D->ptab, &D->state, &D->xn_s, &D->attn_s,
                          &D->ff_s, &D->h2_s, &D->o_tok, &D->o_lp, &D->o_ent, &D->masks_buf, &D->forced_buf, &D->seen,
                          &D->dlog, &D->sm_part, &D->row_map})
    sig = sig * 1000003u + reinterpret_cast<size_t>(b->p);
  if (sig != D->ws_signature || D->tune_epoch_seen != tune_epoch()) {
    D->drop_graphs(); D->ws_signature = sig; D->tune_epoch_seen = tune_epoch();
  }
  return 0;
}

// Phase list of the small-batch persistent step: the same operations, buffers and order as decode_rows() below.
static int build_small_phases(scv_decoder* D, const scv_generate_args* A, int B, int M) {
  const scv_decoder_config& c = D->cfg;
  const int d = c.d_model, hd = d / c.nhead, dff = c.dim_feedforward, pps = ceil_div(c.pe_len, kPagePos);
  D->small_active = false;
  static const int env = [] { const char* e = getenv("SCV_SMALL"); return e ? atoi(e) : 1; }();   // SCV_SMALL=0: per-projection path
  // up to 32 rows by default: above that a phase walks a second group of 32 rows and the step (1.21 ms at 64 rows) is no
  // faster than the per-projection path; SCV_SMALL_MAX_ROWS raises the limit up to kSmallMaxRows
  static const int max_rows = [] { const char* e = getenv("SCV_SMALL_MAX_ROWS"); return std::min(e ? atoi(e) : 32, kSmallMaxRows); }();
  if (!env || B > max_rows || prof_enabled() || (A->flags & SCV_FLAG_SYNC_EVERY_STEP)) return 0;
  if (D->sm_grid == 0) {
    int dev = 0, sms = 0;
    SCV_CUDA(cudaGetDevice(&dev));
    SCV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    D->sm_grid = sms;
  }
  if (D->sm_grid > 256) return 0;               // one barrier flag per CTA, 256 flags (decode_small.cu)
  // fixed sizes: a captured step holds these pointers, so they must never be reallocated
  SCV_TRY(D->sm_bar.ensure(256 * sizeof(unsigned)));      // one barrier flag per CTA (decode_small.cu kBarWords)
  SCV_TRY(D->sm_h2b.ensure((size_t)kSmallMaxRows * d * sizeof(float)));
  SCV_TRY(D->sm_t3s.ensure((size_t)kSmallMaxRows * d * sizeof(float)));
  SCV_TRY(D->sm_t3d.ensure((size_t)kSmallMaxRows * d * sizeof(float)));
  SCV_TRY(D->sm_phases.ensure((size_t)(8 * c.num_layers + 3) * sizeof(SmallPhase)));
  // SCV_SMALL_FUSE_FFN=1 (opt-in): fused feed-forward phases (decode_small.cu, kinds 2 and 3), one row group only.  797-802 us
  // per step against 826 us at 32 rows; the golden greedy decodes (masked, plain 63 steps) and 2048 / 2048 rows of the
  // exactness probe are token-identical, but the different fp32 summation order of linear2 (128 slabs) flips the sign of
  // a near-zero logit in the temperature = 0.0 golden (argmax of logits / 0 = first positive logit), so the
  // two projection phases stay the default.
  static const int fuse_env = [] { const char* e = getenv("SCV_SMALL_FUSE_FFN"); return e ? atoi(e) : 0; }();
  const bool fuse_ffn = fuse_env != 0 && B <= 32;
  if (fuse_ffn) SCV_TRY(D->sm_part.ensure((size_t)D->sm_grid * 32 * d * sizeof(float)));
  StepState* st = D->state.as<StepState>();
  float* x = D->x.as<float>(); float* qkv = D->qkv.as<float>(); float* attn = D->attn.as<float>();
  float* q2 = D->q2.as<float>(); float* ff = D->ff.as<float>(); float* h2 = D->h2.as<float>();
  float* t3 = D->t3.as<float>();
  std::vector<SmallPhase>& P = D->sm_host;
  P.clear();
  const int grid = D->sm_grid;
  auto op = [grid](const float* in, int ld_in, int K, const LNp* ln, const __nv_bfloat16* w, int ldw, const float* bias,
               int N, int act, const float* res, int ldr, float* out, int ldo) {
    SmallOp o;
    o.in = in; o.ld_in = ld_in; o.K = K; o.ln_g = ln ? ln->g : nullptr; o.ln_b = ln ? ln->b : nullptr;
    o.w = w; o.ldw = ldw; o.bias = bias; o.N = N; o.act = act; o.cpc = small_cols_per_cta(N, grid); o.res = res; o.ldr = ldr; o.out = out; o.ldo = ldo;
    return o;
  };
  auto lin = [&](const float* in, int ld_in, const LNp* ln, const Lin& L, int act, const float* res, float* out, int ldo) {
    return op(in, ld_in, L.K, ln, L.w, L.ldw, L.b, L.N, act, res, ldo, out, ldo);
  };
  auto gemv = [&](std::initializer_list<SmallOp> ops) {
    SmallPhase ph;
    ph.kind = 0; ph.nops = 0;
    for (const SmallOp& o : ops) ph.op[ph.nops++] = o;
    P.push_back(ph);
  };
  const float scale = (float)(1.0 / std::sqrt((double)hd));
  const long long page_stride = (long long)c.num_layers * 2 * kPagePos * d;
  for (int li = 0; li < c.num_layers; ++li) {
    const DecLayer& L = D->layers[li];
    gemv({op(x, d, d, &L.n1, L.sa_in_w, L.sa_in_ld, L.sa_in_b, 3 * d, ACT_NONE, nullptr, 0, qkv, 3 * d)});
    SmallPhase sa;
    sa.kind = 1; sa.nops = 0;
    sa.attn.q = qkv; sa.attn.ldq = 3 * d; sa.attn.knew = qkv + d; sa.attn.vnew = qkv + 2 * d; sa.attn.ldn = 3 * d;
    sa.attn.kcache = D->kvpool.as<float>() + (size_t)(li * 2 + 0) * kPagePos * d;
    sa.attn.vcache = D->kvpool.as<float>() + (size_t)(li * 2 + 1) * kPagePos * d;
    sa.attn.page_table = D->ptab.as<int>(); sa.attn.pages_per_seq = pps; sa.attn.page_stride = page_stride;
    sa.attn.row_stride = d; sa.attn.out = attn; sa.attn.ldo = d; sa.attn.B = B; sa.attn.nhead = c.nhead; sa.attn.hd = hd;
    sa.attn.scale = scale; sa.attn.fixed_len = -1; sa.attn.max_n = std::max(c.pe_len, M); sa.attn.st = st;
    P.push_back(sa);
    gemv({lin(attn, d, nullptr, L.sa_out, ACT_NONE, x, x, d)});
    gemv({op(x, d, d, &L.n2, L.ca_in_w, L.ca_in_ld, L.ca_in_b, d, ACT_NONE, nullptr, 0, q2, d)});
    SmallPhase ca;
    ca.kind = 1; ca.nops = 0;
    float* ckv = D->ckv.as<float>() + (size_t)li * B * M * 2 * d;
    ca.attn.q = q2; ca.attn.ldq = d; ca.attn.kcache = ckv; ca.attn.vcache = ckv + d; ca.attn.seq_stride = (long long)M * 2 * d;
    ca.attn.row_stride = 2 * d; ca.attn.out = attn; ca.attn.ldo = d; ca.attn.B = B; ca.attn.nhead = c.nhead; ca.attn.hd = hd;
    ca.attn.scale = scale; ca.attn.fixed_len = M; ca.attn.max_n = std::max(c.pe_len, M); ca.attn.st = st;
    P.push_back(ca);
    gemv({lin(attn, d, nullptr, L.ca_out, ACT_NONE, x, x, d)});
    if (fuse_ffn) {
      // feed-forward block as two phases without the [B, dff] round trip: linear1 + GELU on 16 hidden units per CTA and
      // linear2's partial sums over them (one slab per CTA), then the sum of the slabs + bias + residual
      SmallPhase f;
      f.kind = 2; f.nops = 2;
      f.op[0] = lin(x, d, &L.n3, L.ff1, ACT_GELU, nullptr, ff, dff);
      f.op[0].cpc = 16;
      f.op[1] = lin(ff, dff, nullptr, L.ff2, ACT_NONE, nullptr, D->sm_part.as<float>(), d);
      P.push_back(f);
      SmallPhase r;
      r.kind = 3; r.nops = 1;
      r.op[0] = op(D->sm_part.as<float>(), d, ceil_div(dff, 16), nullptr, nullptr, 0, L.ff2.b, d, ACT_NONE, x, d, x, d);
      r.op[0].cpc = 4;
      P.push_back(r);
    } else {
      gemv({lin(x, d, &L.n3, L.ff1, ACT_GELU, nullptr, ff, dff)});
      gemv({lin(ff, dff, nullptr, L.ff2, ACT_NONE, x, x, d)});
    }
  }
  const bool type = A->type_masks != nullptr, stop = A->stop_boost > 0.f, dup = A->site_dup_threshold > 0.f;
  float* h2b = D->sm_h2b.as<float>(); float* t3s = D->sm_t3s.as<float>(); float* t3d = D->sm_t3d.as<float>();
  {
    SmallPhase ph;
    ph.kind = 0; ph.nops = 0;
    ph.op[ph.nops++] = lin(x, d, &D->out_ln, D->out_a, ACT_GELU, nullptr, h2, d);
    if (type) ph.op[ph.nops++] = lin(x, d, &D->tt_ln, D->tt_a, ACT_GELU, nullptr, h2b, d);
    if (stop) ph.op[ph.nops++] = lin(x, d, nullptr, D->stop_a, ACT_GELU, nullptr, t3s, d / 4);
    if (dup) ph.op[ph.nops++] = lin(x, d, nullptr, D->dup_a, ACT_GELU, nullptr, t3d, d / 4);
    P.push_back(ph);
  }
  {
    SmallPhase ph;
    ph.kind = 0; ph.nops = 0;
    ph.op[ph.nops++] = lin(h2, d, nullptr, D->out_b, ACT_NONE, nullptr, D->logits.as<float>(), c.vocab_size);
    if (type) ph.op[ph.nops++] = lin(h2b, d, nullptr, D->tt_b, ACT_GELU, nullptr, t3, d / 4);
    if (stop) ph.op[ph.nops++] = lin(t3s, d / 4, nullptr, D->stop_b, ACT_NONE, nullptr, D->slog.as<float>(), 1);
    if (dup) ph.op[ph.nops++] = lin(t3d, d / 4, nullptr, D->dup_b, ACT_NONE, nullptr, D->dlog.as<float>(), 1);
    P.push_back(ph);
  }
  if (type) gemv({lin(t3, d / 4, nullptr, D->tt_c, ACT_NONE, nullptr, D->tlog.as<float>(), 8)});
  for (const SmallPhase& ph : P)
    if (!small_phase_fits(ph, D->sm_grid)) return 0;          // some shape does not fit: the per-projection path decodes
  SCV_REQUIRE(P.size() <= (size_t)(8 * c.num_layers + 3), "small-batch step: phase list overflow");
  D->sm_n_phases = (int)P.size();
  D->small_active = true;
  return 0;
}

// Program of the cluster-parallel small-batch step (decode_cluster.cu): the same operations as decode_rows() below, every
// projection split by output columns over the 8 CTAs of a cluster, activations exchanged through distributed shared memory.
static int build_cluster_program(scv_decoder* D, const scv_generate_args* A, int B, int M) {
  const scv_decoder_config& c = D->cfg;
  D->cl_active = false;
  const int d = c.d_model, hd = d / c.nhead, dff = c.dim_feedforward, V = c.vocab_size, pps = ceil_div(c.pe_len, kPagePos);
  if (A->memory_rows > 0 || tun().cluster == 0 || B > tun().cluster_max_rows || prof_enabled() || (A->flags & SCV_FLAG_SYNC_EVERY_STEP)) return 0;
  if (!cluster_shape_ok(d, c.nhead, dff, V, c.pe_len, M)) return 0;
  // Clusters that can be resident at once (8 CTAs of ~200 KB each inside one GPC: 15-16 on a B200); all of a decode's
  // clusters should run together (a late cluster is still decoded correctly, just later), so R is the smallest rows-per-
  // cluster count that fits the batch into them.
  static int max_clusters = 0;
  if (max_clusters == 0) max_clusters = std::max(1, cluster_max_active(4, 215 * 1024));
  int R = tun().cluster_rows;
  if (R != 1 && R != 2 && R != 4) R = B <= max_clusters ? 1 : (B <= 2 * max_clusters ? 2 : 4);
  if (ceil_div(B, R) > max_clusters) return 0;
  ClProgram& P = D->cl_prog;
  P = ClProgram{};
  P.B = B; P.d = d; P.nhead = c.nhead; P.hd = hd; P.dff = dff; P.V = V; P.pe_len = c.pe_len; P.rows_per_cluster = R;
  {
    auto seg = [](int K) { return K <= 768 ? K : K / 4; };               // KSEG_MAX of decode_cluster.cu
    P.cpl = std::max(seg(d), std::max(seg(dff), seg(d / 4))) <= 512 ? 2 : 3;
  }
  const int hpc = c.nhead / kClSize;
  const int widths[CB_COUNT] = {d, d, d, d, d, dff, d / 4, d / 4, d / 4, hd * hpc, hd * hpc, hd * hpc};
  int off = 0;
  for (int b = 0; b < CB_COUNT; ++b) { P.buf_off[b] = off; P.buf_ld[b] = round_up(widths[b], 4); off += R * P.buf_ld[b]; }
  P.buf_floats = off;
  P.page_table = D->ptab.as<int>(); P.pages_per_seq = pps;
  P.scale = (float)(1.0 / std::sqrt((double)hd));
  P.x_global = D->x.as<float>();
  P.st = D->state.as<StepState>();
  std::vector<ClInstr> prog;
  auto blank = [](int kind) { ClInstr I; memset(&I, 0, sizeof(I)); I.kind = kind; return I; };
  auto ln = [&](const LNp& p, int src, int dst) {
    ClInstr I = blank(CL_LN); I.ln.gamma = p.g; I.ln.beta = p.b; I.ln.src_buf = src; I.ln.dst_buf = dst; I.ln.n = d; prog.push_back(I);
  };
  auto gemv = [&](const __nv_bfloat16* w, int ldw, int K, int N, const float* bias, int act, int in_buf, bool residual, int out_kind,
                  int out_buf, float* out_global, int out_ld, bool split) {
    ClInstr I = blank(CL_GEMV); I.act = act; I.g.w = w; I.g.ldw = ldw; I.g.K = K; I.g.N = N; I.g.split = split ? 1 : 0; I.g.bias = bias;
    I.g.in_buf = (short)in_buf; I.g.residual = residual ? 1 : 0; I.g.out_kind = (short)out_kind; I.g.out_buf = (short)out_buf;
    I.g.out_global = out_global; I.g.out_ld = out_ld;
    prog.push_back(I);
  };
  auto lin = [&](const Lin& L, int act, int in_buf, bool residual, int out_kind, int out_buf) {
    gemv(L.w, L.ldw, L.K, L.N, L.b, act, in_buf, residual, out_kind, out_buf, nullptr, 0, true);
  };
  auto sync = [&] { prog.push_back(blank(CL_SYNC)); };
  const long long page_stride = (long long)c.num_layers * 2 * kPagePos * d;
  for (int li = 0; li < c.num_layers; ++li) {
    const DecLayer& L = D->layers[li];
    ln(L.n1, CB_X, CB_XN);                                                                   // self attention (:1244-1296)
    for (int part = 0; part < 3; ++part)
      gemv(L.sa_in_w + (size_t)part * d * L.sa_in_ld, L.sa_in_ld, d, d, L.sa_in_b + part * d, ACT_NONE, CB_XN, false, CO_LOCAL,
           CB_Q + part, nullptr, 0, true);
    { ClInstr I = blank(CL_ATTN_SELF); I.at.kcache = D->kvpool.as<float>() + (size_t)(li * 2 + 0) * kPagePos * d;
      I.at.vcache = D->kvpool.as<float>() + (size_t)(li * 2 + 1) * kPagePos * d; I.at.page_stride = page_stride; I.at.row_stride = d;
      I.at.fixed_len = -1; prog.push_back(I); }
    sync();
    lin(L.sa_out, ACT_NONE, CB_A, true, CO_GATHER, CB_X);
    sync();
    ln(L.n2, CB_X, CB_XN);                                                                   // cross attention (:1299-1308)
    gemv(L.ca_in_w, L.ca_in_ld, d, d, L.ca_in_b, ACT_NONE, CB_XN, false, CO_LOCAL, CB_Q, nullptr, 0, true);
    { ClInstr I = blank(CL_ATTN_CROSS); float* ckv = D->ckv.as<float>() + (size_t)li * B * M * 2 * d;
      I.at.kcache = ckv; I.at.vcache = ckv + d; I.at.seq_stride = (long long)M * 2 * d; I.at.row_stride = 2 * d; I.at.fixed_len = M;
      prog.push_back(I); }
    sync();
    lin(L.ca_out, ACT_NONE, CB_A, true, CO_GATHER, CB_X);
    sync();
    ln(L.n3, CB_X, CB_XN);                                                                   // feed forward (:1311-1313)
    lin(L.ff1, ACT_GELU, CB_XN, false, CO_GATHER, CB_H);
    sync();
    lin(L.ff2, ACT_NONE, CB_H, true, CO_GATHER, CB_X);
    sync();
  }
  const bool type = A->type_masks != nullptr, stop = A->stop_boost > 0.f, dup = A->site_dup_threshold > 0.f;
  ln(D->out_ln, CB_X, CB_XN);                                                                // heads (:1413, 1417, 1439)
  lin(D->out_a, ACT_GELU, CB_XN, false, CO_GATHER, CB_A);
  if (type) { ln(D->tt_ln, CB_X, CB_XN2); lin(D->tt_a, ACT_GELU, CB_XN2, false, CO_GATHER, CB_B); }
  if (stop) lin(D->stop_a, ACT_GELU, CB_X, false, CO_GATHER, CB_C);
  if (dup) lin(D->dup_a, ACT_GELU, CB_X, false, CO_GATHER, CB_D);
  sync();
  gemv(D->out_b.w, D->out_b.ldw, d, V, D->out_b.b, ACT_NONE, CB_A, false, CO_GLOBAL, 0, D->logits.as<float>(), V, true);
  if (type) lin(D->tt_b, ACT_GELU, CB_B, false, CO_GATHER, CB_E);
  if (stop) gemv(D->stop_b.w, D->stop_b.ldw, d / 4, 1, D->stop_b.b, ACT_NONE, CB_C, false, CO_GLOBAL, 0, D->slog.as<float>(), 1, false);
  if (dup) gemv(D->dup_b.w, D->dup_b.ldw, d / 4, 1, D->dup_b.b, ACT_NONE, CB_D, false, CO_GLOBAL, 0, D->dlog.as<float>(), 1, false);
  if (type) {
    sync();
    gemv(D->tt_c.w, D->tt_c.ldw, d / 4, 5, D->tt_c.b, ACT_NONE, CB_E, false, CO_GLOBAL, 0, D->tlog.as<float>(), 8, false);
  }
  P.n_instr = (int)prog.size();
  if (prog.size() > 320 || cluster_smem_bytes(P) > 227 * 1024) return 0;      // MAX_INSTR of decode_cluster.cu
  SCV_TRY(D->cl_instr.ensure(1024 * sizeof(ClInstr)));      // fixed size: a captured step holds this pointer
  // (pageable host memory: the copy is staged by the driver before the call returns)
  SCV_CUDA(cudaMemcpyAsync(D->cl_instr.p, prog.data(), prog.size() * sizeof(ClInstr), cudaMemcpyHostToDevice, D->main));
  P.instr = D->cl_instr.as<ClInstr>(); P.n_instr = (int)prog.size();
  static const int exp_env = [] { const char* e = getenv("SCV_CLUSTER_EXP"); return e ? atoi(e) : 0; }();
  P.exp = exp_env;
  static const int dbg_env = [] { const char* e = getenv("SCV_CLUSTER_DEBUG"); return e ? atoi(e) : 0; }();
  if (dbg_env) {
    SCV_TRY(D->cl_dbg.ensure(1024 * sizeof(unsigned long long)));
    SCV_CUDA(cudaMemsetAsync(D->cl_dbg.p, 0, 1024 * sizeof(unsigned long long), D->main));
    P.dbg = D->cl_dbg.as<unsigned long long>();
    D->cl_kinds.assign(prog.size(), 0);
    for (size_t i = 0; i < prog.size(); ++i) D->cl_kinds[i] = prog[i].kind * 10000 + (prog[i].kind == CL_GEMV ? prog[i].g.K / 8 : 0);
  }
  D->cl_active = true;
  return 0;
}

// One decode step for rows [r0, r0 + B) of the call's batch (a sub-batch; every buffer is row-major by batch row and
// the SplitTile buffers are tiled by 128 rows, so a sub-batch is a pointer offset).  phase 1 = everything up to and
// including the first sampler kernel, phase 2 = the second sampler kernel (sampling / entropy only).
static int decode_rows(scv_decoder* D, const scv_generate_args* A, int steps_max, int host_step, int r0, int B,
                       int phase, cudaStream_t s) {
  const scv_decoder_config& c = D->cfg;
  const int Bfull = A->batch, M = A->n_memory, d = c.d_model, hd = d / c.nhead, dff = c.dim_feedforward;
  StepState* st = D->state.as<StepState>();
  const int* done = &st->done;
  const int pps = ceil_div(c.pe_len, kPagePos);
  const float scale = (float)(1.0 / std::sqrt((double)hd));
  float* x = D->x.as<float>() + (size_t)r0 * d; float* xn = D->xn.as<float>() + (size_t)r0 * d;
  float* qkv = D->qkv.as<float>() + (size_t)r0 * 3 * d; float* attn = D->attn.as<float>() + (size_t)r0 * d;
  float* q2 = D->q2.as<float>() + (size_t)r0 * d; float* ff = D->ff.as<float>() + (size_t)r0 * dff;
  int* page_table = D->ptab.as<int>() + (size_t)(((A->flags & SCV_FLAG_COMPACT_FINISHED) != 0 && !D->small_active && !D->cl_active) ? 0 : r0) * pps;
  // Tensor-core step: every projection input is handed over as a bf16 hi/lo SplitTile written by its producer
  // (LayerNorm, attention, previous GEMM epilogue); nullptr selects the fp32 CUDA-core path.
  const bool tc = use_tensor_cores(c, Bfull);
  // Opt-in retirement of finished rows (SCV_FLAG_COMPACT_FINISHED, per-projection path only): per-step buffers are indexed
  // by SLOT, everything that lives for the whole call (KV pages, projected memory, outputs, flags) by row_map[slot].
  const bool compact = (A->flags & SCV_FLAG_COMPACT_FINISHED) != 0 && !D->small_active && !D->cl_active;
  const int* row_map = compact ? D->row_map.as<int>() : nullptr;
  const int ob = compact ? 0 : r0;            // offset of row-indexed arrays: none when rows are looked up through row_map
  auto tile_off = [&](const DevBuf& b, int K) -> void* {
    return static_cast<unsigned char*>(b.p) + (size_t)(r0 / 128) * ceil_div(K, 64) * 32768;
  };
  // each tensor-core projection prefetches the tiled weights of the projection that follows it (common.cuh next_w)
  auto next = [&](LinearArgs& l, const __nv_bfloat16* wt, int N, int K) {
    l.row_base = r0;
    if (tc && wt != nullptr) { l.next_w = wt; l.next_w_bytes = tc_packed_elems(N, K) * sizeof(__nv_bfloat16); }
  };
  void* xn_s = tc ? tile_off(D->xn_s, d) : nullptr; void* attn_s = tc ? tile_off(D->attn_s, d) : nullptr;
  void* ff_s = tc ? tile_off(D->ff_s, dff) : nullptr; void* h2_s = tc ? tile_off(D->h2_s, d) : nullptr;
  SamplerArgs sp;
  sp.logits = D->logits.as<float>() + (size_t)r0 * c.vocab_size; sp.ldl = c.vocab_size;
  sp.type_logits = D->tlog.as<float>() + (size_t)r0 * 8; sp.ldt = 8; sp.stop_logits = D->slog.as<float>() + r0;
  sp.type_masks = A->type_masks; sp.B = B; sp.V = c.vocab_size; sp.max_len = steps_max + 1;
  sp.temperature = A->temperature; sp.top_k = A->top_k; sp.top_p = A->top_p;
  sp.stop_boost = A->stop_boost; sp.hard_stop = A->hard_stop_threshold;
  if (A->site_dup_threshold > 0.f) {
    sp.dup_logits = D->dlog.as<float>() + r0; sp.dup_threshold = A->site_dup_threshold;
    sp.seen = D->seen.as<unsigned char>() + (size_t)ob * c.vocab_size;
  }
  sp.want_logprobs = A->want_log_probs; sp.want_entropy = A->want_entropy; sp.flags = A->flags;
  sp.row_base = r0; sp.row_map = row_map; sp.slot_base = r0;
  sp.out_tokens = reinterpret_cast<long long*>(A->out_tokens) + (size_t)ob * steps_max;
  sp.out_logprobs = A->want_log_probs ? A->out_log_probs + (size_t)ob * steps_max : nullptr;
  sp.out_entropy = A->want_entropy ? A->out_entropy + (size_t)ob * steps_max : nullptr;
  sp.out_ld = steps_max; sp.cur_tokens = D->cur.as<int>() + r0; sp.finished = D->fin.as<unsigned char>() + ob;
  sp.forced = A->forced_tokens ? reinterpret_cast<const long long*>(A->forced_tokens) + (size_t)ob * steps_max : nullptr;
  sp.st = st;
  if (phase == 2) return launch_sampler(sp, 2, s);
  auto norm = [&](const LNp& P, float* fp32_out) -> int {
    return tc ? launch_layernorm_split(x, d, P.g, P.b, xn_s, B, d, 1, done, s)
              : launch_layernorm(x, d, P.g, P.b, fp32_out, d, B, d, ACT_NONE, done, s);
  };

  // A residual projection followed by LayerNorm P of the row it wrote: one cluster kernel when the shape allows
  // (gemm_tcgen05_ln.cu), else the projection and then the LayerNorm kernel.
  auto residual_then_norm = [&](LinearArgs& l, const LNp& P, float* fp32_out) -> int {
    if (tc) { l.ln_gamma = P.g; l.ln_beta = P.b; l.ln_out_split = xn_s; }
    if (tc && tc_res_ln_ok(l)) return launch_linear_res_ln(l, s);
    l.ln_out_split = nullptr;
    SCV_TRY(launch_linear(l, 0, s));
    return norm(P, fp32_out);
  };

  EmbedArgs e;
  e.table = D->emb; e.ld_table = D->ld_emb; e.pe = D->pe; e.d = d; e.cur_tokens = D->cur.as<int>() + r0; e.x = x; e.B = B;
  e.page_table = page_table; e.pages_per_seq = pps; e.st = st; e.row_map = row_map; e.slot_base = r0;
  SCV_TRY(launch_embed(e, s));
  if (D->cl_active) {                      // cluster-parallel small-batch decode (decode_cluster.cu)
    if (phase == 3) {                      // the whole plain-greedy decode in one launch
      SCV_REQUIRE(sampler_plain_greedy(sp), "persistent decode: the call is not plain greedy");
      SmallTail t;
      t.sp = sp; t.emb = D->emb; t.ld_emb = D->ld_emb; t.pe = D->pe; t.d = d; t.x = x; t.page_table = page_table; t.pages_per_seq = pps;
      t.max_steps = steps_max;
      return launch_decode_cluster_persist(D->cl_prog, t, s);
    }
    SCV_TRY(launch_decode_cluster_step(D->cl_prog, s));
    return launch_sampler(sp, 1, s);
  }
  if (D->small_active && phase == 3) {     // the whole decode in one launch (plain greedy): step loop inside the kernel
    SCV_REQUIRE(sampler_plain_greedy(sp), "persistent decode: the call is not plain greedy");
    SCV_CUDA(cudaMemsetAsync(D->sm_bar.p, 0, 256 * sizeof(unsigned), s));
    SmallTail t;
    t.sp = sp; t.emb = D->emb; t.ld_emb = D->ld_emb; t.pe = D->pe; t.d = d; t.x = x; t.page_table = page_table; t.pages_per_seq = pps;
    t.max_steps = steps_max;
    return launch_decode_small_persist(D->sm_phases.as<SmallPhase>(), D->sm_n_phases, B, D->sm_bar.as<unsigned>(), D->sm_grid, t, s);
  }
  if (D->small_active) {     // layers and heads in one persistent kernel (decode_small.cu); its barrier counter starts at 0
    SCV_CUDA(cudaMemsetAsync(D->sm_bar.p, 0, 256 * sizeof(unsigned), s));
    SCV_TRY(launch_decode_small(D->sm_phases.as<SmallPhase>(), D->sm_n_phases, B, st, D->sm_bar.as<unsigned>(), D->sm_grid, s));
    return launch_sampler(sp, 1, s);
  }

  const long long page_stride = (long long)c.num_layers * 2 * kPagePos * d;
  float* h1_ = D->h1.as<float>() + (size_t)r0 * d;     // fp32 output of the head LayerNorm (CUDA-core path)
  for (int li = 0; li < c.num_layers; ++li) {
    const DecLayer& L = D->layers[li];
    // ---- self attention (:1244-1296); norm1 of layers >= 1 was produced by the previous layer's linear2 kernel
    if (li == 0) SCV_TRY(norm(L.n1, xn));
    LinearArgs a;
    a.x = xn; a.ldx = d; a.a_split = xn_s; a.w = L.sa_in_w; a.ldw = L.sa_in_ld; a.wt = L.sa_in_wt; a.bias = L.sa_in_b; a.y = qkv; a.ldy = 3 * d;
    a.M = B; a.N = 3 * d; a.K = d; a.done_flag = done;
    next(a, L.sa_out.wt, d, d);
    SCV_TRY(launch_linear(a, 0, s));
    AttnArgs sa;
    sa.q = qkv; sa.ldq = 3 * d; sa.knew = qkv + d; sa.vnew = qkv + 2 * d; sa.ldn = 3 * d;
    sa.kcache = D->kvpool.as<float>() + (size_t)(li * 2 + 0) * kPagePos * d;
    sa.vcache = D->kvpool.as<float>() + (size_t)(li * 2 + 1) * kPagePos * d;
    sa.page_stride = page_stride; sa.row_stride = d;
    sa.page_table = page_table; sa.pages_per_seq = pps;
    sa.out = attn; sa.ldo = d; sa.B = B; sa.nhead = c.nhead; sa.hd = hd; sa.scale = scale; sa.fixed_len = -1;
    sa.max_n = std::max(c.pe_len, M); sa.st = st; sa.host_len_hint = host_step + 1;
    sa.out_split = static_cast<unsigned char*>(attn_s); sa.kb_out = d / 64;
    sa.row_map = row_map; sa.slot_base = r0;
    SCV_TRY(launch_attention(sa, s));
    LinearArgs o = lin_args(attn, d, L.sa_out, x, d, B, ACT_NONE, done);
    o.residual = x; o.ldr = d; o.a_split = attn_s;
    next(o, L.ca_q_wt, d, d);
    // ---- cross attention to the memory tokens (:1299-1308)
    SCV_TRY(residual_then_norm(o, L.n2, xn));
    LinearArgs q;
    q.x = xn; q.ldx = d; q.a_split = xn_s; q.w = L.ca_in_w; q.ldw = L.ca_in_ld; q.wt = L.ca_q_wt; q.bias = L.ca_in_b; q.y = q2; q.ldy = d;
    q.M = B; q.N = d; q.K = d; q.done_flag = done;
    next(q, L.ca_out.wt, d, d);
    SCV_TRY(launch_linear(q, 0, s));
    AttnArgs ca;
    ca.q = q2; ca.ldq = d;
    const int Bm = A->memory_rows > 0 ? A->memory_rows : Bfull;          // rows of the projected memory (RLOO samples share it)
    float* ckv = D->ckv.as<float>() + ((size_t)li * Bm + (A->memory_rows > 0 ? 0 : ob)) * M * 2 * d;
    ca.kcache = ckv; ca.vcache = ckv + d; ca.seq_stride = (long long)M * 2 * d; ca.row_stride = 2 * d;
    ca.out = attn; ca.ldo = d; ca.B = B; ca.nhead = c.nhead; ca.hd = hd; ca.scale = scale;
    ca.fixed_len = M; ca.max_n = std::max(c.pe_len, M); ca.st = st;
    ca.out_split = static_cast<unsigned char*>(attn_s); ca.kb_out = d / 64;
    ca.row_map = row_map; ca.slot_base = r0; ca.seq_mod = A->memory_rows > 0 ? A->memory_rows : 0;
    SCV_TRY(launch_attention(ca, s));
    LinearArgs co = lin_args(attn, d, L.ca_out, x, d, B, ACT_NONE, done);
    co.residual = x; co.ldr = d; co.a_split = attn_s;
    next(co, L.ff1.wt, dff, d);
    // ---- feed forward (:1311-1313)
    SCV_TRY(residual_then_norm(co, L.n3, xn));
    LinearArgs f1 = lin_args(xn, d, L.ff1, ff, dff, B, ACT_GELU, done);
    f1.a_split = xn_s; f1.y_split = ff_s;
    next(f1, L.ff2.wt, d, dff);
    SCV_TRY(launch_linear(f1, 0, s));
    LinearArgs f2 = lin_args(ff, dff, L.ff2, x, d, B, ACT_NONE, done);
    f2.residual = x; f2.ldr = d; f2.a_split = ff_s;
    if (li + 1 < c.num_layers) next(f2, D->layers[li + 1].sa_in_wt, 3 * d, d); else next(f2, D->out_a.wt, d, d);
    SCV_TRY(residual_then_norm(f2, li + 1 < c.num_layers ? D->layers[li + 1].n1 : D->out_ln, li + 1 < c.num_layers ? xn : h1_));
  }
  // ---- heads (:1413, 1417, 1439)
  float* h1 = D->h1.as<float>() + (size_t)r0 * d; float* h2 = D->h2.as<float>() + (size_t)r0 * d;
  float* t3 = D->t3.as<float>() + (size_t)r0 * (d / 4);
  float* logits = D->logits.as<float>() + (size_t)r0 * c.vocab_size;
  float* tlog = D->tlog.as<float>() + (size_t)r0 * 8; float* slog = D->slog.as<float>() + r0;
  LinearArgs oa = lin_args(h1, d, D->out_a, h2, d, B, ACT_GELU, done);
  oa.a_split = xn_s; oa.y_split = h2_s; oa.row_base = r0;
  next(oa, D->out_b.wt, c.vocab_size, d);
  SCV_TRY(launch_linear(oa, 0, s));
  // Sampling modes run two sampler kernels because the reference's uniform fallback is BATCH-global (H2, :1464-1466): the
  // first only finds out whether any adjusted logit of the batch is NaN / +-inf.  Without a type mask and without the hard
  // stop (the two adjustments that write -inf) that is the case exactly when a raw logit is non-finite or a stop logit is
  // NaN, which the kernels producing them report on the fly - so the first sampler kernel (a pass over all B x V logits)
  // is not launched.  With the row-local fallback (flags bit 0 clear) its result is never read at all.
  const bool two_phase_here = !(A->temperature < 0.01f) || A->want_entropy;
  const bool skip_phase1 = two_phase_here && (((A->flags & 1u) == 0) ||
                                              (A->type_masks == nullptr && !(A->stop_boost > 0.f && A->hard_stop_threshold > 0.f)));
  LinearArgs ob_ = lin_args(h2, d, D->out_b, logits, c.vocab_size, B, ACT_NONE, done);
  ob_.a_split = h2_s; ob_.row_base = r0;
  if (skip_phase1 && (A->flags & 1u) != 0) { ob_.nonfinite_flag = &st->degenerate; ob_.nonfinite_mode = 1; }
  SCV_TRY(launch_linear(ob_, 0, s));
  if (A->type_masks != nullptr) {
    SCV_TRY(norm(D->tt_ln, h1));
    LinearArgs ta = lin_args(h1, d, D->tt_a, h2, d, B, ACT_GELU, done);
    ta.a_split = xn_s; ta.y_split = h2_s; ta.row_base = r0;
    SCV_TRY(launch_linear(ta, 0, s));
    LinearArgs tb = lin_args(h2, d, D->tt_b, t3, d / 4, B, ACT_GELU, done);
    tb.a_split = h2_s; tb.row_base = r0;
    SCV_TRY(launch_linear(tb, 0, s));
    SCV_TRY(launch_linear(lin_args(t3, d / 4, D->tt_c, tlog, 8, B, ACT_NONE, done), 0, s));
  }
  if (A->stop_boost > 0.f) {
    SCV_TRY(launch_linear(lin_args(x, d, D->stop_a, t3, d / 4, B, ACT_GELU, done), 0, s));
    LinearArgs sb_ = lin_args(t3, d / 4, D->stop_b, slog, 1, B, ACT_NONE, done);
    if (skip_phase1 && (A->flags & 1u) != 0) { sb_.nonfinite_flag = &st->degenerate; sb_.nonfinite_mode = 2; }   // sigmoid(+-inf) is finite
    SCV_TRY(launch_linear(sb_, 0, s));
  }
  if (A->site_dup_threshold > 0.f) {      // site_dup_head (:1427); position 0 never gates but the launch list stays fixed
    SCV_TRY(launch_linear(lin_args(x, d, D->dup_a, t3, d / 4, B, ACT_GELU, done), 0, s));
    SCV_TRY(launch_linear(lin_args(t3, d / 4, D->dup_b, D->dlog.as<float>() + r0, 1, B, ACT_NONE, done), 0, s));
  }
  if (skip_phase1) return 0;
  return launch_sampler(sp, 1, s);
}

int scv_decoder_generate(scv_decoder* D, const scv_generate_args* A, void* stream) {
  SCV_REQUIRE(D && A, "generate: null argument");
  const scv_generate_args* A_user = A;
  SCV_REQUIRE(A->batch > 0 && A->memory && A->out_tokens && A->out_steps, "generate: bad arguments");
  SCV_REQUIRE(A->n_memory > 0, "generate: n_memory must be positive");
  SCV_REQUIRE(!(A->site_dup_threshold > 0.f) || (D->ws.loaded("site_dup_head.0.weight") && D->ws.loaded("site_dup_head.2.weight")),
              "generate: site_dup_threshold > 0 needs the site_dup_head weights, which this checkpoint did not provide");
  SCV_REQUIRE(!A->want_log_probs || A->out_log_probs, "generate: out_log_probs is NULL");
  SCV_REQUIRE(!A->want_entropy || A->out_entropy, "generate: out_entropy is NULL");
  SCV_REQUIRE(scv_decoder_missing_weights(D) == 0, "generate: weights missing");
  // The decode runs on an engine-owned stream (the caller's may be the legacy default stream, which cannot be
  // captured into a CUDA graph); it is ordered after the caller's stream here and synchronised before returning.
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  SCV_TRY(D->ensure_streams());
  cudaStream_t s = D->main;
  SCV_CUDA(cudaEventRecord(D->ev_fork, caller));
  SCV_CUDA(cudaStreamWaitEvent(s, D->ev_fork, 0));
  const scv_decoder_config& c = D->cfg;
  const int max_len = std::min(A->max_len, c.pe_len);        // silent clamp (:1372-1375)
  const int steps_max = max_len - 1;
  SCV_REQUIRE(steps_max >= 1, "generate: max_len %d leaves no step to run", A->max_len);
  const int B = A->batch, M = A->n_memory, d = c.d_model;
  set_pdl_for_call(true);
  const int Bm = A->memory_rows > 0 ? A->memory_rows : B;
  SCV_REQUIRE(A->memory_rows == 0 || (B % A->memory_rows == 0 && B > 64 && B > A->memory_rows),
              "generate: memory_rows %d needs batch %d to be a larger multiple of it (and > 64 rows)", A->memory_rows, B);
  SCV_TRY(ensure_workspace(D, B, M, A->site_dup_threshold > 0.f, Bm));
  StepState* st = D->state.as<StepState>();
  const bool want_compact = (A->flags & SCV_FLAG_COMPACT_FINISHED) != 0;
  SCV_TRY(launch_init_rows(D->cur.as<int>(), D->fin.as<unsigned char>(), B, st, A->seed, A->offset, s,
                           want_compact ? D->row_map.as<int>() : nullptr));
  // the step kernels read / write engine-owned buffers only (so a captured step can be replayed for any call);
  // caller tensors are copied in here and out after the loop
  const size_t io = (size_t)B * steps_max;
  scv_generate_args G = *A;
  G.out_tokens = D->o_tok.as<int64_t>(); G.out_log_probs = D->o_lp.as<float>(); G.out_entropy = D->o_ent.as<float>();
  if (A->type_masks != nullptr) {
    SCV_CUDA(cudaMemcpyAsync(D->masks_buf.p, A->type_masks, (size_t)5 * c.vocab_size, cudaMemcpyDeviceToDevice, s));
    G.type_masks = D->masks_buf.as<uint8_t>();
  }
  if (A->forced_tokens != nullptr) {
    SCV_CUDA(cudaMemcpyAsync(D->forced_buf.p, A->forced_tokens, io * sizeof(long long), cudaMemcpyDeviceToDevice, s));
    G.forced_tokens = D->forced_buf.as<int64_t>();
  }
  if (A->site_dup_threshold > 0.f) SCV_CUDA(cudaMemsetAsync(D->seen.p, 0, (size_t)B * c.vocab_size, s));
  SCV_CUDA(cudaMemsetAsync(D->o_tok.p, 0, io * sizeof(long long), s));
  if (A->want_log_probs) SCV_CUDA(cudaMemsetAsync(D->o_lp.p, 0, io * sizeof(float), s));
  if (A->want_entropy) SCV_CUDA(cudaMemsetAsync(D->o_ent.p, 0, io * sizeof(float), s));
  A = &G;
  SCV_TRY(build_cluster_program(D, A, B, M));
  if (D->cl_active) D->small_active = false; else SCV_TRY(build_small_phases(D, A, B, M));
  if (D->small_active)
    SCV_CUDA(cudaMemcpyAsync(D->sm_phases.p, D->sm_host.data(), D->sm_host.size() * sizeof(SmallPhase), cudaMemcpyHostToDevice, s));
  // per-layer K/V projection of the memory tokens, once per call instead of once per step and layer
  // (the reference re-projects them inside nn.MultiheadAttention at every step, :1302-1307)
  void* mem_split = nullptr;
  if (use_tensor_cores(c, B)) {      // split the memory tokens once instead of once per layer and column tile
    SCV_TRY(D->msplit.ensure(split_tile_bytes(Bm * M, d)));
    mem_split = D->msplit.p;
    SCV_TRY(launch_layernorm_split(A->memory, d, nullptr, nullptr, mem_split, Bm * M, d, 0, nullptr, s));
  }
  for (int li = 0; li < c.num_layers; ++li) {
    const DecLayer& L = D->layers[li];
    LinearArgs a;
    a.a_split = mem_split;
    a.x = A->memory; a.ldx = d; a.w = L.ca_in_w + (size_t)d * L.ca_in_ld; a.ldw = L.ca_in_ld; a.wt = L.ca_kv_wt;
    a.bias = L.ca_in_b + d; a.y = D->ckv.as<float>() + (size_t)li * Bm * M * 2 * d; a.ldy = 2 * d;
    a.M = Bm * M; a.N = 2 * d; a.K = d;
    SCV_TRY(launch_linear(a, 0, s));
  }
  D->pinned[0] = D->pinned[1] = 0;
  bool used[2] = {false, false};
  // Profiling pass (bench.py roofline): the host waits for every step, so no step is enqueued after the batch has
  // finished (no-op launches would be timed and their flops counted); see the gate kernel in the step loop.
  const bool prof = prof_enabled();
  const bool sync_each = (A->flags & SCV_FLAG_SYNC_EVERY_STEP) != 0 || prof;
  if (prof) { SCV_CUDA(cudaStreamSynchronize(s)); SCV_TRY(prof_harvest()); }    // memory K/V projections above
  const bool two_phase = !(A->temperature < 0.01f) || A->want_entropy;
  // Sub-batches: the ~150 dependent kernels of a step are short at these sizes (a few microseconds of tensor work
  // behind fixed launch / prologue / tail costs), so the batch is decoded as independent row ranges on separate
  // streams whose kernels overlap on the GPU.  Semantics stay batch-global: one shared StepState (step counter,
  // unfinished-row count, H2 flag), sub-batches re-join before the second sampler kernel and before step_end.
  int n_sub = 1;
  {
    const int forced = tun().subbatches;
    n_sub = forced > 0 ? forced : (B >= tun().sub_min_rows ? 2 : 1);
    // shared memory tokens: a launch over ALL rows lets one warp serve every sample of a (latent, head); row ranges would
    // split a latent's samples between the streams (config 3: 177.9 ms against 185.4 with two ranges)
    if (forced <= 0 && Bm < B && tun().attn_shared != 0) n_sub = 1;
    if (prof_enabled()) n_sub = 1;       // per-kernel timing wants one kernel at a time
    n_sub = std::max(1, std::min(n_sub, kMaxSub));
    while (n_sub > 1 && round_up(ceil_div(B, n_sub), 128) * (n_sub - 1) >= B) --n_sub;
  }
  const int sub_rows = n_sub > 1 ? round_up(ceil_div(B, n_sub), 128) : B;
  auto enqueue_step = [&](int step) -> int {
    if (prof) { prof_step_begin(); ProfScope null_pair(PC_NULL, s, 0.0, 0.0); }
    if (n_sub == 1) {
      SCV_TRY(decode_rows(D, A, steps_max, step, 0, B, 1, s));
      if (two_phase) SCV_TRY(decode_rows(D, A, steps_max, step, 0, B, 2, s));
    } else {
      SCV_CUDA(cudaEventRecord(D->ev_fork, s));
      for (int i = 0; i < n_sub; ++i) {
        const int r0 = i * sub_rows, nb = std::min(sub_rows, B - r0);
        SCV_CUDA(cudaStreamWaitEvent(D->sub[i], D->ev_fork, 0));
        SCV_TRY(decode_rows(D, A, steps_max, step, r0, nb, 1, D->sub[i]));
        SCV_CUDA(cudaEventRecord(D->ev_sub[i], D->sub[i]));
      }
      for (int i = 0; i < n_sub; ++i) SCV_CUDA(cudaStreamWaitEvent(s, D->ev_sub[i], 0));
      if (two_phase) {       // the H2 flag is batch-global: every row's first sampler kernel precedes any second one
        SCV_CUDA(cudaEventRecord(D->ev_fork, s));
        for (int i = 0; i < n_sub; ++i) {
          const int r0 = i * sub_rows, nb = std::min(sub_rows, B - r0);
          SCV_CUDA(cudaStreamWaitEvent(D->sub[i], D->ev_fork, 0));
          SCV_TRY(decode_rows(D, A, steps_max, step, r0, nb, 2, D->sub[i]));
          SCV_CUDA(cudaEventRecord(D->ev_sub[i], D->sub[i]));
        }
        for (int i = 0; i < n_sub; ++i) SCV_CUDA(cudaStreamWaitEvent(s, D->ev_sub[i], 0));
      }
    }
    if ((A->flags & SCV_FLAG_COMPACT_FINISHED) != 0 && !D->small_active && !D->cl_active)
      SCV_TRY(launch_compact_rows(D->row_map.as<int>(), D->cur.as<int>(), D->fin.as<unsigned char>(), st, B, s));
    SCV_TRY(launch_step_end(st, steps_max, s));
    return 0;
  };
  // CUDA graph of one step: the launch sequence is identical for every step (position, done flag, RNG seed all live
  // in device memory), so step 0 runs eagerly and steps >= 1 replay one instantiated graph (one launch per step
  // instead of ~150-300); cached per call configuration.
  const bool use_graph = tun().graph != 0 && !sync_each && steps_max >= 3;
  const GraphKey key{B, M, steps_max, n_sub, A->top_k, A->type_masks != nullptr, A->want_log_probs, A->want_entropy,
                     A->forced_tokens != nullptr, prof ? 1 : 0, A->memory_rows, A->flags, A->temperature, A->top_p, A->stop_boost,
                     A->hard_stop_threshold, A->site_dup_threshold};
  cudaGraphExec_t exec = nullptr;
  if (use_graph)
    for (auto& g : D->graphs) if (g.key == key) exec = g.exec;
  // Small batches, plain greedy: the step loop runs inside one persistent kernel (decode_small.cu), no host polling.
  static const int persist_env = [] { const char* e = getenv("SCV_SMALL_PERSIST"); return e ? atoi(e) : 1; }();
  const bool persist = (D->small_active || D->cl_active) && persist_env != 0 && !two_phase && !sync_each && A->temperature < 0.01f &&
                       c.vocab_size % 4 == 0 && (reinterpret_cast<uintptr_t>(A->type_masks) & 3u) == 0;   // sampler_plain_greedy
  if (persist) SCV_TRY(decode_rows(D, A, steps_max, 0, 0, B, 3, s));
  for (int step = 0; step < steps_max && !persist; ++step) {
    if (use_graph && step >= 1) {
      if (exec == nullptr) {
        cudaGraph_t graph = nullptr;
        SCV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_step(step);
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
        SCV_CUDA(ce);
        SCV_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        if (D->graphs.size() >= 8) D->drop_graphs();
        D->graphs.push_back({key, exec});
      }
      SCV_CUDA(cudaGraphLaunch(exec, s));
      count_launch(D->launches_per_step);
    } else {
      const long long before = launch_total();
      if (prof) {
        // Hold the stream behind a gate kernel while the whole step (kernels + timing events) is enqueued, then open it:
        // the device runs the step back to back, so no interval between two events contains host launch latency.
        // (Events recorded inside a captured graph cannot be read with cudaEventElapsedTime, hence no graph here.)
        volatile int* gate = D->pinned + 20;
        *gate = 0;
        SCV_TRY(launch_host_gate(D->pinned + 20, s));
        const int rc = enqueue_step(step);
        __atomic_store_n(D->pinned + 20, 1, __ATOMIC_RELEASE);
        if (rc != 0) return rc;
      } else {
        SCV_TRY(enqueue_step(step));
      }
      D->launches_per_step = (int)(launch_total() - before);
    }
    if (sync_each) {
      SCV_CUDA(cudaStreamSynchronize(s));
      if (prof) {
        SCV_TRY(prof_harvest());
        int done_now = 0;
        SCV_CUDA(cudaMemcpy(&done_now, &st->done, sizeof(int), cudaMemcpyDeviceToHost));
        if (done_now != 0) break;
        continue;
      }
    }
    const int poll = use_graph ? 1 : 2;
    if ((step % poll) == poll - 1 && step + 1 < steps_max) {
      // bound the host's run-ahead to <= 2 * poll steps and stop enqueueing once every row has finished
      const int k = (step / poll) & 1;
      if (used[k]) {
        SCV_CUDA(cudaEventSynchronize(D->ev[k]));
        if (D->pinned[k] != 0) break;
      }
      SCV_CUDA(cudaMemcpyAsync(&D->pinned[k], &st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
      SCV_CUDA(cudaEventRecord(D->ev[k], s));
      used[k] = true;
    }
  }
  SCV_CUDA(cudaMemcpyAsync(A_user->out_tokens, D->o_tok.p, io * sizeof(long long), cudaMemcpyDeviceToDevice, s));
  if (A->want_log_probs) SCV_CUDA(cudaMemcpyAsync(A_user->out_log_probs, D->o_lp.p, io * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (A->want_entropy) SCV_CUDA(cudaMemcpyAsync(A_user->out_entropy, D->o_ent.p, io * sizeof(float), cudaMemcpyDeviceToDevice, s));
  SCV_CUDA(cudaMemcpyAsync(&D->pinned[2], st, sizeof(StepState), cudaMemcpyDeviceToHost, s));
  SCV_CUDA(cudaStreamSynchronize(s));
  const StepState* hs = reinterpret_cast<const StepState*>(&D->pinned[2]);
  *A_user->out_steps = hs->done ? hs->out_len : hs->step;
  if (D->cl_active && D->cl_prog.dbg != nullptr) {       // SCV_CLUSTER_DEBUG: where one CTA's step time goes
    std::vector<unsigned long long> h(D->cl_kinds.size());
    SCV_CUDA(cudaMemcpy(h.data(), D->cl_prog.dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const char* names[] = {"layernorm", "projection", "self-attention", "cross-attention", "cluster barrier"};
    double tot[5] = {0, 0, 0, 0, 0}; int cnt[5] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < h.size(); ++i) { tot[D->cl_kinds[i] / 10000] += (double)h[i]; cnt[D->cl_kinds[i] / 10000] += 1; }
    const int steps = std::max(1, (int)*A_user->out_steps);
    for (int k = 0; k < 5; ++k)
      fprintf(stderr, "cluster decode: %-16s %3d per step, %9.1f us per step in total (cluster 0, rank 0)\n", names[k], cnt[k], tot[k] / 1e3 / steps);
    for (size_t i = 0; i < h.size() && i < 24; ++i)
      fprintf(stderr, "  instr %2zu kind %d K/8 %4d: %8.2f us per step\n", i, D->cl_kinds[i] / 10000, D->cl_kinds[i] % 10000, (double)h[i] / 1e3 / steps);
  }
  D->last_B = B;
  return 0;
}

// Teacher-forced forward (:901-985, teacher_forcing_ratio = 1.0): the layer stack of decode_rows() applied to
// R = batch * seq_len rows at once.  Row r = b * L + t is position t of sequence b; the self-attention of a row
// attends over rows b * L .. r of the qkv buffer itself (no cache), skipping padded keys.
int scv_decoder_forward(scv_decoder* D, const scv_forward_args* A, void* stream) {
  SCV_REQUIRE(D && A, "forward: null argument");
  SCV_REQUIRE(A->batch > 0 && A->seq_len > 0 && A->n_memory > 0 && A->memory && A->tokens && A->out_logits &&
                  A->ld_tokens >= A->seq_len, "forward: bad arguments");
  SCV_REQUIRE(scv_decoder_missing_weights(D) == 0, "forward: weights missing");
  const scv_decoder_config& c = D->cfg;
  SCV_REQUIRE(A->seq_len <= c.pe_len, "forward: %d positions exceed the positional-encoding buffer (%d)", A->seq_len, c.pe_len);
  const bool have_dup = D->ws.loaded("site_dup_head.0.weight") && D->ws.loaded("site_dup_head.2.weight");
  SCV_REQUIRE(A->out_dup == nullptr || have_dup, "forward: out_dup needs the site_dup_head weights");
  const long long Rll = (long long)A->batch * A->seq_len;
  SCV_REQUIRE(Rll <= 262144, "forward: %lld rows (batch * seq_len) exceed one call's limit of 262144", Rll);
  const int B = A->batch, L = A->seq_len, M = A->n_memory, R = (int)Rll;
  const int d = c.d_model, hd = d / c.nhead, dff = c.dim_feedforward;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  set_pdl_for_call(true);
  const size_t f = sizeof(float);
  SCV_TRY(D->x.ensure((size_t)R * d * f)); SCV_TRY(D->xn.ensure((size_t)R * d * f));
  SCV_TRY(D->qkv.ensure((size_t)R * 3 * d * f)); SCV_TRY(D->attn.ensure((size_t)R * d * f));
  SCV_TRY(D->q2.ensure((size_t)R * d * f)); SCV_TRY(D->ff.ensure((size_t)R * dff * f));
  SCV_TRY(D->h1.ensure((size_t)R * d * f)); SCV_TRY(D->h2.ensure((size_t)R * d * f));
  SCV_TRY(D->t3.ensure((size_t)R * d * f)); SCV_TRY(D->tlog.ensure((size_t)R * 8 * f));
  SCV_TRY(D->ckv.ensure((size_t)c.num_layers * B * M * 2 * d * f));
  SCV_TRY(D->fw_skip.ensure((size_t)R));
  const bool tc = use_tensor_cores(c, R);
  if (tc) {
    for (DevBuf* b : {&D->xn_s, &D->attn_s, &D->h2_s}) {
      const size_t need = split_tile_bytes(R, d);
      if (need > b->cap) { SCV_TRY(b->ensure(need)); SCV_CUDA(cudaMemset(b->p, 0, need)); }
    }
    const size_t need = split_tile_bytes(R, dff);
    if (need > D->ff_s.cap) { SCV_TRY(D->ff_s.ensure(need)); SCV_CUDA(cudaMemset(D->ff_s.p, 0, need)); }
  }
  float* x = D->x.as<float>(); float* xn = D->xn.as<float>(); float* qkv = D->qkv.as<float>();
  float* attn = D->attn.as<float>(); float* q2 = D->q2.as<float>(); float* ff = D->ff.as<float>();
  float* h1 = D->h1.as<float>(); float* h2 = D->h2.as<float>(); float* t3 = D->t3.as<float>();
  void* xn_s = tc ? D->xn_s.p : nullptr; void* attn_s = tc ? D->attn_s.p : nullptr;
  void* ff_s = tc ? D->ff_s.p : nullptr; void* h2_s = tc ? D->h2_s.p : nullptr;
  unsigned char* skip = D->fw_skip.as<unsigned char>();
  const float scale = (float)(1.0 / std::sqrt((double)hd));
  auto norm = [&](const LNp& P, float* fp32_out) -> int {
    return tc ? launch_layernorm_split(x, d, P.g, P.b, xn_s, R, d, 1, nullptr, s)
              : launch_layernorm(x, d, P.g, P.b, fp32_out, d, R, d, ACT_NONE, nullptr, s);
  };
  SCV_TRY(launch_embed_sequence(D->emb, D->ld_emb, D->pe, d, reinterpret_cast<const long long*>(A->tokens), A->ld_tokens,
                                B, L, c.vocab_size, x, skip, s));
  // K / V projection of the memory tokens, once per layer (shared by all L positions of a sequence)
  void* mem_split = nullptr;
  if (use_tensor_cores(c, B * M)) {
    SCV_TRY(D->msplit.ensure(split_tile_bytes(B * M, d)));
    mem_split = D->msplit.p;
    SCV_TRY(launch_layernorm_split(A->memory, d, nullptr, nullptr, mem_split, B * M, d, 0, nullptr, s));
  }
  for (int li = 0; li < c.num_layers; ++li) {
    const DecLayer& Lr = D->layers[li];
    LinearArgs a;
    a.a_split = mem_split;
    a.x = A->memory; a.ldx = d; a.w = Lr.ca_in_w + (size_t)d * Lr.ca_in_ld; a.ldw = Lr.ca_in_ld; a.wt = Lr.ca_kv_wt;
    a.bias = Lr.ca_in_b + d; a.y = D->ckv.as<float>() + (size_t)li * B * M * 2 * d; a.ldy = 2 * d;
    a.M = B * M; a.N = 2 * d; a.K = d;
    SCV_TRY(launch_linear(a, 0, s));
  }
  for (int li = 0; li < c.num_layers; ++li) {
    const DecLayer& Lr = D->layers[li];
    SCV_TRY(norm(Lr.n1, xn));
    LinearArgs a;
    a.x = xn; a.ldx = d; a.a_split = xn_s; a.w = Lr.sa_in_w; a.ldw = Lr.sa_in_ld; a.wt = Lr.sa_in_wt; a.bias = Lr.sa_in_b;
    a.y = qkv; a.ldy = 3 * d; a.M = R; a.N = 3 * d; a.K = d;
    SCV_TRY(launch_linear(a, 0, s));
    AttnArgs sa;
    sa.q = qkv; sa.ldq = 3 * d; sa.kcache = qkv + d; sa.vcache = qkv + 2 * d; sa.seq_stride = (long long)L * 3 * d;
    sa.row_stride = 3 * d; sa.out = attn; sa.ldo = d; sa.B = R; sa.nhead = c.nhead; sa.hd = hd; sa.scale = scale;
    sa.fixed_len = -1; sa.rows_per_seq = L; sa.key_skip = (A->flags & SCV_FORWARD_NO_KEY_PADDING) ? nullptr : skip;
    sa.max_n = std::max(L, M); sa.host_len_hint = (L + 1) / 2;
    sa.out_split = static_cast<unsigned char*>(attn_s); sa.kb_out = d / 64;
    SCV_TRY(launch_attention(sa, s));
    LinearArgs o = lin_args(attn, d, Lr.sa_out, x, d, R, ACT_NONE);
    o.residual = x; o.ldr = d; o.a_split = attn_s;
    SCV_TRY(launch_linear(o, 0, s));
    SCV_TRY(norm(Lr.n2, xn));
    LinearArgs q;
    q.x = xn; q.ldx = d; q.a_split = xn_s; q.w = Lr.ca_in_w; q.ldw = Lr.ca_in_ld; q.wt = Lr.ca_q_wt; q.bias = Lr.ca_in_b;
    q.y = q2; q.ldy = d; q.M = R; q.N = d; q.K = d;
    SCV_TRY(launch_linear(q, 0, s));
    AttnArgs ca;
    float* ckv = D->ckv.as<float>() + (size_t)li * B * M * 2 * d;
    ca.q = q2; ca.ldq = d; ca.kcache = ckv; ca.vcache = ckv + d; ca.seq_stride = (long long)M * 2 * d; ca.row_stride = 2 * d;
    ca.out = attn; ca.ldo = d; ca.B = R; ca.nhead = c.nhead; ca.hd = hd; ca.scale = scale; ca.fixed_len = M;
    ca.rows_per_seq = L; ca.max_n = std::max(L, M);
    ca.out_split = static_cast<unsigned char*>(attn_s); ca.kb_out = d / 64;
    SCV_TRY(launch_attention(ca, s));
    LinearArgs co = lin_args(attn, d, Lr.ca_out, x, d, R, ACT_NONE);
    co.residual = x; co.ldr = d; co.a_split = attn_s;
    SCV_TRY(launch_linear(co, 0, s));
    SCV_TRY(norm(Lr.n3, xn));
    LinearArgs f1 = lin_args(xn, d, Lr.ff1, ff, dff, R, ACT_GELU);
    f1.a_split = xn_s; f1.y_split = ff_s;
    SCV_TRY(launch_linear(f1, 0, s));
    LinearArgs f2 = lin_args(ff, dff, Lr.ff2, x, d, R, ACT_NONE);
    f2.residual = x; f2.ldr = d; f2.a_split = ff_s;
    SCV_TRY(launch_linear(f2, 0, s));
  }
  // heads at every position (:975-979): logits straight into the caller's tensor
  SCV_TRY(norm(D->out_ln, h1));
  LinearArgs oa = lin_args(h1, d, D->out_a, h2, d, R, ACT_GELU);
  oa.a_split = xn_s; oa.y_split = h2_s;
  SCV_TRY(launch_linear(oa, 0, s));
  LinearArgs ob = lin_args(h2, d, D->out_b, A->out_logits, c.vocab_size, R, ACT_NONE);
  ob.a_split = h2_s;
  SCV_TRY(launch_linear(ob, 0, s));
  if (A->out_type != nullptr) {
    SCV_TRY(norm(D->tt_ln, h1));
    LinearArgs ta = lin_args(h1, d, D->tt_a, h2, d, R, ACT_GELU);
    ta.a_split = xn_s; ta.y_split = h2_s;
    SCV_TRY(launch_linear(ta, 0, s));
    LinearArgs tb = lin_args(h2, d, D->tt_b, t3, d / 4, R, ACT_GELU);
    tb.a_split = h2_s;
    SCV_TRY(launch_linear(tb, 0, s));
    SCV_TRY(launch_linear(lin_args(t3, d / 4, D->tt_c, D->tlog.as<float>(), 8, R, ACT_NONE), 0, s));
    SCV_CUDA(cudaMemcpy2DAsync(A->out_type, 5 * f, D->tlog.p, 8 * f, 5 * f, (size_t)R, cudaMemcpyDeviceToDevice, s));
  }
  if (A->out_stop != nullptr) {
    SCV_TRY(launch_linear(lin_args(x, d, D->stop_a, t3, d / 4, R, ACT_GELU), 0, s));
    SCV_TRY(launch_linear(lin_args(t3, d / 4, D->stop_b, A->out_stop, 1, R, ACT_NONE), 0, s));
  }
  if (A->out_dup != nullptr) {
    SCV_TRY(launch_linear(lin_args(x, d, D->dup_a, t3, d / 4, R, ACT_GELU), 0, s));
    SCV_TRY(launch_linear(lin_args(t3, d / 4, D->dup_b, A->out_dup, 1, R, ACT_NONE), 0, s));
  }
  D->last_B = 0;     // the step taps no longer describe a decode
  return 0;
}

int scv_decoder_debug_read(scv_decoder* D, int32_t what, float* dst, int64_t numel, void* stream) {
  SCV_REQUIRE(D && dst && D->last_B > 0, "debug_read: nothing has been decoded yet");
  const int B = D->last_B;
  const float* src = nullptr;
  int64_t n = 0;
  switch (what) {
    case 0: src = D->x.as<float>(); n = (int64_t)B * D->cfg.d_model; break;
    case 1: src = D->logits.as<float>(); n = (int64_t)B * D->cfg.vocab_size; break;
    case 2: src = D->tlog.as<float>(); n = (int64_t)B * 8; break;
    case 3: src = D->slog.as<float>(); n = B; break;
    default: SCV_REQUIRE(false, "debug_read: unknown tap %d", what);
  }
  SCV_REQUIRE(numel == n, "debug_read: tap %d holds %lld floats, caller asked for %lld", what, (long long)n,
              (long long)numel);
  return launch_copy_f32(src, dst, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
