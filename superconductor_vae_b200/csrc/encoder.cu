// FullMaterialsVAE engine: three-branch encoder -> z, and the heads that feed the decoder's memory tokens.
// Reference: src/superconductor/models/attention_vae.py:350-606 (module), :625-676 (encode),
// :678-709 + :733-770 (heads), :236-307 (hierarchical family head);
// src/superconductor/encoders/element_attention.py:73-98, 152-214 (element embedding + attention).
#include <string>
#include <vector>

#include "../../include/scvae_b200.h"
#include "weights.cuh"

using namespace scv;

namespace {

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
  // like ensure(), but a (re)allocated buffer starts as zeros (padding columns nobody writes must read as 0)
  int ensure_zeroed(size_t bytes) {
    if (bytes <= cap) return 0;
    SCV_TRY(ensure(bytes));
    SCV_CUDA(cudaMemset(p, 0, bytes));
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  float* f() const { return static_cast<float*>(p); }
};

// weighted[b, e, :] = element_embed[idx[b, e]] * frac[b, e]      (attention_vae.py:113-118)
__global__ void elem_embed_kernel(const __nv_bfloat16* __restrict__ table, int ld, int n_rows,
                                  const long long* __restrict__ idx, const float* __restrict__ frac,
                                  float* __restrict__ out, int rows, int dim) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  long long id = idx[r];
  if (id < 0 || id >= n_rows) id = 0;
  const float w = frac[r];
  for (int i = threadIdx.x; i < dim; i += blockDim.x)
    out[(size_t)r * dim + i] = __bfloat162float(table[(size_t)id * ld + i]) * w;
}

// Learned-query attention over the element slots (element_attention.py:182-208).
// One CTA per row, thread t <-> hidden unit t = head * head_dim + j.
__global__ void elem_attention_kernel(const float* __restrict__ keys, const float* __restrict__ values,
                                      const __nv_bfloat16* __restrict__ query, int ldq,
                                      const unsigned char* __restrict__ mask, float* __restrict__ attended,
                                      float* __restrict__ attn_w, int E, int nh, int hd) {
  extern __shared__ float sm[];      // scores [nh * E]
  const int b = blockIdx.x, hidden = nh * hd;
  const float inv = 1.0f / (sqrtf((float)hd) * 1.0f);
  for (int i = threadIdx.x; i < nh * E; i += blockDim.x) {
    const int h = i / E, e = i % E;
    float s = 0.f;
    for (int j = 0; j < hd; ++j)
      s = fmaf(__bfloat162float(query[h * ldq + j]), keys[((size_t)b * E + e) * hidden + h * hd + j], s);
    s = s * inv;
    if (mask[(size_t)b * E + e] == 0) s = -INFINITY;
    sm[i] = s;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < nh; h += blockDim.x) {
    float m = -INFINITY;
    for (int e = 0; e < E; ++e) m = fmaxf(m, sm[h * E + e]);
    float sum = 0.f;
    for (int e = 0; e < E; ++e) { const float v = expf(sm[h * E + e] - m); sm[h * E + e] = v; sum += v; }
    for (int e = 0; e < E; ++e) sm[h * E + e] = sm[h * E + e] / sum;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < hidden; t += blockDim.x) {
    const int h = t / hd;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(sm[h * E + e], values[((size_t)b * E + e) * hidden + t], acc);
    attended[(size_t)b * hidden + t] = acc;
  }
  if (attn_w != nullptr)
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      float s = 0.f;
      for (int h = 0; h < nh; ++h) s += sm[h * E + e];
      attn_w[(size_t)b * E + e] = s / (float)nh;
    }
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows,
                                 int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
  }
}

__global__ void sigmoid_col_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) dst[(size_t)r * ldd] = sigmoidf_(src[(size_t)r * lds]);
}

struct FinalizeArgs {
  const float* sc_in; int ld_sc; int latent; int magpie_dim; int max_el;    // the assembled sc_head input row
  const float* sc_pred; const float* coarse; const float* cup; const float* iron;
  scv_encoder_heads_out out;
  int B;
};

__device__ void softmax_small(const float* x, int n, float* p) {
  float m = -INFINITY;
  for (int i = 0; i < n; ++i) m = fmaxf(m, x[i]);
  float s = 0.f;
  for (int i = 0; i < n; ++i) { p[i] = expf(x[i] - m); s += p[i]; }
  for (int i = 0; i < n; ++i) p[i] = p[i] / s;
}

// composed_14 (attention_vae.py:268-300) and the decoder conditioning rows
// (stoich_pred scripts/train_v12_clean.py:5249; heads order models/autoregressive_decoder.py:845-858).
__global__ void heads_finalize_kernel(FinalizeArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const float* row = a.sc_in + (size_t)b * a.ld_sc;
  const int c_tc = a.latent, c_mag = c_tc + 1, c_hp = c_mag + a.magpie_dim, c_fr = c_hp + 1,
            c_cnt = c_fr + a.max_el, c_comp = c_cnt + 1, c_cls = c_comp + 1;
  const float tc = row[c_tc], hp = row[c_hp], cnt = row[c_cnt], comp = row[c_comp], sc = a.sc_pred[b];
  float cp[7], up[6], ip[2], fam[14];
  softmax_small(a.coarse + (size_t)b * 8, 7, cp);
  softmax_small(a.cup + (size_t)b * 8, 6, up);
  softmax_small(a.iron + (size_t)b * 2, 2, ip);
  const float sp = sigmoidf_(sc);
  fam[0] = 1.0f - sp;
  fam[1] = sp * cp[0];
  const float cup_p = sp * cp[1], iron_p = sp * cp[2];
  for (int i = 0; i < 6; ++i) fam[2 + i] = cup_p * up[i];
  for (int i = 0; i < 2; ++i) fam[8 + i] = iron_p * ip[i];
  fam[10] = sp * cp[3]; fam[11] = sp * cp[4]; fam[12] = sp * cp[5]; fam[13] = sp * cp[6];
  const scv_encoder_heads_out& o = a.out;
  if (o.tc_pred) o.tc_pred[b] = tc;
  if (o.hp_pred) o.hp_pred[b] = hp;
  if (o.sc_pred) o.sc_pred[b] = sc;
  if (o.competence) o.competence[b] = comp;
  if (o.element_count_pred) o.element_count_pred[b] = cnt;
  if (o.tc_class_logits) for (int i = 0; i < 5; ++i) o.tc_class_logits[(size_t)b * 5 + i] = row[c_cls + i];
  if (o.fraction_pred) for (int i = 0; i < a.max_el; ++i) o.fraction_pred[(size_t)b * a.max_el + i] = row[c_fr + i];
  if (o.stoich_pred) for (int i = 0; i <= a.max_el; ++i) o.stoich_pred[(size_t)b * (a.max_el + 1) + i] = row[c_fr + i];
  if (o.family_coarse_logits) for (int i = 0; i < 7; ++i) o.family_coarse_logits[(size_t)b * 7 + i] = a.coarse[(size_t)b * 8 + i];
  if (o.family_cuprate_sub_logits) for (int i = 0; i < 6; ++i) o.family_cuprate_sub_logits[(size_t)b * 6 + i] = a.cup[(size_t)b * 8 + i];
  if (o.family_iron_sub_logits) for (int i = 0; i < 2; ++i) o.family_iron_sub_logits[(size_t)b * 2 + i] = a.iron[(size_t)b * 2 + i];
  if (o.family_composed_14) for (int i = 0; i < 14; ++i) o.family_composed_14[(size_t)b * 14 + i] = fam[i];
  if (o.heads_input) {
    float* h = o.heads_input + (size_t)b * 24;
    h[0] = tc; h[1] = sc; h[2] = hp;
    for (int i = 0; i < 5; ++i) h[3 + i] = row[c_cls + i];
    h[8] = comp; h[9] = cnt;
    for (int i = 0; i < 14; ++i) h[10 + i] = fam[i];
  }
}

struct Mlp2 { Lin a, b; };

}  // namespace

struct scv_encoder {
  scv_encoder_config cfg{};
  WeightStore ws;
  __nv_bfloat16* elem_table = nullptr; int ld_table = 0;
  __nv_bfloat16* query = nullptr; int ld_query = 0;
  Lin key_proj, value_proj, att_out; LNp att_ln;
  Lin elem_out; LNp elem_out_ln;
  Lin mag_a, mag_b; LNp mag_ln_a, mag_ln_b;
  Lin tc_a, tc_b; LNp tc_ln;
  Lin fusion; LNp fusion_ln;
  std::vector<Lin> enc_lin; std::vector<LNp> enc_ln;
  Lin fc_mean;
  std::vector<Lin> bb_lin; std::vector<LNp> bb_ln;
  Lin tc_proj, res_a, res_b; LNp res_ln, tc_out_ln; Lin tc_out_a, tc_out_b;
  Lin mag_head_a, mag_head_b, att_head; LNp att_head_ln;
  Lin comp_a, comp_b, frac_a, frac_b, frac_c; LNp frac_ln;
  Lin hp_a, hp_b, cls_a, cls_b;
  Lin sc_a, sc_b, sc_c; LNp sc_ln;
  Lin co_a, co_b, co_c, cu_a, cu_b, cu_c, ir_a, ir_b; LNp co_ln, cu_ln, ir_ln;
  // never executed by the reference forward (element_properties=None) but present in its state_dict
  Lin prop_enc, combiner; LNp prop_ln;
  DevBuf t0, t1, t2, fused_in, cond, sc_in, small, zsplit, scsplit, condsplit;

  ~scv_encoder() { for (DevBuf* b : {&t0, &t1, &t2, &fused_in, &cond, &sc_in, &small, &zsplit, &scsplit, &condsplit}) b->release(); }
};

// nn.Linear with an extra tcgen05-tiled weight copy whenever the tensor-core path can take the projection
static int add_lin(WeightStore& W, const std::string& prefix, int N, int K, Lin* out) {
  const bool tiled = K >= 64 && K % 4 == 0 && N % 4 == 0;
  return W.add_linear(prefix, N, K, out, true, tiled);
}

static int enc_register(scv_encoder* E) {
  const scv_encoder_config& c = E->cfg;
  WeightStore& W = E->ws;
  const int e = c.element_embed_dim, f = c.fusion_dim;
  std::string p = "element_encoder.element_embedding.";
  E->elem_table = W.add_matrix(p + "element_embed.weight", c.n_element_rows, e, &E->ld_table);
  if (!E->elem_table) return 2;
  SCV_TRY(add_lin(W, p + "property_encoder.0", e, 11, &E->prop_enc));
  SCV_TRY(W.add_layernorm(p + "property_encoder.1", e, &E->prop_ln));
  SCV_TRY(add_lin(W, p + "combiner", e, 2 * e, &E->combiner));
  for (const char* n : {"property_encoder.0.weight", "property_encoder.0.bias", "property_encoder.1.weight",
                        "property_encoder.1.bias", "combiner.weight", "combiner.bias"})
    W.mark_optional(p + n);
  p = "element_encoder.element_attention.";
  E->query = W.add_matrix(p + "query", c.n_attention_heads, e / c.n_attention_heads, &E->ld_query);
  if (!E->query) return 2;
  SCV_TRY(add_lin(W, p + "key_proj", e, e, &E->key_proj));
  SCV_TRY(add_lin(W, p + "value_proj", e, e, &E->value_proj));
  SCV_TRY(add_lin(W, p + "output_proj", e, e, &E->att_out));
  SCV_TRY(W.add_layernorm(p + "layer_norm", e, &E->att_ln));
  SCV_TRY(add_lin(W, "element_encoder.output_projection.0", f, e, &E->elem_out));
  SCV_TRY(W.add_layernorm("element_encoder.output_projection.1", f, &E->elem_out_ln));
  SCV_TRY(add_lin(W, "magpie_encoder.0", 2 * f, c.magpie_dim, &E->mag_a));
  SCV_TRY(W.add_layernorm("magpie_encoder.1", 2 * f, &E->mag_ln_a));
  SCV_TRY(add_lin(W, "magpie_encoder.4", f, 2 * f, &E->mag_b));
  SCV_TRY(W.add_layernorm("magpie_encoder.5", f, &E->mag_ln_b));
  SCV_TRY(add_lin(W, "tc_encoder.0", f / 2, 1, &E->tc_a));
  SCV_TRY(add_lin(W, "tc_encoder.2", f, f / 2, &E->tc_b));
  SCV_TRY(W.add_layernorm("tc_encoder.3", f, &E->tc_ln));
  SCV_TRY(add_lin(W, "fusion.0", 3 * f, 3 * f, &E->fusion));
  SCV_TRY(W.add_layernorm("fusion.1", 3 * f, &E->fusion_ln));
  int prev = 3 * f;
  E->enc_lin.resize(c.n_encoder_hidden); E->enc_ln.resize(c.n_encoder_hidden);
  for (int j = 0; j < c.n_encoder_hidden; ++j) {
    SCV_TRY(add_lin(W, "vae_encoder.encoder." + std::to_string(3 * j), c.encoder_hidden[j], prev, &E->enc_lin[j]));
    SCV_TRY(W.add_layernorm("vae_encoder.encoder." + std::to_string(3 * j + 1), c.encoder_hidden[j], &E->enc_ln[j]));
    prev = c.encoder_hidden[j];
  }
  SCV_TRY(add_lin(W, "vae_encoder.fc_mean", c.latent_dim, prev, &E->fc_mean));
  prev = c.latent_dim;
  E->bb_lin.resize(c.n_decoder_hidden); E->bb_ln.resize(c.n_decoder_hidden);
  for (int j = 0; j < c.n_decoder_hidden; ++j) {
    SCV_TRY(add_lin(W, "decoder_backbone." + std::to_string(4 * j), c.decoder_hidden[j], prev, &E->bb_lin[j]));
    SCV_TRY(W.add_layernorm("decoder_backbone." + std::to_string(4 * j + 1), c.decoder_hidden[j], &E->bb_ln[j]));
    prev = c.decoder_hidden[j];
  }
  const int bb = prev, L = c.latent_dim;
  SCV_TRY(add_lin(W, "tc_proj", 256, bb, &E->tc_proj));
  SCV_TRY(add_lin(W, "tc_res_block.0", 256, 256, &E->res_a));
  SCV_TRY(W.add_layernorm("tc_res_block.1", 256, &E->res_ln));
  SCV_TRY(add_lin(W, "tc_res_block.4", 256, 256, &E->res_b));
  SCV_TRY(W.add_layernorm("tc_out.0", 256, &E->tc_out_ln));
  SCV_TRY(add_lin(W, "tc_out.2", 128, 256, &E->tc_out_a));
  SCV_TRY(add_lin(W, "tc_out.4", 1, 128, &E->tc_out_b));
  SCV_TRY(add_lin(W, "magpie_head.0", bb, bb, &E->mag_head_a));
  SCV_TRY(add_lin(W, "magpie_head.2", c.magpie_dim, bb, &E->mag_head_b));
  SCV_TRY(add_lin(W, "attended_head.0", f, bb, &E->att_head));
  SCV_TRY(W.add_layernorm("attended_head.1", f, &E->att_head_ln));
  SCV_TRY(add_lin(W, "competence_head.0", L / 4, L, &E->comp_a));
  SCV_TRY(add_lin(W, "competence_head.2", 1, L / 4, &E->comp_b));
  SCV_TRY(add_lin(W, "fraction_head.0", 256, L, &E->frac_a));
  SCV_TRY(W.add_layernorm("fraction_head.1", 256, &E->frac_ln));
  SCV_TRY(add_lin(W, "fraction_head.4", 128, 256, &E->frac_b));
  SCV_TRY(add_lin(W, "fraction_head.6", c.max_elements + 1, 128, &E->frac_c));
  SCV_TRY(add_lin(W, "hp_head.0", 256, L, &E->hp_a));
  SCV_TRY(add_lin(W, "hp_head.2", 1, 256, &E->hp_b));
  SCV_TRY(add_lin(W, "tc_class_head.0", 256, bb, &E->cls_a));
  SCV_TRY(add_lin(W, "tc_class_head.3", 5, 256, &E->cls_b));
  const int sc_in = L + 1 + c.magpie_dim + 1 + c.max_elements + 1 + 1 + 5;
  // K = 2214 / 513 are not multiples of 4: no fp32-row tensor path, but the tiled weights are zero padded to whole 64-wide
  // k-blocks, so a SplitTile copy of the (zero padded) input row feeds the tensor cores at large batch (scv_encoder_heads)
  SCV_TRY(W.add_linear("sc_head.0", 512, sc_in, &E->sc_a, true, true));
  SCV_TRY(W.add_layernorm("sc_head.2", 512, &E->sc_ln));
  SCV_TRY(add_lin(W, "sc_head.4", 128, 512, &E->sc_b));
  SCV_TRY(add_lin(W, "sc_head.6", 1, 128, &E->sc_c));
  p = "hierarchical_family_head.";
  SCV_TRY(W.add_linear(p + "coarse_head.0", 256, bb + 1, &E->co_a, true, true));
  SCV_TRY(W.add_layernorm(p + "coarse_head.1", 256, &E->co_ln));
  SCV_TRY(add_lin(W, p + "coarse_head.4", 128, 256, &E->co_b));
  SCV_TRY(add_lin(W, p + "coarse_head.6", 7, 128, &E->co_c));
  SCV_TRY(W.add_linear(p + "cuprate_sub_head.0", 128, bb + 1, &E->cu_a, true, true));
  SCV_TRY(W.add_layernorm(p + "cuprate_sub_head.1", 128, &E->cu_ln));
  SCV_TRY(add_lin(W, p + "cuprate_sub_head.4", 64, 128, &E->cu_b));
  SCV_TRY(add_lin(W, p + "cuprate_sub_head.6", 6, 64, &E->cu_c));
  SCV_TRY(W.add_linear(p + "iron_sub_head.0", 64, bb + 1, &E->ir_a, true, true));
  SCV_TRY(W.add_layernorm(p + "iron_sub_head.1", 64, &E->ir_ln));
  SCV_TRY(add_lin(W, p + "iron_sub_head.4", 2, 64, &E->ir_b));
  return 0;
}

static int lin(const float* x, int ldx, const Lin& L, float* y, int ldy, int M, int act, cudaStream_t s,
               const float* residual = nullptr, int ldr = 0, const void* x_split = nullptr) {
  LinearArgs a;
  a.a_split = (x_split != nullptr && L.wt != nullptr && L.N % 4 == 0) ? x_split : nullptr;   // SplitTile copy of x
  a.x = x; a.ldx = ldx; a.w = L.w; a.ldw = L.ldw; a.wt = L.wt; a.bias = L.b; a.y = y; a.ldy = ldy; a.M = M; a.N = L.N; a.K = L.K;
  a.act = act; a.residual = residual; a.ldr = ldr;
  return launch_linear(a, 0, s);
}
static int ln(const float* x, int ldx, const LNp& P, float* y, int ldy, int M, int act, cudaStream_t s) {
  return launch_layernorm(x, ldx, P.g, P.b, y, ldy, M, P.N, act, nullptr, s);
}

extern "C" {

int scv_encoder_create(const scv_encoder_config* cfg, scv_encoder** out) {
  SCV_REQUIRE(cfg && out, "null argument");
  SCV_REQUIRE(cfg->n_attention_heads > 0 && cfg->element_embed_dim % cfg->n_attention_heads == 0, "bad head split");
  SCV_REQUIRE(cfg->n_encoder_hidden >= 0 && cfg->n_encoder_hidden <= 4 && cfg->n_decoder_hidden >= 1 &&
                  cfg->n_decoder_hidden <= 4, "hidden lists must have <= 4 entries");
  SCV_REQUIRE(cfg->max_elements > 0 && cfg->max_elements <= 64, "bad max_elements");
  scv_encoder* E = new scv_encoder();
  E->cfg = *cfg;
  const int rc = enc_register(E);
  if (rc != 0) { delete E; return rc; }
  *out = E;
  return 0;
}

void scv_encoder_destroy(scv_encoder* enc) { delete enc; }

int scv_encoder_load_weight(scv_encoder* enc, const char* name, const float* src, int64_t numel, void* stream) {
  SCV_REQUIRE(enc && name && src, "null argument");
  return enc->ws.load(name, src, numel, static_cast<cudaStream_t>(stream));
}

int scv_encoder_missing_weights(scv_encoder* enc) {
  std::string first;
  const int n = enc->ws.missing(&first);
  if (n > 0) set_error("%d state_dict entries not loaded, first: %s", n, first.c_str());
  return n;
}

int scv_encoder_encode(scv_encoder* E, int32_t B, const int64_t* idx, const float* frac, const uint8_t* mask,
                       const float* magpie, const float* tc, float* z_out, float* attn_w_out, float* fused_out,
                       void* stream) {
  SCV_REQUIRE(E && idx && frac && mask && magpie && tc && z_out && B > 0, "encode: bad arguments");
  SCV_REQUIRE(scv_encoder_missing_weights(E) == 0, "encode: weights missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const scv_encoder_config& c = E->cfg;
  const int e = c.element_embed_dim, f = c.fusion_dim, El = c.max_elements, F3 = 3 * f;
  int wide = std::max(std::max(El * e, 2 * f), F3);
  for (int j = 0; j < c.n_encoder_hidden; ++j) wide = std::max(wide, c.encoder_hidden[j]);
  SCV_TRY(E->t0.ensure((size_t)B * wide * sizeof(float)));
  SCV_TRY(E->t1.ensure((size_t)B * wide * sizeof(float)));
  SCV_TRY(E->t2.ensure((size_t)B * wide * sizeof(float)));
  SCV_TRY(E->fused_in.ensure((size_t)B * F3 * sizeof(float)));
  float *t0 = E->t0.f(), *t1 = E->t1.f(), *t2 = E->t2.f(), *fin = E->fused_in.f();
  // --- branch 1: element attention
  elem_embed_kernel<<<B * El, 128, 0, s>>>(E->elem_table, E->ld_table, c.n_element_rows,
                                           reinterpret_cast<const long long*>(idx), frac, t0, B * El, e);
  SCV_LAUNCH_CHECK();
  SCV_TRY(lin(t0, e, E->key_proj, t1, e, B * El, ACT_NONE, s));
  SCV_TRY(lin(t0, e, E->value_proj, t2, e, B * El, ACT_NONE, s));
  const int nh = c.n_attention_heads;
  elem_attention_kernel<<<B, 128, (size_t)nh * El * sizeof(float), s>>>(t1, t2, E->query, E->ld_query, mask, t0,
                                                                          attn_w_out, El, nh, e / nh);
  SCV_LAUNCH_CHECK();
  SCV_TRY(lin(t0, e, E->att_out, t1, e, B, ACT_NONE, s));
  SCV_TRY(ln(t1, e, E->att_ln, t1, e, B, ACT_NONE, s));
  SCV_TRY(lin(t1, e, E->elem_out, t0, f, B, ACT_NONE, s));
  SCV_TRY(ln(t0, f, E->elem_out_ln, fin, F3, B, ACT_GELU, s));
  // --- branch 2: magpie MLP
  SCV_TRY(lin(magpie, c.magpie_dim, E->mag_a, t0, 2 * f, B, ACT_NONE, s));
  SCV_TRY(ln(t0, 2 * f, E->mag_ln_a, t0, 2 * f, B, ACT_GELU, s));
  SCV_TRY(lin(t0, 2 * f, E->mag_b, t1, f, B, ACT_NONE, s));
  SCV_TRY(ln(t1, f, E->mag_ln_b, fin + f, F3, B, ACT_GELU, s));
  // --- branch 3: Tc embedding
  SCV_TRY(lin(tc, 1, E->tc_a, t0, f / 2, B, ACT_GELU, s));
  SCV_TRY(lin(t0, f / 2, E->tc_b, t1, f, B, ACT_NONE, s));
  SCV_TRY(ln(t1, f, E->tc_ln, fin + 2 * f, F3, B, ACT_GELU, s));
  // --- fusion + VAE encoder
  SCV_TRY(lin(fin, F3, E->fusion, t0, F3, B, ACT_NONE, s));
  float* fused = fused_out ? fused_out : t1;
  SCV_TRY(ln(t0, F3, E->fusion_ln, fused, F3, B, ACT_GELU, s));
  const float* h = fused;
  int hdim = F3;
  float* ping[2] = {t0, t2};
  const void* h_split = nullptr;
  for (int j = 0; j < c.n_encoder_hidden; ++j) {
    float* o = ping[j & 1];
    SCV_TRY(lin(h, hdim, E->enc_lin[j], o, c.encoder_hidden[j], B, ACT_NONE, s));
    const bool last = j == c.n_encoder_hidden - 1;
    if (last && c.encoder_hidden[j] % 64 == 0 && c.encoder_hidden[j] <= 1024) {
      // the only consumer is fc_mean: normalise + GELU straight into the tensor-core operand form
      SCV_TRY(E->zsplit.ensure(split_tile_bytes(B, c.encoder_hidden[j])));
      SCV_TRY(launch_layernorm_split(o, c.encoder_hidden[j], E->enc_ln[j].g, E->enc_ln[j].b, E->zsplit.p, B,
                                     c.encoder_hidden[j], 1, nullptr, s, ACT_GELU));
      h_split = E->zsplit.p;
    } else {
      SCV_TRY(ln(o, c.encoder_hidden[j], E->enc_ln[j], o, c.encoder_hidden[j], B, ACT_GELU, s));
    }
    h = o; hdim = c.encoder_hidden[j];
  }
  SCV_TRY(lin(h, hdim, E->fc_mean, z_out, c.latent_dim, B, ACT_NONE, s, nullptr, 0, h_split));
  return 0;
}

int scv_encoder_heads(scv_encoder* E, int32_t B, const float* z, const scv_encoder_heads_out* out, void* stream) {
  SCV_REQUIRE(E && z && out && B > 0, "heads: bad arguments");
  SCV_REQUIRE(scv_encoder_missing_weights(E) == 0, "heads: weights missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const scv_encoder_config& c = E->cfg;
  const int L = c.latent_dim, f = c.fusion_dim, md = c.magpie_dim, El = c.max_elements;
  const int bb = c.decoder_hidden[c.n_decoder_hidden - 1];
  const int ld_cond = round_up(bb + 1, 8);
  const int sc_dim = L + 1 + md + 1 + El + 1 + 1 + 5, ld_sc = round_up(sc_dim, 8);
  // Large batches: the two ragged-K projections (sc_head.0: K = 2214, the family heads: K = 513) go to the tensor cores
  // through a SplitTile copy of their zero-padded input rows (52.8 K rows: 4.4 ms of fp32 CUDA-core work otherwise).
  // Smaller calls keep the fp32 CUDA-core kernel (every product exact): the heads feed the decoder's conditioning, and the
  // greedy decode is held to token-exactness against the fp32 oracle.
  const bool ragged_tc = B >= tun().cond_tc_min_rows;      // default 16384 (common.cuh)
  int wide = std::max(std::max(L / 4, 512), bb);
  for (int j = 0; j < c.n_decoder_hidden; ++j) wide = std::max(wide, c.decoder_hidden[j]);
  SCV_TRY(E->t0.ensure((size_t)B * wide * sizeof(float)));
  SCV_TRY(E->t1.ensure((size_t)B * wide * sizeof(float)));
  SCV_TRY(E->t2.ensure((size_t)B * wide * sizeof(float)));
  SCV_TRY(E->cond.ensure_zeroed((size_t)B * ld_cond * sizeof(float)));
  SCV_TRY(E->sc_in.ensure_zeroed((size_t)B * ld_sc * sizeof(float)));
  SCV_TRY(E->small.ensure((size_t)B * 32 * sizeof(float)));
  float *t0 = E->t0.f(), *t1 = E->t1.f(), *t2 = E->t2.f(), *cond = E->cond.f(), *sci = E->sc_in.f();
  float* sc_pred = E->small.f();            // [B]
  float* coarse = sc_pred + B;              // [B, 8]
  float* cup = coarse + (size_t)B * 8;      // [B, 8]
  float* iron = cup + (size_t)B * 8;        // [B, 2]
  const int c_tc = L, c_mag = c_tc + 1, c_hp = c_mag + md, c_fr = c_hp + 1, c_comp = c_fr + El + 1, c_cls = c_comp + 1;
  // sc_head input row starts with z itself (attention_vae.py:756-765)
  copy_cols_kernel<<<std::min(ceil_div(B * L, 256), 148 * 8), 256, 0, s>>>(z, L, sci, ld_sc, B, L);
  SCV_LAUNCH_CHECK();
  // z feeds the backbone, competence, fraction and high-pressure heads: split it once for all four projections
  const void* z_split = nullptr;
  if (L % 64 == 0) {
    SCV_TRY(E->zsplit.ensure(split_tile_bytes(B, L)));
    SCV_TRY(launch_layernorm_split(z, L, nullptr, nullptr, E->zsplit.p, B, L, 0, nullptr, s));
    z_split = E->zsplit.p;
  }
  // backbone h -> cond[:, :bb]  (decode, :689)
  const float* h = z;
  int hdim = L;
  for (int j = 0; j < c.n_decoder_hidden; ++j) {
    const bool last = j == c.n_decoder_hidden - 1;
    float* o = (j & 1) ? t1 : t0;
    SCV_TRY(lin(h, hdim, E->bb_lin[j], o, c.decoder_hidden[j], B, ACT_NONE, s, nullptr, 0, j == 0 ? z_split : nullptr));
    float* dst = last ? cond : o;
    const int ldd = last ? ld_cond : c.decoder_hidden[j];
    SCV_TRY(ln(o, c.decoder_hidden[j], E->bb_ln[j], dst, ldd, B, ACT_GELU, s));
    h = dst; hdim = c.decoder_hidden[j];
  }
  // Tc head (:693-695)
  SCV_TRY(lin(cond, ld_cond, E->tc_proj, t0, 256, B, ACT_NONE, s));            // tc_h
  SCV_TRY(lin(t0, 256, E->res_a, t1, 256, B, ACT_NONE, s));
  SCV_TRY(ln(t1, 256, E->res_ln, t1, 256, B, ACT_GELU, s));
  SCV_TRY(lin(t1, 256, E->res_b, t2, 256, B, ACT_NONE, s, t0, 256));           // tc_h + res_block(tc_h)
  SCV_TRY(ln(t2, 256, E->tc_out_ln, t2, 256, B, ACT_GELU, s));
  SCV_TRY(lin(t2, 256, E->tc_out_a, t1, 128, B, ACT_GELU, s));
  SCV_TRY(lin(t1, 128, E->tc_out_b, sci + c_tc, ld_sc, B, ACT_NONE, s));
  // magpie head (:697)
  SCV_TRY(lin(cond, ld_cond, E->mag_head_a, t0, bb, B, ACT_GELU, s));
  SCV_TRY(lin(t0, bb, E->mag_head_b, sci + c_mag, ld_sc, B, ACT_NONE, s));
  if (out->magpie_pred) {
    copy_cols_kernel<<<std::min(ceil_div(B * md, 256), 148 * 8), 256, 0, s>>>(sci + c_mag, ld_sc, out->magpie_pred, md, B, md);
    SCV_LAUNCH_CHECK();
  }
  // attended head (:698)
  if (out->attended_input) {
    SCV_TRY(lin(cond, ld_cond, E->att_head, t0, f, B, ACT_NONE, s));
    SCV_TRY(ln(t0, f, E->att_head_ln, out->attended_input, f, B, ACT_NONE, s));
  }
  // Tc class head (:701)
  SCV_TRY(lin(cond, ld_cond, E->cls_a, t0, 256, B, ACT_GELU, s));
  SCV_TRY(lin(t0, 256, E->cls_b, sci + c_cls, ld_sc, B, ACT_NONE, s));
  // competence (:734)
  SCV_TRY(lin(z, L, E->comp_a, t0, L / 4, B, ACT_GELU, s, nullptr, 0, z_split));
  SCV_TRY(lin(t0, L / 4, E->comp_b, sci + c_comp, ld_sc, B, ACT_SIGMOID, s));
  // fraction head (:738-740): 12 fractions + count land contiguously in the sc_head input
  SCV_TRY(lin(z, L, E->frac_a, t0, 256, B, ACT_NONE, s, nullptr, 0, z_split));
  SCV_TRY(ln(t0, 256, E->frac_ln, t0, 256, B, ACT_GELU, s));
  SCV_TRY(lin(t0, 256, E->frac_b, t1, 128, B, ACT_GELU, s));
  SCV_TRY(lin(t1, 128, E->frac_c, sci + c_fr, ld_sc, B, ACT_NONE, s));
  // high-pressure head (:747)
  SCV_TRY(lin(z, L, E->hp_a, t0, 256, B, ACT_RELU, s, nullptr, 0, z_split));
  SCV_TRY(lin(t0, 256, E->hp_b, sci + c_hp, ld_sc, B, ACT_NONE, s));
  // SC head over the concatenation (:756-766): Linear, GELU, LayerNorm, Linear, GELU, Linear
  if (ragged_tc) {
    SCV_TRY(E->scsplit.ensure(split_tile_bytes(B, ld_sc)));
    SCV_TRY(launch_layernorm_split(sci, ld_sc, nullptr, nullptr, E->scsplit.p, B, ld_sc, 0, nullptr, s));
    SCV_TRY(lin(sci, ld_sc, E->sc_a, t0, 512, B, ACT_GELU, s, nullptr, 0, E->scsplit.p));
  } else {
    SCV_TRY(lin(sci, ld_sc, E->sc_a, t0, 512, B, ACT_GELU, s));
  }
  SCV_TRY(ln(t0, 512, E->sc_ln, t0, 512, B, ACT_NONE, s));
  SCV_TRY(lin(t0, 512, E->sc_b, t1, 128, B, ACT_GELU, s));
  SCV_TRY(lin(t1, 128, E->sc_c, sc_pred, 1, B, ACT_NONE, s));
  // hierarchical family head on cat[h, sigmoid(sc_pred)] (:256-266)
  sigmoid_col_kernel<<<ceil_div(B, 256), 256, 0, s>>>(sc_pred, 1, cond + bb, ld_cond, B);
  SCV_LAUNCH_CHECK();
  const void* cond_split = nullptr;
  if (ragged_tc) {
    SCV_TRY(E->condsplit.ensure(split_tile_bytes(B, ld_cond)));
    SCV_TRY(launch_layernorm_split(cond, ld_cond, nullptr, nullptr, E->condsplit.p, B, ld_cond, 0, nullptr, s));
    cond_split = E->condsplit.p;
  }
  SCV_TRY(lin(cond, ld_cond, E->co_a, t0, 256, B, ACT_NONE, s, nullptr, 0, cond_split));
  SCV_TRY(ln(t0, 256, E->co_ln, t0, 256, B, ACT_GELU, s));
  SCV_TRY(lin(t0, 256, E->co_b, t1, 128, B, ACT_GELU, s));
  SCV_TRY(lin(t1, 128, E->co_c, coarse, 8, B, ACT_NONE, s));
  SCV_TRY(lin(cond, ld_cond, E->cu_a, t0, 128, B, ACT_NONE, s, nullptr, 0, cond_split));
  SCV_TRY(ln(t0, 128, E->cu_ln, t0, 128, B, ACT_GELU, s));
  SCV_TRY(lin(t0, 128, E->cu_b, t1, 64, B, ACT_GELU, s));
  SCV_TRY(lin(t1, 64, E->cu_c, cup, 8, B, ACT_NONE, s));
  SCV_TRY(lin(cond, ld_cond, E->ir_a, t0, 64, B, ACT_NONE, s, nullptr, 0, cond_split));
  SCV_TRY(ln(t0, 64, E->ir_ln, t0, 64, B, ACT_GELU, s));
  SCV_TRY(lin(t0, 64, E->ir_b, iron, 2, B, ACT_NONE, s));
  FinalizeArgs fa;
  fa.sc_in = sci; fa.ld_sc = ld_sc; fa.latent = L; fa.magpie_dim = md; fa.max_el = El;
  fa.sc_pred = sc_pred; fa.coarse = coarse; fa.cup = cup; fa.iron = iron; fa.out = *out; fa.B = B;
  heads_finalize_kernel<<<ceil_div(B, 128), 128, 0, s>>>(fa);
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
