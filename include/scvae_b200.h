/*
 * scvae_b200 - C ABI of the B200-native KV-cache decode engine.
 *
 * The reference (jamesconde/superconductor-vae) has no FFI: its boundary for this path is the
 * Python method surface of two nn.Modules.  This header is the boundary a
 * Python host binds with ctypes: plain pointers and sizes, no torch types.
 * Every entry point cites the reference interface it replaces
 * (paths relative to the reference root, src/superconductor/...).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller unless the comment says HOST;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work on it unless stated;
 *   - return value 0 = ok, non-zero = error, message via scv_last_error() (thread local);
 *   - no exceptions cross this boundary; one host thread per engine.
 */
#ifndef SCVAE_B200_H
#define SCVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCV_ABI_VERSION 1

typedef struct scv_decoder scv_decoder;
typedef struct scv_encoder scv_encoder;

/* ---------------------------------------------------------------- general */
int scv_abi_version(void);
const char* scv_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py "gpu_launches") */
int64_t scv_launch_count(void);

/* Run-time tunables of the launch configuration (no reference counterpart: the reference has no launch configuration).
 * Keys: "attn_ctas_per_sm" (0 = one CTA per 8 (row, head) items; k > 0 = persistent attention grid of k CTAs per SM so
 * that another sub-batch's projections can run on the same SMs), "gemm_stages" (0 auto / 2 / 4), "subbatches"
 * (0 auto; row ranges decoded on separate streams), "sub_min_rows", "graph" (0 / 1: CUDA-graph step replay),
 * "gemm_bn64" / "gemm_bn64_max_ctas" (128 x 64 projection tiles for launches of at most that many row tiles / CTAs),
 * "gemm_mc" / "gemm_mc_min_row_tiles" / "gemm_mc_min_kblocks" (opt-in: column-tile pairs share the A tile by multicast),
 * "attn_bulk" / "attn_bulk_min_rows" / "attn_bulk_piece_kb" (opt-in: bulk-copy staged cross-attention), "attn_shared"
 * (shared memory tokens: one warp per (latent, head) serves all samples), "attn_forward" / "attn_forward_min_ctas"
 * (teacher-forced passes: K / V of a (sequence, head) staged in shared memory), "attn_pages_regs", "cond_tc_min_rows"
 * (rows from which the small conditioning projections use the tensor cores), "cluster" / "cluster_max_rows" /
 * "cluster_rows" (opt-in cluster-parallel small-batch kernel).  No setting changes a token (tests/test_gpu_parity.py
 * test_tunables_never_change_tokens).  Defaults come from the SCV_* environment variables read at load (common.cuh
 * Tunables).  Unknown keys return 1. */
int scv_tune(const char* key, int32_t value);

/* Debug: CTA residency trace of the step kernels (projection and attention CTAs log kernel id, SM, start and end time).
 * scv_trace_begin arms it for up to max_records CTAs; scv_trace_read synchronises the device, disarms it and copies the
 * records (32 bytes each: u64 t0_ns, u64 t1_ns, u32 sm, u32 kernel id, u32 blockIdx.x, u32 blockIdx.y) to HOST memory. */
int scv_trace_begin(int32_t max_records);
int scv_trace_read(void* records_host, int32_t max_records, int32_t* n_out);

/* Per-category kernel timing with CUDA events on the launch stream (bench.py roofline pass; adds overhead,
 * never enabled inside a timed region).  Categories are indexed 0..n-1, see scv_profile_category_name. */
int scv_profile_begin(void);
int scv_profile_end(int32_t n_cats, int32_t* counts, double* ms, double* flops, double* bytes);
const char* scv_profile_category_name(int32_t cat);

/* ---------------------------------------------------------------- decoder */
/* Constructor arguments of EnhancedTransformerDecoder (models/autoregressive_decoder.py:564-598). */
typedef struct scv_decoder_config {
  int32_t d_model, nhead, num_layers, dim_feedforward, vocab_size;
  int32_t pe_len;                 /* rows of pos_encoding.pe (= constructor max_len) */
  int32_t latent_dim, n_memory_tokens, memory_bottleneck_dim; /* 0 = pre-V15 direct MLP (:638-644) */
  int32_t stoich_input_dim, n_stoich_tokens;                  /* 0 tokens = no stoich conditioning */
  int32_t heads_input_dim, heads_n_tokens;                    /* :756-765 */
  int32_t encoder_skip_dim, skip_n_tokens;                    /* 0 tokens = use_skip_connection=False */
} scv_decoder_config;

int scv_decoder_create(const scv_decoder_config* cfg, scv_decoder** out);
void scv_decoder_destroy(scv_decoder* dec);

/* Copy one state_dict entry (fp32, contiguous, device) into engine-owned storage; matrices are
 * rounded to bf16 once here.  `name` is the reference's state_dict key
 * (SURVEY.md section 8b "State-dict layout"), e.g.
 * "transformer_decoder.layers.3.self_attn.in_proj_weight".  Unknown names return 1. */
int scv_decoder_load_weight(scv_decoder* dec, const char* name, const float* src, int64_t numel, void* stream);
/* 0 when every tensor the config requires has been loaded, else the count missing
 * (first missing name in scv_last_error()). */
int scv_decoder_missing_weights(scv_decoder* dec);

/* _create_memory / precompute_memory (models/autoregressive_decoder.py:779-899).
 * z [B, latent_dim]; skip [B, encoder_skip_dim] or NULL; stoich [B, stoich_input_dim] or NULL;
 * heads_in [B, heads_input_dim] or NULL (already concatenated in the order of :845-858).
 * memory_out [B, n_tokens, d_model] fp32 where n_tokens = n_memory_tokens (+skip) (+stoich) (+heads)
 * for the non-NULL inputs; *n_tokens_out (HOST) receives it. */
int scv_decoder_build_memory(scv_decoder* dec, int32_t batch, const float* z, const float* skip,
                             const float* stoich, const float* heads_in, float* memory_out,
                             int32_t* n_tokens_out, void* stream);

#define SCV_FLAG_H2_UNIFORM_FALLBACK 1u /* reproduce :1464-1466/:1512-1513 (batch-global degenerate guard) */
#define SCV_FLAG_SYNC_EVERY_STEP 2u     /* debugging: synchronise after every step */
#define SCV_FLAG_COMPACT_FINISHED 4u    /* opt-in: rows that have emitted END are retired (the reference keeps decoding them until
                                           every row has finished, :1541-1548): outputs are identical up to and including
                                           each row's first END - what every caller consumes - and PAD / 0 after it */

/* generate_with_kv_cache (models/autoregressive_decoder.py:1321-1557). */
typedef struct scv_generate_args {
  int32_t batch;
  int32_t n_memory;        /* memory tokens per row (16, 20, 24, +8 with skip) */
  int32_t max_len;         /* reference max_len; executed steps <= min(max_len, pe_len) - 1 */
  const float* memory;     /* [batch, n_memory, d_model] fp32 (cached_memory) */
  float temperature;       /* < 0.01 -> argmax (:1506); 0.0 reproduces the reference's divide */
  int32_t top_k;           /* <= 0: off */
  float top_p;             /* >= 1: off */
  float stop_boost, hard_stop_threshold, site_dup_threshold;
  const uint8_t* type_masks; /* [5, vocab] 0/1 or NULL */
  int32_t want_log_probs, want_entropy;
  uint32_t flags;
  uint64_t seed, offset;   /* Philox key / counter offset for multinomial */
  int64_t* out_tokens;     /* [batch, max_len-1] row stride = max_len-1 (clamped) */
  float* out_log_probs;    /* same shape or NULL */
  float* out_entropy;      /* same shape or NULL */
  int32_t* out_steps;      /* HOST: number of executed steps L (columns >= L are not written) */
  const int64_t* forced_tokens; /* optional [batch, max_len-1]: teacher-forced replay (tests) */
  int32_t memory_rows;     /* 0: `memory` has `batch` rows.  B0 > 0: `memory` has B0 rows and row r of the batch is conditioned on
                              memory row r % B0 - the reference's RLOO layout, z.repeat(k, 1) (scripts/train_v12_clean.py:2677-2688),
                              without materialising the k copies: the projected memory K / V are built once per latent and the
                              k samples of a latent read the same K / V.  batch % B0 == 0, batch > 64. */
} scv_generate_args;

/* Enqueues the decode loop and synchronises `stream` once at the end to read *out_steps. */
int scv_decoder_generate(scv_decoder* dec, const scv_generate_args* args, void* stream);

/* Teacher-forced forward, teacher_forcing_ratio = 1.0 (EnhancedTransformerDecoder.forward,
 * models/autoregressive_decoder.py:901-985): every position of every row in one pass -- projections with
 * batch * seq_len rows, causal self-attention with the key padding mask of :952 (input id == PAD), cross-attention to
 * the row's memory tokens, then output_proj / stop_head / token_type_head / site_dup_head at every position. */
typedef struct scv_forward_args {
  int32_t batch;
  int32_t seq_len;          /* L = target length - 1 (<= pe_len) */
  int32_t n_memory;
  const float* memory;      /* [batch, n_memory, d_model] fp32 */
  const int64_t* tokens;    /* input ids target[:, :-1]: [batch, L], row stride ld_tokens elements */
  int32_t ld_tokens;
  float* out_logits;        /* [batch, L, vocab] */
  float* out_stop;          /* [batch, L] or NULL */
  float* out_type;          /* [batch, L, 5] or NULL */
  float* out_dup;           /* [batch, L] or NULL; needs the site_dup_head weights */
  uint32_t flags;           /* SCV_FORWARD_NO_KEY_PADDING: attend PAD inputs like any other key (generation semantics) */
} scv_forward_args;
#define SCV_FORWARD_NO_KEY_PADDING 1u
int scv_decoder_forward(scv_decoder* dec, const scv_forward_args* args, void* stream);

/* The greedy epilogue of generate_with_kv_cache (:1415-1509: type mask, stop boost, hard stop, length boost, / temperature,
 * first-occurrence argmax) applied to EVERY position of a teacher-forced pass at once (draft verification):
 * row r = (sequence r / seq_len, position r % seq_len).  logits [R, vocab], type_logits [R, 5] or NULL (then type_masks is
 * ignored), stop_logits [R] or NULL (then stop_boost is ignored), type_masks [5, vocab] bytes or NULL, finished_before [R]
 * bytes (1 = the sequence emitted END at an earlier position: no hard stop), out_tokens [R]. */
int scv_greedy_positions(const float* logits, const float* type_logits, const float* stop_logits, const uint8_t* type_masks,
                         const uint8_t* finished_before, int64_t n_rows, int32_t seq_len, int32_t vocab, int32_t max_len,
                         float temperature, float stop_boost, float hard_stop_threshold, int64_t* out_tokens, void* stream);

/* Debug taps (tests): copy engine-internal fp32 state of the LAST executed step to `dst`.
 * what: 0 final hidden [B,d]; 1 raw logits [B,V]; 2 type logits [B,5]; 3 stop logit [B] */
int scv_decoder_debug_read(scv_decoder* dec, int32_t what, float* dst, int64_t numel, void* stream);

/* ---------------------------------------------------------------- encoder */
/* Constructor arguments of FullMaterialsVAE (models/attention_vae.py:350-362). */
typedef struct scv_encoder_config {
  int32_t n_element_rows;   /* rows of element_embed.weight (n_elements + 1) */
  int32_t element_embed_dim, n_attention_heads, max_elements;
  int32_t magpie_dim, fusion_dim, latent_dim;
  int32_t n_encoder_hidden; int32_t encoder_hidden[4];
  int32_t n_decoder_hidden; int32_t decoder_hidden[4];
} scv_encoder_config;

int scv_encoder_create(const scv_encoder_config* cfg, scv_encoder** out);
void scv_encoder_destroy(scv_encoder* enc);
int scv_encoder_load_weight(scv_encoder* enc, const char* name, const float* src, int64_t numel, void* stream);
int scv_encoder_missing_weights(scv_encoder* enc);

/* FullMaterialsVAE.encode (models/attention_vae.py:625-676).
 * element_indices int64 [B, E]; fractions fp32 [B, E]; mask uint8 [B, E]; magpie [B, magpie_dim];
 * tc [B].  Outputs: z [B, latent], attention_weights [B, E] (may be NULL), fused [B, 3*fusion] (may be NULL). */
int scv_encoder_encode(scv_encoder* enc, int32_t batch, const int64_t* element_indices,
                       const float* element_fractions, const uint8_t* element_mask, const float* magpie,
                       const float* tc, float* z_out, float* attention_weights_out, float* fused_out,
                       void* stream);

/* FullMaterialsVAE.decode + head section of forward (attention_vae.py:678-709, 733-770, 236-307).
 * Every output may be NULL. */
typedef struct scv_encoder_heads_out {
  float* tc_pred;            /* [B] */
  float* magpie_pred;        /* [B, magpie_dim] */
  float* attended_input;     /* [B, fusion_dim] */
  float* tc_class_logits;    /* [B, 5] */
  float* competence;         /* [B] */
  float* fraction_pred;      /* [B, max_elements] */
  float* element_count_pred; /* [B] */
  float* hp_pred;            /* [B] */
  float* sc_pred;            /* [B] */
  float* family_coarse_logits;      /* [B, 7] */
  float* family_cuprate_sub_logits; /* [B, 6] */
  float* family_iron_sub_logits;    /* [B, 2] */
  float* family_composed_14;        /* [B, 14] */
  float* stoich_pred;        /* [B, max_elements+1] = cat(fraction_pred, count) (train_v12_clean.py:5249) */
  float* heads_input;        /* [B, 24] in the order the decoder concatenates (autoregressive_decoder.py:845-858) */
} scv_encoder_heads_out;

int scv_encoder_heads(scv_encoder* enc, int32_t batch, const float* z, const scv_encoder_heads_out* out,
                      void* stream);

/* slerp (scripts/holdout/holdout_search.py:128-146), rows z1[i1[r]], z2[i2[r]], t[r] -> out[r].
 * The reference's batch-global lerp fallback (:137-138) is applied when *any* row is near-parallel. */
int scv_slerp_rows(const float* anchors, int32_t dim, const int32_t* i1, const int32_t* i2, const float* t,
                   int64_t n_rows, float* out, int32_t* fallback_flag_dev, void* stream);

/* Candidate post-processing (replaces the per-row Python loop of tokens_to_formula,
 * scripts/holdout/holdout_search.py:88-99, and FractionAwareTokenizer.decode, tokenizer/fraction_tokenizer.py:478-519,
 * at scale): tokens [n_rows, row_len] int64 -> canonical [n_rows, row_len] int16 (ids up to the first END, PAD after),
 * hash [n_rows] (64-bit FNV-1a of the canonical ids) and, if not NULL, length [n_rows] (position of the first END). */
int scv_tokens_canonical_hash(const int64_t* tokens, int64_t n_rows, int32_t row_len, int16_t* canonical,
                              uint64_t* hash, int32_t* length, void* stream);

/* element_similarity (scripts/holdout/holdout_search.py:149-182) of every candidate x target pair: 0.5 * Jaccard of the
 * element sets + 0.5 * sum over shared elements of min(normalised amounts).  comp_a [n_a, n_elements], comp_b
 * [n_b, n_elements]: amounts per element column as parse_formula_elements (:109-125) returns them, a NEGATIVE value marks
 * an element the formula does not contain; doubles like the reference's Python floats.  out [n_a, n_b]. */
int scv_element_similarity(const double* comp_a, const double* comp_b, int32_t n_a, int32_t n_b, int32_t n_elements,
                           double* out, void* stream);

/* Token-level rollout reward (replaces compute_reward_gpu_native,
 * src/superconductor/losses/reward_gpu_native.py:448-722, called after every rollout at
 * scripts/train_v12_clean.py:2745-2752, 2829-2836, 2942-2950).  The struct carries the fields of GPURewardConfig
 * (:42-79) and GPURewardConfigV14 (:82-131) under their reference names; `v14` = "the caller passed a
 * GPURewardConfigV14" (the reference branches on isinstance, :564). */
typedef struct scv_reward_config {
  float exact_match, near_exact_1, near_exact_2, near_exact_3, token_correct, token_penalty, length_mismatch_penalty;
  float fraction_digit_penalty, fraction_structure_penalty;
  int32_t use_semantic_digit_penalty;
  float semantic_digit_scale, length_only_base_reward, length_only_per_extra, length_only_floor;
  int32_t v14, use_continuous_reward;
  float max_reward, sharpness, element_error_penalty, integer_error_penalty, fraction_error_penalty, special_error_penalty;
  float too_short_base_reward, too_short_per_missing, too_short_floor;
  int32_t use_phased_curriculum, reward_phase;
  float phase3_sharpness;
  int32_t v14_element_start, v14_element_end, v14_integer_start, v14_integer_end, v14_fraction_start;
} scv_reward_config;
/* sampled, target: int64 [batch, seq_len] with `row_stride` elements between rows; mask: one byte per position
 * (nonzero = valid), same stride; fraction_values: float [n_fraction_values] or NULL (then, as in the reference, the
 * digit-level penalties of the pre-V13 vocabulary apply even if use_semantic_fractions is set); rewards: float [batch].
 * All pointers are device pointers. */
int scv_reward_tokens(const int64_t* sampled, const int64_t* target, const uint8_t* mask, int32_t batch,
                      int32_t seq_len, int64_t row_stride, const scv_reward_config* config, int32_t end_idx,
                      int32_t use_semantic_fractions, int32_t fraction_token_start, const float* fraction_values,
                      int32_t n_fraction_values, float* rewards, void* stream);

/* Chemistry-constraint rewards of a rollout (replaces compute_constraint_rewards,
 * src/superconductor/losses/constraint_rewards.py:629-676 = A1 :270-303 + A2 :306-379 + A4 :382-459 + A7 :462-507 +
 * B1-B8 :510-626 over the parser :172-267; call sites scripts/train_v12_clean.py:2754-2766, 2990-3007).  The struct
 * carries VocabConfig (:29-56), ConstraintRewardConfig (:132-149) and FamilyConstraintConfig (:152-167) under their
 * reference names; penalties and the threshold are doubles like the Python floats they replace; b_penalty[i] = b{i+1}. */
typedef struct scv_constraint_config {
  int32_t element_start, element_end, digit_start, digit_end, lparen_idx, rparen_idx, slash_idx, pad_idx, end_idx;
  int32_t use_semantic_fractions, fraction_token_start;
  int32_t a1_enabled, a2_enabled, a4_enabled, a7_enabled, family_enabled;
  double a1_penalty, a2_penalty_per_violation, a4_penalty, a7_penalty, confidence_threshold;
  double b_penalty[8];
} scv_constraint_config;
/* sampled: int64 [batch, seq_len] (non-negative ids), mask: one byte per position, both with `row_stride` elements
 * between rows; fraction_values: float [n_fraction_values] or NULL; family_probs: float [batch, n_families] or NULL (no
 * family rules); rewards: float [batch].  Device pointers. */
int scv_constraint_rewards(const int64_t* sampled, const uint8_t* mask, int32_t batch, int32_t seq_len,
                           int64_t row_stride, const scv_constraint_config* config, const float* fraction_values,
                           int32_t n_fraction_values, const float* family_probs, int32_t n_families, float* rewards,
                           void* stream);

/* ---------------------------------------------------------------- kernel-level taps (tests) */
/* y[M,N] = act(x[M,K] * w[N,K]^T + bias) (+ residual); w is bf16 with row stride ldw (elements).
 * act: 0 none, 1 gelu(erf), 2 relu, 3 sigmoid.  impl: 1 SIMT fp32 (w_bf16 row-major), 2 tcgen05 hi/lo bf16
 * (w_bf16 in the tile layout of scv_op_pack_tiled; ldw ignored). */
int scv_op_linear(const float* x, int32_t ldx, const uint16_t* w_bf16, int32_t ldw, const float* bias,
                  const float* residual, int32_t ldr, float* y, int32_t ldy, int32_t M, int32_t N, int32_t K,
                  int32_t act, int32_t impl, void* stream);
/* impl 2 takes its weights in the tcgen05 tile layout: [ceil(N/128)][ceil(K/64)][128 x 64 bf16, 128-byte swizzle] */
int64_t scv_op_tiled_elems(int32_t N, int32_t K);
int scv_op_pack_tiled(const float* src, uint16_t* dst, int32_t N, int32_t K, void* stream);
/* SplitTile activations (bf16 hi/lo, [m_tile][k_block][hi 16 KB | lo 16 KB], 128-byte swizzle): the form in which
 * LayerNorm / attention / GEMM epilogues hand activations to the next tcgen05 projection. */
int64_t scv_op_split_tile_bytes(int32_t M, int32_t K);
/* LayerNorm (normalize = 1) or plain split (normalize = 0) of fp32 rows into a zero-initialised SplitTile buffer */
int scv_op_split_rows(const float* x, int32_t ldx, const float* gamma, const float* beta, void* out_split, int32_t M,
                      int32_t N, int32_t normalize, void* stream);
/* tcgen05 projection with SplitTile input; output fp32 rows (y) or, when y_split != NULL, SplitTile */
int scv_op_linear_split(const void* a_split, const uint16_t* w_tiled, const float* bias, const float* residual,
                        int32_t ldr, float* y, int32_t ldy, void* y_split, int32_t M, int32_t N, int32_t K, int32_t act,
                        void* stream);
int scv_op_pack_bf16(const float* src, uint16_t* dst, int32_t rows, int32_t cols, int32_t ld_dst, void* stream);
int scv_op_layernorm(const float* x, int32_t ldx, const float* gamma, const float* beta, float* y, int32_t ldy,
                     int32_t M, int32_t N, int32_t act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCVAE_B200_H */
