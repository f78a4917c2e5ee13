// Device-side state and launchers of the per-step decode kernels.
#pragma once
#include "common.cuh"

namespace scv {

constexpr int kPageShift = 4;               // 16 cached positions per KV page
constexpr int kPagePos = 1 << kPageShift;

// Lives in device memory; every per-step kernel reads it instead of taking the position as a launch
// argument, so the same launch sequence (or CUDA graph) serves every step and the host never has to
// synchronise inside the loop (the reference syncs 3-5 times per step: finished.all(), .any() ...).
struct StepState {
  int step;            // position being decoded (range(max_len - 1), autoregressive_decoder.py:1403)
  int done;            // set once every row has emitted END (:1547) or the step budget is used up
  int n_unfinished;    // rows that have not emitted END yet
  int out_len;         // number of executed steps L
  int degenerate;      // batch-global "logits contain nan/inf" flag of the current step (:1464-1466)
  int next_free_page;  // bump allocator over the KV page pool
  int pad[2];          // pad[0]: 1 + last step at which a row emitted END (decode_cluster.cu); pad[1]: rows still being decoded
                       // when finished rows are compacted away (0 = no compaction): kernels skip slots >= pad[1]
  unsigned long long seed, offset;   // Philox key / counter offset of this call (kept out of the kernel arguments)
};

struct EmbedArgs {
  const __nv_bfloat16* table; int ld_table;     // token_embedding.weight [V, d] (bf16, padded rows)
  const float* pe; int d;                        // pos_encoding.pe [pe_len, d]
  const int* cur_tokens;                         // [B]
  float* x; int B;
  int* page_table; int pages_per_seq;
  StepState* st;
  const int* row_map = nullptr; int slot_base = 0;   // compaction: slot -> row of the call's batch (page_table is indexed by row)
};
int launch_embed(const EmbedArgs& a, cudaStream_t s);
// x[b * L + t] = token_embedding[tokens[b, t]] + pe[t] for every position (teacher-forced forward, :947-949)
int launch_embed_sequence(const __nv_bfloat16* table, int ld_table, const float* pe, int d, const long long* tokens,
                          int ld_tokens, int B, int L, int vocab, float* x, unsigned char* key_skip, cudaStream_t s);

struct AttnArgs {
  const float* q = nullptr; int ldq = 0;
  const float* knew = nullptr; const float* vnew = nullptr; int ldn = 0;   // self-attention: row to append
  float* kcache = nullptr; float* vcache = nullptr;     // layer / k-v offsets already applied
  const int* page_table = nullptr; int pages_per_seq = 0; long long page_stride = 0;   // paged (self)
  long long seq_stride = 0; int row_stride = 0;          // contiguous (cross): b*seq_stride + p*row_stride
  float* out = nullptr; int ldo = 0;
  unsigned char* out_split = nullptr; int kb_out = 0;     // SplitTile output (bf16 hi/lo) for the tcgen05 out-projection
  int B = 0, nhead = 0, hd = 0; float scale = 1.f;
  int fixed_len = -1;                                    // cross: memory tokens; self: -1 -> step + 1
  int max_n = 0;                                         // smem scores per warp
  int host_len_hint = 0;                                 // host's view of step + 1 (profiling byte counts only)
  // Teacher-forced forward (all positions at once): query row r belongs to sequence r / rows_per_seq, position
  // r % rows_per_seq; it attends over positions <= its own (self) of that sequence's contiguous K/V, skipping keys
  // whose key_skip[sequence, position] byte is set (tgt_key_padding_mask, :952).  0 = one row per sequence (decode).
  int rows_per_seq = 0;
  const unsigned char* key_skip = nullptr;
  const StepState* st = nullptr;                         // null: no done flag / step counter (forward)
  void* trace = nullptr;                                 // CTA residency trace (common.cuh), normally null
  // compaction of finished rows: query row b is slot slot_base + b; its sequence (page table row, projected memory) is
  // row_map[slot]; page_table / kcache / vcache are then NOT offset to the sub-batch
  const int* row_map = nullptr; int slot_base = 0;
  // shared memory tokens (RLOO): the sequence of query row r is (slot_base + r) % seq_mod (0 = off); the launch's rows
  // are walked so that the samples of one latent are adjacent (they then hit the same K / V in L2)
  int seq_mod = 0;
  int pages_regs = 1;                                    // set by the launcher (tunable attn_pages_regs)
};
int launch_attention(const AttnArgs& a, cudaStream_t s);

// ---- small-batch persistent step (decode_small.cu): the step as a list of phases separated by grid barriers
struct SmallOp {               // y = act(LN?(x) W^T + b) (+ residual) for all B rows
  const float* in; int ld_in; int K;
  const float* ln_g; const float* ln_b;          // LayerNorm of the input rows (null: none)
  const __nv_bfloat16* w; int ldw;               // [N, ldw] row-major bf16, zero padded beyond K, ldw % 8 == 0
  const float* bias; int N; int act;
  int cpc;                                        // output columns per CTA (small_cols_per_cta)
  const float* res; int ldr;                      // residual added after the activation (may alias out)
  float* out; int ldo;
};
struct SmallPhase {
  int kind;                                       // 0: up to four independent projections, 1: attention
  int nops;
  SmallOp op[4];
  AttnArgs attn;
};
constexpr int kSmallMaxRows = 64;                // what the kernel can take (row groups of 32); the default limit is 32
size_t small_step_smem_bytes();
int small_cols_per_cta(int N, int grid);
bool small_phase_fits(const SmallPhase& ph, int grid);
int launch_decode_small(const SmallPhase* phases_dev, int n_phases, int B, const StepState* st, unsigned* bar, int grid,
                        cudaStream_t s);

struct SamplerArgs {
  const float* logits = nullptr; int ldl = 0;
  const float* type_logits = nullptr; int ldt = 0;
  const float* stop_logits = nullptr;
  const uint8_t* type_masks = nullptr;
  int B = 0, V = 0, max_len = 0;
  float temperature = 1.f; int top_k = 0; float top_p = 1.f;
  int sort_n = 0;                                        // set by the launcher: padded vocabulary for top-k/top-p
  float stop_boost = 0.f, hard_stop = 0.f;
  // site-duplication gating (reference :1424-1435, :1525-1539): seen[b, v] marks "element" ids already emitted
  const float* dup_logits = nullptr; float dup_threshold = 0.f; unsigned char* seen = nullptr;
  int want_logprobs = 0, want_entropy = 0; unsigned flags = 0;
  int row_base = 0;                                      // first row of this sub-batch (Philox counter)
  long long* out_tokens = nullptr; float* out_logprobs = nullptr; float* out_entropy = nullptr; int out_ld = 0;
  int* cur_tokens = nullptr; unsigned char* finished = nullptr;
  const long long* forced = nullptr;
  StepState* st = nullptr;
  // compaction of finished rows: logits / head outputs / cur_tokens are indexed by slot, everything else by row_map[slot]
  const int* row_map = nullptr; int slot_base = 0;
};
// which = 1: first kernel (stages logits, publishes the H2 flag, picks the token when the call is plain greedy);
// which = 2: second kernel (entropy / temperature sampling / log-prob), only when sampling or entropy is requested
int launch_sampler(const SamplerArgs& a, int which, cudaStream_t s);
int launch_step_end(StepState* st, int max_steps, cudaStream_t s);
// stable compaction of the slots whose row has not emitted END (row_map, cur_tokens in place); st->pad[1] = rows left
int launch_compact_rows(int* row_map, int* cur_tokens, const unsigned char* finished, StepState* st, int B, cudaStream_t s);
// one thread that spins until *host_flag (pinned host memory) becomes non-zero: holds a stream while the host enqueues
int launch_host_gate(int* host_flag, cudaStream_t s);
// true when launch_sampler(a, 1, ...) is the warp-per-row greedy kernel and nothing else (no second sampler kernel)
bool sampler_plain_greedy(const SamplerArgs& a);

// Whole-decode persistent kernel (decode_small.cu): what the step loop needs besides the phase list
struct SmallTail {
  SamplerArgs sp;                                        // greedy sampling epilogue (sp.st = the call's StepState)
  const __nv_bfloat16* emb; int ld_emb;                  // token_embedding.weight [V, ld_emb]
  const float* pe; int d; float* x;                      // pos_encoding.pe [pe_len, d]; residual stream [B, d]
  int* page_table; int pages_per_seq;
  int max_steps;
};
int launch_decode_small_persist(const SmallPhase* phases_dev, int n_phases, int B, unsigned* bar, int grid, const SmallTail& tail,
                                cudaStream_t s);
// ---- cluster-parallel small-batch decode (decode_cluster.cu): the batch's rows are dealt to thread-block clusters of
// kClSize CTAs; a cluster runs the whole layer stack for its rows as a tensor-parallel group (every CTA streams 1 / kClSize
// of every weight matrix with cp.async.bulk, activations are exchanged through distributed shared memory, phases are
// separated by mbarrier-based CLUSTER barriers, never by a grid barrier).  The step is described by a host-built program.
constexpr int kClSize = 8;                  // CTAs per cluster (portable maximum)
constexpr int kClMaxRows = 4;               // rows per cluster
enum ClBuf : int { CB_X = 0, CB_XN, CB_XN2, CB_A, CB_B, CB_H, CB_C, CB_D, CB_E, CB_Q, CB_K, CB_V, CB_COUNT };
enum ClKind : int { CL_LN = 0, CL_GEMV, CL_ATTN_SELF, CL_ATTN_CROSS, CL_SYNC };
enum ClOut : int { CO_LOCAL = 0, CO_GATHER, CO_GLOBAL };
struct ClInstr {                             // 64 bytes: the whole program of a step sits in shared memory
  int kind; int act;
  union {
    // CL_GEMV: out[r][n] = act(in[r] . w[n] + bias[n]) (+ x[r][n]);  split = 1: CTA `rank` computes columns
    // [rank * N / kClSize, ...), 0: rank 0 computes all N;  CO_LOCAL: out_buf[r][n - n0], CO_GATHER: out_buf[r][n] in EVERY
    // CTA of the cluster, CO_GLOBAL: out_global[row * out_ld + n]
    struct { const __nv_bfloat16* w; const float* bias; float* out_global; int ldw, K, N, out_ld;
             short split, in_buf, residual, out_kind, out_buf, pad0, pad1, pad2; } g;
    // CL_LN: dst = LayerNorm(src) * gamma + beta over n columns (every CTA normalises its own copy of the rows)
    struct { const float* gamma; const float* beta; int src_buf, dst_buf, n; } ln;
    // CL_ATTN_*: K / V of this layer (paged self-attention cache, or the projected memory tokens)
    struct { float* kcache; float* vcache; long long page_stride; long long seq_stride; int row_stride, fixed_len; } at;
  };
};
static_assert(sizeof(ClInstr) == 64, "ClInstr is copied into shared memory as 64-byte records");
struct ClProgram {
  const ClInstr* instr; int n_instr;         // device array
  int B, d, nhead, hd, dff, V, pe_len;
  int rows_per_cluster;                      // R: 1, 2 or 4
  int cpl;                                   // 16-byte weight chunks per lane and K segment: 2 (segments <= 512 wide) or 3 (<= 768)
  int buf_off[CB_COUNT]; int buf_ld[CB_COUNT]; int buf_floats;     // shared-memory layout of the activation buffers (floats)
  const int* page_table; int pages_per_seq;
  float scale;
  float* x_global;                           // [B, d] residual stream: read at kernel entry (step mode / first step), written back
  StepState* st;
  int exp;                                   // SCV_CLUSTER_EXP (timing experiments, wrong results): 1 stream only, 2 no epilogue
  unsigned long long* dbg;                   // SCV_CLUSTER_DEBUG: per-instruction nanoseconds of cluster 0 / rank 0 / thread 0
};
size_t cluster_smem_bytes(const ClProgram& p);
int cluster_max_active(int rows_per_cluster, size_t smem);      // clusters of the kernel that can be resident at once
// true when the decoder shape can run on the cluster kernel (head / column counts divisible by the cluster size, ...)
bool cluster_shape_ok(int d, int nhead, int dff, int V, int pe_len, int n_memory);
// one decode step (sampling calls: the sampler kernels follow) / the whole plain-greedy decode in one launch
int launch_decode_cluster_step(const ClProgram& p, cudaStream_t s);
int launch_decode_cluster_persist(const ClProgram& p, const SmallTail& tail, cudaStream_t s);

int launch_init_rows(int* cur_tokens, unsigned char* finished, int B, StepState* st, unsigned long long seed,
                     unsigned long long offset, cudaStream_t s, int* row_map = nullptr);

}  // namespace scv
