// Shared helpers for the scvae_b200 engine (sm_100a only).
#pragma once
#ifndef SCV_SPLIT_FP16
#define SCV_SPLIT_FP16 0
#endif
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace scv {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
long long launch_total();

#define SCV_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::scv::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

#define SCV_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::scv::set_error(__VA_ARGS__);           \
      return 1;                                \
    }                                          \
  } while (0)

#define SCV_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

#define SCV_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    ::scv::count_launch();                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) {                                                                   \
      ::scv::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

// Optional per-category kernel timing (CUDA events on the launch stream), used by bench.py's roofline pass.
enum ProfCat : int { PC_LINEAR = 0, PC_LAYERNORM, PC_ATTN_SELF, PC_ATTN_CROSS, PC_SAMPLER, PC_EMBED, PC_MISC,
                     PC_GEMM_TC, PC_NULL /* two events back to back: the overhead one ProfScope adds */, PC_COUNT };
bool prof_enabled();
// Profiling protocol of a step (decoder.cu): prof_step_begin() forgets the previous step's records, the step is enqueued
// behind a gate kernel (every ProfScope adds a record), the gate opens, and after a stream synchronise prof_harvest() adds
// the elapsed time of every record to the category totals.
void prof_step_begin();
void prof_mark_pending();
int prof_harvest();
struct ProfScope {
  ProfScope(int cat, cudaStream_t s, double flops, double bytes);
  ~ProfScope();
  int idx_;
  cudaStream_t s_;
};

// Programmatic dependent launch (PDL): every per-step kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel of the stream may be scheduled while this one
// is still running; griddepcontrol.wait blocks until all prerequisite grids have completed and flushed their
// memory, so everything that reads a predecessor's output comes after pdl_wait().  What runs before it (barrier
// init, TMEM allocation, index math) overlaps the predecessor's tail.  SCV_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
void set_pdl_for_call(bool on);

// Run-time tunables (scv_tune in the C ABI; defaults from SCV_* environment variables read once at load).  A change bumps
// tune_epoch(), which makes every decoder drop its captured step graphs.
struct Tunables {
  int attn_ctas_per_sm;    // SCV_ATTN_CTAS: 0 = one CTA per 8 (row, head) items; > 0 = persistent grid, that many CTAs per SM
  int gemm_stages;         // SCV_GEMM_STAGES: 0 = by grid size (4 stages up to one CTA per SM, else 2 x 2 CTAs per SM); 2 / 4 forces
  int subbatches;          // SCV_SUBBATCHES: 0 = automatic
  int graph;               // SCV_GRAPH: replay one captured CUDA graph per step
  int sub_min_rows;        // SCV_SUB_MIN_ROWS: batches at least this large are decoded as sub-batches on separate streams
  int attn_bulk;           // SCV_ATTN_BULK: cross-attention through the bulk-copy (cp.async.bulk) staged kernel
  int attn_bulk_min_rows;  // SCV_ATTN_BULK_MIN_ROWS: ... for launches with at least this many rows
  int attn_bulk_piece_kb;  // SCV_ATTN_BULK_PIECE_KB: a sequence's block travels as copies of this size (0 = one copy)
  int cond_tc_min_rows;    // SCV_COND_TC_MIN_ROWS: calls with at least this many rows run the small conditioning projections
                           // (stoich / heads / skip memory branches, sc_head.0, family heads) on the tensor cores
  int cluster;             // SCV_CLUSTER: small batches through the cluster-parallel kernel (decode_cluster.cu); opt-in: measured
                           // 0.76-0.92 ms per step against 0.77-0.84 ms for the grid-barrier kernel
  int cluster_max_rows;    // SCV_CLUSTER_MAX_ROWS: ... up to this many rows (<= 64)
  int cluster_rows;        // SCV_CLUSTER_ROWS: rows per cluster (0 = automatic: 1 up to 16 rows, 2 up to 32, else 4)
  int gemm_bn64;           // SCV_GEMM_BN64: projections with at most this many 128-row tiles (0 = never) whose grid of 128 x 64
  int gemm_bn64_max_ctas;  // SCV_GEMM_BN64_MAX_CTAS: tiles has at most this many CTAs use that narrower tile
  int gemm_mc;             // SCV_GEMM_MC: projections over SplitTile input run as clusters of two column tiles that share the A tile:
                           // each CTA loads one half (hi / lo) and multicasts it to both (2/3 of the L2 -> SM operand traffic)
  int gemm_mc_min_row_tiles;   // SCV_GEMM_MC_MIN_ROW_TILES: ... for launches with at least this many 128-row tiles
  int gemm_mc_min_kblocks;     // SCV_GEMM_MC_MIN_KBLOCKS: ... and at least this many 64-wide k-blocks
  int attn_pages_regs;     // SCV_ATTN_PAGES_REGS: self-attention reads a sequence's page ids once into registers (lane l = page l)
  int attn_forward;        // SCV_ATTN_FORWARD: teacher-forced passes stage a (sequence, head)'s K / V in shared memory once
  int attn_forward_min_ctas;   // SCV_ATTN_FORWARD_MIN_CTAS: ... when there are at least this many (sequence, head) pairs
  int attn_shared;         // SCV_ATTN_SHARED: shared memory tokens (RLOO) go through attention_cross_shared_kernel: one warp per
                           // (latent, head) serves all of the latent's samples (0 = the per-row kernel, samples adjacent for L2)
};
Tunables& tun();
unsigned tune_epoch();
int sm_count();            // multiprocessors of the current device (cached per device)
// cudaFuncSetAttribute is per DEVICE: launchers keep one flag per device ordinal and set their attributes the first time
// they run on each (a process-wide flag would leave a second GPU of the same process without the opt-in shared memory).
inline bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// launch_k with thread-block clusters of cluster_x CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, unsigned cluster_x,
                                    Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- CTA residency trace (debug; scv_trace_begin / scv_trace_read): thread 0 of every CTA of the big step kernels logs
// (kernel id, SM, start, end) with %globaltimer, which shows which kernels of the sub-batch streams really share SMs.
struct TraceBuf { unsigned n, cap; unsigned long long pad; };
struct TraceRec { unsigned long long t0, t1; unsigned smid, kid, bx, by; };
void* trace_ptr();         // device TraceBuf* while tracing, else nullptr
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ TraceRec* trace_begin(void* tb, unsigned kid) {
  if (tb == nullptr) return nullptr;
  TraceBuf* b = static_cast<TraceBuf*>(tb);
  const unsigned i = atomicAdd(&b->n, 1u);
  if (i >= b->cap) return nullptr;
  TraceRec* r = reinterpret_cast<TraceRec*>(b + 1) + i;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  r->smid = smid; r->kid = kid; r->bx = blockIdx.x; r->by = blockIdx.y; r->t0 = globaltimer_ns(); r->t1 = 0;
  return r;
}
__device__ __forceinline__ void trace_end(TraceRec* r) { if (r != nullptr) r->t1 = globaltimer_ns(); }

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };

constexpr int kEndIdx = 2;    // END_IDX (models/autoregressive_decoder.py:97)
constexpr int kStartIdx = 1;  // START_IDX (:96)

// erf with one code path (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 + fp32 rounding): the library erff has
// two branches that a warp usually executes both of (~38 instructions per element), and the FFN1 epilogue applies GELU
// to 128 x 128 values per tile with only eight warps (measured: 9.3 of 31 us of that projection).  An absolute error
// of ~2e-7 on erf is the same order as fp32's own rounding of 1 + erf(x) in the reference and 30x below the 2^-17
// relative precision the hi/lo activation split keeps.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  return copysignf(1.0f - p * t * e, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(v);
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

// ACT >= 0: the activation is known at compile time; ACT < 0: decided by `act` at run time.
template <int ACT>
__device__ __forceinline__ float apply_act_t(float v, int act) {
  if constexpr (ACT == ACT_NONE) return v;
  else if constexpr (ACT == ACT_GELU) return gelu_erf(v);
  else if constexpr (ACT == ACT_RELU) return v > 0.f ? v : 0.f;
  else if constexpr (ACT == ACT_SIGMOID) return sigmoidf_(v);
  else return apply_act(v, act);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }

// Two fp32 values -> packed 16-bit pairs hi = r16(x) and lo = r16(x - hi) (element 0 in the low half), the operands of the
// two MMAs the tensor-core projections issue per weight tile.
//   bf16 (default): hi + lo carries 16 mantissa bits of x; products against exact-bf16 weights are exact in fp32.
//   kSplitFp16 (-DSCV_SPLIT_FP16=1): r16 = fp16, 22 mantissa bits of x, weight tiles stored as fp16 too (the bf16-rounded
//     weights are exact in fp16 down to 2^-17 in magnitude), |x| > 65504 saturates.  Measured (profiles/split_error_r02.log):
//     the full GPU suite passes and the error against fp64 halves at K = 512 (rms 2.7e-6 -> 1.25e-6 on O(1) outputs) but
//     stays 5e-6 at K = 2048: from there on the error is the tensor core's own fp32 accumulation (hundreds of truncating
//     accumulate steps; an fp32 FMA loop has 3e-7), which no split can remove.  Not worth the range limit: off.
//   (A bf16 hi with an fp16 residual would keep the range, but the hardware rejects an f16 x bf16 MMA for kind::f16 even
//   though the instruction descriptor has separate A / B format fields: "illegal instruction" on B200.)
constexpr bool kSplitFp16 = SCV_SPLIT_FP16 != 0;
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  if constexpr (kSplitFp16) {
    const __half ha = __float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f)), hb = __float2half_rn(fminf(fmaxf(b, -65504.f), 65504.f));
    const __half2 h = __halves2half2(ha, hb);
    const __half2 l = __floats2half2_rn(a - __half2float(ha), b - __half2float(hb));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const __nv_bfloat162 h = __halves2bfloat162(ha, hb);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __bfloat162float(ha), b - __bfloat162float(hb));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG for the multinomial sampler ----
struct Philox {
  __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ static inline void run(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
  }
  // uniform in [0, 1) with 24 random bits
  __device__ static inline float uniform(uint64_t seed, uint64_t offset, uint32_t row, uint32_t step) {
    uint32_t c[4] = {step, row, (uint32_t)offset, (uint32_t)(offset >> 32)};
    run(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (float)(c[0] >> 8) * (1.0f / 16777216.0f);
  }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---------------- kernel launchers shared between translation units ----------------
struct LinearArgs {
  const float* x = nullptr; int ldx = 0;
  const void* a_split = nullptr;                   // activations already in SplitTile form (tcgen05 path only)
  const __nv_bfloat16* w = nullptr; int ldw = 0;   // [N, ldw] bf16, zero padded beyond K
  const __nv_bfloat16* wt = nullptr;               // same weights, tiled + swizzled for tcgen05 (or null)
  const float* bias = nullptr;
  const float* residual = nullptr; int ldr = 0;    // may alias y
  float* y = nullptr; int ldy = 0;
  void* y_split = nullptr;                         // write the output in SplitTile form instead of fp32 rows
  int M = 0, N = 0, K = 0;
  int act = ACT_NONE;
  const int* done_flag = nullptr;                  // &StepState::done: skip the work when it is set; with compaction of
                                                   // finished rows (done_flag[6] = StepState::pad[1] > 0) row tiles at or
                                                   // beyond that many rows are skipped as well
  int row_base = 0;                                // first row of this launch inside the call's batch (sub-batches)
  const void* next_w = nullptr; size_t next_w_bytes = 0;   // tiled weights of the NEXT projection: prefetched into L2
  // LayerNorm(y) written as a SplitTile by the same kernel (gemm_tcgen05_ln.cu; callers check tc_res_ln_ok first)
  const float* ln_gamma = nullptr; const float* ln_beta = nullptr; void* ln_out_split = nullptr;
  // *nonfinite_flag |= 1 when an output (fp32 rows only) is NaN / +-inf (mode 1) or NaN (mode 2); null: no check.  Lets the
  // sampling path learn "some adjusted logit of the batch is not finite" (reference :1464-1466) from the kernels that
  // produce the logits and the stop logit instead of a separate pass over all logits.
  int* nonfinite_flag = nullptr; int nonfinite_mode = 0;
};
__device__ __forceinline__ bool nonfinite_hit(float v, int mode) { return mode == 2 ? isnan(v) : (isnan(v) || isinf(v)); }
int launch_linear_simt(const LinearArgs& a, cudaStream_t s);
int launch_linear_tcgen05(const LinearArgs& a, cudaStream_t s);
bool tc_persistent_ok(const LinearArgs& a);
int launch_linear_tcgen05_persistent(const LinearArgs& a, cudaStream_t s);
// CTA-pair variant for wide projections (gemm_tcgen05_2cta.cu, SCV_GEMM_2CTA)
bool tc_2cta_ok(const LinearArgs& a);
int launch_linear_tcgen05_2cta(const LinearArgs& a, cudaStream_t s);
bool tc_shape_ok(const LinearArgs& a);
bool tc_res_ln_ok(const LinearArgs& a);
int launch_linear_res_ln(const LinearArgs& a, cudaStream_t s);
size_t tc_packed_elems(int N, int K);
size_t split_tile_bytes(int M, int K);             // bytes of a SplitTile buffer for an [M, K] activation
// LayerNorm (or, with normalize = 0, a plain copy) of fp32 rows straight into SplitTile form
int launch_layernorm_split(const float* x, int ldx, const float* gamma, const float* beta, void* out_split, int M,
                           int N, int normalize, const int* done_flag, cudaStream_t s, int act = ACT_NONE);
int launch_pack_tiled(const float* src, __nv_bfloat16* dst, int N, int K, cudaStream_t s);
int launch_linear(const LinearArgs& a, int impl, cudaStream_t s);   // impl 0 auto, 1 simt, 2 tcgen05

int launch_pack_bf16(const float* src, __nv_bfloat16* dst, int rows, int cols, int ld_dst, cudaStream_t s);
int launch_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t s);
int launch_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy, int M,
                     int N, int act, const int* done_flag, cudaStream_t s);

}  // namespace scv
