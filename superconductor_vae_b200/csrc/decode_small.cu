// Small-batch decode (<= 32 rows by default, <= 64 with SCV_SMALL_MAX_ROWS): ONE persistent cooperative kernel for the whole decode (plain greedy calls) or per
// step (sampling calls) instead of ~140 launches per step.
//
// At a few dozen rows a decode step is bound by streaming the 107 MB of bf16 weights once (16 us of HBM
// time) but the per-projection kernels spend ~7 us each on launch, prologue and drain: 1.17 ms per step.  Here one CTA
// per SM (16 warps) stays resident and walks a list of phases separated by grid-wide barriers:
//   projection phase: up to four projections of the same dependency level (y = act(LN?(x) W^T + b) (+ residual)).  A
//                 CTA owns 8-48 contiguous output columns; per group of 32 rows it reads the input rows ([32, K] fp32)
//                 through L2 into registers (a warp = two rows, every load in flight at once), applies the LayerNorm
//                 there, writes them as three bf16 terms (x = hi + mid + lo exactly) into shared memory
//                 and multiplies with mma.sync.m16n8k16 (warp = row tile x k-group, partial sums meet in shared
//                 memory), then bias / activation / residual.  The CTA's weight rows and LayerNorm weights are copied
//                 with cp.async BEFORE the barrier that precedes the phase: they do not depend on activations.
//   feed-forward: (opt-in, SCV_SMALL_FUSE_FFN=1) linear1 (+ LayerNorm, GELU) as a projection phase on 16 hidden units per CTA whose epilogue keeps the
//                 hidden activations in shared memory (three bf16 terms again) and multiplies them with the matching
//                 16 columns of linear2: a [32, d] slab of partial sums per CTA; the next phase adds up the slabs
//                 (fixed order) + bias + residual.  The [32, dff] activations never travel.
//   attention   : one warp per (row, head); cached K / V rows are copied before the barrier too, the step's new row is
//                 used from registers.
//   sampling    : (whole-decode kernel) greedy epilogue, END bookkeeping, embedding of the chosen token.
// Arithmetic: bf16 weights, activations as three bf16 terms (all 24 significant bits: every product exact), fp32
// accumulate; LayerNorm, softmax,
// residual stream and KV cache fp32.
//
// Measured on B200: 0.80-0.85 ms per step at 8-32 rows against 1.17-1.20 ms for the per-projection
// path, 2048 of 2048 rows token-identical to the fp32 oracle when decoded 32 at a time (tests/kv_probe.py).  A
// step is ~99 dependent phases; a barrier costs ~2 us, staging 1-2 us, MMA + epilogue ~1.2 us.  Per-phase times of
// the per-step kernel are printed with SCV_GRAPH=0 SCV_SMALL_PERSIST=0 SCV_SMALL_DEBUG=<launch index>.
// Reference call sites: models/autoregressive_decoder.py:1244-1313 (layer), :1413-1441 (heads), :1505-1548 (sampling).
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "decode_kernels.cuh"
#include "sampler_device.cuh"

namespace scv {

namespace {

constexpr int SM_THREADS = 512, SM_WARPS = 16;   // 4 warps per scheduler: the staging is issue / latency bound
constexpr int KQ = SM_WARPS / 2;              // k-groups of the MMA loop (warp = row tile x k-group)
constexpr int RPW = 32 / SM_WARPS;            // staged rows per warp
constexpr int KC = 512;                       // widest input staged in one piece (a LayerNorm input must be: d_model <= 512)
constexpr int KCH = 512;                      // wider inputs (the feed-forward width) are staged in chunks of 512 columns
constexpr int A_PITCH = KC + 8;               // bf16 per staged row: 1040 B, consecutive rows 16 B apart in the banks
constexpr int W_BUF_BYTES = 56 * 1024;        // weight rows of one phase for one CTA
constexpr int W_ROW_PAD = 16;                 // bytes added to every staged weight row (same bank rotation as A)
constexpr int NT_MAX = 6;                     // 8-column MMA tiles per CTA and projection (<= 48 output columns)
constexpr int RED_PITCH = NT_MAX * 8;         // floats per row of the cross-warp partial sums
constexpr int FF2_PITCH = 48;                 // bytes per staged linear2 row slice (16 hidden units = 32 B + pad)
constexpr int FFH_PITCH = 24;                 // bf16 per row of the staged hidden activations (16 + pad: 48 B)
constexpr int FFH_OFFSET = 64 * 1024;         // byte offset of the hidden-activation terms inside the a_hi.. region
constexpr int MAX_N_SCORES = 256;             // positions per (row, head) the score buffer can hold

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D += A (16x16 bf16, row) * B (16x8 bf16, col), fp32 accumulate
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Grid-wide barrier over co-resident CTAs (the launch is cooperative: one CTA per SM).  The arrivals are spread over
// kBarShards counters in separate 128-byte lines (atomics on one address are serialised by the L2); CTA b adds to
// counter b % kBarShards and lanes 0 .. kBarShards-1 of warp 0 poll one counter each.  `bar` is zeroed by a memset
// before the launch; the counters only grow (epoch * CTAs of the shard).  Measured at 32 rows: 1, 4 and 8 shards are
// within 1 % of each other (782 / 781 / 789 us per step), so the default is one counter (SCV_SMALL_BAR_SHARDS): the
// barrier's ~2 us are the release fence plus two L2 round trips, not the serialised atomics.  One release-store flag
// per CTA polled by every CTA was also tried: 148 x 148 acquire loads per poll round on five lines, 1190 us per step.
constexpr int kBarWords = 256;                    // 8 counters, 32 words (128 B) apart
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch, int shards) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x < 32) {
    const int lane = (int)threadIdx.x;
    if (lane == 0)                                // release: the CTA's writes (ordered before this by the barrier above)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 32 * ((int)blockIdx.x % shards)) : "memory");
    const unsigned members = lane < shards ? ((unsigned)gridDim.x - (unsigned)lane + (unsigned)shards - 1u) / (unsigned)shards : 0u;
    const unsigned want = members * epoch;
    for (;;) {
      const bool ok = lane >= shards || ld_acquire_u32(bar + 32 * lane) >= want;
      if (__all_sync(0xffffffffu, ok)) break;
    }
  }
  __syncthreads();
}

struct Smem {
  SmallPhase ph[2];                               // descriptors of the running and of the next phase (copied from global)
  // the staged input rows as three bf16 terms, x = hi + mid + lo EXACTLY (3 x 8 significant bits = the 24 of an fp32):
  // with bf16 weights every product is exact and only the fp32 accumulation rounds, like the fp32 CUDA-core path.
  // (Two terms, the split of the tcgen05 path, measured 2046 of 2048 rows token-identical to the fp32 oracle when
  // decoded 32 at a time, tests/kv_probe.py; the per-projection path these batch sizes used before is fp32-exact.)
  // a_hi doubles as the cross-warp partial-sum buffer of a projection and as the score buffer of an attention phase
  __align__(16) __nv_bfloat16 a_hi[32 * A_PITCH];
  __align__(16) __nv_bfloat16 a_mid[32 * A_PITCH];
  __align__(16) __nv_bfloat16 a_lo[32 * A_PITCH];
  __align__(16) unsigned char w[2][W_BUF_BYTES];
};
static_assert(KQ * 32 * RED_PITCH * sizeof(float) <= 3 * 32 * A_PITCH * sizeof(__nv_bfloat16), "partial sums alias a_hi .. a_lo");
static_assert(SM_WARPS * MAX_N_SCORES * sizeof(float) <= 32 * A_PITCH * sizeof(__nv_bfloat16), "scores alias a_hi");
static_assert(KQ * 32 * RED_PITCH * sizeof(float) <= FFH_OFFSET && FFH_OFFSET + 3 * 32 * FFH_PITCH * 2 <= 3 * 32 * A_PITCH * 2, "hidden terms sit behind the partial sums");

// Output columns are dealt to the CTAs in contiguous blocks of op.cpc (a multiple of 8 = whole MMA column tiles, see
// small_cols_per_cta): CTA c owns [c * cpc, c * cpc + nc); CTAs beyond N / cpc skip the projection (and its staging).
__device__ __forceinline__ void cols_of_cta(const SmallOp& op, int& n0, int& nc) {
  n0 = (int)blockIdx.x * op.cpc;
  nc = max(0, min(op.cpc, op.N - n0));
}

__device__ void prefetch_attention(const AttnArgs& a, Smem& sm, int buf, int step);

// Issue the cp.async copies of this CTA's weight rows of every op of `ph` into buffer `buf`.
__device__ void prefetch_weights(const SmallPhase& ph, Smem& sm, int buf, int step) {
  if (ph.kind == 1) { prefetch_attention(ph.attn, sm, buf, step); return; }
  if (ph.kind == 3) { cp_async_commit(); return; }             // partial-sum reduction: nothing to stage
  unsigned char* dst = sm.w[buf];
  const int nops = ph.kind == 2 ? 1 : ph.nops;                  // fused feed-forward: op[0] = linear1, op[1] = linear2
  for (int o = 0; o < nops; ++o) {
    const SmallOp& op = ph.op[o];
    int n0, nc;
    cols_of_cta(op, n0, nc);
    const int row_bytes = op.ldw * 2, chunks = row_bytes / 16, wp = row_bytes + W_ROW_PAD;
    for (int i = threadIdx.x; i < nc * chunks; i += SM_THREADS) {
      const int j = i / chunks, c = i - j * chunks;
      cp_async16(dst + (size_t)j * wp + 16 * c, reinterpret_cast<const unsigned char*>(op.w + (size_t)(n0 + j) * op.ldw) + 16 * c);
    }
    dst += (size_t)nc * wp;
    if (op.ln_g != nullptr && nc > 0) {           // LayerNorm weight and bias of the input rows: K floats each
      const int q4 = op.K >> 2;
      for (int i = threadIdx.x; i < 2 * q4; i += SM_THREADS)
        cp_async16(dst + 16 * i, (i < q4 ? op.ln_g : op.ln_b) + 4 * (i < q4 ? i : i - q4));
      dst += (size_t)op.K * 8;
    }
  }
  if (ph.kind == 2) {
    // linear2's weights for this CTA's hidden units: W2[n, h0 .. h0 + 15] = 32 contiguous bytes of every output row n
    // (the k-contiguous "col" operand of the second MMA), staged at a 48-byte pitch (bank rotation)
    const SmallOp& f1 = ph.op[0];
    const SmallOp& f2 = ph.op[1];
    int h0, nh;
    cols_of_cta(f1, h0, nh);
    if (nh > 0)
      for (int i = threadIdx.x; i < 2 * f2.N; i += SM_THREADS) {
        const int n = i >> 1, c = i & 1;
        cp_async16(dst + (size_t)n * FF2_PITCH + 16 * c, reinterpret_cast<const unsigned char*>(f2.w + (size_t)n * f2.ldw + h0) + 16 * c);
      }
  }
  cp_async_commit();
}

__device__ __forceinline__ uint32_t bf162_bits(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// Stage columns [k0, k0 + kc) of the B input rows as bf16 hi / mid / lo, LayerNorm applied on the way (then the chunk is the
// whole row).  Warp w owns rows RPW * w .. RPW * w + RPW - 1, R rows at a time, a row spread over the lanes as NJ float4s; every load of
// a round (and the LayerNorm weights) is in flight at once: the input was written by other CTAs in the previous phase
// and comes through L2, so a round costs one L2 round trip.
template <int NJ, int R>
__device__ __forceinline__ void stage_rows_t(const SmallOp& op, Smem& sm, int k0, int kc, int rg, int B, const float* gb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool ln = op.ln_g != nullptr;
#pragma unroll 1
  for (int rr = 0; rr < RPW; rr += R) {
    float4 v[R][NJ];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int r = warp * RPW + rr + u;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int col = 4 * (lane + 32 * j);
        v[u][j] = (rg + r < B && col < kc) ? __ldcg(reinterpret_cast<const float4*>(op.in + (size_t)(rg + r) * op.ld_in + k0 + col))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (ln) {                                     // two passes over the registers, eps 1e-5 (nn.LayerNorm)
      float s[R], q[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        s[u] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) s[u] += (v[u][j].x + v[u][j].y) + (v[u][j].z + v[u][j].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < R; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
      for (int u = 0; u < R; ++u) {
        s[u] = s[u] / (float)kc;                  // mean
        q[u] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (4 * (lane + 32 * j) < kc) {
            const float d0 = v[u][j].x - s[u], d1 = v[u][j].y - s[u], d2 = v[u][j].z - s[u], d3 = v[u][j].w - s[u];
            q[u] = fmaf(d0, d0, q[u]); q[u] = fmaf(d1, d1, q[u]); q[u] = fmaf(d2, d2, q[u]); q[u] = fmaf(d3, d3, q[u]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < R; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
#pragma unroll
      for (int u = 0; u < R; ++u) q[u] = 1.0f / sqrtf(q[u] / (float)kc + 1e-5f);      // rstd
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int col = 4 * (lane + 32 * j);
        if (col < kc) {
          const float4 g = *reinterpret_cast<const float4*>(gb + col), bt = *reinterpret_cast<const float4*>(gb + kc + col);
#pragma unroll
          for (int u = 0; u < R; ++u) {
            const float mean = s[u], rstd = q[u];
            v[u][j].x = (v[u][j].x - mean) * rstd * g.x + bt.x; v[u][j].y = (v[u][j].y - mean) * rstd * g.y + bt.y;
            v[u][j].z = (v[u][j].z - mean) * rstd * g.z + bt.z; v[u][j].w = (v[u][j].w - mean) * rstd * g.w + bt.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int r = warp * RPW + rr + u;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int col = 4 * (lane + 32 * j);
        if (col < kc) {
          const float4 x = v[u][j];
          // hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid) (both differences exact in fp32): packed converts
          const uint32_t h01 = bf162_bits(__floats2bfloat162_rn(x.x, x.y)), h23 = bf162_bits(__floats2bfloat162_rn(x.z, x.w));
          const float r0 = x.x - __uint_as_float(h01 << 16), r1 = x.y - __uint_as_float(h01 & 0xffff0000u);
          const float r2 = x.z - __uint_as_float(h23 << 16), r3 = x.w - __uint_as_float(h23 & 0xffff0000u);
          const uint32_t m01 = bf162_bits(__floats2bfloat162_rn(r0, r1)), m23 = bf162_bits(__floats2bfloat162_rn(r2, r3));
          const uint32_t l01 = bf162_bits(__floats2bfloat162_rn(r0 - __uint_as_float(m01 << 16), r1 - __uint_as_float(m01 & 0xffff0000u)));
          const uint32_t l23 = bf162_bits(__floats2bfloat162_rn(r2 - __uint_as_float(m23 << 16), r3 - __uint_as_float(m23 & 0xffff0000u)));
          *reinterpret_cast<uint2*>(sm.a_hi + r * A_PITCH + col) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(sm.a_mid + r * A_PITCH + col) = make_uint2(m01, m23);
          *reinterpret_cast<uint2*>(sm.a_lo + r * A_PITCH + col) = make_uint2(l01, l23);
        }
      }
    }
  }
}

__device__ __forceinline__ void stage_rows(const SmallOp& op, Smem& sm, int k0, int kc, int rg, int B, const float* gb) {
  stage_rows_t<KC / 128, RPW>(op, sm, k0, kc, rg, B, gb);        // kc <= KC = 512: the warp's rows in one round
}

// One projection phase.  Per op: the CTA's nc <= 48 output columns for all 32 (padded) rows as mma.sync m16n8k16 tiles,
// A = the staged rows (hi, mid and lo against the same weight fragment, one accumulator), B = the CTA's weight rows
// [nc, K] (already k-contiguous = "col" operand).  Warp = (row tile, quarter of the k-steps); the four partial sums per
// output meet in shared memory, then bias / activation / residual and nc contiguous floats per row go to global.
__device__ void run_gemv_phase(const SmallPhase& ph, Smem& sm, int buf, int B, unsigned long long* tdbg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 1, kq = warp >> 1;      // row tile, k-group
  const unsigned char* wbase = sm.w[buf];
  float* red = reinterpret_cast<float*>(sm.a_hi);
  const bool fused = ph.kind == 2;                // feed-forward block: linear1 + GELU here, linear2 as partial sums
  for (int o = 0; o < (fused ? 1 : ph.nops); ++o) {
    const SmallOp op = ph.op[o];                  // by value: the descriptor lives in shared memory, the hot fields in registers
    int n0, nc;
    cols_of_cta(op, n0, nc);
    if (nc == 0) continue;                        // CTA-uniform
    const int K = op.K, wp = op.ldw * 2 + W_ROW_PAD;
    const int ntiles = (nc + 7) >> 3;
    for (int rg = 0; rg < B; rg += 32) {            // groups of 32 rows: the staged weights serve every group
      // this thread's outputs of the epilogue: fetch bias and residual now, their latency hides behind the staging
      constexpr int EP = (32 * NT_MAX * 8 + SM_THREADS - 1) / SM_THREADS;
      float bias_v[EP], res_v[EP];
#pragma unroll
      for (int e = 0; e < EP; ++e) {
        const int idx = threadIdx.x + e * SM_THREADS;
        const int r = idx / nc, c = idx - r * nc;
        bias_v[e] = 0.f; res_v[e] = 0.f;
        if (r < 32 && rg + r < B) {
          if (op.bias != nullptr) bias_v[e] = __ldg(op.bias + n0 + c);
          if (op.res != nullptr) res_v[e] = __ldcg(op.res + (size_t)(rg + r) * op.ldr + n0 + c);
        }
      }
      float acc[NT_MAX][4];                         // the three terms' products of a k-step go to the same accumulator
#pragma unroll
      for (int j = 0; j < NT_MAX; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
      const float* gb = reinterpret_cast<const float*>(wbase + (size_t)nc * wp);   // staged LayerNorm weight | bias
      // ldmatrix row addresses: A x4 = (rows 0-7 | 8-15) x (k 0-7 | 8-15); B x4 = (tile j | j+1) x (k 0-7 | 8-15)
      const int mi = lane >> 3;
      const int a_row = mt * 16 + (lane & 7) + (mi & 1) * 8, a_kofs = (mi >> 1) * 8;
      const int b_kofs = (mi & 1) * 8;
      const int chunk = K > KC ? KCH : KC;
      for (int k0 = 0; k0 < K; k0 += chunk) {
        const int kc = min(chunk, K - k0);
        cp_async_wait_all();                        // this phase's weight rows and LayerNorm weights, copied before the barrier ...
        __syncthreads();                            // ... by every thread; the previous chunk / op / phase is done with a_hi, a_lo
        if (tdbg && threadIdx.x == 0 && o == 0 && k0 == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tdbg[0]));
        stage_rows(op, sm, k0, kc, rg, B, gb);
        __syncthreads();                            // the staged rows are complete
        if (tdbg && threadIdx.x == 0 && o == 0 && k0 == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tdbg[1]));
        const int steps = kc >> 4;
        for (int ks = kq; ks < steps; ks += KQ) {
          uint32_t ah[4], am[4], al[4];
          ldmatrix_x4(ah, sm.a_hi + a_row * A_PITCH + ks * 16 + a_kofs);
          ldmatrix_x4(am, sm.a_mid + a_row * A_PITCH + ks * 16 + a_kofs);
          ldmatrix_x4(al, sm.a_lo + a_row * A_PITCH + ks * 16 + a_kofs);
#pragma unroll
          for (int jp = 0; jp < NT_MAX / 2; ++jp) {
            if (2 * jp < ntiles) {
              const int wrow = min((2 * jp + (mi >> 1)) * 8 + (lane & 7), nc - 1);      // rows beyond nc: a duplicate, never stored
              uint32_t b[4];
              ldmatrix_x4(b, wbase + (size_t)wrow * wp + 2 * (k0 + ks * 16 + b_kofs));
              mma_bf16(acc[2 * jp], al, b[0], b[1]);                  // smallest term first
              if (2 * jp + 1 < ntiles) mma_bf16(acc[2 * jp + 1], al, b[2], b[3]);
              mma_bf16(acc[2 * jp], am, b[0], b[1]);
              if (2 * jp + 1 < ntiles) mma_bf16(acc[2 * jp + 1], am, b[2], b[3]);
              mma_bf16(acc[2 * jp], ah, b[0], b[1]);
              if (2 * jp + 1 < ntiles) mma_bf16(acc[2 * jp + 1], ah, b[2], b[3]);
            }
          }
        }
      }
      __syncthreads();                              // every warp is done reading a_hi: it becomes the partial-sum buffer
      if (tdbg && threadIdx.x == 0 && o == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tdbg[2]));
      {
        const int g = lane >> 2, t = lane & 3;
        float* r0 = red + ((size_t)kq * 32 + mt * 16 + g) * RED_PITCH + 2 * t;
#pragma unroll
        for (int j = 0; j < NT_MAX; ++j) {
          if (j < ntiles) {
            *reinterpret_cast<float2*>(r0 + 8 * j) = make_float2(acc[j][0], acc[j][1]);
            *reinterpret_cast<float2*>(r0 + 8 * RED_PITCH + 8 * j) = make_float2(acc[j][2], acc[j][3]);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < EP; ++e) {
        const int idx = threadIdx.x + e * SM_THREADS;
        const int r = idx / nc, c = idx - r * nc;
        if (r < 32 && (fused || rg + r < B)) {
          const float* p = red + (size_t)r * RED_PITCH + c;
          float v = 0.f;
#pragma unroll
          for (int g = 0; g < KQ; ++g) v += p[g * 32 * RED_PITCH];
          v = apply_act(v + bias_v[e], op.act);
          if (fused) {                             // hidden activation -> three bf16 terms, the A operand of linear2
            if (rg + r >= B) v = 0.f;
            __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sm.a_hi) + FFH_OFFSET);
            const __nv_bfloat16 t0 = __float2bfloat16_rn(v);
            const float r1 = v - __bfloat162float(t0);
            const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
            hb[r * FFH_PITCH + c] = t0; hb[(32 + r) * FFH_PITCH + c] = t1; hb[(64 + r) * FFH_PITCH + c] = t2;
          } else {
            if (op.res != nullptr) v += res_v[e];
            __stcg(op.out + (size_t)(rg + r) * op.ldo + n0 + c, v);
          }
        }
      }
      if (fused) {
        // linear2 on the CTA's 16 hidden units: partial[32, d] = h[32, 16] * W2[:, h0 .. h0 + 15]^T, one k16 step, warp =
        // (row tile, eighth of the output columns); the partial sums go to this CTA's slab, summed by the next phase
        __syncthreads();
        const SmallOp f2 = ph.op[1];
        const unsigned char* w2s = wbase + (size_t)nc * wp + (op.ln_g != nullptr ? (size_t)op.K * 8 : 0);
        const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const unsigned char*>(sm.a_hi) + FFH_OFFSET);
        const int ng = warp >> 1, tiles = f2.N >> 6;            // n-tiles of this warp (d / 64 <= 8)
        uint32_t h_t[3][4];
#pragma unroll
        for (int t3 = 0; t3 < 3; ++t3) ldmatrix_x4(h_t[t3], hb + (t3 * 32 + a_row) * FFH_PITCH + a_kofs);
        float* pslab = f2.out + ((size_t)blockIdx.x * 32) * f2.ldo;
        const int g4 = lane >> 2, t4 = lane & 3;
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          if (2 * jp < tiles) {
            float c2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            const int nrow = (ng * tiles + 2 * jp + (mi >> 1)) * 8 + (lane & 7);
            uint32_t b[4];
            ldmatrix_x4(b, w2s + (size_t)nrow * FF2_PITCH + 2 * b_kofs);
#pragma unroll
            for (int t3 = 2; t3 >= 0; --t3) {                   // smallest term first
              mma_bf16(c2[0], h_t[t3], b[0], b[1]);
              mma_bf16(c2[1], h_t[t3], b[2], b[3]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int col = (ng * tiles + 2 * jp + u) * 8 + 2 * t4;
              const int r_lo = mt * 16 + g4;
              if (rg + r_lo < B) __stcg(reinterpret_cast<float2*>(pslab + (size_t)r_lo * f2.ldo + col), make_float2(c2[u][0], c2[u][1]));
              if (rg + r_lo + 8 < B) __stcg(reinterpret_cast<float2*>(pslab + (size_t)(r_lo + 8) * f2.ldo + col), make_float2(c2[u][2], c2[u][3]));
            }
          }
        }
      }
    }
    if (tdbg && threadIdx.x == 0 && o == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tdbg[3]));
    wbase += (size_t)nc * wp + (op.ln_g != nullptr ? (size_t)op.K * 8 : 0);
  }
}

// Second half of the fused feed-forward block: x[r, n] += b2[n] + sum over the hidden-unit slabs of partial[slab][r, n].
// A CTA owns four output columns; thread = (row, sixteenth of the slabs): 16-byte loads all in flight, a sequential sum
// per thread, then a fixed shuffle tree over the 16 threads of a row (deterministic order).
__device__ void run_reduce_phase(const SmallPhase& ph, int B) {
  const SmallOp op = ph.op[0];                    // in = slabs [K][32][ld_in], K = number of slabs, N = d, res = out = x
  const int c0 = (int)blockIdx.x * 4;
  if (c0 >= op.N) return;
  const int r = threadIdx.x >> 4, q = threadIdx.x & 15;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < B) {
    float4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int p = q + 16 * i;
      v[i] = p < op.K ? __ldcg(reinterpret_cast<const float4*>(op.in + ((size_t)p * 32 + r) * op.ld_in + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (q == 0 && r < B) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias + c0));
    const float4 xv = __ldcg(reinterpret_cast<const float4*>(op.res + (size_t)r * op.ldr + c0));
    __stcg(reinterpret_cast<float4*>(op.out + (size_t)r * op.ldo + c0),
           make_float4(acc.x + bv.x + xv.x, acc.y + bv.y + xv.y, acc.z + bv.z + xv.z, acc.w + bv.w + xv.w));
  }
}

// Attention phase: pair p = (row, head) goes to CTA p % grid, warp p / grid (one pair per warp: B * nhead <= 8 * grid).
// Cached K / V rows do not depend on this step's activations (self-attention: positions < step; cross-attention: the
// projected memory tokens), so the pair's rows are copied into the idle weight buffer BEFORE the barrier that precedes
// the phase, as many as fit; after the barrier only q and the step's new k, v row come through L2 (one round trip),
// and the new row is used from registers.
struct AttnPlan { int ppc, cap; };               // pairs per CTA, prefetchable positions per pair
__device__ __forceinline__ AttnPlan attn_plan(const AttnArgs& a) {
  AttnPlan pl;
  pl.ppc = (a.B * a.nhead + (int)gridDim.x - 1) / (int)gridDim.x;
  pl.cap = min(W_BUF_BYTES / (pl.ppc * 2 * a.hd * (int)sizeof(float)), 4 * kPagePos);   // four page indices in registers
  return pl;
}
__device__ __forceinline__ size_t attn_row_off(const AttnArgs& a, int b, int h, int p) {
  if (a.page_table != nullptr)
    return (size_t)__ldcg(a.page_table + (size_t)b * a.pages_per_seq + (p >> kPageShift)) * a.page_stride +
           (size_t)(p & (kPagePos - 1)) * a.row_stride + h * a.hd;
  return (size_t)b * a.seq_stride + (size_t)p * a.row_stride + h * a.hd;
}

__device__ void prefetch_attention(const AttnArgs& a, Smem& sm, int buf, int step) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AttnPlan pl = attn_plan(a);
  const int pair = (int)blockIdx.x + warp * (int)gridDim.x;
  if (warp < pl.ppc && pair < a.B * a.nhead) {
    const int b = pair / a.nhead, h = pair % a.nhead;
    const int n = a.fixed_len >= 0 ? a.fixed_len : step + 1;
    const int np = min(a.knew != nullptr ? n - 1 : n, pl.cap);          // self: the last row is produced in the phase itself
    float* kv = reinterpret_cast<float*>(sm.w[buf]) + (size_t)warp * 2 * pl.cap * a.hd;
    const int cpr = a.hd >> 2;                                           // 16-byte pieces per row
    // paged cache: the (<= 4, cap <= 64 positions) page indices of the row, all in flight at once, instead of one
    // dependent L2 round trip per copy
    int pg0 = 0, pg1 = 0, pg2 = 0, pg3 = 0;
    if (a.page_table != nullptr) {
      const int* pt = a.page_table + (size_t)b * a.pages_per_seq;
      if (np > 0) pg0 = __ldcg(pt);
      if (np > kPagePos) pg1 = __ldcg(pt + 1);
      if (np > 2 * kPagePos) pg2 = __ldcg(pt + 2);
      if (np > 3 * kPagePos) pg3 = __ldcg(pt + 3);
    }
    const bool pow2 = cpr == 16;
    for (int i = lane; i < np * cpr; i += 32) {
      const int p = pow2 ? i >> 4 : i / cpr, c = pow2 ? i & 15 : i - p * cpr;
      size_t off;
      if (a.page_table != nullptr) {
        const int q = p >> kPageShift;
        const int pg = q == 0 ? pg0 : q == 1 ? pg1 : q == 2 ? pg2 : pg3;
        off = (size_t)pg * a.page_stride + (size_t)(p & (kPagePos - 1)) * a.row_stride + h * a.hd + 4 * c;
      } else {
        off = (size_t)b * a.seq_stride + (size_t)p * a.row_stride + h * a.hd + 4 * c;
      }
      cp_async16(kv + (size_t)p * a.hd + 4 * c, a.kcache + off);
      cp_async16(kv + (size_t)(pl.cap + p) * a.hd + 4 * c, a.vcache + off);
    }
  }
  cp_async_commit();
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

// One warp per (row, head): q.k * scale -> softmax -> sum w v (autoregressive_decoder.py:1270-1296, 1302-1307).  Eight
// lanes share a cached row (16-byte pieces sub, sub + 8, ...), four rows per warp instruction.
__device__ void run_attention_phase(const AttnArgs& a_in, Smem& sm, int buf, int step) {
  const AttnArgs a = a_in;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 7, pg = lane >> 3;
  const int hd = a.hd, cpr = hd >> 2;
  const AttnPlan pl = attn_plan(a);
  const int pair = (int)blockIdx.x + warp * (int)gridDim.x;
  if (warp >= pl.ppc || pair >= a.B * a.nhead) { cp_async_wait_all(); return; }
  const int b = pair / a.nhead, h = pair % a.nhead;
  const bool self = a.knew != nullptr;
  const int n = a.fixed_len >= 0 ? a.fixed_len : step + 1;
  const int ncached = self ? n - 1 : n, np = min(ncached, pl.cap);
  float* sc = reinterpret_cast<float*>(sm.a_hi) + warp * MAX_N_SCORES;
  const float* kv = reinterpret_cast<const float*>(sm.w[buf]) + (size_t)warp * 2 * pl.cap * hd;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 q[4], kn[4], vn[4];
  size_t new_off = 0;
  if (self) new_off = attn_row_off(a, b, h, n - 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = sub + 8 * i;
    q[i] = kn[i] = vn[i] = zero4;
    if (c < cpr) {
      q[i] = __ldcg(reinterpret_cast<const float4*>(a.q + (size_t)b * a.ldq + h * hd) + c);
      if (self) {
        kn[i] = __ldcg(reinterpret_cast<const float4*>(a.knew + (size_t)b * a.ldn + h * hd) + c);
        vn[i] = __ldcg(reinterpret_cast<const float4*>(a.vnew + (size_t)b * a.ldn + h * hd) + c);
      }
    }
  }
  cp_async_wait_all();                             // this warp's staged rows (its own copies) ...
  __syncwarp();                                    // ... of every lane
  if (self && pg == 0) {                           // append this step's key / value (:1266-1267)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = sub + 8 * i;
      if (c < cpr) {
        *(reinterpret_cast<float4*>(a.kcache + new_off) + c) = kn[i];
        *(reinterpret_cast<float4*>(a.vcache + new_off) + c) = vn[i];
      }
    }
  }
#pragma unroll 2
  for (int p0 = 0; p0 < ncached; p0 += 4) {
    const int p = min(p0 + pg, ncached - 1);
    const float4* kp = reinterpret_cast<const float4*>(p < np ? kv + (size_t)p * hd : a.kcache + attn_row_off(a, b, h, p));
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = sub + 8 * i;
      if (c < cpr) d = dot4(q[i], kp[c], d);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1); d += __shfl_xor_sync(0xffffffffu, d, 2); d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (sub == 0 && p0 + pg < ncached) sc[p0 + pg] = d * a.scale;
  }
  if (self) {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) d = dot4(q[i], kn[i], d);
    d += __shfl_xor_sync(0xffffffffu, d, 1); d += __shfl_xor_sync(0xffffffffu, d, 2); d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (lane == 0) sc[n - 1] = d * a.scale;
  }
  __syncwarp();
  float m = -INFINITY;
  for (int p = lane; p < n; p += 32) m = fmaxf(m, sc[p]);
  m = warp_max(m);
  float sum = 0.f;
  for (int p = lane; p < n; p += 32) { const float e = expf(sc[p] - m); sc[p] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  for (int p = lane; p < n; p += 32) sc[p] = sc[p] / sum;
  __syncwarp();
  float4 acc[4] = {zero4, zero4, zero4, zero4};
#pragma unroll 2
  for (int p0 = 0; p0 < ncached; p0 += 4) {
    const int p = min(p0 + pg, ncached - 1);
    const float w = p0 + pg < ncached ? sc[p] : 0.f;
    const float4* vp = reinterpret_cast<const float4*>(p < np ? kv + (size_t)(pl.cap + p) * hd : a.vcache + attn_row_off(a, b, h, p));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = sub + 8 * i;
      if (c < cpr) {
        const float4 v = vp[c];
        acc[i].x = fmaf(w, v.x, acc[i].x); acc[i].y = fmaf(w, v.y, acc[i].y); acc[i].z = fmaf(w, v.z, acc[i].z); acc[i].w = fmaf(w, v.w, acc[i].w);
      }
    }
  }
  if (self && pg == 0) {
    const float w = sc[n - 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i].x = fmaf(w, vn[i].x, acc[i].x); acc[i].y = fmaf(w, vn[i].y, acc[i].y); acc[i].z = fmaf(w, vn[i].z, acc[i].z); acc[i].w = fmaf(w, vn[i].w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {                    // the four row groups meet
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, o); acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, o);
      acc[i].z += __shfl_xor_sync(0xffffffffu, acc[i].z, o); acc[i].w += __shfl_xor_sync(0xffffffffu, acc[i].w, o);
    }
  }
  if (pg == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = sub + 8 * i;
      if (c < cpr) __stcg(reinterpret_cast<float4*>(a.out + (size_t)b * a.ldo + h * hd) + c, acc[i]);
    }
  }
}

__global__ void __launch_bounds__(SM_THREADS, 1)
decode_small_kernel(const SmallPhase* __restrict__ phases, int n_phases, int B, const StepState* st, unsigned* bar,
                    int bar_shards, unsigned long long* dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  if (st->done) return;                           // uniform: written by step_end_kernel of an earlier launch
  const int step = st->step;
  unsigned target = 0;                            // barrier epoch
  // phase descriptors are read through shared memory (ph[p & 1]); the next one travels through a register while the
  // current phase runs (one word per thread), so its L2 latency is never waited for
  constexpr int kDescWords = (int)(sizeof(SmallPhase) / 4);
  static_assert(kDescWords <= SM_THREADS, "one descriptor word per thread");
  auto desc_load = [&](int p) -> uint32_t {
    return (int)threadIdx.x < kDescWords ? __ldg(reinterpret_cast<const uint32_t*>(phases + p) + threadIdx.x) : 0u;
  };
  auto desc_store = [&](int p, uint32_t v) {
    if ((int)threadIdx.x < kDescWords) reinterpret_cast<uint32_t*>(&sm.ph[p & 1])[threadIdx.x] = v;
  };
  desc_store(0, desc_load(0));
  __syncthreads();
  prefetch_weights(sm.ph[0], sm, 0, step);
  for (int p = 0; p < n_phases; ++p) {
    const SmallPhase& ph = sm.ph[p & 1];
    uint32_t next_word = 0;
    if (p + 1 < n_phases) next_word = desc_load(p + 1);
    unsigned long long t0 = 0, t1 = 0, t2 = 0;
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));   // SCV_SMALL_DEBUG
    if (ph.kind == 0 || ph.kind == 2) run_gemv_phase(ph, sm, p & 1, B, (dbg && blockIdx.x == 0) ? dbg + 2048 + 4 * p : nullptr);
    else if (ph.kind == 3) run_reduce_phase(ph, B);
    else run_attention_phase(ph.attn, sm, p & 1, step);
    if (dbg && blockIdx.x == 0) { __syncthreads(); if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); }
    if (p + 1 < n_phases) {
      desc_store(p + 1, next_word);                        // slot (p + 1) & 1 was last read in phase p - 1
      __syncthreads();                                     // the staged descriptor of phase p + 1 is complete
      prefetch_weights(sm.ph[(p + 1) & 1], sm, (p + 1) & 1, step);   // weights / cached rows do not depend on this phase's results
      grid_barrier(bar, target, bar_shards);
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
      dbg[2 * p] = t1 - t0; dbg[2 * p + 1] = t2 - t1;
    }
  }
  cp_async_wait_all();
}

// ---- the whole decode in ONE launch (plain greedy calls): the step loop, the sampling epilogue, the embedding of the
// chosen token and the "every row has emitted END" exit run inside the kernel, so a decode costs two launches (the
// embedding of the start token + this kernel) instead of five per step, and the host never polls.
// Everything one step writes and a later step reads is read through L2 (ld.cg / cp.async.cg): L1 is not coherent
// across SMs and is no longer flushed by kernel boundaries.
__device__ __forceinline__ void embed_row(const SmallTail& t, int b, int tok, int pos, int lane) {
  if (lane == 0 && (pos & (kPagePos - 1)) == 0)            // KV page of positions pos .. pos + 15 (embed_kernel)
    t.page_table[b * t.pages_per_seq + (pos >> kPageShift)] = atomicAdd(&t.sp.st->next_free_page, 1);
  const __nv_bfloat16* row = t.emb + (size_t)tok * t.ld_emb;
  const float* pe = t.pe + (size_t)pos * t.d;
  float* x = t.x + (size_t)b * t.d;
  for (int i = lane; i < t.d; i += 32) __stcg(x + i, __bfloat162float(row[i]) + __ldg(pe + i));
}

__global__ void __launch_bounds__(SM_THREADS, 1)
decode_small_persist_kernel(const SmallPhase* __restrict__ phases, int n_phases, int B, unsigned* bar, int bar_shards, SmallTail tail) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  StepState* st = tail.sp.st;
  if (st->done) return;
  int step = st->step;
  unsigned target = 0;                            // barrier epoch
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kDescWords = (int)(sizeof(SmallPhase) / 4);
  auto desc_load = [&](int p) -> uint32_t {
    return (int)threadIdx.x < kDescWords ? __ldg(reinterpret_cast<const uint32_t*>(phases + p) + threadIdx.x) : 0u;
  };
  auto desc_store = [&](int slot, uint32_t v) {
    if ((int)threadIdx.x < kDescWords) reinterpret_cast<uint32_t*>(&sm.ph[slot])[threadIdx.x] = v;
  };
  desc_store(0, desc_load(0));
  __syncthreads();
  prefetch_weights(sm.ph[0], sm, 0, step);
  int it = 0;                                      // phases executed so far: descriptor slot / weight buffer = it & 1
  for (;;) {
    for (int p = 0; p < n_phases; ++p, ++it) {
      const int cur = it & 1, nxt = cur ^ 1;
      const uint32_t next_word = desc_load(p + 1 < n_phases ? p + 1 : 0);
      const SmallPhase& ph = sm.ph[cur];
      if (ph.kind == 0 || ph.kind == 2) run_gemv_phase(ph, sm, cur, B, nullptr);
      else if (ph.kind == 3) run_reduce_phase(ph, B);
      else run_attention_phase(ph.attn, sm, cur, step);
      desc_store(nxt, next_word);                  // slot nxt was last read in the previous phase
      __syncthreads();
      // the next phase's weights / cached rows do not depend on this phase's results (after the last phase: the first
      // projection of the next step, harmless if the decode ends here)
      prefetch_weights(sm.ph[nxt], sm, nxt, step);
      grid_barrier(bar, target, bar_shards);
    }
    // sampling epilogue (:1415-1548, plain greedy) + embedding of the chosen token at the next position: warp per row
    for (int b = (int)blockIdx.x + warp * (int)gridDim.x; b < B; b += (int)gridDim.x * SM_WARPS) {
      int tok = greedy_row_token(tail.sp, b, step, lane);
      if (lane == 0) tok = commit_token(tail.sp, b, step, tok, 0.f);
      tok = __shfl_sync(0xffffffffu, tok, 0);
      if (step + 1 < tail.max_steps) embed_row(tail, b, tok, step + 1, lane);
    }
    grid_barrier(bar, target, bar_shards);
    // step_end_kernel: every CTA takes the same decision from the same counter (nobody changes it before the next epilogue)
    const int unfinished = (int)ld_acquire_u32(reinterpret_cast<const unsigned*>(&st->n_unfinished));
    const int s1 = step + 1;
    const bool done = unfinished <= 0 || s1 >= tail.max_steps;      // finished.all() -> break (:1547-1548)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->step = s1;
      st->degenerate = 0;
      if (done) { st->done = 1; st->out_len = s1; }
    }
    if (done) break;
    step = s1;
  }
  cp_async_wait_all();
}

}  // namespace

size_t small_step_smem_bytes() { return sizeof(Smem); }

static int bar_shards_env() {
  static const int v = [] { const char* e = getenv("SCV_SMALL_BAR_SHARDS"); return std::max(1, std::min(e ? atoi(e) : 1, kBarWords / 32)); }();
  return v;
}

int small_cols_per_cta(int N, int grid) {
  // whole 8-column MMA tiles: rounding up never adds a tile to the busiest CTA, and fewer CTAs stage the input rows
  // (every active CTA reads all of them through L2, which is what bounds a phase)
  static const int min_cpc = [] { const char* e = getenv("SCV_SMALL_CPC"); return e ? atoi(e) : 8; }();
  return std::max(round_up(ceil_div(N, grid), 8), std::min(round_up(min_cpc, 8), NT_MAX * 8));
}

bool small_phase_fits(const SmallPhase& ph, int grid) {
  if (ph.kind == 3) {                                           // partial-sum reduction of the fused feed-forward block
    const SmallOp& op = ph.op[0];
    return op.N % 4 == 0 && op.N <= 4 * grid && op.K <= 256 && op.ld_in % 4 == 0 && op.ldr % 4 == 0 && op.ldo % 4 == 0 &&
           op.bias != nullptr && op.res != nullptr;
  }
  if (ph.kind == 2) {                                           // linear1 (+ LayerNorm, GELU) and linear2 partial sums
    const SmallOp& f1 = ph.op[0];
    const SmallOp& f2 = ph.op[1];
    if (f1.cpc != 16 || f1.N % 16 != 0 || f1.N > 16 * grid || f1.K > KC || f1.K % 16 != 0 || f1.ldw != f1.K) return false;
    if (f1.ld_in % 4 != 0 || (reinterpret_cast<uintptr_t>(f1.in) & 15u) != 0) return false;
    if (f2.N % 128 != 0 || f2.N > 512 || f2.K != f1.N || f2.ldw % 8 != 0 || f2.ldo % 2 != 0) return false;   // whole tile pairs per warp
    const size_t bytes = (size_t)16 * (f1.ldw * 2 + W_ROW_PAD) + (f1.ln_g != nullptr ? (size_t)f1.K * 8 : 0) + (size_t)f2.N * FF2_PITCH;
    return bytes <= (size_t)W_BUF_BYTES;
  }
  if (ph.kind != 0)
    return ph.attn.hd <= 128 && ph.attn.hd % 4 == 0 && ph.attn.max_n <= MAX_N_SCORES && ph.attn.B * ph.attn.nhead <= SM_WARPS * grid &&
           ph.attn.ldq % 4 == 0 && ph.attn.ldn % 4 == 0 && ph.attn.ldo % 4 == 0 && ph.attn.row_stride % 4 == 0;
  size_t bytes = 0;
  for (int o = 0; o < ph.nops; ++o) {
    const SmallOp& op = ph.op[o];
    if (op.K % 16 != 0 || op.ldw != op.K) return false;      // whole k16 MMA steps, rows copied in 16-byte pieces
    if (op.ld_in % 4 != 0 || (reinterpret_cast<uintptr_t>(op.in) & 15u) != 0) return false;   // float4 staging
    if (op.cpc % 8 != 0 || op.cpc > NT_MAX * 8 || (long long)op.cpc * grid < op.N) return false;   // accumulators: 6 tiles of 8 columns
    if (op.ln_g != nullptr && (op.K > KC || ((reinterpret_cast<uintptr_t>(op.ln_g) | reinterpret_cast<uintptr_t>(op.ln_b)) & 15u) != 0))
      return false;                                           // the LayerNorm needs the whole row in one chunk
    bytes += (size_t)op.cpc * (op.ldw * 2 + W_ROW_PAD) + (op.ln_g != nullptr ? (size_t)op.K * 8 : 0);
  }
  return bytes <= (size_t)W_BUF_BYTES;
}

int launch_decode_small(const SmallPhase* phases_dev, int n_phases, int B, const StepState* st, unsigned* bar, int grid,
                        cudaStream_t s) {
  SCV_REQUIRE(B >= 1 && B <= kSmallMaxRows, "small-batch step: %d rows (1..%d supported)", B, kSmallMaxRows);
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
    SCV_CUDA(cudaFuncSetAttribute(decode_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(SM_THREADS); cfg.dynamicSmemBytes = sizeof(Smem); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident, or the launch fails: the barrier cannot hang
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static unsigned long long* dbg = nullptr;
  static int dbg_calls = 0;
  static const int dbg_env = [] { const char* e = getenv("SCV_SMALL_DEBUG"); return e ? atoi(e) : 0; }();
  if (dbg_env && dbg == nullptr) { SCV_CUDA(cudaMalloc(&dbg, 4096 * 8)); SCV_CUDA(cudaMemset(dbg, 0, 4096 * 8)); }
  SCV_CUDA(cudaLaunchKernelEx(&cfg, decode_small_kernel, phases_dev, n_phases, B, st, bar, bar_shards_env(), dbg));
  if (dbg_env && ++dbg_calls == dbg_env) {
    SCV_CUDA(cudaStreamSynchronize(s));
    std::vector<unsigned long long> h(4096);
    SCV_CUDA(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int p = 0; p < n_phases; ++p) {
      const unsigned long long* t = h.data() + 2048 + 4 * p;
      fprintf(stderr, "phase %3d: run %6llu ns (stage %6llu, mma %6llu, epilogue %6llu), prefetch+barrier %6llu ns\n", p, h[2 * p],
              t[1] - t[0], t[2] - t[1], t[3] - t[2], h[2 * p + 1]);
    }
  }
  SCV_LAUNCH_CHECK();
  return 0;
}

int launch_decode_small_persist(const SmallPhase* phases_dev, int n_phases, int B, unsigned* bar, int grid, const SmallTail& tail,
                                cudaStream_t s) {
  SCV_REQUIRE(B >= 1 && B <= kSmallMaxRows, "small-batch decode: %d rows (1..%d supported)", B, kSmallMaxRows);
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
    SCV_CUDA(cudaFuncSetAttribute(decode_small_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(SM_THREADS); cfg.dynamicSmemBytes = sizeof(Smem); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident, or the launch fails: the barrier cannot hang
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SCV_CUDA(cudaLaunchKernelEx(&cfg, decode_small_persist_kernel, phases_dev, n_phases, B, bar, bar_shards_env(), tail));
  count_launch();
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
