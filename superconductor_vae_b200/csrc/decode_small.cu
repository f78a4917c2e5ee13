// Small-batch decode step (B <= 32 rows): ONE persistent kernel per step instead of ~140 launches.
//
// At a few dozen rows a decode step is bound by streaming the 107 MB of bf16 weights once (16 us of
// HBM time) but the per-projection kernels spend ~7 us each on launch, prologue and drain: 1.05 ms per step.  Here
// one CTA per SM stays resident for the whole step and walks a list of phases separated by grid-wide barriers:
//   GEMV phase  : up to four projections of the same dependency level (y = act(LN?(x) W^T + b) (+ residual)); the
//                 output columns are dealt round-robin to the CTAs, the input rows ([B, K] fp32, a few KB) are staged
//                 in shared memory with the LayerNorm applied on the fly, lane r of every warp owns batch row r, so a
//                 warp produces one output column for all rows with no cross-lane reduction.  The CTA's weight rows
//                 (bf16, <= 40 KB) are fetched with cp.async BEFORE the barrier that precedes the phase: they do not
//                 depend on activations, so the weight stream overlaps the barrier and the previous phase's tail.
//   attention   : one warp per (row, head), same arithmetic as attention_decode_kernel (decode_kernels.cu).
// Arithmetic: bf16 weights, fp32 activations / accumulate, like linear_simt.cu (every product exact in the FMA).
//
// STATUS: opt-in (SCV_SMALL=1).  Measured on B200 at 32 rows: 1.33 ms per step against 1.05 ms for the per-projection
// path (CUDA graph + PDL): with ~100 dependent phases per step, each phase pays a grid barrier (1.5-2.5 us), an L2
// round trip to stage its input (2.4 us) and a latency-bound compute / epilogue tail (3-5 us; two warps per scheduler
// cannot hide the shared-memory and L2 latencies).  Per-phase times are printed with SCV_SMALL_DEBUG=<launch index>.
// What it needs to win (next round): fewer, fatter phases (q projection fused into the out-projection phase, both
// attention phases overlapped with the next weight stream), mma.sync on the hi/lo split instead of scalar FMAs.
// Reference call sites: models/autoregressive_decoder.py:1244-1313 (layer), :1413-1441 (heads).
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "decode_kernels.cuh"

namespace scv {

namespace {

constexpr int SM_THREADS = 256, SM_WARPS = 8;
constexpr int KC = 1024;                      // columns of the input staged per chunk (LayerNorm inputs fit in one)
constexpr int XS_PITCH = KC + 4;              // floats; row r starts 16 B further in the banks than row r-1
constexpr int W_BUF_BYTES = 40 * 1024;        // weight rows of one phase for one CTA
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

// Grid-wide barrier over co-resident CTAs (the launch is cooperative: one CTA per SM).  `bar` counts arrivals and is
// reset to zero by step_end_kernel after the kernel has finished.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_u32(bar) < target) { __nanosleep(32); }
    __threadfence();
  }
  __syncthreads();
}

struct Smem {
  SmallPhase ph[2];                               // descriptors of the running and of the next phase (copied from global)
  float xs[32 * XS_PITCH];
  __align__(16) unsigned char w[2][W_BUF_BYTES];
  float sc[SM_WARPS * MAX_N_SCORES];
  float gb[2 * KC];                               // LayerNorm weight and bias of the op being staged
};

__device__ __forceinline__ int cols_of_cta(int N) { return (N - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x; }

// Issue the cp.async copies of this CTA's weight rows of every op of `ph` into buffer `buf`.
__device__ void prefetch_weights(const SmallPhase& ph, Smem& sm, int buf) {
  if (ph.kind != 0) return;
  unsigned char* dst = sm.w[buf];
  for (int o = 0; o < ph.nops; ++o) {
    const SmallOp& op = ph.op[o];
    const int nc = cols_of_cta(op.N), row_bytes = op.ldw * 2, chunks = row_bytes / 16;
    for (int i = threadIdx.x; i < nc * chunks; i += SM_THREADS) {
      const int j = i / chunks, c = i - j * chunks;
      const int n = (int)blockIdx.x + j * (int)gridDim.x;
      cp_async16(dst + (size_t)j * row_bytes + 16 * c, reinterpret_cast<const unsigned char*>(op.w + (size_t)n * op.ldw) + 16 * c);
    }
    dst += (size_t)nc * row_bytes;
  }
  cp_async_commit();
}

__device__ void run_gemv_phase(const SmallPhase& ph, Smem& sm, int buf, int B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned char* wbase = sm.w[buf];
  for (int o = 0; o < ph.nops; ++o) {
    const SmallOp op = ph.op[o];                  // by value: the descriptor lives in shared memory, the hot fields in registers
    const int K = op.K, nc = cols_of_cta(op.N);
    const int row_bytes = op.ldw * 2;
    const bool multi_chunk = K > KC;              // then nc <= 8 (small_phase_fits)
    float acc_mc = 0.f;
    for (int k0 = 0; k0 < K; k0 += KC) {
      const int kc = min(KC, K - k0), kc8 = (kc + 7) & ~7;
      __syncthreads();                            // the previous chunk / op / phase is done with xs
      if ((kc & 3) == 0 && (op.ld_in & 3) == 0 && (reinterpret_cast<uintptr_t>(op.in) & 15u) == 0) {
        // asynchronous 16-byte copies L2 -> shared memory (cp.async.cg reads through L2, where the other CTAs'
        // results of the previous phase are): every copy of the chunk is in flight at once
        const int q4 = kc >> 2, total = B * q4;
        for (int i = threadIdx.x; i < total; i += SM_THREADS) {
          const int r = i / q4, c = i - r * q4;
          cp_async16(sm.xs + r * XS_PITCH + 4 * c, op.in + (size_t)r * op.ld_in + k0 + 4 * c);
        }
        if (op.ln_g != nullptr) {                // K <= KC and K % 4 == 0 here
          if (((reinterpret_cast<uintptr_t>(op.ln_g) | reinterpret_cast<uintptr_t>(op.ln_b)) & 15u) == 0) {
            for (int i = threadIdx.x; i < 2 * q4; i += SM_THREADS)
              cp_async16(sm.gb + (i < q4 ? 0 : KC) + 4 * (i < q4 ? i : i - q4), (i < q4 ? op.ln_g : op.ln_b) + 4 * (i < q4 ? i : i - q4));
          } else {
            for (int k = threadIdx.x; k < K; k += SM_THREADS) { sm.gb[k] = __ldg(op.ln_g + k); sm.gb[KC + k] = __ldg(op.ln_b + k); }
          }
        }
        cp_async_commit();
        if (kc8 != kc)
          for (int r = threadIdx.x; r < B; r += SM_THREADS)
            for (int k = kc; k < kc8; ++k) sm.xs[r * XS_PITCH + k] = 0.f;
      } else {
        for (int i = threadIdx.x; i < B * kc8; i += SM_THREADS) {
          const int r = i / kc8, k = i - r * kc8;
          sm.xs[r * XS_PITCH + k] = k < kc ? __ldcg(op.in + (size_t)r * op.ld_in + k0 + k) : 0.f;
        }
        if (op.ln_g != nullptr)
          for (int k = threadIdx.x; k < K; k += SM_THREADS) { sm.gb[k] = __ldg(op.ln_g + k); sm.gb[KC + k] = __ldg(op.ln_b + k); }
      }
      cp_async_wait_all();                        // the chunk and this phase's weight rows have landed (own copies) ...
      __syncthreads();                            // ... and everybody else's, and xs is complete
      if (op.ln_g != nullptr) {
        // LayerNorm in place (the whole row is in this chunk): 8 threads per row, two passes, eps 1e-5
        // (rows >= B hold stale shared memory: normalised like the rest so the warp shuffles stay convergent, never used)
        const int r = threadIdx.x >> 3, sub = threadIdx.x & 7;
        {
          float* xr = sm.xs + r * XS_PITCH;
          float s = 0.f;
          for (int k = sub; k < K; k += 8) s += xr[k];
          s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
          const float mean = s / (float)K;
          float q = 0.f;
          for (int k = sub; k < K; k += 8) { const float d = xr[k] - mean; q = fmaf(d, d, q); }
          q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
          const float rstd = 1.0f / sqrtf(q / (float)K + 1e-5f);
          for (int k = sub; k < K; k += 64) {       // eight elements per batch: loads first, then the in-place stores
            float xv[8], gv[8], bv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int kk = k + 8 * u;
              if (kk < K) { xv[u] = xr[kk]; gv[u] = sm.gb[kk]; bv[u] = sm.gb[KC + kk]; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int kk = k + 8 * u;
              if (kk < K) xr[kk] = (xv[u] - mean) * rstd * gv[u] + bv[u];
            }
          }
        }
        __syncthreads();
      }
      // lane = batch row, warp = K slice: every warp works on every column of the CTA (eight at a time, sharing each
      // x load), the eight partial sums per (row, column) meet in shared memory.  (One warp per column left most
      // warps idle and ran a dependent load -> FMA chain 64-256 steps long: 13-50 us per phase.)
      const int kslice = ((kc8 / 8 + SM_WARPS - 1) / SM_WARPS) * 8;      // multiple of 8 columns per warp
      const int kbeg = min(warp * kslice, kc8), kend = min(kbeg + kslice, kc8);
      const float* xr = sm.xs + lane * XS_PITCH;
      for (int cg = 0; cg < nc; cg += 8) {
        const unsigned char* wr = wbase + (size_t)cg * row_bytes + 2 * k0;
        float a0[8], a1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
        // warp c will finish column cg + c: fetch its bias and residual now, their L2 latency hides behind the FMAs
        const bool fin = warp < 8 && cg + warp < nc && lane < B && k0 + KC >= K;
        const int n_fin = (int)blockIdx.x + (cg + warp) * (int)gridDim.x;
        float bias_v = 0.f, res_v = 0.f;
        if (fin) {
          if (op.bias != nullptr) bias_v = __ldg(op.bias + n_fin);
          if (op.res != nullptr) res_v = __ldcg(op.res + (size_t)lane * op.ldr + n_fin);
        }
#pragma unroll 2
        for (int k = kbeg; k < kend; k += 8) {
          const float4 xa = *reinterpret_cast<const float4*>(xr + k), xb = *reinterpret_cast<const float4*>(xr + k + 4);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (cg + c < nc) {
              const uint4 wv = *reinterpret_cast<const uint4*>(wr + (size_t)c * row_bytes + 2 * k);
              a0[c] = fmaf(xa.x, bf16_bits_to_float(wv.x & 0xffffu), a0[c]); a1[c] = fmaf(xa.y, __uint_as_float(wv.x & 0xffff0000u), a1[c]);
              a0[c] = fmaf(xa.z, bf16_bits_to_float(wv.y & 0xffffu), a0[c]); a1[c] = fmaf(xa.w, __uint_as_float(wv.y & 0xffff0000u), a1[c]);
              a0[c] = fmaf(xb.x, bf16_bits_to_float(wv.z & 0xffffu), a0[c]); a1[c] = fmaf(xb.y, __uint_as_float(wv.z & 0xffff0000u), a1[c]);
              a0[c] = fmaf(xb.z, bf16_bits_to_float(wv.w & 0xffffu), a0[c]); a1[c] = fmaf(xb.w, __uint_as_float(wv.w & 0xffff0000u), a1[c]);
            }
          }
        }
        if (cg > 0 || k0 > 0) __syncthreads();     // the previous group's partial sums have been consumed
#pragma unroll
        for (int c = 0; c < 8; ++c) sm.sc[(warp * 8 + c) * 32 + lane] = a0[c] + a1[c];
        __syncthreads();
        if (warp < 8 && cg + warp < nc) {          // warp c finishes column cg + c: lane = row
          float t = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < SM_WARPS; ++w2) t += sm.sc[(w2 * 8 + warp) * 32 + lane];
          if (multi_chunk) {
            acc_mc += t;                           // K spans several chunks (nc <= 8): one column per warp, kept in a register
          } else if (lane < B) {
            float v = apply_act(t + bias_v, op.act);
            if (op.res != nullptr) v += res_v;
            __stcg(op.out + (size_t)lane * op.ldo + n_fin, v);
          }
        }
      }
    }
    if (multi_chunk && warp < nc && lane < B) {
      const int n = (int)blockIdx.x + warp * (int)gridDim.x;
      float v = acc_mc;
      if (op.bias != nullptr) v += __ldg(op.bias + n);
      v = apply_act(v, op.act);
      if (op.res != nullptr) v += __ldcg(op.res + (size_t)lane * op.ldr + n);
      __stcg(op.out + (size_t)lane * op.ldo + n, v);
    }
    wbase += (size_t)nc * row_bytes;
  }
}

// One warp per (row, head); same operation order as attention_decode_kernel (q.k * scale -> softmax -> sum w v).
__device__ void run_attention_phase(const AttnArgs& a_in, Smem& sm) {
  const AttnArgs a = a_in;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = a.hd;
  const int n = a.fixed_len >= 0 ? a.fixed_len : a.st->step + 1;
  float* sc = sm.sc + warp * MAX_N_SCORES;
  const int pairs = a.B * a.nhead;
  for (int gw = (int)blockIdx.x * SM_WARPS + warp; gw < pairs; gw += (int)gridDim.x * SM_WARPS) {
    const int b = gw / a.nhead, h = gw % a.nhead;
    const bool paged = a.page_table != nullptr;
    const int* pt = paged ? a.page_table + (size_t)b * a.pages_per_seq : nullptr;
    auto row_off = [&](int p) -> size_t {
      if (paged) return (size_t)pt[p >> kPageShift] * a.page_stride + (size_t)(p & (kPagePos - 1)) * a.row_stride + h * hd;
      return (size_t)b * a.seq_stride + (size_t)p * a.row_stride + h * hd;
    };
    float qv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = lane + 32 * j;
      qv[j] = e < hd ? __ldcg(a.q + (size_t)b * a.ldq + h * hd + e) : 0.f;
    }
    if (a.knew != nullptr) {                       // append this step's key / value (:1266-1267)
      const size_t off = row_off(n - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = lane + 32 * j;
        if (e < hd) {
          a.kcache[off + e] = __ldcg(a.knew + (size_t)b * a.ldn + h * hd + e);
          a.vcache[off + e] = __ldcg(a.vnew + (size_t)b * a.ldn + h * hd + e);
        }
      }
      __syncwarp();
    }
    for (int p0 = 0; p0 < n; p0 += 8) {           // eight cached rows in flight per warp
      float kk[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* kp = a.kcache + row_off(min(p0 + u, n - 1));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = lane + 32 * j;
          kk[u][j] = e < hd ? kp[e] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) d = fmaf(qv[j], kk[u][j], d);
        d = warp_sum(d);
        if (lane == 0 && p0 + u < n) sc[p0 + u] = d * a.scale;
      }
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
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = 0; p0 < n; p0 += 8) {
      float vv[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* vp = a.vcache + row_off(min(p0 + u, n - 1));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = lane + 32 * j;
          vv[u][j] = e < hd ? vp[e] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (p0 + u < n) {
          const float w = sc[p0 + u];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(w, vv[u][j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = lane + 32 * j;
      if (e < hd) __stcg(a.out + (size_t)b * a.ldo + h * hd + e, acc[j]);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(SM_THREADS, 1)
decode_small_kernel(const SmallPhase* __restrict__ phases, int n_phases, int B, const StepState* st, unsigned* bar,
                    unsigned long long* dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  if (st->done) return;                           // uniform: written by step_end_kernel of an earlier launch
  unsigned target = 0;
  // phase descriptors are read through shared memory (ph[p & 1]); the next one is copied while the current one runs
  auto stage_desc = [&](int p) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(phases + p);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.ph[p & 1]);
    for (int i = threadIdx.x; i < (int)(sizeof(SmallPhase) / 4); i += SM_THREADS) dst[i] = __ldg(src + i);
  };
  stage_desc(0);
  __syncthreads();
  prefetch_weights(sm.ph[0], sm, 0);
  for (int p = 0; p < n_phases; ++p) {
    const SmallPhase& ph = sm.ph[p & 1];
    if (p + 1 < n_phases) stage_desc(p + 1);      // visible after the barriers inside / after this phase
    unsigned long long t0 = 0, t1 = 0, t2 = 0;
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));   // SCV_SMALL_DEBUG
    if (ph.kind == 0) run_gemv_phase(ph, sm, p & 1, B);
    else run_attention_phase(ph.attn, sm);
    if (dbg && blockIdx.x == 0) { __syncthreads(); if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); }
    if (p + 1 < n_phases) {
      __syncthreads();                                     // the staged descriptor of phase p + 1 is complete
      prefetch_weights(sm.ph[(p + 1) & 1], sm, (p + 1) & 1);   // weights do not depend on this phase's results
      grid_barrier(bar, target);
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
      dbg[2 * p] = t1 - t0; dbg[2 * p + 1] = t2 - t1;
    }
  }
  cp_async_wait_all();
}

}  // namespace

size_t small_step_smem_bytes() { return sizeof(Smem); }

bool small_phase_fits(const SmallPhase& ph, int grid) {
  if (ph.kind != 0) return ph.attn.hd <= 128 && ph.attn.max_n <= MAX_N_SCORES;
  size_t bytes = 0;
  for (int o = 0; o < ph.nops; ++o) {
    const SmallOp& op = ph.op[o];
    const int nc = ceil_div(op.N, grid);
    if (op.ldw % 8 != 0 || op.ldw < op.K) return false;
    if (op.K > KC && nc > 8) return false;                    // several chunks: one register accumulator per warp
    if (op.ln_g != nullptr && op.K > KC) return false;       // the LayerNorm needs the whole row in one chunk
    bytes += (size_t)nc * op.ldw * 2;
  }
  return bytes <= (size_t)W_BUF_BYTES;
}

int launch_decode_small(const SmallPhase* phases_dev, int n_phases, int B, const StepState* st, unsigned* bar, int grid,
                        cudaStream_t s) {
  SCV_REQUIRE(B >= 1 && B <= 32, "small-batch step: %d rows (1..32 supported)", B);
  static bool attr_set = false;
  if (!attr_set) {
    SCV_CUDA(cudaFuncSetAttribute(decode_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    attr_set = true;
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
  SCV_CUDA(cudaLaunchKernelEx(&cfg, decode_small_kernel, phases_dev, n_phases, B, st, bar, dbg));
  if (dbg_env && ++dbg_calls == dbg_env) {
    SCV_CUDA(cudaStreamSynchronize(s));
    std::vector<unsigned long long> h(2 * n_phases);
    SCV_CUDA(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int p = 0; p < n_phases; ++p) fprintf(stderr, "phase %3d: run %6llu ns, prefetch+barrier %6llu ns\n", p, h[2 * p], h[2 * p + 1]);
  }
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
