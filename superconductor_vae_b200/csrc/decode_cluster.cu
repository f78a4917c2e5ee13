// Cluster-parallel small-batch decode (<= 16 clusters x 4 rows): the north-star's persistent decode kernel with
// TMA-staged (cp.async.bulk) bf16 weight streaming, WITHOUT grid-wide barriers.
//
// The grid-barrier kernel in decode_small.cu partitions every projection's output columns over all 148 SMs, so a step is
// ~99 dependent phases separated by grid barriers and runs at ~0.8 ms per step where the HBM time is 16 us.  Here the
// ROWS are dealt to thread-block clusters of 8 CTAs (one cluster = a small tensor-parallel group for its 1-4 rows):
//   * CTA `rank` of a cluster owns 1/8 of every projection's output columns (its attention head(s), 64 of the 512
//     residual columns, 256 of the 2048 hidden units, 594 of the 4752 logits) and streams exactly those weight rows:
//     a producer warp walks the step's static weight list and copies 32 KB chunks into a 3-stage shared-memory ring with
//     cp.async.bulk + mbarrier transaction counts.  The stream does not depend on activations, so it keeps running
//     through attention phases and barriers; eight compute warps consume the ring (weights as bf16 from shared memory,
//     the rows' activations held in registers, fp32 FMA - every bf16 x fp32 product enters an fp32 accumulator, the
//     arithmetic of the fp32 CUDA-core path).
//   * activations travel through DISTRIBUTED SHARED MEMORY: a CTA writes its slice of a projection's output straight into
//     the buffers of all 8 CTAs (st.shared::cluster), phases are separated by a cluster barrier built from mbarriers
//     (remote mbarrier.arrive.release.cluster + local try_wait.acquire.cluster): 6 per layer, ~0.5 us each, against
//     8 grid barriers of ~2 us plus the L2 round trips of the staged inputs.
//   * every cluster reads all weights, so the chip streams n_clusters x 107 MB per step out of L2 (the matrices are read by
//     all clusters at about the same time, HBM sees them once); that trade - L2 bandwidth for barriers - is what wins.
//   * self-attention / cross-attention of head h run on the CTA that owns head h, on its own q / k / v (never exchanged).
// The whole plain-greedy decode is one launch (step loop, greedy epilogue, END bookkeeping and next-token embedding
// inside); sampling calls run one launch per step followed by the sampler kernels.  Clusters never wait for each other:
// "every row of the batch has emitted END" (reference :1547-1548) is taken from two global counters (rows not finished,
// last finishing step), and a cluster that runs a step too far only writes positions the caller never sees.
// Reference call sites: models/autoregressive_decoder.py:1244-1313 (layer), :1413-1441 (heads), :1505-1548 (sampling).
#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "decode_kernels.cuh"
#include "sampler_device.cuh"

namespace scv {

namespace {

constexpr int CS = kClSize;
constexpr int NW = 8;                          // compute warps
constexpr int CL_THREADS = (NW + 1) * 32;      // + the producer warp
constexpr int STAGE_BYTES = 32 * 1024;
constexpr int NSTAGE = 3;
constexpr int KSEG_MAX = 768;                  // k-values one warp covers per output column: 3 chunks of 8 per lane
constexpr int CPL_MAX = KSEG_MAX / 256;        // 16-byte weight chunks per lane and segment (template parameter CPL: 2 when every segment is <= 512 wide)
constexpr int PART_ROWS = 32;                  // weight rows per stage of a split-K projection (K > 768: at most 10)
constexpr int MAX_SC = 256;                    // attention positions per (row, head)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Shared-memory control block (after the ring and the activation buffers)
struct Ctl {
  unsigned long long full[NSTAGE], empty[NSTAGE], sync[2];
  int go;                 // producer gate: step the consumers have committed to (or -1: stop)
  int cont;               // decision of the cluster's rank 0 after a step: 1 continue, 0 stop
  int tok[kClMaxRows];    // tokens chosen for the cluster's rows (written by the CTA that sampled the row)
  alignas(16) float part[2][PART_ROWS * 4 * kClMaxRows];     // split-K partial sums of one stage (double buffered)
  alignas(16) float sc[NW][MAX_SC];   // attention scores of the CTA's (row, head) items
  alignas(16) float pacc[NW][128];    // partial P*V sums, one row per compute warp
  int pg[NW][16];         // KV page ids of a warp's row
};

constexpr int MAX_INSTR = 320;                // program length the shared-memory copy can hold (19 per layer + heads)

struct Ctx {
  unsigned char* ring; float* bufs; Ctl* ctl;
  const ClInstr* prog;          // the step's program, copied into shared memory once per launch
  uint32_t peer_bufs[CS];       // shared::cluster address of every CTA's activation-buffer base
  uint32_t peer_ctl[CS];        // ... of every CTA's control block
  int rank, row_base, rows_valid;
  uint32_t cons_it;             // ring chunks consumed so far
  uint32_t sync_it;             // cluster barriers passed so far
};

// which weight rows CTA `rank` streams / computes for a GEMV instruction
__device__ __forceinline__ void cols_of_rank(const ClInstr& I, int rank, int& n0, int& nc) {
  if (I.g.split) { nc = I.g.N / CS; n0 = rank * nc; }
  else { nc = rank == 0 ? I.g.N : 0; n0 = 0; }
}
__device__ __forceinline__ int rows_per_stage(const ClInstr& I) { return STAGE_BYTES / (I.g.ldw * 2); }

// ------------------------------------------------------------------------------------------------ producer
// One lane streams the weight rows of every GEMV of the program, in program order, chunk by chunk.
__device__ void producer_step(const ClProgram& P, const ClInstr* prog, Ctl* ctl, unsigned char* ring, int rank, uint32_t& it) {
  for (int i = 0; i < P.n_instr; ++i) {
    const ClInstr& I = prog[i];
    if (I.kind != CL_GEMV) continue;
    int n0, nc;
    cols_of_rank(I, rank, n0, nc);
    const int rps = rows_per_stage(I), row_bytes = I.g.ldw * 2;
    for (int r0 = 0; r0 < nc; r0 += rps, ++it) {
      const int rows = min(rps, nc - r0);
      const uint32_t st = it % NSTAGE, use = it / NSTAGE;
      if (use > 0) mbar_wait(s_u32(&ctl->empty[st]), (use - 1) & 1u);
      const uint32_t bytes = (uint32_t)rows * row_bytes;
      mbar_arrive_expect_tx(s_u32(&ctl->full[st]), bytes);
      bulk_g2s(s_u32(ring + (size_t)st * STAGE_BYTES), I.g.w + (size_t)(n0 + r0) * I.g.ldw, bytes, s_u32(&ctl->full[st]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ cluster barrier
__device__ __forceinline__ void cl_sync(Ctx& c) {
  compute_bar();                                    // this CTA's remote stores of the phase are issued
  const uint32_t k = c.sync_it++;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < CS; ++p)
      mbar_arrive_remote(c.peer_ctl[p] + (uint32_t)offsetof(Ctl, sync) + 8u * (k & 1u));
    // ONE thread acquires at cluster scope (an acquire of that scope invalidates the L1: 256 spinning threads would do
    // it on every poll); the CTA barrier below hands the ordering on to the other threads
    mbar_wait_cluster(s_u32(&c.ctl->sync[k & 1u]), (k >> 1) & 1u);
  }
  compute_bar();
}

// ------------------------------------------------------------------------------------------------ LayerNorm
constexpr int LN_PER_LANE = 24;                // columns per lane: d_model <= 768
template <int R>
__device__ void run_ln(const ClProgram& P, const ClInstr& I, Ctx& c, int warp, int lane) {
  if (warp < R) {
    const float* src = c.bufs + P.buf_off[I.ln.src_buf] + warp * P.buf_ld[I.ln.src_buf];
    float* dst = c.bufs + P.buf_off[I.ln.dst_buf] + warp * P.buf_ld[I.ln.dst_buf];
    const int n = I.ln.n;
    float v[LN_PER_LANE], g[LN_PER_LANE], bt[LN_PER_LANE];
#pragma unroll
    for (int j = 0; j < LN_PER_LANE; ++j) {          // every load of the row and of the affine parameters is independent
      const int i = lane + 32 * j;
      v[j] = i < n ? src[i] : 0.f;
      g[j] = i < n ? __ldg(I.ln.gamma + i) : 0.f;
      bt[j] = i < n ? __ldg(I.ln.beta + i) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_PER_LANE; ++j) s += v[j];
    const float mean = warp_sum(s) / (float)n;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_PER_LANE; ++j) { const float dlt = (lane + 32 * j < n) ? v[j] - mean : 0.f; q = fmaf(dlt, dlt, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)n + 1e-5f);
#pragma unroll
    for (int j = 0; j < LN_PER_LANE; ++j) {
      const int i = lane + 32 * j;
      if (i < n) dst[i] = (v[j] - mean) * rstd * g[j] + bt[j];
    }
  }
  compute_bar();
}

// ------------------------------------------------------------------------------------------------ projections
template <int R>
__device__ __forceinline__ void finish_value(const ClProgram& P, const ClInstr& I, Ctx& c, int n, int n0, int r, float v) {
  if (I.g.bias != nullptr) v += __ldg(I.g.bias + n);
  v = apply_act(v, I.act);
  if (I.g.residual) v += c.bufs[P.buf_off[CB_X] + r * P.buf_ld[CB_X] + n];
  if (I.g.out_kind == CO_LOCAL) {
    c.bufs[P.buf_off[I.g.out_buf] + r * P.buf_ld[I.g.out_buf] + (n - n0)] = v;
  } else if (I.g.out_kind == CO_GATHER) {
    const uint32_t off = (uint32_t)(P.buf_off[I.g.out_buf] + r * P.buf_ld[I.g.out_buf] + n) * 4u;
#pragma unroll
    for (int p = 0; p < CS; ++p) st_cluster_f32(c.peer_bufs[p] + off, v);
  } else if (r < c.rows_valid) {
    __stcg(I.g.out_global + (size_t)(c.row_base + r) * I.g.out_ld + n, v);
  }
}

template <int R, int CPL>
__device__ void run_gemv(const ClProgram& P, const ClInstr& I, Ctx& c, int warp, int lane) {
  int n0, nc;
  cols_of_rank(I, c.rank, n0, nc);
  if (nc == 0) return;
  const int K = I.g.K, row_bytes = I.g.ldw * 2;
  const int nseg = K <= KSEG_MAX ? 1 : 4, kseg = K / nseg, nch = kseg >> 3;
  const int seg = warp % nseg;
  const int rps = rows_per_stage(I);
  // this lane's slice of the R input rows, kept in registers for every weight row of the instruction
  float xr[R][CPL][8];
  {
    const float* in = c.bufs + P.buf_off[I.g.in_buf] + seg * kseg;
    const int ld = P.buf_ld[I.g.in_buf];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int ch = lane + 32 * i;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (ch < nch) {
          const float4 a = *reinterpret_cast<const float4*>(in + r * ld + ch * 8);
          const float4 b = *reinterpret_cast<const float4*>(in + r * ld + ch * 8 + 4);
          xr[r][i][0] = a.x; xr[r][i][1] = a.y; xr[r][i][2] = a.z; xr[r][i][3] = a.w;
          xr[r][i][4] = b.x; xr[r][i][5] = b.y; xr[r][i][6] = b.z; xr[r][i][7] = b.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) xr[r][i][e] = 0.f;
        }
      }
    }
  }
  constexpr int IPW = R == 4 ? 2 : 4;          // weight rows (items) a warp multiplies at once: IPW x R independent FMA chains
  constexpr int S = IPW * R;                   // sums a warp reduces together (4 or 8)
  for (int r0 = 0; r0 < nc; r0 += rps) {
    const int rows = min(rps, nc - r0);
    const uint32_t it = c.cons_it++, st = it % NSTAGE;
    mbar_wait(s_u32(&c.ctl->full[st]), (it / NSTAGE) & 1u);
    const unsigned char* sbase = c.ring + (size_t)st * STAGE_BYTES + (size_t)seg * kseg * 2;
    float* part = c.ctl->part[it & 1u];
    const int n_items = rows * nseg;
    for (int item0 = warp; item0 < n_items && P.exp != 1; item0 += NW * IPW) {
      // items item0 + NW * j (j < IPW): same K segment (NW % nseg == 0), weight rows (item / nseg)
      float acc[S];
#pragma unroll
      for (int q = 0; q < S; ++q) acc[q] = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nch) {
          uint4 w[IPW];
#pragma unroll
          for (int j = 0; j < IPW; ++j) {
            const int item = item0 + NW * j;
            w[j] = item < n_items ? *reinterpret_cast<const uint4*>(sbase + (size_t)(item / nseg) * row_bytes + ch * 16) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int j = 0; j < IPW; ++j) {
            const float wf[8] = {__uint_as_float(w[j].x << 16), __uint_as_float(w[j].x & 0xffff0000u), __uint_as_float(w[j].y << 16),
                                 __uint_as_float(w[j].y & 0xffff0000u), __uint_as_float(w[j].z << 16), __uint_as_float(w[j].z & 0xffff0000u),
                                 __uint_as_float(w[j].w << 16), __uint_as_float(w[j].w & 0xffff0000u)};
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[j * R + r] = fmaf(wf[e], xr[r][i][e], acc[j * R + r]);
          }
        }
      }
      // transposing reduction: at every level a lane keeps half of its sums and hands the other half to lane ^ off, so
      // S sums cost S - 1 + (5 - log2 S) shuffles instead of 5 S; sum q ends up in the lanes whose bits 4.. spell q
      int cur = S;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        if (cur > 1) {
          const bool upper = (lane & off) != 0;
          const int half = cur >> 1;
#pragma unroll
          for (int q = 0; q < S / 2; ++q) {
            if (q < half) {
              const float send = upper ? acc[q] : acc[q + half];
              const float keep = upper ? acc[q + half] : acc[q];
              acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          cur = half;
        } else {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
        }
      }
      constexpr int LOGS = S == 4 ? 2 : 3;
      const int q_mine = lane >> (5 - LOGS);                           // index of the sum this lane holds
      if ((lane & ((1 << (5 - LOGS)) - 1)) == 0) {
        const int j = q_mine / R, r = q_mine % R;
        const int item = item0 + NW * j;
        if (item < n_items) {
          const int rr = item / nseg;
          if (P.exp == 2) { if (acc[0] == 123.456f) part[0] = acc[0]; }      // experiment: multiply, do not finish
          else if (nseg == 1) finish_value<R>(P, I, c, n0 + r0 + rr, n0, r, acc[0]);
          else part[(rr * 4 + seg) * R + r] = acc[0];
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(s_u32(&c.ctl->empty[st]));           // this warp no longer reads the stage
    if (nseg > 1) {
      compute_bar();
      const int t = (int)threadIdx.x;
      if (t < rows * R) {
        const int rr = t / R, r = t % R;
        const float* pp = part + (rr * 4) * R + r;
        const float v = ((pp[0] + pp[R]) + pp[2 * R]) + pp[3 * R];       // fixed order over the four K segments
        finish_value<R>(P, I, c, n0 + r0 + rr, n0, r, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// (row, head) items of this CTA, NW / items warps each: the warps of an item split its positions, so every K / V row of
// the item is requested in ONE round of loads (the phase is pure load latency).  fp32, reference order:
// (q.k) * scale -> softmax (exp(x - max) / sum) -> sum_p w_p v_p; the warps' partial P*V sums are added in a fixed order.
constexpr int AUN = 8;                         // positions in flight per lane group
template <int R>
__device__ void run_attention(const ClProgram& P, const ClInstr& I, Ctx& c, int step, int warp, int lane) {
  const int hpc = P.nhead / CS, hd = P.hd;
  const bool self = I.kind == CL_ATTN_SELF;
  const int n_items = R * hpc, wpi = NW / n_items;                // warps per item (NW is a multiple of every R * hpc we accept)
  const int item = warp / wpi, part = warp % wpi;
  const int r = item / hpc, hh = item % hpc, h = c.rank * hpc + hh;
  const int b = c.row_base + min(r, c.rows_valid - 1);             // rows beyond the batch repeat the last valid row (never stored)
  const int n = self ? step + 1 : I.at.fixed_len;
  int lpp = 1;
  while (lpp * 4 < hd) lpp <<= 1;                                   // lanes per position (power of two >= hd / 4)
  const int ppi = 32 / lpp, grp = lane / lpp, e0 = 4 * (lane % lpp);
  const bool e_ok = e0 < hd;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float* sc = c.ctl->sc[item];                                      // scores of the item, shared by its warps
  float* pacc = c.ctl->pacc[item * wpi];                            // partial P*V sums of the item's warps
  int* pg = c.ctl->pg[warp];
  const float* qb = c.bufs + P.buf_off[CB_Q] + r * P.buf_ld[CB_Q] + hh * hd;
  const float4 q4 = e_ok ? *reinterpret_cast<const float4*>(qb + e0) : zero4;
  if (self) {                                                       // page ids of the row: one load per page
    const int* pt = P.page_table + (size_t)b * P.pages_per_seq;
    const int npg = (n + kPagePos - 1) >> kPageShift;
    if (lane < npg) pg[lane] = __ldcg(pt + lane);
    __syncwarp();
  }
  auto row_off = [&](int p) -> size_t {
    if (self) return (size_t)pg[p >> kPageShift] * I.at.page_stride + (size_t)(p & (kPagePos - 1)) * I.at.row_stride + h * hd;
    return (size_t)b * I.at.seq_stride + (size_t)p * I.at.row_stride + h * hd;
  };
  float4 knew = zero4, vnew = zero4;
  if (self) {
    const float* kb = c.bufs + P.buf_off[CB_K] + r * P.buf_ld[CB_K] + hh * hd;
    const float* vb = c.bufs + P.buf_off[CB_V] + r * P.buf_ld[CB_V] + hh * hd;
    if (e_ok) { knew = *reinterpret_cast<const float4*>(kb + e0); vnew = *reinterpret_cast<const float4*>(vb + e0); }
    if (part == 0 && grp == 0 && e_ok && r < c.rows_valid) {        // append (torch.cat in the reference, :1266-1267)
      const size_t off = row_off(n - 1) + e0;
      __stcg(reinterpret_cast<float4*>(I.at.kcache + off), knew);
      __stcg(reinterpret_cast<float4*>(I.at.vcache + off), vnew);
    }
  }
  // this warp's positions: p = start + (k * AUN + u) * stride; the K (and, when one batch covers the item, the V) rows of
  // a batch are requested together
  const int stride = wpi * ppi, start = part * ppi + grp, per_batch = stride * AUN;
  const int nb = (n + per_batch - 1) / per_batch;                   // same for every warp of the CTA
  float4 v4[AUN];
  for (int k = 0; k < nb; ++k) {
    float4 k4[AUN];
#pragma unroll
    for (int u = 0; u < AUN; ++u) {
      const int p = start + (k * AUN + u) * stride;
      k4[u] = zero4;
      if (nb == 1) v4[u] = zero4;
      if (p < n && e_ok) {
        if (self && p == n - 1) { k4[u] = knew; if (nb == 1) v4[u] = vnew; }
        else {
          const size_t off = row_off(p) + e0;
          k4[u] = __ldcg(reinterpret_cast<const float4*>(I.at.kcache + off));
          if (nb == 1) v4[u] = __ldcg(reinterpret_cast<const float4*>(I.at.vcache + off));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < AUN; ++u) {
      const int p = start + (k * AUN + u) * stride;
      float dsum = fmaf(q4.x, k4[u].x, fmaf(q4.y, k4[u].y, fmaf(q4.z, k4[u].z, q4.w * k4[u].w)));
      for (int o = lpp >> 1; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      if ((lane % lpp) == 0 && p < n) sc[p] = dsum * P.scale;
    }
  }
  compute_bar();                                // all scores of every item are in shared memory
  float m = -INFINITY;
  for (int p = lane; p < n; p += 32) m = fmaxf(m, sc[p]);
  m = warp_max(m);
  float sum = 0.f;
  for (int p = lane; p < n; p += 32) sum += expf(sc[p] - m);
  sum = warp_sum(sum);
  float4 acc = zero4;
  for (int k = 0; k < nb; ++k) {
    if (nb > 1) {
#pragma unroll
      for (int u = 0; u < AUN; ++u) {
        const int p = start + (k * AUN + u) * stride;
        v4[u] = zero4;
        if (p < n && e_ok) v4[u] = (self && p == n - 1) ? vnew : __ldcg(reinterpret_cast<const float4*>(I.at.vcache + row_off(p) + e0));
      }
    }
#pragma unroll
    for (int u = 0; u < AUN; ++u) {
      const int p = start + (k * AUN + u) * stride;
      if (p < n && e_ok) {
        const float w = expf(sc[p] - m) / sum;
        acc.x = fmaf(w, v4[u].x, acc.x); acc.y = fmaf(w, v4[u].y, acc.y); acc.z = fmaf(w, v4[u].z, acc.z); acc.w = fmaf(w, v4[u].w, acc.w);
      }
    }
  }
  for (int o = lpp; o < 32; o <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  float* mine = pacc + part * 128;              // [wpi][128] partial sums of the item
  if (grp == 0 && e_ok) *reinterpret_cast<float4*>(mine + e0) = acc;
  compute_bar();
  if (part == 0 && grp == 0 && e_ok) {          // fixed order over the item's warps; the head's slice goes to every CTA
    float4 o4 = *reinterpret_cast<const float4*>(pacc + e0);
    for (int w = 1; w < wpi; ++w) {
      const float4 t4 = *reinterpret_cast<const float4*>(pacc + w * 128 + e0);
      o4.x += t4.x; o4.y += t4.y; o4.z += t4.z; o4.w += t4.w;
    }
    const uint32_t off = (uint32_t)(P.buf_off[CB_A] + r * P.buf_ld[CB_A] + h * hd + e0) * 4u;
#pragma unroll
    for (int p = 0; p < CS; ++p) st_cluster_v4(c.peer_bufs[p] + off, o4);
  }
}

// ------------------------------------------------------------------------------------------------ one step
template <int R, int CPL>
__device__ void run_program(const ClProgram& P, Ctx& c, int step, int warp, int lane) {
  const bool timing = P.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t_prev = timing ? globaltimer_ns() : 0ull;
  for (int i = 0; i < P.n_instr; ++i) {
    const ClInstr& I = c.prog[i];
    if (timing && i > 0) { const unsigned long long t = globaltimer_ns(); P.dbg[i - 1] += t - t_prev; t_prev = t; }
    switch (I.kind) {
      case CL_LN: run_ln<R>(P, I, c, warp, lane); break;
      case CL_GEMV: run_gemv<R, CPL>(P, I, c, warp, lane); break;
      case CL_ATTN_SELF:
      case CL_ATTN_CROSS:
        compute_bar();                 // q / k / v of this CTA's head(s) are complete
        run_attention<R>(P, I, c, step, warp, lane);
        break;
      default: cl_sync(c); break;
    }
  }
  if (timing) P.dbg[P.n_instr - 1] += globaltimer_ns() - t_prev;
}

template <int R>
__device__ void load_x_rows(const ClProgram& P, Ctx& c) {
  float* X = c.bufs + P.buf_off[CB_X];
  for (int i = threadIdx.x; i < R * P.d; i += NW * 32) {
    const int r = i / P.d, col = i - r * P.d;
    X[r * P.buf_ld[CB_X] + col] = r < c.rows_valid ? __ldcg(P.x_global + (size_t)(c.row_base + r) * P.d + col) : 0.f;
  }
  compute_bar();
}
template <int R>
__device__ void store_x_rows(const ClProgram& P, Ctx& c) {
  if (c.rank != 0) return;
  const float* X = c.bufs + P.buf_off[CB_X];
  for (int i = threadIdx.x; i < c.rows_valid * P.d; i += NW * 32) {
    const int r = i / P.d, col = i - r * P.d;
    __stcg(P.x_global + (size_t)(c.row_base + r) * P.d + col, X[r * P.buf_ld[CB_X] + col]);
  }
}

// Common prologue: shared-memory carve-up, barrier init, peer addresses.  Returns false for clusters without rows.
template <int R>
__device__ __forceinline__ bool cl_setup(const ClProgram& P, unsigned char* smem, Ctx& c) {
  c.ring = smem;
  c.bufs = reinterpret_cast<float*>(smem + (size_t)NSTAGE * STAGE_BYTES);
  c.ctl = reinterpret_cast<Ctl*>(smem + (size_t)NSTAGE * STAGE_BYTES + (size_t)P.buf_floats * 4);
  {                             // the program is read thousands of times per step: keep it in shared memory
    ClInstr* dst = reinterpret_cast<ClInstr*>(reinterpret_cast<unsigned char*>(c.ctl) + sizeof(Ctl));
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(P.instr);
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(dst);
    const int words = P.n_instr * (int)(sizeof(ClInstr) / 4);
    for (int i = threadIdx.x; i < words; i += CL_THREADS) dst32[i] = __ldg(src32 + i);
    c.prog = dst;
  }
  c.rank = (int)cluster_rank();
  c.row_base = ((int)blockIdx.x / CS) * R;
  c.rows_valid = min(R, P.B - c.row_base);
  c.cons_it = 0; c.sync_it = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(s_u32(&c.ctl->full[s]), 1); mbar_init(s_u32(&c.ctl->empty[s]), NW); }
    mbar_init(s_u32(&c.ctl->sync[0]), CS);
    mbar_init(s_u32(&c.ctl->sync[1]), CS);
    c.ctl->go = 0; c.ctl->cont = 1;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#pragma unroll
  for (int p = 0; p < CS; ++p) {
    c.peer_bufs[p] = mapa_u32(s_u32(c.bufs), (uint32_t)p);
    c.peer_ctl[p] = mapa_u32(s_u32(c.ctl), (uint32_t)p);
  }
  for (int i = threadIdx.x; i < P.buf_floats; i += CL_THREADS) c.bufs[i] = 0.f;      // rows beyond the batch stay finite
  cluster_sync_all();                         // every CTA's barriers exist before anybody arrives on them remotely
  return c.rows_valid > 0;
}

template <int R, int CPL>
__global__ void __launch_bounds__(CL_THREADS, 1) decode_cluster_step_kernel(ClProgram P) {
  extern __shared__ __align__(128) unsigned char cl_smem[];
  Ctx c;
  const bool active = cl_setup<R>(P, cl_smem, c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool done = P.st->done != 0;
  if (active && !done) {
    if (warp == NW) {
      uint32_t it = 0;
      if (lane == 0) producer_step(P, c.prog, c.ctl, c.ring, c.rank, it);
      __syncwarp();                           // the cluster barrier below is .aligned: the warp must be converged
    } else {
      const int step = P.st->step;
      load_x_rows<R>(P, c);
      run_program<R, CPL>(P, c, step, warp, lane);
      store_x_rows<R>(P, c);
    }
  }
  cluster_sync_all();                         // nobody exits while a peer may still write into its shared memory
}

// Next-token embedding of the cluster's rows into this CTA's copy of the residual stream (reference :1405, :1233)
template <int R>
__device__ void embed_rows(const ClProgram& P, const SmallTail& t, Ctx& c, int pos) {
  float* X = c.bufs + P.buf_off[CB_X];
  for (int i = threadIdx.x; i < R * P.d; i += NW * 32) {
    const int r = i / P.d, col = i - r * P.d;
    const int tok = c.ctl->tok[r];
    X[r * P.buf_ld[CB_X] + col] = r < c.rows_valid ? __bfloat162float(t.emb[(size_t)tok * t.ld_emb + col]) + __ldg(t.pe + (size_t)pos * P.d + col) : 0.f;
  }
  compute_bar();
}

template <int R, int CPL>
__global__ void __launch_bounds__(CL_THREADS, 1) decode_cluster_persist_kernel(ClProgram P, SmallTail tail) {
  extern __shared__ __align__(128) unsigned char cl_smem[];
  Ctx c;
  const bool active = cl_setup<R>(P, cl_smem, c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  StepState* st = P.st;
  // (no "done" test here: when fewer clusters are resident than launched, a late cluster starts after the others have
  // finished and set the flag; its rows still have to be decoded.  One launch = one decode, the flag is 0 at launch.)
  if (active) {
    const int step0 = 0;
    if (warp == NW) {
      // producer: stream a step's weights, then wait until the consumers commit to the next step (or stop)
      if (lane == 0) {
        uint32_t it = 0;
        for (int step = step0;; ++step) {
          producer_step(P, c.prog, c.ctl, c.ring, c.rank, it);
          volatile int* go = &c.ctl->go;
          int g;
          while ((g = *go) == step - step0) __nanosleep(64);
          if (g < 0) break;
        }
      }
      __syncwarp();
    } else {
      load_x_rows<R>(P, c);
      for (int step = step0;; ++step) {
        run_program<R, CPL>(P, c, step, warp, lane);
        store_x_rows<R>(P, c);
        cl_sync(c);                            // logits / head outputs of every CTA are in global memory
        // greedy epilogue (:1415-1548) of row r on CTA r % 8, one warp; the token goes to every CTA of the cluster
        for (int r = c.rank; r < c.rows_valid; r += CS) {
          if (warp == 0) {
            const int b = c.row_base + r;
            int tok = greedy_row_token(tail.sp, b, step, lane);
            if (lane == 0) {
              if (tail.sp.forced != nullptr) {
                const long long f = tail.sp.forced[(size_t)b * tail.sp.out_ld + step];
                if (f >= 0) tok = (int)f;
              }
              if (tok == kEndIdx && tail.sp.finished[b] == 0) {      // publish the finishing step BEFORE the row count drops
                atomicMax(&st->pad[0], step + 1);
                __threadfence();
              }
              tok = commit_token(tail.sp, b, step, tok, 0.f);
              if (step + 1 < tail.max_steps && ((step + 1) & (kPagePos - 1)) == 0)       // KV page of the next 16 positions
                tail.page_table[b * tail.pages_per_seq + ((step + 1) >> kPageShift)] = atomicAdd(&st->next_free_page, 1);
              __threadfence();
#pragma unroll
              for (int p = 0; p < CS; ++p) st_cluster_u32(c.peer_ctl[p] + (uint32_t)offsetof(Ctl, tok) + 4u * r, (uint32_t)tok);
            }
          }
        }
        cl_sync(c);                            // tokens (and KV pages) of the cluster's rows are published
        if (c.rank == 0 && threadIdx.x == 0) {
          // every row of the BATCH has emitted END and this cluster has decoded up to the last finishing step
          const int unfinished = ld_acquire_s32(&st->n_unfinished);
          const int last = ld_acquire_s32(&st->pad[0]);
          const bool stop = (unfinished <= 0 && step + 1 >= last) || step + 1 >= tail.max_steps;
#pragma unroll
          for (int p = 0; p < CS; ++p) st_cluster_u32(c.peer_ctl[p] + (uint32_t)offsetof(Ctl, cont), stop ? 0u : 1u);
          if (stop) {                          // every cluster ends with the same values
            st->step = step + 1;
            st->out_len = unfinished <= 0 ? max(last, 1) : step + 1;
            st->degenerate = 0;
            __threadfence();
            st->done = 1;
          }
        }
        cl_sync(c);
        const bool cont = *reinterpret_cast<volatile int*>(&c.ctl->cont) != 0;
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&c.ctl->go) = cont ? (step + 1 - step0) : -1;
        if (!cont) break;
        embed_rows<R>(P, tail, c, step + 1);
      }
    }
  }
  cluster_sync_all();
}

size_t smem_bytes_for(const ClProgram& p) {
  return (size_t)NSTAGE * STAGE_BYTES + (size_t)p.buf_floats * 4 + sizeof(Ctl) + (size_t)p.n_instr * sizeof(ClInstr) + 128;
}

template <typename K, typename... Args>
int launch_cluster(K kernel, int n_clusters, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * CS); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SCV_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  return 0;
}

}  // namespace

size_t cluster_smem_bytes(const ClProgram& p) { return smem_bytes_for(p); }

int cluster_max_active(int rows_per_cluster, size_t smem) {
  (void)rows_per_cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(64 * CS); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  cudaFuncSetAttribute(decode_cluster_persist_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (cudaOccupancyMaxActiveClusters(&n, decode_cluster_persist_kernel<4, 3>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

bool cluster_shape_ok(int d, int nhead, int dff, int V, int pe_len, int n_memory) {
  auto k_ok = [](int K) { return K % 8 == 0 && K >= 64 && (K <= KSEG_MAX || (K % 32 == 0 && K / 4 <= KSEG_MAX)); };
  const int hd = nhead > 0 ? d / nhead : 0;
  return nhead % CS == 0 && (nhead / CS) * kClMaxRows <= NW && hd % 4 == 0 && hd <= 128 && d % (8 * CS) == 0 && dff % (8 * CS) == 0 &&
         V % CS == 0 && (d / 4) % CS == 0 && k_ok(d) && k_ok(dff) && k_ok(d / 4) && std::max(pe_len, n_memory) <= 256 &&
         d <= 32 * LN_PER_LANE;
}

#define SCV_CL_FOR_EACH(X) X(1, 2) X(1, 3) X(2, 2) X(2, 3) X(4, 2) X(4, 3)

int launch_decode_cluster_step(const ClProgram& p, cudaStream_t s) {
  const int R = p.rows_per_cluster, ncl = ceil_div(p.B, R);
  const size_t smem = smem_bytes_for(p);
  SCV_REQUIRE((R == 1 || R == 2 || R == 4) && (p.cpl == 2 || p.cpl == 3) && smem <= 227 * 1024,
              "cluster decode: %d rows per cluster / %zu bytes of shared memory", R, smem);
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
#define X(RR, CC) SCV_CUDA(cudaFuncSetAttribute(decode_cluster_step_kernel<RR, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SCV_CL_FOR_EACH(X)
#undef X
  }
#define X(RR, CC) if (R == RR && p.cpl == CC) SCV_TRY(launch_cluster(decode_cluster_step_kernel<RR, CC>, ncl, smem, s, p));
  SCV_CL_FOR_EACH(X)
#undef X
  SCV_LAUNCH_CHECK();
  return 0;
}

int launch_decode_cluster_persist(const ClProgram& p, const SmallTail& tail, cudaStream_t s) {
  const int R = p.rows_per_cluster, ncl = ceil_div(p.B, R);
  const size_t smem = smem_bytes_for(p);
  SCV_REQUIRE((R == 1 || R == 2 || R == 4) && (p.cpl == 2 || p.cpl == 3) && smem <= 227 * 1024,
              "cluster decode: %d rows per cluster / %zu bytes of shared memory", R, smem);
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
#define X(RR, CC) SCV_CUDA(cudaFuncSetAttribute(decode_cluster_persist_kernel<RR, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SCV_CL_FOR_EACH(X)
#undef X
  }
#define X(RR, CC) if (R == RR && p.cpl == CC) SCV_TRY(launch_cluster(decode_cluster_persist_kernel<RR, CC>, ncl, smem, s, p, tail));
  SCV_CL_FOR_EACH(X)
#undef X
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
