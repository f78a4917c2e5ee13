// Per-step decode kernels: token embedding + positional encoding, warp-level flash-decode attention over
// the paged self-attention KV cache (and over the per-layer projected memory tokens for cross-attention),
// and the fused sampling epilogue (type mask, stop head, degenerate guard, entropy, temperature,
// argmax / Philox multinomial, log-prob, finished bookkeeping).
#include <limits.h>

#include <cstdlib>

#include "../../include/scvae_b200.h"
#include "decode_kernels.cuh"
#include "sampler_device.cuh"

namespace scv {

// ---------------------------------------------------------------------------------------------
// x = token_embedding[cur] + pe[step]   (models/autoregressive_decoder.py:1405, 1233)
// Also grows the row's KV page list when the step crosses a page boundary.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_kernel(EmbedArgs a) {
  pdl_wait();
  if (a.st->done) return;
  pdl_launch_dependents();
  const int b = blockIdx.x;
  const int step = a.st->step;
  int ob = b;
  if (a.row_map != nullptr) {                 // finished rows are compacted away: slot -> row
    if (a.slot_base + b >= a.st->pad[1]) return;
    ob = a.row_map[a.slot_base + b];
  }
  if (threadIdx.x == 0 && (step & (kPagePos - 1)) == 0) {
    a.page_table[ob * a.pages_per_seq + (step >> kPageShift)] = atomicAdd(&a.st->next_free_page, 1);
  }
  const int tok = a.cur_tokens[b];
  const __nv_bfloat16* row = a.table + (size_t)tok * a.ld_table;
  const float* pe = a.pe + (size_t)step * a.d;
  float* x = a.x + (size_t)b * a.d;
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) x[i] = __bfloat162float(row[i]) + pe[i];
}

int launch_embed(const EmbedArgs& a, cudaStream_t s) {
  ProfScope prof(PC_EMBED, s, 1.0 * a.B * a.d, 6.0 * a.B * a.d);
  SCV_CUDA(launch_k(embed_kernel, dim3(a.B), dim3(128), 0, s, a));
  SCV_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(128) embed_sequence_kernel(const __nv_bfloat16* __restrict__ table, int ld_table,
                                                             const float* __restrict__ pe, int d,
                                                             const long long* __restrict__ tokens, int ld_tokens, int L,
                                                             int vocab, float* __restrict__ x,
                                                             unsigned char* __restrict__ key_skip) {
  const int r = blockIdx.x, b = r / L, t = r - b * L;
  long long tok = tokens[(size_t)b * ld_tokens + t];
  if (threadIdx.x == 0) key_skip[r] = tok == 0 ? 1 : 0;          // PAD_IDX keys are masked (:952)
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);          // the host validated the ids; never index outside the table
  const __nv_bfloat16* row = table + (size_t)tok * ld_table;
  const float* per = pe + (size_t)t * d;
  float* xr = x + (size_t)r * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) xr[i] = __bfloat162float(row[i]) + per[i];
}

int launch_embed_sequence(const __nv_bfloat16* table, int ld_table, const float* pe, int d, const long long* tokens,
                          int ld_tokens, int B, int L, int vocab, float* x, unsigned char* key_skip, cudaStream_t s) {
  embed_sequence_kernel<<<B * L, 128, 0, s>>>(table, ld_table, pe, d, tokens, ld_tokens, L, vocab, x, key_skip);
  SCV_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Flash-decode attention for ONE query token per (row, head): one warp per (row, head).
//   self  (:1259-1290): append this step's k,v to the paged cache, attend over positions 0..step
//   cross (:1302-1307): attend over the n_memory projected memory tokens of this layer
// fp32 throughout (scores, softmax, accumulation), same operation order as the reference:
// (q.k) * scale -> softmax (exp(x-max)/sum) -> sum_p w_p v_p.
// Lane l owns head-dim elements l, l+32, ... so every cache row is read as coalesced 128-byte segments.
// ---------------------------------------------------------------------------------------------
template <int EPL>
__global__ void __launch_bounds__(256) attention_decode_kernel(AttnArgs a) {
  extern __shared__ float sc_all[];
  pdl_wait();
  if (a.st->done) return;
  pdl_launch_dependents();
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  if (gw >= a.B * a.nhead) return;
  const int b = gw / a.nhead, h = gw % a.nhead;
  const int hd = a.hd;
  const int n = a.fixed_len >= 0 ? a.fixed_len : a.st->step + 1;
  float* sc = sc_all + (size_t)warp_in_block * a.max_n;

  float qv[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int e = lane + 32 * j;
    qv[j] = e < hd ? a.q[(size_t)b * a.ldq + h * hd + e] : 0.f;
  }
  const bool paged = a.page_table != nullptr;
  const int* pt = paged ? a.page_table + (size_t)b * a.pages_per_seq : nullptr;
  auto row_off = [&](int p) -> size_t {
    if (paged) return (size_t)pt[p >> kPageShift] * a.page_stride + (size_t)(p & (kPagePos - 1)) * a.row_stride + h * hd;
    return (size_t)b * a.seq_stride + (size_t)p * a.row_stride + h * hd;
  };

  if (a.knew != nullptr) {   // append (torch.cat in the reference, :1266-1267)
    const size_t off = row_off(n - 1);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int e = lane + 32 * j;
      if (e < hd) {
        a.kcache[off + e] = a.knew[(size_t)b * a.ldn + h * hd + e];
        a.vcache[off + e] = a.vnew[(size_t)b * a.ldn + h * hd + e];
      }
    }
  }

  // scores
  for (int p0 = 0; p0 < n; p0 += 4) {
    float part[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u;
      float d = 0.f;
      if (p < n) {
        const float* kp = a.kcache + row_off(p);
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          const int e = lane + 32 * j;
          if (e < hd) d = fmaf(qv[j], kp[e], d);
        }
      }
      part[u] = d;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) part[u] = warp_sum(part[u]);
    float sel = part[0];
    if (lane == 1) sel = part[1];
    if (lane == 2) sel = part[2];
    if (lane == 3) sel = part[3];
    if (lane < 4 && p0 + lane < n) sc[p0 + lane] = sel * a.scale;
  }
  __syncwarp();
  float m = -INFINITY;
  for (int p = lane; p < n; p += 32) m = fmaxf(m, sc[p]);
  m = warp_max(m);
  float sum = 0.f;
  for (int p = lane; p < n; p += 32) {
    const float e = expf(sc[p] - m);
    sc[p] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  for (int p = lane; p < n; p += 32) sc[p] = sc[p] / sum;
  __syncwarp();

  float acc[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) acc[j] = 0.f;
  int p = 0;
  for (; p + 4 <= n; p += 4) {
    float vv[4][EPL];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* vp = a.vcache + row_off(p + u);
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const int e = lane + 32 * j;
        vv[u][j] = e < hd ? vp[e] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float w = sc[p + u];
#pragma unroll
      for (int j = 0; j < EPL; ++j) acc[j] = fmaf(w, vv[u][j], acc[j]);
    }
  }
  for (; p < n; ++p) {
    const float* vp = a.vcache + row_off(p);
    const float w = sc[p];
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int e = lane + 32 * j;
      if (e < hd) acc[j] = fmaf(w, vp[e], acc[j]);
    }
  }
  if (a.out_split == nullptr) {
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int e = lane + 32 * j;
      if (e < hd) a.out[(size_t)b * a.ldo + h * hd + e] = acc[j];
    }
  } else {
    // the only consumer is the tensor-core out-projection: emit bf16 hi/lo chunks in SplitTile form
    __syncwarp();
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int e = lane + 32 * j;
      if (e < hd) sc[e] = acc[j];
    }
    __syncwarp();
    if (lane * 8 < hd) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) split_pair(sc[lane * 8 + 2 * p], sc[lane * 8 + 2 * p + 1], hi[p], lo[p]);
      const int col = h * hd + lane * 8;           // head_dim is a multiple of 8 on this path
      const int mt = b >> 7, ri = b & 127, kb = col >> 6, cj = (col & 63) >> 3;
      uint8_t* dst = a.out_split + ((size_t)mt * a.kb_out + kb) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// Vectorised variant (head_dim % 4 == 0): LPP lanes cover one cache row with float4 loads, so one warp instruction
// reads 32 / LPP positions (512 B for head_dim 64) and the dot-product reduction runs over LPP lanes only.  The
// first version above issued one 4-byte load per lane and position and was instruction-issue bound (ncu r01b:
// issue slots 49 % busy at 47 % of DRAM peak); this one moves the same bytes with ~4x fewer instructions.
// Registers: left alone the compiler takes 48 (5 CTAs = 40 warps per SM); the kernel waits on memory (ncu: 10-17 warps stalled
// on the long scoreboard per issued instruction, 55 % warps active), so it is capped at 40 registers = 6 CTAs per SM (8 bytes
// spilled).  Same box, alternating builds: 67.1 ms per 4096-latent decode against 68.9 (63-step decode 225.5 against 230.8);
// 32 registers (8 CTAs, 52 bytes spilled) 67.2-67.9.  -DSCV_ATTN_MIN_CTAS=0 restores the uncapped build.
#ifndef SCV_ATTN_MIN_CTAS
#define SCV_ATTN_MIN_CTAS 6
#endif
#if SCV_ATTN_MIN_CTAS > 0
#define SCV_ATTN_BOUNDS __launch_bounds__(256, SCV_ATTN_MIN_CTAS)
#else
#define SCV_ATTN_BOUNDS __launch_bounds__(256)
#endif
template <int LPP>
__global__ void SCV_ATTN_BOUNDS attention_decode_v4_kernel(AttnArgs a) {
  extern __shared__ float sc_all[];
  pdl_wait();
  if (a.st != nullptr && a.st->done) return;
  constexpr int PPI = 32 / LPP;
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The grid may be smaller than the work (a few CTAs per SM, launch_attention): CTAs walk the (row, head) items with a
  // grid stride, so this HBM-bound kernel leaves registers / shared memory on every SM for the tensor-core projections of
  // another sub-batch's stream to run beside it ("co-residency").
  const int n_items = a.B * a.nhead, wpb = blockDim.x >> 5;
  TraceRec* trc = threadIdx.x == 0 ? trace_begin(a.trace, a.fixed_len >= 0 ? 2u : 1u) : nullptr;
  // shared memory tokens: when the launch holds whole groups of seq_mod rows, item i is sample i % k of latent i / k
  const int k_local = (a.seq_mod > 0 && a.row_map == nullptr && a.B % a.seq_mod == 0 && a.slot_base % a.seq_mod == 0) ? a.B / a.seq_mod : 0;
  for (int gw = blockIdx.x * wpb + warp_in_block; gw < n_items; gw += gridDim.x * wpb) {
  const int bi = gw / a.nhead, h = gw % a.nhead;
  const int b = k_local > 1 ? (bi % k_local) * a.seq_mod + bi / k_local : bi;
  const int hd = a.hd;
  if (a.row_map != nullptr && a.slot_base + b >= a.st->pad[1]) continue;       // slot of a finished row (compaction)
  // sequence whose K / V this query row attends to
  int sb = a.row_map != nullptr ? a.row_map[a.slot_base + b] : (a.rows_per_seq > 0 ? b / a.rows_per_seq : b);
  if (a.seq_mod > 0) sb = (a.row_map != nullptr ? sb : a.slot_base + b) % a.seq_mod;
  const int n = a.fixed_len >= 0 ? a.fixed_len : (a.rows_per_seq > 0 ? b - sb * a.rows_per_seq + 1 : a.st->step + 1);
  float* sc = sc_all + (size_t)warp_in_block * a.max_n;
  const int grp = lane / LPP, e0 = 4 * (lane % LPP);
  const bool e_ok = e0 < hd;

  const bool paged = a.page_table != nullptr;
  const int* pt = paged ? a.page_table + (size_t)(a.row_map != nullptr ? a.row_map[a.slot_base + b] : b) * a.pages_per_seq : nullptr;
  // The sequence's page ids are read ONCE, lane l holding page l (the table was filled by the step's first kernel): a lookup
  // per iteration put a dependent L2 round trip in front of every batch of K / V loads, five per item at 12 positions, in a
  // kernel whose time is (dependent round trips per item) x (waves of warps).
  const bool pages_in_regs = paged && a.pages_per_seq <= 32 && a.pages_regs != 0;
  int my_page = 0;
  if (pages_in_regs && lane < a.pages_per_seq && lane <= ((n - 1) >> kPageShift)) my_page = pt[lane];
  auto page_of = [&](int p) -> int {          // p uniform over the warp
    return pages_in_regs ? __shfl_sync(0xffffffffu, my_page, p >> kPageShift) : pt[p >> kPageShift];
  };
  auto row_off = [&](int p) -> size_t {
    if (paged) return (size_t)pt[p >> kPageShift] * a.page_stride + (size_t)(p & (kPagePos - 1)) * a.row_stride + h * hd;
    return (size_t)sb * a.seq_stride + (size_t)p * a.row_stride + h * hd;
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 q4 = e_ok ? *reinterpret_cast<const float4*>(a.q + (size_t)b * a.ldq + h * hd + e0) : zero4;

  if (a.knew != nullptr) {   // append (torch.cat in the reference, :1266-1267)
    const size_t new_off = paged ? (size_t)page_of(n - 1) * a.page_stride + (size_t)((n - 1) & (kPagePos - 1)) * a.row_stride + h * hd
                                 : row_off(n - 1);
    if (grp == 0 && e_ok) {
      const size_t off = new_off + e0;
      *reinterpret_cast<float4*>(a.kcache + off) = *reinterpret_cast<const float4*>(a.knew + (size_t)b * a.ldn + h * hd + e0);
      *reinterpret_cast<float4*>(a.vcache + off) = *reinterpret_cast<const float4*>(a.vnew + (size_t)b * a.ldn + h * hd + e0);
      // (plain stores: the row is read back a few lines below by the other lanes of this warp)
    }
    __syncwarp();            // the appended row is read below by other lanes of this warp
  }

  // scores: UNR * PPI positions in flight per warp (UNR 2 and 8 both measured ~2.5 % slower per decode: fewer
  // loads in flight, or fewer resident warps)
  constexpr int UNR = 4;
  // ncu (r01n): issue slots 61-73 % busy at 46-77 % of the DRAM peak, i.e. instruction issue limits this kernel as much
  // as HBM does.  Two savings: (1) one page lookup and one 64-bit base per iteration (the UNR * PPI positions of an
  // iteration never straddle a 16-position page), (2) for 16 lanes per position the four dot products of an iteration
  // are reduced together (5 shuffles instead of 16; same pairing order, so the sums are bit-identical).
  constexpr bool kFastAddr = UNR * PPI <= kPagePos;
  const long long kv_delta = a.vcache - a.kcache;
  auto iter_base = [&](int p0) -> const float* {     // K row of position p0 for this lane's head slice
    if (paged) return a.kcache + (size_t)page_of(p0) * a.page_stride + (size_t)(p0 & (kPagePos - 1)) * a.row_stride + h * hd + e0;
    return a.kcache + (size_t)sb * a.seq_stride + (size_t)p0 * a.row_stride + h * hd + e0;
  };
  for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
    float4 kv[UNR];
    const float* kb = kFastAddr ? iter_base(p0) : nullptr;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * PPI + grp;
      if (kFastAddr) kv[u] = (p < n && e_ok) ? __ldcs(reinterpret_cast<const float4*>(kb + (u * PPI + grp) * a.row_stride)) : zero4;
      else kv[u] = (p < n && e_ok) ? __ldcs(reinterpret_cast<const float4*>(a.kcache + row_off(p) + e0)) : zero4;
    }
    float d[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) d[u] = fmaf(q4.x, kv[u].x, fmaf(q4.y, kv[u].y, fmaf(q4.z, kv[u].z, q4.w * kv[u].w)));
    if constexpr (LPP == 16 && UNR == 4) {
      const int sub = lane & 15;
      const bool up8 = (sub & 8) != 0, up4 = (sub & 4) != 0;
      // level 1 (partner lane ^ 8): lower half keeps positions u = 0, 1, upper half keeps u = 2, 3
      const float r0 = __shfl_xor_sync(0xffffffffu, up8 ? d[0] : d[2], 8);
      const float r1 = __shfl_xor_sync(0xffffffffu, up8 ? d[1] : d[3], 8);
      const float e0s = (up8 ? d[2] : d[0]) + r0, e1s = (up8 ? d[3] : d[1]) + r1;
      // level 2 (partner lane ^ 4): keep one of the two
      float f = (up4 ? e1s : e0s) + __shfl_xor_sync(0xffffffffu, up4 ? e0s : e1s, 4);
      f += __shfl_xor_sync(0xffffffffu, f, 2);
      f += __shfl_xor_sync(0xffffffffu, f, 1);
      const int u_mine = 2 * (sub >> 3) + ((sub >> 2) & 1);
      const int p = p0 + u_mine * PPI + grp;
      if ((sub & 3) == 0 && p < n) sc[p] = f * a.scale;
    } else {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float dd = d[u];
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
        const int p = p0 + u * PPI + grp;
        if ((lane % LPP) == 0 && p < n) sc[p] = dd * a.scale;
      }
    }
  }
  __syncwarp();
  if (a.key_skip != nullptr) {                 // padded keys take no weight (masked_fill(-inf) before the softmax)
    const unsigned char* ks = a.key_skip + (size_t)sb * a.rows_per_seq;
    for (int p = lane; p < n; p += 32)
      if (ks[p] != 0) sc[p] = -INFINITY;
    __syncwarp();
  }
  float m = -INFINITY;
  for (int p = lane; p < n; p += 32) m = fmaxf(m, sc[p]);
  m = warp_max(m);
  float sum = 0.f;
  for (int p = lane; p < n; p += 32) {
    const float e = expf(sc[p] - m);
    sc[p] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  for (int p = lane; p < n; p += 32) sc[p] = sc[p] / sum;
  __syncwarp();

  float4 acc = zero4;
  for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
    float4 vv[UNR];
    float w[UNR];
    const float* vb = kFastAddr ? iter_base(p0) + kv_delta : nullptr;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * PPI + grp;
      const bool ok = p < n && e_ok;
      if (kFastAddr) vv[u] = ok ? __ldcs(reinterpret_cast<const float4*>(vb + (u * PPI + grp) * a.row_stride)) : zero4;
      else vv[u] = ok ? __ldcs(reinterpret_cast<const float4*>(a.vcache + row_off(p) + e0)) : zero4;
      w[u] = ok ? sc[p] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      acc.x = fmaf(w[u], vv[u].x, acc.x); acc.y = fmaf(w[u], vv[u].y, acc.y);
      acc.z = fmaf(w[u], vv[u].z, acc.z); acc.w = fmaf(w[u], vv[u].w, acc.w);
    }
  }
  // all K / V of this CTA's first warp are in: let the next kernel ramp up (a grid-stride CTA triggers in its last round)
  if (threadIdx.x == 0 && gw + gridDim.x * wpb >= n_items) pdl_launch_dependents();
#pragma unroll
  for (int o = LPP; o < 32; o <<= 1) {       // combine the PPI position groups
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (grp == 0 && e_ok) {
    if (a.out_split == nullptr) {
      *reinterpret_cast<float4*>(a.out + (size_t)b * a.ldo + h * hd + e0) = acc;
    } else {
      // SplitTile output: 4 floats = half of an 8-element chunk -> 8 bytes of hi and 8 bytes of lo
      uint32_t hi0, lo0, hi1, lo1;
      split_pair(acc.x, acc.y, hi0, lo0);
      split_pair(acc.z, acc.w, hi1, lo1);
      const int col = h * hd + e0;
      const int mt = b >> 7, ri = b & 127, kb = col >> 6, cj = (col & 63) >> 3;
      uint8_t* dst = a.out_split + ((size_t)mt * a.kb_out + kb) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4) +
                     (size_t)((col & 7) >> 2) * 8;
      *reinterpret_cast<uint2*>(dst) = make_uint2(hi0, hi1);
      *reinterpret_cast<uint2*>(dst + 16384) = make_uint2(lo0, lo1);
    }
  }
  __syncwarp();          // the warp's score buffer is reused by its next item
  }
  trace_end(trc);
}

// ---------------------------------------------------------------------------------------------
// Attention of the teacher-forced pass (rows_per_seq = L > 0: query row r is position r % L of sequence r / L).
// All L queries of a (sequence, head) read the same keys, so one CTA per (sequence, head) stages that head's K / V rows in
// shared memory ONCE (the per-row kernel re-read them per query: 8.9 TB/s of L2 traffic, 56 % of a forward pass) and its
// warps take the positions round robin.  Per query the arithmetic and its order are attention_decode_v4_kernel's (same
// chunks of UNR * PPI positions, same reduction tree, softmax and accumulation order): identical bits.
// ---------------------------------------------------------------------------------------------
template <int LPP>
__global__ void __launch_bounds__(256) attention_forward_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float fw_smem[];
  pdl_wait();
  constexpr int PPI = 32 / LPP, UNR = 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int L = a.rows_per_seq, hd = a.hd, sb = blockIdx.x / a.nhead, h = blockIdx.x % a.nhead;
  const int n_keys = a.fixed_len >= 0 ? a.fixed_len : L;
  float* Ks = fw_smem;
  float* Vs = Ks + (size_t)n_keys * hd;
  float* sc = Vs + (size_t)n_keys * hd + (size_t)warp * a.max_n;
  TraceRec* trc = threadIdx.x == 0 ? trace_begin(a.trace, 5u) : nullptr;
  {
    const int c4n = hd >> 2;
    const float* kg = a.kcache + (size_t)sb * a.seq_stride + h * hd;
    const float* vg = a.vcache + (size_t)sb * a.seq_stride + h * hd;
    for (int i = threadIdx.x; i < n_keys * c4n; i += blockDim.x) {
      const int p = i / c4n, c = i - p * c4n;
      reinterpret_cast<float4*>(Ks)[i] = __ldcs(reinterpret_cast<const float4*>(kg + (size_t)p * a.row_stride) + c);
      reinterpret_cast<float4*>(Vs)[i] = __ldcs(reinterpret_cast<const float4*>(vg + (size_t)p * a.row_stride) + c);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int grp = lane / LPP, e0 = 4 * (lane % LPP);
  const bool e_ok = e0 < hd;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const unsigned char* ks = a.key_skip != nullptr ? a.key_skip + (size_t)sb * L : nullptr;
  for (int pos = warp; pos < L; pos += nwarps) {
    const int b = sb * L + pos;
    const int n = a.fixed_len >= 0 ? a.fixed_len : pos + 1;
    // (fetching the warp's queries eight positions ahead was measured: 80 registers, 8 % slower)
    const float4 q4 = e_ok ? *reinterpret_cast<const float4*>(a.q + (size_t)b * a.ldq + h * hd + e0) : zero4;
    for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
      float4 kv[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int p = p0 + u * PPI + grp;
        kv[u] = (p < n && e_ok) ? *reinterpret_cast<const float4*>(Ks + (size_t)p * hd + e0) : zero4;
      }
      float d[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) d[u] = fmaf(q4.x, kv[u].x, fmaf(q4.y, kv[u].y, fmaf(q4.z, kv[u].z, q4.w * kv[u].w)));
      if constexpr (LPP == 16) {
        const int sub = lane & 15;
        const bool up8 = (sub & 8) != 0, up4 = (sub & 4) != 0;
        const float r0 = __shfl_xor_sync(0xffffffffu, up8 ? d[0] : d[2], 8);
        const float r1 = __shfl_xor_sync(0xffffffffu, up8 ? d[1] : d[3], 8);
        const float e0s = (up8 ? d[2] : d[0]) + r0, e1s = (up8 ? d[3] : d[1]) + r1;
        float f = (up4 ? e1s : e0s) + __shfl_xor_sync(0xffffffffu, up4 ? e0s : e1s, 4);
        f += __shfl_xor_sync(0xffffffffu, f, 2);
        f += __shfl_xor_sync(0xffffffffu, f, 1);
        const int u_mine = 2 * (sub >> 3) + ((sub >> 2) & 1);
        const int p = p0 + u_mine * PPI + grp;
        if ((sub & 3) == 0 && p < n) sc[p] = f * a.scale;
      } else {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          float dd = d[u];
#pragma unroll
          for (int o = LPP / 2; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
          const int p = p0 + u * PPI + grp;
          if ((lane % LPP) == 0 && p < n) sc[p] = dd * a.scale;
        }
      }
    }
    __syncwarp();
    if (ks != nullptr) {
      for (int p = lane; p < n; p += 32)
        if (ks[p] != 0) sc[p] = -INFINITY;
      __syncwarp();
    }
    float m = -INFINITY;
    for (int p = lane; p < n; p += 32) m = fmaxf(m, sc[p]);
    m = warp_max(m);
    float sum = 0.f;
    for (int p = lane; p < n; p += 32) {
      const float e = expf(sc[p] - m);
      sc[p] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int p = lane; p < n; p += 32) sc[p] = sc[p] / sum;
    __syncwarp();
    float4 acc = zero4;
    for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
      float4 vv[UNR];
      float w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int p = p0 + u * PPI + grp;
        const bool ok = p < n && e_ok;
        vv[u] = ok ? *reinterpret_cast<const float4*>(Vs + (size_t)p * hd + e0) : zero4;
        w[u] = ok ? sc[p] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        acc.x = fmaf(w[u], vv[u].x, acc.x); acc.y = fmaf(w[u], vv[u].y, acc.y);
        acc.z = fmaf(w[u], vv[u].z, acc.z); acc.w = fmaf(w[u], vv[u].w, acc.w);
      }
    }
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (grp == 0 && e_ok) {
      if (a.out_split == nullptr) {
        *reinterpret_cast<float4*>(a.out + (size_t)b * a.ldo + h * hd + e0) = acc;
      } else {
        uint32_t hi0, lo0, hi1, lo1;
        split_pair(acc.x, acc.y, hi0, lo0);
        split_pair(acc.z, acc.w, hi1, lo1);
        const int col = h * hd + e0;
        const int mt = b >> 7, ri = b & 127, kb = col >> 6, cj = (col & 63) >> 3;
        uint8_t* dst = a.out_split + ((size_t)mt * a.kb_out + kb) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4) +
                       (size_t)((col & 7) >> 2) * 8;
        *reinterpret_cast<uint2*>(dst) = make_uint2(hi0, hi1);
        *reinterpret_cast<uint2*>(dst + 16384) = make_uint2(lo0, lo1);
      }
    }
    __syncwarp();
  }
  trace_end(trc);
}

// ---------------------------------------------------------------------------------------------
// Cross-attention over SHARED memory tokens (RLOO: the k samples of a latent attend to the same projected K / V).
// One warp per (latent, head): every 16-byte piece of the latent's K / V is loaded ONCE into registers and used for up to
// KS query rows (the samples, rows j, j + seq_mod, j + 2 seq_mod, ... of the launch), so the stream through L2 and the
// load instructions shrink by the group size, not only the HBM traffic.  Per query row the arithmetic and its order are
// those of attention_decode_v4_kernel (same chunks of UNR * PPI positions, same reduction tree, same accumulation
// order), so the output bits are identical.
// ---------------------------------------------------------------------------------------------
template <int LPP, int KS>
__global__ void __launch_bounds__(256, 4) attention_cross_shared_kernel(AttnArgs a, int k_local) {
  extern __shared__ float sc_all[];
  pdl_wait();
  if (a.st != nullptr && a.st->done) return;
  constexpr int PPI = 32 / LPP, UNR = 4;
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_items = a.seq_mod * a.nhead, n = a.fixed_len, hd = a.hd;
  TraceRec* trc = threadIdx.x == 0 ? trace_begin(a.trace, 4u) : nullptr;
  float* sc = sc_all + (size_t)warp_in_block * KS * a.max_n;
  const int grp = lane / LPP, e0 = 4 * (lane % LPP);
  const bool e_ok = e0 < hd;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long kv_delta = a.vcache - a.kcache;
  for (int gw = blockIdx.x * wpb + warp_in_block; gw < n_items; gw += gridDim.x * wpb) {
    const int j = gw / a.nhead, h = gw % a.nhead;
    const float* kbase = a.kcache + (size_t)j * a.seq_stride + h * hd + e0;
    for (int s0 = 0; s0 < k_local; s0 += KS) {
      const int ns = min(KS, k_local - s0);
      float4 q4[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s)
        q4[s] = (s < ns && e_ok) ? *reinterpret_cast<const float4*>(a.q + (size_t)((s0 + s) * a.seq_mod + j) * a.ldq + h * hd + e0) : zero4;
      for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
        float4 kv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int p = p0 + u * PPI + grp;
          kv[u] = (p < n && e_ok) ? __ldcs(reinterpret_cast<const float4*>(kbase + (size_t)p * a.row_stride)) : zero4;
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if (s < ns) {
            float d[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
              d[u] = fmaf(q4[s].x, kv[u].x, fmaf(q4[s].y, kv[u].y, fmaf(q4[s].z, kv[u].z, q4[s].w * kv[u].w)));
            if constexpr (LPP == 16) {
              const int sub = lane & 15;
              const bool up8 = (sub & 8) != 0, up4 = (sub & 4) != 0;
              const float r0 = __shfl_xor_sync(0xffffffffu, up8 ? d[0] : d[2], 8);
              const float r1 = __shfl_xor_sync(0xffffffffu, up8 ? d[1] : d[3], 8);
              const float e0s = (up8 ? d[2] : d[0]) + r0, e1s = (up8 ? d[3] : d[1]) + r1;
              float f = (up4 ? e1s : e0s) + __shfl_xor_sync(0xffffffffu, up4 ? e0s : e1s, 4);
              f += __shfl_xor_sync(0xffffffffu, f, 2);
              f += __shfl_xor_sync(0xffffffffu, f, 1);
              const int u_mine = 2 * (sub >> 3) + ((sub >> 2) & 1);
              const int p = p0 + u_mine * PPI + grp;
              if ((sub & 3) == 0 && p < n) sc[s * a.max_n + p] = f * a.scale;
            } else {
#pragma unroll
              for (int u = 0; u < UNR; ++u) {
                float dd = d[u];
#pragma unroll
                for (int o = LPP / 2; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
                const int p = p0 + u * PPI + grp;
                if ((lane % LPP) == 0 && p < n) sc[s * a.max_n + p] = dd * a.scale;
              }
            }
          }
        }
      }
      __syncwarp();
      for (int s = 0; s < ns; ++s) {
        float* scs = sc + s * a.max_n;
        float m = -INFINITY;
        for (int p = lane; p < n; p += 32) m = fmaxf(m, scs[p]);
        m = warp_max(m);
        float sum = 0.f;
        for (int p = lane; p < n; p += 32) {
          const float e = expf(scs[p] - m);
          scs[p] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        for (int p = lane; p < n; p += 32) scs[p] = scs[p] / sum;
      }
      __syncwarp();
      float4 acc[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) acc[s] = zero4;
      for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
        float4 vv[UNR];
        bool ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int p = p0 + u * PPI + grp;
          ok[u] = p < n && e_ok;
          vv[u] = ok[u] ? __ldcs(reinterpret_cast<const float4*>(kbase + kv_delta + (size_t)p * a.row_stride)) : zero4;
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if (s < ns) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              const float w = ok[u] ? sc[s * a.max_n + p0 + u * PPI + grp] : 0.f;
              acc[s].x = fmaf(w, vv[u].x, acc[s].x); acc[s].y = fmaf(w, vv[u].y, acc[s].y);
              acc[s].z = fmaf(w, vv[u].z, acc[s].z); acc[s].w = fmaf(w, vv[u].w, acc[s].w);
            }
          }
        }
      }
      if (threadIdx.x == 0 && gw + gridDim.x * wpb >= n_items && s0 + KS >= k_local) pdl_launch_dependents();
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        if (s < ns) {
          float4 r = acc[s];
#pragma unroll
          for (int o = LPP; o < 32; o <<= 1) {
            r.x += __shfl_xor_sync(0xffffffffu, r.x, o); r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
            r.z += __shfl_xor_sync(0xffffffffu, r.z, o); r.w += __shfl_xor_sync(0xffffffffu, r.w, o);
          }
          if (grp == 0 && e_ok) {
            const int b = (s0 + s) * a.seq_mod + j;
            if (a.out_split == nullptr) {
              *reinterpret_cast<float4*>(a.out + (size_t)b * a.ldo + h * hd + e0) = r;
            } else {
              uint32_t hi0, lo0, hi1, lo1;
              split_pair(r.x, r.y, hi0, lo0);
              split_pair(r.z, r.w, hi1, lo1);
              const int col = h * hd + e0;
              const int mt = b >> 7, ri = b & 127, kb = col >> 6, cj = (col & 63) >> 3;
              uint8_t* dst = a.out_split + ((size_t)mt * a.kb_out + kb) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4) +
                             (size_t)((col & 7) >> 2) * 8;
              *reinterpret_cast<uint2*>(dst) = make_uint2(hi0, hi1);
              *reinterpret_cast<uint2*>(dst + 16384) = make_uint2(lo0, lo1);
            }
          }
        }
      }
      __syncwarp();          // the score buffers are reused by the next group of samples / the next item
    }
  }
  trace_end(trc);
}

// ---------------------------------------------------------------------------------------------
// Cross-attention with bulk-copy staging (decode only: one query row per sequence, contiguous K / V).
// A sequence's projected memory tokens of one layer are ONE contiguous block ([M][k(d) | v(d)] fp32 = 96 KB for
// M = 24, d = 512), so instead of thousands of per-thread 16-byte loads a single elected thread streams whole blocks
// into a two-stage shared-memory ring with cp.async.bulk (UBLKCP) + mbarrier transaction counts, one persistent CTA per
// SM, while the warps (one per head) compute scores / softmax / P*V out of shared memory.  The load side is then pure
// DMA: no address arithmetic, no load instructions in the issue slots (the per-thread kernel above runs at 61 % issue
// utilisation and 57 % warps active while it streams; ncu, profiles/ncu_full_r01n_summary.csv), and up to two blocks
// (192 KB) per SM are in flight independent of how the warps are scheduled.  Same arithmetic, in the same order, as
// attention_decode_v4_kernel, so the two kernels produce identical bits.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx_(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "XWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra XWAIT_DONE;\n\t"
      "bra XWAIT_LOOP;\n\t"
      "XWAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s_(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int LPP>
__global__ void __launch_bounds__(256, 1) attention_cross_bulk_kernel(AttnArgs a, int kv_floats, int q_floats, int piece_bytes) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  pdl_wait();
  if (a.st != nullptr && a.st->done) return;
  constexpr int PPI = 32 / LPP, UNR = 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int n = a.fixed_len, hd = a.hd;
  float* kv_s[2] = {reinterpret_cast<float*>(bulk_smem), reinterpret_cast<float*>(bulk_smem) + kv_floats};
  float* q_s[2] = {kv_s[1] + kv_floats, kv_s[1] + kv_floats + q_floats};
  float* sc = q_s[1] + q_floats + (size_t)warp * a.max_n;
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_s[1] + q_floats + (size_t)nwarps * a.max_n);
  const uint32_t bar[2] = {smem_addr_u32(&bars[0]), smem_addr_u32(&bars[1])};
  TraceRec* trc = tid == 0 ? trace_begin(a.trace, 3u) : nullptr;
  if (tid == 0) {
    mbar_init_(bar[0], 1);
    mbar_init_(bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t kv_bytes = (uint32_t)kv_floats * 4u, q_bytes = (uint32_t)q_floats * 4u;
  auto issue = [&](int row, int st) {
    mbar_arrive_expect_tx_(bar[st], kv_bytes + q_bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.kcache + (size_t)row * a.seq_stride);
    const uint32_t dst = smem_addr_u32(kv_s[st]);
    for (uint32_t off = 0; off < kv_bytes; off += (uint32_t)piece_bytes)      // several copies in flight per block
      bulk_g2s_(dst + off, src + off, min((uint32_t)piece_bytes, kv_bytes - off), bar[st]);
    bulk_g2s_(smem_addr_u32(q_s[st]), a.q + (size_t)row * a.ldq, q_bytes, bar[st]);
  };
  const int G = gridDim.x;
  if (tid == 0) {
    if ((int)blockIdx.x < a.B) issue(blockIdx.x, 0);
    if ((int)blockIdx.x + G < a.B) issue(blockIdx.x + G, 1);
  }
  const int grp = lane / LPP, e0 = 4 * (lane % LPP);
  const bool e_ok = e0 < hd;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long v_off = a.vcache - a.kcache;          // d floats: V follows K inside a token's row
  int it = 0;
  for (int b = blockIdx.x; b < a.B; b += G, ++it) {
    const int st = it & 1;
    mbar_wait_(bar[st], (uint32_t)(it >> 1) & 1u);
    for (int h = warp; h < a.nhead; h += nwarps) {
      const float* kb = kv_s[st] + h * hd + e0;          // K of token 0, this lane's slice of head h
      const float4 q4 = e_ok ? *reinterpret_cast<const float4*>(q_s[st] + h * hd + e0) : zero4;
      for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
        float4 kv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int p = p0 + u * PPI + grp;
          kv[u] = (p < n && e_ok) ? *reinterpret_cast<const float4*>(kb + (size_t)p * a.row_stride) : zero4;
        }
        float d[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) d[u] = fmaf(q4.x, kv[u].x, fmaf(q4.y, kv[u].y, fmaf(q4.z, kv[u].z, q4.w * kv[u].w)));
        if constexpr (LPP == 16 && UNR == 4) {
          const int sub = lane & 15;
          const bool up8 = (sub & 8) != 0, up4 = (sub & 4) != 0;
          const float r0 = __shfl_xor_sync(0xffffffffu, up8 ? d[0] : d[2], 8);
          const float r1 = __shfl_xor_sync(0xffffffffu, up8 ? d[1] : d[3], 8);
          const float e0s = (up8 ? d[2] : d[0]) + r0, e1s = (up8 ? d[3] : d[1]) + r1;
          float f = (up4 ? e1s : e0s) + __shfl_xor_sync(0xffffffffu, up4 ? e0s : e1s, 4);
          f += __shfl_xor_sync(0xffffffffu, f, 2);
          f += __shfl_xor_sync(0xffffffffu, f, 1);
          const int u_mine = 2 * (sub >> 3) + ((sub >> 2) & 1);
          const int p = p0 + u_mine * PPI + grp;
          if ((sub & 3) == 0 && p < n) sc[p] = f * a.scale;
        } else {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            float dd = d[u];
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
            const int p = p0 + u * PPI + grp;
            if ((lane % LPP) == 0 && p < n) sc[p] = dd * a.scale;
          }
        }
      }
      __syncwarp();
      float m = -INFINITY;
      for (int p = lane; p < n; p += 32) m = fmaxf(m, sc[p]);
      m = warp_max(m);
      float sum = 0.f;
      for (int p = lane; p < n; p += 32) {
        const float e = expf(sc[p] - m);
        sc[p] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      __syncwarp();
      for (int p = lane; p < n; p += 32) sc[p] = sc[p] / sum;
      __syncwarp();
      float4 acc = zero4;
      const float* vb = kb + v_off;
      for (int p0 = 0; p0 < n; p0 += UNR * PPI) {
        float4 vv[UNR];
        float w[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int p = p0 + u * PPI + grp;
          const bool ok = p < n && e_ok;
          vv[u] = ok ? *reinterpret_cast<const float4*>(vb + (size_t)p * a.row_stride) : zero4;
          w[u] = ok ? sc[p] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          acc.x = fmaf(w[u], vv[u].x, acc.x); acc.y = fmaf(w[u], vv[u].y, acc.y);
          acc.z = fmaf(w[u], vv[u].z, acc.z); acc.w = fmaf(w[u], vv[u].w, acc.w);
        }
      }
#pragma unroll
      for (int o = LPP; o < 32; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
      if (grp == 0 && e_ok) {
        if (a.out_split == nullptr) {
          *reinterpret_cast<float4*>(a.out + (size_t)b * a.ldo + h * hd + e0) = acc;
        } else {
          uint32_t hi0, lo0, hi1, lo1;
          split_pair(acc.x, acc.y, hi0, lo0);
          split_pair(acc.z, acc.w, hi1, lo1);
          const int col = h * hd + e0;
          const int mt = b >> 7, ri = b & 127, kb2 = col >> 6, cj = (col & 63) >> 3;
          uint8_t* dst = a.out_split + ((size_t)mt * a.kb_out + kb2) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4) +
                         (size_t)((col & 7) >> 2) * 8;
          *reinterpret_cast<uint2*>(dst) = make_uint2(hi0, hi1);
          *reinterpret_cast<uint2*>(dst + 16384) = make_uint2(lo0, lo1);
        }
      }
      __syncwarp();
    }
    __syncthreads();                         // every warp is done with stage st: refill it
    if (tid == 0) {
      if (b + 2 * G < a.B) issue(b + 2 * G, st);
      else if (b + G >= a.B) pdl_launch_dependents();   // nothing left to stream for this CTA: let the next kernel ramp up
    }
  }
  trace_end(trc);
}

// bytes of dynamic shared memory of the bulk kernel, or 0 when the shape does not fit two stages
static size_t cross_bulk_smem(const AttnArgs& a, int warps) {
  if (a.fixed_len < 0 || a.page_table != nullptr || a.rows_per_seq != 0 || a.key_skip != nullptr || a.row_map != nullptr || a.seq_mod != 0) return 0;
  const size_t kv = (size_t)a.fixed_len * a.row_stride * sizeof(float), q = (size_t)a.nhead * a.hd * sizeof(float);
  if (a.seq_stride != (long long)a.fixed_len * a.row_stride) return 0;                  // one contiguous block per sequence
  if (a.vcache - a.kcache <= 0 || a.vcache - a.kcache >= a.row_stride) return 0;        // V inside the token's row
  if (kv % 16 != 0 || q % 16 != 0 || (a.ldq * sizeof(float)) % 16 != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(a.kcache) & 15u) != 0 || (reinterpret_cast<uintptr_t>(a.q) & 15u) != 0) return 0;
  const size_t total = 2 * (kv + q) + (size_t)warps * a.max_n * sizeof(float) + 16;
  return total <= 227 * 1024 ? total : 0;
}

int launch_attention(const AttnArgs& a_in, cudaStream_t s) {
  SCV_REQUIRE(a_in.hd >= 1 && a_in.hd <= 128, "attention: head_dim %d not in 1..128", a_in.hd);
  SCV_REQUIRE(a_in.max_n >= 1, "attention: max_n must be positive");
  SCV_REQUIRE(a_in.out_split == nullptr || a_in.hd % 8 == 0, "attention: SplitTile output needs head_dim %% 8 == 0");
  AttnArgs a = a_in;
  a.trace = trace_ptr();
  a.pages_regs = tun().attn_pages_regs;
  a.max_n = std::max(a_in.max_n, a_in.hd);     // the score buffer doubles as the staging row of the split output
  const int warps = 8;
  int blocks = ceil_div(a.B * a.nhead, warps);
  const int blocks_v1 = blocks;
  if (tun().attn_ctas_per_sm > 0) blocks = std::min(blocks, tun().attn_ctas_per_sm * sm_count());
  const size_t smem = (size_t)warps * a.max_n * sizeof(float);
  const int epl = ceil_div(a.hd, 32);
  const bool v4 = a.hd % 4 == 0 && a.ldq % 4 == 0 && a.row_stride % 4 == 0 && a.seq_stride % 4 == 0 && a.page_stride % 4 == 0 &&
                  (a.knew == nullptr || a.ldn % 4 == 0) && (a.out_split != nullptr || a.ldo % 4 == 0);
  const double n_hint = a.fixed_len >= 0 ? a.fixed_len : a.host_len_hint;
  // algorithmic traffic: K and V rows of every attended position (fp32) + q, out, and the appended row
  ProfScope prof(a.fixed_len >= 0 ? PC_ATTN_CROSS : PC_ATTN_SELF, s, 4.0 * a.B * a.nhead * a.hd * n_hint,
                 4.0 * a.B * a.nhead * a.hd * (2.0 * n_hint + 2.0 + (a.knew ? 4.0 : 0.0)));
  SCV_REQUIRE(a.rows_per_seq == 0 || v4, "attention: the teacher-forced layout needs head_dim %% 4 == 0 and 16-byte aligned rows");
  SCV_REQUIRE(a.seq_mod == 0 || v4, "attention: shared memory tokens need head_dim %% 4 == 0 and 16-byte aligned rows");
  SCV_REQUIRE(a.row_map == nullptr || v4, "attention: compaction of finished rows needs head_dim %% 4 == 0 and 16-byte aligned rows");
  // teacher-forced pass: one CTA per (sequence, head) with that head's K / V staged in shared memory
  // (from 128 sequences on: below that its (sequence, head) CTAs are too few and the per-row kernel's warps fill the GPU better)
  if (a.rows_per_seq > 0 && v4 && tun().attn_forward != 0 && a.page_table == nullptr && a.knew == nullptr && a.row_map == nullptr &&
      a.seq_mod == 0 && a.B % a.rows_per_seq == 0 && (a.B / a.rows_per_seq) * a.nhead >= tun().attn_forward_min_ctas) {
    const int n_keys = a.fixed_len >= 0 ? a.fixed_len : a.rows_per_seq;
    const size_t smem_f = ((size_t)2 * n_keys * a.hd + (size_t)warps * a.max_n) * sizeof(float);
    if (smem_f <= 200 * 1024) {
      static bool attr_f[64] = {};
      if (first_use_on_device(attr_f)) {
        SCV_CUDA(cudaFuncSetAttribute(attention_forward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SCV_CUDA(cudaFuncSetAttribute(attention_forward_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SCV_CUDA(cudaFuncSetAttribute(attention_forward_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SCV_CUDA(cudaFuncSetAttribute(attention_forward_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      }
      const dim3 grid_f((a.B / a.rows_per_seq) * a.nhead), block_f(warps * 32);
      const int lanes = a.hd / 4;
      if (lanes <= 4) SCV_CUDA(launch_k(attention_forward_kernel<4>, grid_f, block_f, smem_f, s, a));
      else if (lanes <= 8) SCV_CUDA(launch_k(attention_forward_kernel<8>, grid_f, block_f, smem_f, s, a));
      else if (lanes <= 16) SCV_CUDA(launch_k(attention_forward_kernel<16>, grid_f, block_f, smem_f, s, a));
      else SCV_CUDA(launch_k(attention_forward_kernel<32>, grid_f, block_f, smem_f, s, a));
      SCV_LAUNCH_CHECK();
      return 0;
    }
  }
  // shared memory tokens, whole groups of seq_mod rows in this launch: one warp serves the k samples of a (latent, head)
  const int k_shared = (a.seq_mod > 0 && a.row_map == nullptr && a.fixed_len >= 0 && a.knew == nullptr && a.page_table == nullptr &&
                        a.rows_per_seq == 0 && a.key_skip == nullptr && a.B % a.seq_mod == 0 && a.slot_base % a.seq_mod == 0)
                           ? a.B / a.seq_mod : 0;
  if (k_shared > 1 && v4 && tun().attn_shared != 0) {
    constexpr int KS = 4;
    const size_t smem_s = (size_t)warps * KS * a.max_n * sizeof(float);
    int blocks_s = ceil_div(a.seq_mod * a.nhead, warps);
    if (tun().attn_ctas_per_sm > 0) blocks_s = std::min(blocks_s, tun().attn_ctas_per_sm * sm_count());
    const int lanes = a.hd / 4;
    if (lanes <= 4) SCV_CUDA(launch_k(attention_cross_shared_kernel<4, KS>, dim3(blocks_s), dim3(warps * 32), smem_s, s, a, k_shared));
    else if (lanes <= 8) SCV_CUDA(launch_k(attention_cross_shared_kernel<8, KS>, dim3(blocks_s), dim3(warps * 32), smem_s, s, a, k_shared));
    else if (lanes <= 16) SCV_CUDA(launch_k(attention_cross_shared_kernel<16, KS>, dim3(blocks_s), dim3(warps * 32), smem_s, s, a, k_shared));
    else SCV_CUDA(launch_k(attention_cross_shared_kernel<32, KS>, dim3(blocks_s), dim3(warps * 32), smem_s, s, a, k_shared));
    SCV_LAUNCH_CHECK();
    return 0;
  }
  const size_t bulk_smem = (v4 && tun().attn_bulk != 0 && a.B >= tun().attn_bulk_min_rows) ? cross_bulk_smem(a, warps) : 0;
  if (bulk_smem != 0) {
    static bool attr_dev[64] = {};
    if (first_use_on_device(attr_dev)) {
      SCV_CUDA(cudaFuncSetAttribute(attention_cross_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      SCV_CUDA(cudaFuncSetAttribute(attention_cross_bulk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      SCV_CUDA(cudaFuncSetAttribute(attention_cross_bulk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      SCV_CUDA(cudaFuncSetAttribute(attention_cross_bulk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int kv_floats = a.fixed_len * a.row_stride, q_floats = a.nhead * a.hd;
    const int piece = tun().attn_bulk_piece_kb > 0 ? tun().attn_bulk_piece_kb * 1024 : kv_floats * 4;
    const dim3 grid(std::min(a.B, sm_count())), block(warps * 32);
    const int lanes = a.hd / 4;
    if (lanes <= 4) SCV_CUDA(launch_k(attention_cross_bulk_kernel<4>, grid, block, bulk_smem, s, a, kv_floats, q_floats, piece));
    else if (lanes <= 8) SCV_CUDA(launch_k(attention_cross_bulk_kernel<8>, grid, block, bulk_smem, s, a, kv_floats, q_floats, piece));
    else if (lanes <= 16) SCV_CUDA(launch_k(attention_cross_bulk_kernel<16>, grid, block, bulk_smem, s, a, kv_floats, q_floats, piece));
    else SCV_CUDA(launch_k(attention_cross_bulk_kernel<32>, grid, block, bulk_smem, s, a, kv_floats, q_floats, piece));
    SCV_LAUNCH_CHECK();
    return 0;
  }
  if (v4) {
    const int lanes = a.hd / 4;
    if (lanes <= 4) SCV_CUDA(launch_k(attention_decode_v4_kernel<4>, dim3(blocks), dim3(warps * 32), smem, s, a));
    else if (lanes <= 8) SCV_CUDA(launch_k(attention_decode_v4_kernel<8>, dim3(blocks), dim3(warps * 32), smem, s, a));
    else if (lanes <= 16) SCV_CUDA(launch_k(attention_decode_v4_kernel<16>, dim3(blocks), dim3(warps * 32), smem, s, a));
    else SCV_CUDA(launch_k(attention_decode_v4_kernel<32>, dim3(blocks), dim3(warps * 32), smem, s, a));
    SCV_LAUNCH_CHECK();
    return 0;
  }
  switch (epl) {
    case 1: SCV_CUDA(launch_k(attention_decode_kernel<1>, dim3(blocks_v1), dim3(warps * 32), smem, s, a)); break;
    case 2: SCV_CUDA(launch_k(attention_decode_kernel<2>, dim3(blocks_v1), dim3(warps * 32), smem, s, a)); break;
    case 3: SCV_CUDA(launch_k(attention_decode_kernel<3>, dim3(blocks_v1), dim3(warps * 32), smem, s, a)); break;
    default: SCV_CUDA(launch_k(attention_decode_kernel<4>, dim3(blocks_v1), dim3(warps * 32), smem, s, a)); break;
  }
  SCV_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Sampling epilogue (:1415-1548).  One CTA per row, the row's logits staged once in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kSamplerThreads = 256;

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kSamplerThreads / 32; ++i) t += red[i];
  return t;
}
__device__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSamplerThreads / 32; ++i) t = fmaxf(t, red[i]);
  return t;
}
__device__ int block_argmax(const float* sl, int V, float* redv, int* redi) {
  float bv = -INFINITY;
  int bi = INT_MAX;
  for (int v = threadIdx.x; v < V; v += kSamplerThreads)
    if (arg_better(sl[v], v, bv, bi)) { bv = sl[v]; bi = v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { redv[w] = bv; redi[w] = bi; }
  __syncthreads();
  bv = redv[0]; bi = redi[0];
#pragma unroll
  for (int i = 1; i < kSamplerThreads / 32; ++i)
    if (arg_better(redv[i], redi[i], bv, bi)) { bv = redv[i]; bi = redi[i]; }
  return bi;
}

__device__ __forceinline__ int logit_flags(float l) {         // bit0: nan/+-inf, bit1: nan/+inf
  int bad = 0;
  if (isnan(l) || isinf(l)) bad |= 1;
  if (isnan(l) || (isinf(l) && l > 0.f)) bad |= 2;
  return bad;
}

// Stage the row's adjusted logits in shared memory.  Returns bit0: row has nan/+-inf, bit1: row has nan/+inf (a
// genuine numerical failure rather than a mask).
__device__ int stage_logits(const SamplerArgs& a, int b, int step, float* sl) {
  const RowCtx c = make_row_ctx(a, b, step);
  const float* lg = a.logits + (size_t)b * a.ldl;
  int bad = 0;
  for (int v = threadIdx.x; v < a.V; v += kSamplerThreads) {
    const float l = adjust_logit(c, v, lg[v], c.mk != nullptr ? c.mk[v] : 1u, c.seen_row != nullptr ? c.seen_row[v] : 0u);
    sl[v] = l;
    bad |= logit_flags(l);
  }
  return bad;
}

// Phase 1: stage logits, publish the degenerate flag; when the call is plain greedy, also pick the token.
__global__ void __launch_bounds__(kSamplerThreads) sampler_phase1_kernel(SamplerArgs a, int finalize) {
  extern __shared__ float sl[];
  __shared__ float redv[kSamplerThreads / 32];
  __shared__ int redi[kSamplerThreads / 32];
  pdl_wait();
  if (a.st->done) return;
  pdl_launch_dependents();
  const int b = blockIdx.x, step = a.st->step;
  if (a.row_map != nullptr && a.slot_base + b >= a.st->pad[1]) return;
  const int bad = stage_logits(a, b, step, sl);
  const int row_bad = __syncthreads_or(bad & 1);
  // test before the atomic: with a type mask every row is "bad" and they would all serialise on this one word
  if (threadIdx.x == 0 && row_bad && *reinterpret_cast<volatile int*>(&a.st->degenerate) == 0) atomicOr(&a.st->degenerate, 1);
  if (!finalize) return;
  if (a.temperature != 1.0f) {
    for (int v = threadIdx.x; v < a.V; v += kSamplerThreads) sl[v] = sl[v] / a.temperature;   // (:1485-1486)
    __syncthreads();
  }
  const int tok = block_argmax(sl, a.V, redv, redi);                                          // (:1507)
  if (threadIdx.x == 0) commit_token(a, b, step, tok, 0.f);
}

// Plain greedy decoding (temperature < 0.01, no entropy): one warp per row, no staging -- the adjusted logits are
// consumed as they stream in (16-byte loads, 4 mask bytes at a time) and only the running argmax is kept.  Same
// per-element arithmetic as phase 1 above (adjust, / temperature, first-occurrence argmax with NaN as maximum).
// Four warps per row (two rows per CTA): a row's 19 KB of logits is ten dependent rounds of loads for one warp; dealt to
// four warps the rounds run side by side (the logits sit in L2, the kernel is load-latency bound) and the warps' candidates
// are merged with the same first-occurrence rule.
constexpr int kGreedyWarpsPerRow = 4;
__global__ void __launch_bounds__(256) sampler_greedy_kernel(SamplerArgs a) {
  __shared__ float cand_v[8];
  __shared__ int cand_i[8];
  pdl_wait();
  if (a.st->done) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * (8 / kGreedyWarpsPerRow) + warp / kGreedyWarpsPerRow, part = warp % kGreedyWarpsPerRow;
  const bool live = b < a.B && !(a.row_map != nullptr && a.slot_base + b >= a.st->pad[1]);
  const int step = a.st->step;
  float best = -INFINITY;
  int tok = INT_MAX;
  if (live) tok = greedy_row_token(a, b, step, lane, part, kGreedyWarpsPerRow, &best);
  if (lane == 0) { cand_v[warp] = best; cand_i[warp] = tok; }
  __syncthreads();
  if (threadIdx.x == 0) pdl_launch_dependents();
  // (the batch-global degenerate flag, :1464-1466, only changes how probabilities are sampled; argmax ignores it,
  // and with a type mask every row would hit the same atomic: 4096 serialised updates cost ~75 us per step)
  if (live && part == 0 && lane == 0) {
    float bv = cand_v[warp];
    int bi = cand_i[warp];
    for (int w = 1; w < kGreedyWarpsPerRow; ++w)
      if (arg_better(cand_v[warp + w], cand_i[warp + w], bv, bi)) { bv = cand_v[warp + w]; bi = cand_i[warp + w]; }
    commit_token(a, b, step, bi, 0.f);                          // (:1507)
  }
}

// Phase 2: entropy, temperature, multinomial (or argmax) and log-prob, given the batch-global flag.
__global__ void __launch_bounds__(kSamplerThreads) sampler_phase2_kernel(SamplerArgs a) {
  extern __shared__ float sl[];
  __shared__ float redv[kSamplerThreads / 32];
  __shared__ int redi[kSamplerThreads / 32];
  __shared__ float scan[kSamplerThreads];
  __shared__ int pick;
  pdl_wait();
  if (a.st->done) return;
  pdl_launch_dependents();
  const int b = blockIdx.x, step = a.st->step, V = a.V, tid = threadIdx.x;
  if (a.row_map != nullptr && a.slot_base + b >= a.st->pad[1]) return;
  const int ob = orig_row(a, b);
  const int bad = stage_logits(a, b, step, sl);
  const int row_real_bad = __syncthreads_or(bad & 2);
  const bool degenerate = (a.flags & 1u) ? (a.st->degenerate != 0) : (row_real_bad != 0);

  if (a.out_entropy != nullptr) {                                                   // (:1470-1482)
    float H;
    if (degenerate) {
      H = logf((float)max(V, 1));
    } else {
      float m = -INFINITY;
      for (int v = tid; v < V; v += kSamplerThreads) m = fmaxf(m, sl[v]);
      m = block_max(m, redv);
      float s = 0.f;
      for (int v = tid; v < V; v += kSamplerThreads) s += expf(sl[v] - m);
      s = block_sum(s, redv);
      float h = 0.f;
      for (int v = tid; v < V; v += kSamplerThreads) {
        const float p = fmaxf(expf(sl[v] - m) / s, 1e-8f);
        h += p * logf(p);
      }
      H = -block_sum(h, redv);
    }
    if (tid == 0) a.out_entropy[(size_t)ob * a.out_ld + step] = H;
  }
  if (a.temperature != 1.0f) {
    __syncthreads();
    for (int v = tid; v < V; v += kSamplerThreads) sl[v] = sl[v] / a.temperature;
  }
  __syncthreads();
  if (a.temperature < 0.01f) {
    const int tok = block_argmax(sl, V, redv, redi);
    if (tid == 0) commit_token(a, b, step, tok, 0.f);
    return;
  }
  if (a.sort_n > 0) {
    // top-k (:1489-1491) and nucleus (:1494-1503) filtering need the descending order: bitonic sort of
    // (logit, id) pairs in shared memory (ties broken by id; padding sorts last)
    const int NP = a.sort_n;
    float* key = sl + V;
    int* sid = reinterpret_cast<int*>(key + NP);
    for (int i = tid; i < NP; i += kSamplerThreads) { key[i] = i < V ? sl[i] : -INFINITY; sid[i] = i; }
    __syncthreads();
    auto before = [](float x, int ix, float y, int iy) { return x > y || (x == y && ix < iy); };
    for (int k = 2; k <= NP; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < NP; i += kSamplerThreads) {
          const int o = i ^ j;
          if (o > i) {
            const float x = key[i], y = key[o];
            const int ix = sid[i], iy = sid[o];
            const bool desc = (i & k) == 0;
            if (desc ? before(y, iy, x, ix) : before(x, ix, y, iy)) { key[i] = y; key[o] = x; sid[i] = iy; sid[o] = ix; }
          }
        }
        __syncthreads();
      }
    }
    if (a.top_k > 0) {
      const float thr = key[min(a.top_k, V) - 1];                 // k-th largest, duplicates counted
      __syncthreads();
      for (int i = tid; i < NP; i += kSamplerThreads) {
        if (i < V && sl[i] < thr) sl[i] = -INFINITY;
        if (key[i] < thr) key[i] = -INFINITY;
      }
      __syncthreads();
    }
    if (a.top_p < 1.0f) {
      const float mx = key[0];
      const int per = NP / kSamplerThreads;                       // contiguous sorted positions per thread
      float loc = 0.f;
      for (int i = tid * per; i < (tid + 1) * per; ++i) { const float e = expf(key[i] - mx); key[i] = e; loc += e; }
      scan[tid] = loc;
      __syncthreads();
      if (tid < 32) {
        float part[kSamplerThreads / 32], run = 0.f;
#pragma unroll
        for (int i = 0; i < kSamplerThreads / 32; ++i) { part[i] = run; run += scan[tid * (kSamplerThreads / 32) + i]; }
        float incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
#pragma unroll
        for (int i = 0; i < kSamplerThreads / 32; ++i) scan[tid * (kSamplerThreads / 32) + i] = incl - run + part[i];
        if (tid == 31) redv[0] = incl;
      }
      __syncthreads();
      const float total = redv[0];
      float cum = scan[tid];                                      // exclusive prefix of this thread's first position
      for (int i = tid * per; i < (tid + 1) * per; ++i) {
        // position i is dropped when the cumulative probability of positions < i already exceeds top_p
        if (i >= 1 && cum / total > a.top_p && sid[i] < V) sl[sid[i]] = -INFINITY;
        cum += key[i];
      }
      __syncthreads();
    }
  }
  // probs = softmax(logits) (:1511); uniform when degenerate (:1512-1513)
  const float u = Philox::uniform(a.st->seed, a.st->offset, (uint32_t)(a.row_map != nullptr ? ob : b + a.row_base), (uint32_t)step);
  if (degenerate) {
    if (tid == 0) {
      int tok = min((int)(u * (float)V), V - 1);
      commit_token(a, b, step, tok, logf(fmaxf(1.0f / (float)V, 1e-8f)));
    }
    return;
  }
  float m = -INFINITY;
  for (int v = tid; v < V; v += kSamplerThreads) m = fmaxf(m, sl[v]);
  m = block_max(m, redv);
  // contiguous chunk per thread so that the inverse-CDF walk is in token-id order
  const int chunk = (V + kSamplerThreads - 1) / kSamplerThreads;
  const int v0 = min(tid * chunk, V), v1 = min(v0 + chunk, V);
  float csum = 0.f;
  int last_pos = -1;
  for (int v = v0; v < v1; ++v) {
    const float e = expf(sl[v] - m);
    sl[v] = e;
    csum += e;
    if (e > 0.f) last_pos = v;
  }
  scan[tid] = csum;
  if (tid == 0) pick = -1;
  __syncthreads();
  // exclusive scan over 256 chunk sums by one warp (8 values per lane)
  if (tid < 32) {
    float loc[kSamplerThreads / 32];
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < kSamplerThreads / 32; ++i) { loc[i] = run; run += scan[tid * (kSamplerThreads / 32) + i]; }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += t;
    }
    const float excl = incl - run;
#pragma unroll
    for (int i = 0; i < kSamplerThreads / 32; ++i) scan[tid * (kSamplerThreads / 32) + i] = excl + loc[i];
    if (tid == 31) redv[0] = incl;   // total
  }
  __syncthreads();
  const float total = redv[0];
  const float target = u * total;
  const float excl = scan[tid];
  const float next_excl = (tid + 1 < kSamplerThreads) ? scan[tid + 1] : INFINITY;
  if (target >= excl && target < next_excl && v0 < v1) {
    float run = excl;
    int found = -1, cand = -1;
    for (int v = v0; v < v1; ++v) {
      run += sl[v];
      if (sl[v] > 0.f) {
        cand = v;
        if (run > target) { found = v; break; }
      }
    }
    if (found < 0) found = cand;   // summation-order rounding at the chunk edge
    if (found >= 0) pick = found;
  }
  // fallback for rounding at chunk edges: the last token with non-zero probability at or before ... any
  int lp = last_pos;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lp = max(lp, __shfl_xor_sync(0xffffffffu, lp, o));
  __syncthreads();
  if ((tid & 31) == 0) redi[tid >> 5] = lp;
  __syncthreads();
  if (tid == 0) {
    int tok = pick;
    if (tok < 0) {
      tok = 0;
      for (int i = 0; i < kSamplerThreads / 32; ++i) tok = max(tok, redi[i]);
    }
    if (a.forced != nullptr) {           // teacher-forced replay; a negative entry leaves that position sampled
      const long long f = a.forced[(size_t)ob * a.out_ld + step];
      if (f >= 0) tok = (int)f;
    }
    const float p = sl[tok] / total;
    commit_token(a, b, step, tok, logf(fmaxf(p, 1e-8f)));                            // (:1518)
  }
}

bool sampler_plain_greedy(const SamplerArgs& a) {
  const bool two_phase = !(a.temperature < 0.01f) || a.want_entropy;
  const bool vec_ok = a.V % 4 == 0 && a.ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(a.logits) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(a.type_masks) & 3u) == 0 && (reinterpret_cast<uintptr_t>(a.seen) & 3u) == 0;
  return !two_phase && vec_ok;
}

int launch_sampler(const SamplerArgs& a_in, int which, cudaStream_t s) {
  SamplerArgs a = a_in;
  const bool two_phase = !(a.temperature < 0.01f) || a.want_entropy;
  const bool filter = !(a.temperature < 0.01f) && (a.top_k > 0 || a.top_p < 1.0f);   // argmax ignores the filters
  a.sort_n = 0;
  if (filter) { a.sort_n = kSamplerThreads; while (a.sort_n < a.V) a.sort_n <<= 1; }
  const size_t smem = (size_t)a.V * sizeof(float) + (size_t)a.sort_n * 8;
  ProfScope prof(PC_SAMPLER, s, 4.0 * a.B * a.V, 4.0 * a.B * a.V * (two_phase ? 2 : 1));
  static bool attr_dev[64] = {};
  if (smem > 40 * 1024 && first_use_on_device(attr_dev)) {
    SCV_CUDA(cudaFuncSetAttribute(sampler_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SCV_CUDA(cudaFuncSetAttribute(sampler_phase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  SCV_REQUIRE(smem <= 200 * 1024, "sampler: vocabulary of %d tokens does not fit in shared memory", a.V);
  const bool vec_ok = a.V % 4 == 0 && a.ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(a.logits) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(a.type_masks) & 3u) == 0 && (reinterpret_cast<uintptr_t>(a.seen) & 3u) == 0;
  if (which == 1 && !two_phase && vec_ok) {
    SCV_CUDA(launch_k(sampler_greedy_kernel, dim3(ceil_div(a.B, 8 / kGreedyWarpsPerRow)), dim3(256), 0, s, a));
    SCV_LAUNCH_CHECK();
  } else if (which == 1) {
    SCV_CUDA(launch_k(sampler_phase1_kernel, dim3(a.B), dim3(kSamplerThreads), smem, s, a, two_phase ? 0 : 1));
    SCV_LAUNCH_CHECK();
  } else if (two_phase) {
    SCV_CUDA(launch_k(sampler_phase2_kernel, dim3(a.B), dim3(kSamplerThreads), smem, s, a));
    SCV_LAUNCH_CHECK();
  }
  return 0;
}

__global__ void step_end_kernel(StepState* st, int max_steps) {
  pdl_wait();
  if (st->done) return;
  const int s = st->step + 1;
  st->step = s;
  st->degenerate = 0;
  if (st->n_unfinished <= 0 || s >= max_steps) {   // finished.all() -> break (:1547-1548)
    st->done = 1;
    st->out_len = s;
  }
}

int launch_step_end(StepState* st, int max_steps, cudaStream_t s) {
  SCV_CUDA(launch_k(step_end_kernel, dim3(1), dim3(1), 0, s, st, max_steps));
  SCV_LAUNCH_CHECK();
  return 0;
}

// Greedy epilogue over every position of a teacher-forced pass (scv_greedy_positions): one warp per (sequence, position)
// row; step = position.  Same per-element arithmetic as the decode's greedy sampler (adjust_logit, / temperature,
// first-occurrence argmax with NaN as maximum); nothing is committed, the chosen ids go to out_tokens.
__global__ void __launch_bounds__(256) greedy_positions_kernel(SamplerArgs a, int seq_len, long long n_rows, int vec_ok,
                                                               long long* __restrict__ out_tokens) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int step = (int)(r % seq_len);
  int tok;
  if (vec_ok) {
    tok = greedy_row_token(a, (int)r, step, lane);
  } else {
    const RowCtx c = make_row_ctx(a, (int)r, step);
    const float* lg = a.logits + (size_t)r * a.ldl;
    float bv = -INFINITY;
    int bi = INT_MAX;
    for (int v = lane; v < a.V; v += 32) {
      float l = adjust_logit(c, v, lg[v], c.mk != nullptr ? c.mk[v] : 1u, 0u);
      if (a.temperature != 1.0f) l = l / a.temperature;
      if (arg_better(l, v, bv, bi)) { bv = l; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    tok = bi;
  }
  if (lane == 0) out_tokens[r] = tok;
}

}  // namespace scv

extern "C" int scv_greedy_positions(const float* logits, const float* type_logits, const float* stop_logits, const uint8_t* type_masks,
                                    const uint8_t* finished_before, int64_t n_rows, int32_t seq_len, int32_t vocab, int32_t max_len,
                                    float temperature, float stop_boost, float hard_stop_threshold, int64_t* out_tokens, void* stream) {
  using namespace scv;
  SCV_REQUIRE(logits && finished_before && out_tokens && n_rows > 0 && seq_len > 0 && vocab > kEndIdx, "greedy_positions: bad arguments");
  SCV_REQUIRE(n_rows < (1ll << 31), "greedy_positions: %lld rows (at most 2^31 - 1)", (long long)n_rows);
  SamplerArgs a;
  a.logits = logits; a.ldl = vocab; a.type_logits = type_logits; a.ldt = 5; a.stop_logits = stop_logits;
  a.type_masks = type_logits != nullptr ? type_masks : nullptr;
  a.B = (int)n_rows; a.V = vocab; a.max_len = max_len; a.temperature = temperature;
  a.stop_boost = stop_logits != nullptr ? stop_boost : 0.f; a.hard_stop = hard_stop_threshold;
  a.finished = const_cast<unsigned char*>(finished_before);
  const int vec_ok = vocab % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.type_masks) & 3u) == 0;
  const long long blocks = (n_rows + 7) / 8;
  greedy_positions_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, seq_len, n_rows, vec_ok,
                                                                                            reinterpret_cast<long long*>(out_tokens));
  SCV_LAUNCH_CHECK();
  return 0;
}

namespace scv {

__global__ void host_gate_kernel(volatile int* flag) {
  // bounded wait (50 ms): if the host ever blocked in a launch call while the gate is closed, the gate opens by itself
  const unsigned long long t0 = globaltimer_ns();
  while (*flag == 0 && globaltimer_ns() - t0 < 50000000ull) __nanosleep(200);
}

int launch_host_gate(int* host_flag, cudaStream_t s) {
  host_gate_kernel<<<1, 1, 0, s>>>(host_flag);
  SCV_LAUNCH_CHECK();
  return 0;
}

__global__ void init_rows_kernel(int* cur_tokens, unsigned char* finished, int B, StepState* st,
                                 unsigned long long seed, unsigned long long offset, int* row_map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    cur_tokens[i] = kStartIdx;     // (:1392)
    finished[i] = 0;
    if (row_map != nullptr) row_map[i] = i;
  }
  if (i == 0) {
    st->step = 0; st->done = 0; st->n_unfinished = B; st->out_len = 0; st->degenerate = 0; st->next_free_page = 0;
    st->pad[0] = 0;                        // 1 + the last step at which a row emitted its END (decode_cluster.cu)
    st->pad[1] = row_map != nullptr ? B : 0;   // rows still being decoded (compaction of finished rows), 0 = off
    st->seed = seed; st->offset = offset;
  }
}

// Opt-in retirement of finished rows: after the sampler, the slots whose row has not emitted END move to the front (stable),
// so the next step's projections / attention / sampler run on st->pad[1] rows only.  One CTA, <= 16 slots per thread.
__global__ void __launch_bounds__(1024) compact_rows_kernel(int* row_map, int* cur_tokens, const unsigned char* finished,
                                                            StepState* st, int B) {
  __shared__ int warp_tot[32];
  __shared__ int total;
  pdl_wait();
  if (st->done) return;
  const int n = st->pad[1], tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int PER = 16;                       // B <= 16384
  int rows[PER], toks[PER];
  int keep = 0;
  const int s0 = tid * PER;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int sl = s0 + j;
    rows[j] = -1;
    if (sl < n) {
      const int r = row_map[sl];
      if (finished[r] == 0) { rows[keep] = r; toks[keep] = cur_tokens[sl]; ++keep; }
    }
  }
  int incl = keep;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();                              // every slot has been read before any is overwritten
  if (w == 0) {
    int v = warp_tot[lane], in2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, in2, o); if (lane >= o) in2 += t; }
    warp_tot[lane] = in2 - v;
    if (lane == 31) total = in2;
  }
  __syncthreads();
  const int base = warp_tot[w] + incl - keep;
  for (int j = 0; j < keep; ++j) { row_map[base + j] = rows[j]; cur_tokens[base + j] = toks[j]; }
  if (tid == 0) st->pad[1] = total;
  (void)B;
}

int launch_compact_rows(int* row_map, int* cur_tokens, const unsigned char* finished, StepState* st, int B, cudaStream_t s) {
  SCV_REQUIRE(B <= 16384, "compaction of finished rows: %d rows (at most 16384 per call)", B);
  SCV_CUDA(launch_k(compact_rows_kernel, dim3(1), dim3(1024), 0, s, row_map, cur_tokens, finished, st, B));
  SCV_LAUNCH_CHECK();
  return 0;
}

int launch_init_rows(int* cur_tokens, unsigned char* finished, int B, StepState* st, unsigned long long seed,
                     unsigned long long offset, cudaStream_t s, int* row_map) {
  init_rows_kernel<<<ceil_div(B, 256), 256, 0, s>>>(cur_tokens, finished, B, st, seed, offset, row_map);
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
