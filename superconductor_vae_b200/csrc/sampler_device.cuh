// Device-side pieces of the sampling epilogue shared by the per-step sampler kernels (decode_kernels.cu) and the
// persistent small-batch decode (decode_small.cu).  Everything a step writes and a later step reads (logits, head
// outputs) is loaded through L2 (ld.cg), so the functions are also correct inside a kernel that spans several steps.
#pragma once
#include <limits.h>

#include "decode_kernels.cuh"

namespace scv {

__device__ __forceinline__ bool arg_better(float a, int ia, float b, int ib) {
  const bool an = isnan(a), bn = isnan(b);     // torch.argmax treats NaN as the maximum
  if (an != bn) return an;
  if (!an && a != b) return a > b;
  return ia < ib;                              // first occurrence wins ties
}


// Row context of the logit adjustments: type mask (:1416-1422), site-dup gate (:1426-1435), stop boost (:1438-1441),
// hard stop (:1444-1448) and length boost (:1455-1457), applied per element in the reference's order.
struct RowCtx {
  const uint8_t* mk; const unsigned char* seen_row;
  bool stop_on, force, dup_suppress, late;
  float boost, length_boost;
};
// b = row of the per-step buffers (logits, head outputs: the SLOT when finished rows are compacted away), ob = the row of
// the call's batch it stands for (outputs, finished flags, site-dup bitmap); ob == b without compaction.
__device__ __forceinline__ int orig_row(const SamplerArgs& a, int b) { return a.row_map != nullptr ? a.row_map[a.slot_base + b] : b; }
__device__ __forceinline__ RowCtx make_row_ctx(const SamplerArgs& a, int b, int step) {
  const int ob = orig_row(a, b);
  RowCtx c;
  int pred_type = 0;
  if (a.type_masks != nullptr) {
    const float* tl = a.type_logits + (size_t)b * a.ldt;
    float bv = __ldcg(tl);
    for (int t = 1; t < 5; ++t)
      { const float v = __ldcg(tl + t); if (arg_better(v, t, bv, pred_type)) { bv = v; pred_type = t; } }
  }
  c.stop_on = a.stop_boost > 0.f;
  float sp = 0.f;
  c.force = false;
  if (c.stop_on) {
    sp = sigmoidf_(__ldcg(a.stop_logits + b));
    c.force = a.hard_stop > 0.f && sp > a.hard_stop && a.finished[ob] == 0;
  }
  c.boost = a.stop_boost * sp;
  c.late = c.stop_on && step > 10;
  c.length_boost = c.late ? 10.0f * (float)(step - 10) / (float)max(a.max_len - 10, 1) : 0.f;
  c.dup_suppress = a.seen != nullptr && step > 0 && sigmoidf_(__ldcg(a.dup_logits + b)) < a.dup_threshold;
  c.seen_row = c.dup_suppress ? a.seen + (size_t)ob * a.V : nullptr;
  c.mk = a.type_masks != nullptr ? a.type_masks + (size_t)pred_type * a.V : nullptr;
  return c;
}
// allowed = type-mask byte of v (1 when no mask is given), seen = site-dup byte of v (0 when the gate is off)
__device__ __forceinline__ float adjust_logit(const RowCtx& c, int v, float l, unsigned allowed, unsigned seen) {
  if (allowed == 0) l = -INFINITY;
  if (seen != 0) l = -30.0f;                                  // masked_fill(-30.0), even over a -inf
  if (v == kEndIdx && c.stop_on) l = l + c.boost;
  if (c.force) l = (v == kEndIdx) ? 100.0f : -INFINITY;
  if (v == kEndIdx && c.late) l = l + c.length_boost;
  return l;
}

__device__ __forceinline__ int commit_token(const SamplerArgs& a, int b, int step, int token, float logprob) {
  // single thread
  const int ob = orig_row(a, b);
  if (a.forced != nullptr) {               // teacher-forced replay; a negative entry leaves the position to the sampler
    const long long f = a.forced[(size_t)ob * a.out_ld + step];
    if (f >= 0) token = (int)f;
  }
  a.out_tokens[(size_t)ob * a.out_ld + step] = (long long)token;
  if (a.out_logprobs != nullptr) a.out_logprobs[(size_t)ob * a.out_ld + step] = logprob;
  a.cur_tokens[b] = token;
  // ids 20..137 are the element range of the pre-V13 vocabulary; the reference still uses it
  if (a.seen != nullptr && token >= 20 && token <= 137 && a.finished[ob] == 0) a.seen[(size_t)ob * a.V + token] = 1;
  if (token == kEndIdx && a.finished[ob] == 0) {
    a.finished[ob] = 1;
    atomicSub(&a.st->n_unfinished, 1);
  }
  return token;
}


// Plain greedy choice of row b (temperature < 0.01, no entropy): one warp per row, 16-byte loads, the adjustments
// applied per element, running first-occurrence argmax.  Every lane returns the token.
// part / nparts: the row's vocabulary is dealt to nparts warps in interleaved blocks of 512 ids; best_out (optional)
// receives the value of the returned id so that the warps' candidates can be merged with arg_better.
__device__ __forceinline__ int greedy_row_token(const SamplerArgs& a, int b, int step, int lane, int part = 0, int nparts = 1,
                                                float* best_out = nullptr) {
  const RowCtx c = make_row_ctx(a, b, step);
  const float4* lg = reinterpret_cast<const float4*>(a.logits + (size_t)b * a.ldl);
  const uint32_t* mk4 = reinterpret_cast<const uint32_t*>(c.mk);
  const uint32_t* sn4 = reinterpret_cast<const uint32_t*>(c.seen_row);
  const bool scale = a.temperature != 1.0f;
  // The IEEE division below is the expensive part of this loop: with a type mask most elements are -inf, which takes the
  // division's slow path (ncu: 71 us per step at 4096 rows, an order of magnitude above the traffic).  For T > 0,
  // x -> fl(x / T) is monotone and a lane meets its ids in increasing order, so an element whose RAW value does not exceed
  // the raw value of the lane's current best can neither beat it nor win a tie (its id is larger): only new raw maxima
  // (and NaNs, which argmax treats as the maximum) are divided and compared.  Same result, bit for bit.
  const bool lazy = a.temperature > 0.f;
  float bv = -INFINITY, braw = -INFINITY;
  int bi = INT_MAX;
  const int n4 = a.V >> 2;
  constexpr int UNR = 4;
  for (int i0 = lane + 32 * UNR * part; i0 < n4; i0 += 32 * UNR * nparts) {
    float4 x[UNR];
    uint32_t m[UNR], sn[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + 32 * u;
      const bool ok = i < n4;
      x[u] = ok ? __ldcg(lg + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      m[u] = (ok && mk4 != nullptr) ? mk4[i] : 0x01010101u;
      sn[u] = (ok && sn4 != nullptr) ? sn4[i] : 0u;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int i = i0 + 32 * u;
      if (i < n4) {
        const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int v = 4 * i + e;
          const float raw = adjust_logit(c, v, xs[e], (m[u] >> (8 * e)) & 0xffu, (sn[u] >> (8 * e)) & 0xffu);
          if (lazy && bi != INT_MAX && !(raw > braw) && !isnan(raw)) continue;
          const float l = scale ? raw / a.temperature : raw;    // (:1485-1486)
          if (arg_better(l, v, bv, bi)) { bv = l; bi = v; braw = raw; }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (best_out != nullptr) *best_out = bv;
  return bi;
}

}  // namespace scv
