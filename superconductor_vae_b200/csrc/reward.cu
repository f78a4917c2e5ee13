// Token-level rollout reward on the device.
//
// Reference: compute_reward_gpu_native, src/superconductor/losses/reward_gpu_native.py:448-722 with its helpers
// (:144-445), called on every RLOO / SCST rollout right after sample_for_reinforce
// (scripts/train_v12_clean.py:2745-2752, 2829-2836, 2942-2950).  There it is ~60 whole-batch tensor expressions over
// [B, L] int64 tokens (each a kernel launch and an HBM round trip); here it is ONE pass: a warp per row, positions
// across the lanes (coalesced 8-byte loads), counts by ballot + popc, the two per-position penalty sums by warp
// reduction, then one lane evaluates the branch structure of the reference on the row's scalars.  Integer / boolean
// work is exact; the float results differ from the reference's only by summation order and powf (tests: 1e-3 abs).
#include "common.cuh"
#include "../../include/scvae_b200.h"

namespace scv {
namespace {

// pre-V13 vocabulary ids of the digit-level fraction penalties (reward_gpu_native.py:34-39)
constexpr long long kLParen = 4, kRParen = 5, kSlash = 16, kDigit0 = 138, kDigit9 = 147;

struct RowScan {                      // everything the branch structure needs about one row
  int n_matches, n_mism, n_valid;
  int s_end, t_end;                   // first masked END position, -1 if none
  int n_el, n_int, n_fr, n_sp;        // mismatches by TARGET token type (:383-389)
  int ph_match, ph_total;             // phased curriculum counts (:609-632)
  int structure_errors, digit_errors; // (:541, :553-554)
  float frac_value_pen;               // compute_fraction_value_penalty (:279-342) or compute_semantic_digit_penalty (:215-276)
};

__device__ __forceinline__ int warp_isum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
reward_rows_kernel(const long long* __restrict__ sampled, const long long* __restrict__ target,
                   const unsigned char* __restrict__ mask, int B, int L, long long ld, scv_reward_config c, int end_idx,
                   int semantic, int frac_start, const float* __restrict__ frac_values, int n_frac_values,
                   float* __restrict__ rewards) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const long long* sr = sampled + (long long)row * ld;
  const long long* tr = target + (long long)row * ld;
  const unsigned char* mr = mask + (long long)row * ld;
  RowScan a = {};
  a.s_end = a.t_end = -1;
  int depth_carry = 0;                                   // parenthesis depth of the target before this chunk (:168-170)
  float pen = 0.f;
  // ---- pass 1: counts, END positions, per-position penalties
  for (int p0 = 0; p0 < L; p0 += 32) {
    const int p = p0 + lane;
    const bool in = p < L;
    const long long s = in ? sr[p] : 0, t = in ? tr[p] : 0;
    const bool m = in && mr[p] != 0;
    const bool eq = s == t, mis = m && !eq;
    a.n_matches += __popc(__ballot_sync(0xffffffffu, m && eq));
    a.n_mism += __popc(__ballot_sync(0xffffffffu, mis));
    a.n_valid += __popc(__ballot_sync(0xffffffffu, m));
    const unsigned se = __ballot_sync(0xffffffffu, m && s == end_idx), te = __ballot_sync(0xffffffffu, m && t == end_idx);
    if (a.s_end < 0 && se != 0u) a.s_end = p0 + __ffs(se) - 1;
    if (a.t_end < 0 && te != 0u) a.t_end = p0 + __ffs(te) - 1;
    const bool t_el = t >= c.v14_element_start && t <= c.v14_element_end;
    const bool t_int = t >= c.v14_integer_start && t <= c.v14_integer_end;
    const bool t_fr = t >= c.v14_fraction_start;
    a.n_el += __popc(__ballot_sync(0xffffffffu, mis && t_el));
    a.n_int += __popc(__ballot_sync(0xffffffffu, mis && t_int));
    a.n_fr += __popc(__ballot_sync(0xffffffffu, mis && t_fr));
    a.n_sp += __popc(__ballot_sync(0xffffffffu, mis && !t_el && !t_int && !t_fr));
    const bool pm = m && (c.reward_phase == 1 ? t_el : c.reward_phase == 2 ? (t_el || t_int || t_fr) : true);
    a.ph_match += __popc(__ballot_sync(0xffffffffu, pm && eq));
    a.ph_total += __popc(__ballot_sync(0xffffffffu, pm));
    if (semantic) {
      if (in && m && !eq && t >= frac_start) {
        const long long hi = n_frac_values - 1;
        const float sv = __ldg(frac_values + (s < 0 ? 0 : s > hi ? hi : s)), tv = __ldg(frac_values + (t < 0 ? 0 : t > hi ? hi : t));
        const float scale = 1.0f + c.semantic_digit_scale * fminf(fabsf(sv - tv), 20.0f) / 20.0f;
        pen += c.fraction_digit_penalty * scale;
      }
    } else {
      const bool lp = m && t == kLParen, rp = m && t == kRParen;
      int d = (lp ? 1 : 0) - (rp ? 1 : 0);               // inclusive scan of the depth deltas over the chunk
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, d, o);
        if (lane >= o) d += v;
      }
      const int depth = depth_carry + d;
      depth_carry += __shfl_sync(0xffffffffu, d, 31);
      const bool in_frac = m && (depth > 0 || lp);
      const bool structure = m && (t == kLParen || t == kRParen || t == kSlash);
      a.structure_errors += __popc(__ballot_sync(0xffffffffu, mis && structure));
      const bool digit_mis = mis && t >= kDigit0 && t <= kDigit9 && in_frac;
      a.digit_errors += __popc(__ballot_sync(0xffffffffu, digit_mis));
      if (digit_mis && c.use_semantic_digit_penalty) {
        const long long sdl = s - kDigit0, tdl = t - kDigit0;
        const float sd = (float)(sdl < 0 ? 0 : sdl > 9 ? 9 : sdl), td = (float)(tdl < 0 ? 0 : tdl > 9 ? 9 : tdl);
        pen += c.fraction_digit_penalty * (1.0f + c.semantic_digit_scale * fabsf(sd - td) / 9.0f);
      }
    }
  }
  pen = warp_sum(pen);
  a.frac_value_pen = pen;
  const float s_end_pos = a.s_end >= 0 ? (float)a.s_end : (float)a.n_valid;     // (:506-515)
  const float t_end_pos = a.t_end >= 0 ? (float)a.t_end : (float)a.n_valid;
  const int t_end_col = (int)t_end_pos, s_end_col = (int)s_end_pos;
  // ---- pass 2: what depends on the END positions (:576-581, 592-600, 638-643); the row is in L1 / L2 now
  bool prefix_ok = true, prefix2_ok = true;
  int content_matches = 0;
  for (int p0 = 0; p0 < L; p0 += 32) {
    const int p = p0 + lane;
    const bool in = p < L;
    const bool m = in && mr[p] != 0;
    const bool eq = in ? sr[p] == tr[p] : true;
    prefix_ok = prefix_ok && __all_sync(0xffffffffu, !in || eq || !(p < t_end_col) || !m);
    prefix2_ok = prefix2_ok && __all_sync(0xffffffffu, !in || eq || !(p < s_end_col) || !m);
    content_matches += __popc(__ballot_sync(0xffffffffu, in && eq && p <= t_end_col && m));
  }
  if (lane != 0) return;
  // ---- the branch structure of the reference on the row's scalars
  const bool exact = a.n_mism == 0;
  const float length_diff = fabsf(s_end_pos - t_end_pos);
  float frac_pen;
  if (semantic) frac_pen = a.frac_value_pen;
  else frac_pen = (c.use_semantic_digit_penalty ? a.frac_value_pen : (float)a.digit_errors * c.fraction_digit_penalty) +
                  (float)a.structure_errors * c.fraction_structure_penalty;
  const bool length_only = prefix_ok && s_end_pos > t_end_pos && !exact;
  const float length_only_reward = fmaxf(c.length_only_base_reward - fmaxf(s_end_pos - t_end_pos, 0.f) * c.length_only_per_extra,
                                         c.length_only_floor);
  const float length_pen = length_diff * c.length_mismatch_penalty;
  float r = exact ? c.exact_match : 0.f;
  if (c.v14 && c.use_continuous_reward) {                // (:564-660)
    if (length_only) r = length_only_reward;
    const bool too_short = prefix2_ok && s_end_pos < t_end_pos && a.s_end >= 0 && !exact && !length_only;
    if (too_short)
      r = fmaxf(c.too_short_base_reward - fmaxf(t_end_pos - s_end_pos, 0.f) * c.too_short_per_missing, c.too_short_floor);
    if (!exact && !length_only && !too_short) {
      float n_correct, n_total;
      if (c.use_phased_curriculum && c.reward_phase < 3) { n_correct = (float)a.ph_match; n_total = (float)a.ph_total; }
      else { n_correct = (float)content_matches; n_total = fmaxf(t_end_pos + 1.0f, 1.0f); }
      float ratio = n_correct / fmaxf(n_total, 1.0f);     // _compute_continuous_reward (:406-445)
      ratio = fminf(fmaxf(ratio, 0.f), 1.f);
      const float sharp = (c.use_phased_curriculum && c.reward_phase >= 3) ? c.phase3_sharpness : c.sharpness;
      const float base = c.max_reward * powf(ratio, sharp);
      float tp = (float)a.n_el * c.element_error_penalty + (float)a.n_int * c.integer_error_penalty +
                 (float)a.n_sp * c.special_error_penalty;                       // (:392-396)
      if (!semantic) tp += (float)a.n_fr * c.fraction_error_penalty;            // (:398-403)
      r = fmaxf(base + tp + frac_pen + length_pen, -100.0f);
    }
  } else {                                               // tiered rewards (:662-722)
    if (length_only) r = length_only_reward;
    const bool not_handled = !exact && !length_only;
    if (not_handled && a.n_mism == 1) r = c.near_exact_1 + frac_pen + length_pen;
    if (not_handled && a.n_mism == 2) r = c.near_exact_2 + frac_pen + length_pen;
    if (not_handled && a.n_mism == 3) r = c.near_exact_3 + frac_pen + length_pen;
    if (not_handled && a.n_mism > 3) {
      float tr_ = (float)a.n_matches * c.token_correct + (float)a.n_mism * c.token_penalty;
      tr_ += length_diff * c.length_mismatch_penalty;
      tr_ += frac_pen;
      r = fminf(fmaxf(tr_, -100.0f), 5.0f);
    }
  }
  rewards[row] = r;
}

}  // namespace
}  // namespace scv

extern "C" int scv_reward_tokens(const int64_t* sampled, const int64_t* target, const uint8_t* mask, int32_t batch,
                                 int32_t seq_len, int64_t row_stride, const scv_reward_config* config, int32_t end_idx,
                                 int32_t use_semantic_fractions, int32_t fraction_token_start,
                                 const float* fraction_values, int32_t n_fraction_values, float* rewards, void* stream) {
  SCV_REQUIRE(sampled && target && mask && config && rewards && batch > 0 && seq_len > 0 && row_stride >= seq_len,
              "reward: bad arguments");
  const int semantic = use_semantic_fractions != 0 && fraction_values != nullptr;      // (:520)
  SCV_REQUIRE(!semantic || n_fraction_values > 0, "reward: empty fraction value table");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long threads = (long long)batch * 32;
  SCV_CUDA(scv::launch_k(scv::reward_rows_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, s,
                         reinterpret_cast<const long long*>(sampled), reinterpret_cast<const long long*>(target), mask,
                         (int)batch, (int)seq_len, (long long)row_stride, *config, (int)end_idx, semantic,
                         (int)fraction_token_start, fraction_values, (int)n_fraction_values, rewards));
  SCV_LAUNCH_CHECK();
  return 0;
}
