// Chemistry-constraint rewards of a rollout on the device (second half).
//
// Reference: compute_constraint_rewards, src/superconductor/losses/constraint_rewards.py:629-676, which adds up A1
// duplicate elements (:270-303), A2 non-canonical fractions (:306-379), A4 reducible integer formulas (:382-459), A7
// impossible combinations (:462-507) and the family rules B1-B8 (:510-626) over the formula parser (:172-267).
// Except for A1 the reference copies the tokens to the host and walks every row with Python loops and .item() calls,
// inside the RL step (scripts/train_v12_clean.py:2754-2766, 2990-3007).  Here a CTA stages 128 rows of tokens + mask
// into shared memory with coalesced loads (one word per position, odd row pitch: conflict-free when every thread walks
// its own row) and one thread per row runs the same four scans and the rules on the parsed composition.
// Amounts and thresholds are doubles like the Python floats of the reference; integer work (gcd, digit runs) is exact
// up to 19 digits.
#include "common.cuh"
#include "../../include/scvae_b200.h"

namespace scv {
namespace {

constexpr int kRows = 128;                   // rows (= threads) per CTA
constexpr int kMaskBit = 1 << 30;

// composition entries the rules look at; slot order is the order of kTrackedZ
constexpr int kTracked = 18;
__constant__ int kTrackedZ[kTracked] = {8, 38, 20, 29, 23, 3, 6, 13, 41, 50, 14, 32, 25, 26, 27, 28, 9, 81};
enum Slot { S_O, S_Sr, S_Ca, S_Cu, S_V, S_Li, S_C, S_Al, S_Nb, S_Sn, S_Si, S_Ge, S_Mn, S_Fe, S_Co, S_Ni, S_F, S_Tl };

struct Row {
  const int* w; int L;
  __device__ __forceinline__ bool masked(int i) const { return (w[i] & kMaskBit) != 0; }
  __device__ __forceinline__ int tok(int i) const { return w[i] & (kMaskBit - 1); }
};

__device__ __forceinline__ unsigned long long gcd64(unsigned long long a, unsigned long long b) {
  while (b != 0) { const unsigned long long t = a % b; a = b; b = t; }
  return a;
}
__device__ __forceinline__ unsigned long long push_digit(unsigned long long v, int d) {
  return v > (0xffffffffffffffffull - 9) / 10 ? 0xffffffffffffffffull : v * 10 + (unsigned long long)d;   // saturates
}

struct Composition {
  unsigned long long present[2];             // bit z: element with atomic number z occurs
  double amt[kTracked];                      // amount attached to its last occurrence
  __device__ __forceinline__ bool has(int z) const { return (present[z >> 6] >> (z & 63)) & 1ull; }
};

__global__ void __launch_bounds__(kRows)
constraint_rows_kernel(const long long* __restrict__ tokens, const unsigned char* __restrict__ mask, int B, int L, long long ld,
                       scv_constraint_config c, const float* __restrict__ frac_values, int n_frac_values,
                       const float* __restrict__ family, int n_fam, float* __restrict__ out) {
  extern __shared__ int sw[];
  const int Ls = L | 1;
  const int row0 = blockIdx.x * kRows;
  const int rows = min(kRows, B - row0);
  for (int i = threadIdx.x; i < rows * L; i += kRows) {            // coalesced: consecutive threads, consecutive positions
    const int r = i / L, p = i - r * L;
    const long long t = tokens[(long long)(row0 + r) * ld + p];
    const int tt = t < 0 ? kMaskBit - 1 : t > kMaskBit - 2 ? kMaskBit - 2 : (int)t;      // ids outside every range stay outside
    sw[r * Ls + p] = tt | (mask[(long long)(row0 + r) * ld + p] != 0 ? kMaskBit : 0);
  }
  __syncthreads();
  if ((int)threadIdx.x >= rows) return;
  const Row R{sw + threadIdx.x * Ls, L};
  const bool sem = c.use_semantic_fractions != 0;
  auto is_el = [&](int t) { return t >= c.element_start && t <= c.element_end; };
  auto is_dg = [&](int t) { return t >= c.digit_start && t <= c.digit_end; };

  // ---- the formula parser (:172-267): elements up to the first END / unmasked position, amount of the last occurrence
  Composition cp;
  cp.present[0] = cp.present[1] = 0ull;
#pragma unroll
  for (int s = 0; s < kTracked; ++s) cp.amt[s] = 0.0;
  for (int i = 0; i < L && R.masked(i) && R.tok(i) != c.end_idx;) {
    const int t = R.tok(i);
    if (!is_el(t)) { ++i; continue; }
    const int z = t - c.element_start + 1;
    double a = 1.0;
    int j = i + 1;
    if (j < L && R.masked(j)) {
      const int n = R.tok(j);
      if (sem) {
        if (is_dg(n)) { a = (double)(n - c.digit_start + 1); ++j; }
        else if (n >= c.fraction_token_start && frac_values != nullptr) {
          if (n < n_frac_values) a = (double)__ldg(frac_values + n);
          ++j;
        }
      } else if (n == c.lparen_idx) {
        ++j;
        unsigned long long num = 0, den = 0;
        int n_num = 0, n_den = 0;
        bool in_num = true;
        while (j < L) {                                          // (no mask test in here, as in the reference)
          const int ch = R.tok(j);
          if (ch == c.slash_idx) in_num = false;
          else if (ch == c.rparen_idx) { ++j; break; }
          else if (is_dg(ch)) {
            if (in_num) { num = push_digit(num, ch - c.digit_start); ++n_num; }
            else { den = push_digit(den, ch - c.digit_start); ++n_den; }
          } else break;
          ++j;
        }
        if (n_num > 0 && n_den > 0 && den > 0) a = (double)num / (double)den;
      } else if (is_dg(n)) {
        unsigned long long v = 0;
        while (j < L && is_dg(R.tok(j))) { v = push_digit(v, R.tok(j) - c.digit_start); ++j; }
        a = (double)v;
      }
      i = j;
    } else {
      ++i;
    }
    if (z >= 1 && z < 128) {
      cp.present[z >> 6] |= 1ull << (z & 63);
#pragma unroll
      for (int s = 0; s < kTracked; ++s)
        if (kTrackedZ[s] == z) cp.amt[s] = a;
    }
  }

  float total = 0.f;
  // ---- A1: an element id at two masked positions anywhere in the row (:270-303)
  if (c.a1_enabled) {
    unsigned long long seen[2] = {0ull, 0ull};
    bool dup = false;
    for (int i = 0; i < L; ++i) {
      const int t = R.tok(i);
      if (R.masked(i) && is_el(t)) {
        const int z = (t - c.element_start) & 127;
        const unsigned long long bit = 1ull << (z & 63);
        dup = dup || (seen[z >> 6] & bit) != 0;
        seen[z >> 6] |= bit;
      }
    }
    total += (float)((dup ? 1.0 : 0.0) * c.a1_penalty);
  }
  // ---- A2: closed "( num / den )" groups that are not in lowest terms; pre-V13 vocabulary only (:306-379)
  if (c.a2_enabled) {
    int violations = 0;
    if (!sem) {
      for (int i = 0; i < L && R.masked(i);) {
        if (R.tok(i) != c.lparen_idx) { ++i; continue; }
        int j = i + 1, n_num = 0, n_den = 0;
        unsigned long long num = 0, den = 0;
        bool in_num = true, closed = false;
        while (j < L && R.masked(j)) {
          const int ch = R.tok(j);
          if (ch == c.slash_idx) in_num = false;
          else if (ch == c.rparen_idx) { closed = true; ++j; break; }
          else if (is_dg(ch)) {
            if (in_num) { num = push_digit(num, ch - c.digit_start); ++n_num; }
            else { den = push_digit(den, ch - c.digit_start); ++n_den; }
          } else break;
          ++j;
        }
        if (closed && n_num > 0 && n_den > 0 && den > 0 && gcd64(num, den) > 1) ++violations;
        i = j > i + 1 ? j : i + 1;
      }
    }
    total += (float)((double)violations * c.a2_penalty_per_violation);      // (0 violations: the reference adds zeros)
  }
  // ---- A4: integer-only formula with >= 2 elements whose subscripts share a factor (:382-459)
  if (c.a4_enabled) {
    bool fractions = false;
    int n_sub = 0;
    unsigned long long g = 0;
    for (int i = 0; i < L && R.masked(i) && R.tok(i) != c.end_idx;) {
      const int t = R.tok(i);
      if ((sem && t >= c.fraction_token_start) || (!sem && t == c.lparen_idx)) { fractions = true; break; }
      if (!is_el(t)) { ++i; continue; }
      int j = i + 1;
      unsigned long long v = 1;
      if (sem) {
        if (j < L && R.masked(j) && is_dg(R.tok(j))) { v = (unsigned long long)(R.tok(j) - c.digit_start + 1); ++j; }
      } else {
        unsigned long long d = 0;
        int nd = 0;
        while (j < L && R.masked(j) && is_dg(R.tok(j))) { d = push_digit(d, R.tok(j) - c.digit_start); ++nd; ++j; }
        if (nd > 0) v = d;
      }
      g = n_sub == 0 ? v : gcd64(g, v);
      ++n_sub;
      i = j;
    }
    total += (float)((!fractions && n_sub >= 2 && g > 1) ? c.a4_penalty : 0.0);
  }
  const bool mag_present[4] = {cp.has(25), cp.has(26), cp.has(27), cp.has(28)};
  const double mag_amt[4] = {cp.amt[S_Mn], cp.amt[S_Fe], cp.amt[S_Co], cp.amt[S_Ni]};
  // ---- A7: F with Tl; magnetic 3d metal above 2 % and above half the Cu amount next to Cu (:462-507)
  if (c.a7_enabled) {
    bool bad = cp.has(9) && cp.has(81);
    if (!bad && cp.has(29) && cp.amt[S_Cu] > 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) bad = bad || (mag_present[k] && mag_amt[k] > 0.02 && mag_amt[k] > 0.5 * cp.amt[S_Cu]);
    }
    total += (float)(bad ? c.a7_penalty : 0.0);
  }
  // ---- B1-B8: rules of the predicted family, when the classifier is confident (:510-626)
  if (family != nullptr && c.family_enabled) {
    const float* pr = family + (long long)(row0 + threadIdx.x) * n_fam;
    int fam = 0;
    float best = pr[0];
    for (int k = 1; k < n_fam; ++k)
      if (pr[k] > best) { best = pr[k]; fam = k; }               // first maximum
    if (!((double)best < c.confidence_threshold)) {
      auto mag_over = [&](double lim) {
        bool r = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) r = r || (mag_present[k] && mag_amt[k] > lim);
        return r;
      };
      double p = 0.0;
      if (fam == 2) {                                            // YBCO: oxygen content
        if (cp.amt[S_O] > 0 && cp.amt[S_O] < 6.35) p += c.b_penalty[0];
      } else if (fam == 3) {                                     // LSCO: Sr doping window
        if (cp.has(38) && (cp.amt[S_Sr] < 0.055 || cp.amt[S_Sr] > 0.27)) p += c.b_penalty[1];
      } else if (fam == 4) {                                     // BSCCO: Ca against Cu - 1
        if (cp.has(20) && cp.has(29) && fabs(cp.amt[S_Ca] - (cp.amt[S_Cu] - 1)) > 0.3) p += c.b_penalty[2];
      } else if (fam == 6) {                                     // Hg cuprates
        if (cp.amt[S_V] > 0.30) p += c.b_penalty[3];
      } else if (fam == 5) {                                     // Tl cuprates
        if (cp.amt[S_V] > 0.30) p += c.b_penalty[4];
        if (cp.amt[S_Li] > 0.10) p += c.b_penalty[4];
        if (mag_over(0.10)) p += c.b_penalty[4];
      } else if (fam == 8) {                                     // iron pnictides
        if (cp.has(8) && cp.amt[S_O] < 0.7 && cp.amt[S_O] != 1.0) p += c.b_penalty[5];
      } else if (fam == 10) {                                    // MgB2
        if (cp.amt[S_C] > 0.125) p += c.b_penalty[6];
        if (cp.amt[S_Al] > 0.50) p += c.b_penalty[6];
        if (mag_over(0.05)) p += c.b_penalty[6];
      } else if (fam == 1) {                                     // A15: (Nb + V) : (Sn + Al + Si + Ge) = 3 : 1 within 10 %
        double a_tot = 0.0, b_tot = 0.0;
        if (cp.has(41)) a_tot += cp.amt[S_Nb];
        if (cp.has(23)) a_tot += cp.amt[S_V];
        if (cp.has(50)) b_tot += cp.amt[S_Sn];
        if (cp.has(13)) b_tot += cp.amt[S_Al];
        if (cp.has(14)) b_tot += cp.amt[S_Si];
        if (cp.has(32)) b_tot += cp.amt[S_Ge];
        if (a_tot > 0 && b_tot > 0 && fabs(a_tot / b_tot - 3.0) > 0.3) p += c.b_penalty[7];
      }
      if (p < 0) total += (float)p;
    }
  }
  out[row0 + threadIdx.x] = total;
}

}  // namespace
}  // namespace scv

extern "C" int scv_constraint_rewards(const int64_t* sampled, const uint8_t* mask, int32_t batch, int32_t seq_len,
                                      int64_t row_stride, const scv_constraint_config* config,
                                      const float* fraction_values, int32_t n_fraction_values, const float* family_probs,
                                      int32_t n_families, float* rewards, void* stream) {
  SCV_REQUIRE(sampled && mask && config && rewards && batch > 0 && seq_len > 0 && row_stride >= seq_len,
              "constraint rewards: bad arguments");
  SCV_REQUIRE(family_probs == nullptr || n_families > 0, "constraint rewards: empty family probabilities");
  const size_t smem = (size_t)scv::kRows * (seq_len | 1) * sizeof(int);
  SCV_REQUIRE(smem <= 200 * 1024, "constraint rewards: rows of %d positions do not fit in shared memory", seq_len);
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SCV_CUDA(cudaFuncSetAttribute(scv::constraint_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_smem = 200 * 1024;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SCV_CUDA(scv::launch_k(scv::constraint_rows_kernel, dim3((unsigned)((batch + scv::kRows - 1) / scv::kRows)), dim3(scv::kRows),
                         smem, s, reinterpret_cast<const long long*>(sampled), mask, (int)batch, (int)seq_len,
                         (long long)row_stride, *config, fraction_values, (int)n_fraction_values, family_probs,
                         (int)n_families, rewards));
  SCV_LAUNCH_CHECK();
  return 0;
}
