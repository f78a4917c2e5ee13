// Spherical interpolation of latent rows for latent-space candidate generation.
// Reference: scripts/holdout/holdout_search.py:128-146 (identical copy in
// notebooks/generative_evaluation.ipynb cell 12).  One CTA per output row.
#include "../../include/scvae_b200.h"
#include "common.cuh"

using namespace scv;

namespace {

constexpr int kThreads = 256;

struct RowStats { float n1, n2, dot; };

__device__ RowStats row_stats(const float* a, const float* b, int dim, float* red) {
  float s1 = 0.f, s2 = 0.f, d = 0.f;
  for (int i = threadIdx.x; i < dim; i += kThreads) {
    const float x = a[i], y = b[i];
    s1 = fmaf(x, x, s1); s2 = fmaf(y, y, s2); d = fmaf(x, y, d);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2); d = warp_sum(d);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { red[w * 3 + 0] = s1; red[w * 3 + 1] = s2; red[w * 3 + 2] = d; }
  __syncthreads();
  float t1 = 0.f, t2 = 0.f, td = 0.f;
  for (int i = 0; i < kThreads / 32; ++i) { t1 += red[i * 3]; t2 += red[i * 3 + 1]; td += red[i * 3 + 2]; }
  RowStats r;
  r.n1 = sqrtf(t1); r.n2 = sqrtf(t2);
  // dot of the F.normalize()d rows: x / max(||x||, 1e-12)
  r.dot = td / (fmaxf(r.n1, 1e-12f) * fmaxf(r.n2, 1e-12f));
  return r;
}

__device__ float omega_of(float dot) { return fmaxf(acosf(fminf(fmaxf(dot, -1.0f), 1.0f)), 1e-6f); }

__global__ void __launch_bounds__(kThreads)
slerp_flag_kernel(const float* anchors, int dim, const int* i1, const int* i2, long long n, int* flag) {
  __shared__ float red[kThreads / 32 * 3];
  const long long r = blockIdx.x;
  if (r >= n) return;
  const RowStats st = row_stats(anchors + (size_t)i1[r] * dim, anchors + (size_t)i2[r] * dim, dim, red);
  if (threadIdx.x == 0 && fabsf(sinf(omega_of(st.dot))) < 1e-6f) atomicOr(flag, 1);   // (:137)
}

__global__ void __launch_bounds__(kThreads)
slerp_rows_kernel(const float* anchors, int dim, const int* i1, const int* i2, const float* t, long long n,
                  float* out, const int* flag) {
  __shared__ float red[kThreads / 32 * 3];
  const long long r = blockIdx.x;
  if (r >= n) return;
  const float* a = anchors + (size_t)i1[r] * dim;
  const float* b = anchors + (size_t)i2[r] * dim;
  float* o = out + (size_t)r * dim;
  const float tt = t[r];
  if (*flag != 0) {                                   // batch-global lerp fallback (:137-138)
    for (int i = threadIdx.x; i < dim; i += kThreads) o[i] = (1.0f - tt) * a[i] + tt * b[i];
    return;
  }
  const RowStats st = row_stats(a, b, dim, red);
  const float om = omega_of(st.dot), so = sinf(om);
  const float s1 = sinf((1.0f - tt) * om) / so, s2 = sinf(tt * om) / so;
  const float mag = (1.0f - tt) * st.n1 + tt * st.n2;
  const float inv1 = 1.0f / fmaxf(st.n1, 1e-12f), inv2 = 1.0f / fmaxf(st.n2, 1e-12f);
  for (int i = threadIdx.x; i < dim; i += kThreads) o[i] = (s1 * (a[i] * inv1) + s2 * (b[i] * inv2)) * mag;
}

}  // namespace

extern "C" int scv_slerp_rows(const float* anchors, int32_t dim, const int32_t* i1, const int32_t* i2, const float* t,
                              int64_t n_rows, float* out, int32_t* fallback_flag_dev, void* stream) {
  SCV_REQUIRE(anchors && i1 && i2 && t && out && fallback_flag_dev && dim > 0 && n_rows > 0, "slerp: bad arguments");
  SCV_REQUIRE(n_rows < (1ll << 31), "slerp: too many rows for one call");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SCV_CUDA(cudaMemsetAsync(fallback_flag_dev, 0, sizeof(int), s));
  slerp_flag_kernel<<<(unsigned)n_rows, kThreads, 0, s>>>(anchors, dim, i1, i2, n_rows, fallback_flag_dev);
  SCV_LAUNCH_CHECK();
  slerp_rows_kernel<<<(unsigned)n_rows, kThreads, 0, s>>>(anchors, dim, i1, i2, t, n_rows, out, fallback_flag_dev);
  SCV_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Candidate post-processing at scale (scripts/holdout/holdout_search.py:88-99 turns every generated row
// into a string with a Python loop, the wall-clock bottleneck at 1 M candidates).  A formula is the tokens up to the
// first END with PAD / START skipped (tokenizer/fraction_tokenizer.py:478-519), so rows are canonicalised (everything
// from the first END on -> PAD, ids narrowed to int16) and hashed on the device; the host then groups identical rows
// and builds strings for the unique ones only.  One warp per row; hash = 64-bit FNV-1a over the canonical ids.
// ---------------------------------------------------------------------------------------------
namespace scv {
static __global__ void canonical_hash_kernel(const long long* __restrict__ tokens, long long n_rows, int L,
                                      short* __restrict__ canon, unsigned long long* __restrict__ hash,
                                      int* __restrict__ length) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const long long* tr = tokens + row * L;
  // first END position of the row (L if none)
  int first_end = L;
  for (int p0 = 0; p0 < L; p0 += 32) {
    const int p = p0 + lane;
    const bool is_end = p < L && tr[p] == kEndIdx;
    const unsigned m = __ballot_sync(0xffffffffu, is_end);
    if (m != 0u) { first_end = p0 + __ffs(m) - 1; break; }
  }
  for (int p = lane; p < L; p += 32) canon[row * L + p] = p < first_end ? (short)tr[p] : (short)0;
  if (lane == 0) {
    unsigned long long h = 1469598103934665603ull;
    for (int p = 0; p < first_end; ++p) {
      const unsigned long long t = (unsigned long long)tr[p];
      h = (h ^ (t & 0xffull)) * 1099511628211ull;
      h = (h ^ ((t >> 8) & 0xffull)) * 1099511628211ull;
    }
    hash[row] = h;
    if (length != nullptr) length[row] = first_end;
  }
}
}  // namespace scv

extern "C" int scv_tokens_canonical_hash(const int64_t* tokens, int64_t n_rows, int32_t row_len, int16_t* canonical,
                                         uint64_t* hash, int32_t* length, void* stream) {
  SCV_REQUIRE(tokens && canonical && hash && n_rows > 0 && row_len > 0, "canonical_hash: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long threads = n_rows * 32;
  scv::canonical_hash_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const long long*>(tokens), (long long)n_rows, row_len, canonical,
      reinterpret_cast<unsigned long long*>(hash), length);
  SCV_LAUNCH_CHECK();
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Compositional similarity of candidate x target formulas (scripts/holdout/holdout_search.py:149-182):
//   0.5 * |shared elements| / |all elements| + 0.5 * sum over shared elements of min(a_e / sum a, b_e / sum b).
// Compositions arrive as dense rows over the batch's element columns, absent elements marked by a negative amount
// (amounts parsed from a formula are never negative).  Doubles, like the reference's Python floats.  One CTA per
// candidate row: the row and its total are staged in shared memory once, threads walk the targets.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128)
element_similarity_kernel(const double* __restrict__ a, const double* __restrict__ b, int n_a, int n_b, int n_el,
                          double* __restrict__ out) {
  extern __shared__ double row_a[];
  __shared__ double total_a;
  __shared__ int present_a;
  const int i = blockIdx.x;
  if (threadIdx.x == 0) { total_a = 0.0; present_a = 0; }
  __syncthreads();
  for (int e = threadIdx.x; e < n_el; e += blockDim.x) row_a[e] = a[(size_t)i * n_el + e];
  __syncthreads();
  if (threadIdx.x == 0) {                      // sequential, in column order: the same sum for every thread
    double t = 0.0; int c = 0;
    for (int e = 0; e < n_el; ++e) if (row_a[e] >= 0.0) { t += row_a[e]; ++c; }
    total_a = t; present_a = c;
  }
  __syncthreads();
  const double ta = fmax(total_a, 1e-8);
  for (int j = threadIdx.x; j < n_b; j += blockDim.x) {
    const double* rb = b + (size_t)j * n_el;
    double tb = 0.0; int cb = 0;
    for (int e = 0; e < n_el; ++e) if (rb[e] >= 0.0) { tb += rb[e]; ++cb; }
    double sim = 0.0;
    if (present_a > 0 && cb > 0) {
      tb = fmax(tb, 1e-8);
      int shared = 0; double frac = 0.0;
      for (int e = 0; e < n_el; ++e)
        if (row_a[e] >= 0.0 && rb[e] >= 0.0) { ++shared; frac += fmin(row_a[e] / ta, rb[e] / tb); }
      sim = 0.5 * (double)shared / (double)(present_a + cb - shared) + 0.5 * frac;
    }
    out[(size_t)i * n_b + j] = sim;
  }
}
}  // namespace

extern "C" int scv_element_similarity(const double* comp_a, const double* comp_b, int32_t n_a, int32_t n_b, int32_t n_elements,
                                      double* out, void* stream) {
  SCV_REQUIRE(comp_a && comp_b && out && n_a > 0 && n_b > 0 && n_elements > 0, "element_similarity: bad arguments");
  SCV_REQUIRE(n_elements <= 4096, "element_similarity: %d element columns (at most 4096)", n_elements);
  element_similarity_kernel<<<n_a, 128, (size_t)n_elements * sizeof(double), static_cast<cudaStream_t>(stream)>>>(
      comp_a, comp_b, n_a, n_b, n_elements, out);
  SCV_LAUNCH_CHECK();
  return 0;
}
