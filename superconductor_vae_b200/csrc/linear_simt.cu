// y = act(x * W^T + bias) (+ residual): fp32 activations, bf16 weights, fp32 accumulate, CUDA cores.
//
// This is the exactness baseline for every projection on the decode path
// (F.linear call sites: models/autoregressive_decoder.py:1259, 1293, 1302-1307, 1312, 1413, 1417, 1439
// and the memory builders :800-859).  Products bf16 x fp32 are exact in the FMA's internal width, so the
// only difference from the fp32 oracle is summation order.  The tcgen05 path (gemm_tcgen05.cu) replaces it
// where the batch makes the projection a dense contraction; this kernel remains for ragged/small shapes
// (K = 1, 13, 24, 145, 513, 2214 ...) and as the in-library cross-check.
#include <cstdlib>

#include "common.cuh"

namespace scv {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_simt_kernel(LinearArgs a, int x_vec_ok) {
  pdl_wait();
  if (a.done_flag != nullptr && *a.done_flag != 0) return;
  pdl_launch_dependents();
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TMH = TM / 2, TNH = TN / 2;
  constexpr int PADM = BM + 4, PADN = BN + 4;
  constexpr int A_F4 = BM * BK / 4, W_U4 = BN * BK / 8;
  constexpr int A_PER = (A_F4 + NT - 1) / NT, W_PER = (W_U4 + NT - 1) / NT;
  __shared__ __align__(16) float As[2][BK][PADM];
  __shared__ __align__(16) float Ws[2][BK][PADN];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = a.M, N = a.N, K = a.K;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_PER];
  uint4 rw[W_PER];

  auto load_tiles = [&](int kt) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int idx = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < A_F4) {
        const int row = idx / (BK / 4), kq = idx % (BK / 4);
        const int gm = m0 + row, gk = kt * BK + kq * 4;
        if (gm < M && gk < K) {
          const float* p = a.x + (size_t)gm * a.ldx + gk;
          if (x_vec_ok && gk + 3 < K) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            v.x = p[0];
            if (gk + 1 < K) v.y = p[1];
            if (gk + 2 < K) v.z = p[2];
            if (gk + 3 < K) v.w = p[3];
          }
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < W_PER; ++i) {
      const int idx = tid + i * NT;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (idx < W_U4) {
        const int row = idx / (BK / 8), kh = idx % (BK / 8);
        const int gn = n0 + row, gk = kt * BK + kh * 8;
        if (gn < N && gk < a.ldw) v = *reinterpret_cast<const uint4*>(a.w + (size_t)gn * a.ldw + gk);
      }
      rw[i] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int idx = tid + i * NT;
      if (idx < A_F4) {
        const int row = idx / (BK / 4), kq = idx % (BK / 4);
        As[buf][kq * 4 + 0][row] = ra[i].x;
        As[buf][kq * 4 + 1][row] = ra[i].y;
        As[buf][kq * 4 + 2][row] = ra[i].z;
        As[buf][kq * 4 + 3][row] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < W_PER; ++i) {
      const int idx = tid + i * NT;
      if (idx < W_U4) {
        const int row = idx / (BK / 8), kh = idx % (BK / 8);
        const uint32_t u[4] = {rw[i].x, rw[i].y, rw[i].z, rw[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          Ws[buf][kh * 8 + 2 * j + 0][row] = __uint_as_float(u[j] << 16);
          Ws[buf][kh * 8 + 2 * j + 1][row] = __uint_as_float(u[j] & 0xffff0000u);
        }
      }
    }
  };

  const int KT = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < TMH; ++i) av[h * TMH + i] = As[buf][k][h * (BM / 2) + ty * TMH + i];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < TNH; ++j) bv[h * TNH + j] = Ws[buf][k][h * (BN / 2) + tx * TNH + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
      store_tiles(buf ^ 1);   // the other buffer was last read in iteration kt-1, fenced by the barrier below
      __syncthreads();
    }
  }

#pragma unroll
  for (int hi = 0; hi < 2; ++hi)
#pragma unroll
    for (int i = 0; i < TMH; ++i) {
      const int gm = m0 + hi * (BM / 2) + ty * TMH + i;
      if (gm >= M) continue;
#pragma unroll
      for (int hj = 0; hj < 2; ++hj)
#pragma unroll
        for (int j = 0; j < TNH; ++j) {
          const int gn = n0 + hj * (BN / 2) + tx * TNH + j;
          if (gn >= N) continue;
          float v = acc[hi * TMH + i][hj * TNH + j];
          if (a.bias != nullptr) v += a.bias[gn];
          v = apply_act(v, a.act);
          if (a.residual != nullptr) v += a.residual[(size_t)gm * a.ldr + gn];
          a.y[(size_t)gm * a.ldy + gn] = v;
          if (a.nonfinite_flag != nullptr && nonfinite_hit(v, a.nonfinite_mode)) atomicOr(a.nonfinite_flag, 1);
        }
    }
}

// Narrow heads (N <= 16: token-type logits, stop / site-dup logit, Tc, competence, family heads ...): one warp per
// row, the row in registers, one warp reduction per output.  The tiled kernel above would spend a 128 x 128 tile on
// 1-13 useful columns (ncu r01c: 45 us for the 5 token-type logits of 4096 rows).
constexpr int kRowWarpMaxK = 1024;
__global__ void __launch_bounds__(256) linear_rowwarp_kernel(LinearArgs a) {
  pdl_wait();
  if (a.done_flag != nullptr && *a.done_flag != 0) return;
  pdl_launch_dependents();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= a.M) return;
  float xv[kRowWarpMaxK / 32];
  const float* xr = a.x + (size_t)row * a.ldx;
#pragma unroll
  for (int j = 0; j < kRowWarpMaxK / 32; ++j) {
    const int k = lane + 32 * j;
    xv[j] = k < a.K ? xr[k] : 0.f;
  }
  for (int n = 0; n < a.N; ++n) {
    const __nv_bfloat16* wr = a.w + (size_t)n * a.ldw;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kRowWarpMaxK / 32; ++j) {
      const int k = lane + 32 * j;
      if (k < a.K) acc = fmaf(xv[j], __bfloat162float(wr[k]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc;
      if (a.bias != nullptr) v += a.bias[n];
      v = apply_act(v, a.act);
      if (a.residual != nullptr) v += a.residual[(size_t)row * a.ldr + n];
      a.y[(size_t)row * a.ldy + n] = v;
      if (a.nonfinite_flag != nullptr && nonfinite_hit(v, a.nonfinite_mode)) atomicOr(a.nonfinite_flag, 1);
    }
  }
}

int launch_linear_simt(const LinearArgs& a, cudaStream_t s) {
  static const bool rowwarp = [] { const char* e = getenv("SCV_ROWWARP"); return e ? atoi(e) != 0 : true; }();
  if (rowwarp && a.N <= 16 && a.K <= kRowWarpMaxK && a.M > 0 && a.K > 0) {
    ProfScope prof(PC_LINEAR, s, 2.0 * a.M * a.N * a.K, 2.0 * a.N * a.K + 4.0 * a.M * a.K + 4.0 * a.M * a.N);
    SCV_CUDA(launch_k(linear_rowwarp_kernel, dim3(ceil_div(a.M, 8)), dim3(256), 0, s, a));
    SCV_LAUNCH_CHECK();
    return 0;
  }
  SCV_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "linear: empty shape M=%d N=%d K=%d", a.M, a.N, a.K);
  SCV_REQUIRE(a.ldw % 8 == 0 && a.ldw >= a.K, "linear: ldw=%d must be a multiple of 8 and >= K=%d", a.ldw, a.K);
  SCV_REQUIRE((reinterpret_cast<uintptr_t>(a.w) & 15u) == 0, "linear: weight pointer must be 16-byte aligned");
  const int vec_ok = (a.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0);
  ProfScope prof(PC_LINEAR, s, 2.0 * a.M * a.N * a.K,
                 2.0 * a.N * a.K + 4.0 * a.M * a.K + 4.0 * a.M * a.N * (a.residual ? 2 : 1));
  if (a.M > 48) {
    dim3 grid(ceil_div(a.N, 128), ceil_div(a.M, 128));
    SCV_CUDA(launch_k(linear_simt_kernel<128, 128, 16, 8, 8>, grid, dim3(256), 0, s, a, vec_ok));
  } else {
    dim3 grid(ceil_div(a.N, 32), ceil_div(a.M, 32));
    SCV_CUDA(launch_k(linear_simt_kernel<32, 32, 16, 2, 2>, grid, dim3(256), 0, s, a, vec_ok));
  }
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
