// Small bandwidth-bound kernels: weight packing, LayerNorm(+activation), row utilities.
#include "common.cuh"

namespace scv {

// fp32 [rows, cols] -> bf16 [rows, ld_dst] (round-to-nearest-even), zero padding columns >= cols.
__global__ void pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows,
                                 int cols, int ld_dst) {
  const int64_t total = (int64_t)rows * ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_dst), c = (int)(i % ld_dst);
    dst[i] = __float2bfloat16_rn(c < cols ? src[(int64_t)r * cols + c] : 0.f);
  }
}

int launch_pack_bf16(const float* src, __nv_bfloat16* dst, int rows, int cols, int ld_dst, cudaStream_t s) {
  const int64_t total = (int64_t)rows * ld_dst;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 16);
  pack_bf16_kernel<<<blocks, 256, 0, s>>>(src, dst, rows, cols, ld_dst);
  SCV_LAUNCH_CHECK();
  return 0;
}

int launch_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t s) {
  SCV_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

// One warp per row.  Matches nn.LayerNorm(eps=1e-5): biased variance, two-pass in fp32.
// (norm1/2/3 of nn.TransformerDecoderLayer, output_proj.0, token_type_head.0, *_to_memory.1 ...)
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* x, int ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* y, int ldy, int M, int N, int act,
                 const int* done_flag) {
  pdl_wait();
  if (done_flag != nullptr && *done_flag != 0) return;
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + (size_t)warp * ldx;
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += xr[i];
  const float mean = warp_sum(s) / (float)N;
  float v = 0.f;
  for (int i = lane; i < N; i += 32) {
    const float d = xr[i] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)N + 1e-5f);
  float* yr = y + (size_t)warp * ldy;
  for (int i = lane; i < N; i += 32) {
    const float o = (xr[i] - mean) * rstd * gamma[i] + beta[i];
    yr[i] = apply_act(o, act);
  }
}

// LayerNorm whose only consumer is a tensor-core projection: one warp per row, the row held in registers
// (single global read), output written as bf16 hi/lo chunks straight into the SplitTile layout of
// gemm_tcgen05.cu ([m_tile][k_block][hi 16 KB | lo 16 KB], 128-byte swizzle).  normalize = 0 only splits.
constexpr int kLnMaxChunks = 4;      // 8-element chunks per lane -> rows up to 1024 wide
__global__ void __launch_bounds__(256)
layernorm_split_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                       const float* __restrict__ beta, uint8_t* __restrict__ out, int M, int N, int normalize,
                       int act, const int* done_flag) {
  pdl_wait();
  if (done_flag != nullptr && *done_flag != 0) return;
  if (!normalize) pdl_launch_dependents();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int n_chunks = N >> 3, KB = (N + 63) >> 6;
  const float* xr = x + (size_t)row * ldx;
  if (!normalize) {
    // plain split of a row of any width (latents, memory tokens): stream the chunks, no row statistics needed
    const int mt = row >> 7, ri = row & 127;
    for (int ch = lane; ch < KB * 8; ch += 32) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
      if (ch < n_chunks) {
        const float4 a = *reinterpret_cast<const float4*>(xr + ch * 8), b = *reinterpret_cast<const float4*>(xr + ch * 8 + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) split_pair(o[2 * p], o[2 * p + 1], hi[p], lo[p]);
      const int kb = ch >> 3, cj = ch & 7;
      uint8_t* dst = out + ((size_t)mt * KB + kb) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    return;
  }
  float v[kLnMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kLnMaxChunks; ++j) {
    const int ch = lane + 32 * j;
    if (ch < n_chunks) {
      const float4 a = *reinterpret_cast<const float4*>(xr + ch * 8);
      const float4 b = *reinterpret_cast<const float4*>(xr + ch * 8 + 4);
      v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
      v[j][4] = b.x; v[j][5] = b.y; v[j][6] = b.z; v[j][7] = b.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[j][e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (normalize) {
    mean = warp_sum(s) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxChunks; ++j) {
      if (lane + 32 * j < n_chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[j][e] - mean; q = fmaf(d, d, q); }
      }
    }
    rstd = 1.0f / sqrtf(warp_sum(q) / (float)N + 1e-5f);
  }
  if (threadIdx.x == 0) pdl_launch_dependents();   // the row is in registers: only the stores are left
  const int mt = row >> 7, ri = row & 127;
#pragma unroll
  for (int j = 0; j < kLnMaxChunks; ++j) {
    const int ch = lane + 32 * j;
    if (ch < KB * 8) {                      // chunks in [n_chunks, KB*8) are the zero padding of the last k-block
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
      if (ch < n_chunks) {
        if (normalize) {
          const float4 g0 = *reinterpret_cast<const float4*>(gamma + ch * 8), g1 = *reinterpret_cast<const float4*>(gamma + ch * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(beta + ch * 8), b1 = *reinterpret_cast<const float4*>(beta + ch * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = apply_act((v[j][e] - mean) * rstd * gg[e] + bb[e], act);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = v[j][e];
        }
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) split_pair(o[2 * p], o[2 * p + 1], hi[p], lo[p]);
      const int kb = ch >> 3, cj = ch & 7;
      uint8_t* dst = out + ((size_t)mt * KB + kb) * 32768 + (size_t)ri * 128 + (size_t)((cj ^ (ri & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

int launch_layernorm_split(const float* x, int ldx, const float* gamma, const float* beta, void* out_split, int M,
                           int N, int normalize, const int* done_flag, cudaStream_t s, int act) {
  SCV_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && (!normalize || N <= kLnMaxChunks * 256),
              "layernorm_split: N=%d must be a multiple of 8 (and <= %d when normalising)", N, kLnMaxChunks * 256);
  SCV_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0, "layernorm_split: unaligned input");
  ProfScope prof(PC_LAYERNORM, s, 8.0 * M * N, 8.0 * M * N);
  SCV_CUDA(launch_k(layernorm_split_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, s, x, ldx, gamma, beta,
                    static_cast<uint8_t*>(out_split), M, N, normalize, act, done_flag));
  SCV_LAUNCH_CHECK();
  return 0;
}

int launch_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy, int M,
                     int N, int act, const int* done_flag, cudaStream_t s) {
  SCV_REQUIRE(M > 0 && N > 0, "layernorm: empty shape");
  ProfScope prof(PC_LAYERNORM, s, 8.0 * M * N, 8.0 * M * N);
  SCV_CUDA(launch_k(layernorm_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, s, x, ldx, gamma, beta, y, ldy, M, N, act, done_flag));
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
