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
  if (done_flag != nullptr && *done_flag != 0) return;
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

int launch_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy, int M,
                     int N, int act, const int* done_flag, cudaStream_t s) {
  SCV_REQUIRE(M > 0 && N > 0, "layernorm: empty shape");
  ProfScope prof(PC_LAYERNORM, s, 8.0 * M * N, 8.0 * M * N);
  layernorm_kernel<<<ceil_div(M, 8), 256, 0, s>>>(x, ldx, gamma, beta, y, ldy, M, N, act, done_flag);
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
