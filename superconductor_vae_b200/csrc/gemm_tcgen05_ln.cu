// x_new = a * W^T + bias + x, and LayerNorm(x_new) as the next projection's SplitTile, in ONE kernel (sm_100a).
//
// The three residual projections of a decoder layer (self-attention out_proj, cross-attention out_proj, linear2;
// models/autoregressive_decoder.py:1293-1296, 1307-1308, 1312-1313) are each followed by a LayerNorm of the row they
// just wrote (norm2, norm3, the next layer's norm1 / the head LayerNorms, :1299, :1311, :1244).  Their N = d_model
// columns span four 128-wide tiles, so the four CTAs of a row block form a thread-block cluster: each computes its
// tile exactly like gemm_tcgen05.cu, keeps the 128 x 128 result in registers, and the row statistics are exchanged
// through distributed shared memory -- two passes (sum -> mean, then sum of squared deviations -> variance), the same
// arithmetic as layernorm_split_kernel (misc_kernels.cu), only the order of the partial sums differs.  This removes
// 36 of the 38 LayerNorm launches of a decode step and the re-read of x.
//
// Main loop = gemm_tcgen05.cu with A_SPLIT, BN = 128: one producer thread (bulk copies), one MMA thread (two UMMAs,
// hi and lo, per weight tile), fp32 accumulator in TMEM.
#include <cstdlib>

#include "tcgen05_common.cuh"

namespace scv {

using namespace tc;

namespace {

constexpr int LN_CLUSTER = 4;                                     // column tiles per row block: N = 4 * 128
constexpr int STATS_BYTES = 2 * LN_CLUSTER * 2 * BM * 4;          // [pass][source rank][column half][row] floats

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_remote_f32(uint32_t local_saddr, uint32_t rank, float v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, STAGES == 2 ? 2 : 1) gemm_res_ln_kernel(TcArgs a) {
  constexpr int BN = 128;
  constexpr int STAGE_BYTES = stage_bytes(BN);
  constexpr uint32_t TMEM_COLS = BN;
  constexpr uint32_t kIdesc = idesc_for(BN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;            // full[STAGES], empty[STAGES], accum, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t accum_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));
  const uint32_t stats = bars + 128u;                           // never aliased by the pipeline stages: peers write here
  const float* stats_ptr = reinterpret_cast<const float*>(base_ptr + STAGES * STAGE_BYTES + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, m0 = m_tile * BM, n0 = n_tile * BN;
  const int KB = a.kblocks;
  const uint32_t rank = cluster_rank();

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  const bool skip = a.done_flag != nullptr && *a.done_flag != 0;     // uniform over the grid: the whole cluster skips
  // cluster barrier #1: every CTA of the cluster is running (its shared memory exists) before any peer writes to it
  __syncwarp();
  if (!skip) cluster_arrive();
  const __nv_bfloat16* wtile0 = a.wt + (size_t)n_tile * KB * (TILE_BYTES / 2);

  if (skip) {
    // fall through to the teardown
  } else if (warp < 8) {
    if (tid == 0) {
      const uint8_t* atile0 = a.a_split + (size_t)m_tile * KB * (2 * TILE_BYTES);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(empty_bar(s), phase ^ 1u);
        const uint32_t st_base = base + s * STAGE_BYTES;
        mbar_arrive_expect_tx(full_bar(s), 3 * TILE_BYTES);
        bulk_copy_g2s(st_base, atile0 + (size_t)kb * (2 * TILE_BYTES), 2 * TILE_BYTES, full_bar(s));
        bulk_copy_g2s(st_base + 2 * TILE_BYTES, wtile0 + (size_t)kb * (TILE_BYTES / 2), TILE_BYTES, full_bar(s));
      }
    }
    // ===================== epilogue =====================
    const int quad = warp & 3, chalf = warp >> 2;
    const int c4 = lane & 7, rr = lane >> 3;
    float* stg = reinterpret_cast<float*>(base_ptr) + warp * (32 * STG_PITCH);
    float4 o[2][8];                                  // this thread's 2 x 8 x 4 outputs; first holds the residual
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int gn = n0 + chalf * 64 + cc * 32 + 4 * c4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int gm = m0 + quad * 32 + it * 4 + rr;
        o[cc][it] = gm < a.M ? *reinterpret_cast<const float4*>(a.residual + (size_t)gm * a.ldr + gn)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (tid == 0) pdl_launch_dependents();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = chalf * 64 + cc * 32;
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      const int gn = n0 + c0 + 4 * c4;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.bias != nullptr) bv = *reinterpret_cast<const float4*>(a.bias + gn);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rr, gm = m0 + quad * 32 + row;
        const float4 v = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 4 * c4);
        float4 x;
        x.x = (v.x + bv.x) + o[cc][it].x; x.y = (v.y + bv.y) + o[cc][it].y;
        x.z = (v.z + bv.z) + o[cc][it].z; x.w = (v.w + bv.w) + o[cc][it].w;
        o[cc][it] = x;
        if (gm < a.M) *reinterpret_cast<float4*>(a.y + (size_t)gm * a.ldy + gn) = x;
      }
    }
    // ---- row statistics across the 4 CTAs x 2 column halves that share a row
    // slot(pass, source rank, column half, row); lanes c4 = 0..3 of each 8-lane group deliver to cluster rank c4
    auto slot = [&](int pass, uint32_t src, int half, int row) -> uint32_t {
      return stats + 4u * (uint32_t)(((pass * LN_CLUSTER + (int)src) * 2 + half) * BM + row);
    };
    auto gather = [&](int pass, int row) -> float {
      float t = 0.f;
#pragma unroll
      for (int s = 0; s < LN_CLUSTER; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) t += stats_ptr[((pass * LN_CLUSTER + s) * 2 + h) * BM + row];
      return t;
    };
    const float inv_n = 1.0f / (float)a.N;
    __syncwarp();
    cluster_wait();                                   // barrier #1 (see above)
    float mean[8], rstd[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      float s = (o[0][it].x + o[0][it].y + o[0][it].z + o[0][it].w) + (o[1][it].x + o[1][it].y + o[1][it].z + o[1][it].w);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (c4 < LN_CLUSTER) st_remote_f32(slot(0, rank, chalf, quad * 32 + it * 4 + rr), (uint32_t)c4, s);
    }
    __syncwarp();
    cluster_arrive();                                 // barrier #2: every partial sum has been delivered
    cluster_wait();
#pragma unroll
    for (int it = 0; it < 8; ++it) mean[it] = gather(0, quad * 32 + it * 4 + rr) * inv_n;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      float q = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const float d0 = o[cc][it].x - mean[it], d1 = o[cc][it].y - mean[it];
        const float d2 = o[cc][it].z - mean[it], d3 = o[cc][it].w - mean[it];
        q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
      }
      q += __shfl_xor_sync(0xffffffffu, q, 1);
      q += __shfl_xor_sync(0xffffffffu, q, 2);
      q += __shfl_xor_sync(0xffffffffu, q, 4);
      if (c4 < LN_CLUSTER) st_remote_f32(slot(1, rank, chalf, quad * 32 + it * 4 + rr), (uint32_t)c4, q);
    }
    __syncwarp();
    cluster_arrive();                                 // barrier #3
    cluster_wait();
#pragma unroll
    for (int it = 0; it < 8; ++it) rstd[it] = 1.0f / sqrtf(gather(1, quad * 32 + it * 4 + rr) * inv_n + 1e-5f);
    // ---- normalise, scale, shift, split to bf16 hi/lo, store as the next projection's A tile
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = n0 + chalf * 64 + cc * 32 + 4 * c4;
      const float4 g = *reinterpret_cast<const float4*>(a.ln_gamma + col);
      const float4 bt = *reinterpret_cast<const float4*>(a.ln_beta + col);
      const int kb2 = col >> 6, cj = (col & 63) >> 3;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int ri = quad * 32 + it * 4 + rr;
        if (m0 + ri < a.M) {
          const float y0 = (o[cc][it].x - mean[it]) * rstd[it] * g.x + bt.x;
          const float y1 = (o[cc][it].y - mean[it]) * rstd[it] * g.y + bt.y;
          const float y2 = (o[cc][it].z - mean[it]) * rstd[it] * g.z + bt.z;
          const float y3 = (o[cc][it].w - mean[it]) * rstd[it] * g.w + bt.w;
          uint32_t hi0, lo0, hi1, lo1;
          split_pair(y0, y1, hi0, lo0);
          split_pair(y2, y3, hi1, lo1);
          uint8_t* dst = a.ln_out + ((size_t)m_tile * a.kb_out + kb2) * (2 * TILE_BYTES) + (size_t)ri * 128 +
                         (size_t)((cj ^ (ri & 7)) << 4) + (size_t)((col & 7) >> 2) * 8;
          *reinterpret_cast<uint2*>(dst) = make_uint2(hi0, hi1);
          *reinterpret_cast<uint2*>(dst + TILE_BYTES) = make_uint2(lo0, lo1);
        }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // ===================== MMA issuer (warp 8) =====================
    if (lane == 0) {
      if (a.next_w_bytes != 0) {
        const uint32_t ncta = gridDim.x * gridDim.y, cid = blockIdx.y * gridDim.x + blockIdx.x;
        const uint32_t chunk = ((a.next_w_bytes + ncta - 1) / ncta + 127u) & ~127u;
        const uint32_t off = cid * chunk;
        if (off < a.next_w_bytes) bulk_prefetch_l2(a.next_w + off, min(chunk, a.next_w_bytes - off) & ~15u);
      }
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(full_bar(s), phase);
        tc_fence_after();
        const uint32_t st_base = base + s * STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t bd = umma_desc_sw128(st_base + 2 * TILE_BYTES + kk * 32);
          umma_bf16(tmem_base, umma_desc_sw128(st_base + kk * 32), bd, (kb | kk) != 0 ? 1u : 0u, kIdesc);
          umma_bf16(tmem_base, umma_desc_sw128(st_base + TILE_BYTES + kk * 32), bd, 1u, kIdesc);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
    // the cluster barriers count every thread of every CTA: this warp takes part in all three
    cluster_wait();
    cluster_arrive();
    cluster_wait();
    cluster_arrive();
    cluster_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

bool tc_res_ln_ok(const LinearArgs& a) {
  // Opt-in (SCV_FUSE_LN=1).  Measured on B200 at 4096 rows: 18.4 us against 13.3 us (projection) + 5.8 us (LayerNorm
  // kernel) in isolation, and 2.4 % SLOWER per decode with two sub-batch streams -- the cluster barriers wait for the
  // slowest of the four CTAs and a 4-CTA cluster needs four free SMs of one GPC at once.
  static const int env = [] { const char* e = getenv("SCV_FUSE_LN"); return e ? atoi(e) : 0; }();
  return env != 0 && a.ln_out_split != nullptr && a.ln_gamma != nullptr && a.ln_beta != nullptr && a.a_split != nullptr &&
         a.y_split == nullptr && a.residual != nullptr && a.act == ACT_NONE && a.N == LN_CLUSTER * 128 && a.K % 64 == 0 &&
         a.wt != nullptr && a.y != nullptr && a.ldy % 4 == 0 && a.ldr % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(a.y) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15u) == 0 &&
         (a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15u) == 0) &&
         (reinterpret_cast<uintptr_t>(a.ln_gamma) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.ln_beta) & 15u) == 0;
}

int launch_linear_res_ln(const LinearArgs& a, cudaStream_t s) {
  SCV_REQUIRE(tc_res_ln_ok(a), "fused residual + LayerNorm projection: unsupported shape (M=%d N=%d K=%d)", a.M, a.N, a.K);
  constexpr int SMEM2 = smem_bytes(2, 128) + STATS_BYTES, SMEM4 = smem_bytes(4, 128) + STATS_BYTES;
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
    SCV_CUDA(cudaFuncSetAttribute(gemm_res_ln_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
    SCV_CUDA(cudaFuncSetAttribute(gemm_res_ln_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM4));
  }
  TcArgs t = {};
  t.a_split = reinterpret_cast<const uint8_t*>(a.a_split); t.wt = a.wt; t.kblocks = ceil_div(a.K, BK);
  t.bias = a.bias; t.residual = a.residual; t.ldr = a.ldr; t.y = a.y; t.ldy = a.ldy;
  t.kb_out = ceil_div(a.N, BK); t.M = a.M; t.N = a.N; t.K = a.K; t.act = a.act; t.done_flag = a.done_flag;
  t.next_w = static_cast<const uint8_t*>(a.next_w); t.next_w_bytes = (uint32_t)a.next_w_bytes;
  t.ln_gamma = a.ln_gamma; t.ln_beta = a.ln_beta; t.ln_out = reinterpret_cast<uint8_t*>(a.ln_out_split);
  // bytes: the plain projection's + the SplitTile written instead of a separate LayerNorm pass
  ProfScope prof(PC_GEMM_TC, s, 2.0 * a.M * a.N * a.K, 2.0 * a.N * a.K + 4.0 * a.M * a.K + 12.0 * a.M * a.N);
  const int mt = ceil_div(a.M, BM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LN_CLUSTER, mt); cfg.blockDim = dim3(NUM_THREADS); cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LN_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (LN_CLUSTER * mt <= 148) {
    cfg.dynamicSmemBytes = SMEM4;
    SCV_CUDA(cudaLaunchKernelEx(&cfg, gemm_res_ln_kernel<4>, t));
  } else {
    cfg.dynamicSmemBytes = SMEM2;
    SCV_CUDA(cudaLaunchKernelEx(&cfg, gemm_res_ln_kernel<2>, t));
  }
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
