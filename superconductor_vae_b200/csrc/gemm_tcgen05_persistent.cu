// Persistent variant of the tcgen05 projection (SplitTile activations only) for grids of more than one tile per SM.
//
// Why: the non-persistent kernel gets its epilogue/main-loop overlap from two CTAs per SM, but CTAs that start
// together run in lock-step (both in the main loop, tensor pipe shared; then both in the epilogue, tensor pipe
// idle), so a QKV projection of 4096 rows measured 19.6 us against ~8 us of tensor work, FFN1 31.5 us against ~10.
// Here ONE CTA per SM walks a static list of 128 x 128 output tiles with the three roles decoupled:
//   warp 9  lane 0  producer: bulk copies (A hi|lo 32 KB + W 16 KB per k-block) into a 3-stage ring that keeps
//                   running across tile boundaries;
//   warp 8  lane 0  MMA issuer: accumulates tile i into TMEM buffer i & 1 (2 x 128 columns) as soon as the epilogue
//                   has drained that buffer;
//   warps 0-7       epilogue of tile i-1 (tcgen05.ld, bias / activation / residual, staged through its own shared
//                   memory for row-contiguous stores) while tile i is in the tensor core.
#include <cstdlib>

#include "tcgen05_common.cuh"

namespace scv {

using namespace tc;

namespace {

constexpr int P_STAGES = 3;
constexpr int P_BN = 128;
constexpr int P_STAGE_BYTES = 3 * TILE_BYTES;
constexpr int P_STG_BYTES = 8 * 32 * STG_PITCH * 4;                 // epilogue staging, 8 warps x 32 rows
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + P_STG_BYTES + 1024 /*align*/ + 128 /*barriers*/;
constexpr int P_THREADS = 320;
constexpr uint32_t P_TMEM_COLS = 256;

template <bool OUT_SPLIT, bool HAS_RES>
__global__ void __launch_bounds__(P_THREADS, 1) gemm_tcgen05_persistent_kernel(TcArgs a, int n_tiles_n, int n_tiles) {
  constexpr uint32_t kIdesc = idesc_for(P_BN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t stg_base = P_STAGES * P_STAGE_BYTES;
  const uint32_t bars = base + stg_base + P_STG_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (P_STAGES + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * P_STAGES + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * P_STAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * P_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + stg_base + P_STG_BYTES + 8 * (2 * P_STAGES + 4));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = a.kblocks;

  if (tid == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(P_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  const bool skip = a.done_flag != nullptr && *a.done_flag != 0;
  pdl_launch_dependents();

  if (skip) {
    // teardown only
  } else if (warp == 9) {
    // ===================== producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m_tile = tile / n_tiles_n, n_tile = tile % n_tiles_n;
        const uint8_t* atile0 = a.a_split + (size_t)m_tile * KB * (2 * TILE_BYTES);
        const __nv_bfloat16* wtile0 = a.wt + (size_t)n_tile * KB * (TILE_BYTES / 2);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % P_STAGES;
          const uint32_t phase = (it / P_STAGES) & 1u;
          mbar_wait(empty_bar(s), phase ^ 1u);
          const uint32_t st_base = base + s * P_STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(s), 3 * TILE_BYTES);
          bulk_copy_g2s(st_base, atile0 + (size_t)kb * (2 * TILE_BYTES), 2 * TILE_BYTES, full_bar(s));
          bulk_copy_g2s(st_base + 2 * TILE_BYTES, wtile0 + (size_t)kb * (TILE_BYTES / 2), TILE_BYTES, full_bar(s));
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1u, use = ti >> 1;              // use-th time this accumulator buffer is filled
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);              // epilogue has drained it (first use passes)
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * P_BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % P_STAGES;
          const uint32_t phase = (it / P_STAGES) & 1u;
          mbar_wait(full_bar(s), phase);
          tc_fence_after();
          const uint32_t st_base = base + s * P_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t bd = umma_desc_sw128(st_base + 2 * TILE_BYTES + kk * 32);
            umma_bf16(acc, umma_desc_sw128(st_base + kk * 32), bd, (kb | kk) != 0 ? 1u : 0u, kIdesc);
            umma_bf16(acc, umma_desc_sw128(st_base + TILE_BYTES + kk * 32), bd, 1u, kIdesc);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(buf));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps 0-7 =====================
    const int quad = warp & 3, chalf = warp >> 2;
    float* stg = reinterpret_cast<float*>(base_ptr + stg_base) + warp * (32 * STG_PITCH);
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int m_tile = tile / n_tiles_n, n_tile = tile % n_tiles_n;
      const int m0 = m_tile * BM, n0 = n_tile * P_BN;
      const uint32_t buf = ti & 1u, use = ti >> 1;
      float4 rv[HAS_RES ? 2 : 1][HAS_RES ? 8 : 1];
      if constexpr (HAS_RES) {      // residual is an input: fetch it while the tile is still in the tensor core
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int gn = n0 + chalf * 64 + cc * 32 + 4 * (lane & 7);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int gm = m0 + quad * 32 + i * 4 + (lane >> 3);
            rv[cc][i] = (gm < a.M && gn < a.N) ? *reinterpret_cast<const float4*>(a.residual + (size_t)gm * a.ldr + gn)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      mbar_wait(tfull_bar(buf), use & 1u);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = chalf * 64 + cc * 32;
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * P_BN + (uint32_t)c0, r);
        if (cc == 1) {              // both halves of this warp's columns are in registers: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int gn0 = n0 + c0;
        if constexpr (!OUT_SPLIT) {
          const int c4 = lane & 7, rr = lane >> 3;
          const int gn = gn0 + 4 * c4;
          if (gn < a.N) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.bias != nullptr) bv = *reinterpret_cast<const float4*>(a.bias + gn);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rr, gm = m0 + quad * 32 + row;
              if (gm < a.M) {
                const float4 v = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 4 * c4);
                float o0 = apply_act(v.x + bv.x, a.act), o1 = apply_act(v.y + bv.y, a.act);
                float o2 = apply_act(v.z + bv.z, a.act), o3 = apply_act(v.w + bv.w, a.act);
                if constexpr (HAS_RES) { o0 += rv[cc][i].x; o1 += rv[cc][i].y; o2 += rv[cc][i].z; o3 += rv[cc][i].w; }
                *reinterpret_cast<float4*>(a.y + (size_t)gm * a.ldy + gn) = make_float4(o0, o1, o2, o3);
              }
            }
          }
        } else {
          const int kb2 = gn0 >> 6, chunk0 = (gn0 & 63) >> 3;
          const int ch = lane & 3, rr = lane >> 2;
          const int gn = gn0 + 8 * ch;
          if (kb2 < a.kb_out) {
            float bb[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) bb[e] = (a.bias != nullptr && gn + e < a.N) ? a.bias[gn + e] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = i * 8 + rr, ri = quad * 32 + row, gm = m0 + ri;
              if (gm < a.M) {
                const float4 v0 = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 8 * ch);
                const float4 v1 = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 8 * ch + 4);
                const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int p2 = 0; p2 < 4; ++p2) {
                  const float o0 = gn + 2 * p2 < a.N ? apply_act(vv[2 * p2] + bb[2 * p2], a.act) : 0.f;
                  const float o1 = gn + 2 * p2 + 1 < a.N ? apply_act(vv[2 * p2 + 1] + bb[2 * p2 + 1], a.act) : 0.f;
                  split_pair(o0, o1, hi[p2], lo[p2]);
                }
                uint8_t* dst = a.y_split + ((size_t)m_tile * a.kb_out + kb2) * (2 * TILE_BYTES) + (size_t)ri * 128 +
                               (size_t)(((chunk0 + ch) ^ (ri & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(dst + TILE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P_TMEM_COLS) : "memory");
  }
}

}  // namespace

// SplitTile input, more than one 128 x 128 tile per SM: the persistent kernel pays off
bool tc_persistent_ok(const LinearArgs& a) {
  // Opt-in: in isolation it is faster for multi-wave grids (M = 4096: QKV 19.6 -> 15.8 us, logits 48.6 -> 41.7 us),
  // inside the decode step it measured 1.5-5 % slower than two 2-stage CTAs per SM (profiles/README.md, r01s).
  static const int env = [] { const char* e = getenv("SCV_GEMM_PERSISTENT"); return e ? atoi(e) : 0; }();
  return env != 0 && a.a_split != nullptr && (long long)ceil_div(a.N, P_BN) * ceil_div(a.M, BM) > 148;
}

int launch_linear_tcgen05_persistent(const LinearArgs& a, cudaStream_t s) {
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
  }
  TcArgs t;
  t.x = a.x; t.ldx = a.ldx; t.a_split = reinterpret_cast<const uint8_t*>(a.a_split); t.wt = a.wt;
  t.kblocks = ceil_div(a.K, BK); t.bias = a.bias; t.residual = a.residual; t.ldr = a.ldr; t.y = a.y; t.ldy = a.ldy;
  t.y_split = reinterpret_cast<uint8_t*>(a.y_split); t.kb_out = ceil_div(a.N, BK);
  t.M = a.M; t.N = a.N; t.K = a.K; t.act = a.act; t.done_flag = a.done_flag;
  t.next_w = nullptr; t.next_w_bytes = 0;
  const int ntn = ceil_div(a.N, P_BN), ntm = ceil_div(a.M, BM), n_tiles = ntn * ntm;
  const int grid = std::min(n_tiles, 148);
  ProfScope prof(PC_GEMM_TC, s, 2.0 * a.M * a.N * a.K,
                 2.0 * a.N * a.K + 4.0 * a.M * a.K + 4.0 * a.M * a.N * (a.residual ? 2 : 1));
  if (a.y_split != nullptr)
    SCV_CUDA(launch_k(gemm_tcgen05_persistent_kernel<true, false>, dim3(grid), dim3(P_THREADS), (size_t)P_SMEM_BYTES, s, t, ntn, n_tiles));
  else if (a.residual != nullptr)
    SCV_CUDA(launch_k(gemm_tcgen05_persistent_kernel<false, true>, dim3(grid), dim3(P_THREADS), (size_t)P_SMEM_BYTES, s, t, ntn, n_tiles));
  else
    SCV_CUDA(launch_k(gemm_tcgen05_persistent_kernel<false, false>, dim3(grid), dim3(P_THREADS), (size_t)P_SMEM_BYTES, s, t, ntn, n_tiles));
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
