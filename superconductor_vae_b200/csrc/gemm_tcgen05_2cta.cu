// Wide projections (N a multiple of 256, SplitTile input) on CTA pairs: tcgen05.mma.cta_group::2, one 256 x 256 output
// tile per pair of SMs.
//
// Why: the 128 x 128 kernel (gemm_tcgen05.cu) is bound by operand delivery from L2 into the SMs, not by the tensor pipe
// (QKV moves 144 MB in 21 us): per 64-wide k-block a CTA pulls 32 KB of A (hi | lo) and 16 KB of W
// for 128 x 128 outputs.  In a pair, CTA r holds rows [128 r, 128 r + 128) of the A tile and HALF of the 256 weight
// rows; the UMMA of the leader reads both halves from both shared memories and accumulates 128 x 256 outputs in each
// CTA's TMEM, so the same 48 KB per CTA and k-block now feed twice the outputs (the A re-reads of the grid halve) while
// two 2-stage CTAs per SM still overlap one CTA's epilogue with the other's main loop (the single-CTA 128 x 256 tile
// needs 64 KB stages = one CTA per SM and was slower, profiles/README.md r01e).
//
// Pipeline per CTA: thread 0 streams its A tile and its weight half with cp.async.bulk into a 2-stage ring (full /
// empty mbarriers); bulk copies can only signal a barrier of the CTA they write to, so in the follower CTA a second
// thread waits for each stage and forwards one arrive to the leader's `peer_full` barrier (mapa + remote arrive).  The
// leader's MMA thread waits for both, issues the 8 UMMAs of the k-block (A hi and A lo against the same weight
// fragment) and commits to the `empty` barriers of BOTH CTAs (multicast commit), at the end to both accumulator
// barriers.  Epilogue as in gemm_tcgen05.cu (TMEM -> registers -> shared staging -> row-contiguous stores), 256 columns.
//
// STATUS: opt-in (SCV_GEMM_2CTA=k: projections at least 256 k columns wide; covered by
// test_optin_kernels_match_default_tokens).  Results are identical to the single-CTA kernel (max |err| 1.5e-5 against
// fp64 on the QKV / logits shapes) but it is not faster, which also corrects the diagnosis above: halving the A traffic
// buys nothing.  Measured on B200 (tests/gemm_bench.py, tests/gemm2cta_sweep.sh): 4096 rows, isolated: QKV 20.0 us
// (single-CTA 20.2), FFN1 27.5 (28.2), logits 54.3 (48.3); 2048 rows: QKV 14.5 (12.1), FFN1 20.7 (15.1), logits 32.2
// (26.9); whole decode 55.9 K formulas/s with k = 3 against 57.5 K; a 4-stage one-CTA-per-SM ring
// (SCV_GEMM_2CTA_STAGES=4) is slower still.  What bounds these projections is per-tile latency (K = 512 is 8 k-blocks:
// fill, drain and the epilogue weigh as much as the MMAs, and the forwarded `peer_full` arrive adds a hop per k-block)
// and wave quantisation (QKV at 4096 rows = 384 tiles on 296 CTA slots), not operand bandwidth.
#include <cstdlib>

#include <type_traits>

#include "tcgen05_common.cuh"

namespace scv {

using namespace tc;

namespace {

constexpr int BN2 = 256;                                  // output columns of a pair (and of each CTA's accumulator)
constexpr int STAGE2_BYTES = 3 * TILE_BYTES;              // A hi | A lo | this CTA's 128 weight rows
__host__ __device__ constexpr int smem2_bytes(int stages) { return stages * STAGE2_BYTES + 1024 + 128; }
// kind::f16: D = f32, A = B = bf16, K-major, M = 256 (pair), N = 256
constexpr uint32_t kIdesc2 = (1u << 4) | kIdescFormats | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {        // acquire at cluster scope
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc2), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {       // arrives on `bar` of both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <bool OUT_SPLIT, int STAGES2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, STAGES2 == 2 ? 2 : 1) gemm_tcgen05_2cta_kernel(TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SW128 tiles need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES2 * STAGE2_BYTES;          // full[2], empty[2], peer_full[2], accum, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES2 + s); };
  auto peer_full_bar = [&](int s) { return bars + 8u * (2 * STAGES2 + s); };
  const uint32_t accum_bar = bars + 8u * (3 * STAGES2);
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES2 + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES2 * STAGE2_BYTES + 8 * (3 * STAGES2 + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                       // 0 = leader (issues the MMAs)
  const int n_tile = blockIdx.x >> 1, m_tile = blockIdx.y * 2 + (int)rank, m0 = m_tile * BM, n0 = n_tile * BN2;
  const int KB = a.kblocks;

  if (tid == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(peer_full_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN2) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                            // both CTAs' barriers exist before anybody signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  const bool skip = a.done_flag != nullptr && *a.done_flag != 0;
  // weight tiles are packed per 128 output rows: the pair's 256 columns are two of them, CTA r streams tile r
  const __nv_bfloat16* wtile = a.wt + ((size_t)n_tile * 2 + rank) * KB * (TILE_BYTES / 2);

  if (skip) {
    // fall through to the teardown
  } else if (warp < 8) {
    if (tid == 0) {                                              // producer
      const uint8_t* atile0 = a.a_split + (size_t)m_tile * KB * (2 * TILE_BYTES);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES2;
        const uint32_t phase = (uint32_t)(kb / STAGES2) & 1u;
        mbar_wait_cluster(empty_bar(s), phase ^ 1u);             // freed by the leader's multicast commit
        const uint32_t st_base = base + s * STAGE2_BYTES;
        mbar_arrive_expect_tx(full_bar(s), 3 * TILE_BYTES);
        bulk_copy_g2s(st_base, atile0 + (size_t)kb * (2 * TILE_BYTES), 2 * TILE_BYTES, full_bar(s));
        bulk_copy_g2s(st_base + 2 * TILE_BYTES, wtile + (size_t)kb * (TILE_BYTES / 2), TILE_BYTES, full_bar(s));
      }
    } else if (tid == 32 && rank == 1) {                         // follower: tell the leader when a stage has landed here
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES2;
        mbar_wait(full_bar(s), (uint32_t)(kb / STAGES2) & 1u);
        mbar_arrive_remote(map_to_cta(peer_full_bar(s), 0));
      }
    }
    // ===================== epilogue (both CTAs: rows m0 .. m0 + 127, the pair's 256 columns) =====================
    const int quad = warp & 3, chalf = warp >> 2;
    float* stg = reinterpret_cast<float*>(base_ptr) + warp * (32 * STG_PITCH);
    mbar_wait_cluster(accum_bar, 0);
    tc_fence_after();
    if (tid == 0) pdl_launch_dependents();
    auto epilogue = [&](auto act_tag) {
      constexpr int ACT = decltype(act_tag)::value;
#pragma unroll 1
      for (int cc = 0; cc < BN2 / 64; ++cc) {
        const int c0 = chalf * (BN2 / 2) + cc * 32;
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
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
            for (int it = 0; it < 8; ++it) {
              const int row = it * 4 + rr, gm = m0 + quad * 32 + row;
              if (gm < a.M) {
                const float4 v = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 4 * c4);
                *reinterpret_cast<float4*>(a.y + (size_t)gm * a.ldy + gn) =
                    make_float4(apply_act_t<ACT>(v.x + bv.x, a.act), apply_act_t<ACT>(v.y + bv.y, a.act),
                                apply_act_t<ACT>(v.z + bv.z, a.act), apply_act_t<ACT>(v.w + bv.w, a.act));
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
            for (int it = 0; it < 4; ++it) {
              const int row = it * 8 + rr, ri = quad * 32 + row, gm = m0 + ri;
              if (gm < a.M) {
                const float4 v0 = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 8 * ch);
                const float4 v1 = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 8 * ch + 4);
                const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int p2 = 0; p2 < 4; ++p2) {
                  const float o0 = gn + 2 * p2 < a.N ? apply_act_t<ACT>(vv[2 * p2] + bb[2 * p2], a.act) : 0.f;
                  const float o1 = gn + 2 * p2 + 1 < a.N ? apply_act_t<ACT>(vv[2 * p2 + 1] + bb[2 * p2 + 1], a.act) : 0.f;
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
    };
    if (a.act == ACT_NONE) epilogue(std::integral_constant<int, ACT_NONE>{});
    else if (a.act == ACT_GELU) epilogue(std::integral_constant<int, ACT_GELU>{});
    else epilogue(std::integral_constant<int, -1>{});
    __syncwarp();
    tc_fence_before();
  } else {
    // ===================== MMA issuer: warp 8 of the leader =====================
    if (lane == 0 && rank == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES2;
        const uint32_t phase = (uint32_t)(kb / STAGES2) & 1u;
        mbar_wait(full_bar(s), phase);                           // this CTA's operands
        mbar_wait_cluster(peer_full_bar(s), phase);              // the follower's
        tc_fence_after();
        const uint32_t st_base = base + s * STAGE2_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t bd = umma_desc_sw128(st_base + 2 * TILE_BYTES + kk * 32);
          umma2_bf16(tmem_base, umma_desc_sw128(st_base + kk * 32), bd, (kb | kk) != 0 ? 1u : 0u);
          umma2_bf16(tmem_base, umma_desc_sw128(st_base + TILE_BYTES + kk * 32), bd, 1u);
        }
        umma2_commit_both(empty_bar(s));                         // frees the stage in both CTAs
      }
      umma2_commit_both(accum_bar);                              // accumulators of both CTAs complete
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();                                            // nobody leaves while the other CTA may still signal it
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN2) : "memory");
  }
}

}  // namespace

bool tc_2cta_ok(const LinearArgs& a) {
  static const int on = [] { const char* e = getenv("SCV_GEMM_2CTA"); return e ? atoi(e) : 0; }();
  if (!on || a.a_split == nullptr || a.residual != nullptr) return false;
  if (ceil_div(a.N, 128) % 2 != 0 || ceil_div(a.M, BM) % 2 != 0) return false;       // whole 256 x 256 pair tiles
  return ceil_div(a.N, 128) >= on * 2;                           // SCV_GEMM_2CTA=k: projections at least 256 k columns wide
}

int launch_linear_tcgen05_2cta(const LinearArgs& a, cudaStream_t s) {
  // SCV_GEMM_2CTA_STAGES: 2 (default, two CTAs per SM) or 4 (one CTA per SM, deeper ring)
  static const int stages = [] { const char* e = getenv("SCV_GEMM_2CTA_STAGES"); return e && atoi(e) == 4 ? 4 : 2; }();
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes(2)));
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes(2)));
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes(4)));
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes(4)));
  }
  TcArgs t = {};
  t.a_split = reinterpret_cast<const uint8_t*>(a.a_split); t.wt = a.wt;
  t.kblocks = ceil_div(a.K, BK); t.bias = a.bias; t.y = a.y; t.ldy = a.ldy;
  t.y_split = reinterpret_cast<uint8_t*>(a.y_split); t.kb_out = ceil_div(a.N, BK);
  t.M = a.M; t.N = a.N; t.K = a.K; t.act = a.act; t.done_flag = a.done_flag;
  ProfScope prof(PC_GEMM_TC, s, 2.0 * a.M * a.N * a.K, 2.0 * a.N * a.K + 4.0 * a.M * a.K + 4.0 * a.M * a.N);
  const dim3 grid(2 * (ceil_div(a.N, 128) / 2), ceil_div(a.M, BM) / 2);
  const bool os = a.y_split != nullptr;
  if (stages == 4) {
    if (os) SCV_CUDA(launch_k(gemm_tcgen05_2cta_kernel<true, 4>, grid, dim3(NUM_THREADS), (size_t)smem2_bytes(4), s, t));
    else SCV_CUDA(launch_k(gemm_tcgen05_2cta_kernel<false, 4>, grid, dim3(NUM_THREADS), (size_t)smem2_bytes(4), s, t));
  } else {
    if (os) SCV_CUDA(launch_k(gemm_tcgen05_2cta_kernel<true, 2>, grid, dim3(NUM_THREADS), (size_t)smem2_bytes(2), s, t));
    else SCV_CUDA(launch_k(gemm_tcgen05_2cta_kernel<false, 2>, grid, dim3(NUM_THREADS), (size_t)smem2_bytes(2), s, t));
  }
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
