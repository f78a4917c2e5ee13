// PTX wrappers and constants shared by the tcgen05 projection kernels (sm_100a).
#pragma once
#include "common.cuh"

namespace scv {
namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int TILE_BYTES = 128 * 128;                       // 128 rows x 64 bf16 = 16 KB
__host__ __device__ constexpr int stage_bytes(int bn) { return 2 * TILE_BYTES + bn * 128; }       // A_hi, A_lo, W (bn rows x 128 B)
__host__ __device__ constexpr int smem_bytes(int stages, int bn) { return stages * stage_bytes(bn) + 1024 /*align*/ + 128 /*barriers*/; }
constexpr int STG_PITCH = 36;                               // floats per staged epilogue row (32 + pad, 16-byte aligned)
constexpr int GROUP_THREADS = 128;
constexpr int NUM_THREADS = 288;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// ... delivered to the same CTA-relative address (and signalled on the same CTA-relative barrier) in every CTA of cta_mask
__device__ __forceinline__ void bulk_copy_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle: 8-row atoms of 1024 B (SBO = 1024), LBO unused, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = bn.
// (A and B formats: 0 = f16, 1 = bf16; common.cuh kSplitFp16)
constexpr uint32_t kIdescFormats = kSplitFp16 ? 0u : ((1u << 7) | (1u << 10));
__host__ __device__ constexpr uint32_t idesc_for(int bn) {
  return (1u << 4) | kIdescFormats | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t kIdesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ... arriving on the barrier at this CTA-relative address in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

struct TcArgs {
  const float* x; int ldx;          // fp32 activations (A_SPLIT = false)
  const uint8_t* a_split;           // SplitTile activations (A_SPLIT = true)
  const __nv_bfloat16* wt;          // tiled + swizzled weights
  int kblocks;                      // ceil(K / 64)
  const float* bias;
  const float* residual; int ldr;
  float* y; int ldy;                // fp32 output (OUT_SPLIT = false)
  uint8_t* y_split; int kb_out;     // SplitTile output with kb_out = ceil(N / 64) k-blocks per row tile
  int M, N, K, act;
  const int* done_flag;
  const uint8_t* next_w; uint32_t next_w_bytes;   // L2 prefetch of the next projection's weights (0 = none)
  const float* ln_gamma; const float* ln_beta; uint8_t* ln_out;   // gemm_tcgen05_ln.cu: LayerNorm of the output row
  void* trace;                                                    // CTA residency trace (common.cuh), normally null
  int row_base;                                                   // first row of the launch inside the call's batch
  int* nonfinite_flag; int nonfinite_mode;                        // LinearArgs::nonfinite_flag (fp32 output only)
};


}  // namespace tc
}  // namespace scv
