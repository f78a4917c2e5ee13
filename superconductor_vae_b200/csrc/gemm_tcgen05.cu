// y = act(x * W^T + bias) (+ residual) on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Exactness: the decode path must reproduce the fp32 reference's greedy decisions while streaming bf16
// weights (precision-sensitivity table).  Plain bf16 activations flip ~1 % of the
// decisions; splitting every fp32 activation into hi = bf16(x) and lo = bf16(x - hi) and issuing two MMAs per
// weight tile keeps 16 mantissa bits of x, every product bf16 x bf16 is exact in fp32, and the probe in the
// survey shows 100 % agreement.  So per 64-wide k-block this kernel issues 4 x {A_hi, A_lo} x W UMMAs of
// shape 128 x 128 x 16 (kind::f16, bf16 inputs, fp32 accumulate in TMEM).
//
// Operand formats
//   W        packed once at load time: [n_tile][k_block][128 rows x 64 bf16, 128-byte swizzle] (16 KB tiles)
//   A split  "SplitTile": [m_tile][k_block][hi 16 KB | lo 16 KB], same swizzle; written by the producing
//            kernel (LayerNorm, attention, a previous GEMM's epilogue), so one 32 KB + one 16 KB bulk copy
//            (cp.async.bulk -> UBLKCP) per k-block feeds the tensor core with no ALU work in this kernel;
//   A fp32   row-major activations are split inside the kernel by eight producer warps (two groups that
//            alternate k-blocks so one group's global-load latency overlaps the other's conversion work).
// One 128 x 128 output tile per CTA, 2 CTAs per SM (2 stages x 48 KB each) so one CTA's epilogue overlaps the
// other's main loop; 9 warps: 0-7 producers/epilogue (warp w reads TMEM lanes 32*(w%4).., columns 64*(w/4)..),
// warp 8 allocates TMEM and its lane 0 issues the UMMAs.
#include <cstdlib>

#include <type_traits>

#include "tcgen05_common.cuh"

namespace scv {

using namespace tc;

namespace {

// MC (multicast): the launch is made of clusters of two CTAs with neighbouring column tiles of the same row tile.  They need
// the same A tile, so CTA `rank` loads only its half of it (rank 0 the hi tile, rank 1 the lo tile) and the copy is delivered
// to both CTAs' shared memory (cp.async.bulk ... .multicast::cluster, signalling both full barriers); a stage is free when
// the MMAs of BOTH CTAs that read it are done (each commit arrives on both empty barriers).  Per k-block a CTA then pulls
// 32 KB instead of 48 KB out of L2 (ncu: the K = 2048 projection moves 210 MB through the crossbar in 21 us, i.e. it runs at
// the L2's ~10 TB/s).
// NFC: report NaN / +-inf outputs through TcArgs::nonfinite_flag (logits projection of the sampling modes).  A template
// flag, not a run-time test: the run-time test alone, never taken, cost 2.5 % of the 4096-latent decode (68.9 against 67.1-67.6
// ms on one box with alternating builds) by lengthening every fp32 epilogue.
template <bool A_SPLIT, bool OUT_SPLIT, bool HAS_RES, int STAGES, int BN, bool MC = false, bool NFC = false>
__global__ void __launch_bounds__(NUM_THREADS, (STAGES == 2 && BN <= 128) ? 2 : 1) gemm_tcgen05_kernel(TcArgs a) {
  static_assert(!MC || (A_SPLIT && BN == 128), "multicast: SplitTile input, 128-wide tiles");
  constexpr int STAGE_BYTES = stage_bytes(BN);
  constexpr uint32_t TMEM_COLS = BN;
  constexpr uint32_t W_BYTES = BN * 128;
  constexpr uint32_t kIdesc = idesc_for(BN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SW128 tiles need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;            // full[STAGES], empty[STAGES], accum, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t accum_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, m0 = m_tile * BM, n0 = n_tile * BN;
  const int KB = a.kblocks;
  // Compaction of finished rows: this row tile holds no live row.  (StepState::pad[1] was written several kernels ago -
  // by the compaction kernel of the previous step - so it may be read before griddepcontrol.wait.)
  if (a.done_flag != nullptr && a.done_flag[6] > 0 && a.row_base + m0 >= a.done_flag[6]) return;
  TraceRec* trc = tid == 0 ? trace_begin(a.trace, 100u + (uint32_t)(a.N >> 7)) : nullptr;

  uint32_t rank = 0;
  if constexpr (MC) rank = cluster_rank();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), A_SPLIT ? 1 : GROUP_THREADS);
      mbar_init(empty_bar(s), MC ? 2 : 1);
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
  if constexpr (MC) cluster_sync();       // the peer's barriers exist before anything is copied into it or signalled on it
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Everything above (barrier init, TMEM allocation) touched no global memory and overlapped the previous kernel's
  // tail; from here on this kernel reads what its predecessors wrote.
  pdl_wait();
  const bool skip = a.done_flag != nullptr && *a.done_flag != 0;     // every row finished: nothing left to compute
  // weight tiles are packed per 128 output rows: a 256-wide CTA tile is two of them (same k-block, KB tiles apart)
  // (a 64-wide CTA tile is the upper or lower half of one: 64 rows x 128 B are contiguous and keep the swizzle pattern)
  constexpr int NW = BN >= 128 ? BN / 128 : 1;
  constexpr uint32_t W_COPY = BN >= 128 ? TILE_BYTES : W_BYTES;
  const __nv_bfloat16* wtile0 = BN >= 128 ? a.wt + (size_t)n_tile * (BN / 128) * KB * (TILE_BYTES / 2)
                                          : a.wt + (size_t)(n_tile >> 1) * KB * (TILE_BYTES / 2) + (size_t)(n_tile & 1) * (W_BYTES / 2);

  if (skip) {
    // fall through to the teardown
  } else if (warp < 8) {
    // ===================== producers =====================
    if constexpr (A_SPLIT) {
      if (tid == 0) {
        const uint8_t* atile0 = a.a_split + (size_t)m_tile * KB * (2 * TILE_BYTES);
        for (int kb = 0; kb < KB; ++kb) {
          const int s = kb % STAGES;
          const uint32_t phase = (uint32_t)(kb / STAGES) & 1u;
          mbar_wait(empty_bar(s), phase ^ 1u);
          const uint32_t st_base = base + s * STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(s), 2 * TILE_BYTES + W_BYTES);
          if constexpr (MC)
            bulk_copy_g2s_multicast(st_base + rank * TILE_BYTES, atile0 + (size_t)kb * (2 * TILE_BYTES) + (size_t)rank * TILE_BYTES,
                                    TILE_BYTES, full_bar(s), (uint16_t)3);
          else
            bulk_copy_g2s(st_base, atile0 + (size_t)kb * (2 * TILE_BYTES), 2 * TILE_BYTES, full_bar(s));
#pragma unroll
          for (int h = 0; h < NW; ++h)
            bulk_copy_g2s(st_base + (2 + h) * TILE_BYTES, wtile0 + ((size_t)h * KB + kb) * (TILE_BYTES / 2), W_COPY, full_bar(s));
        }
      }
    } else {
      const int grp = warp >> 2, gtid = tid & (GROUP_THREADS - 1);
      for (int kb = grp; kb < KB; kb += 2) {
        const int s = kb % STAGES;                  // == grp: each group owns one stage
        const uint32_t phase = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(empty_bar(s), phase ^ 1u);
        const uint32_t st_base = base + s * STAGE_BYTES;
        if (gtid == 0) {
          mbar_expect_tx(full_bar(s), W_BYTES);
#pragma unroll
          for (int h = 0; h < NW; ++h)
            bulk_copy_g2s(st_base + (2 + h) * TILE_BYTES, wtile0 + ((size_t)h * KB + kb) * (TILE_BYTES / 2), W_COPY, full_bar(s));
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int idx = (half * 8 + it) * GROUP_THREADS + gtid;
            const int row = idx >> 4, c4 = idx & 15;
            const int gm = m0 + row, gk = kb * BK + c4 * 4;
            v[it] = (gm < a.M && gk < a.K) ? *reinterpret_cast<const float4*>(a.x + (size_t)gm * a.ldx + gk)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int idx = (half * 8 + it) * GROUP_THREADS + gtid;
            const int row = idx >> 4, c4 = idx & 15;
            const float4 f = v[it];
            uint32_t hi0, hi1, lo0, lo1;
            split_pair(f.x, f.y, hi0, lo0);
            split_pair(f.z, f.w, hi1, lo1);
            const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)c4 >> 1) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)c4 & 1u) * 8u;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_base + off), "r"(hi0), "r"(hi1) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_base + TILE_BYTES + off), "r"(lo0), "r"(lo1) : "memory");
          }
        }
        fence_proxy_async();              // generic-proxy stores -> visible to the tensor core's async proxy
        mbar_arrive(full_bar(s));
      }
    }
    // ===================== epilogue =====================
    // TMEM -> registers (thread = one row, 32 columns) -> shared staging (the pipeline stages are free once the
    // accumulator barrier fires) -> row-contiguous global accesses: each warp instruction touches whole 128-byte
    // (fp32 output / residual) or 64-byte (SplitTile) row segments instead of 32 scattered 16-byte pieces.
    const int quad = warp & 3, chalf = warp >> 2;
    float* stg = reinterpret_cast<float*>(base_ptr) + warp * (32 * STG_PITCH);
    // The residual is an input of the kernel: fetch this thread's share while the tensor core is still working
    // (it may alias the output, x += f(x), so the compiler could not hoist these loads above earlier stores).
    constexpr int NCC = BN / 64;
    float4 rv[HAS_RES ? NCC : 1][HAS_RES ? 8 : 1];
    if constexpr (HAS_RES) {
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        const int gn = n0 + chalf * (BN / 2) + cc * 32 + 4 * (lane & 7);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int gm = m0 + quad * 32 + it * 4 + (lane >> 3);
          rv[cc][it] = (gm < a.M && gn < a.N) ? *reinterpret_cast<const float4*>(a.residual + (size_t)gm * a.ldr + gn)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // The main loop of this CTA is over: once every CTA is here (or gone) the next kernel of the stream may start
    // launching, so its launch latency and prologue overlap this epilogue and the grid's drain.  (Triggering at kernel
    // entry instead let early dependents take shared memory / TMEM from CTAs of this grid that had not started yet.)
    if (tid == 0) pdl_launch_dependents();
    // The activation is uniform over the launch: pick it once (a per-element switch costs branches and, for
    // GELU, kept ~200 call sites alive in the unrolled epilogue).
    auto epilogue = [&](auto act_tag) {
    constexpr int ACT = decltype(act_tag)::value;
#pragma unroll(HAS_RES ? BN / 64 : 1)
    for (int cc = 0; cc < BN / 64; ++cc) {
      const int c0 = chalf * (BN / 2) + cc * 32;
      uint32_t r[32];
      __syncwarp();                        // tcgen05.ld is .sync.aligned; also fences the previous staging pass
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      const int gn0 = n0 + c0;
      if constexpr (!OUT_SPLIT) {
        const int c4 = lane & 7, rr = lane >> 3;
        const int gn = gn0 + 4 * c4;
        if (gn < a.N) {                    // N is a multiple of 4 on this path
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias != nullptr) bv = *reinterpret_cast<const float4*>(a.bias + gn);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + rr, gm = m0 + quad * 32 + row;
            if (gm < a.M) {
              const float4 v = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + 4 * c4);
              float o0 = apply_act_t<ACT>(v.x + bv.x, a.act), o1 = apply_act_t<ACT>(v.y + bv.y, a.act);
              float o2 = apply_act_t<ACT>(v.z + bv.z, a.act), o3 = apply_act_t<ACT>(v.w + bv.w, a.act);
              if constexpr (HAS_RES) { o0 += rv[cc][it].x; o1 += rv[cc][it].y; o2 += rv[cc][it].z; o3 += rv[cc][it].w; }
              *reinterpret_cast<float4*>(a.y + (size_t)gm * a.ldy + gn) = make_float4(o0, o1, o2, o3);
              if constexpr (NFC) {
                if (nonfinite_hit(o0, a.nonfinite_mode) || nonfinite_hit(o1, a.nonfinite_mode) || nonfinite_hit(o2, a.nonfinite_mode) ||
                    nonfinite_hit(o3, a.nonfinite_mode))
                  atomicOr(a.nonfinite_flag, 1);
              }
            }
          }
        }
      } else {
        // SplitTile output: these 32 columns are chunks chunk0..chunk0+3 of output k-block gn0 / 64
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
                // columns >= N are the zero padding of the next projection's K
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
    // ===================== MMA issuer (warp 8) =====================
    if (lane == 0) {
      // While this projection runs, pull the next one's weights into L2 (they were evicted by a whole decode step of
      // activations and K/V traffic): its first k-blocks then start from L2 instead of a DRAM round trip.
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
        if constexpr (MC) umma_commit_multicast(empty_bar(s), (uint16_t)3);
        else umma_commit(empty_bar(s));  // frees the stage once the MMAs that read it have finished
      }
      umma_commit(accum_bar);            // accumulator complete -> epilogue
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  if constexpr (MC) cluster_sync();       // nobody leaves while the peer's commits may still arrive on its barriers
  trace_end(trc);
}

// fp32 [N, K] row-major -> bf16 tiles [ceil(N/128)][ceil(K/64)][128 x 64, SW128], zero padded.
__global__ void pack_tiled_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int K,
                                  int n_tiles, int kblocks) {
  const int64_t total = (int64_t)n_tiles * kblocks * 8192;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = i >> 13;
    const int within = (int)(i & 8191);
    const int r = within >> 6, pos = within & 63;           // physical position inside the row
    const int chunk_phys = pos >> 3, e = pos & 7;
    const int chunk = chunk_phys ^ (r & 7);                 // logical 16-byte chunk stored at this position
    const int nt = (int)(tile / kblocks), kb = (int)(tile % kblocks);
    const int n = nt * 128 + r, k = kb * 64 + chunk * 8 + e;
    const float v = (n < N && k < K) ? src[(int64_t)n * K + k] : 0.f;
    if constexpr (kSplitFp16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);     // 16-bit storage either way
    else dst[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace

size_t tc_packed_elems(int N, int K) { return (size_t)ceil_div(N, 128) * ceil_div(K, 64) * 8192; }
size_t split_tile_bytes(int M, int K) { return (size_t)ceil_div(M, 128) * ceil_div(K, 64) * (2 * TILE_BYTES); }

int launch_pack_tiled(const float* src, __nv_bfloat16* dst, int N, int K, cudaStream_t s) {
  const int nt = ceil_div(N, 128), kb = ceil_div(K, 64);
  const int64_t total = (int64_t)nt * kb * 8192;
  pack_tiled_kernel<<<(int)std::min<int64_t>(ceil_div64(total, 256), 148 * 16), 256, 0, s>>>(src, dst, N, K, nt, kb);
  SCV_LAUNCH_CHECK();
  return 0;
}

bool tc_shape_ok(const LinearArgs& a) {
  static const int min_rows = [] { const char* e = getenv("SCV_TC_MIN_ROWS"); return e ? atoi(e) : 1; }();
  if (a.wt == nullptr || a.M < min_rows || a.K < 64 || a.N % 4 != 0) return false;
  if (a.a_split == nullptr) {
    if (a.x == nullptr || a.K % 4 != 0 || a.ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15u) != 0) return false;
  }
  if (a.y_split == nullptr) {
    if (a.y == nullptr || a.ldy % 4 != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15u) != 0) return false;
    if (a.bias != nullptr && (reinterpret_cast<uintptr_t>(a.bias) & 15u) != 0) return false;
    if (a.residual != nullptr && (a.ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(a.residual) & 15u) != 0)) return false;
  } else if (a.residual != nullptr) {
    return false;
  }
  return true;
}

int launch_linear_tcgen05(const LinearArgs& a, cudaStream_t s) {
  SCV_REQUIRE(tc_shape_ok(a), "tcgen05 linear: shape/alignment not supported (M=%d N=%d K=%d)", a.M, a.N, a.K);
  if (a.nonfinite_flag == nullptr) {       // (the opt-in variants do not carry the non-finite check)
    if (tc_2cta_ok(a)) return launch_linear_tcgen05_2cta(a, s);
    if (tc_persistent_ok(a)) return launch_linear_tcgen05_persistent(a, s);
  }
  static bool attr_dev[64] = {};
  if (first_use_on_device(attr_dev)) {
#define SCV_SET_SMEM(A, O, R, S, N) \
  SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<A, O, R, S, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(S, N)))
#define SCV_SET_ALL(S, N)                                                                             \
  SCV_SET_SMEM(false, false, false, S, N); SCV_SET_SMEM(false, true, false, S, N); SCV_SET_SMEM(true, false, false, S, N); \
  SCV_SET_SMEM(true, true, false, S, N); SCV_SET_SMEM(false, false, true, S, N); SCV_SET_SMEM(true, false, true, S, N)
    SCV_SET_ALL(2, 128); SCV_SET_ALL(4, 128); SCV_SET_ALL(3, 256); SCV_SET_ALL(4, 64);
    SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true, false, false, 2, 128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  smem_bytes(2, 128)));
#define SCV_SET_MC(O, R, S) \
  SCV_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true, O, R, S, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(S, 128)))
    SCV_SET_MC(false, false, 2); SCV_SET_MC(true, false, 2); SCV_SET_MC(false, true, 2);
    SCV_SET_MC(false, false, 4); SCV_SET_MC(true, false, 4); SCV_SET_MC(false, true, 4);
#undef SCV_SET_MC
#undef SCV_SET_ALL
#undef SCV_SET_SMEM
  }
  TcArgs t;
  t.x = a.x; t.ldx = a.ldx; t.a_split = reinterpret_cast<const uint8_t*>(a.a_split); t.wt = a.wt;
  t.kblocks = ceil_div(a.K, BK); t.bias = a.bias; t.residual = a.residual; t.ldr = a.ldr; t.y = a.y; t.ldy = a.ldy;
  t.y_split = reinterpret_cast<uint8_t*>(a.y_split); t.kb_out = ceil_div(a.N, BK);
  t.M = a.M; t.N = a.N; t.K = a.K; t.act = a.act; t.done_flag = a.done_flag;
  t.next_w = static_cast<const uint8_t*>(a.next_w); t.next_w_bytes = (uint32_t)a.next_w_bytes;
  t.trace = trace_ptr();
  t.row_base = a.row_base;
  t.nonfinite_flag = a.y_split == nullptr ? a.nonfinite_flag : nullptr; t.nonfinite_mode = a.nonfinite_mode;
  // 2 MMAs (hi, lo) per weight tile: algorithmic flops stay 2MNK, the tensor pipe executes twice that
  ProfScope prof(PC_GEMM_TC, s, 2.0 * a.M * a.N * a.K,
                 2.0 * a.N * a.K + 4.0 * a.M * a.K + 4.0 * a.M * a.N * (a.residual ? 2 : 1));
  const bool as = a.a_split != nullptr, os = a.y_split != nullptr, res = a.residual != nullptr;
  // Tile choice: 128 x 128 with a deep 4-stage pipeline when the grid is at most one CTA per SM, else two 2-stage
  // CTAs per SM so that one CTA's epilogue overlaps the other's main loop; 128 x 256 (SCV_GEMM_BN=256) halves the
  // A re-reads but runs one CTA per SM.
  static const int force_bn = [] { const char* e = getenv("SCV_GEMM_BN"); return e ? atoi(e) : 0; }();
  const int n128 = ceil_div(a.N, 128), mt = ceil_div(a.M, BM);
  // measured on B200 (profiles/README.md, r01e): the 256-wide tile is slower at these K (8-32 k-blocks, one CTA per
  // SM so no epilogue overlap), so it stays opt-in
  const bool wide = force_bn == 256 && n128 % 2 == 0;
#define SCV_LAUNCH(A, O, R, S, N) SCV_CUDA(launch_k(gemm_tcgen05_kernel<A, O, R, S, N>, grid, dim3(NUM_THREADS), (size_t)smem_bytes(S, N), s, t))
#define SCV_LAUNCH_MODE(S, N)                                                                        \
  do {                                                                                               \
    if (as && os) SCV_LAUNCH(true, true, false, S, N);                                               \
    else if (os) SCV_LAUNCH(false, true, false, S, N);                                               \
    else if (as && res) SCV_LAUNCH(true, false, true, S, N);                                         \
    else if (as) SCV_LAUNCH(true, false, false, S, N);                                               \
    else if (res) SCV_LAUNCH(false, false, true, S, N);                                              \
    else SCV_LAUNCH(false, false, false, S, N);                                                      \
  } while (0)
  // Few row tiles (a few hundred rows): a 128-wide tile leaves most SMs idle while each CTA's MMAs and epilogue are the
  // critical path of a ~6 us kernel; 128 x 64 tiles halve both (64 columns is where the UMMA's math time equals its
  // A-operand read, so narrower tiles gain nothing).  Same k order per output element: identical bits.
  const int n64 = ceil_div(a.N, 64);
  // Measured in the decode (tests/batch_sweep.py): 9-10 % per step at 64-512 rows, 2-5 % at 1024-2048, but 3 % SLOWER at 4096
  // rows (two streams x 16 row tiles: the A tiles are re-read by twice as many CTAs out of an L2 that is the bottleneck
  // there), hence the bound on the row tiles.
  const bool narrow = !wide && mt <= tun().gemm_bn64 && n64 * mt <= tun().gemm_bn64_max_ctas;
  if (t.nonfinite_flag != nullptr) {
    SCV_REQUIRE(as && !os && !res, "tcgen05 linear: the non-finite check needs SplitTile input, fp32 output, no residual");
    dim3 grid(n128, mt);
    SCV_CUDA(launch_k(gemm_tcgen05_kernel<true, false, false, 2, 128, false, true>, grid, dim3(NUM_THREADS), (size_t)smem_bytes(2, 128), s, t));
    SCV_LAUNCH_CHECK();
    return 0;
  }
  const bool mc = !narrow && !wide && as && tun().gemm_mc != 0 && n128 % 2 == 0 && mt >= tun().gemm_mc_min_row_tiles &&
                  t.kblocks >= tun().gemm_mc_min_kblocks;
  if (mc) {
    dim3 grid(n128, mt);
    const int stages = tun().gemm_stages == 2 || tun().gemm_stages == 4 ? tun().gemm_stages : (n128 * mt <= 148 ? 4 : 2);
#define SCV_LAUNCH_MC(O, R, S) \
  SCV_CUDA(launch_k_cluster(gemm_tcgen05_kernel<true, O, R, S, 128, true>, grid, dim3(NUM_THREADS), (size_t)smem_bytes(S, 128), s, 2u, t))
    if (stages == 4) { if (os) SCV_LAUNCH_MC(true, false, 4); else if (res) SCV_LAUNCH_MC(false, true, 4); else SCV_LAUNCH_MC(false, false, 4); }
    else { if (os) SCV_LAUNCH_MC(true, false, 2); else if (res) SCV_LAUNCH_MC(false, true, 2); else SCV_LAUNCH_MC(false, false, 2); }
#undef SCV_LAUNCH_MC
  } else if (narrow) {
    dim3 grid(n64, mt);
    SCV_LAUNCH_MODE(4, 64);
  } else if (wide) {
    dim3 grid(n128 / 2, mt);
    SCV_LAUNCH_MODE(3, 256);
  } else {
    dim3 grid(n128, mt);
    const int stages = tun().gemm_stages == 2 || tun().gemm_stages == 4 ? tun().gemm_stages : (n128 * mt <= 148 ? 4 : 2);
    if (stages == 4) SCV_LAUNCH_MODE(4, 128); else SCV_LAUNCH_MODE(2, 128);
  }
#undef SCV_LAUNCH_MODE
#undef SCV_LAUNCH
  SCV_LAUNCH_CHECK();
  return 0;
}

}  // namespace scv
