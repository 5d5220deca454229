// CTA-pair (cta_group::2) variant of the tcgen05 GEMM (included by gemm.cu) -- ROUND-2 EXPERIMENT, NOT YET RUN ON HARDWARE.
// Selected only by TtbGemmArgs.variant == 6; nothing on the product path uses it.
//
// Why: round-1 measurements (profiles/gemm_trace_r01*.txt) show that one CTA receives its TMA operands at ~46 B/clk
// whatever the pipeline depth, so a 128x128 tile (32 KB per 64-wide k-block, 256 tensor clocks) is delivery-bound at
// ~65 % of the tensor rate even with two CTAs per SM. A CTA PAIR computing a 256 x BN tile with one
// tcgen05.mma.cta_group::2 (UMMA M = 256) halves the weight bytes per CTA: each CTA loads its own 128 activation
// rows and only HALF of the BN weight rows, and the tensor cores of both SMs read both halves. With BN = 256 a CTA
// moves 32 KB per k-block for 128 x 256 outputs, twice the arithmetic per delivered byte.
//
// Structure (one 256 x BN tile per pair, cluster {2,1,1}, 192 threads per CTA):
//   warp 0 lane 0 (both CTAs): TMA producer; loads land in the CTA's own shared memory but complete_tx on the LEADER's
//     (cluster rank 0) full barrier (`.cta_group::2` TMA form, barrier address with the peer bit cleared); the leader
//     arms that barrier with the bytes of both CTAs, the peer adds a plain remote arrive (count 2).
//   warp 1 lane 0 (leader only): issues tcgen05.mma.cta_group::2; tcgen05.commit.cta_group::2 multicasts the arrive to
//     the empty barrier / accumulator barrier of BOTH CTAs.
//   warp 1 (both CTAs): tcgen05.alloc.cta_group::2 / dealloc (same warp index in both CTAs, as the ISA requires).
//   warps 2..5 (both CTAs): epilogue of the CTA's own 128 x BN accumulator (gemm_epilogue.cuh).
// PTX forms follow the CUTLASS sm100 headers shipped in this image (cute/arch/copy_sm100_tma.hpp:104-128,
// cutlass/arch/barrier.h:811-921, cute/arch/tmem_allocator_sm100.hpp:98-170).
#pragma once

namespace ttb {

constexpr uint32_t SM100_PEER_BIT_MASK = 0xFEFFFFFFu;      // clears the CTA-pair bit of a shared::cluster address

template <int kCols>
TTB_DEVINL void tmem_alloc_2sm(uint32_t* smem_dst) {        // one full warp, same warp index, in BOTH CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
TTB_DEVINL void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on the pair leader's barrier at the same offset
TTB_DEVINL void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & SM100_PEER_BIT_MASK), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}
TTB_DEVINL void mbar_arrive_on_leader(uint64_t* bar) {      // plain arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & SM100_PEER_BIT_MASK) : "memory");
}
TTB_DEVINL void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
TTB_DEVINL void umma_commit_2sm(uint64_t* bar) {            // arrive on `bar` in BOTH CTAs when the MMAs have retired
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int BN, int STAGES>
struct Gemm2Smem {
  static constexpr int A_BYTES = BM * BK * 2;               // this CTA's 128 activation rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;         // this CTA's half of the weight rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
  static_assert(TOTAL <= 227 * 1024, "2-CTA GEMM shared memory");
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N,
                         int K, int taps, int pad, int a_batch_mul, GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  // NOTE: both CTAs must carve shared memory identically (the MMA addresses the peer's tiles at the same offsets)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using L = Gemm2Smem<BN, STAGES>;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int m0 = pair * (2 * BM) + (int)rank * BM;          // this CTA's 128 output rows
  const int n0 = blockIdx.y * BN;                           // the pair's BN output columns
  const int nb0 = n0 + (int)rank * (BN / 2);                // this CTA's half of the weight rows
  const int bz = blockIdx.z;
  const int kblocks_per_tap = K / BK;
  const int num_kb = kblocks_per_tap * taps;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  cluster_sync_all();                                       // barriers of both CTAs exist before anyone signals them
  if (warp == 1) tmem_alloc_2sm<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const int tap = kb / kblocks_per_tap;
        const int kk = (kb - tap * kblocks_per_tap) * BK;
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);   // bytes of both CTAs
        tma_load_3d_2sm(sa, &map_a, &full_bar[stage], kk, m0 + tap - pad, bz * a_batch_mul);
        tma_load_3d_2sm(sb, &map_b, &full_bar[stage], tap * K + kk, nb0, 0);
        if (rank != 0) mbar_arrive_on_leader(&full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16_ss_2sm(tmem_base, umma_desc_kmajor_sw128(sa + k * 32), umma_desc_kmajor_sw128(sb + k * 32), idesc,
                           (kb | k) != 0 ? 1u : 0u);
        umma_commit_2sm(&empty_bar[stage]);                 // frees the slot in both CTAs
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(accum_bar);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    gemm_epilogue_dispatch<BN>(tmem_base + ((uint32_t)(q * 32) << 16), n0, N, m0 + q * 32, M, lane, (long long)bz, ep,
                               smem_u32(smem + (warp - 2) * EPI_SCRATCH_BYTES), nullptr);
    tc_fence_before();
  }
  __syncthreads();
  __syncwarp();
  cluster_sync_all();                                       // nobody frees TMEM or exits while the peer still computes
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<BN>(tmem_base);
  }
}

template <int BN, int STAGES>
static int launch_2cta(const TtbGemmArgs& g, const GemmEpilogue& ep, cudaStream_t st) {
  CUtensorMap ma, mb;
  const bool bcast = (g.batch == 1) || (g.a_bstride == 0);
  const uint64_t a_d2 = bcast ? 1 : (uint64_t)g.batch;
  const uint64_t a_s2 = bcast ? (uint64_t)g.rows * g.lda : (uint64_t)g.a_bstride;
  if (get_tensor_map_bf16(&ma, g.A, (uint64_t)g.K, (uint64_t)g.rows, a_d2, (uint64_t)g.lda, a_s2, BK, BM)) return -1;
  if (get_tensor_map_bf16(&mb, g.W, (uint64_t)g.K * g.taps, (uint64_t)g.N, 1, (uint64_t)g.K * g.taps,
                          (uint64_t)g.K * g.taps * g.N, BK, BN / 2)) return -1;
  using L = Gemm2Smem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(gemm 2cta)");
    attr_set = true;
  }
  const int m_pairs = (g.M + 2 * BM - 1) / (2 * BM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * m_pairs, (g.N + BN - 1) / BN, g.batch);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tc_2cta_kernel<BN, STAGES>, ma, mb, g.M, g.N, g.K, g.taps, g.pad,
                                     bcast ? 0 : 1, ep);
  if (e != cudaSuccess) return check_cuda(e, "gemm_bf16_tc_2cta_kernel launch");
  TTB_CHECK_LAUNCH("gemm_bf16_tc_2cta_kernel");
  return 0;
}

}  // namespace ttb
