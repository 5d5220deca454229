// tcgen05 flash attention, head_dim 64, bf16 in/out, fp32 softmax statistics.
//
//   out = softmax(q k^T * scale + relpos_bias[h][kj - qi] (+ causal mask)) v      per (sequence, head)
//
// One CTA = 128 queries of one (sequence, head); KV streamed in tiles of 64 keys by TMA (2 stages) straight out of
// the token-major qkv buffer (3-D tensor map, out-of-range rows zero-filled).
//   warp 0      : TMA producer (Q once, then K/V tiles)
//   warp 1      : TMEM allocator + MMA issuer:  S = Q K^T (UMMA 128x64x16, K-major A/B)  and
//                 O_j = P V (A = P written to smem by the softmax warps, B = V tile used MN-major as loaded)
//   warps 2..5  : softmax; thread r owns query row r (TMEM lane r): tcgen05.ld S row -> scale/bias/mask -> running
//                 max/sum in registers -> P (bf16) to swizzled smem -> after the PV MMA, tcgen05.ld O_j and
//                 accumulate o = o*corr + O_j in registers (no TMEM read-modify-write).
// Replaces the materialised [H, T, T] score tensors of QKVAttentionLegacy (arch_util.py:60-77), HF GPT2Attention._attn
// and xtransformers Attention (xtransformers.py:660-712).
#include <type_traits>

#include "common.cuh"
#include "ttb_internal.h"

namespace ttb {

constexpr int FA_BM = 128;   // queries per CTA
constexpr int FA_BN = 64;    // keys per tile
constexpr int FA_THREADS = 192;
constexpr int FA_STAGES = 2;

struct FaSmem {
  static constexpr int Q_BYTES = FA_BM * 64 * 2;          // 16 KB
  static constexpr int K_BYTES = FA_BN * 64 * 2;          // 8 KB
  static constexpr int V_BYTES = FA_BN * 64 * 2;          // 8 KB
  static constexpr int P_BYTES = FA_BM * FA_BN * 2;       // 16 KB
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = Q_BYTES;
  static constexpr int P_OFF = KV_OFF + FA_STAGES * (K_BYTES + V_BYTES);
  static constexpr int BIAS_OFF = P_OFF + P_BYTES;        // 192 floats
  static constexpr int BAR_OFF = BIAS_OFF + 192 * 4;
  static constexpr int TOTAL = BAR_OFF + 16 * 8 + 1024;
};

// DBUF (round-2 experiment, TTB_FA_DBUF=1, not yet run on hardware): S is double-buffered in TMEM and Q K_{j+1}^T is
// issued BEFORE the MMA thread waits for P_j, so the softmax of tile j overlaps the score MMA of tile j+1 instead of
// waiting behind [P V_j, Q K_{j+1}^T]. The softmax threads then wait for o_full(j-1) themselves before they touch O
// (lazy rescale) or overwrite the single P buffer. With DBUF = false the kernel is unchanged.
template <int MINB, bool DBUF = false>
__global__ void __launch_bounds__(FA_THREADS, MINB)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, TtbAttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FaSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;            // [FA_STAGES]
  uint64_t* kv_empty = bars + 1 + FA_STAGES;
  constexpr int NS = DBUF ? 2 : 1;         // score buffers in TMEM
  uint64_t* s_full = bars + 1 + 2 * FA_STAGES;   // [NS]
  uint64_t* p_full = s_full + NS;
  uint64_t* o_full = s_full + NS + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + NS + 2);
  float* sbias = reinterpret_cast<float*>(smem + FaSmem::BIAS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, seq = blockIdx.z;
  const int Tq = a.T;                          // queries per sequence
  const int T = a.Tk > 0 ? a.Tk : a.T;         // keys per sequence (bias / causal need Tk == T)
  const int kv_end = a.causal ? min(T, q0 + FA_BM) : T;
  // KV addressing: packed qkv rows (k_off/v_off + 64*h columns, sequence = 3rd coordinate) or a head-major cache
  // [H][Tk][64] shared by every sequence (kv_headmajor): head = 3rd coordinate.
  const int kc0 = a.kv_headmajor ? 0 : a.k_off + h * 64;
  const int vc0 = a.kv_headmajor ? 0 : a.v_off + h * 64;
  const int kvc2 = a.kv_headmajor ? h : seq;
  const int ntiles = (kv_end + FA_BN - 1) / FA_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int b = 0; b < NS; ++b) mbar_init(&s_full[b], 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<DBUF ? 256 : 128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // barrier init, TMEM allocation and descriptor prefetch above overlap the tail of the qkv GEMM
  const uint32_t tmem_s = tmem_base;                        // columns [0, 64) (+ [64, 128) with DBUF)
  const uint32_t tmem_o = tmem_base + (DBUF ? 128 : 64);    // 64 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, FaSmem::Q_BYTES);
      tma_load_3d(smem + FaSmem::Q_OFF, &map_q, q_full, h * 64, q0, seq);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = smem + FaSmem::KV_OFF + stage * (FaSmem::K_BYTES + FaSmem::V_BYTES);
        uint8_t* sv = sk + FaSmem::K_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], FaSmem::K_BYTES + FaSmem::V_BYTES);
        tma_load_3d(sk, &map_k, &kv_full[stage], kc0, j * FA_BN, kvc2);
        tma_load_3d(sv, &map_v, &kv_full[stage], vc0, j * FA_BN, kvc2);
        if (++stage == FA_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN, 0, 0);   // S[128x64] = Q[128x64d] * K[64keys x 64d]^T
      constexpr uint32_t idesc_o = umma_idesc_bf16(FA_BM, 64, 0, 1);      // O[128x64d] = P[128x64keys] * V (MN-major B)
      const uint32_t sq = smem_u32(smem + FaSmem::Q_OFF);
      const uint32_t sp = smem_u32(smem + FaSmem::P_OFF);
      mbar_wait(q_full, 0);
      int stage = 0; uint32_t phase = 0;
      if constexpr (DBUF) {
        // S_0 first; then per tile: S_{j+1} (other score buffer, next KV stage) BEFORE waiting for P_j, then P V_j
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        {
          const uint32_t sk = smem_u32(smem + FaSmem::KV_OFF);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_s, umma_desc_kmajor_sw128(sq + k * 32), umma_desc_kmajor_sw128(sk + k * 32), idesc_s, k != 0);
          umma_commit(&s_full[0]);
        }
        for (int j = 0; j < ntiles; ++j) {
          const int nstage = (stage + 1 == FA_STAGES) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == FA_STAGES) ? (phase ^ 1) : phase;
          if (j + 1 < ntiles) {
            // score buffer (j+1)&1 was last read by the softmax of tile j-1, which arrived on p_full(j-1) afterwards
            mbar_wait(&kv_full[nstage], nphase);
            tc_fence_after();
            const uint32_t sk = smem_u32(smem + FaSmem::KV_OFF + nstage * (FaSmem::K_BYTES + FaSmem::V_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_s + (uint32_t)(((j + 1) & 1) * 64), umma_desc_kmajor_sw128(sq + k * 32),
                           umma_desc_kmajor_sw128(sk + k * 32), idesc_s, k != 0);
            umma_commit(&s_full[(j + 1) & 1]);
          }
          mbar_wait(p_full, j & 1);
          tc_fence_after();
          const uint32_t sv = smem_u32(smem + FaSmem::KV_OFF + stage * (FaSmem::K_BYTES + FaSmem::V_BYTES)) + FaSmem::K_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_o, umma_desc_kmajor_sw128(sp + k * 32), umma_desc_mnmajor_sw128(sv + k * 2048, 0), idesc_o, (j | k) != 0);
          umma_commit(o_full);
          umma_commit(&kv_empty[stage]);
          stage = nstage; phase = nphase;
        }
      } else
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&kv_full[stage], phase);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + FaSmem::KV_OFF + stage * (FaSmem::K_BYTES + FaSmem::V_BYTES));
        const uint32_t sv = sk + FaSmem::K_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_s, umma_desc_kmajor_sw128(sq + k * 32), umma_desc_kmajor_sw128(sk + k * 32), idesc_s, k != 0);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);            // softmax consumed S_j and wrote P_j
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)           // K (= keys) advances 16 rows of 128 B per step in the MN-major V tile
          umma_bf16_ss(tmem_o, umma_desc_kmajor_sw128(sp + k * 32), umma_desc_mnmajor_sw128(sv + k * 2048, 0), idesc_o, (j | k) != 0);
        umma_commit(o_full);
        umma_commit(&kv_empty[stage]);
        if (++stage == FA_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;            // TMEM lane == query row inside the tile
    const int st = threadIdx.x - 64;           // 0..127 among the softmax threads
    const int qi = q0 + row;
    const float sl2 = a.scale * 1.4426950408889634f;
    const float* bias_h = a.bias ? a.bias + (long long)h * (2 * T - 1) + (T - 1) : nullptr;
    // Running statistics in the log2 domain. O accumulates in TMEM across tiles (the PV MMA runs with
    // enable-input-d); it is rescaled in place only when this row's max grows by more than 2^8 ("lazy rescale"),
    // so P values stay <= 256 and the common case needs no TMEM round trip.
    float m = -INFINITY, l = 0.f;
    uint8_t* sp = smem + FaSmem::P_OFF;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    for (int j = 0; j < ntiles; ++j) {
      const int k0 = j * FA_BN;
      // Tile classification (uniform over the CTA): far from the diagonal the T5 bias is saturated, i.e. constant
      // over the whole tile, and no key is masked -> one FFMA + one EX2 per score ("fast" tiles).
      const bool nomask = (k0 + FA_BN <= T) && (!a.causal || k0 + FA_BN - 1 <= q0);
      bool cbias_ok = (bias_h == nullptr);
      float cb = 0.f;
      if (bias_h && a.bias_sat > 0) {
        if (k0 - q0 + (FA_BN - 1) <= -a.bias_sat) { cbias_ok = true; cb = __ldg(bias_h - (T - 1)) * 1.4426950408889634f; }
        else if (k0 - q0 - (FA_BM - 1) >= a.bias_sat) { cbias_ok = true; cb = __ldg(bias_h + (T - 1)) * 1.4426950408889634f; }
      }
      const bool fast = nomask && cbias_ok;
      if (bias_h && !fast) {
        // window of the Toeplitz bias table needed by this (q-tile, k-tile): rel = kj - qi in [k0-q0-127, k0-q0+63]
        asm volatile("bar.sync 1, 128;" ::: "memory");   // previous tile's readers are done
        for (int i = st; i < 191; i += 128) {
          const int rel = k0 - q0 - 127 + i;
          sbias[i] = (rel > -T && rel < T) ? __ldg(bias_h + rel) * 1.4426950408889634f : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      // The rest of the tile is instantiated once per tile class: with one shared body the score registers of the two
      // classes had to agree at the join, which cost the common (fast) class 64 register moves per tile (SASS of
      // profiles/ncu_r02/fa_dbuf.ncu-rep: 42 IMAD.MOV + 22 MOV in a 100-instruction max block).
      auto tile_body = [&](auto fast_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      // S_j ready (without DBUF this also means that every earlier MMA, incl. P V of tile j-1, has retired)
      mbar_wait(&s_full[DBUF ? (j & 1) : 0], DBUF ? ((j >> 1) & 1) : (j & 1));
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_s + (DBUF ? (uint32_t)((j & 1) * 64) : 0u) + lane_base, r0);
      tmem_ld_32x32b_x32(tmem_s + (DBUF ? (uint32_t)((j & 1) * 64) : 0u) + lane_base + 32, r1);
      tmem_ld_wait();
      float mx = -INFINITY;
      if constexpr (FAST) {
        float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains (ILP)
#pragma unroll
        for (int c = 0; c < 32; ++c) mq[c & 3] = fmaxf(mq[c & 3], fmaxf(__uint_as_float(r0[c]), __uint_as_float(r1[c])));
        mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
        mx = fmaf(mx, sl2, cb);
      } else {
        const int klim = min(T - k0, a.causal ? (qi - k0 + 1) : FA_BN);   // keys c < klim are visible to this row
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float v = __uint_as_float(c < 32 ? r0[c] : r1[c - 32]) * sl2;
          if (bias_h) v += sbias[c - row + 127];
          v = (c < klim) ? v : -INFINITY;
          if (c < 32) r0[c] = __float_as_uint(v); else r1[c - 32] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      }
      if constexpr (DBUF) {
        // O (lazy rescale below) and the P buffer belong to P V_{j-1} until it has retired
        if (j > 0) { mbar_wait(o_full, (j - 1) & 1); tc_fence_after(); }
      }
      const bool grow = mx > m + 8.0f;                 // also true for the first finite max (m = -inf)
      const float m_new = grow ? mx : m;
      const float corr = grow ? exp2f(m - m_new) : 1.0f;   // exp2(-inf) = 0 on the first tile
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        uint32_t t0[32];
        tmem_ld_32x32b_x32(tmem_o + lane_base, t0);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) t0[d] = __float_as_uint(__uint_as_float(t0[d]) * corr);
        tmem_st_32x32b_x32(tmem_o + lane_base, t0);
        tmem_ld_32x32b_x32(tmem_o + lane_base + 32, t0);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) t0[d] = __float_as_uint(__uint_as_float(t0[d]) * corr);
        tmem_st_32x32b_x32(tmem_o + lane_base + 32, t0);
        tmem_st_wait();
      }
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      float psum = 0.f;
      uint32_t pk[32];
      if constexpr (FAST) {
        const float cm = cb - m_use;
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float p0 = exp2f(fmaf(__uint_as_float(c < 32 ? r0[c] : r1[c - 32]), sl2, cm));
          const float p1 = exp2f(fmaf(__uint_as_float(c < 32 ? r0[c + 1] : r1[c - 31]), sl2, cm));
          ps[(c >> 1) & 3] += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
        psum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float p0 = exp2f(__uint_as_float(c < 32 ? r0[c] : r1[c - 32]) - m_use);
          const float p1 = exp2f(__uint_as_float(c < 32 ? r0[c + 1] : r1[c - 31]) - m_use);
          psum += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
      }
      l = l * corr + psum;
      m = m_new;
      // P row (64 bf16 = 128 B) into the K-major SWIZZLE_128B layout: 16-byte chunk c8 lands at (c8 ^ (row & 7))
      uint4* prow = reinterpret_cast<uint4*>(sp + row * 128);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        prow[c8 ^ (row & 7)] = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
      };
      if (fast) tile_body(std::true_type{}); else tile_body(std::false_type{});
    }
    // all P V products have been issued; wait for the last one and read the accumulated O row
    mbar_wait(o_full, (ntiles - 1) & 1);
    tc_fence_after();
    float o[64];
    {
      uint32_t t0[32], t1[32];
      tmem_ld_32x32b_x32(tmem_o + lane_base, t0);
      tmem_ld_32x32b_x32(tmem_o + lane_base + 32, t1);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 32; ++d) { o[d] = __uint_as_float(t0[d]); o[d + 32] = __uint_as_float(t1[d]); }
    }
    tc_fence_before();
    if (qi < Tq && a.lse) {
      // partial-attention mode: normalised O in fp32 + log2-sum-exp, to be merged with another key range
      const float inv = 1.0f / l;
      float* op = a.out_f32 + ((long long)seq * Tq + qi) * a.ldo + h * 64;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        reinterpret_cast<float4*>(op)[i] = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
      a.lse[((long long)seq * Tq + qi) * a.H + h] = m + log2f(l);
    } else if (qi < Tq) {
      const float inv = 1.0f / l;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((long long)seq * Tq + qi) * a.ldo + h * 64;
      uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o4[i] = make_uint4(pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                           pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
    }
  }
  pdl_launch_dependents();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<DBUF ? 256 : 128>(tmem_base);
  }
}

static int g_fa_mode = -1;  // 1 = tcgen05 (default), 0 = SIMT (TTB_ATTN_IMPL=simt)

bool flash_attention_supported(const TtbAttnArgs& a) {
  if (g_fa_mode < 0) {
    const char* e = getenv("TTB_ATTN_IMPL");
    g_fa_mode = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  if (!g_fa_mode) return false;
  if (a.kv || a.lse) return true;      // split Q / KV tensors and LSE output exist only in the tcgen05 kernel
  return a.T >= 64 && (a.ld % 8) == 0;
}

int flash_attention_launch(const TtbAttnArgs& a, cudaStream_t st) {
  CUtensorMap mq, mk, mv;
  if (get_tensor_map_bf16(&mq, a.qkv, (uint64_t)a.ld, (uint64_t)a.T, (uint64_t)a.nseq, (uint64_t)a.ld,
                          (uint64_t)a.T * a.ld, 64, FA_BM)) return -1;
  if (a.kv_headmajor) {
    if (!a.kv || !a.kv_v || a.Tk <= 0 || a.causal || a.bias) { set_error("flash attention: bad head-major KV arguments"); return -1; }
    if (get_tensor_map_bf16(&mk, a.kv, 64, (uint64_t)a.Tk, (uint64_t)a.H, 64, (uint64_t)a.Tk * 64, 64, FA_BN)) return -1;
    if (get_tensor_map_bf16(&mv, a.kv_v, 64, (uint64_t)a.Tk, (uint64_t)a.H, 64, (uint64_t)a.Tk * 64, 64, FA_BN)) return -1;
  } else {
    if (a.Tk > 0 && a.Tk != a.T) { set_error("flash attention: Tk != T needs the head-major KV form"); return -1; }
    if (get_tensor_map_bf16(&mk, a.qkv, (uint64_t)a.ld, (uint64_t)a.T, (uint64_t)a.nseq, (uint64_t)a.ld,
                            (uint64_t)a.T * a.ld, 64, FA_BN)) return -1;
    mv = mk;
  }
  if (a.lse && !a.out_f32) { set_error("flash attention: lse output needs out_f32"); return -1; }
  static int dbuf = -1;
  if (dbuf < 0) { const char* e = getenv("TTB_FA_DBUF"); dbuf = (e && e[0] == '0') ? 0 : 1; }   // default on (0.109 -> 0.098 ms)
  if (dbuf) {
    static bool attr = false;
    if (!attr) {
      cudaError_t r = cudaFuncSetAttribute(flash_attn_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaSmem::TOTAL);
      if (r != cudaSuccess) return check_cuda(r, "cudaFuncSetAttribute(flash_attn dbuf)");
      attr = true;
    }
    dim3 grid_d((a.T + FA_BM - 1) / FA_BM, a.H, a.nseq);
    const cudaError_t led = launch_pdl(flash_attn_tc_kernel<2, true>, grid_d, dim3(FA_THREADS), (size_t)FaSmem::TOTAL, st, mq, mk, mv, a);
    if (led != cudaSuccess) return check_cuda(led, "flash_attn_tc_kernel<dbuf> launch");
    TTB_CHECK_LAUNCH("flash_attn_tc_kernel<dbuf>");
    return 0;
  }
  static int occ = 0;
  if (!occ) {
    const char* e = getenv("TTB_FA_OCC");       // CTAs per SM the kernel is compiled for (register cap): 2 (default) or 3
    occ = (e && atoi(e) == 3) ? 3 : 2;
    cudaError_t r = (occ == 3)
        ? cudaFuncSetAttribute(flash_attn_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaSmem::TOTAL)
        : cudaFuncSetAttribute(flash_attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaSmem::TOTAL);
    if (r != cudaSuccess) { occ = 0; return check_cuda(r, "cudaFuncSetAttribute(flash_attn)"); }
  }
  dim3 grid((a.T + FA_BM - 1) / FA_BM, a.H, a.nseq);
  const cudaError_t le = (occ == 3) ? launch_pdl(flash_attn_tc_kernel<3>, grid, dim3(FA_THREADS), (size_t)FaSmem::TOTAL, st, mq, mk, mv, a)
                                    : launch_pdl(flash_attn_tc_kernel<2>, grid, dim3(FA_THREADS), (size_t)FaSmem::TOTAL, st, mq, mk, mv, a);
  if (le != cudaSuccess) return check_cuda(le, "flash_attn_tc_kernel launch");
  TTB_CHECK_LAUNCH("flash_attn_tc_kernel");
  return 0;
}

}  // namespace ttb
