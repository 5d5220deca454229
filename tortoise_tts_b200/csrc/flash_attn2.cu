// tcgen05 flash attention, ping-pong form: TWO 128-query tiles per CTA share every K/V tile (head_dim 64).
//
//   out = softmax(q k^T * scale + relpos_bias[h][kj - qi] (+ causal mask)) v      per (sequence, head)
//
// Round 1's kernel (flash_attn.cu) ran S = Q K^T -> softmax -> O += P V strictly one after the other inside a CTA: the
// tensor pipe was 15 % busy and the exponent unit 30 % (profiles/ncu_summary_r01_run10.txt), i.e. the kernel was bound
// by the dependent chain MMA -> tcgen05.ld -> softmax -> st.shared -> MMA of ONE tile. Here a CTA owns two query tiles A
// and B with their own score / output accumulators in TMEM and their own softmax warpgroup; the single MMA thread issues
//     S_A(0) S_B(0) | P V_A(j), S_A(j+1) | P V_B(j), S_B(j+1) | ...
// so that while warpgroup A does the softmax of tile j the tensor core serves B, and vice versa. K/V tiles are loaded
// once for both query tiles (half the L2 -> SM traffic of the one-tile kernel).
//   warp 0        : TMA producer (Q_A, Q_B once; K/V tiles of 64 keys, 3 stages)
//   warp 1        : TMEM allocator + MMA issuer
//   warps 2..5    : softmax warpgroup A (thread r <-> query row r of tile A = TMEM lane r)
//   warps 6..9    : softmax warpgroup B
// O accumulates in TMEM over all key tiles; it is rescaled in place only when a row's running max grows by more than 2^8.
// Replaces the materialised [H, T, T] score tensors of QKVAttentionLegacy (arch_util.py:60-77), HF GPT2Attention._attn
// and xtransformers Attention (xtransformers.py:660-712).
#include "common.cuh"
#include "ttb_internal.h"

namespace ttb {

constexpr int F2_BM = 128;
constexpr int F2_BN = 64;
constexpr int F2_THREADS = 320;
constexpr int F2_STAGES = 3;

struct F2Smem {
  static constexpr int Q_BYTES = F2_BM * 64 * 2;          // 16 KB per query tile
  static constexpr int KV_BYTES = 2 * F2_BN * 64 * 2;     // K + V tile: 16 KB
  static constexpr int P_BYTES = F2_BM * F2_BN * 2;       // 16 KB per query tile
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = 2 * Q_BYTES;
  static constexpr int P_OFF = KV_OFF + F2_STAGES * KV_BYTES;
  static constexpr int BIAS_OFF = P_OFF + 2 * P_BYTES;    // 2 x 192 floats
  static constexpr int BAR_OFF = BIAS_OFF + 2 * 192 * 4;
  static constexpr int TOTAL = BAR_OFF + 32 * 8 + 1024;
};

// exp2 on the FMA pipe (Cody-Waite split + degree-4 polynomial on [-0.5, 0.5], rel. error 4e-6 << bf16 rounding of P):
// takes a share of the exponentials off the 16-lane/clk exponent unit, which bounds the softmax at head_dim 64.
TTB_DEVINL float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float r = rintf(x);
  const float f = x - r;                    // [-0.5, 0.5]
  float p = 1.3333558146e-3f;
  p = fmaf(p, f, 9.6181291076e-3f);
  p = fmaf(p, f, 5.5504108665e-2f);
  p = fmaf(p, f, 2.4022650696e-1f);
  p = fmaf(p, f, 6.9314718056e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((int)r << 23));
}

template <bool POLY>
__global__ void __launch_bounds__(F2_THREADS, 1)
flash_attn2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, TtbAttnArgs a,
                   int turns) {
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F2Smem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                       // [F2_STAGES]
  uint64_t* kv_empty = kv_full + F2_STAGES;           // [F2_STAGES]
  uint64_t* s_full = kv_empty + F2_STAGES;            // [2]
  uint64_t* p_full = s_full + 2;                      // [2]
  uint64_t* o_full = p_full + 2;                      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, seq = blockIdx.z;
  const int T = a.T;
  const int q0A = blockIdx.x * 2 * F2_BM, q0B = q0A + F2_BM;
  // number of key tiles each query tile needs (causal: up to its own diagonal block); B >= A
  const int endA = a.causal ? min(T, q0A + F2_BM) : T;
  const int endB = (q0B < T) ? (a.causal ? min(T, q0B + F2_BM) : T) : 0;
  const int nA = (endA + F2_BN - 1) / F2_BN;
  const int nB = (endB + F2_BN - 1) / F2_BN;
  const int ntiles = max(nA, nB);
  const int kc0 = a.k_off + h * 64, vc0 = a.v_off + h * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < F2_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 128); mbar_init(&o_full[g], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S_A [0,64) S_B [64,128) O_A [128,192) O_B [192,256)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * F2Smem::Q_BYTES);
      tma_load_3d(smem + F2Smem::Q_OFF, &map_q, q_full, h * 64, q0A, seq);
      tma_load_3d(smem + F2Smem::Q_OFF + F2Smem::Q_BYTES, &map_q, q_full, h * 64, q0B, seq);   // rows >= T zero-filled
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = smem + F2Smem::KV_OFF + stage * F2Smem::KV_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], F2Smem::KV_BYTES);
        tma_load_3d(sk, &map_kv, &kv_full[stage], kc0, j * F2_BN, seq);
        tma_load_3d(sk + F2Smem::KV_BYTES / 2, &map_kv, &kv_full[stage], vc0, j * F2_BN, seq);
        if (++stage == F2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(F2_BM, F2_BN, 0, 0);   // S[128x64] = Q[128x64d] K[64 keys x 64d]^T
      constexpr uint32_t idesc_o = umma_idesc_bf16(F2_BM, 64, 0, 1);      // O[128x64d] += P[128x64 keys] V (MN-major B)
      const uint32_t sq = smem_u32(smem + F2Smem::Q_OFF);
      const uint32_t sp = smem_u32(smem + F2Smem::P_OFF);
      const uint32_t skv = smem_u32(smem + F2Smem::KV_OFF);
      auto issue_s = [&](int g, int stage) {
        const uint32_t sk = skv + stage * F2Smem::KV_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + g * 64, umma_desc_kmajor_sw128(sq + g * F2Smem::Q_BYTES + k * 32),
                       umma_desc_kmajor_sw128(sk + k * 32), idesc_s, k != 0);
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int stage, int j) {
        const uint32_t sv = skv + stage * F2Smem::KV_BYTES + F2Smem::KV_BYTES / 2;
#pragma unroll
        for (int k = 0; k < 4; ++k)           // the key index advances 16 rows of 128 B per step in the MN-major V tile
          umma_bf16_ss(tmem_base + 128 + g * 64, umma_desc_kmajor_sw128(sp + g * F2Smem::P_BYTES + k * 32),
                       umma_desc_mnmajor_sw128(sv + k * 2048, 0), idesc_o, (j | k) != 0);
        umma_commit(&o_full[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (0 < nA) issue_s(0, 0);
      if (0 < nB) issue_s(1, 0);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int nstage = (stage + 1 == F2_STAGES) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == F2_STAGES) ? (phase ^ 1) : phase;
        bool next_ready = false;
        if (j < nA) {
          mbar_wait(&p_full[0], j & 1);        // warpgroup A consumed S_A(j) and wrote P_A(j)
          tc_fence_after();
          issue_pv(0, stage, j);
          if (j + 1 < nA) {
            mbar_wait(&kv_full[nstage], nphase);
            tc_fence_after();
            next_ready = true;
            issue_s(0, nstage);
          }
        }
        if (j < nB) {
          mbar_wait(&p_full[1], j & 1);
          tc_fence_after();
          issue_pv(1, stage, j);
          umma_commit(&kv_empty[stage]);       // every MMA that reads K/V stage j has been issued
          if (j + 1 < nB) {
            if (!next_ready) { mbar_wait(&kv_full[nstage], nphase); tc_fence_after(); }
            issue_s(1, nstage);
          }
        } else {
          umma_commit(&kv_empty[stage]);
        }
        stage = nstage; phase = nphase;
      }
    }
  } else {
    const int g = (warp - 2) >> 2;               // softmax warpgroup: 0 = tile A, 1 = tile B
    const int qd = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = qd * 32 + lane;              // query row inside the tile == TMEM lane
    const int st = (threadIdx.x - 64) & 127;     // 0..127 inside the warpgroup
    const int q0 = g ? q0B : q0A;
    const int nt = g ? nB : nA;
    const int qi = q0 + row;
    const float sl2 = a.scale * 1.4426950408889634f;
    const float* bias_h = a.bias ? a.bias + (long long)h * (2 * T - 1) + (T - 1) : nullptr;
    float* sbias = reinterpret_cast<float*>(smem + F2Smem::BIAS_OFF) + g * 192;
    const uint32_t tmem_s = tmem_base + g * 64 + ((uint32_t)(qd * 32) << 16);
    const uint32_t tmem_o = tmem_base + 128 + g * 64 + ((uint32_t)(qd * 32) << 16);
    uint8_t* sp = smem + F2Smem::P_OFF + g * F2Smem::P_BYTES;
    const int bar_id = 1 + g;
    // Softmax turn-taking (named barriers 3 / 4, 256 threads = both warpgroups): without it the two warpgroups run in
    // phase - both do their softmax at the same time, fighting for the issue slots and the exponent unit, then both wait
    // for the tensor core - and the kernel is no faster than one tile per CTA (measured: 0.110 ms either way). With it,
    // group B's softmax of tile j runs while group A waits for S_A(j+1) and vice versa.
    const int my_turn = 3 + g, other_turn = 4 - g;
    if (turns && g == 1) asm volatile("bar.arrive %0, 256;" ::"r"(3) : "memory");      // group A goes first
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      if (j >= nt) {                         // this group has no tile j (causal: fewer key tiles; tile B out of range)
        if (turns) {
          asm volatile("bar.sync %0, 256;" ::"r"(my_turn) : "memory");
          asm volatile("bar.arrive %0, 256;" ::"r"(other_turn) : "memory");
        }
        continue;
      }
      const int k0 = j * F2_BN;
      // tile classes (uniform over the warpgroup): far from the diagonal the T5 bias is constant over the tile
      // (buckets saturate at max_distance) and no key is masked -> one FFMA + one EX2 per score
      const bool nomask = (k0 + F2_BN <= T) && (!a.causal || k0 + F2_BN - 1 <= q0);
      bool cbias_ok = (bias_h == nullptr);
      float cb = 0.f;
      if (bias_h && a.bias_sat > 0) {
        if (k0 - q0 + (F2_BN - 1) <= -a.bias_sat) { cbias_ok = true; cb = __ldg(bias_h - (T - 1)) * 1.4426950408889634f; }
        else if (k0 - q0 - (F2_BM - 1) >= a.bias_sat) { cbias_ok = true; cb = __ldg(bias_h + (T - 1)) * 1.4426950408889634f; }
      }
      const bool fast = nomask && cbias_ok;
      if (bias_h && !fast) {
        // window of the Toeplitz bias table for this (q tile, k tile): rel = kj - qi in [k0-q0-127, k0-q0+63]
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // previous tile's readers are done
        for (int i = st; i < 191; i += 128) {
          const int rel = k0 - q0 - 127 + i;
          sbias[i] = (rel > -T && rel < T) ? __ldg(bias_h + rel) * 1.4426950408889634f : 0.f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      // S(j) complete; the commit behind it also covers P V(j-1): O and the P buffer are free
      mbar_wait(&s_full[g], j & 1);
      if (turns) asm volatile("bar.sync %0, 256;" ::"r"(my_turn) : "memory");
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_s, r0);
      tmem_ld_32x32b_x32(tmem_s + 32, r1);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (fast) {
        float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 32; ++c) mq[c & 3] = fmaxf(mq[c & 3], fmaxf(__uint_as_float(r0[c]), __uint_as_float(r1[c])));
        mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
        mx = fmaf(mx, sl2, cb);
      } else {
        const int klim = min(T - k0, a.causal ? (qi - k0 + 1) : F2_BN);   // keys c < klim are visible to this row
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float v = __uint_as_float(c < 32 ? r0[c] : r1[c - 32]) * sl2;
          if (bias_h) v += sbias[c - row + 127];
          v = (c < klim) ? v : -INFINITY;
          if (c < 32) r0[c] = __float_as_uint(v); else r1[c - 32] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      }
      const bool grow = mx > m + 8.0f;                 // also true for the first finite max (m = -inf)
      const float m_new = grow ? mx : m;
      const float corr = grow ? exp2f(m - m_new) : 1.0f;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        uint32_t t0[32];
        tmem_ld_32x32b_x32(tmem_o, t0);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) t0[d] = __float_as_uint(__uint_as_float(t0[d]) * corr);
        tmem_st_32x32b_x32(tmem_o, t0);
        tmem_ld_32x32b_x32(tmem_o + 32, t0);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) t0[d] = __float_as_uint(__uint_as_float(t0[d]) * corr);
        tmem_st_32x32b_x32(tmem_o + 32, t0);
        tmem_st_wait();
      }
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
      if (fast) {
        const float cm = cb - m_use;
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float x0 = fmaf(__uint_as_float(c < 32 ? r0[c] : r1[c - 32]), sl2, cm);
          const float x1 = fmaf(__uint_as_float(c < 32 ? r0[c + 1] : r1[c - 31]), sl2, cm);
          // every 4th pair on the FMA pipe (POLY): balances the exponent unit against the FMA pipe
          const bool poly = POLY && ((c & 6) == 6);
          const float p0 = poly ? exp2_poly(x0) : exp2f(x0);
          const float p1 = poly ? exp2_poly(x1) : exp2f(x1);
          ps[(c >> 1) & 3] += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float p0 = exp2f(__uint_as_float(c < 32 ? r0[c] : r1[c - 32]) - m_use);
          const float p1 = exp2f(__uint_as_float(c < 32 ? r0[c + 1] : r1[c - 31]) - m_use);
          ps[(c >> 1) & 3] += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
      }
      l = l * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
      m = m_new;
      // P row (64 bf16 = 128 B) into the K-major SWIZZLE_128B layout: 16-byte chunk c8 lands at (c8 ^ (row & 7))
      uint4* prow = reinterpret_cast<uint4*>(sp + row * 128);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        prow[c8 ^ (row & 7)] = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      if (turns) asm volatile("bar.arrive %0, 256;" ::"r"(other_turn) : "memory");
    }
    // balance the last arrive of the other group (every bar.arrive needs its bar.sync before the barrier id is reused)
    if (turns && g == 0) asm volatile("bar.sync %0, 256;" ::"r"(3) : "memory");
    if (nt > 0) {
      mbar_wait(&o_full[g], (nt - 1) & 1);
      tc_fence_after();
      uint32_t t0[32], t1[32];
      tmem_ld_32x32b_x32(tmem_o, t0);
      tmem_ld_32x32b_x32(tmem_o + 32, t1);
      tmem_ld_wait();
      tc_fence_before();
      if (qi < T) {
        const float inv = 1.0f / l;
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((long long)seq * T + qi) * a.ldo + h * 64;
        uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[i] = make_uint4(pack_bf16(__uint_as_float(t0[8 * i]) * inv, __uint_as_float(t0[8 * i + 1]) * inv),
                             pack_bf16(__uint_as_float(t0[8 * i + 2]) * inv, __uint_as_float(t0[8 * i + 3]) * inv),
                             pack_bf16(__uint_as_float(t0[8 * i + 4]) * inv, __uint_as_float(t0[8 * i + 5]) * inv),
                             pack_bf16(__uint_as_float(t0[8 * i + 6]) * inv, __uint_as_float(t0[8 * i + 7]) * inv));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[4 + i] = make_uint4(pack_bf16(__uint_as_float(t1[8 * i]) * inv, __uint_as_float(t1[8 * i + 1]) * inv),
                                 pack_bf16(__uint_as_float(t1[8 * i + 2]) * inv, __uint_as_float(t1[8 * i + 3]) * inv),
                                 pack_bf16(__uint_as_float(t1[8 * i + 4]) * inv, __uint_as_float(t1[8 * i + 5]) * inv),
                                 pack_bf16(__uint_as_float(t1[8 * i + 6]) * inv, __uint_as_float(t1[8 * i + 7]) * inv));
      }
    }
  }
  pdl_launch_dependents();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

bool flash_attention2_supported(const TtbAttnArgs& a) {
  static int on = -1;
  // Measured on B200 at 2 x 16 heads, S = 1872 (tools/dbg/fa_time.py): one-tile kernel 0.109 ms, + double-buffered scores
  // (TTB_FA_DBUF, default) 0.098 ms; this kernel 0.102 ms (0.111 with the polynomial exp2 share, 0.121 with softmax
  // turn-taking): no gain from the ping-pong form - two co-resident CTAs of the one-tile kernel already interleave two
  // tiles per SM, and the softmax warps are bound by their own dependent chains (ncu: issue 39 %, XU 36 %, tensor 15 %).
  // Kept selectable (TTB_FA2=1) and tested; not the default.
  if (on < 0) { const char* e = getenv("TTB_FA2"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on) return false;
  if (a.kv || a.lse || a.out_f32) return false;       // partial-attention form stays with flash_attn.cu
  if (a.Tk > 0 && a.Tk != a.T) return false;
  return a.T >= 192 && (a.ld % 8) == 0;
}

int flash_attention2_launch(const TtbAttnArgs& a, cudaStream_t st) {
  CUtensorMap mq, mkv;
  if (get_tensor_map_bf16(&mq, a.qkv, (uint64_t)a.ld, (uint64_t)a.T, (uint64_t)a.nseq, (uint64_t)a.ld,
                          (uint64_t)a.T * a.ld, 64, F2_BM)) return -1;
  if (get_tensor_map_bf16(&mkv, a.qkv, (uint64_t)a.ld, (uint64_t)a.T, (uint64_t)a.nseq, (uint64_t)a.ld,
                          (uint64_t)a.T * a.ld, 64, F2_BN)) return -1;
  static int poly = -1, turns = -1;
  if (poly < 0) { const char* e = getenv("TTB_FA2_POLY"); poly = (e && e[0] == '1') ? 1 : 0; }
  if (turns < 0) { const char* e = getenv("TTB_FA2_TURNS"); turns = (e && e[0] == '1') ? 1 : 0; }
  static bool attr = false;
  if (!attr) {
    cudaError_t r = cudaFuncSetAttribute(flash_attn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2Smem::TOTAL);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(flash_attn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2Smem::TOTAL);
    if (r != cudaSuccess) return check_cuda(r, "cudaFuncSetAttribute(flash_attn2)");
    attr = true;
  }
  dim3 grid((a.T + 2 * F2_BM - 1) / (2 * F2_BM), a.H, a.nseq);
  const cudaError_t le = poly ? launch_pdl(flash_attn2_kernel<true>, grid, dim3(F2_THREADS), (size_t)F2Smem::TOTAL, st, mq, mkv, a, turns)
                              : launch_pdl(flash_attn2_kernel<false>, grid, dim3(F2_THREADS), (size_t)F2Smem::TOTAL, st, mq, mkv, a, turns);
  if (le != cudaSuccess) return check_cuda(le, "flash_attn2_kernel launch");
  TTB_CHECK_LAUNCH("flash_attn2_kernel");
  return 0;
}

}  // namespace ttb
