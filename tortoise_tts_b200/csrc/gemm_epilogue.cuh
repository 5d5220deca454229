// GEMM epilogue shared by the tcgen05 GEMM kernels (included by gemm.cu after GemmEpilogue is defined).
//
// tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns). Writing that row straight to global
// memory makes each warp-wide store touch 32 different cache lines, so the 32x32 chunk goes through a padded
// shared-memory tile: 8 lanes then cover one row's 128 bytes and a warp instruction moves 4 complete rows, for the
// residual read and the fp32 / bf16 writes alike.
//
// Measured with ttb_debug_gemm_trace (profiles/gemm_trace_r01.txt): the first version of this epilogue -- one
// function with every activation and every ragged-edge case behind run-time branches, inside the chunk loop -- took
// ~2.3 us per 32-column chunk whatever it stored (9-10 us per 128x128 tile against a 5 us mainloop): the executed
// path was spread over ~39 KB of code, and residual loads serialised behind possibly-aliasing stores. Hence:
//   * the activation is a template parameter and the run-time switch sits OUTSIDE the chunk loop;
//   * a chunk that is complete (32 rows, 32 columns, 16-byte aligned rows) takes a branch-free FAST path of ~100
//     instructions; ragged edges take a separate, simple per-row SLOW path;
//   * the residual (which may alias the output: x += f(x) in place) is loaded for the whole chunk before any store;
//   * the transpose tile uses a 36-float pitch so both sides are conflict-free 128-bit shared accesses.
#pragma once

namespace ttb {

constexpr int EPI_PITCH = 36;                           // floats per scratch row (144 B: 16-B aligned, conflict-free v4)
constexpr int EPI_SCRATCH_BYTES = 32 * EPI_PITCH * 4;   // per epilogue warp

TTB_DEVINL void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
TTB_DEVINL float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// tcgen05.wait::ld that also carries a data dependency on the loaded registers, so that no use of r[] can be scheduled
// above the wait.
TTB_DEVINL void tmem_ld_wait_dep(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

template <int ACT>
TTB_DEVINL float epi_act(float x) {
  if (ACT == TTB_ACT_GELU_NEW) return gelu_new(x);
  if (ACT == TTB_ACT_SILU) return silu(x);
  if (ACT == TTB_ACT_LRELU02) return leaky(x, 0.2f);
  if (ACT == TTB_ACT_TANH) return tanhf(x);
  return x;
}

// What a tile's epilogue needs to know once (warp-uniform).
struct EpiFlags {
  bool aligned;     // bias pointer, ldr, ldo, ldob all allow 128-bit (fp32) / 64-bit (bf16) row accesses
};

TTB_DEVINL EpiFlags epi_flags(const GemmEpilogue& ep) {
  EpiFlags f;
  f.aligned = (!ep.bias || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) &&
              (!ep.residual || ((ep.ldr & 3) == 0 && (ep.res_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0)) &&
              (!ep.out_f32 || ((ep.ldo & 3) == 0 && (ep.outf_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0)) &&
              (!ep.out_bf16 || ((ep.ldob & 3) == 0 && (ep.outb_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out_bf16) & 7) == 0));
  return f;
}

// (sum, sum of squares) of one 32-row x 32-column block of the output = one GroupNorm partial (32 channels per group):
// fixed xor tree, one writer, no atomics (bit-reproducible).
TTB_DEVINL void epi_gn_store(float gs, float gq, int nb, int m_base, int lane, long long bz, const GemmEpilogue& ep) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    gs += __shfl_xor_sync(0xffffffffu, gs, off);
    gq += __shfl_xor_sync(0xffffffffu, gq, off);
  }
  if (lane == 0)
    ep.gn_part[((long long)bz * ep.gn_groups + (nb >> 5)) * TTB_GN_SPLITS + (m_base >> 5)] = make_float2(gs, gq);
}

// ---- FAST path: all 32 columns exist, everything aligned. Branch-free apart from the uniform pointer tests. Only the
// first `rows_valid` (1..32) of the 32 rows exist (< 32: the last row block of a ragged M): the row-wise residual loads
// and stores are predicated per row -- one code path, so the kernels' register allocation and code size stay as they were. (The per-lane scalar SLOW path costs ~18 us for such a block -- profiles/gemm_trace_r01_fine.txt,
// "SM 77" -- and with M = 1872 = 14.6 tiles it sat on the tail of every denoiser GEMM.)
template <int ACT>
TTB_DEVINL void epi_chunk_fast(const uint32_t* r, int nb, int m_base, int rows_valid, int lane, long long bz,
                               const GemmEpilogue& ep, uint32_t scratch) {
  const uint32_t wrow = scratch + (uint32_t)lane * (EPI_PITCH * 4);
  const bool GN = ep.gn_part != nullptr;
  float v[32];
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + nb) + j);
      v[4 * j] = fmaf(__uint_as_float(r[4 * j]), ep.alpha, b.x);
      v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), ep.alpha, b.y);
      v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), ep.alpha, b.z);
      v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), ep.alpha, b.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
  }
  if (ACT == TTB_ACT_GEGLU) {
    // columns interleaved (u0,g0,u1,g1,...): out[j] = u * gelu_erf(g); output width N/2 -> 16 columns per chunk
#pragma unroll
    for (int k = 0; k < 4; ++k)
      sts128(wrow + k * 16, v[8 * k] * gelu_erf(v[8 * k + 1]), v[8 * k + 2] * gelu_erf(v[8 * k + 3]),
             v[8 * k + 4] * gelu_erf(v[8 * k + 5]), v[8 * k + 6] * gelu_erf(v[8 * k + 7]));
    __syncwarp();
    const int rr0 = lane >> 2, c = (lane & 3) * 4;   // 8 rows per instruction: 4 lanes x 4 columns per row
    const long long col = (nb >> 1) + c;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + rr0;
      const float4 o = lds128(scratch + (uint32_t)(rr * EPI_PITCH + c) * 4);
      const long long m = m_base + rr;
      if (rr >= rows_valid) continue;
      if (ep.out_bf16)
        *reinterpret_cast<uint2*>(ep.out_bf16 + bz * ep.outb_bstride + m * ep.ldob + col) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + bz * ep.outf_bstride + m * ep.ldo + col) = o;
    }
    __syncwarp();                                    // scratch is reused by the next chunk
    return;
  }
  const int rr0 = lane >> 3, c = (lane & 7) * 4;     // 4 rows per instruction: 8 lanes x 4 columns per row
  const long long col = nb + c;
  // residual for the positions this lane stores, all issued before the first store
  float4 rs[8];
  if (ep.residual) {
    const float* rp = ep.residual + bz * ep.res_bstride + (long long)(m_base + rr0) * ep.ldr + col;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (it * 4 + rr0 < rows_valid) rs[it] = *reinterpret_cast<const float4*>(rp + (long long)it * 4 * ep.ldr);
      else rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    sts128(wrow + k * 16, epi_act<ACT>(v[4 * k]), epi_act<ACT>(v[4 * k + 1]), epi_act<ACT>(v[4 * k + 2]),
           epi_act<ACT>(v[4 * k + 3]));
  __syncwarp();
  float gs = 0.f, gq = 0.f;                          // GroupNorm partial of this 32 x 32 block (ep.gn_part)
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rr0;
    float4 o = lds128(scratch + (uint32_t)(rr * EPI_PITCH + c) * 4);
    if (ep.residual) { o.x += rs[it].x; o.y += rs[it].y; o.z += rs[it].z; o.w += rs[it].w; }
    const long long m = m_base + rr;
    if (rr >= rows_valid) continue;
    if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + bz * ep.outf_bstride + m * ep.ldo + col) = o;
    if (ep.out_bf16)
      *reinterpret_cast<uint2*>(ep.out_bf16 + bz * ep.outb_bstride + m * ep.ldob + col) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    if (GN) {
      gs += (o.x + o.y) + (o.z + o.w);
      gq += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
    }
  }
  if (GN) epi_gn_store(gs, gq, nb, m_base, lane, bz, ep);
  __syncwarp();                                      // scratch is reused by the next chunk
}

// ---- SLOW path (ragged M / N edges, unaligned rows): each lane finishes its own accumulator row with scalar accesses.
template <int ACT>
TTB_DEVINL void epi_chunk_slow(const uint32_t* r, int nb, int N, int m, int M, long long bz, const GemmEpilogue& ep,
                               float& gs, float& gq) {
  if (m >= M) return;
  if (ACT == TTB_ACT_GEGLU) {
    const int nout = N >> 1, ob = nb >> 1;
    float* of = ep.out_f32 ? ep.out_f32 + bz * ep.outf_bstride + (long long)m * ep.ldo : nullptr;
    __nv_bfloat16* obf = ep.out_bf16 ? ep.out_bf16 + bz * ep.outb_bstride + (long long)m * ep.ldob : nullptr;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (ob + j < nout) {
        float u = __uint_as_float(r[2 * j]) * ep.alpha, g = __uint_as_float(r[2 * j + 1]) * ep.alpha;
        if (ep.bias) { u += __ldg(ep.bias + nb + 2 * j); g += __ldg(ep.bias + nb + 2 * j + 1); }
        const float o = u * gelu_erf(g);
        if (of) of[ob + j] = o;
        if (obf) obf[ob + j] = __float2bfloat16(o);
      }
    }
    return;
  }
  const float* rp = ep.residual ? ep.residual + bz * ep.res_bstride + (long long)m * ep.ldr : nullptr;
  float* of = ep.out_f32 ? ep.out_f32 + bz * ep.outf_bstride + (long long)m * ep.ldo : nullptr;
  __nv_bfloat16* obf = ep.out_bf16 ? ep.out_bf16 + bz * ep.outb_bstride + (long long)m * ep.ldob : nullptr;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (nb + j < N) {
      float x = __uint_as_float(r[j]) * ep.alpha;
      if (ep.bias) x += __ldg(ep.bias + nb + j);
      x = epi_act<ACT>(x);
      if (rp) x += rp[nb + j];
      if (of) of[nb + j] = x;
      if (obf) obf[nb + j] = __float2bfloat16(x);
      gs += x;
      gq += x * x;
    }
  }
}

// Whole accumulator tile of one epilogue warp: taddr = TMEM address of (lane group, first column). When release_bar
// is given it is arrived on once the last tcgen05.ld has completed (persistent kernel: the MMA warp may then overwrite
// this accumulator buffer while the stores are still going out).
template <int BN, int ACT>
TTB_DEVINL void gemm_epilogue_tile(uint32_t taddr, int n0, int N, int m_base, int M, int lane, long long bz,
                                   const GemmEpilogue& ep, uint32_t scratch, uint64_t* release_bar) {
  constexpr int NCH = BN / 32;
  const bool aligned = epi_flags(ep).aligned;
  const int rows_valid = min(32, M - m_base);        // warp-uniform; <= 0: this warp's rows are all past M
  // rolled on purpose: one chunk's worth of code and registers (the 128-wide kernel must stay <= 168 registers for
  // two CTAs per SM); the co-resident warps cover the tcgen05.ld latency
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    const int nb = n0 + ch * 32;
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + (uint32_t)ch * 32, r);
    tmem_ld_wait_dep(r);
    if (release_bar && ch == NCH - 1) { tc_fence_before(); mbar_arrive(release_bar); }
    if (nb >= N || rows_valid <= 0) continue;        // warp-uniform
    if (aligned && nb + 32 <= N) {
      epi_chunk_fast<ACT>(r, nb, m_base, rows_valid, lane, bz, ep, scratch);
    } else {
      float gs = 0.f, gq = 0.f;                      // rows past M contribute nothing
      epi_chunk_slow<ACT>(r, nb, N, m_base + lane, M, bz, ep, gs, gq);
      if (ep.gn_part) epi_gn_store(gs, gq, nb, m_base, lane, bz, ep);
    }
  }
}

template <int BN>
TTB_DEVINL void gemm_epilogue_dispatch(uint32_t taddr, int n0, int N, int m_base, int M, int lane, long long bz,
                                       const GemmEpilogue& ep, uint32_t scratch, uint64_t* release_bar) {
  switch (ep.act) {
    case TTB_ACT_GEGLU: gemm_epilogue_tile<BN, TTB_ACT_GEGLU>(taddr, n0, N, m_base, M, lane, bz, ep, scratch, release_bar); break;
    case TTB_ACT_GELU_NEW: gemm_epilogue_tile<BN, TTB_ACT_GELU_NEW>(taddr, n0, N, m_base, M, lane, bz, ep, scratch, release_bar); break;
    case TTB_ACT_SILU: gemm_epilogue_tile<BN, TTB_ACT_SILU>(taddr, n0, N, m_base, M, lane, bz, ep, scratch, release_bar); break;
    case TTB_ACT_LRELU02: gemm_epilogue_tile<BN, TTB_ACT_LRELU02>(taddr, n0, N, m_base, M, lane, bz, ep, scratch, release_bar); break;
    case TTB_ACT_TANH: gemm_epilogue_tile<BN, TTB_ACT_TANH>(taddr, n0, N, m_base, M, lane, bz, ep, scratch, release_bar); break;
    default: gemm_epilogue_tile<BN, TTB_ACT_NONE>(taddr, n0, N, m_base, M, lane, bz, ep, scratch, release_bar); break;
  }
}

}  // namespace ttb
