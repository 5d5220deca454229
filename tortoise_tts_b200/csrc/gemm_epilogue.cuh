// Coalesced GEMM epilogue (included by gemm.cu after GemmEpilogue is defined).
//
// tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns). Writing that row straight to global
// memory makes each warp-wide store touch 32 different cache lines with 16 bytes each: the LSU serialises them
// (8x more transactions than needed) and the epilogue, not the tensor pipe, bounds every K<=1024 GEMM on the path
// (ncu: tensor pipe 14-20 % active, long-scoreboard stalls; profiles/ncu_summary_r01_run10.txt).
// Here the 32x32 chunk goes through a padded shared-memory tile so that 8 lanes cover one row's 128 bytes: a warp
// instruction reads/writes 4 complete rows (4 x 128 B lines), for the residual read and the fp32 / bf16 writes alike.
#pragma once

namespace ttb {

constexpr int EPI_PITCH = 33;                       // floats per scratch row (bank-conflict-free column writes)
constexpr int EPI_SCRATCH_BYTES = 32 * EPI_PITCH * 4;   // per epilogue warp

// r: 32 accumulator columns [nb, nb+32) of row (m_base + lane). scratch: this warp's [32][33] float tile.
TTB_DEVINL void gemm_epilogue_coalesced(const uint32_t* r, int nb, int N, int m_base, int M, int lane, long long bz,
                                        const GemmEpilogue& ep, float* scratch) {
  if (nb >= N) return;                               // warp-uniform
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = __uint_as_float(r[j]) * ep.alpha;
    if (ep.bias && nb + j < N) x += __ldg(ep.bias + nb + j);
    v[j] = x;
  }
  if (ep.act == TTB_ACT_GEGLU) {
    // columns interleaved (u0,g0,u1,g1,...): out[j] = u * gelu_erf(g); output width N/2 -> 16 columns per chunk
#pragma unroll
    for (int j = 0; j < 16; ++j) scratch[lane * EPI_PITCH + j] = v[2 * j] * gelu_erf(v[2 * j + 1]);
    __syncwarp();
    const int ob = nb >> 1, nout = N >> 1;
#pragma unroll
    for (int it = 0; it < 4; ++it) {                 // 8 rows per instruction: 4 lanes x 4 columns per row
      const int rr = it * 8 + (lane >> 2), c = (lane & 3) * 4;
      const int m = m_base + rr, n = ob + c;
      if (m < M && n < nout) {
        const float* s = scratch + rr * EPI_PITCH + c;
        const float o0 = s[0], o1 = s[1], o2 = s[2], o3 = s[3];
        if (ep.out_bf16) {
          __nv_bfloat16* p = ep.out_bf16 + bz * ep.outb_bstride + (long long)m * ep.ldob + n;
          if (n + 4 <= nout && (ep.ldob & 3) == 0) *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
          else { const float o[4] = {o0, o1, o2, o3}; for (int j = 0; j < 4 && n + j < nout; ++j) p[j] = __float2bfloat16(o[j]); }
        }
        if (ep.out_f32) {
          float* p = ep.out_f32 + bz * ep.outf_bstride + (long long)m * ep.ldo + n;
          const float o[4] = {o0, o1, o2, o3};
          for (int j = 0; j < 4 && n + j < nout; ++j) p[j] = o[j];
        }
      }
    }
    __syncwarp();
    return;
  }
  if (ep.act == TTB_ACT_GELU_NEW) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_new(v[j]);
  } else if (ep.act == TTB_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
  } else if (ep.act == TTB_ACT_LRELU02) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = leaky(v[j], 0.2f);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) scratch[lane * EPI_PITCH + j] = v[j];
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {                   // 4 rows per instruction: 8 lanes x 4 columns per row
    const int rr = it * 4 + (lane >> 3), c = (lane & 7) * 4;
    const int m = m_base + rr, n = nb + c;
    if (m < M && n < N) {
      const float* s = scratch + rr * EPI_PITCH + c;
      float o[4] = {s[0], s[1], s[2], s[3]};
      const bool full4 = (n + 4 <= N);
      if (ep.residual) {
        const float* p = ep.residual + bz * ep.res_bstride + (long long)m * ep.ldr + n;
        if (full4 && (ep.ldr & 3) == 0) {
          const float4 t = *reinterpret_cast<const float4*>(p);
          o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
        } else {
          for (int j = 0; j < 4 && n + j < N; ++j) o[j] += p[j];
        }
      }
      if (ep.out_f32) {
        float* p = ep.out_f32 + bz * ep.outf_bstride + (long long)m * ep.ldo + n;
        if (full4 && (ep.ldo & 3) == 0) *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
        else for (int j = 0; j < 4 && n + j < N; ++j) p[j] = o[j];
      }
      if (ep.out_bf16) {
        __nv_bfloat16* p = ep.out_bf16 + bz * ep.outb_bstride + (long long)m * ep.ldob + n;
        if (full4 && (ep.ldob & 3) == 0) *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
        else for (int j = 0; j < 4 && n + j < N; ++j) p[j] = __float2bfloat16(o[j]);
      }
    }
  }
  __syncwarp();                                      // scratch is reused by the next chunk
}

}  // namespace ttb
