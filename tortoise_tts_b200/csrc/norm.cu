// Normalisation kernels (HBM/L2-bound): LayerNorm (single or chained), RMSNorm, GroupNorm32 with the fused
// scale-shift / SiLU consumers. fp32 statistics via warp-shuffle reductions; outputs feed the tcgen05 GEMM as bf16.
#include "common.cuh"
#include "ttb_internal.h"

namespace ttb {

// one block (256 threads) per row; D <= 4096; row cached in registers (up to 16 per thread)
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* x, int D, const float* __restrict__ g1, const float* __restrict__ b1,
                 const float* __restrict__ g2, const float* __restrict__ b2, __nv_bfloat16* __restrict__ ob,
                 float* __restrict__ of, float* xw, const float* __restrict__ partials, int nsplit,
                 long long split_stride, const float* __restrict__ rbias) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * D;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * 256;
    v[i] = (c < D) ? xr[c] : 0.f;
    if (xw && c < D) {          // fused residual update: x += bias + sum of split-K partials (fixed order)
      float t = rbias ? rbias[c] : 0.f;
      for (int sp = 0; sp < nsplit; ++sp) t += partials[sp * split_stride + row * D + c];
      v[i] += t;
      xw[row * D + c] = v[i];
    }
    s += v[i];
  }
  float mean = block_sum(s, red) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * 256;
    float d = (c < D) ? v[i] - mean : 0.f;
    q += d * d;
  }
  float rstd = rsqrtf(block_sum(q, red) / D + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * 256;
    if (c < D) v[i] = (v[i] - mean) * rstd * g1[c] + b1[c];
  }
  if (g2) {
    s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) { int c = threadIdx.x + i * 256; s += (c < D) ? v[i] : 0.f; }
    mean = block_sum(s, red) / D;
    q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = threadIdx.x + i * 256;
      float d = (c < D) ? v[i] - mean : 0.f;
      q += d * d;
    }
    rstd = rsqrtf(block_sum(q, red) / D + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = threadIdx.x + i * 256;
      if (c < D) v[i] = (v[i] - mean) * rstd * g2[c] + b2[c];
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * 256;
    if (c < D) {
      if (ob) ob[row * D + c] = __float2bfloat16(v[i]);
      if (of) of[row * D + c] = v[i];
    }
  }
}

template <int MAXV>
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const float* __restrict__ x, int D, const float* __restrict__ g, __nv_bfloat16* __restrict__ ob) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * D;
  float v[MAXV];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * 256;
    v[i] = (c < D) ? xr[c] : 0.f;
    q += v[i] * v[i];
  }
  float nrm = sqrtf(block_sum(q, red)) * rsqrtf((float)D);
  float inv = 1.0f / fmaxf(nrm, 1e-8f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * 256;
    if (c < D) ob[row * D + c] = __float2bfloat16(v[i] * inv * g[c]);
  }
}

// GroupNorm statistics, token-major x [B, S, C]: grid (groups, splits, B); block 256 threads.
// Each block reduces rows [s0, s1) x channels of one group; partial (sum, sumsq) -> partials[b][g][split][2].
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, int S, int C, int cpg, int splits, float* __restrict__ partials) {
  __shared__ float red[32];
  const int g = blockIdx.x, sp = blockIdx.y, b = blockIdx.z;
  const int rows_per = (S + splits - 1) / splits;
  const int s0 = sp * rows_per, s1 = min(S, s0 + rows_per);
  const float* xb = x + (long long)b * S * C + g * cpg;
  float s = 0.f, q = 0.f;
  if ((cpg & 3) == 0) {
    const int vec = cpg >> 2;  // float4 per row
    const int total = (s1 - s0) * vec;
    for (int i = threadIdx.x; i < total; i += 256) {
      int r = i / vec, c = i - r * vec;
      float4 t = *reinterpret_cast<const float4*>(xb + (long long)(s0 + r) * C + c * 4);
      s += t.x + t.y + t.z + t.w;
      q += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
  } else {
    const int total = (s1 - s0) * cpg;
    for (int i = threadIdx.x; i < total; i += 256) {
      int r = i / cpg, c = i - r * cpg;
      float t = xb[(long long)(s0 + r) * C + c];
      s += t; q += t * t;
    }
  }
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    float* p = partials + (((long long)b * gridDim.x + g) * splits + sp) * 2;
    p[0] = s; p[1] = q;
  }
}

// apply: each thread handles 4 consecutive channels of one token
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, int S, int C, int groups, int cpg, int splits,
                const float* __restrict__ partials, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ ss, int ss_bstride, const int* __restrict__ ss_row, int ss_row_stride,
                int do_silu, __nv_bfloat16* __restrict__ ob, int ldo,
                float* __restrict__ of, int ldof) {
  const int b = blockIdx.y;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;  // over S * C/4
  const int c4 = C >> 2;
  if (idx >= (long long)S * c4) return;
  const int srow = (int)(idx / c4);
  const int c = (int)(idx - (long long)srow * c4) * 4;
  const int g = c / cpg;
  const float* p = partials + ((long long)b * groups + g) * splits * 2;
  float sum = 0.f, sq = 0.f;
  for (int i = 0; i < splits; ++i) { sum += p[2 * i]; sq += p[2 * i + 1]; }
  const float n = (float)S * cpg;
  const float mean = sum / n;
  const float var = fmaxf(sq / n - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  if (ss && ss_row) ss += (long long)(*ss_row) * ss_row_stride;
  float4 t = *reinterpret_cast<const float4*>(x + ((long long)b * S + srow) * C + c);
  float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y = (v[j] - mean) * rstd * gamma[c + j] + beta[c + j];
    if (ss) y = y * (1.0f + ss[(long long)b * ss_bstride + c + j]) + ss[(long long)b * ss_bstride + C + c + j];
    if (do_silu) y = silu(y);
    v[j] = y;
  }
  if (ob) {
    uint2 o = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    *reinterpret_cast<uint2*>(ob + ((long long)b * S + srow) * ldo + c) = o;
  }
  if (of) *reinterpret_cast<float4*>(of + ((long long)b * S + srow) * ldof + c) = make_float4(v[0], v[1], v[2], v[3]);
}

}  // namespace ttb
using namespace ttb;

extern "C" int ttb_layernorm(const float* x, int M, int D, const float* g1, const float* b1, const float* g2,
                             const float* b2, void* out_bf16, float* out_f32, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0) return 0;
  auto ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (D <= 1024) layernorm_kernel<4><<<M, 256, 0, st>>>(x, D, g1, b1, g2, b2, ob, out_f32, nullptr, nullptr, 0, 0, nullptr);
  else if (D <= 4096) layernorm_kernel<16><<<M, 256, 0, st>>>(x, D, g1, b1, g2, b2, ob, out_f32, nullptr, nullptr, 0, 0, nullptr);
  else { set_error("ttb_layernorm: D=%d > 4096", D); return -1; }
  TTB_CHECK_LAUNCH("layernorm_kernel");
  return 0;
}

extern "C" int ttb_residual_layernorm(float* x, int M, int D, const float* partials, int nsplit, long long split_stride,
                                      const float* bias, const float* g1, const float* b1, const float* g2,
                                      const float* b2, void* out_bf16, float* out_f32, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0) return 0;
  auto ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (D <= 1024) layernorm_kernel<4><<<M, 256, 0, st>>>(x, D, g1, b1, g2, b2, ob, out_f32, x, partials, nsplit, split_stride, bias);
  else if (D <= 4096) layernorm_kernel<16><<<M, 256, 0, st>>>(x, D, g1, b1, g2, b2, ob, out_f32, x, partials, nsplit, split_stride, bias);
  else { set_error("ttb_residual_layernorm: D=%d > 4096", D); return -1; }
  TTB_CHECK_LAUNCH("layernorm_kernel(residual)");
  return 0;
}

extern "C" int ttb_rmsnorm(const float* x, int M, int D, const float* g, void* out_bf16, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0) return 0;
  if (D > 1024) { set_error("ttb_rmsnorm: D=%d > 1024", D); return -1; }
  rmsnorm_kernel<4><<<M, 256, 0, st>>>(x, D, g, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  TTB_CHECK_LAUNCH("rmsnorm_kernel");
  return 0;
}

extern "C" int ttb_groupnorm(const float* x, int B, int S, int C, int groups, const float* gamma, const float* beta,
                             const float* scale_shift, int ss_bstride, const int* ss_row, int ss_row_stride,
                             int do_silu, float* partials, void* out_bf16, int ldo, float* out_f32, int ldof,
                             void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C % groups != 0 || (C & 3)) { set_error("ttb_groupnorm: C=%d groups=%d unsupported", C, groups); return -1; }
  const int cpg = C / groups;
  if (cpg % 4 != 0 && (cpg & 3)) { set_error("ttb_groupnorm: channels per group must be a multiple of 4"); return -1; }
  const int splits = TTB_GN_SPLITS;
  dim3 g1(groups, splits, B);
  gn_stats_kernel<<<g1, 256, 0, st>>>(x, S, C, cpg, splits, partials);
  TTB_CHECK_LAUNCH("gn_stats_kernel");
  const long long total = (long long)S * (C >> 2);
  dim3 g2((unsigned)((total + 255) / 256), B);
  gn_apply_kernel<<<g2, 256, 0, st>>>(x, S, C, groups, cpg, splits, partials, gamma, beta, scale_shift, ss_bstride,
                                      ss_row, ss_row_stride, do_silu, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, out_f32, ldof);
  TTB_CHECK_LAUNCH("gn_apply_kernel");
  return 0;
}
