// UnivNet vocoder kernels (models/vocoder.py). Activations are channel-major fp32 [C, L] as in the reference;
// every kernel is HBM/L2-bound or small-channel SIMT work (C = 32/64), so there is no tensor-core path here
// except the kernel-predictor GEMM (ttb_gemm), whose output this file's LVC kernel consumes directly.
#include "common.cuh"
#include "ttb_internal.h"

namespace ttb {

// ------------------------------------------------------------------ small-channel Conv1d
// thread: VT consecutive time steps x VCO output channels. grid (ceil(L / (128*VT)), ceil(Cout / VCO)).
constexpr int VT = 4;
constexpr int VCO = 8;

__global__ void __launch_bounds__(128)
voc_conv1d_kernel(const float* __restrict__ x, int Cin, int L, const float* __restrict__ w, const float* __restrict__ b,
                  int Cout, int ks, int dil, int reflect, float lrelu_in, float lrelu_out, int tanh_out,
                  const float* __restrict__ residual, float* __restrict__ out) {
  extern __shared__ float sw[];  // [Cin][ks][VCO]
  const int co0 = blockIdx.y * VCO;
  for (int i = threadIdx.x; i < Cin * ks * VCO; i += 128) {
    const int c = i % VCO, r = i / VCO;  // r = ci*ks + k
    const int co = co0 + c;
    sw[i] = (co < Cout) ? w[(long long)co * Cin * ks + r] : 0.f;
  }
  __syncthreads();
  const int t0 = (blockIdx.x * 128 + threadIdx.x) * VT;
  if (t0 >= L) return;
  float acc[VT][VCO];
#pragma unroll
  for (int i = 0; i < VT; ++i)
#pragma unroll
    for (int c = 0; c < VCO; ++c) acc[i][c] = 0.f;
  const int half = ks / 2;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xr = x + (long long)ci * L;
    for (int k = 0; k < ks; ++k) {
      float xv[VT];
#pragma unroll
      for (int i = 0; i < VT; ++i) {
        int t = t0 + i + (k - half) * dil;
        float v = 0.f;
        if (reflect) {
          if (t < 0) t = -t;
          if (t >= L) t = 2 * (L - 1) - t;
          v = (t0 + i < L) ? xr[t] : 0.f;
        } else if (t >= 0 && t < L) {
          v = xr[t];
        }
        xv[i] = (lrelu_in != 1.0f) ? leaky(v, lrelu_in) : v;
      }
      const float4* wp = reinterpret_cast<const float4*>(sw + (ci * ks + k) * VCO);
      const float4 w0 = wp[0], w1 = wp[1];
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < VT; ++i)
#pragma unroll
        for (int c = 0; c < VCO; ++c) acc[i][c] += xv[i] * ww[c];
    }
  }
#pragma unroll
  for (int c = 0; c < VCO; ++c) {
    const int co = co0 + c;
    if (co >= Cout) break;
    const float bb = b ? b[co] : 0.f;
#pragma unroll
    for (int i = 0; i < VT; ++i) {
      const int t = t0 + i;
      if (t >= L) break;
      float v = acc[i][c] + bb;
      if (lrelu_out != 1.0f) v = leaky(v, lrelu_out);
      if (tanh_out) v = tanhf(v);
      if (residual) v += residual[(long long)co * L + t];
      out[(long long)co * L + t] = v;
    }
  }
}

// ------------------------------------------------------------------ LeakyReLU + ConvTranspose1d (k = 2*stride)
// out[co, t] = b[co] + sum_ci sum_{j in {jhi, jhi-1}} lrelu(x[ci, j]) * w[ci, co, t + p - j*stride]
__global__ void __launch_bounds__(128)
voc_convt_kernel(const float* __restrict__ x, int C, int L, const float* __restrict__ w, const float* __restrict__ b,
                 int stride, float lrelu_in, float* __restrict__ out) {
  extern __shared__ float sw[];  // [C][C][2*stride]
  const int ks = 2 * stride;
  for (int i = threadIdx.x; i < C * C * ks; i += 128) sw[i] = w[i];
  __syncthreads();
  const int Lo = L * stride;
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= Lo) return;
  const int p = stride / 2 + stride % 2;
  const int jhi = (t + p) / stride;
  const int khi = t + p - jhi * stride;      // in [0, stride)
  const int jlo = jhi - 1;
  const int klo = khi + stride;              // in [stride, 2*stride)
  const bool vhi = jhi < L, vlo = jlo >= 0;
  for (int co0 = 0; co0 < C; co0 += 16) {
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    for (int ci = 0; ci < C; ++ci) {
      const float xh = vhi ? leaky(x[(long long)ci * L + jhi], lrelu_in) : 0.f;
      const float xl = vlo ? leaky(x[(long long)ci * L + jlo], lrelu_in) : 0.f;
      const float* wr = sw + ((long long)ci * C + co0) * ks;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] += xh * wr[c * ks + khi] + xl * wr[c * ks + klo];
    }
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (co0 + c < C) out[(long long)(co0 + c) * Lo + t] = acc[c] + b[co0 + c];
  }
}

// ------------------------------------------------------------------ LVC + gated activation (fused)
// block: one (frame f, chunk) pair; 256 threads = 64 output channels x 4 sample groups of NS samples.
template <int NS>
__global__ void __launch_bounds__(256)
voc_lvc_gate_kernel(const float* __restrict__ y, int C, int L, int hop, const float* __restrict__ kernels, int ldk,
                    int koff, const float* __restrict__ bias, int ldb, int boff, float* __restrict__ x) {
  constexpr int CH = 4 * NS;               // samples per block
  __shared__ float ytile[32][CH + 2 + 2];  // y[-1 .. CH], padded
  __shared__ float otile[64][CH + 1];
  const int chunks = hop / CH;
  const int f = blockIdx.x / chunks, ch = blockIdx.x - f * chunks;
  const int s0 = f * hop + ch * CH;        // first global sample of this block
  for (int i = threadIdx.x; i < 32 * (CH + 2); i += 256) {
    const int c = i / (CH + 2), j = i - c * (CH + 2);
    const int t = s0 - 1 + j;
    ytile[c][j] = (t >= 0 && t < L) ? y[(long long)c * L + t] : 0.f;
  }
  __syncthreads();
  const int oc = threadIdx.x & 63, sg = threadIdx.x >> 6;
  const float* kf = kernels + (long long)f * ldk + koff + oc;  // [i][k][oc]
  float acc[NS];
  const float bv = bias[(long long)f * ldb + boff + oc];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = bv;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    float yr[NS + 2];
#pragma unroll
    for (int s = 0; s < NS + 2; ++s) yr[s] = ytile[i][sg * NS + s];
    const float k0 = __ldg(kf + (i * 3 + 0) * 64), k1 = __ldg(kf + (i * 3 + 1) * 64), k2 = __ldg(kf + (i * 3 + 2) * 64);
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] += yr[s] * k0 + yr[s + 1] * k1 + yr[s + 2] * k2;
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) otile[oc][sg * NS + s] = acc[s];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * CH; i += 256) {
    const int c = i / CH, s = i - c * CH;
    const float a = otile[c][s], g = otile[c + 32][s];
    const float sig = 1.0f / (1.0f + __expf(-a));
    x[(long long)c * L + s0 + s] += sig * tanhf(g);
  }
}

// channel-major fp32 -> token-major bf16. split != 0 writes the error-compensated triple [hi | lo | hi] per row
// (x = hi + lo to ~16 mantissa bits) so that one bf16 GEMM against [Wh | Wh | Wl] reproduces fp32-grade accuracy.
__global__ void voc_to_tokens_kernel(const float* __restrict__ x, int C, int L, __nv_bfloat16* __restrict__ out, int ldo,
                                     int split) {
  __shared__ float tile[32][33];
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && l < L) ? x[(long long)c * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + threadIdx.x;
    if (l >= L) continue;
    const float v = (c < C) ? tile[threadIdx.x][i] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    if (!split) {
      if (c < ldo) out[(long long)l * ldo + c] = hi;
    } else if (c < C) {
      const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
      out[(long long)l * ldo + c] = hi;
      out[(long long)l * ldo + C + c] = lo;
      out[(long long)l * ldo + 2 * C + c] = hi;
    }
  }
}

}  // namespace ttb
using namespace ttb;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int ttb_voc_conv1d(const float* x, int Cin, int L, const float* w, const float* b, int Cout, int ksize,
                              int dilation, int reflect, float lrelu_in, float lrelu_out, int tanh_out,
                              const float* residual, float* out, void* stream) {
  const size_t smem = (size_t)Cin * ksize * VCO * sizeof(float);
  if (smem > 48 * 1024) { set_error("ttb_voc_conv1d: weights slice too large"); return -1; }
  if (reflect && L <= ksize / 2) { set_error("ttb_voc_conv1d: reflect pad needs L > %d", ksize / 2); return -1; }
  dim3 grid((L + 128 * VT - 1) / (128 * VT), (Cout + VCO - 1) / VCO);
  voc_conv1d_kernel<<<grid, 128, smem, ST>>>(x, Cin, L, w, b, Cout, ksize, dilation, reflect, lrelu_in, lrelu_out,
                                             tanh_out, residual, out);
  TTB_CHECK_LAUNCH("voc_conv1d_kernel");
  return 0;
}

extern "C" int ttb_voc_convt(const float* x, int C, int L, const float* w, const float* b, int stride, float lrelu_in,
                             float* out, void* stream) {
  const size_t smem = (size_t)C * C * 2 * stride * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(voc_convt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(voc_convt)");
    attr = true;
  }
  if (smem > 96 * 1024 || (C % 16) != 0) { set_error("ttb_voc_convt: unsupported C=%d stride=%d", C, stride); return -1; }
  const int Lo = L * stride;
  voc_convt_kernel<<<(Lo + 127) / 128, 128, smem, ST>>>(x, C, L, w, b, stride, lrelu_in, out);
  TTB_CHECK_LAUNCH("voc_convt_kernel");
  return 0;
}

extern "C" int ttb_voc_lvc_gate(const float* y, int C, int L, int hop, const float* kernels, int ldk, int koff,
                                const float* bias, int ldb, int boff, float* x, void* stream) {
  if (C != 32) { set_error("ttb_voc_lvc_gate: C must be 32"); return -1; }
  if (L % hop != 0) { set_error("ttb_voc_lvc_gate: L %% hop != 0"); return -1; }
  const int F = L / hop;
  if (hop % 64 == 0) {
    voc_lvc_gate_kernel<16><<<F * (hop / 64), 256, 0, ST>>>(y, C, L, hop, kernels, ldk, koff, bias, ldb, boff, x);
  } else if (hop % 8 == 0) {
    voc_lvc_gate_kernel<2><<<F * (hop / 8), 256, 0, ST>>>(y, C, L, hop, kernels, ldk, koff, bias, ldb, boff, x);
  } else { set_error("ttb_voc_lvc_gate: hop=%d unsupported", hop); return -1; }
  TTB_CHECK_LAUNCH("voc_lvc_gate_kernel");
  return 0;
}

extern "C" int ttb_voc_to_tokens_bf16(const float* x, int C, int L, void* out, int ldo, int split, void* stream) {
  if (split && ldo < 3 * C) { set_error("ttb_voc_to_tokens_bf16: split needs ldo >= 3*C"); return -1; }
  dim3 grid((L + 31) / 32, ((split ? C : ldo) + 31) / 32), block(32, 8);
  voc_to_tokens_kernel<<<grid, block, 0, ST>>>(x, C, L, reinterpret_cast<__nv_bfloat16*>(out), ldo, split);
  TTB_CHECK_LAUNCH("voc_to_tokens_kernel");
  return 0;
}
