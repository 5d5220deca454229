// Attention kernels, head_dim 64.
//  * attn_simt_kernel: general softmax(q k^T * scale + relpos_bias [+ causal]) v, online-softmax (flash) form,
//    one query per thread, K/V tiles staged in shared memory as fp32. Used for prompt prefill, latents, CLVP and
//    diffusion attention (until the tcgen05 path in flash_attn.cu takes the large shapes).
//  * ar_decode_attn_kernel: single-query attention of every (candidate, head) over [shared prefix | own KV]
//    for the GPT-2 sampler; memory-bound on the KV stream (HBM roofline), appends the new K/V.
#include "common.cuh"
#include "ttb_internal.h"

namespace ttb {

constexpr int ATT_Q = 128;   // queries per block (one per thread)
constexpr int ATT_KT = 32;   // keys per tile

__global__ void __launch_bounds__(ATT_Q)
attn_simt_kernel(TtbAttnArgs a) {
  __shared__ __align__(16) float sk[ATT_KT][64];
  __shared__ __align__(16) float sv[ATT_KT][64];
  const int h = blockIdx.y, seq = blockIdx.z;
  const int q0 = blockIdx.x * ATT_Q;
  const int qi = q0 + threadIdx.x;
  const bool active = qi < a.T;
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a.qkv) + (long long)seq * a.T * a.ld + h * 64;
  float q[64], o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
  if (active) {
    const uint4* qp = reinterpret_cast<const uint4*>(base + (long long)qi * a.ld);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = qp[i];
      float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
      q[8 * i] = f0.x * a.scale; q[8 * i + 1] = f0.y * a.scale; q[8 * i + 2] = f1.x * a.scale; q[8 * i + 3] = f1.y * a.scale;
      q[8 * i + 4] = f2.x * a.scale; q[8 * i + 5] = f2.y * a.scale; q[8 * i + 6] = f3.x * a.scale; q[8 * i + 7] = f3.y * a.scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < 64; ++d) q[d] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  const int kend = a.causal ? min(a.T, q0 + ATT_Q) : a.T;
  const float* bias = a.bias ? a.bias + (long long)h * (2 * a.T - 1) + (a.T - 1) : nullptr;
  for (int k0 = 0; k0 < kend; k0 += ATT_KT) {
    __syncthreads();
    // cooperative load of K and V tiles: ATT_KT rows x 64 dims, 8 bf16 per uint4 -> 8 uint4 per row
    for (int i = threadIdx.x; i < ATT_KT * 8; i += ATT_Q) {
      const int r = i >> 3, c = (i & 7) * 8;
      const int kj = k0 + r;
      uint4 uk = make_uint4(0, 0, 0, 0), uv = make_uint4(0, 0, 0, 0);
      if (kj < a.T) {
        uk = *reinterpret_cast<const uint4*>(base + (long long)kj * a.ld + a.k_off + c);
        uv = *reinterpret_cast<const uint4*>(base + (long long)kj * a.ld + a.v_off + c);
      }
      float2 t;
      t = unpack_bf16(uk.x); sk[r][c] = t.x; sk[r][c + 1] = t.y;
      t = unpack_bf16(uk.y); sk[r][c + 2] = t.x; sk[r][c + 3] = t.y;
      t = unpack_bf16(uk.z); sk[r][c + 4] = t.x; sk[r][c + 5] = t.y;
      t = unpack_bf16(uk.w); sk[r][c + 6] = t.x; sk[r][c + 7] = t.y;
      t = unpack_bf16(uv.x); sv[r][c] = t.x; sv[r][c + 1] = t.y;
      t = unpack_bf16(uv.y); sv[r][c + 2] = t.x; sv[r][c + 3] = t.y;
      t = unpack_bf16(uv.z); sv[r][c + 4] = t.x; sv[r][c + 5] = t.y;
      t = unpack_bf16(uv.w); sv[r][c + 6] = t.x; sv[r][c + 7] = t.y;
    }
    __syncthreads();
    if (!active) continue;
    float s[ATT_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < ATT_KT; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(&sk[j][0]);
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        float4 kk = kr[d];
        acc += q[4 * d] * kk.x + q[4 * d + 1] * kk.y + q[4 * d + 2] * kk.z + q[4 * d + 3] * kk.w;
      }
      const int kj = k0 + j;
      if (bias && kj < a.T) acc += __ldg(bias + (kj - qi));
      const bool ok = (kj < a.T) && (!a.causal || kj <= qi);
      s[j] = ok ? acc : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mnew = fmaxf(mrun, tmax);
    if (mnew == -INFINITY) continue;
    const float corr = __expf(mrun - mnew);
    lrun *= corr;
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < ATT_KT; ++j) {
      const float p = __expf(s[j] - mnew);
      lrun += p;
      const float4* vr = reinterpret_cast<const float4*>(&sv[j][0]);
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        float4 vv = vr[d];
        o[4 * d] += p * vv.x; o[4 * d + 1] += p * vv.y; o[4 * d + 2] += p * vv.z; o[4 * d + 3] += p * vv.w;
      }
    }
    mrun = mnew;
  }
  if (active) {
    const float inv = 1.0f / lrun;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((long long)seq * a.T + qi) * a.ldo + h * 64;
    uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o4[i] = make_uint4(pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                         pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
  }
}

// ------------------------------------------------------------------ AR decode attention
// grid (H, B), 128 threads. ctx = P + step + 1 (the new token included).
constexpr int DEC_THREADS = 128;

__global__ void __launch_bounds__(DEC_THREADS)
ar_decode_attn_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ pk,
                      const __nv_bfloat16* __restrict__ pv, __nv_bfloat16* __restrict__ ck,
                      __nv_bfloat16* __restrict__ cv, const TtbArState* __restrict__ state, int H, int P, int Nmax,
                      __nv_bfloat16* __restrict__ out) {
  extern __shared__ float dsm[];  // scores [P + Nmax]
  __shared__ float sq[64];
  __shared__ float red[32];
  __shared__ float part[DEC_THREADS / 32][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * 64;
  // decode step t (>= 1) feeds token t-1 whose K/V land in candidate slot t-1
  const int slot = state->step - 1;
  const int nc = slot + 1;            // candidate entries incl. the new one
  const int ctx = P + nc;
  const __nv_bfloat16* row = qkv + (long long)b * 3 * D + h * 64;
  __nv_bfloat16* ckb = ck + ((long long)b * H + h) * Nmax * 64;
  __nv_bfloat16* cvb = cv + ((long long)b * H + h) * Nmax * 64;
  const __nv_bfloat16* pkb = pk + (long long)h * P * 64;
  const __nv_bfloat16* pvb = pv + (long long)h * P * 64;
  if (threadIdx.x < 64) {
    sq[threadIdx.x] = __bfloat162float(row[threadIdx.x]) * 0.125f;
    ckb[(long long)slot * 64 + threadIdx.x] = row[D + threadIdx.x];
    cvb[(long long)slot * 64 + threadIdx.x] = row[2 * D + threadIdx.x];
  }
  __syncthreads();
  float q[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) q[d] = sq[d];
  float lmax = -INFINITY;
  for (int j = threadIdx.x; j < ctx; j += DEC_THREADS) {
    const __nv_bfloat16* kr = (j < P) ? pkb + (long long)j * 64 : ckb + (long long)(j - P) * 64;
    const uint4* k4 = reinterpret_cast<const uint4*>(kr);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = k4[i];
      float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
      acc += q[8 * i] * f0.x + q[8 * i + 1] * f0.y + q[8 * i + 2] * f1.x + q[8 * i + 3] * f1.y + q[8 * i + 4] * f2.x +
             q[8 * i + 5] * f2.y + q[8 * i + 6] * f3.x + q[8 * i + 7] * f3.y;
    }
    dsm[j] = acc;
    lmax = fmaxf(lmax, acc);
  }
  const float m = block_max(lmax, red);
  float lsum = 0.f;
  for (int j = threadIdx.x; j < ctx; j += DEC_THREADS) {
    float p = __expf(dsm[j] - m);
    dsm[j] = p;
    lsum += p;
  }
  const float denom = block_sum(lsum, red);  // contains __syncthreads -> dsm visible
  // PV: 16-byte loads; a warp covers 4 positions x 8 dim-chunks per instruction, 4 instructions in flight.
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int psub = lane >> 3, dch = lane & 7;       // position within the group of 4, 8-dim chunk
  float acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;
  constexpr int NW = DEC_THREADS / 32;
  for (int j0 = w * 16; j0 < ctx; j0 += NW * 16) {
    uint4 u[4];
    float pr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i * 4 + psub;
      const bool ok = j < ctx;
      const __nv_bfloat16* vr = (j < P) ? pvb + (long long)j * 64 : cvb + (long long)(j - P) * 64;
      u[i] = ok ? reinterpret_cast<const uint4*>(vr)[dch] : make_uint4(0, 0, 0, 0);
      pr[i] = ok ? dsm[j] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f0 = unpack_bf16(u[i].x), f1 = unpack_bf16(u[i].y), f2 = unpack_bf16(u[i].z), f3 = unpack_bf16(u[i].w);
      acc[0] += pr[i] * f0.x; acc[1] += pr[i] * f0.y; acc[2] += pr[i] * f1.x; acc[3] += pr[i] * f1.y;
      acc[4] += pr[i] * f2.x; acc[5] += pr[i] * f2.y; acc[6] += pr[i] * f3.x; acc[7] += pr[i] * f3.y;
    }
  }
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], 8);
    acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], 16);
  }
  if (psub == 0) {
#pragma unroll
    for (int d = 0; d < 8; ++d) part[w][dch * 8 + d] = acc[d];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += part[i][threadIdx.x];
    out[(long long)b * D + h * 64 + threadIdx.x] = __float2bfloat16(s / denom);
  }
}

__global__ void ar_store_prefix_kernel(const __nv_bfloat16* __restrict__ qkv, int P, int H, __nv_bfloat16* __restrict__ pk,
                                       __nv_bfloat16* __restrict__ pv) {
  const int D = H * 64;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over P * D
  if (i >= (long long)P * D) return;
  const int p = (int)(i / D), c = (int)(i - (long long)p * D);
  const int h = c >> 6, d = c & 63;
  pk[((long long)h * P + p) * 64 + d] = qkv[(long long)p * 3 * D + D + c];
  pv[((long long)h * P + p) * 64 + d] = qkv[(long long)p * 3 * D + 2 * D + c];
}

}  // namespace ttb
using namespace ttb;

extern "C" int ttb_attention(const TtbAttnArgs* ap, void* stream) {
  const TtbAttnArgs& a = *ap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a.T <= 0 || a.nseq <= 0) return 0;
  if ((a.ld & 7) || (a.ldo & 7) || (a.k_off & 7) || (a.v_off & 7)) { set_error("ttb_attention: strides must be multiples of 8"); return -1; }
  if (flash_attention_supported(a)) return flash_attention_launch(a, st);
  dim3 grid((a.T + ATT_Q - 1) / ATT_Q, a.H, a.nseq);
  attn_simt_kernel<<<grid, ATT_Q, 0, st>>>(a);
  TTB_CHECK_LAUNCH("attn_simt_kernel");
  return 0;
}

extern "C" int ttb_ar_decode_attention(const void* qkv, const void* prefix_k, const void* prefix_v, void* cand_k,
                                       void* cand_v, const TtbArState* state, int B, int H, int P, int Nmax, void* out,
                                       void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)(P + Nmax) * sizeof(float);
  if (smem > 40 * 1024) { set_error("ttb_ar_decode_attention: context %d too long", P + Nmax); return -1; }
  dim3 grid(H, B);
  ar_decode_attn_kernel<<<grid, DEC_THREADS, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(prefix_k),
      reinterpret_cast<const __nv_bfloat16*>(prefix_v), reinterpret_cast<__nv_bfloat16*>(cand_k),
      reinterpret_cast<__nv_bfloat16*>(cand_v), state, H, P, Nmax, reinterpret_cast<__nv_bfloat16*>(out));
  TTB_CHECK_LAUNCH("ar_decode_attn_kernel");
  return 0;
}

extern "C" int ttb_ar_store_prefix(const void* qkv, int P, int H, void* prefix_k, void* prefix_v, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)P * H * 64;
  ar_store_prefix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), P, H,
                                                                     reinterpret_cast<__nv_bfloat16*>(prefix_k),
                                                                     reinterpret_cast<__nv_bfloat16*>(prefix_v));
  TTB_CHECK_LAUNCH("ar_store_prefix_kernel");
  return 0;
}
