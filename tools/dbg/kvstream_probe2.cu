// Development probe 2 (no product code): the same 226 MB candidate-KV stream as kvstream_probe.cu, layout 0, but staged into
// SHARED memory the way the attention kernel does, to find out which copy engine limits the bytes in flight per SM:
//   mode 1: cp.async.bulk (TMA unit), one 4 KB copy per tile, mbarrier ring of `depth` tiles per warp
//   mode 2: cp.async.cg 16 B per lane (LSU path), `depth` tiles per warp in flight via commit groups
//   mode 3: half the warps use mode 1, the other half mode 2
// 148 CTAs x `warps` warps; each warp owns candidates as in ar_attn_only_kernel; consuming a tile = xor of 8 x 16 B per lane.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

template <int DEPTH>
__global__ void __launch_bounds__(512) stage_kernel(const uint4* __restrict__ kv, int mode, int B, int H, int Nmax, int nc, int ncph,
                                                    unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // [warps][DEPTH]
  uint8_t* ring = smem + 1024 + (size_t)warp * DEPTH * 4096;
  if (lane == 0) for (int s = 0; s < DEPTH; ++s) mbar_init(&bars[warp * DEPTH + s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const bool use_bulk = mode == 1 || (mode == 3 && (warp & 1) == 0);
  uint32_t par = 0;                                                   // bit s = parity of stage s
  uint4 acc = make_uint4(0, 0, 0, 0);
  const int nt = (nc + 15) / 16, NT = Nmax / 16;
  const int units = H * ncph;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int h = u % H, ci = u / H;
    const int b0 = (int)((long long)ci * B / ncph), b1 = (int)((long long)(ci + 1) * B / ncph);
    for (int b = b0 + warp; b < b1; b += warps) {
      const uint4* base = kv + (((long long)b * H + h) * NT) * 256;
      // prologue: DEPTH tiles in flight
      for (int s = 0; s < DEPTH; ++s) {
        if (s < nt) {
          if (use_bulk) {
            if (lane == 0) { mbar_expect(&bars[warp * DEPTH + s], 4096); bulk_g2s(ring + s * 4096, base + (long long)s * 256, 4096, &bars[warp * DEPTH + s]); }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) cp16(ring + s * 4096 + (i * 32 + lane) * 16, base + (long long)s * 256 + i * 32 + lane);
          }
        }
        if (!use_bulk) asm volatile("cp.async.commit_group;" ::: "memory");
      }
      int s = 0;
      for (int t = 0; t < nt; ++t) {
        if (use_bulk) { mbar_wait(&bars[warp * DEPTH + s], (par >> s) & 1); par ^= 1u << s; }
        else { asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory"); __syncwarp(); }
        const uint4* tp = reinterpret_cast<const uint4*>(ring + s * 4096);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const uint4 v = tp[i * 32 + lane]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
        __syncwarp();
        const int tn = t + DEPTH;
        if (use_bulk) {
          if (lane == 0 && tn < nt) { mbar_expect(&bars[warp * DEPTH + s], 4096); bulk_g2s(ring + s * 4096, base + (long long)tn * 256, 4096, &bars[warp * DEPTH + s]); }
        } else {
          if (tn < nt) {
#pragma unroll
            for (int i = 0; i < 8; ++i) cp16(ring + s * 4096 + (i * 32 + lane) * 16, base + (long long)tn * 256 + i * 32 + lane);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (++s == DEPTH) s = 0;
      }
      if (!use_bulk) asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = 1;
}

template <int DEPTH>
void run(const uint4* kv, size_t layer_bytes, int L, int warps, unsigned* sink, cudaEvent_t e0, cudaEvent_t e1) {
  const int B = 256, H = 16, Nmax = 432, nc = 215;
  const double bytes = (double)B * H * ((nc + 15) / 16) * 4096.0;
  const size_t smem = 1024 + (size_t)warps * DEPTH * 4096;
  if (smem > 227 * 1024) return;
  cudaFuncSetAttribute(stage_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 1; mode <= 3; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      for (int l = 0; l < L; ++l)
        stage_kernel<DEPTH><<<148, warps * 32, smem>>>(kv + (size_t)l * layer_bytes / 16, mode, B, H, Nmax, nc, 9, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms / L < best) best = ms / L;
    }
    printf("warps %2d depth %d (%3zu KB ring) mode %d (%s): %.1f us per layer = %.0f GB/s  [%s]\n", warps, DEPTH, smem / 1024, mode,
           mode == 1 ? "bulk" : mode == 2 ? "cp.async 16B" : "half/half", best * 1e3, bytes / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  const int L = 30;
  const size_t layer_bytes = (size_t)256 * 16 * 432 * 256;
  uint4* kv;
  unsigned* sink;
  cudaMalloc(&kv, layer_bytes * L);
  cudaMalloc(&sink, 4);
  cudaMemset(kv, 1, layer_bytes * L);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 8; warps <= 16; warps += 4) {
    run<1>(kv, layer_bytes, L, warps, sink, e0, e1);
    run<2>(kv, layer_bytes, L, warps, sink, e0, e1);
    run<3>(kv, layer_bytes, L, warps, sink, e0, e1);
    run<4>(kv, layer_bytes, L, warps, sink, e0, e1);
    run<6>(kv, layer_bytes, L, warps, sink, e0, e1);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
