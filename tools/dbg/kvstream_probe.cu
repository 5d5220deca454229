// Development probe (no product code): how fast can 148 CTAs x 16 warps stream the candidate KV cache of one decode-attention
// layer (B = 256 candidates x 16 heads, 215 positions x 256 B each = 226 MB) under different memory layouts?
//   layout 0: [B][H][Nmax][256 B]           (the cache layout of round 2: 2k concurrent 55 KB streams, 110 KB apart)
//   layout 1: [Nt][H][B][16 pos][256 B]     (position-tile major: the tiles the GPU reads at one moment are adjacent)
//   layout 2: [H][B][Nmax][256 B]           (head major: a CTA's candidates are adjacent 110 KB regions)
// Work assignment as in ar_attn_only_kernel: CTA u -> head u % H, candidate range u / H of ncph; warp w -> one candidate.
// Each warp reads 4 KB tiles with 8 x 16-byte loads per lane (all issued before use) and xors them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o kvstream_probe kvstream_probe.cu && ./kvstream_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) stream_kernel(const uint4* __restrict__ kv, int layout, int B, int H, int Nmax, int nc,
                                                     int ncph, int depth, unsigned* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = H * ncph;
  uint4 acc = make_uint4(0, 0, 0, 0);
  const int nt = (nc + 15) / 16, NT = Nmax / 16;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int h = u % H, ci = u / H;
    const int b0 = (int)((long long)ci * B / ncph), b1 = (int)((long long)(ci + 1) * B / ncph);
    for (int b = b0 + warp; b < b1; b += 16) {
      for (int t = 0; t < nt; t += depth) {
        uint4 v[4][8];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          if (d < depth && t + d < nt) {
            long long tile;   // in 4 KB units
            if (layout == 0) tile = ((long long)b * H + h) * NT + (t + d);
            else if (layout == 1) tile = ((long long)(t + d) * H + h) * B + b;
            else tile = ((long long)h * B + b) * NT + (t + d);
            const uint4* p = kv + tile * 256 + lane;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[d][i] = __ldcs(p + i * 32);
          }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d)
          if (d < depth && t + d < nt)
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc.x ^= v[d][i].x; acc.y ^= v[d][i].y; acc.z ^= v[d][i].z; acc.w ^= v[d][i].w; }
      }
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = 1;
}

int main() {
  const int B = 256, H = 16, Nmax = 432, nc = 215, L = 30;
  const size_t layer_bytes = (size_t)B * H * Nmax * 256;
  uint4* kv;
  unsigned* sink;
  cudaMalloc(&kv, layer_bytes * L);
  cudaMalloc(&sink, 4);
  cudaMemset(kv, 1, layer_bytes * L);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double bytes = (double)B * H * ((nc + 15) / 16) * 4096.0;
  for (int ncph = 9; ncph <= 18; ncph += 9)
    for (int depth = 1; depth <= 4; depth *= 2)
      for (int layout = 0; layout < 3; ++layout) {
        float best = 1e9f, tot = 0.f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          for (int l = 0; l < L; ++l)
            stream_kernel<<<148, 512>>>(kv + (size_t)l * layer_bytes / 16, layout, B, H, Nmax, nc, ncph, depth, sink);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          tot = ms / L;
          if (tot < best) best = tot;
        }
        printf("ncph %2d depth %d layout %d: %.1f us per layer = %.0f GB/s\n", ncph, depth, layout, best * 1e3, bytes / best / 1e6);
      }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
