// Conditioning front-end kernels: what turns reference clips into the mel spectrograms the two
// conditioning encoders consume. All fp32 (audio dynamic range); tables (window, twiddles, mel filterbank, resampling
// kernels) are built on the host in float64 exactly as the reference's libraries build them and passed in.
//   * ttb_audio_resample   = torchaudio.functional.resample (polyphase sinc, api.py:284)
//   * ttb_audio_stft_mel   = torchaudio MelSpectrogram + log + mel_norms (arch_util.py:295-331) and
//                            TacotronSTFT.mel_spectrogram (utils/audio.py:177-191, utils/stft.py:133-157)
//   * ttb_mean_rows        = the means over positions / clips (autoregressive.py:451, diffusion_decoder.py:229)
#include "common.cuh"
#include "ttb_internal.h"

namespace ttb {

// out[i * up + j] = sum_k xpad[i * down + k] * kern[j][k], xpad = x zero-padded by `width` on the left
// (torchaudio _apply_sinc_resample_kernel: pad (width, width + orig), conv1d stride orig, interleave the `new` phases)
__global__ void audio_resample_kernel(const float* __restrict__ x, int n, const float* __restrict__ kern, int down, int up,
                                      int klen, int width, float* __restrict__ out, int m) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= m) return;
  const int i = o / up, j = o - i * up;
  const float* kj = kern + (long long)j * klen;
  const int base = i * down - width;
  float acc = 0.f;
  for (int k = 0; k < klen; ++k) {
    const int s = base + k;
    if (s >= 0 && s < n) acc += x[s] * kj[k];
  }
  out[o] = acc;
}

// One CTA per frame. frame t covers samples [t*hop - n_fft/2, +n_fft) of x with reflect padding (center=True).
//   spec[k] = | sum_n x[n] w[n] e^{-2 pi i k n / n_fft} |^power ;  mel[c] = sum_k fb[c][k] spec[k]
//   y = log(max(mel, floor)) / div[c]
constexpr int STFT_THREADS = 256;
__global__ void __launch_bounds__(STFT_THREADS)
audio_stft_mel_kernel(const float* __restrict__ x, int n, int n_fft, int hop, const float* __restrict__ window,
                      const float2* __restrict__ twiddle, const float* __restrict__ fb, int n_mels, int power, int clip,
                      float floor_v, const float* __restrict__ div, __nv_bfloat16* __restrict__ out_bf16, int ldo,
                      float* __restrict__ out_f32, int frames) {
  extern __shared__ float sm[];
  float* fr = sm;                       // n_fft windowed samples
  float* spec = sm + n_fft;             // n_fft/2 + 1
  const int t = blockIdx.x;
  const int nfreq = n_fft / 2 + 1;
  for (int i = threadIdx.x; i < n_fft; i += STFT_THREADS) {
    int s = t * hop - n_fft / 2 + i;
    if (s < 0) s = -s;
    if (s >= n) s = 2 * (n - 1) - s;
    float v = (s >= 0 && s < n) ? x[s] : 0.f;
    if (clip) v = fminf(fmaxf(v, -1.f), 1.f);
    fr[i] = v * window[i];
  }
  __syncthreads();
  const int mask = n_fft - 1;           // n_fft is a power of two
  for (int k = threadIdx.x; k < nfreq; k += STFT_THREADS) {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int i = 0; i < n_fft; ++i) {
      const float2 w = __ldg(twiddle + idx);      // (cos, sin)(2 pi idx / n_fft)
      re = fmaf(fr[i], w.x, re);
      im = fmaf(fr[i], -w.y, im);
      idx = (idx + k) & mask;
    }
    const float p2 = re * re + im * im;
    spec[k] = (power == 2) ? p2 : sqrtf(p2);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_mels; c += STFT_THREADS) {
    const float* f = fb + (long long)c * nfreq;
    float acc = 0.f;
    for (int k = 0; k < nfreq; ++k) acc = fmaf(f[k], spec[k], acc);
    float y = logf(fmaxf(acc, floor_v));
    if (div) y /= div[c];
    if (out_f32) out_f32[(long long)c * frames + t] = y;
    if (out_bf16) out_bf16[(long long)t * ldo + c] = __float2bfloat16(y);
  }
  if (out_bf16)
    for (int c = n_mels + threadIdx.x; c < ldo; c += STFT_THREADS) out_bf16[(long long)t * ldo + c] = __float2bfloat16(0.f);
}

// out[c] = scale * sum_r x[r * ld + c]  (+ accumulate)
__global__ void mean_rows_kernel(const float* __restrict__ x, int R, int C, int ld, float scale, int accumulate,
                                 float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc += x[(long long)r * ld + c];
  out[c] = (accumulate ? out[c] : 0.f) + acc * scale;
}

// y[m, :] = leaky_relu(x[m, :] @ (W * wscale)^T + b * bscale, slope) * gain     (EqualLinear, random_latent_generator.py:21-37;
// slope = 1, gain = 1, scales = 1 gives nn.Linear). One warp per output element row-block; fp32.
__global__ void equal_linear_kernel(const float* __restrict__ x, int K, const float* __restrict__ W, const float* __restrict__ b,
                                    int N, float wscale, float bscale, float slope, float gain, float* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* w = W + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(x[k], w[k] * wscale, acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (b ? b[n] * bscale : 0.f);
    v = (v > 0.f ? v : v * slope) * gain;
    out[n] = v;
  }
}

}  // namespace ttb
using namespace ttb;

extern "C" int ttb_audio_resample(const float* x, int n, const float* kernels, int down, int up, int klen, int width,
                                  float* out, int m, void* stream) {
  if (m <= 0) return 0;
  audio_resample_kernel<<<(m + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, kernels, down, up, klen, width, out, m);
  TTB_CHECK_LAUNCH("audio_resample_kernel");
  return 0;
}

extern "C" int ttb_audio_stft_mel(const float* x, int n, int n_fft, int hop, const float* window, const float* twiddle,
                                  const float* fb, int n_mels, int power, int clip, float floor_v, const float* div,
                                  void* out_bf16, int ldo, float* out_f32, void* stream) {
  if (n_fft <= 0 || (n_fft & (n_fft - 1)) || hop <= 0 || n < n_fft / 2 + 1) { set_error("ttb_audio_stft_mel: bad shape"); return -1; }
  if (power != 1 && power != 2) { set_error("ttb_audio_stft_mel: power must be 1 or 2"); return -1; }
  const int frames = 1 + n / hop;
  const size_t smem = (size_t)(n_fft + n_fft / 2 + 1) * sizeof(float);
  audio_stft_mel_kernel<<<frames, STFT_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      x, n, n_fft, hop, window, reinterpret_cast<const float2*>(twiddle), fb, n_mels, power, clip, floor_v, div,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, out_f32, frames);
  TTB_CHECK_LAUNCH("audio_stft_mel_kernel");
  return 0;
}

extern "C" int ttb_mean_rows(const float* x, int R, int C, int ld, float scale, int accumulate, float* out, void* stream) {
  if (C <= 0) return 0;
  mean_rows_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, R, C, ld, scale, accumulate, out);
  TTB_CHECK_LAUNCH("mean_rows_kernel");
  return 0;
}

extern "C" int ttb_equal_linear(const float* x, int K, const float* W, const float* b, int N, float wscale, float bscale,
                                float slope, float gain, float* out, void* stream) {
  if (N <= 0) return 0;
  equal_linear_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, K, W, b, N, wscale, bscale, slope, gain, out);
  TTB_CHECK_LAUNCH("equal_linear_kernel");
  return 0;
}
