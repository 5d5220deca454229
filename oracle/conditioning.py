"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of the conditioning front-end.

Follows, in plain torch-fp32 / numpy-f64:
  * TextToSpeech.get_conditioning_latents                  api.py:258-299
  * format_conditioning (pad / crop to 132300 samples)     api.py:73-84
  * TorchMelSpectrogram (torchaudio MelSpectrogram + log + mel_norms)   models/arch_util.py:295-331
  * ConditioningEncoder / UnifiedVoice.get_conditioning    models/autoregressive.py:204-228, 444-452
  * torchaudio.functional.resample(22050 -> 24000), pad_or_truncate(102400)   api.py:284-285
  * TacotronSTFT.mel_spectrogram / STFT.transform          utils/audio.py:151-191, utils/stft.py:94-157
  * DiffusionTts.contextual_embedder / get_conditioning    models/diffusion_decoder.py:186-192, 222-230
  * RandomLatentConverter / EqualLinear                    models/random_latent_generator.py:8-50

`torchaudio` is a dependency of the reference that IS present in this image; it is called directly where the
reference calls it (MelSpectrogram, resample). `librosa` is absent: `slaney_mel_basis` restates
librosa.filters.mel(htk=False, norm='slaney') and is pinned against torchaudio's independent implementation of the
same formula (tests/test_oracle_conditioning.py). Parity against the reference modules: same file.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import diffusion as od

COND_LENGTH = 132300        # api.py:73
DIFF_COND_LENGTH = 102400   # api.py:285


# ------------------------------------------------------------------ mel filterbanks (f64)
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def _hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def _mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax, mel_scale):
    """Triangular filters with Slaney area normalisation, [n_mels, n_fft//2+1] float64.
    mel_scale 'slaney' = librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) defaults (utils/audio.py:158-160);
    mel_scale 'htk' = torchaudio melscale_fbanks(norm='slaney', mel_scale='htk') (arch_util.py:307-311)."""
    to_mel, to_hz = (_hz_to_mel_slaney, _mel_to_hz_slaney) if mel_scale == "slaney" else (_hz_to_mel_htk, _mel_to_hz_htk)
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    pts = to_hz(np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2))
    fdiff = np.diff(pts)
    ramps = pts[:, None] - freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (pts[2:n_mels + 2] - pts[:n_mels])
    return w * enorm[:, None]


# ------------------------------------------------------------------ AR conditioning mel
def format_conditioning_clip(clip, start=None):
    """api.py:73-84 up to the spectrogram: clip [1, n] -> [1, 132300]. `start` replaces random.randint(0, gap)."""
    gap = clip.shape[-1] - COND_LENGTH
    if gap < 0:
        return F.pad(clip, (0, -gap))
    if gap > 0:
        s = 0 if start is None else int(start)
        return clip[:, s:s + COND_LENGTH]
    return clip


def torch_mel_spectrogram(wav, mel_norms):
    """TorchMelSpectrogram.forward (arch_util.py:318-331): wav [1, n] -> [1, 80, 1 + n // 256]."""
    import torchaudio
    tr = torchaudio.transforms.MelSpectrogram(n_fft=1024, hop_length=256, win_length=1024, power=2, normalized=False,
                                              sample_rate=22050, f_min=0, f_max=8000, n_mels=80, norm="slaney")
    mel = torch.log(torch.clamp(tr(wav), min=1e-5))
    return mel / mel_norms.reshape(1, -1, 1)


def conditioning_encoder(sd, cfg, mel):
    """ConditioningEncoder.forward (autoregressive.py:223-228): mel [1, 80, T] -> [1, D] (position 0)."""
    h = F.conv1d(mel, sd["conditioning_encoder.init.weight"], sd["conditioning_encoder.init.bias"])
    for i in range(cfg.cond_enc_blocks):
        h = od.attention_block(sd, f"conditioning_encoder.attn.{i}.", h, cfg.ar_heads, rel_pos=False)
    return h[:, :, 0]


def ar_conditioning_latent(sd, cfg, clips, mel_norms, starts=None):
    """get_conditioning_latents, first half (api.py:268-276) + UnifiedVoice.get_conditioning (autoregressive.py:444-452)."""
    outs = []
    for i, c in enumerate(clips):
        w = format_conditioning_clip(c, None if starts is None else starts[i])
        outs.append(conditioning_encoder(sd, cfg, torch_mel_spectrogram(w, mel_norms)))
    return torch.stack(outs, dim=1).mean(dim=1)


# ------------------------------------------------------------------ diffusion conditioning mel
def resample_22k_24k(wav):
    import torchaudio
    return torchaudio.functional.resample(wav, 22050, 24000)


def tacotron_mel(wav, n_mels=100, sr=24000, fmax=12000.0):
    """TacotronSTFT(1024, 256, 1024, 100, 24000, 0, 12000).mel_spectrogram (audio.py:177-191) with STFT.transform
    (stft.py:133-157): conv with the hann-windowed DFT basis after reflect padding; magnitude; mel; log(clamp 1e-5)."""
    y = torch.clip(wav, -1, 1)
    n_fft, hop = 1024, 256
    basis = np.fft.fft(np.eye(n_fft))
    cutoff = n_fft // 2 + 1
    fb = np.vstack([np.real(basis[:cutoff]), np.imag(basis[:cutoff])])
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)          # scipy get_window('hann', fftbins=True)
    fwd = torch.from_numpy((fb * win[None, :]).astype(np.float32)).unsqueeze(1)
    x = F.pad(y.unsqueeze(1), (n_fft // 2, n_fft // 2), mode="reflect")
    ft = F.conv1d(x, fwd, stride=hop)
    mag = torch.sqrt(ft[:, :cutoff] ** 2 + ft[:, cutoff:] ** 2)
    mb = torch.from_numpy(mel_filterbank(sr, n_fft, n_mels, 0.0, fmax, "slaney")).float()
    return torch.log(torch.clamp(torch.matmul(mb, mag), min=1e-5))


def contextual_embedder(sd, cfg, mel):
    """DiffusionTts.contextual_embedder (diffusion_decoder.py:186-192): mel [1, 100, T] -> [1, 2C, T//4 + 1]."""
    h = F.conv1d(mel, sd["contextual_embedder.0.weight"], sd["contextual_embedder.0.bias"], stride=2, padding=1)
    h = F.conv1d(h, sd["contextual_embedder.1.weight"], sd["contextual_embedder.1.bias"], stride=2, padding=1)
    for i in range(2, 7):
        h = od.attention_block(sd, f"contextual_embedder.{i}.", h, cfg.diff_heads, rel_pos=True)
    return h


def diffusion_conditioning_latent(sd, cfg, clips):
    """get_conditioning_latents, second half (api.py:278-294) + DiffusionTts.get_conditioning (diffusion_decoder.py:222-230)."""
    conds = []
    for c in clips:
        s = resample_22k_24k(c)
        s = s[..., :DIFF_COND_LENGTH] if s.shape[-1] >= DIFF_COND_LENGTH else F.pad(s, (0, DIFF_COND_LENGTH - s.shape[-1]))
        conds.append(contextual_embedder(sd, cfg, tacotron_mel(s)))
    return torch.cat(conds, dim=-1).mean(dim=-1)


# ------------------------------------------------------------------ random voice
def random_latent(sd_rlg, r):
    """RandomLatentConverter.forward with the normal draw `r` [1, C] injected (random_latent_generator.py:40-50)."""
    C = r.shape[-1]
    y = r
    for i in range(5):
        w, b = sd_rlg[f"layers.{i}.weight"], sd_rlg[f"layers.{i}.bias"]
        y = F.leaky_relu(F.linear(y, w * ((1 / math.sqrt(C)) * 0.1)) + b * 0.1, 0.2) * (2 ** 0.5)
    return F.linear(y, sd_rlg["layers.5.weight"], sd_rlg["layers.5.bias"])
