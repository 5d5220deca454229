"""Model/shape configuration of the four networks on the hot path.

Full-size values are the constructor arguments the reference passes in
`tortoise/api.py:217-237` (UnifiedVoice, DiffusionTts, CLVP, UnivNetGenerator).
`small()` is a reduced configuration (head_dim stays 64) used by parity tests so that
the CPU oracle finishes in seconds.
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class ModelConfig:
    # UnifiedVoice (api.py:217-220, autoregressive.py:293-357)
    ar_layers: int = 30
    ar_dim: int = 1024
    ar_heads: int = 16
    max_mel_tokens: int = 604
    max_text_tokens: int = 402
    max_conditioning_inputs: int = 2
    number_text_tokens: int = 255
    start_text_token: int = 255
    stop_text_token: int = 0
    number_mel_codes: int = 8194
    start_mel_token: int = 8192
    stop_mel_token: int = 8193
    cond_enc_blocks: int = 6
    # DiffusionTts (api.py:224-226)
    diff_dim: int = 1024
    diff_layers: int = 10
    diff_heads: int = 16
    diff_in_channels: int = 100
    diff_out_channels: int = 200
    diff_in_tokens: int = 8193
    # CLVP (api.py:229-232)
    clvp_dim: int = 768
    clvp_depth: int = 20
    clvp_heads: int = 12
    clvp_text_tokens: int = 256
    clvp_speech_tokens: int = 8192
    # CVVP (api.py:252-256; only used when tts(cvvp_amount > 0, voice_samples=...))
    cvvp_dim: int = 512
    cvvp_depth: int = 8
    cvvp_heads: int = 8
    # HifiganGenerator of the `api_fast` path (api_fast.py:221-224): in_channels = ar_dim, cond_channels = ar_dim
    hifi_channels: int = 512
    # UnivNet (vocoder.py:232-233)
    voc_noise_dim: int = 64
    voc_channels: int = 32
    voc_mel: int = 100
    voc_kp_hidden: int = 64

    @property
    def mel_pos_rows(self):  # autoregressive.py:339: max_mel_tokens + 2 + max_conditioning_inputs
        return self.max_mel_tokens + 2 + self.max_conditioning_inputs

    @property
    def text_pos_rows(self):
        return self.max_text_tokens + 2

    def to_dict(self):
        return asdict(self)

    @staticmethod
    def full():
        return ModelConfig()

    @staticmethod
    def small():
        return ModelConfig(ar_layers=2, ar_dim=128, ar_heads=2, cond_enc_blocks=1,
                           diff_dim=128, diff_layers=2, diff_heads=2,
                           clvp_dim=128, clvp_depth=2, clvp_heads=2, cvvp_dim=128, cvvp_depth=2, cvvp_heads=2, hifi_channels=128)

    @staticmethod
    def medium():
        """Full widths, few layers: exercises every full-size tile shape cheaply."""
        return ModelConfig(ar_layers=2, cond_enc_blocks=1, diff_layers=1, clvp_depth=2, cvvp_depth=2)


HIFI_UP_FACTORS = (8, 8, 2, 2)   # api_fast.py:223 (kernel sizes 16, 16, 4, 4 = 2 x factor)
HIFI_RES_KERNELS = (3, 7, 11)    # api_fast.py:222
HIFI_RES_DILATIONS = (1, 3, 5)   # api_fast.py:222 (the same for the three kernel sizes)
HIFI_LRELU = 0.1                 # hifigan_decoder.py:8
VOC_STRIDES = (8, 8, 4)          # vocoder.py:232
VOC_DILATIONS = (1, 3, 9, 27)    # vocoder.py:232
VOC_LRELU = 0.2
