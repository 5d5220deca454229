"""GPU: conditioning front-end through the C-ABI against the CPU oracle (oracle/conditioning.py, pinned
against the reference modules by tests/test_oracle_conditioning.py): resampler, both mel spectrograms, head_dim-128
attention, the two conditioning encoders, the random-voice converter, and tts() fed with voice samples / a random voice
(BASELINE configs[0] / [1] plumbing at a reduced config).

Tolerances: audio kernels are fp32 (log-mel within 2e-3 abs: direct DFT vs FFT summation order + fast-math log); the
encoders use bf16 GEMM operands (latents within 3 % of their scale, as every other stage)."""
import pytest
import torch

from gpu_util import report

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.float().abs().max().item(), 1e-6)


def _clips(lengths=(140000, 60000), seed=4):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in lengths:
        t = torch.arange(n) / 22050.0
        tone = 0.3 * torch.sin(2 * 3.14159265 * 220.0 * t) + 0.1 * torch.sin(2 * 3.14159265 * 1760.0 * t)
        out.append((tone + 0.05 * torch.randn(n, generator=g)).clamp(-1, 1).reshape(1, n))
    return out


@pytest.fixture(scope="module")
def env():
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_all
    cfg = ModelConfig.small()
    return cfg, synth_all(cfg, seed=0, suppress_stop=False)


def test_resample_and_mels(env):
    from tortoise_tts_b200.conditioning_engine import ConditioningEngine, COND_LENGTH
    from oracle import conditioning as oc
    cfg, sds = env
    mel_norms = -(torch.rand(80, generator=torch.Generator().manual_seed(1)) * 5 + 1)
    eng = ConditioningEngine(sds["autoregressive"], sds["diffusion"], cfg, "cuda", mel_norms=mel_norms)
    clips = _clips()
    w0 = oc.format_conditioning_clip(clips[0], 23).reshape(-1)
    mf = torch.empty(80, 1 + COND_LENGTH // 256, device="cuda")
    eng.ar_mel(w0.cuda().contiguous(), mf)
    want = oc.torch_mel_spectrogram(w0.reshape(1, -1), mel_norms)[0]
    e = (mf.cpu() - want).abs().max().item()
    report("cond AR log-mel abs", e)
    assert e < 2e-3
    for c in clips:
        dm = torch.empty(100, 401, device="cuda")
        eng.diffusion_mel(c.reshape(-1).cuda().contiguous(), dm)
        s = oc.resample_22k_24k(c)
        s = s[..., :102400] if s.shape[-1] >= 102400 else torch.nn.functional.pad(s, (0, 102400 - s.shape[-1]))
        e = (dm.cpu() - oc.tacotron_mel(s)[0]).abs().max().item()
        report("cond diffusion log-mel abs (n=%d)" % c.shape[-1], e)
        assert e < 2e-3


@pytest.mark.parametrize("hd,T,H,use_bias,causal", [(128, 101, 16, True, False), (128, 33, 2, False, True), (32, 70, 3, True, False)])
def test_attention_head_dims(hd, T, H, use_bias, causal):
    from tortoise_tts_b200 import lib
    torch.manual_seed(hd + T)
    D = H * hd
    qkv = torch.randn(2 * T, 3 * D, device="cuda").to(torch.bfloat16)
    bias = torch.randn(H, 2 * T - 1, device="cuda") if use_bias else None
    out = torch.empty(2 * T, D, device="cuda", dtype=torch.bfloat16)
    scale = hd ** -0.5
    lib.attention(qkv, out, nseq=2, T=T, H=H, ld=3 * D, ldo=D, k_off=D, v_off=2 * D, scale=scale, causal=causal, bias=bias,
                  head_dim=hd)
    q, k, v = (t.float().view(2, T, H, hd).transpose(1, 2) for t in qkv.split(D, dim=1))
    w = (q @ k.transpose(-1, -2)) * scale
    if use_bias:
        i = torch.arange(T, device="cuda")
        w = w + bias[:, i[None, :] - i[:, None] + T - 1].unsqueeze(0)
    if causal:
        w = w.masked_fill(~torch.ones(T, T, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    want = (torch.softmax(w, -1) @ v).transpose(1, 2).reshape(2 * T, D)
    e = (out.float() - want).abs().max().item()
    report("attention head_dim=%d" % hd, e)
    assert e < 0.03


@pytest.mark.parametrize("which", ["small", "medium"])
def test_conditioning_latents_vs_oracle(which):
    """Both latents from two clips (one cropped, one padded) vs the oracle; `medium` = full widths (1024 / 2048 channels,
    16 heads of 64 / 128), one attention block per encoder."""
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_autoregressive, synth_diffusion
    from tortoise_tts_b200.conditioning_engine import ConditioningEngine
    from oracle import conditioning as oc
    cfg = ModelConfig.small() if which == "small" else ModelConfig.medium()
    sd_ar, sd_df = synth_autoregressive(cfg, 0, False), synth_diffusion(cfg, 0)
    mel_norms = -(torch.rand(80, generator=torch.Generator().manual_seed(1)) * 5 + 1)
    eng = ConditioningEngine(sd_ar, sd_df, cfg, "cuda", mel_norms=mel_norms)
    clips = _clips()
    starts = [23, None]
    with torch.no_grad():
        want_ar = oc.ar_conditioning_latent(sd_ar, cfg, clips, mel_norms, starts)
        want_df = oc.diffusion_conditioning_latent(sd_df, cfg, clips)
    r = _rel(eng.ar_latent(clips, starts), want_ar)
    report("cond AR latent %s" % which, r)
    assert r < 0.03
    r = _rel(eng.diffusion_latent(clips), want_df)
    report("cond diffusion latent %s" % which, r)
    assert r < 0.03


def test_random_latents_and_tts_entry_points(env):
    """get_random_conditioning_latents vs the oracle on the same normal draw; tts() with a random voice (configs[0]) and
    with voice samples (configs[1]) produces the same waveform as tts() fed the latents those entry points return."""
    from tortoise_tts_b200.api import TextToSpeech
    from oracle import conditioning as oc
    cfg, sds = env
    tts = TextToSpeech(state_dicts=sds, config=cfg, kv_cache=True)
    torch.manual_seed(9)
    a, d = tts.get_random_conditioning_latents()
    torch.manual_seed(9)
    ra, rd = torch.randn(1, cfg.ar_dim), torch.randn(1, 2 * cfg.diff_dim)
    assert _rel(a, oc.random_latent(sds["rlg_auto"], ra)) < 1e-4
    assert _rel(d, oc.random_latent(sds["rlg_diffuser"], rd)) < 1e-4
    kw = dict(text_tokens=[42, 2, 194, 91, 24, 2, 243, 190], num_autoregressive_samples=4, diffusion_iterations=5,
              max_mel_tokens=10, verbose=False)
    # random voice: the draw happens inside tts() after deterministic_state(seed) -> reproducible
    w1 = tts.tts("", use_deterministic_seed=3, **kw)
    w2 = tts.tts("", use_deterministic_seed=3, **kw)
    assert w1.shape[0] == 1 and torch.equal(w1, w2)
    # voice samples: crop positions come from `random` seeded by deterministic_state
    clips = _clips()
    w3 = tts.tts("", voice_samples=clips, use_deterministic_seed=5, **kw)
    tts.deterministic_state(5)
    lat = tts.get_conditioning_latents(clips)
    w4 = tts.tts("", conditioning_latents=lat, use_deterministic_seed=5, **kw)
    assert torch.equal(w3, w4)
    al, dl, am, dm = tts.get_conditioning_latents(clips, return_mels=True)
    assert am.shape == (1, 2, 80, 517) and dm.shape == (1, 2, 100, 401)
