"""CPU: the oracle (oracle/*.py) against the committed golden vectors generated from the reference modules
(tests/golden/make_golden.py). Runs anywhere (no GPU, no reference tree)."""
import os

import numpy as np
import pytest
import torch

from tortoise_tts_b200.config import ModelConfig
from tortoise_tts_b200.synth import synth_all

GOLD = os.path.join(os.path.dirname(__file__), "golden", "small_v1.pt")


@pytest.fixture(scope="module")
def env():
    cfg = ModelConfig.small()
    sds = synth_all(cfg, seed=0, suppress_stop=False)
    gold = torch.load(GOLD)
    for k, want in gold["weights_checksum"].items():
        got = float(sum(v.double().sum() for v in sds[k].values()))
        assert abs(got - want) <= 1e-6 * max(1.0, abs(want)), "synthetic checkpoint drifted from the golden run: " + k
    return cfg, sds, gold


def test_ar_logits(env):
    from oracle import ar
    cfg, sds, g = env
    text = g["text"].tolist()
    with torch.no_grad():
        kv = ar.teacher_forced_logits(sds["autoregressive"], cfg, g["ar_cond"], text, g["ar_codes"], "ref_kv_quirk")
        rc = ar.teacher_forced_logits(sds["autoregressive"], cfg, g["ar_cond"], text, g["ar_codes"], "train_consistent")
    assert (kv - g["ar_logits_kv"]).abs().max().item() < 2e-4
    assert (rc - g["ar_logits_recompute"]).abs().max().item() < 2e-4
    # the two position rules really differ
    assert (kv - rc).abs().max().item() > 1e-2


def test_ar_generate_consistent_with_teacher_forcing(env):
    """oracle.generate (KV-cached, injected uniforms) reproduces its own choices under teacher forcing."""
    from oracle import ar
    cfg, sds, g = env
    text = g["text"].tolist()
    torch.manual_seed(0)
    u = torch.rand(2, 5)
    with torch.no_grad():
        codes = ar.generate(sds["autoregressive"], cfg, g["ar_cond"], text, u, 5)
        lg = ar.teacher_forced_logits(sds["autoregressive"], cfg, g["ar_cond"], text, codes[:, :4], "ref_kv_quirk")
    for b in range(2):
        seen = {1, cfg.start_mel_token}
        for n in range(5):
            tok, _, _ = ar.sample_step(lg[b, n], seen, float(u[b, n]))
            assert tok == int(codes[b, n])
            seen.add(tok)


def test_latents(env):
    from oracle import ar
    cfg, sds, g = env
    with torch.no_grad():
        got = ar.latents(sds["autoregressive"], cfg, g["ar_cond"], g["text"].tolist(), g["lat_codes"])
    assert (got - g["latents"]).abs().max().item() < 2e-4


def test_clvp(env):
    from oracle import clvp
    cfg, sds, g = env
    with torch.no_grad():
        got = clvp.scores(sds["clvp"], cfg, g["text"], g["clvp_codes"])
    assert (got - g["clvp_scores"]).abs().max().item() < 1e-5


def test_diffusion(env):
    from oracle import diffusion as od
    cfg, sds, g = env
    sd = sds["diffusion"]
    S = g["code_emb"].shape[-1]
    with torch.no_grad():
        ce = od.timestep_independent(sd, cfg, g["diff_latents"], g["diff_cond"], S)
        assert (ce - g["code_emb"]).abs().max().item() < 1e-4
        t = torch.tensor([3979])
        assert (od.forward(sd, cfg, g["diff_x"], t, code_emb=ce) - g["diff_fwd_cond"]).abs().max().item() < 1e-4
        assert (od.forward(sd, cfg, g["diff_x"], t, conditioning_free=True) - g["diff_fwd_uncond"]).abs().max().item() < 1e-4
        mel = od.spectrogram_diffusion(sd, cfg, g["diff_latents"], g["diff_cond"], g["diff_noise0"], g["diff_step_noise"],
                                       g["diff_iters"], True, 2.0)
    assert (mel - g["diff_mel"]).abs().max().item() < 1e-3


def test_vocoder(env):
    from oracle import vocoder as ov
    cfg, sds, g = env
    with torch.no_grad():
        got = ov.inference(sds["vocoder"], g["voc_mel"], g["voc_z"])
    assert (got - g["voc_wav"]).abs().max().item() < 1e-4


def test_schedule_known_answers():
    """SURVEY App. A3 KATs (derived from the reference's SpacedDiffusion in the survey container)."""
    from oracle import diffusion as od
    kat = {30: ([0, 138, 276], [3861, 3999], 1.526510485e-2, 4.933385471e-1, -10.598246, -0.706601),
           80: ([0, 51, 101], [3948, 3999], 2.920444499e-3, 2.243435051e-1, -10.605134, -1.494589),
           200: ([0, 20, 40], [3979, 3999], 7.609781250e-4, 9.517459723e-2, -10.628935, -2.352047),
           400: ([0, 10, 20], [3989, 3999], 3.183777965e-4, 4.883635492e-2, -10.672204, -3.019282)}
    for iters, (head, tail, b1, blast, p0, plast) in kat.items():
        s = od.make_schedule(iters)
        assert list(s["timestep_map"][:3]) == head and list(s["timestep_map"][-2:]) == tail
        assert abs(s["betas"][0] - 2.5e-5) < 1e-12
        assert abs(s["betas"][1] - b1) < 1e-9 and abs(s["betas"][-1] - blast) < 1e-8
        assert abs(s["posterior_log_variance_clipped"][0] - p0) < 1e-5
        assert abs(s["posterior_log_variance_clipped"][-1] - plast) < 1e-5
        assert abs(s["alphas_cumprod"][-1] - 4.246652276e-5) < 1e-12
        assert abs(s["sqrt_recip_alphas_cumprod"][-1] - 153.453447) < 1e-4


def test_relpos_bucket_and_timestep_embedding_kats():
    from oracle import diffusion as od
    rel = torch.arange(-70, 71, 10)
    assert od.rel_pos_bucket(rel).tolist() == [15, 15, 15, 14, 13, 11, 8, 0, 24, 27, 29, 30, 31, 31, 31]
    e = od.timestep_embedding(torch.tensor([3979]), 1024)[0]
    assert np.allclose(e[:3].numpy(), [-0.172044, 0.996861, 0.803411], atol=2e-4)
    assert np.allclose(e[512:515].numpy(), [0.985089, -0.079166, -0.595425], atol=2e-4)


def test_fix_autoregressive_output_constants():
    from oracle import ar
    c = ar.fix_autoregressive_output(torch.tensor([5, 6, 7, 8193, 8193, 8193, 8193, 8193]))
    assert c.tolist() == [5, 6, 7, 83, 83, 45, 45, 248]
    c = ar.fix_autoregressive_output(torch.tensor([1, 2, 3, 4]))
    assert c.tolist() == [1, 2, 3, 4]
    row = torch.tensor([7] * 5 + [83] * 12 + [9])
    assert ar.calm_trim_length(row) == 13
