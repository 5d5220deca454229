"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of CVVP.forward(return_loss=False) as `tortoise/api.py:464-472`
uses it: one similarity per candidate between the voice-conditioning mel and the candidate's mel codes.

Plain torch-fp32 functional code over the reference `cvvp.pth` state_dict (cvvp.py:64-135):
  cond_emb: Conv1d(80, D/2, k5, s2, p2) -> Conv1d(D/2, D, k3, s2, p1)                               (cvvp.py:83-84)
  CollapsingTransformer: x-transformers Encoder (pre-RMSNorm, rotary 32 on q/k/v, GEGLU with ff_mult 1, final LayerNorm;
    the same block as CLVP's, oracle/clvp.py) -> Conv1d k1 -> AttentionBlock (no relative positions) -> Conv1d k1 ->
    mean over positions (masked_mean with an all-ones mask in eval)                                 (cvvp.py:20-51)
  speech_emb: nn.Embedding(8192, D) (ConvFormatEmbedding)                                           (cvvp.py:54-61,93)
  sim = <normalize(to_conditioning_latent(.)), normalize(to_speech_latent(.))> * exp(temperature)   (cvvp.py:108-124)
Pinned against the reference module by tests/test_oracle_vs_reference.py::test_cvvp.
"""
import torch
import torch.nn.functional as F

from .clvp import encoder
from .diffusion import attention_block


def _collapse(sd, p, x, depth, heads):
    """CollapsingTransformer.forward in eval mode (cvvp.py:43-51): x [B, n, D] -> [B, D]."""
    h = encoder(sd, p + "transformer.", x, depth, heads, wrap="").permute(0, 2, 1)
    h = F.conv1d(h, sd[p + "pre_combiner.0.weight"], sd[p + "pre_combiner.0.bias"])
    h = attention_block(sd, p + "pre_combiner.1.", h, heads, rel_pos=False)
    h = F.conv1d(h, sd[p + "pre_combiner.2.weight"], sd[p + "pre_combiner.2.bias"])
    return h.mean(dim=2)


def cond_latent(sd, cfg, mel):
    """mel [1, 80, T] (one conditioning clip, `auto_conds[:, cl]`) -> normalised [1, D]."""
    e = F.conv1d(mel, sd["cond_emb.0.weight"], sd["cond_emb.0.bias"], stride=2, padding=2)
    e = F.conv1d(e, sd["cond_emb.1.weight"], sd["cond_emb.1.bias"], stride=2, padding=1).permute(0, 2, 1)
    enc = _collapse(sd, "conditioning_transformer.", e, cfg.cvvp_depth, cfg.cvvp_heads)
    return F.normalize(enc @ sd["to_conditioning_latent.weight"].t(), p=2, dim=-1)


def speech_latents(sd, cfg, codes):
    """codes [B, L] -> normalised [B, D]."""
    x = sd["speech_emb.emb.weight"][codes.long()]
    enc = _collapse(sd, "speech_transformer.", x, cfg.cvvp_depth, cfg.cvvp_heads)
    return F.normalize(enc @ sd["to_speech_latent.weight"].t(), p=2, dim=-1)


def scores(sd, cfg, mels, codes):
    """The accumulation of api.py:464-468: mean over the conditioning clips of cvvp(clip.repeat(B), codes) -> [B].
    mels: list of [1, 80, T] (or a tensor [1, n_clips, 80, T] = auto_conds)."""
    if torch.is_tensor(mels):
        mels = [mels[:, i] for i in range(mels.shape[1])]
    s = speech_latents(sd, cfg, codes)
    acc = 0
    for m in mels:
        acc = acc + (s @ cond_latent(sd, cfg, m)[0]) * sd["temperature"].exp()
    return acc / len(mels)
