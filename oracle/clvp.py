"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of CLVP.forward(return_loss=False).

Plain torch-fp32 functional code over the reference `clvp2.pth` state_dict. Follows
clvp.py:99-140 and the x-transformers Encoder configuration the reference instantiates
(clvp.py:54-83): pre-RMSNorm (xtransformers.py:335-344), bias-free q/k/v (519-521), rotary of
dim 32 applied to q, k AND v (625-629, 264-286), softmax(q k^T / 8), GEGLU feed-forward with
erf-GELU (429-474), final LayerNorm (1234), masked mean with an all-ones mask (clvp.py:15-17).
"""

import torch
import torch.nn.functional as F


def _rmsnorm(x, g):
    D = x.shape[-1]
    norm = torch.norm(x, dim=-1, keepdim=True) * (D ** -0.5)
    return x / norm.clamp(min=1e-8) * g


def _rotary(t, freqs):
    """t [..., n, 32]; freqs [n, 32] = cat(f, f)."""
    d = t.shape[-1] // 2
    x1, x2 = t[..., :d], t[..., d:]
    rot = torch.cat((-x2, x1), dim=-1)
    return t * freqs.cos() + rot * freqs.sin()


def encoder(sd, prefix, x, depth, heads, wrap="wrap."):
    """x [B, n, D] -> [B, n, D] (after the final LayerNorm). `wrap`: "wrap." for the CheckpointedXTransformerEncoder CLVP
    uses (arch_util.py:334-373), "" for a plain ContinuousTransformerWrapper (CVVP, cvvp.py:23-36)."""
    B, n, D = x.shape
    hd = D // heads
    inv_freq = sd[prefix + "attn_layers.rotary_pos_emb.inv_freq"]
    tpos = torch.arange(n, dtype=torch.float32)
    f = torch.einsum("i,j->ij", tpos, inv_freq)
    freqs = torch.cat((f, f), dim=-1)  # [n, 32]
    rd = freqs.shape[-1]
    for l in range(depth):
        a = f"{prefix}attn_layers.layers.{2 * l}."
        h = _rmsnorm(x, sd[a + "0.0.g"])
        q = h @ sd[a + "1." + wrap + "to_q.weight"].t()
        k = h @ sd[a + "1." + wrap + "to_k.weight"].t()
        v = h @ sd[a + "1." + wrap + "to_v.weight"].t()
        q, k, v = (t.view(B, n, heads, hd).transpose(1, 2) for t in (q, k, v))
        q, k, v = (torch.cat((_rotary(t[..., :rd], freqs), t[..., rd:]), dim=-1) for t in (q, k, v))
        w = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
        o = (w @ v).transpose(1, 2).reshape(B, n, D)
        x = x + (o @ sd[a + "1." + wrap + "to_out.weight"].t() + sd[a + "1." + wrap + "to_out.bias"])
        fpre = f"{prefix}attn_layers.layers.{2 * l + 1}."
        h = _rmsnorm(x, sd[fpre + "0.0.g"])
        proj = h @ sd[fpre + "1." + wrap + "net.0.proj.weight"].t() + sd[fpre + "1." + wrap + "net.0.proj.bias"]
        u, g = proj.chunk(2, dim=-1)
        h = u * F.gelu(g)
        x = x + (h @ sd[fpre + "1." + wrap + "net.3.weight"].t() + sd[fpre + "1." + wrap + "net.3.bias"])
    return F.layer_norm(x, (D,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-5)


def text_latent(sd, cfg, text_ids):
    """text_ids [T'] (api.py passes the once-zero-padded tokens) -> normalised [D]."""
    x = sd["text_emb.weight"][text_ids.long()].unsqueeze(0)
    e = encoder(sd, "text_transformer.transformer.", x, cfg.clvp_depth, cfg.clvp_heads)
    lat = e.mean(dim=1) @ sd["to_text_latent.weight"].t()
    return F.normalize(lat, p=2, dim=-1)[0]


def speech_latents(sd, cfg, codes):
    """codes [B, L] -> normalised [B, D]."""
    x = sd["speech_emb.weight"][codes.long()]
    e = encoder(sd, "speech_transformer.transformer.", x, cfg.clvp_depth, cfg.clvp_heads)
    lat = e.mean(dim=1) @ sd["to_speech_latent.weight"].t()
    return F.normalize(lat, p=2, dim=-1)


def scores(sd, cfg, text_ids, codes):
    """clvp(text.repeat(B,1), codes, return_loss=False) -> [B] (clvp.py:126-135)."""
    t = text_latent(sd, cfg, text_ids)
    s = speech_latents(sd, cfg, codes)
    return (s @ t) * sd["temperature"].exp()
