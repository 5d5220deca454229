"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of the diffusion decoder path.

Plain torch-fp32 / numpy-f64 functional code over the reference `diffusion_decoder.pth`
state_dict. Follows:
  * get_named_beta_schedule('linear', 4000)              utils/diffusion.py:94-111
  * space_timesteps(4000, [iters])                       utils/diffusion.py:1152-1205
  * SpacedDiffusion / GaussianDiffusion tables           utils/diffusion.py:1093-1115, 192-249
  * p_mean_variance (learned range, CFG ramp), p_sample  utils/diffusion.py:312-418, 487-531
  * DiffusionTts.timestep_independent / forward          models/diffusion_decoder.py:232-322
  * ResBlock / DiffusionLayer                            models/diffusion_decoder.py:60-131
  * AttentionBlock / QKVAttentionLegacy / GroupNorm32    models/arch_util.py:21-123
  * RelativePositionBias (T5 buckets, bidirectional)     models/xtransformers.py:146-186
  * do_spectrogram_diffusion, denormalize_tacotron_mel   api.py:117-130, utils/audio.py:59-64
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

TACOTRON_MEL_MAX = 2.3143386840820312
TACOTRON_MEL_MIN = -11.512925148010254


# --------------------------------------------------------------------------- schedule (f64)
def space_timesteps_single(num_timesteps, count):
    """space_timesteps(num_timesteps, [count]) for a single section -> sorted list."""
    if count <= 1:
        frac = 1
    else:
        frac = (num_timesteps - 1) / (count - 1)
    cur = 0.0
    out = []
    for _ in range(count):
        out.append(round(cur))
        cur += frac
    return sorted(set(out))


def make_schedule(iters, trained_steps=4000):
    """All per-step tables of SpacedDiffusion(space_timesteps(4000,[iters]), linear betas).
    Returns dict of float64 numpy arrays of length len(timestep_map) plus 'timestep_map'."""
    scale = 1000 / trained_steps
    base_betas = np.linspace(scale * 0.0001, scale * 0.02, trained_steps, dtype=np.float64)
    base_ac = np.cumprod(1.0 - base_betas, axis=0)
    use = set(space_timesteps_single(trained_steps, iters))
    last = 1.0
    new_betas, tmap = [], []
    for i, ac in enumerate(base_ac):
        if i in use:
            new_betas.append(1 - ac / last)
            last = ac
            tmap.append(i)
    betas = np.array(new_betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "timestep_map": np.array(tmap, dtype=np.int64),
        "betas": betas,
        "log_betas": np.log(betas),
        "alphas_cumprod": ac,
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


# --------------------------------------------------------------------------- building blocks
def groups_for(C):
    """normalization() group rule (arch_util.py:26-41)."""
    groups = 32
    if C <= 16:
        groups = 8
    elif C <= 64:
        groups = 16
    while C % groups != 0:
        groups = int(groups / 2)
    return groups


def _gn(x, w, b):
    return F.group_norm(x.float(), groups_for(x.shape[1]), w, b, 1e-5)


def rel_pos_bucket(rel, num_buckets=32, max_distance=64):
    """Bidirectional T5 bucket of rel = k_pos - q_pos (xtransformers.py:155-176)."""
    n = -rel
    nb = num_buckets // 2
    ret = (n < 0).long() * nb
    n = torch.abs(n)
    max_exact = nb // 2
    is_small = n < max_exact
    val_if_large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                                * (nb - max_exact)).long()
    val_if_large = torch.min(val_if_large, torch.full_like(val_if_large, nb - 1))
    return ret + torch.where(is_small, n, val_if_large)


def rel_pos_bias(emb_weight, T, scale):
    """[H, T, T] additive bias = scale * E[bucket(j - i), h]."""
    q = torch.arange(T)
    rel = q[None, :] - q[:, None]
    b = rel_pos_bucket(rel)
    return emb_weight[b].permute(2, 0, 1) * scale


def attention_block(sd, p, x, heads, rel_pos=True):
    """AttentionBlock.forward (arch_util.py:117-123), x [B, C, T]."""
    B, C, T = x.shape
    qkv = F.conv1d(_gn(x, sd[p + "norm.weight"], sd[p + "norm.bias"]), sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    ch = C // heads
    q, k, v = qkv.reshape(B * heads, ch * 3, T).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    if rel_pos:
        bias = rel_pos_bias(sd[p + "relative_pos_embeddings.relative_attention_bias.weight"], T, ch ** 0.5)
        w = (w.reshape(B, heads, T, T) + bias.unsqueeze(0)).reshape(B * heads, T, T)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(B, -1, T)
    return x + F.conv1d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def resblock(sd, p, x, time_emb):
    """ResBlock.forward with use_scale_shift_norm (diffusion_decoder.py:107-120)."""
    h = F.conv1d(F.silu(_gn(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"])),
                 sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"])
    e = F.linear(F.silu(time_emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])[..., None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = _gn(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"]) * (1 + scale) + shift
    h = F.conv1d(F.silu(h), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    return x + h


def diffusion_layer(sd, p, x, time_emb, heads):
    return attention_block(sd, p + "attn.", resblock(sd, p + "resblk.", x, time_emb), heads)


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def timestep_independent(sd, cfg, latents, cond_latent, S):
    """latents [1, N, ar_dim], cond_latent [1, 2*C] -> code_emb [1, C, S] (diffusion_decoder.py:232-260)."""
    x = latents.permute(0, 2, 1)
    cond_scale, cond_shift = torch.chunk(cond_latent, 2, dim=1)
    x = F.conv1d(x, sd["latent_conditioner.0.weight"], sd["latent_conditioner.0.bias"], padding=1)
    for i in range(1, 5):
        x = attention_block(sd, f"latent_conditioner.{i}.", x, cfg.diff_heads)
    x = _gn(x, sd["code_norm.weight"], sd["code_norm.bias"]) * (1 + cond_scale.unsqueeze(-1)) + cond_shift.unsqueeze(-1)
    return F.interpolate(x, size=S, mode="nearest")


def forward(sd, cfg, x, t_orig, code_emb=None, conditioning_free=False):
    """DiffusionTts.forward (diffusion_decoder.py:262-322). x [B,100,S], t_orig LongTensor [B]."""
    C = cfg.diff_dim
    if conditioning_free:
        code_emb = sd["unconditioned_embedding"].repeat(x.shape[0], 1, x.shape[-1])
    te = timestep_embedding(t_orig, C)
    time_emb = F.linear(F.silu(F.linear(te, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                        sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    for i in range(3):
        code_emb = diffusion_layer(sd, f"conditioning_timestep_integrator.{i}.", code_emb, time_emb, cfg.diff_heads)
    h = F.conv1d(x, sd["inp_block.weight"], sd["inp_block.bias"], padding=1)
    h = torch.cat([h, code_emb], dim=1)
    h = F.conv1d(h, sd["integrating_conv.weight"], sd["integrating_conv.bias"])
    for i in range(cfg.diff_layers):
        h = diffusion_layer(sd, f"layers.{i}.", h, time_emb, cfg.diff_heads)
    for i in range(cfg.diff_layers, cfg.diff_layers + 3):
        h = resblock(sd, f"layers.{i}.", h, time_emb)
    h = F.silu(_gn(h, sd["out.0.weight"], sd["out.0.bias"]))
    return F.conv1d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def p_sample_loop(sd, cfg, code_emb, noise0, step_noise, iters, cond_free=True, cond_free_k=2.0,
                  return_trace=False):
    """SpacedDiffusion.p_sample_loop with injected randomness.
    noise0 [1,100,S] (already multiplied by the temperature, api.py:126);
    step_noise [iters,1,100,S]: the randn_like drawn at loop index i = iters-1 ... 0 is step_noise[iters-1-i]
    (i.e. in call order). Returns x_0 [1,100,S]."""
    sch = make_schedule(iters)
    f32 = {k: torch.from_numpy(v.astype(np.float32)) for k, v in sch.items() if k != "timestep_map"}
    tmap = sch["timestep_map"]
    n = len(tmap)
    x = noise0
    cin = cfg.diff_in_channels
    trace = []
    for call, i in enumerate(reversed(range(n))):
        t_orig = torch.tensor([int(tmap[i])], dtype=torch.long)
        out = forward(sd, cfg, x, t_orig, code_emb=code_emb)
        eps, var = torch.split(out, cin, dim=1)
        if cond_free:
            out_u = forward(sd, cfg, x, t_orig, conditioning_free=True)
            eps_u = out_u[:, :cin]
            cfk = cond_free_k * (1 - i / n)
            eps = (1 + cfk) * eps - cfk * eps_u
        min_log = f32["posterior_log_variance_clipped"][i]
        max_log = f32["log_betas"][i]
        frac = (var + 1) / 2
        logvar = frac * max_log + (1 - frac) * min_log
        x0 = (f32["sqrt_recip_alphas_cumprod"][i] * x - f32["sqrt_recipm1_alphas_cumprod"][i] * eps).clamp(-1, 1)
        mean = f32["posterior_mean_coef1"][i] * x0 + f32["posterior_mean_coef2"][i] * x
        nz = 0.0 if i == 0 else 1.0
        x = mean + nz * torch.exp(0.5 * logvar) * step_noise[call]
        if return_trace:
            trace.append(x.clone())
    return (x, trace) if return_trace else x


def denormalize_tacotron_mel(x):
    return ((x + 1) / 2) * (TACOTRON_MEL_MAX - TACOTRON_MEL_MIN) + TACOTRON_MEL_MIN


def output_seq_len(n_latents):
    return n_latents * 4 * 24000 // 22050


def spectrogram_diffusion(sd, cfg, latents, cond_latent, noise0, step_noise, iters, cond_free=True, cond_free_k=2.0):
    """do_spectrogram_diffusion (api.py:117-130) with injected randomness -> mel [1,100,S]."""
    S = output_seq_len(latents.shape[1])
    code_emb = timestep_independent(sd, cfg, latents, cond_latent, S)
    x = p_sample_loop(sd, cfg, code_emb, noise0, step_noise, iters, cond_free, cond_free_k)
    return denormalize_tacotron_mel(x)[:, :, :S]
