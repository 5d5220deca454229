"""TEST INFRASTRUCTURE ONLY — a torch-CPU emulation of the libttb.so entry points (same signatures as
tortoise_tts_b200/lib.py), used to check the HOST orchestration (weight packing, op order, layouts, strides) of the
stage engines against the oracle without a GPU.  It mirrors the kernels' numerics where that matters (bf16 rounding
of GEMM operands / outputs).  Never imported by the product path."""
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU_NEW, ACT_SILU, ACT_GEGLU, ACT_LRELU02, ACT_TANH = 0, 1, 2, 3, 4, 5


def _v(t, sizes, strides):
    return torch.as_strided(t, sizes, strides, t.storage_offset())


def gemm(A, W, *, M, N, K, bias=None, residual=None, out_f32=None, out_bf16=None, lda=None, rows=None, batch=1,
         a_bstride=0, res_bstride=0, outf_bstride=0, outb_bstride=0, ldr=None, ldo=None, ldob=None, taps=1, pad=0,
         act=ACT_NONE, alpha=1.0, tile_n=0, force_ref=False, splitk=1, cluster=0, variant=0, gn_partials=None, gn_groups=0, tap_dilation=1, w_static=False):
    n_out = N // 2 if act == ACT_GEGLU else N
    lda = K if lda is None else lda
    rows = M if rows is None else rows
    ldr = n_out if ldr is None else ldr
    ldo = n_out if ldo is None else ldo
    ldob = n_out if ldob is None else ldob
    a = _v(A, (batch, rows, K), (a_bstride, lda, 1)).float()
    w = W.float().reshape(N, taps, K)
    acc = torch.zeros(batch, M, N)
    for tap in range(taps):
        sh = tap * max(tap_dilation, 1) - pad
        lo, hi = max(0, -sh), min(M, rows - sh)
        if hi > lo:
            acc[:, lo:hi] += a[:, lo + sh:hi + sh] @ w[:, tap].t()
    acc = acc * alpha
    if bias is not None:
        acc = acc + bias
    if act == ACT_GEGLU:
        acc = acc[..., 0::2] * F.gelu(acc[..., 1::2])
    elif act == ACT_GELU_NEW:
        acc = F.gelu(acc, approximate="tanh")
    elif act == ACT_SILU:
        acc = F.silu(acc)
    elif act == ACT_LRELU02:
        acc = F.leaky_relu(acc, 0.2)
    elif act == ACT_TANH:
        acc = torch.tanh(acc)
    if residual is not None and act != ACT_GEGLU:
        acc = acc + _v(residual, (batch, M, n_out), (res_bstride, ldr, 1))
    if splitk > 1:   # raw partials: the whole sum in split 0, zeros elsewhere (only the sum is observable)
        kb_total = (K // 64) * taps
        per = (kb_total + splitk - 1) // splitk
        nz = (kb_total + per - 1) // per
        o = _v(out_f32, (nz, M, n_out), (outf_bstride, ldo, 1))
        o.zero_()
        o[0].copy_(acc[0])
        return
    if out_f32 is not None:
        _v(out_f32, (batch, M, n_out), (outf_bstride, ldo, 1)).copy_(acc)
    if out_bf16 is not None:
        _v(out_bf16, (batch, M, n_out), (outb_bstride, ldob, 1)).copy_(acc.to(torch.bfloat16))
    if gn_partials is not None:
        # (sum, sum of squares) per (batch item, group of 32 columns, block of 32 rows), scratch layout of csrc/norm.cu
        assert N == 32 * gn_groups and (M + 31) // 32 <= 128 and act != ACT_GEGLU and splitk <= 1
        nrb = (M + 31) // 32
        pad_rows = nrb * 32 - M
        blk = F.pad(acc, (0, 0, 0, pad_rows)).reshape(batch, nrb, 32, gn_groups, 32)
        part = _v(gn_partials[16:], (batch, gn_groups, 128, 2), (gn_groups * 256, 256, 2, 1))
        part[:, :, :nrb, 0] = blk.sum(dim=(2, 4)).transpose(1, 2)
        part[:, :, :nrb, 1] = (blk * blk).sum(dim=(2, 4)).transpose(1, 2)


def layernorm(x, M, D, g1, b1, g2=None, b2=None, out_bf16=None, out_f32=None):
    y = F.layer_norm(_v(x, (M, D), (D, 1)), (D,), g1, b1, 1e-5)
    if g2 is not None:
        y = F.layer_norm(y, (D,), g2, b2, 1e-5)
    if out_bf16 is not None:
        _v(out_bf16, (M, D), (D, 1)).copy_(y.to(torch.bfloat16))
    if out_f32 is not None:
        _v(out_f32, (M, D), (D, 1)).copy_(y)


def residual_layernorm(x, M, D, partials, nsplit, split_stride, bias, g1, b1, g2=None, b2=None, out_bf16=None,
                       out_f32=None):
    xx = _v(x, (M, D), (D, 1))
    t = _v(partials, (nsplit, M, D), (split_stride, D, 1)).sum(dim=0)
    xx += t + (bias if bias is not None else 0.0)
    layernorm(x, M, D, g1, b1, g2, b2, out_bf16, out_f32)


def rmsnorm(x, M, D, g, out_bf16):
    xx = _v(x, (M, D), (D, 1))
    norm = torch.norm(xx, dim=-1, keepdim=True) * (D ** -0.5)
    _v(out_bf16, (M, D), (D, 1)).copy_((xx / norm.clamp(min=1e-8) * g).to(torch.bfloat16))


def groupnorm_scratch(B, groups, device):
    return torch.zeros(B * groups * (2 * 128 + 2) + 16, dtype=torch.float32, device=device)


def groupnorm(x, B, S, Cc, groups, gamma, beta, partials, scale_shift=None, ss_bstride=0, ss_row=None, ss_row_stride=0,
              silu=False, out_bf16=None, ldo=0, out_f32=None, ldof=0):
    xx = _v(x, (B, S, Cc), (S * Cc, Cc, 1)).transpose(1, 2)
    y = F.group_norm(xx, groups, gamma, beta, 1e-5)
    if scale_shift is not None:
        off = int(ss_row[0]) * ss_row_stride if ss_row is not None else 0
        ss = _v(scale_shift, (B, 2 * Cc), (ss_bstride, 1))
        ss = torch.as_strided(scale_shift, (B, 2 * Cc), (ss_bstride, 1), scale_shift.storage_offset() + off)
        y = y * (1 + ss[:, :Cc, None]) + ss[:, Cc:, None]
    if silu:
        y = F.silu(y)
    y = y.transpose(1, 2)
    if out_bf16 is not None:
        _v(out_bf16, (B, S, Cc), (S * ldo, ldo, 1)).copy_(y.to(torch.bfloat16))
    if out_f32 is not None:
        _v(out_f32, (B, S, Cc), (S * ldof, ldof, 1)).copy_(y)


def groupnorm_apply(x, B, S, Cc, groups, gamma, beta, partials, scale_shift=None, ss_bstride=0, ss_row=None,
                    ss_row_stride=0, silu=False, out_bf16=None, ldo=0, out_f32=None, ldof=0):
    """Uses the statistics the producing gemm(..., gn_partials=) left behind (NOT recomputed from x: a stale-partials
    bug in the host logic must show up as a mismatch)."""
    assert Cc == 32 * groups
    nrb = (S + 31) // 32
    part = _v(partials[16:], (B, groups, 128, 2), (groups * 256, 256, 2, 1))[:, :, :nrb].sum(dim=2)
    n = float(S * 32)
    mean = part[..., 0] / n
    rstd = torch.rsqrt((part[..., 1] / n - mean * mean).clamp(min=0) + 1e-5)
    xx = _v(x, (B, S, groups, 32), (S * Cc, Cc, 32, 1))
    y = ((xx - mean[:, None, :, None]) * rstd[:, None, :, None]).reshape(B, S, Cc) * gamma + beta
    if scale_shift is not None:
        off = int(ss_row[0]) * ss_row_stride if ss_row is not None else 0
        ss = torch.as_strided(scale_shift, (B, 2 * Cc), (ss_bstride, 1), scale_shift.storage_offset() + off)
        y = y * (1 + ss[:, None, :Cc]) + ss[:, None, Cc:]
    if silu:
        y = F.silu(y)
    if out_bf16 is not None:
        _v(out_bf16, (B, S, Cc), (S * ldo, ldo, 1)).copy_(y.to(torch.bfloat16))
    if out_f32 is not None:
        _v(out_f32, (B, S, Cc), (S * ldof, ldof, 1)).copy_(y)


def attention(qkv, out, *, nseq, T, H, ld, ldo, k_off, v_off, scale, causal=False, bias=None, bias_sat=0, head_dim=0):
    hd = head_dim or 64
    base = _v(qkv, (nseq, T, ld), (T * ld, ld, 1)).float()
    q = base[..., :H * hd].reshape(nseq, T, H, hd).transpose(1, 2)
    k = base[..., k_off:k_off + H * hd].reshape(nseq, T, H, hd).transpose(1, 2)
    v = base[..., v_off:v_off + H * hd].reshape(nseq, T, H, hd).transpose(1, 2)
    w = (q @ k.transpose(-1, -2)) * scale
    if bias is not None:
        i = torch.arange(T)
        w = w + bias[:, i[None, :] - i[:, None] + T - 1].unsqueeze(0)
    if causal:
        w = w.masked_fill(~torch.ones(T, T, dtype=torch.bool).tril(), float("-inf"))
    o = (torch.softmax(w, -1) @ v).transpose(1, 2).reshape(nseq, T, H * hd)
    _v(out, (nseq, T, H * hd), (T * ldo, ldo, 1)).copy_(o.to(torch.bfloat16))


def ar_embed_step(codes, ld_codes, state, mel_emb, mel_pos, B, D, pos_mode, x):
    j = int(state[0])
    tok = _v(codes, (B, ld_codes), (ld_codes, 1))[:, j - 1].long()
    _v(x, (B, D), (D, 1)).copy_(mel_emb[tok] + mel_pos[j + 1 if pos_mode else j])


def ar_decode_attention(qkv, pk, pv, ck, cv, state, B, H, P, Nmax, out, scratch_o=None, scratch_lse=None):
    D = H * 64
    slot = int(state[0]) - 1
    r = _v(qkv, (B, 3 * D), (3 * D, 1))
    ckv = _v(ck, (B, H, Nmax, 64), (H * Nmax * 64, Nmax * 64, 64, 1))
    cvv = _v(cv, (B, H, Nmax, 64), (H * Nmax * 64, Nmax * 64, 64, 1))
    ckv[:, :, slot] = r[:, D:2 * D].reshape(B, H, 64)
    cvv[:, :, slot] = r[:, 2 * D:].reshape(B, H, 64)
    q = r[:, :D].reshape(B, H, 1, 64).float() * 0.125
    pkk = _v(pk, (H, P, 64), (P * 64, 64, 1)).float().unsqueeze(0).expand(B, -1, -1, -1)
    pvv = _v(pv, (H, P, 64), (P * 64, 64, 1)).float().unsqueeze(0).expand(B, -1, -1, -1)
    K = torch.cat([pkk, ckv[:, :, :slot + 1].float()], dim=2)
    V = torch.cat([pvv, cvv[:, :, :slot + 1].float()], dim=2)
    w = torch.softmax(q @ K.transpose(-1, -2), -1)
    _v(out, (B, D), (D, 1)).copy_((w @ V).reshape(B, D).to(torch.bfloat16))


def ar_store_prefix(qkv, P, H, pk, pv):
    D = H * 64
    r = _v(qkv, (P, 3 * D), (3 * D, 1))
    _v(pk, (H, P, 64), (P * 64, 64, 1)).copy_(r[:, D:2 * D].reshape(P, H, 64).transpose(0, 1))
    _v(pv, (H, P, 64), (P * 64, 64, 1)).copy_(r[:, 2 * D:].reshape(P, H, 64).transpose(0, 1))


def ar_step_store_prefix(qkv, P, H, pkv):
    D = H * 64
    r = _v(qkv, (P, 3 * D), (3 * D, 1))
    o = _v(pkv, (H, P, 2, 64), (P * 128, 128, 64, 1))
    o[:, :, 0].copy_(r[:, D:2 * D].reshape(P, H, 64).transpose(0, 1))
    o[:, :, 1].copy_(r[:, 2 * D:].reshape(P, H, 64).transpose(0, 1))


def ar_step_supported(B, D, H, P):
    return 0 < B <= 256 and 0 < P <= 352 and D == 64 * H and D % 128 == 0 and D <= 1024


class ArStep:
    """Emulation of the one-kernel decode step (csrc/ar_step.cu): same arguments, same buffers, same KV layout."""

    def __init__(self, **kw):
        self.kw = kw

    def step(self, phase_mask=0, layer_begin=0, layer_end=0):
        k = self.kw
        B, D, H, L, V, P, Nmax = k["B"], k["D"], k["H"], k["L"], k["V"], k["P"], k["Nmax"]
        st = k["state"]
        j = int(st[0])
        slot = j - 1
        if phase_mask == 4:          # attention phase of one layer (the "mixed" mode of AREngine)
            assert layer_end == layer_begin + 1
            l = layer_begin
            pkv = k["prefix_kv"].view(L, H, P, 2, 64)
            ckv = k["cand_kv"].view(L, B, H, Nmax, 2, 64)
            qkv = k["qkv"].view(B, 3 * D)
            ckv[l, :, :, slot, 0] = qkv[:, D:2 * D].reshape(B, H, 64)
            ckv[l, :, :, slot, 1] = qkv[:, 2 * D:].reshape(B, H, 64)
            q = qkv[:, :D].reshape(B, H, 1, 64).float() * 0.125
            Kc = torch.cat([pkv[l, :, :, 0].float().unsqueeze(0).expand(B, -1, -1, -1), ckv[l, :, :, :slot + 1, 0].float()], 2)
            Vc = torch.cat([pkv[l, :, :, 1].float().unsqueeze(0).expand(B, -1, -1, -1), ckv[l, :, :, :slot + 1, 1].float()], 2)
            k["o"].view(B, D).copy_((torch.softmax(q @ Kc.transpose(-1, -2), -1) @ Vc).reshape(B, D).to(torch.bfloat16))
            return
        assert phase_mask == 0 and layer_end == 0, "the emulation runs whole steps or one attention phase"
        tok = k["codes"].view(B, k["ld_codes"])[:, j - 1].long()
        x = k["mel_emb"][tok] + k["mel_pos"][j + 1 if k["pos_mode"] else j]

        def ln(v, g, b):
            return F.layer_norm(v, (D,), g, b, 1e-5)

        def mm(a, w, b=None):                       # bf16 operands, fp32 accumulate
            y = a.to(torch.bfloat16).float() @ w.float().t()
            return y if b is None else y + b
        pkv = k["prefix_kv"].view(L, H, P, 2, 64)
        ckv = k["cand_kv"].view(L, B, H, Nmax, 2, 64)
        for l, lw in enumerate(k["layers"]):
            qkv = mm(ln(x, lw["ln1_g"], lw["ln1_b"]), lw["wqkv"], lw["bqkv"]).to(torch.bfloat16)
            ckv[l, :, :, slot, 0] = qkv[:, D:2 * D].reshape(B, H, 64)
            ckv[l, :, :, slot, 1] = qkv[:, 2 * D:].reshape(B, H, 64)
            q = qkv[:, :D].reshape(B, H, 1, 64).float() * 0.125
            Kc = torch.cat([pkv[l, :, :, 0].float().unsqueeze(0).expand(B, -1, -1, -1), ckv[l, :, :, :slot + 1, 0].float()], 2)
            Vc = torch.cat([pkv[l, :, :, 1].float().unsqueeze(0).expand(B, -1, -1, -1), ckv[l, :, :, :slot + 1, 1].float()], 2)
            o = (torch.softmax(q @ Kc.transpose(-1, -2), -1) @ Vc).reshape(B, D).to(torch.bfloat16)
            x = x + mm(o, lw["wproj"], lw["bproj"])
            hmid = F.gelu(mm(ln(x, lw["ln2_g"], lw["ln2_b"]), lw["wfc"], lw["bfc"]), approximate="tanh").to(torch.bfloat16)
            x = x + mm(hmid, lw["wproj2"], lw["bproj2"])
        hn = ln(ln(x, k["lnf_g"], k["lnf_b"]), k["fn_g"], k["fn_b"])
        k["x"].view(B, D).copy_(x)
        k["logits"].view(B, V).copy_(mm(hn, k["w_head"], k["b_head"]))


def ar_sample(logits, ld_logits, V, B, uniforms, ld_u, seen, codes, ld_codes, finished, state, temperature, top_k,
              top_p, rep_penalty, stop_token, advance=True):
    from oracle.ar import sample_step
    step = int(state[0])
    cd = _v(codes, (B, ld_codes), (ld_codes, 1))
    for b in range(B):
        if int(finished[b]):
            cd[b, step] = stop_token
            continue
        row = torch.as_strided(logits, (V,), (1,), logits.storage_offset() + b * ld_logits)
        prev = [w * 32 + bit for w in range(seen.shape[1]) for bit in range(32)
                if (int(seen[b, w]) >> bit) & 1] if True else []
        tok, _, _ = sample_step(row, prev, float(uniforms.reshape(-1)[b * ld_u + step]), temperature, top_k, top_p,
                                rep_penalty)
        cd[b, step] = tok
        w, bit = divmod(tok, 32)
        seen[b, w] = int(seen[b, w]) | ((1 << bit) if bit < 31 else -(1 << 31))
        if tok == stop_token:
            finished[b] = 1
    if advance:
        state[0] += 1
        state[1] = int(bool(finished.all()))


def ar_fix_codes(codes, B, L, stop_token, trim_len):
    from oracle.ar import fix_autoregressive_output, calm_trim_length
    cd = _v(codes, (B, L), (L, 1))
    for b in range(B):
        cd[b] = fix_autoregressive_output(cd[b].long(), stop_token).to(cd.dtype)
        trim_len[b] = calm_trim_length(cd[b])


def embed(ids, pos, n, D, table, pos_table, out):
    y = table[ids.long()[:n]]
    if pos is not None and pos_table is not None:
        y = y + pos_table[pos.long()[:n]]
    _v(out, (n, D), (D, 1)).copy_(y)


def clvp_rotary(qkv, nseq, T, H):
    x = _v(qkv, (nseq, T, 3 * H, 64), (T * 3 * H * 64, 3 * H * 64, 64, 1))
    xf = x.float()
    inv = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
    f = torch.einsum("i,j->ij", torch.arange(T).float(), inv)
    cs, sn = f.cos()[None, :, None, :], f.sin()[None, :, None, :]
    x1, x2 = xf[..., :16], xf[..., 16:32]
    x[..., :16] = (x1 * cs - x2 * sn).to(torch.bfloat16)
    x[..., 16:32] = (x2 * cs + x1 * sn).to(torch.bfloat16)


def clvp_pool(x, nseq, T, D, g, b, out):
    y = F.layer_norm(_v(x, (nseq, T, D), (T * D, D, 1)), (D,), g, b, 1e-5).mean(dim=1)
    _v(out, (nseq, D), (D, 1)).copy_(y)


def clvp_project(pooled, n, D, W, latents, text_latent, temp_exp, scores):
    y = F.normalize(_v(pooled, (n, D), (D, 1)) @ W.t(), p=2, dim=-1)
    if latents is not None:
        _v(latents, (n, D), (D, 1)).copy_(y)
    if text_latent is not None:
        scores[:n] = (y @ text_latent.reshape(-1)) * temp_exp


def timestep_embedding(t, n, Cc, out):
    half = Cc // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:n, None].float() * freqs[None]
    _v(out, (n, Cc), (Cc, 1)).copy_(torch.cat([torch.cos(args), torch.sin(args)], dim=-1))


def linear_small(x, M, K, W, b, N, out, silu_in=False, silu_out=False):
    xx = _v(x, (M, K), (K, 1))
    y = F.linear(F.silu(xx) if silu_in else xx, W, b)
    _v(out, (M, N), (N, 1)).copy_(F.silu(y) if silu_out else y)


def interp_nearest(x, N, S, Cc, out_bf16=None, ldo=0, out_f32=None, ldof=0):
    y = F.interpolate(_v(x, (N, Cc), (Cc, 1)).t().unsqueeze(0), size=S, mode="nearest")[0].t()
    if out_bf16 is not None:
        _v(out_bf16, (S, Cc), (ldo, 1)).copy_(y.to(torch.bfloat16))
    if out_f32 is not None:
        _v(out_f32, (S, Cc), (ldof, 1)).copy_(y)


def diffusion_step(model_out, out_bstride, ld_out, x, x_bf16, ld_xb, noise, tables, step, S, Cc, iters, cond_free,
                   cond_free_k, mel_out=None, parity_stride=0):
    assert parity_stride == 0, "the peer-exchange area exists on GPUs only"
    call = int(step[0])
    i = iters - 1 - call
    tb = _v(tables, (6, iters), (iters, 1))
    mo = _v(model_out, (2 if cond_free else 1, S, 2 * Cc), (out_bstride, ld_out, 1))
    eps, var = mo[0, :, :Cc], mo[0, :, Cc:]
    if cond_free:
        cfk = cond_free_k * (1 - i / iters)
        eps = (1 + cfk) * eps - cfk * mo[1, :, :Cc]
    frac = (var + 1) / 2
    logvar = frac * tb[3, i] + (1 - frac) * tb[2, i]
    xx = _v(x, (S, Cc), (Cc, 1))
    x0 = (tb[0, i] * xx - tb[1, i] * eps).clamp(-1, 1)
    mean = tb[4, i] * x0 + tb[5, i] * xx
    xn = mean + (0.0 if i == 0 else 1.0) * torch.exp(0.5 * logvar) * _v(noise, (iters, S, Cc), (S * Cc, Cc, 1))[call]
    xx.copy_(xn)
    if x_bf16 is not None:
        _v(x_bf16, (S, Cc), (ld_xb, 1)).copy_(xn.to(torch.bfloat16))
    if mel_out is not None and i == 0:
        _v(mel_out, (Cc, S), (S, 1)).copy_((((xn + 1) / 2) * (2.3143386840820312 + 11.512925148010254)
                                            - 11.512925148010254).t())


def counter_add(counter, delta):
    counter[0] += delta


def transpose_f32(inp, R, Cc, out):
    _v(out, (Cc, R), (R, 1)).copy_(_v(inp, (R, Cc), (Cc, 1)).t())


def cast_pad_bf16(inp, R, Cc, ld_in, out, ldo, ncols_out=None):
    nc = ldo if ncols_out is None else ncols_out
    o = _v(out, (R, nc), (ldo, 1))
    o.zero_()
    o[:, :Cc] = _v(inp, (R, Cc), (ld_in, 1)).to(torch.bfloat16)


def broadcast_rows(row, R, Cc, out_f32, out_bf16, ldo):
    if out_f32 is not None:
        _v(out_f32, (R, Cc), (ldo, 1)).copy_(row[:Cc].expand(R, -1))
    if out_bf16 is not None:
        _v(out_bf16, (R, Cc), (ldo, 1)).copy_(row[:Cc].expand(R, -1).to(torch.bfloat16))


def voc_conv1d(x, Cin, L, w, b, Cout, ksize, out, dilation=1, reflect=False, lrelu_in=1.0, lrelu_out=1.0,
               tanh_out=False, residual=None):
    xi = _v(x, (Cin, L), (L, 1)).unsqueeze(0)
    if lrelu_in != 1.0:
        xi = F.leaky_relu(xi, lrelu_in)
    pad = dilation * (ksize // 2)
    xi = F.pad(xi, (pad, pad), mode="reflect" if reflect else "constant")
    y = F.conv1d(xi, w, b, dilation=dilation)[0]
    if lrelu_out != 1.0:
        y = F.leaky_relu(y, lrelu_out)
    if tanh_out:
        y = torch.tanh(y)
    if residual is not None:
        y = y + _v(residual, (Cout, L), (L, 1))
    _v(out, (Cout, L), (L, 1)).copy_(y)


def voc_convt(x, Cc, L, w, b, stride, lrelu_in, out):
    y = F.conv_transpose1d(F.leaky_relu(_v(x, (Cc, L), (L, 1)), lrelu_in).unsqueeze(0), w, b, stride=stride,
                           padding=stride // 2 + stride % 2, output_padding=stride % 2)[0]
    _v(out, (Cc, L * stride), (L * stride, 1)).copy_(y)


def voc_lvc_gate(y, Cc, L, hop, kernels, ldk, koff, bias, ldb, boff, x):
    from oracle.vocoder import lvc
    Fr = L // hop
    K = torch.as_strided(kernels, (Fr, Cc, 3, 2 * Cc), (ldk, 3 * 2 * Cc, 2 * Cc, 1), kernels.storage_offset() + koff)
    Bi = torch.as_strided(bias, (Fr, 2 * Cc), (ldb, 1), bias.storage_offset() + boff)
    o = lvc(_v(y, (Cc, L), (L, 1)), K.permute(1, 3, 2, 0), Bi.t(), hop)
    xx = _v(x, (Cc, L), (L, 1))
    xx += torch.sigmoid(o[:Cc]) * torch.tanh(o[Cc:])


def voc_to_tokens_bf16(x, Cc, L, out, ldo, split=False):
    o = _v(out, (L, ldo), (ldo, 1))
    xt = _v(x, (Cc, L), (L, 1)).t()
    hi = xt.to(torch.bfloat16)
    if not split:
        o.zero_()
        o[:, :Cc] = hi
    else:
        o[:, :Cc] = hi
        o[:, Cc:2 * Cc] = (xt - hi.float()).to(torch.bfloat16)
        o[:, 2 * Cc:3 * Cc] = hi


def audio_resample(x, n, kernels, down, up, klen, width, out, m):
    xp = F.pad(x[:n].reshape(1, 1, -1), (width, width + down))
    y = F.conv1d(xp, kernels.reshape(up, 1, klen), stride=down).transpose(1, 2).reshape(-1)
    out[:m].copy_(y[:m])


def audio_stft_mel(x, n, n_fft, hop, window, twiddle, fb, n_mels, power, clip, floor_v, div, out_bf16=None, ldo=0,
                   out_f32=None):
    v = x[:n].clamp(-1, 1) if clip else x[:n]
    xp = F.pad(v.reshape(1, 1, -1), (n_fft // 2, n_fft // 2), mode="reflect").reshape(-1)
    fr = xp.unfold(0, n_fft, hop) * window                      # [frames, n_fft]
    idx = (torch.arange(n_fft // 2 + 1)[:, None] * torch.arange(n_fft)[None, :]) % n_fft
    re = fr @ twiddle[:, 0][idx].t()
    im = -(fr @ twiddle[:, 1][idx].t())
    p2 = re * re + im * im
    spec = p2 if power == 2 else p2.sqrt()
    y = torch.log((spec @ fb.t()).clamp(min=floor_v))
    if div is not None:
        y = y / div
    frames = y.shape[0]
    if out_f32 is not None:
        _v(out_f32, (n_mels, frames), (frames, 1)).copy_(y.t())
    if out_bf16 is not None:
        o = _v(out_bf16, (frames, ldo), (ldo, 1))
        o.zero_()
        o[:, :n_mels] = y.to(torch.bfloat16)


def act_split_cast(a, R, Cc, out, ldo, b=None, c=None, scale=1.0, slope=1.0):
    v = _v(a, (R, Cc), (Cc, 1)).clone()
    for t in (b, c):
        if t is not None:
            v = v + _v(t, (R, Cc), (Cc, 1))
    v = v * scale
    v = torch.where(v > 0, v, v * slope)
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    o = _v(out, (R, ldo), (ldo, 1))
    o.zero_()
    o[:, :Cc], o[:, Cc:2 * Cc], o[:, 2 * Cc:3 * Cc] = hi, lo, hi


def interp_linear(x, N, Cc, rscale, S, out):
    src = ((torch.arange(S, dtype=torch.float32) + 0.5) * rscale - 0.5).clamp(min=0)
    i0 = src.floor().long().clamp(max=N - 1)
    i1 = (i0 + 1).clamp(max=N - 1)
    l1 = (src - i0.float()).unsqueeze(1)
    xx = _v(x, (N, Cc), (Cc, 1))
    _v(out, (S, Cc), (Cc, 1)).copy_(xx[i0] * (1 - l1) + xx[i1] * l1)


def mean_rows(x, R, Cc, ld, scale, out, accumulate=False):
    s = _v(x, (R, Cc), (ld, 1)).sum(0) * scale
    out[:Cc].copy_(out[:Cc] + s if accumulate else s)


def equal_linear(x, K, W, b, N, out, wscale=1.0, bscale=1.0, slope=1.0, gain=1.0):
    y = x[:K] @ (W.reshape(N, K) * wscale).t() + (0 if b is None else b * bscale)
    out[:N].copy_(F.leaky_relu(y, slope) * gain)


def load():
    return None


def install():
    """Monkeypatch tortoise_tts_b200.lib with this emulation (tests only)."""
    import sys
    import tortoise_tts_b200.lib as real
    me = sys.modules[__name__]
    saved = {}
    for name in dir(me):
        if name.startswith("_") or name in ("install", "torch", "F", "math"):
            continue
        if hasattr(real, name) and (callable(getattr(me, name)) or isinstance(getattr(me, name), type)):
            saved[name] = getattr(real, name)
            setattr(real, name, getattr(me, name))
    return saved


def uninstall(saved):
    import tortoise_tts_b200.lib as real
    for k, v in saved.items():
        setattr(real, k, v)
