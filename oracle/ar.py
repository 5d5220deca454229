"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of the reference's autoregressive path.

Plain torch-fp32 functional code over the reference `autoregressive.pth` state_dict; no
reference import.  Validated against the reference modules by tests/test_oracle_vs_reference.py
(build container only) and pinned by tests/golden/*.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this.

Follows:
  * UnifiedVoice.inference_speech prompt construction      autoregressive.py:535-549
  * GPT2InferenceModel.forward (kv_cache position rule)    autoregressive.py:134-149
  * HF GPT2Block (transformers 4.31, pinned requirements.txt:3): ln_1 -> c_attn -> causal
    attention (1/sqrt(64)) -> c_proj -> +x ; ln_2 -> c_fc -> gelu_new -> c_proj -> +h
  * lm_head = Sequential(final_norm, mel_head) after gpt.ln_f   autoregressive.py:42,174
  * HF sample(): RepetitionPenalty -> Temperature -> TopK(50) -> TopP -> softmax -> multinomial
    (in-tree 4.31 copy stream_generator.py:943-1000)
  * fix_autoregressive_output                               api.py:87-114
  * UnifiedVoice.forward(return_latent=True, clip_inputs=False)  autoregressive.py:454-512
"""
import math

import torch
import torch.nn.functional as F


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def gpt2_block(sd, l, x, heads, past_kv=None):
    """x: [B, T, D]; past_kv: (k [B,H,Tp,64], v) or None. Returns y, (k, v) (full)."""
    p = f"gpt.h.{l}."
    B, T, D = x.shape
    hd = D // heads
    a = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
    qkv = a @ sd[p + "attn.c_attn.weight"] + sd[p + "attn.c_attn.bias"]
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, T, heads, hd).transpose(1, 2)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=2)
        v = torch.cat([past_kv[1], v], dim=2)
    Tk = k.shape[2]
    w = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    causal = torch.ones(Tk, Tk, dtype=torch.bool).tril()[Tk - T:Tk]
    w = w.masked_fill(~causal, torch.finfo(w.dtype).min)
    w = torch.softmax(w, dim=-1)
    o = (w @ v).transpose(1, 2).reshape(B, T, D)
    h = x + (o @ sd[p + "attn.c_proj.weight"] + sd[p + "attn.c_proj.bias"])
    m = _ln(h, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
    m = gelu_new(m @ sd[p + "mlp.c_fc.weight"] + sd[p + "mlp.c_fc.bias"])
    y = h + (m @ sd[p + "mlp.c_proj.weight"] + sd[p + "mlp.c_proj.bias"])
    return y, (k, v)


def gpt2_trunk(sd, cfg, emb, past=None):
    """30 blocks + ln_f. emb: [B,T,D]. Returns hidden [B,T,D], list of (k,v)."""
    x = emb
    new_past = []
    for l in range(cfg.ar_layers):
        x, kv = gpt2_block(sd, l, x, cfg.ar_heads, None if past is None else past[l])
        new_past.append(kv)
    x = _ln(x, sd["gpt.ln_f.weight"], sd["gpt.ln_f.bias"])
    return x, new_past


def mel_logits(sd, hidden):
    h = _ln(hidden, sd["final_norm.weight"], sd["final_norm.bias"])
    return h @ sd["mel_head.weight"].t() + sd["mel_head.bias"]


def text_ids(cfg, text_tokens):
    """`text_tokens` is what api.py hands to the model, i.e. the tokenizer ids ALREADY padded with one 0
    (api.py:391); inference_speech / forward pad another 0 and prepend start (autoregressive.py:538-539,485-489)."""
    t = [int(x) for x in text_tokens]
    return [cfg.start_text_token] + t + [cfg.stop_text_token]


def prompt_embeddings(sd, cfg, cond_latent, text_tokens):
    """[1, T+4, D] (T = unpadded token count): cond latent + text(T+3), text positions 0.. (autoregressive.py:540-543)."""
    ids = torch.tensor(text_ids(cfg, text_tokens), dtype=torch.long)
    te = sd["text_embedding.weight"][ids] + sd["text_pos_embedding.emb.weight"][: len(ids)]
    return torch.cat([cond_latent.reshape(1, 1, -1), te.unsqueeze(0)], dim=1)


def mel_pos_index(j, pos_mode):
    """j-th mel-segment token (start token j=0). SURVEY App. D-1: the reference's kv-cache path uses
    position j+1 for j>=1 (autoregressive.py:147-149); the recompute path uses j."""
    if pos_mode == "ref_kv_quirk":
        return j + 1 if j >= 1 else 0
    return j


def teacher_forced_logits(sd, cfg, cond_latent, text_tokens, codes, pos_mode="ref_kv_quirk"):
    """Logits the sampler would see at every step when fed `codes` [B, n] (LongTensor).
    Returns [B, n+1, V]: position j predicts mel token j (j=0 from the start token)."""
    B, n = codes.shape
    prompt = prompt_embeddings(sd, cfg, cond_latent, text_tokens).expand(B, -1, -1)
    ids = torch.cat([torch.full((B, 1), cfg.start_mel_token, dtype=torch.long), codes], dim=1)
    pos = torch.tensor([mel_pos_index(j, pos_mode) for j in range(n + 1)], dtype=torch.long)
    me = sd["mel_embedding.weight"][ids] + sd["mel_pos_embedding.emb.weight"][pos]
    hidden, _ = gpt2_trunk(sd, cfg, torch.cat([prompt, me], dim=1))
    return mel_logits(sd, hidden[:, prompt.shape[1]:])


def sample_step(logits, prev_ids, u, temperature=0.8, top_k=50, top_p=0.8, repetition_penalty=2.0):
    """One HF sample() step for one row in closed form.
    logits [V] f32; prev_ids: iterable of all ids seen so far incl. the fake prompt {1, 8192};
    u: uniform in [0,1) used for the inverse-CDF draw over the kept set in DESCENDING order.
    Returns (token, kept_ids (desc), kept_probs)."""
    s = logits.clone().float()
    idx = torch.tensor(sorted(set(int(i) for i in prev_ids)), dtype=torch.long)
    v = s[idx]
    s[idx] = torch.where(v < 0, v * repetition_penalty, v / repetition_penalty)
    s = s / temperature
    k = min(top_k, s.numel())
    vals, ids = torch.topk(s, k)  # descending
    p = torch.softmax(vals, dim=-1)
    excl = torch.cumsum(p, 0) - p
    keep = excl < top_p
    # HF: remove tokens with ascending-cumulative <= 1 - top_p  <=>  keep iff exclusive-descending mass < top_p
    keep[0] = True
    kp = p[keep]
    kp = kp / kp.sum()
    cdf = torch.cumsum(kp, 0)
    j = int(torch.searchsorted(cdf, torch.tensor(float(u)), right=True).clamp(max=kp.numel() - 1))
    return int(ids[keep][j]), ids[keep], kp


def generate(sd, cfg, cond_latent, text_tokens, uniforms, max_new, pos_mode="ref_kv_quirk",
             temperature=0.8, top_k=50, top_p=0.8, repetition_penalty=2.0):
    """Sequential KV-cached sampling with injected uniforms [B, max_new]. Returns codes [B, max_new]
    (finished rows emit stop tokens, as HF pads with pad_token_id = stop)."""
    B = uniforms.shape[0]
    prompt = prompt_embeddings(sd, cfg, cond_latent, text_tokens)
    start = sd["mel_embedding.weight"][cfg.start_mel_token] + sd["mel_pos_embedding.emb.weight"][0]
    emb = torch.cat([prompt, start.reshape(1, 1, -1)], dim=1).expand(B, -1, -1)
    hidden, past = gpt2_trunk(sd, cfg, emb)
    logits = mel_logits(sd, hidden[:, -1])
    codes = torch.full((B, max_new), cfg.stop_mel_token, dtype=torch.long)
    finished = [False] * B
    seen = [{1, cfg.start_mel_token} for _ in range(B)]
    for n in range(max_new):
        toks = []
        for b in range(B):
            if finished[b]:
                toks.append(cfg.stop_mel_token)
                continue
            t, _, _ = sample_step(logits[b], seen[b], float(uniforms[b, n]), temperature, top_k, top_p,
                                  repetition_penalty)
            toks.append(t)
            seen[b].add(t)
            if t == cfg.stop_mel_token:
                finished[b] = True
        codes[:, n] = torch.tensor(toks)
        if all(finished) or n == max_new - 1:
            break
        ids = torch.tensor(toks, dtype=torch.long)
        e = sd["mel_embedding.weight"][ids] + sd["mel_pos_embedding.emb.weight"][mel_pos_index(n + 1, pos_mode)]
        hidden, past = gpt2_trunk(sd, cfg, e.unsqueeze(1), past)
        logits = mel_logits(sd, hidden[:, -1])
    return codes


def fix_autoregressive_output(codes, stop_token=8193):
    """api.py:87-114 on one row (1-D LongTensor); returns a new tensor."""
    codes = codes.clone()
    pos = (codes == stop_token).nonzero()
    if len(pos) == 0:
        return codes
    stm = int(pos.min())
    codes[stm:] = 83
    if stm - 3 < codes.shape[0]:
        codes[-3] = 45
        codes[-2] = 45
        codes[-1] = 248
    return codes


def latents(sd, cfg, cond_latent, text_tokens, codes):
    """UnifiedVoice.forward(..., return_latent=True, clip_inputs=False) (autoregressive.py:454-512;
    called at api.py:521-524 with text_tokens already zero-padded once by api.py:391).
    codes [k, L] -> [k, L, D]."""
    k, L = codes.shape
    ids = torch.tensor(text_ids(cfg, text_tokens), dtype=torch.long)
    te = sd["text_embedding.weight"][ids] + sd["text_pos_embedding.emb.weight"][: len(ids)]
    mel_ids = torch.cat([torch.full((k, 1), cfg.start_mel_token, dtype=torch.long), codes,
                         torch.full((k, 1), cfg.stop_mel_token, dtype=torch.long)], dim=1)
    me = sd["mel_embedding.weight"][mel_ids] + sd["mel_pos_embedding.emb.weight"][: L + 2]
    emb = torch.cat([cond_latent.reshape(1, 1, -1).expand(k, -1, -1), te.unsqueeze(0).expand(k, -1, -1), me], dim=1)
    hidden, _ = gpt2_trunk(sd, cfg, emb)
    enc = _ln(hidden[:, 1:], sd["final_norm.weight"], sd["final_norm.bias"])
    return enc[:, -(L + 2):][:, :-2]


def stream_latents(sd, cfg, cond_latent, text_tokens, codes, pos_mode="ref_kv_quirk"):
    """Latents the streaming generator of the api_fast path yields next to its tokens
    (stream_generator.py:982: `final_norm(outputs.hidden_states[-1][:, -1])`, hidden_states[-1] = after ln_f, of the
    forward that produced the logits token i was sampled from). codes [n] (LongTensor, the yielded tokens, stop token
    included) -> [n, D]: row i belongs to the input [start, c_0 .. c_{i-1}] under the position rule of `pos_mode`."""
    codes = codes.reshape(1, -1)
    n = codes.shape[1]
    prompt = prompt_embeddings(sd, cfg, cond_latent, text_tokens)
    ids = torch.cat([torch.full((1, 1), cfg.start_mel_token, dtype=torch.long), codes[:, : n - 1]], dim=1)
    pos = torch.tensor([mel_pos_index(j, pos_mode) for j in range(n)], dtype=torch.long)
    me = sd["mel_embedding.weight"][ids] + sd["mel_pos_embedding.emb.weight"][pos]
    hidden, _ = gpt2_trunk(sd, cfg, torch.cat([prompt, me], dim=1))
    return _ln(hidden[0, prompt.shape[1]:], sd["final_norm.weight"], sd["final_norm.bias"])


def calm_trim_length(codes_row, calm_token=83):
    """api.py:547-556: index at which latents are cut (first run of >8 calm tokens), or len."""
    c = 0
    for i in range(codes_row.shape[-1]):
        if int(codes_row[i]) == calm_token:
            c += 1
        else:
            c = 0
        if c > 8:
            return i
    return codes_row.shape[-1]
