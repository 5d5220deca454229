"""TEST/BENCH INFRASTRUCTURE ONLY — CPU baseline of the hot path on the host cores (bench.py `cpu_baseline` leg and
`--impl reference` arm). Times the CPU oracle (a restatement of the reference's PyTorch path, see oracle/*.py; the
reference itself is Python and cannot travel to the GPU box) on a BOUNDED sample of the `standard` workload and
extrapolates with the unit counts of BASELINE.md §3:

  total = ceil(B/16) * (prefill(B=16, P) + (N-1) * decode_step(B=16)) + ceil(B/16) * clvp(B=16, N)
          + latents(1 x (P+N+2)) + iters * (cond + uncond denoiser forward at S) + vocoder(S)

(the reference decodes in batches of 16 with a KV cache, recomputes the prompt per batch, api.py:407-427, and runs the
two CFG branches sequentially, utils/diffusion.py:340-342).
"""
import time

import torch

from . import ar, clvp, diffusion as od, vocoder as ov


def host_threads(cap=64):
    """CPU threads this process may really use: the affinity mask and the cgroup CPU quota, not os.cpu_count() (on a
    shared box that is the machine's core count; handing it to OpenMP oversubscribes the quota by 10x+ and the spin-
    waiting workers make every matmul crawl -- the first GPU-box run of this baseline never finished for that reason)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:                                   # cgroup v1
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                quota = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                period = int(f.read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return max(1, min(n, cap))


def pick_threads():
    """The thread count that is actually fastest here: a quota the files above do not show still oversubscribes, so
    time a 1536^3 fp32 matmul at the detected count and at a few smaller ones and keep the best."""
    a = torch.randn(1536, 1536)
    best_n, best_t = 1, 1e30
    for n in sorted({host_threads(), 32, 16, 8, 4}):
        if n > host_threads():
            continue
        torch.set_num_threads(n)
        torch.mm(a, a)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.mm(a, a)
        t = time.perf_counter() - t0
        if t < best_t * 0.9:                   # prefer fewer threads unless clearly slower
            best_n, best_t = n, t
    return best_n


def _t(fn, reps=1):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def measure(cfg, sds, text_tokens, num_candidates=256, n_mel=430, iters=200, cond_free=True, decode_steps=3,
            clvp_rows=2, threads=None):
    """Returns dict(total_s, audio_s, value (audio-s/s), units, sample). text_tokens: api-padded ids."""
    if threads:
        torch.set_num_threads(threads)
    sd_ar, sd_clvp, sd_diff, sd_voc = sds["autoregressive"], sds["clvp"], sds["diffusion"], sds["vocoder"]
    torch.manual_seed(0)
    bs = 16
    cond = torch.randn(1, cfg.ar_dim) * 0.5
    units = {}
    with torch.no_grad():
        # ---- AR: prefill at B=16 then `decode_steps` cached steps (context P + ~n_mel/2 would be the average; the
        # measured steps run at context P+1.., a lower bound on the reference's per-step cost)
        prompt = ar.prompt_embeddings(sd_ar, cfg, cond, text_tokens)
        start = sd_ar["mel_embedding.weight"][cfg.start_mel_token] + sd_ar["mel_pos_embedding.emb.weight"][0]
        emb = torch.cat([prompt, start.reshape(1, 1, -1)], dim=1).expand(bs, -1, -1)
        state = {}

        def prefill():
            h, past = ar.gpt2_trunk(sd_ar, cfg, emb)
            state["past"] = past
            state["logits"] = ar.mel_logits(sd_ar, h[:, -1])
        units["ar_prefill_b16_s"] = _t(prefill)

        def decode():
            tok = torch.randint(0, 8192, (bs,))
            e = sd_ar["mel_embedding.weight"][tok] + sd_ar["mel_pos_embedding.emb.weight"][2]
            h, past = ar.gpt2_trunk(sd_ar, cfg, e.unsqueeze(1), state["past"])
            state["past"] = past
            state["logits"] = ar.mel_logits(sd_ar, h[:, -1])
        t0 = time.perf_counter()
        for _ in range(decode_steps):
            decode()
        units["ar_decode_step_b16_s"] = (time.perf_counter() - t0) / decode_steps
        state.clear()
        # ---- CLVP on `clvp_rows` candidates (linear in rows) + the text encoder once per batch as the reference does
        codes = torch.randint(0, 8192, (clvp_rows, n_mel))
        t_clvp = _t(lambda: clvp.scores(sd_clvp, cfg, torch.tensor(text_tokens), codes))
        units["clvp_b16_s"] = t_clvp * (bs / clvp_rows)
        # ---- latents for k=1
        lc = torch.randint(0, 8192, (1, n_mel))
        units["latents_s"] = _t(lambda: ar.latents(sd_ar, cfg, cond, text_tokens, lc))
        # ---- diffusion: one conditional + one unconditional forward at S
        S = od.output_seq_len(n_mel)
        x = torch.randn(1, 100, S)
        ce = torch.randn(1, cfg.diff_dim, S)
        t = torch.tensor([2000])
        units["diff_forward_cond_s"] = _t(lambda: od.forward(sd_diff, cfg, x, t, code_emb=ce))
        units["diff_forward_uncond_s"] = _t(lambda: od.forward(sd_diff, cfg, x, t, conditioning_free=True)) if cond_free else 0.0
        # ---- vocoder
        mel = torch.randn(1, 100, S) * 2 - 5
        z = torch.randn(1, 64, S + 10)
        units["vocoder_s"] = _t(lambda: ov.inference(sd_voc, mel, z))
    nb = (num_candidates + bs - 1) // bs
    total = (nb * (units["ar_prefill_b16_s"] + (n_mel - 1) * units["ar_decode_step_b16_s"]) + nb * units["clvp_b16_s"]
             + units["latents_s"] + iters * (units["diff_forward_cond_s"] + units["diff_forward_uncond_s"])
             + units["vocoder_s"])
    audio_s = S * 256 / 24000.0
    sample = ("extrapolated from unit costs on real shapes: 1 AR prefill + %d cached decode steps at B=16, CLVP on %d "
              "candidates, 1 latent pass, 1 cond + 1 uncond denoiser forward at S=%d, 1 vocoder pass" %
              (decode_steps, clvp_rows, S))
    return dict(total_s=total, audio_s=audio_s, value=audio_s / total, units=units, sample=sample,
                cores=torch.get_num_threads())


def main():
    """`python -m oracle.cpu_baseline --preset standard --mel-tokens 430 --tokens-json <file> --text para53`
    prints one JSON object; bench.py runs this in a child process under a timeout so that the GPU arm's JSON line
    never depends on how the host's CPU quota behaves."""
    import argparse
    import json
    import os
    import sys
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="standard")
    ap.add_argument("--mel-tokens", type=int, default=430)
    ap.add_argument("--tokens-json", required=True)
    ap.add_argument("--text", default="para53")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1, help="samples to take (one JSON line each)")
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_all
    cfg = ModelConfig.full()
    sds = synth_all(cfg, seed=0, suppress_stop=True)
    with open(a.tokens_json) as f:
        tokens = json.load(f)[a.text]["tokens"]
    B = 256 if a.preset in ("standard", "high_quality") else (96 if a.preset == "fast" else 16)
    iters = {"standard": 200, "fast": 80, "ultra_fast": 30, "high_quality": 400}[a.preset]
    threads = a.threads or pick_threads()
    for _ in range(a.repeat):
        r = measure(cfg, sds, tokens + [0], num_candidates=B, n_mel=a.mel_tokens, iters=iters,
                    cond_free=a.preset != "ultra_fast", threads=threads)
        print(json.dumps(r))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
