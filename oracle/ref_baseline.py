"""TEST/BENCH INFRASTRUCTURE ONLY — the reference arm of bench.py: the UNMODIFIED reference modules
(neonbjb/tortoise-tts, imported from /root/reference or its copy oracle/_ref, see oracle/build_ref.py) timed on the
host cores (`--device cpu`, BASELINE.md "reference CPU path") or in PyTorch eager on the GPU (`--device cuda`,
BASELINE.md §3 "bar to beat"), through the reference's own call path for each stage:

  UnifiedVoice.inference_speech -> HF generate (autoregressive.py:535-563), CLVP.forward (clvp.py:99-140),
  UnifiedVoice.forward(return_latent=True) (autoregressive.py:454-512), DiffusionTts.forward cond / uncond
  (diffusion_decoder.py:262-322), UnivNetGenerator.inference (vocoder.py:300-312).

A full `standard` utterance takes the reference ~40 min on 16 cores, so each sample measures UNIT costs on the real
shapes and extrapolates with the reference's loop counts (api.py:407-427: batches of 16 candidates, prompt recomputed per
batch; utils/diffusion.py:340-342: two forwards per step); the decode-step cost is taken at the START of the sequence,
where the reference's growing KV concatenation is cheapest (a lower bound on its cost). The sample says so.
"""
import json
import os
import sys
import time

import torch


def _sync(dev):
    if dev.type == "cuda":
        torch.cuda.synchronize()


def _t(fn, dev, reps=1):
    best = 1e30
    for _ in range(reps):
        _sync(dev)
        t0 = time.perf_counter()
        fn()
        _sync(dev)
        best = min(best, time.perf_counter() - t0)
    return best


_MODELS = {}


def _models(cfg, sds, dev):
    """The reference modules, built ONCE per process and device: a run of `--repeat n` samples re-times the unit costs n
    times on the same modules instead of constructing 1.2 G parameters per sample (the reference arm of bench.py has to
    fit n = warmup + steps samples into a few minutes)."""
    key = dev.type
    if key not in _MODELS:
        from .ref_build import build_reference_models
        mods = build_reference_models(cfg, sds, kv_cache=True)
        _MODELS[key] = tuple(mods[k].to(dev) for k in ("autoregressive", "clvp", "diffusion", "vocoder"))
    return _MODELS[key]


def measure(cfg, sds, text_tokens, num_candidates=256, n_mel=430, iters=200, cond_free=True, device="cpu", threads=None,
            gen_points=None, clvp_rows=None):
    dev = torch.device(device)
    if threads and dev.type == "cpu":
        torch.set_num_threads(threads)
    if gen_points is None:
        gen_points = (8, 40) if dev.type == "cuda" else (2, 5)
    if clvp_rows is None:
        clvp_rows = 16 if dev.type == "cuda" else 2
    ar, clvp, diff, voc = _models(cfg, sds, dev)
    torch.manual_seed(0)
    bs = 16
    cond = (torch.randn(1, cfg.ar_dim) * 0.5).to(dev)
    text = torch.tensor(text_tokens, dtype=torch.long, device=dev).unsqueeze(0)      # api-padded ids (api.py:391)
    units = {}
    reps = 2 if dev.type == "cuda" else 1
    with torch.no_grad():
        def gen(n):
            return ar.inference_speech(cond, text, do_sample=True, top_p=0.8, top_k=50, temperature=0.8,
                                       num_return_sequences=bs, length_penalty=1.0, repetition_penalty=2.0,
                                       max_generate_length=n)
        if dev.type == "cuda":
            gen(2)                                       # warm-up (cuBLAS handles, HF caches)
        n1, n2 = gen_points
        t1 = _t(lambda: gen(n1), dev, reps)
        t2 = _t(lambda: gen(n2), dev, reps)
        step = max((t2 - t1) / (n2 - n1), 0.0)
        units["ar_decode_step_b16_s"] = step
        units["ar_prefill_b16_s"] = max(t1 - n1 * step, 0.0)
        codes = torch.randint(0, 8192, (clvp_rows, n_mel), device=dev)
        tt = text.repeat(clvp_rows, 1)
        t_clvp = _t(lambda: clvp(tt, codes, return_loss=False), dev, reps)
        units["clvp_b16_s"] = t_clvp * (bs / clvp_rows)
        lc = torch.randint(0, 8192, (1, n_mel), device=dev)
        tl = torch.tensor([text.shape[-1]], device=dev)
        wl = torch.tensor([n_mel * 1024], device=dev)
        units["latents_s"] = _t(lambda: ar(cond, text, tl, lc, wl, return_latent=True, clip_inputs=False), dev, reps)
        S = n_mel * 4 * 24000 // 22050
        x = torch.randn(1, 100, S, device=dev)
        ce = torch.randn(1, cfg.diff_dim, S, device=dev)
        t = torch.tensor([2000], device=dev)
        units["diff_forward_cond_s"] = _t(lambda: diff(x, t, precomputed_aligned_embeddings=ce), dev, reps)
        units["diff_forward_uncond_s"] = _t(lambda: diff(x, t, precomputed_aligned_embeddings=ce, conditioning_free=True),
                                            dev, reps) if cond_free else 0.0
        mel = torch.randn(1, 100, S, device=dev) * 2 - 5
        units["vocoder_s"] = _t(lambda: voc.inference(mel), dev, reps)
    nb = (num_candidates + bs - 1) // bs
    total = (nb * (units["ar_prefill_b16_s"] + n_mel * units["ar_decode_step_b16_s"]) + nb * units["clvp_b16_s"]
             + units["latents_s"] + iters * (units["diff_forward_cond_s"] + units["diff_forward_uncond_s"])
             + units["vocoder_s"])
    audio_s = S * 256 / 24000.0
    sample = ("unmodified reference modules on %s; extrapolated from unit costs on real shapes: HF generate at B=16 for %d and "
              "%d tokens (prefill + per-token cost at the start of the sequence), CLVP on %d candidates, 1 latent pass, 1 cond "
              "+ 1 uncond denoiser forward at S=%d, 1 vocoder pass" % (dev.type, n1, n2, clvp_rows, S))
    cores = torch.get_num_threads() if dev.type == "cpu" else 0
    return dict(total_s=total, audio_s=audio_s, value=audio_s / total, units=units, sample=sample, cores=cores,
                kind="reference", device=dev.type)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="standard")
    ap.add_argument("--mel-tokens", type=int, default=430)
    ap.add_argument("--tokens-json", required=True)
    ap.add_argument("--text", default="para53")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_all
    from .cpu_baseline import pick_threads, host_threads
    cfg = ModelConfig.full()
    sds = synth_all(cfg, seed=0, suppress_stop=True)
    with open(a.tokens_json) as f:
        tokens = json.load(f)[a.text]["tokens"]
    B = 256 if a.preset in ("standard", "high_quality") else (96 if a.preset == "fast" else 16)
    iters = {"standard": 200, "fast": 80, "ultra_fast": 30, "high_quality": 400}[a.preset]
    threads = a.threads or (pick_threads() if a.device == "cpu" else 0)
    for _ in range(a.repeat):
        r = measure(cfg, sds, tokens + [0], num_candidates=B, n_mel=a.mel_tokens, iters=iters,
                    cond_free=a.preset != "ultra_fast", device=a.device, threads=threads)
        r["cores_available"] = host_threads(cap=4096)
        print(json.dumps(r))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
