#!/usr/bin/env python
"""Per-op device time of one AR decode step and one diffusion step at the `standard` shapes, measured live with CUDA
events around every C-ABI call (eager mode, no graph). Development aid: shows where a step's time goes without ncu.

  python tools/op_profile.py [--mel-tokens 430] [--candidates 256]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def instrument(lib, records):
    names = [n for n in dir(lib) if callable(getattr(lib, n)) and n in (
        "gemm", "layernorm", "residual_layernorm", "rmsnorm", "groupnorm", "groupnorm_apply", "attention", "ar_embed_step",
        "ar_decode_attention", "ar_sample", "diffusion_step", "cast_pad_bf16", "counter_add", "clvp_rotary")]
    saved = {}
    for n in names:
        fn = getattr(lib, n)
        saved[n] = fn

        def wrap(*a, _fn=fn, _n=n, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = _fn(*a, **k)
            e1.record()
            tag = _n
            if _n == "gemm":
                tag = "gemm M=%d N=%d K=%d taps=%d b=%d" % (k["M"], k["N"], k["K"], k.get("taps", 1), k.get("batch", 1))
            records.append((tag, e0, e1))
            return r
        setattr(lib, n, wrap)
    return saved


def summarize(records, title):
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for tag, e0, e1 in records:
        a = agg.setdefault(tag, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    tot = sum(a[1] for a in agg.values())
    print("== %s: %.3f ms in %d calls" % (title, tot, sum(a[0] for a in agg.values())))
    for tag, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %8.3f ms %5.1f%%  x%-4d avg %7.1f us  %s" % (a[1], 100 * a[1] / tot, a[0], 1e3 * a[1] / a[0], tag))
    return {t: a for t, a in agg.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mel-tokens", type=int, default=430)
    ap.add_argument("--candidates", type=int, default=256)
    args = ap.parse_args()
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_all
    from tortoise_tts_b200 import lib
    from tortoise_tts_b200.ar_engine import AREngine
    from tortoise_tts_b200.diffusion_engine import DiffusionEngine
    cfg = ModelConfig.full()
    sds = synth_all(cfg, seed=0, suppress_stop=True)
    toks = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_text_tokens.json")))["para53"]["tokens"] + [0]
    B, N = args.candidates, args.mel_tokens
    ar = AREngine(sds["autoregressive"], cfg)
    cond = torch.randn(1, cfg.ar_dim) * 0.5
    # fake a mid-run context for the timing: a bigger cache with step = N/2
    ar._dec = None
    st = ar._decode_state(B, len(toks) + 4, N)
    ar._prefill(cond, toks, st)

    def mid_run():
        for ch in st["chains"]:
            ch["codes"].fill_(100)
            ch["state"][0] = N // 2
    sp = dict(temperature=0.8, top_k=50, top_p=0.8, rep_penalty=2.0, pos_mode=1)
    for _ in range(2):
        mid_run()
        ar._decode_step(st, sp)
    rec = []
    saved = instrument(lib, rec)
    mid_run()
    ar._decode_step(st, sp)
    summarize(rec, "AR decode step at ctx %d+%d, B=%d" % (st["P"], N // 2, B))
    for n, fn in saved.items():
        setattr(lib, n, fn)
    del ar
    torch.cuda.empty_cache()
    # ---- diffusion
    de = DiffusionEngine(sds["diffusion"], cfg)
    S = N * 4 * 24000 // 22050
    lat = torch.randn(N, cfg.ar_dim)
    dc = torch.randn(2 * cfg.diff_dim) * 0.3
    de.sample(lat, dc, 4, torch.randn(100, S), torch.randn(4, 100, S), use_graph=False)
    stt = de._ws
    stt["counter"].zero_()
    rec = []
    saved = instrument(lib, rec)
    de._step(stt)
    summarize(rec, "diffusion step, S=%d, cond+uncond" % S)
    for n, fn in saved.items():
        setattr(lib, n, fn)


if __name__ == "__main__":
    main()
