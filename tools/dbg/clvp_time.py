#!/usr/bin/env python
"""Times CLVPEngine.text_latent / scores at several text lengths (config 5 showed a slow CLVP stage at T = 337)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_clvp
    from tortoise_tts_b200.clvp_engine import CLVPEngine
    cfg = ModelConfig.full()
    eng = CLVPEngine(synth_clvp(cfg, 0), cfg)
    codes = torch.randint(0, 8192, (256, 500))

    def t(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for T in (169, 255, 256, 257, 265, 337, 380):
        toks = [(i * 7) % 250 + 1 for i in range(T)] + [0]
        a = t(lambda: eng.text_latent(toks))
        b = t(lambda: eng.scores(toks, codes))
        print("T=%d: text_latent %.2f ms, scores(256 x 500) %.2f ms" % (T, a, b), flush=True)


if __name__ == "__main__":
    main()
