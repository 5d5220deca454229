#!/usr/bin/env python
"""Launches the hot kernels a few times at the bench shapes, for `ncu -k regex:<name>` captures (profiles/ncu_r02/).
  python tools/prof_kernels.py fa | attn | attnc | step | step32 | lvc      (attnc = the compact attention kernel of the
  two-chain decode step at its batch of 128 candidates)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    from tortoise_tts_b200 import lib
    what = sys.argv[1]
    if what == "fa":
        from tortoise_tts_b200.diffusion_engine import _rel_pos_table
        S, C, H = 1872, 1024, 16
        qkv = torch.randn(2 * S, 3 * C, device="cuda").to(torch.bfloat16)
        o = torch.empty(2 * S, C, device="cuda", dtype=torch.bfloat16)
        bias = _rel_pos_table(torch.randn(32, H, device="cuda"), S, 8.0)
        for _ in range(4):
            lib.attention(qkv, o, nseq=2, T=S, H=H, ld=3 * C, ldo=C, k_off=C, v_off=2 * C, scale=0.125, bias=bias, bias_sat=64)
    elif what in ("attn", "attnc", "step", "step32"):
        from test_gpu_ar_step import _mk
        B = 32 if what == "step32" else (128 if what == "attnc" else 256)
        hd, t = _mk(B, 1024, 16, 30, 8194, 174, 430, 215, seed=1, attn_compact=what == "attnc")
        for _ in range(3):
            if what in ("attn", "attnc"):
                hd.step(phase_mask=4, layer_begin=15, layer_end=16)
            else:
                hd.step()
    elif what == "lvc":
        from tortoise_tts_b200.config import ModelConfig
        from tortoise_tts_b200.synth import synth_vocoder
        from tortoise_tts_b200.vocoder_engine import VocoderEngine
        cfg = ModelConfig.full()
        eng = VocoderEngine(synth_vocoder(cfg, 0), cfg)
        mel = torch.randn(100, 1872, device="cuda") * 2 - 5
        z = torch.randn(64, 1882, device="cuda")
        for _ in range(2):
            eng.inference(mel, z)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
