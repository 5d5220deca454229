#!/usr/bin/env python
"""Times ttb_groupnorm at the denoiser's shape (B=2, S=1872, C=1024, 32 groups, fused scale-shift + SiLU, bf16 out):
row-wise kernels against the generic group-wise ones (TTB_GN_IMPL=group), back-to-back launches (L2-warm, as inside
the per-step CUDA graph) and with an L2 flush between launches. Development aid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from tortoise_tts_b200 import lib
    dev = "cuda"
    B, S, C, G = 2, 1872, 1024, 32
    x = torch.randn(B, S, C, device=dev)
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    ss = torch.randn(4, 2 * C, device=dev) * 0.3
    row = torch.tensor([1], dtype=torch.int32, device=dev)
    part = lib.groupnorm_scratch(B, G, dev)
    ob = torch.empty(B, S, C, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    nbytes = B * S * C * (4 + 4 + 2)          # stats read + apply read + bf16 write

    def call():
        lib.groupnorm(x, B, S, C, G, gamma, beta, part, scale_shift=ss, ss_row=row, ss_row_stride=2 * C, silu=True,
                      out_bf16=ob, ldo=C)

    for impl in ("rows", "group"):
        os.environ["TTB_GN_IMPL"] = impl
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            call()
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 50 * 1e3
        ts = []
        for _ in range(10):
            flush.zero_()
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print("groupnorm %-5s  back-to-back %.1f us (%.0f GB/s algorithmic)   after L2 flush %.1f us (%.0f GB/s)" %
              (impl, warm, nbytes / warm / 1e3, ts[len(ts) // 2], nbytes / ts[len(ts) // 2] / 1e3))


def ln_main():
    """residual + split-K-reduce + LayerNorm of the GPT-2 decode step (256 rows x 1024): warp-per-row vs block-per-row."""
    from tortoise_tts_b200 import lib
    dev = "cuda"
    M, D, ns = 256, 1024, 4
    x = torch.randn(M, D, device=dev)
    part = torch.randn(ns, M, D, device=dev) * 0.1
    bias, g1, b1 = torch.randn(D, device=dev), torch.randn(D, device=dev), torch.randn(D, device=dev)
    ob = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    outs = {}
    for impl in ("warp", "block"):
        os.environ["TTB_LN_IMPL"] = impl
        xx = x.clone()
        lib.residual_layernorm(xx, M, D, part, ns, M * D, bias, g1, b1, out_bf16=ob)
        torch.cuda.synchronize()
        outs[impl] = (xx.clone(), ob.float().clone())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            lib.residual_layernorm(xx, M, D, part, ns, M * D, bias, g1, b1, out_bf16=ob)
        e1.record()
        torch.cuda.synchronize()
        print("residual_layernorm %-5s back-to-back %.2f us" % (impl, e0.elapsed_time(e1) / 100 * 1e3))
    print("warp vs block: max |dx| %.2e, max |dy| %.2e" % ((outs["warp"][0] - outs["block"][0]).abs().max().item(),
                                                           (outs["warp"][1] - outs["block"][1]).abs().max().item()))


if __name__ == "__main__":
    main()
    ln_main()
