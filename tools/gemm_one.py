#!/usr/bin/env python
"""Launches ONE GEMM shape a few times (for `ncu --set full -k regex:gemm_bf16 -s 2 -c 1 python tools/gemm_one.py`).
Shapes: k3 (diffusion conv k=3, in-place fp32 residual), qkv (diffusion qkv, bf16 out), clvp, ar."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from tortoise_tts_b200 import lib
    which = sys.argv[1] if len(sys.argv) > 1 else "k3"
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    M, N, K, taps, batch, res, out, tile = {
        "k3": (1872, 1024, 1024, 3, 2, True, "f32", 128),
        "qkv": (1872, 3072, 1024, 1, 2, False, "bf16", 128),
        "clvp": (27520, 2304, 768, 1, 1, False, "bf16", 128),
        "ar": (256, 3072, 1024, 1, 1, False, "bf16", 32),
    }[which]
    if len(sys.argv) > 3:
        tile = int(sys.argv[3])
    dev = "cuda"
    A = torch.randn(batch, M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, taps * K, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    of = torch.zeros(batch, M, N, device=dev) if out == "f32" else None
    ob = torch.zeros(batch, M, N, device=dev, dtype=torch.bfloat16) if out == "bf16" else None
    for _ in range(4):
        lib.gemm(A, W, M=M, N=N, K=K, taps=taps, pad=(taps - 1) // 2, batch=batch, a_bstride=M * K, res_bstride=M * N,
                 outf_bstride=M * N, outb_bstride=M * N, bias=bias, residual=of if res else None, out_f32=of, out_bf16=ob,
                 tile_n=tile, variant=variant)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
