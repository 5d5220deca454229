"""Shared helpers for the GPU tests: error log (gpurun_out/errors.jsonl) so measured parity margins can be read back."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOG = os.path.join(ROOT, "gpurun_out", "errors.jsonl")


def report(name, value, **extra):
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps(dict(name=name, value=float(value), **extra)) + "\n")
    except OSError:
        pass
    print("[parity] %s: %.3e" % (name, float(value)))


def check_sampled_mel(name, mel, ref):
    """Bound on a mel produced by the DDPM sampling loop against the oracle's loop with the same injected noise.

    The chain amplifies an eps difference up to 153-fold before the clamp (utils/diffusion.py:420-425), so two runs that
    differ by one rounding (bf16 operands, an FMA instead of mul + add) separate into two samples of the same chain: the
    ROOT-MEAN-SQUARE and the 99.9th PERCENTILE of |difference| are stable (0.06-0.08 / < 1 on the 13.8 mel range) and are
    the hard bounds: rms < 0.3, p99.9 < 1.5 (= 1.5 x the drift of the fp32 oracle itself run with bf16-rounded GEMM
    operands, tests/test_host_orchestration.py). The MAXIMUM over ~10^5 elements is a single-element statistic of that
    chaotic process (measured 1.35 and 2.73 for two builds that differ only in how the last row block of a GEMM rounds its
    bias add); it is reported and bounded loosely (< 4.0, i.e. no element may cross a third of the range)."""
    d = (mel.float().cpu() - ref.float().cpu()).abs().flatten()
    mx, rms = d.max().item(), d.pow(2).mean().sqrt().item()
    q = d.kthvalue(max(1, int(0.999 * d.numel()))).values.item()
    report("%s max (range 13.8)" % name, mx)
    report("%s p99.9" % name, q)
    report("%s rms" % name, rms)
    assert rms < 0.3 and q < 1.5 and mx < 4.0, (name, mx, q, rms)
    return mx, q, rms
