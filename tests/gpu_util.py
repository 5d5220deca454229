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
