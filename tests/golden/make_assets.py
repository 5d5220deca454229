"""Copies the reference's DATA assets the front-end needs (numbers, not code) into the package:
tortoise/data/mel_norms.pth (80 per-bin divisors of TorchMelSpectrogram, models/arch_util.py:295-331) -> JSON."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
m = torch.load("/root/reference/tortoise/data/mel_norms.pth")
with open(os.path.join(ROOT, "tortoise_tts_b200", "data", "mel_norms.json"), "w") as f:
    json.dump({"source": "tortoise/data/mel_norms.pth", "mel_norms": [float(v) for v in m.double()]}, f)
print("ok", m.shape)
