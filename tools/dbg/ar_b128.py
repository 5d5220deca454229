import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tortoise_tts_b200.config import ModelConfig
from tortoise_tts_b200.synth import synth_autoregressive
from tortoise_tts_b200 import ar_engine
import json
cfg = ModelConfig.full()
sd = synth_autoregressive(cfg, 0, True)
toks = json.load(open("tests/golden/bench_text_tokens.json"))["para53"]["tokens"] + [0]
cond = torch.randn(1, cfg.ar_dim) * 0.5
for B in (128, 64, 32):
    eng = ar_engine.AREngine(sd, cfg)
    u = torch.rand(B, 430)
    codes = eng.generate(cond, toks, B, 430, uniforms=u)
    torch.cuda.synchronize()
    print("B", B, "mode", eng._dec["mode"], "ok", codes.shape, int(codes.max()))
    del eng
    torch.cuda.empty_cache()
