import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tortoise_tts_b200 import lib
from tortoise_tts_b200.diffusion_engine import _rel_pos_table
S, C, H = 1872, 1024, 16
qkv = torch.randn(2 * S, 3 * C, device="cuda").to(torch.bfloat16)
o = torch.empty(2 * S, C, device="cuda", dtype=torch.bfloat16)
bias = _rel_pos_table(torch.randn(32, H, device="cuda"), S, 8.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]
f = lambda: lib.attention(qkv, o, nseq=2, T=S, H=H, ld=3*C, ldo=C, k_off=C, v_off=2*C, scale=0.125, bias=bias, bias_sat=64)
print("FA %s: %.4f ms" % ({k:v for k,v in os.environ.items() if k.startswith("TTB_FA")}, timeit(f)))
