"""torchrun --nproc-per-node 2 tools/dbg/pair_check.py : the CFG pair (peer-memory exchange and NCCL fallback) against the
single-GPU loop on the same inputs; prints max |d mel| and step times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from tortoise_tts_b200.config import ModelConfig
from tortoise_tts_b200.synth import synth_diffusion
from tortoise_tts_b200.diffusion_engine import DiffusionEngine
from tortoise_tts_b200 import parallel
cfg = ModelConfig.full()
sd = synth_diffusion(cfg, 0)
torch.manual_seed(0)
N, iters = 430, 200
S = N * 4 * 24000 // 22050
lat = torch.randn(N, cfg.ar_dim); cond = torch.randn(2 * cfg.diff_dim) * 0.3
noise0 = torch.randn(100, S); step_noise = torch.randn(iters, 100, S)
rank = dist.get_rank()
groups, _ = parallel.pair_groups()
res = {}
for mode in ("single", "peer", "nccl"):
    os.environ["TTB_PEER_EXCHANGE"] = "0" if mode == "nccl" else "1"
    eng = DiffusionEngine(sd, cfg)
    pair = None if mode == "single" else (groups[0], rank)
    print("rank", rank, "mode", mode, "starting", flush=True)
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        mel = eng.sample(lat, cond, iters, noise0, step_noise, cond_free=True, cond_free_k=2.0, pair=pair)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[mode] = mel.clone()
    if rank == 0:
        xch = eng._ws.get("xch")
        print("%-6s: %.1f ms per 200-step loop (%.3f ms/step)  peer_exchange=%s" % (mode, dt * 1e3, dt * 1e3 / iters, xch is not None))
    del eng
    torch.cuda.empty_cache()
if rank == 0:
    for m in ("peer", "nccl"):
        print("max |mel(%s) - mel(single)| = %.3e" % (m, (res[m] - res["single"]).abs().max().item()))
dist.destroy_process_group()
