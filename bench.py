#!/usr/bin/env python
"""Benchmark of the Tortoise `preset='standard'` hot path (BASELINE.json metric: audio-seconds per wall-second).

  python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A step = one `TextToSpeech.tts_with_preset(preset='standard')` call (256 AR candidates, 200 diffusion iterations,
k=1) on the 53-word paragraph of SURVEY §8d config 3 (T=169 BPE tokens, N=430 mel tokens => S=1872 mel frames,
19.97 s of 24 kHz audio), synthetic seeded checkpoint in the reference layout (EOS suppressed so that every candidate
runs exactly N steps), synthetic conditioning latents. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio-seconds/sec at preset='standard'"
UNIT = "audio-s/s"
# ncu --set full, one launch each at the bench shapes (profiles/ncu_r01/*.ncu-rep): DRAM bytes read + written
NCU_TRAFFIC_BYTES = {
    "AR decode attention": 228.770560e6 + 8.168192e6,      # candidate KV stream kernel, ctx 174+215 (225.4 MB algorithmic)
    "diffusion attention": 23.328768e6,                    # qkv 23.0 MB read once; output stays in L2
    "diffusion conv k=3 GEMM": 29.341696e6 + 6.656e3,      # A 7.7 + W 6.3 + residual 15.3 MB; output stays in L2
}


def load_tokens(name="para53"):
    with open(os.path.join(ROOT, "tests", "golden", "bench_text_tokens.json")) as f:
        return json.load(f)[name]["tokens"]


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


_T0 = time.perf_counter()


def progress(msg):
    """stderr breadcrumbs (stdout carries only the JSON line): a run cut off by a timeout still says where it was."""
    if os.environ.get("RANK", "0") == "0":
        sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg))
        sys.stderr.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def kernel_probes(tts, cfg, n_mel, B, P):
    """Times the candidate dominant kernels live, in isolation at the workload's shapes, with CUDA events on the
    launching stream; returns per-kernel dicts with achieved throughput against the roofline that bounds each."""
    import torch
    from tortoise_tts_b200 import lib
    dev = tts.device
    hbm, tfl, how = peaks()
    out = []

    def timeit(fn, reps=10, flush=None):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if flush is not None:
                flush.zero_()          # > L2 (126 MB) write between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    S = n_mel * 4 * 24000 // 22050
    C, H = cfg.diff_dim, cfg.diff_heads
    # (1) diffusion attention: B=2 (cond+uncond), 16 heads, S x S, relative-position bias
    qkv = torch.randn(2 * S, 3 * C, device=dev).to(torch.bfloat16)
    o = torch.empty(2 * S, C, device=dev, dtype=torch.bfloat16)
    from tortoise_tts_b200.diffusion_engine import _rel_pos_table
    bias = _rel_pos_table(torch.randn(32, H, device=dev), S, 8.0)      # T5 table as the denoiser uses it
    ms = timeit(lambda: lib.attention(qkv, o, nseq=2, T=S, H=H, ld=3 * C, ldo=C, k_off=C, v_off=2 * C, scale=0.125, bias=bias,
                                      bias_sat=64), flush=flush)
    flops = 2 * H * 4.0 * S * S * 64
    out.append(dict(kernel="diffusion attention (2x16 heads, S=%d)" % S, bound="tensor", ms=ms, count=13 * 200,
                    achieved=flops / ms / 1e9, peak=tfl, unit="TFLOP/s"))
    # (2) diffusion conv k=3 as GEMM: [2, S, 1024] x [1024, 3*1024]
    a = torch.randn(2, S, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(C, 3 * C, device=dev) * 0.02).to(torch.bfloat16)
    x = torch.zeros(2, S, C, device=dev)
    bb = torch.zeros(C, device=dev)
    ms = timeit(lambda: lib.gemm(a, w, M=S, N=C, K=C, taps=3, pad=1, bias=bb, residual=x, out_f32=x, batch=2, a_bstride=S * C,
                                 res_bstride=S * C, outf_bstride=S * C), flush=flush)
    flops = 2 * 2.0 * S * C * 3 * C
    out.append(dict(kernel="diffusion conv k=3 GEMM (tcgen05, M=2x%d N=1024 K=3072)" % S, bound="tensor", ms=ms,
                    count=16 * 200, achieved=flops / ms / 1e9, peak=tfl, unit="TFLOP/s"))
    # (3) AR decode attention at the mean context (step n_mel/2): KV stream of all candidates, one layer
    Hh = cfg.ar_heads
    D = cfg.ar_dim
    Nmax = n_mel
    ck = torch.zeros(B, Hh, Nmax, 64, device=dev, dtype=torch.bfloat16)
    cv = torch.zeros_like(ck)
    pk = torch.zeros(Hh, P, 64, device=dev, dtype=torch.bfloat16)
    pv = torch.zeros_like(pk)
    qkv2 = torch.randn(B, 3 * D, device=dev).to(torch.bfloat16)
    o2 = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
    state = torch.zeros(64, dtype=torch.int32, device=dev)
    state[0] = n_mel // 2
    so = torch.zeros(2, B, D, device=dev)
    sl = torch.zeros(2, B, Hh, device=dev)
    ms = timeit(lambda: lib.ar_decode_attention(qkv2, pk, pv, ck, cv, state, B, Hh, P, Nmax, o2, so, sl), flush=flush)
    nbytes = B * Hh * (n_mel // 2) * 64 * 2 * 2 + Hh * P * 64 * 2 * 2   # candidate K+V (bf16) + shared prefix once
    out.append(dict(kernel="AR decode attention (B=%d, ctx=%d+%d)" % (B, P, n_mel // 2), bound="hbm", ms=ms,
                    count=30 * (n_mel - 1), achieved=nbytes / ms / 1e6, peak=hbm, unit="GB/s"))
    for d in out:
        d["frac"] = d["achieved"] / d["peak"]
        d["total_ms_per_utterance"] = d["ms"] * d["count"]
    return out, how


def run_engine(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from tortoise_tts_b200.config import ModelConfig
    from tortoise_tts_b200.synth import synth_all
    from tortoise_tts_b200.api import TextToSpeech
    from tortoise_tts_b200 import lib
    progress("extension built; synthesising the full-size checkpoint")
    cfg = ModelConfig.full()
    sds = synth_all(cfg, seed=0, suppress_stop=True)
    progress("checkpoint ready; loading engines")
    tts = TextToSpeech(state_dicts=sds, config=cfg, kv_cache=True, device="cuda:%d" % local)
    progress("engines ready")
    tokens = load_tokens(args.text)
    n_mel = args.mel_tokens
    g = torch.Generator().manual_seed(0)
    cl_host = ((torch.randn(1, cfg.ar_dim, generator=g) * 0.5).pin_memory(),
               (torch.randn(1, 2 * cfg.diff_dim, generator=g) * 0.3).pin_memory())
    kw = dict(text_tokens=tokens, conditioning_latents=cl_host, max_mel_tokens=n_mel, verbose=False, k=1)
    if args.preset_override:
        kw.update(json.loads(args.preset_override))

    def step(i):
        return tts.tts_with_preset("", preset=args.preset, use_deterministic_seed=1000 + i, **kw)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        wav = step(i)
        progress("warm-up %d done: %s" % (i, {k_: round(v, 1) for k_, v in tts.last_timings.items()}))
    sampler = ClockSampler(local)
    sampler.start()
    sync()
    c0 = lib.CALLS
    dev_ms = []
    stage = {}
    t0 = time.perf_counter()
    for i in range(args.steps):
        wav = step(args.warmup + i)
        dev_ms.append(tts.last_timings["device_total_ms"])
        progress("timed step %d: %.1f ms on device" % (i, dev_ms[-1]))
        for k_, v in tts.last_timings.items():
            stage[k_] = stage.get(k_, 0.0) + v / args.steps
    sync()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    launches = lib.CALLS - c0
    audio_s = wav.shape[-1] / 24000.0
    ms_per_step = wall / args.steps * 1e3
    dev_step = sum(dev_ms) / len(dev_ms)
    if world > 1:
        t = torch.tensor([ms_per_step, dev_step], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, dev_step = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    B = 256 if args.preset in ("standard", "high_quality") else (96 if args.preset == "fast" else 16)
    h2d = sum(t.numel() * 4 for t in cl_host) + 4 * (len(tokens) + 1)
    line = {
        "metric": METRIC, "value": audio_s / (dev_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 tensor-core operands, fp32 accumulate/residual/norm/softmax/scheduler",
        "data": "synthetic (seeded random checkpoint in the reference .pth layout, EOS suppressed; synthetic latents)",
        "config": {"workload": "configs[2]: preset='%s' (%d AR samples, %d diffusion iters), %d-token paragraph, "
                               "N=%d mel tokens -> %.2f s audio, k=1" % (args.preset, B, {"standard": 200, "fast": 80,
                                                                         "ultra_fast": 30, "high_quality": 400}[args.preset],
                                                                         len(tokens), n_mel, audio_s),
                   "l2": "working set (1.9 GB weights + %.1f GB KV cache) >> 126 MB L2; no flush between steps" %
                         (2 * 30 * B * 16 * n_mel * 64 * 2 / 1e9),
                   "parallelism": "candidates sharded %d/GPU" % ((B + world - 1) // world)},
        "e2e": {"value": audio_s / (ms_per_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(wav.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "stage_ms": {k_: round(v, 2) for k_, v in stage.items()},
    }
    if world == 1:
        progress("kernel probes")
        probes, how = kernel_probes(tts, cfg, n_mel, B, len(tokens) + 5)
        progress("probes done")
        dom = max(probes, key=lambda d: d["total_ms_per_utterance"])
        # dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed ncu --set full
        # captures (profiles/ncu_summary_r01_run10.txt, same shapes as the probes); null for a kernel without a capture
        traffic = None
        for key, val in NCU_TRAFFIC_BYTES.items():
            if dom["kernel"].startswith(key):
                traffic = val
        line["roofline"] = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": round(dom["achieved"], 2),
                            "peak": dom["peak"], "unit": dom["unit"], "frac": round(dom["frac"], 4), "traffic": traffic,
                            "traffic_source": "profiles/ncu_summary_r01_run10.txt" if traffic else None,
                            "peak_source": how, "launch_ms": round(dom["ms"], 4)}
        line["kernels"] = [{k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in d.items()} for d in probes]
        if not args.no_cpu_baseline:
            progress("cpu baseline (oracle port on the host cores)")
            line["cpu_baseline"] = cpu_baseline(cfg, sds, tokens, B, n_mel, args)
            progress("cpu baseline done")
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cfg, sds, tokens, B, n_mel, args):
    # child process under a timeout: the oracle's thread pool must not inherit this process's CUDA-side state, and a
    # host whose CPU quota misbehaves must not take the GPU arm's JSON line down with it
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--preset", args.preset, "--mel-tokens", str(n_mel),
           "--tokens-json", os.path.join(ROOT, "tests", "golden", "bench_text_tokens.json"), "--text", args.text]
    try:
        out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=args.cpu_timeout)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except subprocess.TimeoutExpired:
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                "sample": "not finished within %d s on this host" % args.cpu_timeout}
    except (ValueError, IndexError):
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: " + out.stderr[-300:]}
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
            "total_s_extrapolated": round(r["total_s"], 1), "units_s": {k: round(v, 4) for k, v in r["units"].items()}}


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path is Python/PyTorch and does not travel to
    the GPU box, so its restatement (oracle/, pinned against the reference modules) is timed on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tokens = load_tokens(args.text)
    n = args.warmup + args.steps
    # one child process takes all W+K samples (checkpoint synthesis and the thread-count probe are paid once); each
    # sample is the bounded unit-cost measurement of oracle/cpu_baseline.py (~20-40 s of CPU work)
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--preset", args.preset, "--mel-tokens", str(args.mel_tokens),
           "--tokens-json", os.path.join(ROOT, "tests", "golden", "bench_text_tokens.json"), "--text", args.text,
           "--repeat", str(n)]
    t0 = time.perf_counter()
    samples = []
    try:
        out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=args.cpu_timeout * n)
        txt, err = out.stdout, out.stderr
    except subprocess.TimeoutExpired as e:
        txt = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        err = "timeout after %d s" % (args.cpu_timeout * n)
    for ln in txt.strip().splitlines():
        try:
            samples.append(json.loads(ln))
        except ValueError:
            pass
    wall = time.perf_counter() - t0
    timed = samples[args.warmup:] if len(samples) > args.warmup else samples[-1:]
    if not timed:
        print(json.dumps({"impl": "reference", "unavailable": "CPU oracle produced no sample: " + err[-200:].replace("\n", " ")}))
        return
    last = timed[-1]
    v = sum(s["value"] for s in timed) / len(timed)
    audio_s = (args.mel_tokens * 4 * 24000 // 22050) * 256 / 24000.0
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": audio_s / v * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: preset='%s', %d-token paragraph, N=%d mel tokens (CPU oracle port of the "
                                   "reference path, unit costs extrapolated)" % (args.preset, len(tokens), args.mel_tokens)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "bench_wall_s": round(wall, 1)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--preset", default="standard")
    ap.add_argument("--text", default="para53")
    ap.add_argument("--mel-tokens", type=int, default=430)
    ap.add_argument("--preset-override", default=None, help="JSON dict of tts kwargs (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-timeout", type=int, default=240, help="seconds allowed per CPU-oracle sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
