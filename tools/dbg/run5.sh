#!/bin/bash
# round-2 closing measurements (one box): ncu capture of the compact attention kernel, launch list of one decode step and
# one denoiser step, decode mode at 32 candidates, the default bench line with both reference arms
out=gpurun_out/$1; mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:ar_attn_compact -c 3 -o $out/attnc python tools/prof_kernels.py attnc > $out/ncu_attnc.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_op_profile.csv python tools/op_profile.py > $out/op_profile.log 2>&1
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline --no-ref-gpu --steps 3 --warmup 3 $EXTRA > $out/bench_$name.json 2> $out/bench_$name.err; }
EXTRA='--preset-override {"num_autoregressive_samples":32}'
run b32_fused TTB_AR_MODE=fused
run b32_mixed TTB_AR_MODE=mixed
EXTRA=''
python bench.py > $out/bench_full.json 2> $out/bench_full.err
tail -3 $out/bench_full.err
