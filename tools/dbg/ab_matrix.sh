#!/bin/bash
# A/B matrix of launch options (one short bench.py run each); results in gpurun_out/$1/
out=gpurun_out/$1; mkdir -p $out
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline --no-ref-gpu --steps 3 --warmup 3 $EXTRA > $out/bench_$name.json 2> $out/bench_$name.err; }
run diffchains1  TTB_AR_CHAINS=2 TTB_DIFF_CHAINS=1
run diffchains0  TTB_AR_CHAINS=2 TTB_DIFF_CHAINS=0
run archains4    TTB_AR_CHAINS=4 TTB_AR_CHAINS_MIN_B=64 TTB_DIFF_CHAINS=0
