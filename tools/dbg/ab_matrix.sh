#!/bin/bash
# A/B matrix of launch options (one short bench.py run each); results in gpurun_out/$1/
out=gpurun_out/$1; mkdir -p $out
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline --no-ref-gpu --steps 3 --warmup 3 $EXTRA > $out/bench_$name.json 2> $out/bench_$name.err; }
run default      TTB_AR_SPLITK_PROJ=2 TTB_AR_SPLITK_PROJ2=4
run sk_4_4       TTB_AR_SPLITK_PROJ=4 TTB_AR_SPLITK_PROJ2=4
run sk_4_8       TTB_AR_SPLITK_PROJ=4 TTB_AR_SPLITK_PROJ2=8
run sk_2_8       TTB_AR_SPLITK_PROJ=2 TTB_AR_SPLITK_PROJ2=8
