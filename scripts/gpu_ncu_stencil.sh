#!/bin/bash
# ncu --set full of one stencil case of scripts/bench_stencils.py.  usage: scripts/gpu_ncu_stencil.sh <tag> "<only>" <kernel regex> [env...]
tag=$1; only=$2; k=$3; shift 3
out=gpurun_out; mkdir -p $out
env "$@" python scripts/bench_stencils.py --only "$only" --iters 3 > $out/ncu_plain_$tag.log 2>&1 || { tail -5 $out/ncu_plain_$tag.log; exit 2; }
cat $out/ncu_plain_$tag.log | grep -v Warn
env "$@" ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/prof_stencil_$tag python scripts/bench_stencils.py --only "$only" --iters 3 > $out/ncu_stencil_$tag.log 2>&1
tail -2 $out/ncu_stencil_$tag.log
