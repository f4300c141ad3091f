#!/bin/bash
# ncu --set full on single layers of scripts/bench_layers.py; usage: gpu_prof_layers.sh <tag> <layer-substr> [...]
tag=$1; shift
for l in "$@"; do
  name=$(echo $l | tr ' >' '__')
  CMD="python scripts/bench_layers.py --only $l --halo 1 --baseoff 0 --iters 2"
  $CMD > gpurun_out/lp_${name}_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 2 -f -o gpurun_out/prof_${name}_$tag $CMD > gpurun_out/ncu_${name}_$tag.log 2>&1
  tail -1 gpurun_out/lp_${name}_$tag.log | cut -c1-300
done
