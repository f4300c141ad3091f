#!/bin/bash
# r4j visit: more of the small-map layers on the channel-major kernel (halo_min_tiles): A/B of the step and the GEMM kernel times
out=gpurun_out; mkdir -p $out
for cfg in halo_min_tiles=120 halo_min_tiles=90 halo_min_tiles=60 halo_min_tiles=120 halo_min_tiles=90 halo_min_tiles=60; do
  timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg --layer-table $out/layers_r4j_$cfg.json 2> $out/ab_r4j.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_times_ms']; print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], {n: k.get(n) for n in ('dsr_tc_gemm','dsr_tc_gemm2','dsr_tc_gemm3')})"
done
