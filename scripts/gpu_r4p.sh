#!/bin/bash
# r4p visit: confirmation of r4o (DSR_WG2_MIN_TILES 64 / 128 against 8)
out=gpurun_out; mkdir -p $out
for v in 64 8 128 64 8; do
  DSR_WG2_MIN_TILES=$v timeout 200 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --inference 0 --stencils 0 2> $out/ab_r4p.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_times_ms']; print('min_tiles=$v', d['ms_per_step'], d['e2e']['ms_per_step'], k.get('dsr_tc_wgrad2p'))"
done
