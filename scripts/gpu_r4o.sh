#!/bin/bash
# r4o visit: fewer, longer K splits in wgrad_tc2 (DSR_WG2_MIN_TILES): step time and weight-gradient kernel time
out=gpurun_out; mkdir -p $out
for v in 8 32 64 16 8 32; do
  DSR_WG2_MIN_TILES=$v timeout 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --inference 0 --stencils 0 2> $out/ab_r4o.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_times_ms']; print('min_tiles=$v', d['ms_per_step'], k.get('dsr_tc_wgrad2p'), k.get('dsr_tc_unpack_wgrad_splits'))"
done
