#!/bin/bash
# r3c visit: single-thread, division-free role loops in wgrad_tc2: tests, A/B of the step against the previous build
out=gpurun_out; mkdir -p $out
PREV=$PWD/depth-enhancement-and-super-resolution_b200/dsr_b200/libdsr_b200_prev.so
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "conv or tc3 or dgrad or wgrad or fused or cat" > $out/gpu_new_r3c.log 2>&1; tail -4 $out/gpu_new_r3c.log
for lib in "$PREV" "" "$PREV" ""; do
  DSR_B200_LIB=$lib timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --layer-table $out/layers_r3c_$([ -z "$lib" ] && echo new || echo prev).json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$([ -z "$lib" ] && echo new || echo prev)', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['executed_frac'])"
done
