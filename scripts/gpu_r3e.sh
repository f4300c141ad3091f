#!/bin/bash
# r3e visit: gen-1 kernel loops (A/B against the previous build), then the whole GPU suite and the bench line with tables
out=gpurun_out; mkdir -p $out
PREV=$PWD/depth-enhancement-and-super-resolution_b200/dsr_b200/libdsr_b200_prev.so
for lib in "$PREV" "" "$PREV" ""; do
  DSR_B200_LIB=$lib timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$([ -z "$lib" ] && echo new || echo prev)', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
bash scripts/gpu_round.sh r3e
