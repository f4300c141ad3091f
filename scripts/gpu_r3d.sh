#!/bin/bash
# r3d visit: fast operand-preparation kernel: bit-identity tests, prep micro-benchmark, A/B of the step through DSR_PREP_FAST
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "prep_fast or folded or conv2d_fwd_bwd or conv_transpose or fused or cat_conv or dgrad_group" > $out/gpu_new_r3d.log 2>&1; tail -4 $out/gpu_new_r3d.log
for f in 0 1; do echo "DSR_PREP_FAST=$f"; DSR_PREP_FAST=$f timeout 300 python scripts/bench_prep.py 2>&1 | grep -v "fold_or_reps': 1,\|fold_or_reps': False" | cut -c1-200; done
for f in 0 1 0 1; do
  DSR_PREP_FAST=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --layer-table $out/layers_r3d_fast$f.json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fast=$f', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
