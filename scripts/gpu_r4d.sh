#!/bin/bash
# r4d visit: packed weight copies re-made in order of first use (the forward pass follows the packing chain): step parity + A/B,
# in-situ kernel table / timeline of the new default
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 600 > $out/gpu_step_r4d.log 2>&1; tail -3 $out/gpu_step_r4d.log
for cfg in prepack_in_use_order=0 prepack_in_use_order=1 prepack_in_use_order=0 prepack_in_use_order=1; do
  timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg 2> $out/ab_r4d.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
timeout 600 python bench.py --steps 20 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --kernel-table $out/r4d_kernels_in_situ.json > $out/r4d_bench_kt.log 2>&1; tail -2 $out/r4d_bench_kt.log | cut -c1-400
