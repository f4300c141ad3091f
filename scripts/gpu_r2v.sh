#!/bin/bash
# r2v visit: new-kernel tests, prep micro-benchmark, A/B of the step with / without the folded finalize and the replicated sums
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "folded or instance_norm or group_norm or fused_prologue" > $out/gpu_new_r2v.log 2>&1; tail -3 $out/gpu_new_r2v.log
timeout 300 python scripts/bench_prep.py --out $out/prep_r2v.json > $out/prep_r2v.log 2>&1; tail -40 $out/prep_r2v.log | cut -c1-220
for cfg in "fold_finalize=0,csum_reps=1" "fold_finalize=1,csum_reps=1" "fold_finalize=1,csum_reps=8" "fold_finalize=0,csum_reps=1" "fold_finalize=1,csum_reps=8"; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
