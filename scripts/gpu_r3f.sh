#!/bin/bash
# r3f visit: residual-block tail fused with the next operand: tests, step-level parity, A/B of the step through the config
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "residual_block_tail or prep_fast or folded or instance_norm or group_norm or fused" > $out/gpu_new_r3f.log 2>&1; tail -4 $out/gpu_new_r3f.log
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_translation_model.py -m gpu -q -x --timeout 600 > $out/gpu_step_r3f.log 2>&1; tail -3 $out/gpu_step_r3f.log
for cfg in fuse_norm_prep=0 fuse_norm_prep=1 fuse_norm_prep=0 fuse_norm_prep=1; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg --layer-table $out/layers_r3f_$cfg.json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
