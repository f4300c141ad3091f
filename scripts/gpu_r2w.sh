#!/bin/bash
# r2w visit: grouped data gradient (quads on the channel-major kernel): op tests, then A/B of the step
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "dgrad_group or fused_padding or conv2d_fwd_bwd or folded or wgrad" > $out/gpu_new_r2w.log 2>&1; tail -5 $out/gpu_new_r2w.log
for cfg in "dgrad_quad=0" "dgrad_quad=1" "dgrad_quad=0" "dgrad_quad=1"; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg --layer-table $out/layers_r2w_$cfg.json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
