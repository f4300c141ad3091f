#!/bin/bash
# r2z visit: weight-gradient K-split slabs: op tests, A/B of the step
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "wgrad or conv2d_fwd_bwd or conv_transpose or unpack or fused_padding" > $out/gpu_new_r2z.log 2>&1; tail -4 $out/gpu_new_r2z.log
for cfg in wgrad_slabs=0 wgrad_slabs=1 wgrad_slabs=0 wgrad_slabs=1; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg --layer-table $out/layers_r2z_$cfg.json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
