#!/bin/bash
# r2x visit: K-loop rotation in conv_tc3 (A/B through the environment), conv tests
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "dgrad_group or fused_padding or conv2d_fwd_bwd or conv_transpose or folded or fused_prologue" > $out/gpu_new_r2x.log 2>&1; tail -3 $out/gpu_new_r2x.log
for rot in 0 1 0 1; do
  DSR_TC3_ROT=$rot timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --layer-table $out/layers_r2x_rot$rot.json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rot=$rot', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['avg_launch_us'])"
done
