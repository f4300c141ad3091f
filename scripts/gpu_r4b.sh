#!/bin/bash
# r4b visit: skip-gradient fusion, register-blocked one-channel heads: new tests, step parity, A/B through the switches
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "out1 or register_blocked or skip_gradient or in_bwd or norm_backward or residual_block_tail or conv2d_fwd_bwd or conv_transpose2d_fwd_bwd" > $out/gpu_new_r4b.log 2>&1; tail -15 $out/gpu_new_r4b.log
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 600 > $out/gpu_step_r4b.log 2>&1; tail -5 $out/gpu_step_r4b.log
for v in "DSR_OUT1_RB=0 fuse_skip_grad=0" "DSR_OUT1_RB=1 fuse_skip_grad=1" "DSR_OUT1_RB=0 fuse_skip_grad=1" "DSR_OUT1_RB=1 fuse_skip_grad=0" "DSR_OUT1_RB=0 fuse_skip_grad=0" "DSR_OUT1_RB=1 fuse_skip_grad=1"; do
  set -- $v
  env $1 timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --cfg $2 --layer-table "$out/layers_r4b_$1_$2.json" 2> $out/ab_r4b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
