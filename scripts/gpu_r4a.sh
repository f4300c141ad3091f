#!/bin/bash
# r4a visit: InstanceNorm backward fused with the dY operand of the convolution in front of it (dsr_tc_prep_in_bwd):
# kernel-level bitwise test, layer-stack test, step-level parity, A/B of the step through the config switch
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "in_bwd or norm_backward or residual_block_tail or folded or instance_norm" > $out/gpu_new_r4a.log 2>&1; tail -15 $out/gpu_new_r4a.log
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 600 > $out/gpu_step_r4a.log 2>&1; tail -5 $out/gpu_step_r4a.log
for cfg in fuse_bwd_prep=0 fuse_bwd_prep=1 fuse_bwd_prep=0 fuse_bwd_prep=1; do
  timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --cfg $cfg --layer-table $out/layers_r4a_$cfg.json 2> $out/ab_r4a.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
DSR_BENCH_WAITS=1 timeout 300 python scripts/bench_layers.py --halo 1 --only "first 7x7,resnet up2,down 3x3 s2 32->64,unet down 4x4 s2 261" --json $out/r4a_layers_waits.json > $out/r4a_layers_waits.log 2>&1; tail -12 $out/r4a_layers_waits.log
