#!/bin/bash
# r4h visit: running per-sample statistics in the conv_tc2 epilogue (same-address fp64 atomics): tests, step parity, A/B
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "stats or conv2d_fwd_bwd or conv_transpose2d_fwd_bwd or instance_norm or group_norm or folded or residual_block or norm_backward" > $out/gpu_new_r4h.log 2>&1; tail -4 $out/gpu_new_r4h.log
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_translation_model.py -m gpu -q -x --timeout 600 > $out/gpu_step_r4h.log 2>&1; tail -3 $out/gpu_step_r4h.log
for v in "DSR_TC2_RUNSTATS=0 DSR_TC2_WIDE=1" "DSR_TC2_RUNSTATS=1 DSR_TC2_WIDE=1" "DSR_TC2_RUNSTATS=1 DSR_TC2_WIDE=0" "DSR_TC2_RUNSTATS=0 DSR_TC2_WIDE=1" "DSR_TC2_RUNSTATS=1 DSR_TC2_WIDE=1" "DSR_TC2_RUNSTATS=1 DSR_TC2_WIDE=0"; do
  env $v timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --layer-table "$out/layers_r4h_$(echo $v | tr ' =' '__').json" 2> $out/ab_r4h.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_times_ms'].get('dsr_tc_gemm2'))"
done
