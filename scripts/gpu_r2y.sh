#!/bin/bash
# r2y visit: 64-channel stages in the single-pass channel-major kernel: op tests, layer micro-benchmark with wait counters, A/B of the step
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "tc3_single or dgrad_group or default_dgrad or conv2d_fwd_bwd or conv_transpose" > $out/gpu_new_r2y.log 2>&1; tail -4 $out/gpu_new_r2y.log
for kb in 32 64; do DSR_TC3_KB=$kb DSR_BENCH_WAITS=1 timeout 300 python scripts/bench_dgrad_quad.py --g 4 2>&1 | tail -1 | cut -c1-600; done
for kb in 32 64 32 64; do
  DSR_TC3_KB=$kb timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --inference 0 --stencils 0 --layer-table $out/layers_r2y_kb$kb.json 2> $out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kb=$kb', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['avg_launch_us'])"
done
