#!/bin/bash
# r4c visit: conv_tc2 "wide" parity mode (three products in two MMAs per K step): layer errors + times, op tests, step parity, A/B
out=gpurun_out; mkdir -p $out
for wide in 0 1; do
  DSR_TC2_WIDE=$wide DSR_BENCH_WAITS=1 timeout 300 python scripts/bench_layers.py --halo 1 --only "first 7x7,resnet up2,down 3x3 s2 32->64,unet down 4x4 s2 261,unet down 4x4 s2 64->128,unet up convT 4x4 256->64,resnet up convT 3x3 128->64" --json $out/r4c_layers_wide$wide.json > $out/r4c_layers_wide$wide.log 2>&1
  python - <<P
import json
for r in json.load(open("$out/r4c_layers_wide$wide.json")):
    w = r.get("waits") or {}
    print("wide=$wide", r["name"], "err %.2e" % r["err"], "gemm %.4f ms" % r["ms_gemm"], r["kernels"], "mma_total", int(w.get("mma_total", [0])[0]))
P
done
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "out1 or register_blocked or conv2d_fwd_bwd or conv_transpose2d_fwd_bwd or cat_conv or first_layer or stats" > $out/gpu_new_r4c.log 2>&1; tail -8 $out/gpu_new_r4c.log
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 600 > $out/gpu_step_r4c.log 2>&1; tail -3 $out/gpu_step_r4c.log
for v in 0 1 0 1; do
  DSR_TC2_WIDE=$v timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --inference 0 --stencils 0 --layer-table "$out/layers_r4c_wide$v.json" 2> $out/ab_r4c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wide=$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
