#!/bin/bash
# One GPU-box visit: parity tests -> bench -> ncu launch list + full captures of the top kernels.
# usage: scripts/gpu_check.sh <tag> [pytest -k expression for the ops tests]
# env: NCU=0 skips profiling; NCU_KERNELS="k1 k2" picks the kernels of the --set full captures ("" = launch list only)
tag=${1:-x}; kexpr=${2:-}
out=gpurun_out
mkdir -p $out
if [ -n "$kexpr" ]; then
  CUDA_LAUNCH_BLOCKING=1 timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "$kexpr" > $out/gpu_ops_$tag.log 2>&1
else
  CUDA_LAUNCH_BLOCKING=1 timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 > $out/gpu_ops_$tag.log 2>&1
fi
rc1=$?; tail -3 $out/gpu_ops_$tag.log
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 600 > $out/gpu_step_$tag.log 2>&1
rc2=$?; tail -3 $out/gpu_step_$tag.log
if [ $rc1 -ne 0 ] || [ $rc2 -ne 0 ]; then echo "tests failed ($rc1,$rc2): skipping bench/ncu"; exit 1; fi
timeout 900 python bench.py --steps 10 --warmup 3 --layer-table $out/layers_step_$tag.json --kernel-table $out/kernels_step_$tag.json > $out/bench_$tag.json 2> $out/bench_$tag.err || { tail -5 $out/bench_$tag.err; exit 2; }
tail -1 $out/bench_$tag.json | cut -c1-400
if [ "${NCU:-1}" = "1" ]; then
  # eager launches under ncu (a graph replay would hide the library-call boundaries): 3 warm-up + 2 timed + 2 e2e + 1 profile steps
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --inference 0"
  $CMD > $out/plain_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-4600} -c ${NCU_COUNT:-1200} --csv --log-file $out/launches_$tag.csv $CMD > $out/ncu_l_$tag.log 2>&1
  for k in ${NCU_KERNELS-conv_tc3_kernel conv_tc2_kernel}; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_KSKIP:-60} -c 3 -f -o $out/prof_${k}_$tag $CMD > $out/ncu_${k}_$tag.log 2>&1
    tail -2 $out/ncu_${k}_$tag.log
  done
fi
