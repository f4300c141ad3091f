#!/bin/bash
# One GPU-box visit for the profile evidence of a round: plain eager bench (must exit 0) -> ncu launch list of one whole
# step -> `--set full` captures of the named kernels -> (optional) compute-sanitizer over the op tests.
# usage: scripts/gpu_profile.sh <tag>
# env: NCU_KERNELS="k1 k2" (regex per capture), NCU_SKIP / NCU_COUNT (launch-list window), SANITIZE=1
tag=${1:-x}
out=gpurun_out; mkdir -p $out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --inference 0 --stencils 0"
$CMD > $out/plain_$tag.log 2>&1 || { tail -5 $out/plain_$tag.log; exit 2; }
tail -1 $out/plain_$tag.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2800} -c ${NCU_COUNT:-950} --csv \
    --log-file $out/launches_$tag.csv $CMD > $out/ncu_l_$tag.log 2>&1
for k in ${NCU_KERNELS-conv_tc3_kernel conv_tc2_kernel tc_prep_kernel wgrad_tc2_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_KSKIP:-40} -c ${NCU_KCOUNT:-4} -f \
      -o $out/prof_${k}_$tag $CMD > $out/ncu_${k}_$tag.log 2>&1
  tail -1 $out/ncu_${k}_$tag.log
done
if [ "${SANITIZE:-0}" = "1" ]; then
  # memcheck + racecheck over the op-level parity tests at CI sizes (the tests still assert parity under the tool)
  # racecheck only sees shared-memory hazards of ordinary loads / stores: it runs over the stencil / reduction tests
  for tool in memcheck racecheck; do
    kexpr="${SANITIZE_K:-not 512}"
    [ $tool = racecheck ] && kexpr="${RACE_K:-normals or tv or smooth or masked or hole or rect or ssim or instance_norm or pad or layout}"
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $out/sanitizer_${tool}_$tag.log \
      python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 1400 -k "$kexpr" \
      > $out/sanitizer_${tool}_pytest_$tag.log 2>&1
    echo "$tool rc=$?"; tail -2 $out/sanitizer_${tool}_pytest_$tag.log; tail -3 $out/sanitizer_${tool}_$tag.log
  done
fi
