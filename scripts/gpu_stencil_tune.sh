#!/bin/bash
# Stencil iteration on one GPU box: parity tests of the stencil ops, then the HBM table of the chosen cases under a set of
# environment settings (one process per setting: the knobs are read once).  usage: scripts/gpu_stencil_tune.sh <tag> "<only>" "<env1>" "<env2>" ...
tag=$1; only=$2; shift 2
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x --timeout 600 -k "${TEST_K:-normals or tv or stencils or smooth or step}" > $out/stencil_tests_$tag.log 2>&1
rc=$?; tail -4 $out/stencil_tests_$tag.log
[ $rc -ne 0 ] && { grep -E "Error|error|assert" $out/stencil_tests_$tag.log | head -20; exit 1; }
i=0
for e in "" "$@"; do
  echo "== env: $e"
  env $e python scripts/bench_stencils.py --only "$only" --out $out/stencils_${tag}_$i.json 2>&1 | grep -v Warning
  i=$((i+1))
done
