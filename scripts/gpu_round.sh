#!/bin/bash
# One GPU-box visit: the whole GPU test suite, then the bench line (+ per-kernel table).  usage: scripts/gpu_round.sh <tag> [pytest args]
tag=${1:-x}; shift
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 "$@" > $out/gpu_tests_$tag.log 2>&1
rc=$?; tail -15 $out/gpu_tests_$tag.log
[ -f $out/parity_margins.jsonl ] && cat $out/parity_margins.jsonl | cut -c1-600
if [ "${BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 --kernel-table $out/kernels_step_$tag.json --layer-table $out/layers_step_$tag.json > $out/bench_$tag.json 2> $out/bench_$tag.err || { tail -20 $out/bench_$tag.err; exit 2; }
  tail -1 $out/bench_$tag.json | cut -c1-1500
fi
exit $rc
