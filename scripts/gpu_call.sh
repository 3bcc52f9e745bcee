#!/bin/bash
# usage: bash scripts/gpu_call.sh <tag> [steps...]   steps: policy all ab sweep bench benchfull sanitize
# every step writes into gpurun_out/<tag>/ and appends its exit code to status.txt
set -x
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
for step in "$@"; do
  case $step in
    policy) timeout 900 python -m pytest tests/test_gpu_policy.py -x -q -m gpu -s > $O/pytest_policy.log 2>&1;;
    all) timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1;;
    ab) timeout 600 python scripts/ab_options.py --reps 4 ${AB_ARGS} > $O/ab.log 2>&1;;
    sweep) timeout 600 python scripts/kernel_sweep.py --shapes 1:40,24:72 > $O/sweep.log 2>&1;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err;;
    benchfull) timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err;;
    ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err;;
    *) timeout 1800 bash -c "$step" > $O/custom_$(echo "$step" | md5sum | cut -c1-6).log 2>&1;;
  esac
  echo "$step rc=$?" >> $O/status.txt
done
for f in $O/pytest_policy.log $O/pytest_all.log; do [ -f $f ] && tail -4 $f; done
for f in $O/ab.log $O/sweep.log $O/custom_*.log; do [ -f $f ] && tail -30 $f; done
cat $O/status.txt
