#!/bin/bash
# round 2, GPU call 1: parity of the new epilogue / fused LN / wide prefill attention, A/B timings, kernel sweep, short bench
set -x
O=gpurun_out/r02_call1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_policy.py -x -q -m gpu -s > $O/pytest_policy.log 2>&1; echo "policy rc=$?" >> $O/status.txt
timeout 600 python scripts/ab_options.py --reps 4 > $O/ab.log 2>&1; echo "ab rc=$?" >> $O/status.txt
timeout 600 python scripts/kernel_sweep.py --shapes 1:40,24:72 > $O/sweep.log 2>&1; echo "sweep rc=$?" >> $O/status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "all rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
tail -5 $O/pytest_policy.log; cat $O/ab.log; cat $O/sweep.log; tail -5 $O/pytest_all.log; cat $O/status.txt
