#!/bin/bash
# round-2 evidence run on one B200: ncu launch list of one rollout, ncu --set full of the hot kernels (GEMM epilogues, paged attention,
# GEMV path, env kernels), summarised ON THE BOX (the .ncu-rep files are too large to bring back: 64 MiB limit); writes gpurun_out/r02_prof/.
#   bash scripts/gpu_profile.sh [main|ncu|sanitize|all] [out-dir-name]      main = launch list + ncu --set full of the GEMM / attention / LayerNorm kernels only
set -x
WHAT=${1:-all}
O=gpurun_out/${2:-r02_prof}
mkdir -p $O
SKIP=variant_2c,small_batch,maze_po_medium,wordle_ilql,ppo_round,letters,env_only,e2e_materialized
if [ "$WHAT" = ncu ] || [ "$WHAT" = all ] || [ "$WHAT" = main ]; then
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3500 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --skip $SKIP > $O/bench_under_ncu.log 2>&1; echo "launchlist rc=$?" >> $O/status.txt
python scripts/ncu_summarize.py launches $O/launches.csv > $O/launch_summary.txt 2>&1
python scripts/ncu_summarize.py traffic $O/launches.csv $O/traffic.json >> $O/launch_summary.txt 2>&1
gzip -9 $O/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05|attention_|layernorm" -c 44 -o $O/prof_kernels \
    python scripts/kernel_sweep.py --iters 1 --shapes 1:40,24:72 > $O/ncu_kernels.log 2>&1; echo "full_kernels rc=$?" >> $O/status.txt
python scripts/ncu_summarize.py full $O/prof_kernels.ncu-rep > $O/ncu_full_summary.txt 2>&1
ncu -i $O/prof_kernels.ncu-rep --page source --csv -k regex:"attention_prefill" --launch-skip 0 --launch-count 1 > $O/attn_prefill_source.csv 2>/dev/null
ncu -i $O/prof_kernels.ncu-rep --page source --csv -k regex:"gemm_bf16_tcgen05" --launch-skip 23 --launch-count 1 > $O/gemm_qkv_source.csv 2>/dev/null
ncu -i $O/prof_kernels.ncu-rep --page details --csv > $O/ncu_details.csv 2>/dev/null
gzip -9 $O/attn_prefill_source.csv $O/gemm_qkv_source.csv $O/ncu_details.csv
rm -f $O/prof_kernels.ncu-rep
fi
if [ "$WHAT" = ncu ] || [ "$WHAT" = all ]; then
LMRL_NO_GRAPHS=1 timeout 600 ncu --set full --clock-control none -k regex:"gemv_kernel" -s 70 -c 12 -o $O/prof_gemv \
    python scripts/small_batch_bench.py --batches 1 --new 4 > $O/ncu_gemv.log 2>&1; echo "full_gemv rc=$?" >> $O/status.txt
python scripts/ncu_summarize.py full $O/prof_gemv.ncu-rep > $O/ncu_gemv_summary.txt 2>&1
rm -f $O/prof_gemv.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:"wordle_step|wordle_reset" -c 4 -o $O/prof_env \
    python scripts/env_microbench.py > $O/ncu_env.log 2>&1; echo "full_env rc=$?" >> $O/status.txt
python scripts/ncu_summarize.py full $O/prof_env.ncu-rep > $O/ncu_env_summary.txt 2>&1
rm -f $O/prof_env.ncu-rep
fi
if [ "$WHAT" = sanitize ] || [ "$WHAT" = all ]; then
M="compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9"
R="compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 9"
T="python -m pytest -m gpu -q -x"
run() { local name=$1 limit=$2; shift 2; timeout "$limit" "$@" > "$O/sanitizer_$name.log" 2>&1; echo "$name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$O/sanitizer_$name.log" | tail -1)" >> $O/status.txt; tail -c 3000 "$O/sanitizer_$name.log"; }
run mem_small_batch 300 $M $T tests/test_gpu_small_batch.py -k "five_episodes or maze_batch_one"
run mem_paged 300 $M $T tests/test_gpu_paged_kv.py -k "exhausted or small_pool_is"
run mem_rollout 300 $M $T tests/test_gpu_policy.py -k "fused_rollout_matches or bit_identical or joined_history"
LMRL_NO_GRAPHS=1 run race_small_batch 600 $R $T tests/test_gpu_small_batch.py -k five_episodes
LMRL_NO_GRAPHS=1 run race_rollout 1200 $R $T tests/test_gpu_policy.py -k fused_rollout_matches
fi
cat $O/status.txt
du -sh $O; ls -la $O
