#!/usr/bin/env bash
# compute-sanitizer over the small GPU tests that cover every kernel family (run on a B200 box; a few minutes).
#   bash scripts/sanitize.sh [outdir]        -> <outdir>/sanitizer_*.log, verdict lines on stdout
# memcheck runs with CUDA graphs as shipped; racecheck runs with LMRL_NO_GRAPHS=1 (eager launches) — under graph capture the tool
# stopped without a verdict in round 1 (profiles/r01_sanitizer_summary.txt).
set -u
out=${1:-gpurun_out}
mkdir -p "$out"
M="compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9"
R="compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 9"
run() {  # name, limit seconds, tool..., -- pytest args
    local name=$1 limit=$2; shift 2
    timeout "$limit" "$@" > "$out/sanitizer_$name.log" 2>&1
    echo "$name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out/sanitizer_$name.log" | tail -1)"
}
T="python -m pytest -m gpu -q -x"
run mem_rollout   150 $M $T tests/test_gpu_policy.py -k fused_rollout_matches
run mem_backward  120 $M $T tests/test_gpu_backward.py -k "edge_cases or bc_term"
run mem_ppo        90 $M $T tests/test_gpu_ppo.py -k "reference_vectors or forward_from_str"
run mem_value     150 $M $T tests/test_gpu_value_policy.py tests/test_gpu_textnav.py tests/test_gpu_value_forward.py -k "fused_rollout or random_walks or linear"
run mem_env       120 $M $T tests/test_gpu_maze_fused.py tests/test_gpu_env.py -k "early_exit or goldens"
LMRL_NO_GRAPHS=1 run race_rollout  300 $R $T tests/test_gpu_policy.py -k fused_rollout_matches
run race_backward 200 $R $T tests/test_gpu_backward.py -k edge_cases
