#!/bin/bash
set -x
R="compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 9"
export LMRL_NO_GRAPHS=1 LMRL_MBAR_TIMEOUT_S=900
for m in e f g h i; do timeout 600 $R python scripts/debug/dbg_race.py $m 2>&1 | tail -8; echo "mode $m rc=$?"; done
