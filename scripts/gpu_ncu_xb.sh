#!/bin/bash
# launch order of kernel_sweep --iters 1 per shape: qkv x3, proj x3, fc x3, fc2 x3, head x3 (third of each = warm)
# ncu --set full of the LayerNorm-fold residual GEMM (EPI_RESID, XB) on the sweep shapes; source-level stall view of one prefill and one decode launch
O=gpurun_out/${1:-r02_xb}
mkdir -p $O
LMRL_OPTIONS=ln_fold=1 timeout 600 ncu --set full --clock-control none --import-source on -k gemm_bf16_tcgen05_kernel -c 27 -o $O/prof_xb \
    python scripts/kernel_sweep.py --iters 1 --shapes 1:40,18:72 > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/status.txt
python scripts/ncu_summarize.py full $O/prof_xb.ncu-rep > $O/ncu_full_summary.txt 2>&1
ncu -i $O/prof_xb.ncu-rep --page source --csv --launch-skip 20 --launch-count 1 > $O/xb_prefill_proj_source.csv 2>/dev/null
ncu -i $O/prof_xb.ncu-rep --page source --csv --launch-skip 5 --launch-count 1 > $O/xb_decode_proj_source.csv 2>/dev/null
ncu -i $O/prof_xb.ncu-rep --page details --csv --launch-skip 20 --launch-count 1 > $O/xb_prefill_proj_details.csv 2>/dev/null
gzip -9 $O/*.csv
rm -f $O/prof_xb.ncu-rep
cat $O/status.txt; ls -la $O
