#!/bin/bash
# GPU box, under ncu (never a bench number): warm launch list of 3 denoising steps + full captures of the top kernels
OUT=gpurun_out/${1:-r2prof}; mkdir -p $OUT
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $OUT/launches_warm.csv python tools/one_step.py 3 > $OUT/ncu_list.log 2>&1; echo list rc=$?
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tma -s 40 -c 8 -o $OUT/prof_gemm_tma_cold -f python tools/one_step.py 1 > $OUT/ncu_gemm.log 2>&1; echo gemm rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_bwd -s 0 -c 4 -o $OUT/prof_attn_bwd -f python tools/one_step.py 1 > $OUT/ncu_attnb.log 2>&1; echo attn_bwd rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -s 0 -c 2 -o $OUT/prof_attn_fwd -f python tools/one_step.py 1 > $OUT/ncu_attnf.log 2>&1; echo attn_fwd rc=$?
ls -la $OUT
