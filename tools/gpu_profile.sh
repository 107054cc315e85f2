OUT=gpurun_out/${1:-r1prof}; mkdir -p $OUT
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $OUT/launches_warm.csv python tools/one_step.py 3 > $OUT/ncu_list.log 2>&1; echo list rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tma -s 40 -c 6 -o $OUT/prof_gemm_tma -f python tools/one_step.py 1 > $OUT/ncu_gemm.log 2>&1; echo gemm rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_cluster -s 4 -c 6 -o $OUT/prof_gn -f python tools/one_step.py 1 > $OUT/ncu_gn.log 2>&1; echo gn rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -s 0 -c 2 -o $OUT/prof_attn_fwd -f python tools/one_step.py 1 > $OUT/ncu_attn.log 2>&1; echo attn rc=$?
ls -la $OUT
