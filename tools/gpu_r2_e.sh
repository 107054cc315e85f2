OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 300 python tools/gemm_bench.py pair > $OUT/pair.txt 2>&1; echo "pair rc=$?"; cut -c1-90 $OUT/pair.txt
bash tools/gpu_round.sh r2e smoke tests
timeout 900 python bench.py --headline-only > $OUT/bench_head.json 2> $OUT/bench_head.err; echo "bench rc=$?"; tail -c 1500 $OUT/bench_head.json
