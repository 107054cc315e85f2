#!/bin/bash
# GPU box: pair bench, smoke, the GPU test suite and the headline bench in one visit. usage (under gpurun): bash tools/gpu_r2_e.sh <tag>
OUT=gpurun_out/${1:-r2f}; mkdir -p $OUT
timeout 300 python tools/gemm_bench.py pair > $OUT/pair.txt 2>&1; echo "pair rc=$?"; cut -c1-90 $OUT/pair.txt
bash tools/gpu_round.sh ${1:-r2f} smoke tests
timeout 900 python bench.py --headline-only > $OUT/bench_head.json 2> $OUT/bench_head.err; echo "bench rc=$?"; head -c 600 $OUT/bench_head.json
