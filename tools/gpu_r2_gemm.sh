#!/bin/bash
# GPU box: GEMM unit tests, per-shape timings with the chosen tiles, and the deterministic split-K sweep
OUT=gpurun_out/${1:-r2gemm}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > $OUT/pytest_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -3 $OUT/pytest_gemm.log
S2I_GEMM_DEBUG=1 timeout 600 python tools/gemm_bench.py > $OUT/bench_cluster.txt 2> $OUT/choices_cluster.txt; echo cluster rc=$?
grep "can be resident" $OUT/choices_cluster.txt | sort -u
awk '{printf "%-26s %s\n", substr($0,1,26), $(NF-5)}' $OUT/bench_cluster.txt
timeout 900 python tools/gemm_bench.py dsweep > $OUT/dsweep.txt 2>&1; echo dsweep rc=$?
cat $OUT/dsweep.txt
