#!/bin/bash
# GPU box: GEMM unit tests, per-shape timings (single CTAs / CTA pairs / the model's choice), tile x cluster sweep of both forms
OUT=gpurun_out/${1:-r2pair}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > $OUT/pytest_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -3 $OUT/pytest_gemm.log
timeout 300 python tools/gemm_bench.py pair > $OUT/pair.txt 2>&1; echo "pair rc=$?"
cat $OUT/pair.txt
timeout 900 python tools/gemm_bench.py psweep > $OUT/psweep.txt 2>&1; echo "psweep rc=$?"
cat $OUT/psweep.txt
