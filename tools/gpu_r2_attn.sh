#!/bin/bash
# GPU box: attention unit tests, then device times of the attention kernels on the step's shapes
OUT=gpurun_out/${1:-r2attn}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_attn_gpu.py -x -q > $OUT/pytest_attn.log 2>&1; echo "attn tests rc=$?"; tail -3 $OUT/pytest_attn.log
timeout 600 python tools/attn_bench.py > $OUT/attn_bench.txt 2>&1; echo "attn bench rc=$?"
cat $OUT/attn_bench.txt
