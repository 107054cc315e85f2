#!/bin/bash
OUT=gpurun_out/${1:-r2bulk}; mkdir -p $OUT
for m in 0 1 2 3; do
  if [ $m = 0 ]; then unset S2I_GEMM_BULK; else export S2I_GEMM_BULK=$m; fi
  timeout 600 python tools/gemm_bench.py 2>/dev/null | awk '{printf "%-26s %s\n", substr($0,1,26), $(NF-5)}' > $OUT/bulk$m.txt
done
echo "shape / tensor-TMA A+B / bulk B / bulk A+B / bulk A only"
paste $OUT/bulk0.txt <(awk '{print $NF}' $OUT/bulk1.txt) <(awk '{print $NF}' $OUT/bulk2.txt) <(awk '{print $NF}' $OUT/bulk3.txt)
