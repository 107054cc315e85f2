#!/bin/bash
# GPU box: per-shape timings (single CTAs / CTA pairs / the model's choice), then the tile x cluster sweep of both forms
OUT=gpurun_out/${1:-r2pair}; mkdir -p $OUT
timeout 300 python tools/gemm_bench.py pair > $OUT/pair.txt 2>&1; echo "pair rc=$?"
cut -c1-100 $OUT/pair.txt
timeout 900 python tools/gemm_bench.py psweep > $OUT/psweep.txt 2>&1; echo "psweep rc=$?"
cat $OUT/psweep.txt
