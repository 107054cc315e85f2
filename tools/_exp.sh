OUT=gpurun_out/r2early; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_path_gpu.py -x -q > $OUT/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pytest.log
for cfg in "1 220" "0 220" "1 110" "1 140"; do set -- $cfg; S2I_GEMM_EARLY_B=$1 S2I_GEMM_SMEM_CAP=$2 timeout 600 python bench.py --headline-only > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err; echo "early $1 cap $2 rc=$? $(head -c 260 $OUT/bench_$1_$2.json | grep -o '"ms_per_step": [0-9.]*')"; done
