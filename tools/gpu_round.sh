#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench (+ reference arm), ncu launch list and one full capture.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [what...]   what in: smoke tests bench ref ncu_list ncu_full
set -u
TAG=${1:-r1}; shift || true
WHAT=${*:-smoke tests bench ncu_list ncu_full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
for w in $WHAT; do
  case $w in
    smoke) timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" ;;
    tests) timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log ;;
    bench) timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json ;;
    ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" ;;
    ncu_list) timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
                --log-file $OUT/launches.csv python tools/one_step.py 3 > $OUT/ncu_list.log 2>&1; echo "ncu_list rc=$?" ;;
    ncu_full) timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:gemm_tc -s 20 -c 4 -o $OUT/prof_gemm -f python tools/one_step.py 1 > $OUT/ncu_full.log 2>&1; echo "ncu_full rc=$?" ;;
  esac
done
ls -la $OUT
