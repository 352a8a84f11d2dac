#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one ncu --set full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash profiles/gpu_round.sh <tag> [kernel-regex]
# Outputs land in gpurun_out/ (scratch); the summaries worth keeping are copied into profiles/ by hand.
set -u
TAG=${1:-r1}
KREGEX=${2:-"k1m?_cost_volume|k2_conv3d|k3_depth"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log

timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_err_$TAG.log
echo "bench rc=$?"
cat $OUT/bench_$TAG.json

# launch list of the same command shape (short): per-launch device time, cold cache, serialised
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"

# full capture of the first launches of the top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -c 14 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
# memory checker over the kernel parity tests (every kernel of the library at small sizes)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_filter.py tests/test_gpu_cvp.py -x -q -m gpu -k "not cfg1_size and not training_step and not full_size and not cfg4" \
    > $OUT/sanitizer_$TAG.log 2>&1
echo "sanitizer rc=$?" | tee -a $OUT/sanitizer_$TAG.log
tail -3 $OUT/sanitizer_$TAG.log
ls -la $OUT
