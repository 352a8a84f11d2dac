#!/bin/bash
# ncu --set full capture of the first cfg2 launch of each K1 variant (new hypothesis-major kernel, old pixel-major kernel).
# Usage (under gpurun): bash profiles/ncu_k1.sh <tag>
set -u
TAG=${1:-k1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k1h?_cost_volume" -s 2 -c 1 -f -o $OUT/prof_${TAG}_h \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/ncu_${TAG}_h.log 2>&1
echo "ncu h rc=$?"
MVSB200_K1=v1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k1h?_cost_volume" -s 2 -c 1 -f -o $OUT/prof_${TAG}_v1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/ncu_${TAG}_v1.log 2>&1
echo "ncu v1 rc=$?"
ls -la $OUT/*.ncu-rep
