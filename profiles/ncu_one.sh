#!/bin/bash
# ncu --set full capture of the first launches matching a kernel regex in a short bench run.
# Usage (under gpurun): bash profiles/ncu_one.sh <tag> <kernel-regex> [count] [skip]
set -u
TAG=${1:-one}; KREGEX=${2:-k1}; COUNT=${3:-1}; SKIP=${4:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s $SKIP -c $COUNT -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la $OUT/prof_$TAG.ncu-rep
