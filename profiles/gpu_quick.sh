#!/bin/bash
# Short GPU visit while iterating on a kernel: the K2 parity tests, a bench line without the CPU leg, and a full ncu
# capture of the first launches matching a kernel regex.   bash profiles/gpu_quick.sh <tag> <pytest -k expr> [kernel-regex] [count]
set -u
TAG=${1:-q}
KEXPR=${2:-"k2_"}
KREGEX=${3:-"k2_conv3d_zm"}
COUNT=${4:-4}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_err_$TAG.log
echo "bench rc=$?"; tail -3 $OUT/bench_err_$TAG.log
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else None)
for k in d["kernels"]: print("  %-16s %.3f ms %.0f GB/s" % (k["name"], k["ms"], k["GBps"]))
PY
if [ "$COUNT" != "0" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -c $COUNT -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
fi
