#!/bin/bash
# usage: ab_bench.sh <so>   -- runs bench.py (short) with the given library copied over the product library
cp wild_deep_mvs_b200/libmvsb200.so /tmp/lib_keep.so
cp "$1" wild_deep_mvs_b200/libmvsb200.so
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['cfg5']['ms_per_64_views']); print({k['name']: round(k['ms'],4) for k in d['kernels'] if k['name'] in ('conv0','conv11','conv1')})"
cp /tmp/lib_keep.so wild_deep_mvs_b200/libmvsb200.so
