#!/bin/bash
# Experiment build: libmvsb200 with one source recompiled under extra -D flags, linked against the objects of the normal build.
#   bash profiles/build_variant.sh <tag> <source.cu> <nvcc flags...>   ->  wild_deep_mvs_b200/build/libmvsb200_<tag>.so
set -e
cd "$(dirname "$0")/.."
TAG=$1; SRC=$2; shift 2
B=wild_deep_mvs_b200/build
python -m wild_deep_mvs_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I include -I wild_deep_mvs_b200/csrc \
    "$@" -c wild_deep_mvs_b200/csrc/$SRC -o $B/${SRC}_$TAG.o 2>&1 | grep -v "deprecated-gpu-targets" || true
OBJS=$(ls $B/*.cu.o | grep -v "/$SRC.o")
nvcc --shared -o $B/libmvsb200_$TAG.so $OBJS $B/${SRC}_$TAG.o -ldl 2>&1 | grep -v "deprecated-gpu-targets" || true
ls -la $B/libmvsb200_$TAG.so
