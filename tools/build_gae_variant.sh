#!/bin/bash
# Builds rlgym_ppo_b200/build/librlppo_gae_<tag>.so: the product library with gae_scan.cu compiled under other tuning macros
# (A/B runs inside one GPU session: RLPPO_LIB_PATH=<that file>).   usage: tools/build_gae_variant.sh <tag> -DRLPPO_GAE3_STEPS=8 ...
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
B=rlgym_ppo_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
     -I include "$@" -c rlgym_ppo_b200/csrc/gae_scan.cu -o $B/gae_scan_$tag.o 2> $B/ptxas_gae_$tag.log
objs=$(ls $B/*.o | grep -v "gae_scan")
nvcc -shared -o $B/librlppo_gae_$tag.so $objs $B/gae_scan_$tag.o -gencode arch=compute_100a,code=sm_100a
grep -A2 "gae_scan3_kernelILb1ELb1E" $B/ptxas_gae_$tag.log | grep -i "spill\|registers" | head -4
