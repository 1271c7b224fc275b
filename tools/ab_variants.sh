#!/bin/bash
# dev tool: A/B the shipped library against variants/*.so on the north-star shape (two rounds, same box).
# usage: tools/ab_variants.sh [n] [p]
N=${1:-1e8}; P=${2:-64}
for round in 1 2; do
  for lib in anml_b200/lib/libanml_b200.so variants/*.so; do
    echo -n "$lib "; ANML_B200_LIB=$PWD/$lib timeout 200 python tools/time_fused.py $N $P 1 2>&1 | tail -1 | cut -c1-120
  done
done
