#!/bin/bash
# A/B of variant builds (drrt_b200.build.build_variant): bash benchmarks/variants_ab.sh [what] name ...
# runs benchmarks/ab.py <what> once per drrt_b200/variants/<name>/libdrrt_b200.so
what=$1; shift
for v in "$@"; do
  DRRT_B200_LIB=$PWD/drrt_b200/variants/$v/libdrrt_b200.so timeout 300 python benchmarks/ab.py $what 2>&1 | grep -v "^$" | tail -12
done
