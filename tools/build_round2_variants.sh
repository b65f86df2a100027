#!/bin/bash
# Compile-time variants of the pair kernel / predictor pass that were measured in round 2 (profiles/r02/README.md has the
# numbers).  Run HERE (nvcc cross-compiles), then on the GPU box:  bash tools/bench_variants.sh <names>.
cd "$(dirname "$0")/.."
tools/build_variants.sh \
  t128     "-DOSPH_PAIR_THREADS=128 -DPAIR_CAP=512 -DPAIR_MINB64=4 -DPAIR_MINB32=6" \
  newton2  "-DOSPH_NEWTON_STEPS=2" \
  scan4    "-DPAIR_SCAN=4" \
  scan8    "-DPAIR_SCAN=8" \
  list64   "-DPAIR_LIST64=64 -DPAIR_LIST32=64" \
  lean0    "-DPAIR_LEAN=0" \
  scan2    "-DPAIR_SCAN2=1" \
  scan2p   "-DPAIR_SCAN2=1 -DPAIR_LISTPTR=1" \
  scan2p8  "-DPAIR_SCAN2=1 -DPAIR_LISTPTR=1 -DPAIR_SCAN=8" \
  prep1    "-DPREP_ITEMS=1" \
  prep4    "-DPREP_ITEMS=4"
