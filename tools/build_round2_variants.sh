#!/bin/bash
# The variant set worth one GPU measurement at the start of the next round.  Run HERE (nvcc
# cross-compiles), then `gpurun -- 'bash tools/gpu_round_start.sh'`: the script benches every lib/variants/lib_*.so it finds.
cd "$(dirname "$0")/.."
tools/build_variants.sh \
  t128     "-DOSPH_PAIR_THREADS=128 -DPAIR_CAP=512 -DPAIR_MINB64=4 -DPAIR_MINB32=6" \
  newton2  "-DOSPH_NEWTON_STEPS=2" \
  scan4    "-DPAIR_SCAN=4" \
  scan8    "-DPAIR_SCAN=8" \
  list64   "-DPAIR_LIST64=64 -DPAIR_LIST32=64" \
  lean0    "-DPAIR_LEAN=0"
