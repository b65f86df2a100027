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
  prep4    "-DPREP_ITEMS=4" \
  r2gate   "-DPAIR_UH=0 -DPAIR_ISIGN=0 -DPAIR_UH_PIPE=0 -DPAIR_GEN_PIPE=0" \
  isign    "-DPAIR_UH=0 -DPAIR_UH_PIPE=0 -DPAIR_GEN_PIPE=0" \
  uhs      "-DPAIR_UH_PIPE=0 -DPAIR_GEN_PIPE=0" \
  uhp1     "-DPAIR_PIPE_UNROLL=1 -DPAIR_GEN_PIPE=0" \
  uhp3     "-DPAIR_PIPE_UNROLL=3 -DPAIR_GEN_PIPE=0" \
  gp2      "-DPAIR_GEN_PIPE=3 -DPAIR_GEN_PIPE_UNROLL=2" \
  gp1f     "-DPAIR_GEN_PIPE=3"
# (the last seven: the uniform-h / pipelined-flush knobs of the pair kernel against the default build = uniform-h instantiation
#  pipelined in both precisions and unrolled by two, general double instantiation pipelined; r2gate = the kernel before them;
#  measured with tools/gpu_uh_variants*.sh, profiles/r02/variants_uniform_h.log, variants_pipelined*.log)
