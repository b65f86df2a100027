#!/bin/bash
# One B200:  gpurun --timeout 600 -- 'bash tools/gpu_uh_variants2.sh'
# Software-pipelined flush loop of the uniform-h pair kernel (tools/build_variants.sh: uhp, uhpa, uhp2, uhpa2; built in the
# container): the uniform-h parity test on every variant, the whole GPU suite on the most aggressive one, then the bench
# workload on each next to lib_uhs.so (measured in the call before: 286.7 / 165.2 us).
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/uh_variants2.log
: > $LOG
V=$PWD/offshore-sph_b200/lib/variants
{ for v in uhp uhpa uhp2 uhpa2; do
    echo "== uniform-h parity tests on lib_$v.so"
    OSPH_LIB=$V/lib_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "uniform_h or dam_break_vs_oracle or whole_steps" 2>&1 | tail -4
  done
  echo "== pytest -m gpu on lib_uhpa2.so"
  OSPH_LIB=$V/lib_uhpa2.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
  echo "== bench: variants"
  BENCH_ARGS="--no-e2e" bash tools/bench_variants.sh uhs uhp uhpa uhp2 uhpa2
} >> $LOG 2>&1
cat $LOG
