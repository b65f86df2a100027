#!/bin/bash
# One B200:  gpurun --timeout 600 -- 'bash tools/gpu_uh_variants.sh'
# The uniform-h / sign-bit variants of the pair kernel (tools/build_variants.sh uh "-DPAIR_UH=1" uhs "-DPAIR_UH=1 -DPAIR_ISIGN=1"
# isign "-DPAIR_ISIGN=1", built in the container): the whole GPU suite on the candidate build (incl. the bit-identity test
# against the general instantiation), then the bench workload on every variant next to the default build.
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/uh_variants.log
: > $LOG
V=$PWD/offshore-sph_b200/lib/variants
{ echo "== pytest -m gpu on lib_uhs.so"
  OSPH_LIB=$V/lib_uhs.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
  echo "== bench: default build"
  for p in fp64 fp32; do
    timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --precision $p | \
      python -c "import json,sys; d=json.load(sys.stdin); print('default $p value %.4e ms/step %.4f pair_us %.1f e2e %.3e status %s'%(d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['e2e']['value'], d['status_bits']))"
  done
  echo "== bench: variants"
  BENCH_ARGS="" bash tools/bench_variants.sh uh uhs isign
  echo "== bench: uhs, 16 M particles on one GPU"
  OSPH_LIB=$V/lib_uhs.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --particles-per-side 4000 | \
      python -c "import json,sys; d=json.load(sys.stdin); print('uhs fp64 16M value %.4e ms/step %.4f pair_us %.1f status %s'%(d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['status_bits']))"
} >> $LOG 2>&1
cat $LOG
