#!/bin/bash
# Compare pair-kernel build variants (lib/variants/lib_*.so) on the bench workload.
for v in "$@"; do
  for p in fp64 fp32; do
    OSPH_LIB=$PWD/offshore-sph_b200/lib/variants/lib_$v.so python bench.py --steps 40 --warmup 10 --no-cpu-baseline --precision $p ${BENCH_ARGS} | \
      python -c "import json,sys; d=json.load(sys.stdin); print('$v $p value %.4e ms/step %.3f pair_us %.1f'%(d['value'], d['ms_per_step'], d['roofline']['avg_launch_us']))"
  done
done
