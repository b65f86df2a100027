#!/bin/bash
# One B200:  gpurun --timeout 600 -- 'bash tools/gpu_uh_variants3.sh'
# Pipelined flush loop, third round: unroll 2 vs 3 on the uniform-h instantiation; the general instantiation pipelined too
# (gp2, gp1 = unroll 1), measured on the dam break with OSPH_UH=0 (forces the general instantiation) and on the containment
# tank (dynamic h).  The whole GPU suite runs on gp2 first.
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/uh_variants3.log
: > $LOG
V=$PWD/offshore-sph_b200/lib/variants
one() {  # name env args...
  local name=$1; shift
  "$@" | python -c "import json,sys; d=json.load(sys.stdin); print('$name value %.4e ms/step %.4f pair_us %.1f status %s'%(d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['status_bits']))"
}
{ echo "== pytest -m gpu on lib_gp2.so"
  OSPH_LIB=$V/lib_gp2.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
  echo "== uniform-h parity on lib_uhp3.so / lib_gp1.so"
  for v in uhp3 gp1; do OSPH_LIB=$V/lib_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "uniform_h or dam_break_vs_oracle or whole_steps" 2>&1 | tail -2; done
  echo "== bench: dam break (uniform-h instantiation)"
  BENCH_ARGS="--no-e2e" bash tools/bench_variants.sh uhp2 uhp3
  echo "== bench: dam break with OSPH_UH=0 (general instantiation; default build 294.8 / 172.5 us)"
  for v in uhp2 gp2 gp1; do for p in fp64 fp32; do
    one "$v $p general" env OSPH_UH=0 OSPH_LIB=$V/lib_$v.so python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e --precision $p
  done; done
  echo "== bench: containment tank, dynamic h, 1 M particles"
  for v in uhp2 gp2; do for p in fp64 fp32; do
    one "$v $p containment" env OSPH_LIB=$V/lib_$v.so python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e --precision $p --workload containment --particles-per-side 1000
  done; done
} >> $LOG 2>&1
cat $LOG
