#!/bin/bash
# First GPU call of a round (one B200):   gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh'
# Everything the round needs to start from measured ground: GPU tests, bench lines in both precisions, the launch list of
# one bench run, full ncu captures of the kernels changed since the last measurement, the MUFU refinement accuracy.
# Output goes to gpurun_out/ (scratch); copy what should be judged into profiles/rNN/.
set -u
OUT=gpurun_out
mkdir -p $OUT
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench fp64"; timeout 600 python bench.py --steps 100 --warmup 10 | tee $OUT/bench_fp64.json | cut -c1-600
echo "== bench fp32"; timeout 600 python bench.py --steps 100 --warmup 10 --precision fp32 --no-cpu-baseline | tee $OUT/bench_fp32.json | cut -c1-600
echo "== mufu accuracy"; nvcc -arch=sm_100a tools/mufu_accuracy.cu -o /tmp/mufu && /tmp/mufu
} > $OUT/round_start.log 2>&1
# launch list of one short bench run (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
# full captures: pair kernel (both precisions) and the streaming kernels of one fused step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 4 -c 1 -o $OUT/pair_fp64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 4 -c 1 -o $OUT/pair_fp32 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision fp32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_prepare|k_keys|k_sort|k_gather|k_correct' -s 20 -c 14 \
    -o $OUT/stream_fp64 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# compile-time variants built beforehand with tools/build_round2_variants.sh (the .so files travel with gpurun)
if ls offshore-sph_b200/lib/variants/lib_*.so > /dev/null 2>&1; then
  names=$(ls offshore-sph_b200/lib/variants/lib_*.so | sed -E 's/.*lib_(.*)\.so/\1/')
  { echo "== variants"; bash tools/bench_variants.sh $names; } >> $OUT/round_start.log 2>&1
fi
tail -60 $OUT/round_start.log
