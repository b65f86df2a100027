#!/bin/bash
# Hardware gate of HEAD (one B200):   gpurun --timeout 1500 -- 'bash tools/gpu_gate.sh [stage ...]'
# Stages (default: tests sanitize bench):
#   tests     pytest -m gpu (no -x: every failure is listed), smoke()
#   memcheck  the memcheck runs of `sanitize` without the racecheck run
#   sanitize  compute-sanitizer memcheck on the bench workload (1 M particles, both precisions, 2 steps) and on one golden
#             case; racecheck on the golden case ("race detection")
#   sanitize_multi  (needs 2 GPUs) memcheck over tests/multi_gpu_check.py: the three slab sequencers incl. the slab cadence
#   bench     bench.py in both precisions (default K / W) and the reference arm
#   launches  ncu launch list of one short bench run (shares of the step, not absolutes)
#   ncu       ncu --set full of k_pair (both precisions) and of the streaming kernels
# Output goes to gpurun_out/ (scratch); copy what should be judged into profiles/rNN/.
set -u
OUT=gpurun_out
mkdir -p $OUT
STAGES="${*:-tests sanitize bench}"
LOG=$OUT/gate.log
: > $LOG
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has tests; then
  { echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
    echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; } >> $LOG 2>&1
fi
if has sanitize || has memcheck; then
  { for p in fp64 fp32; do
      echo "== memcheck bench $p"
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --precision $p 2>&1 | tail -12
      echo "rc=$?"
    done
    echo "== memcheck golden cases"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python tools/sanitize_small.py 2>&1 | tail -8
    if has sanitize; then echo "== racecheck golden case"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 python tools/sanitize_small.py 2>&1 | tail -8; fi
  } >> $LOG 2>&1
fi
if has sanitize_multi; then
  { echo "== memcheck multi-GPU check"
    timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 99 python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/multi_gpu_check.py --steps 12 2>&1 | grep -E "OK|FAIL|ERROR SUMMARY|Invalid" | tail -12
  } >> $LOG 2>&1
fi
if has bench; then
  { echo "== bench fp64"; timeout 600 python bench.py | tee $OUT/bench_fp64.json | cut -c1-3000
    echo "== bench fp32"; timeout 600 python bench.py --precision fp32 --no-cpu-baseline | tee $OUT/bench_fp32.json | cut -c1-3000
    echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tee $OUT/bench_ref.json | cut -c1-1500
  } >> $LOG 2>&1
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/launches_bench.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_fp32.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --precision fp32 > $OUT/launches_bench32.log 2>&1
fi
if has ncu; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 4 -c 1 -o $OUT/pair_fp64 -f \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 4 -c 1 -o $OUT/pair_fp32 -f \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --precision fp32 > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_prepare|k_bin|k_gather|k_correct|k_grid|k_timestep' -s 14 -c 16 \
      -o $OUT/stream_fp64 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
fi
tail -120 $LOG
