#!/bin/bash
# One B200: the other BASELINE configs as single bench lines with the final build (parity of these set-ups is covered by the
# golden cases tank30_cubic_dynh / tank16_gaussian / tank24_wendland_coupled and the oracle comparisons of tests/).
OUT=gpurun_out; mkdir -p $OUT
A="--no-cpu-baseline --no-e2e"
timeout 45 python bench.py $A --workload containment --particles-per-side 2000 --precision fp32 --steps 30 --warmup 5 > $OUT/bench_containment_4M_fp32.json 2>$OUT/cfg_err.log
timeout 50 python bench.py $A --workload icebreak --particles-per-side 4900 --kernel gaussian --steps 10 --warmup 3 > $OUT/bench_icebreak_4M_gaussian.json 2>>$OUT/cfg_err.log
timeout 60 python bench.py $A --particles-per-side 4000 --steps 20 --warmup 5 > $OUT/bench_1gpu_16M_fp64_pipelined.json 2>>$OUT/cfg_err.log
for f in bench_containment_4M_fp32 bench_icebreak_4M_gaussian bench_1gpu_16M_fp64_pipelined; do
  python -c "import json,sys; d=json.load(open('$OUT/$f.json')); print('$f', d['config']['workload'], 'value %.4e ms/step %.4f pair_us %.1f status %s'%(d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['status_bits']))" 2>&1 | tail -1
done
tail -3 $OUT/cfg_err.log
