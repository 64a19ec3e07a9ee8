#!/bin/bash
# ncu evidence for profiles/: run only after the plain commands have exited 0 (numbers printed under ncu are not bench values)
set -x
OUT=gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/plain_r1e.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r1e.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launch_r1e.log 2>&1
python tools/bench_kernels.py gather --ncu-friendly > $OUT/plain_bk_gather.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gather_inter_kernel --launch-skip 4 -c 1 -o /tmp/p_gather_asia python tools/bench_kernels.py gather --ncu-friendly > $OUT/ncu_g1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_tiles_kernel --launch-skip 2 -c 1 -o /tmp/p_gather_alarm python tools/bench_kernels.py gather --ncu-friendly > $OUT/ncu_g2.log 2>&1
python tools/bench_kernels.py count --ncu-friendly > $OUT/plain_bk_count.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:count_tiles_kernel -c 15 -o /tmp/p_count python tools/bench_kernels.py count --ncu-friendly > $OUT/ncu_c1.log 2>&1
for f in p_gather_asia p_gather_alarm p_count; do python tools/ncu_summary.py /tmp/$f.ncu-rep > $OUT/sum_$f.txt 2>&1; done
cp /tmp/p_gather_asia.ncu-rep /tmp/p_gather_alarm.ncu-rep $OUT/ 2>/dev/null
ls -la /tmp/*.ncu-rep $OUT
