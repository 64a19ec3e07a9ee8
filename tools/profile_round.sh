#!/bin/bash
# ncu evidence for profiles/ (round 2): run only after the plain commands have exited 0 (numbers printed under ncu are not bench values)
OUT=gpurun_out
mkdir -p $OUT
HEAD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$HEAD > $OUT/plain_head.log 2>&1 || exit 1
# every launch of the headline leg with its device time (shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_r2.csv $HEAD > $OUT/ncu_launch_r2.log 2>&1
# the dominant kernel, 3 consecutive launches out of the timed graph: DRAM bytes per launch, stalls
ncu --set full --clock-control none --import-source on -k regex:gather_tiles_kernel -s 40 -c 3 -o $OUT/prof_gather_alarm_r2 $HEAD > $OUT/ncu_head_r2.log 2>&1
python tools/count_one.py ktree200 25 3 > $OUT/plain_count.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:count_tiles -s 2 -c 1 -o $OUT/prof_count_ktree_r2 python tools/count_one.py ktree200 25 3 > $OUT/ncu_count_r2.log 2>&1
python tools/count_one.py alarm 26 3 >> $OUT/plain_count.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:count_tiles -s 2 -c 1 -o $OUT/prof_count_alarm_r2 python tools/count_one.py alarm 26 3 >> $OUT/ncu_count_r2.log 2>&1
python tools/exp_rows_one.py 28 > $OUT/plain_rows.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ve_rows -s 2 -c 1 -o $OUT/prof_rows28_r2 python tools/exp_rows_one.py 28 > $OUT/ncu_rows_r2.log 2>&1
python tools/exp_rows_one.py 40 > $OUT/plain_rows40.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ve_rows -s 2 -c 1 -o $OUT/prof_rows40_r2 python tools/exp_rows_one.py 40 > $OUT/ncu_rows40_r2.log 2>&1
ls -la $OUT/*.ncu-rep
