#!/bin/bash
# round-2 evidence on one B200: bench lines of every workload, blocks-per-call table, reference arm, ncu launch list,
# in-pipeline DRAM capture (application replay, caches untouched) and one --set full capture of the hot kernels
set -x
O=gpurun_out
python bench.py > $O/r2_bench_cfg4.json 2> $O/r2_bench_cfg4.err
for w in cfg2 cfg1 cfg5; do python bench.py --workload $w > $O/r2_bench_$w.json 2> $O/r2_bench_$w.err; done
python bench.py --workload cfg3 > $O/r2_bench_cfg3.json 2> $O/r2_bench_cfg3.err
python bench.py --workload cfg5_activity > $O/r2_bench_cfg5_activity.json 2> $O/r2_bench_cfg5_activity.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
python tools/bench_batch.py cfg4 > $O/r2_blocks_per_call.txt 2>&1
python tools/bench_batch.py cfg2 >> $O/r2_blocks_per_call.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sustained --no-readings"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_ncu_launches.csv $CMD > $O/ncu1.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
$CMD > $O/plain.log 2>&1 && ncu --metrics $M --cache-control none --clock-control none --replay-mode application -k regex:"k_fwd|k_extract" -s 54 -c 18 --csv --log-file $O/r2_ncu_warm_step.csv $CMD > $O/ncu2.log 2>&1
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_fwd|k_extract" -s 57 -c 3 -o $O/r2_prof $CMD > $O/ncu3.log 2>&1
ls -la $O | tail -20
