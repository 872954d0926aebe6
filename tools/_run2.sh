set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_b.log
tail -3 gpurun_out/r2_pytest_b.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_warm.log 2>&1 && ncu --metrics $M --cache-control none --clock-control none -k regex:"k_fwd|k_extract" -s 63 -c 21 --csv --log-file gpurun_out/r2_ncu_warm_step.csv $CMD > gpurun_out/ncu_warm.log 2>&1
export FDC_STREAMS=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --chunk 32"
$CMD > gpurun_out/plain_warm32.log 2>&1 && ncu --metrics $M --cache-control none --clock-control none -k regex:"k_fwd|k_extract" -s 330 -c 66 --csv --log-file gpurun_out/r2_ncu_warm_step_chunk32.csv $CMD > gpurun_out/ncu_warm32.log 2>&1
