FDC_FUSED=1 python -m pytest tests/test_gpu_chan.py -m gpu -x -q -k "chain_matches or device_path" > gpurun_out/r2_pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_e.log
tail -3 gpurun_out/r2_pytest_e.log
for f in 1 2; do FDC_FUSED=$f python bench.py --workload cfg4 --no-cpu --no-e2e > gpurun_out/r2_bench_e_fused$f.json 2> gpurun_out/r2_bench_e_fused$f.err; done
FDC_FUSED=1 FDC_STREAMS=1 FDC_CLUSTER_PROF=1 python bench.py --workload cfg4 --no-cpu --no-e2e --steps 3 --warmup 3 2>&1 | grep "cluster fwd" | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_e_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['roofline']['path']['frac'], d['roofline']['kernels']['forward_fft']['ms'], d['roofline']['kernels']['channel_extract']['ms'], d['clocks'])
    except Exception as e: print(f, 'ERR', e, open(f).read()[-500:])
PY
