set -x
FDC_FUSED=1 python -m pytest tests/test_gpu_chan.py -m gpu -x -q -k "chain_matches or device_path or sliding or golden" > gpurun_out/r2_pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_d.log
tail -5 gpurun_out/r2_pytest_d.log
for f in 0 1; do FDC_FUSED=$f python bench.py --workload cfg4 --no-cpu --no-e2e > gpurun_out/r2_bench_d_fused$f.json 2> gpurun_out/r2_bench_d_fused$f.err; done
FDC_FUSED=1 FDC_STREAMS=1 python bench.py --workload cfg4 --no-cpu --no-e2e > gpurun_out/r2_bench_d_fused1_s1.json 2>&1
FDC_FUSED=1 FDC_STREAMS=2 python bench.py --workload cfg4 --no-cpu --no-e2e > gpurun_out/r2_bench_d_fused1_s2.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['roofline']['path']['frac'], d['roofline']['kernels']['forward_fft']['ms'], d['roofline']['kernels']['channel_extract']['ms'], d['clocks'])
    except Exception as e: print(f, 'ERR', e, open(f).read()[-500:])
PY
