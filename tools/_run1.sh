set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_a.log
tail -5 gpurun_out/r2_pytest_a.log
for w in cfg4 cfg2 cfg1 cfg5 cfg4_ovl75; do python bench.py --workload $w --no-cpu --no-e2e > gpurun_out/r2_bench_a_$w.json 2> gpurun_out/r2_bench_a_$w.err; done
FDC_NO=1 python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['roofline']['path']['frac'], d['roofline']['kernels']['forward_fft']['ms'], d['roofline']['kernels']['channel_extract']['ms'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
