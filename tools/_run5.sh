for st in 2 3 4; do for ch in 16 24 32 48 64 128; do
 echo -n "streams $st chunk $ch: "; FDC_STREAMS=$st python bench.py --workload cfg4 --no-cpu --no-e2e --chunk $ch --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'Ms/s', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
