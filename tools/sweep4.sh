#!/bin/bash
# measurement aid: co-residency of the bandwidth-bound forward kernels and the issue-bound extract kernel
out=gpurun_out/sweep4.txt
: > $out
run() {
  echo -n "$*: " >> $out
  env "$@" python bench.py --workload cfg4 --no-cpu --no-e2e --steps 20 --warmup 3 2>>$out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); k=d['roofline']['kernels']
    print('value %.0f Ms/s  ms/step %.4f  fwd %.3f ms  ext %.3f ms' % (d['value'], d['ms_per_step'], k['forward_fft']['ms'], k['channel_extract']['ms']))
" >> $out
}
run FDC_PREFETCH=1 FDC_STREAMS=2
run FDC_PREFETCH=3 FDC_STREAMS=2
run FDC_PREFETCH=1 FDC_STREAMS=3
run FDC_PREFETCH=1 FDC_STREAMS=2 FDC_CTAS_FWD=1
run FDC_PREFETCH=1 FDC_STREAMS=2 FDC_CTAS_FWD=1 FDC_CTAS_EXT=2
run FDC_PREFETCH=1 FDC_STREAMS=2 FDC_CTAS_FWD=1 FDC_CTAS_EXT=1
run FDC_PREFETCH=1 FDC_STREAMS=3 FDC_CTAS_FWD=1 FDC_CTAS_EXT=1
run FDC_PREFETCH=1 FDC_STREAMS=4 FDC_CTAS_FWD=1 FDC_CTAS_EXT=1
run FDC_PREFETCH=1 FDC_STREAMS=2 FDC_CTAS_EXT=2
run FDC_PREFETCH=1 FDC_STREAMS=3 FDC_CTAS_EXT=2
run FDC_PREFETCH=3 FDC_STREAMS=3 FDC_CTAS_FWD=1 FDC_CTAS_EXT=1
cat $out
