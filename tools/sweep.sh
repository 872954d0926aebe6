#!/bin/bash
# measurement aid: bench.py over the run-time switches (FDC_PREFETCH, FDC_STREAMS, FDC_CTAS_PER_SM) and chunk sizes
wl=${1:-cfg4}
out=gpurun_out/sweep_${wl}.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python bench.py --workload $wl --no-cpu --no-e2e --steps 20 --warmup 3 $CHUNK 2>>$out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); k=d['roofline']['kernels']
    print('value %.0f Ms/s  ms/step %.4f  fwd %.3f ms  ext %.3f ms  path_frac %.3f  chunk %s  launches/step %.0f' % (d['value'], d['ms_per_step'], k['forward_fft']['ms'], k['channel_extract']['ms'], d['roofline']['path']['frac'], d['config']['chunk_blocks'], d['roofline']['launches_per_step']))
" >> $out
}
CHUNK=""
run FDC_PREFETCH=1 FDC_STREAMS=2
run FDC_PREFETCH=0 FDC_STREAMS=2
run FDC_PREFETCH=1 FDC_STREAMS=1
run FDC_PREFETCH=0 FDC_STREAMS=1
run FDC_PREFETCH=1 FDC_STREAMS=2 FDC_CTAS_PER_SM=1
for c in $SWEEP_CHUNKS; do CHUNK="--chunk $c"; run FDC_PREFETCH=1 FDC_STREAMS=2; run FDC_PREFETCH=1 FDC_STREAMS=1; done
cat $out
