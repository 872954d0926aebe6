#!/bin/bash
# measurement aid: worker streams x chunk size (is the K1 -> K2 hand-over L2 resident, and does that pay?)
wl=${1:-cfg4}
out=gpurun_out/sweep2_${wl}.txt
: > $out
for st in 1 2 3 4; do for c in ${SWEEP_CHUNKS:-16 24 32 48 64 128}; do
  echo -n "streams $st chunk $c: " >> $out
  FDC_STREAMS=$st python bench.py --workload $wl --no-cpu --no-e2e --steps 20 --warmup 3 --chunk $c 2>>$out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('value %.0f Ms/s  ms/step %.4f  launches/step %.0f' % (d['value'], d['ms_per_step'], d['roofline']['launches_per_step']))
" >> $out
done; done
cat $out
