#!/bin/bash
# measurement aid: extract kernel variants (16 vs 8 points per thread) with / without prefetch
out=gpurun_out/sweep3.txt
: > $out
for wl in ${SWEEP_WL:-cfg4 cfg2 cfg1}; do for e8 in 0 1; do for pf in 1 0; do
  echo -n "$wl E8=$e8 PF=$pf: " >> $out
  FDC_EXTRACT_E8=$e8 FDC_PREFETCH=$pf python bench.py --workload $wl --no-cpu --no-e2e --steps 20 --warmup 3 2>>$out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); k=d['roofline']['kernels']
    print('value %.0f Ms/s  ms/step %.4f  fwd %.3f ms  ext %.3f ms  path_frac %.3f' % (d['value'], d['ms_per_step'], k['forward_fft']['ms'], k['channel_extract']['ms'], d['roofline']['path']['frac']))
" >> $out
done; done; done
cat $out
