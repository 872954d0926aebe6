#!/usr/bin/env python
"""measurement aid: top stall instructions of one kernel from `ncu -i rep --page source --csv --kernel-name regex:...`"""
import csv
import subprocess
import sys


def main(rep, kernel, top=30):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    out = []
    hdr = None
    for r in rows:
        if r and r[0] == 'Address':
            if hdr is not None:
                break               # only the first launch
            hdr = r; isrc = hdr.index('Source'); ism = hdr.index('Warp Stall Sampling (All Samples)'); continue
        if hdr is None or len(r) <= ism:
            continue
        try:
            out.append((int(r[ism] or 0), len(out), r[isrc].strip()))
        except ValueError:
            pass
    tot = sum(d[0] for d in out) or 1
    print('total samples', tot, 'instructions', len(out))
    for s, i, src in sorted(out, reverse=True)[:top]:
        print('%6d %5.1f%%  #%4d  %s' % (s, 100.0 * s / tot, i, src[:110]))
    # coarse profile along the program: samples per 50 instructions
    print('samples per 50-instruction window:')
    for a in range(0, len(out), 50):
        print('  #%4d-%4d %5.1f%%' % (a, min(a + 49, len(out) - 1), 100.0 * sum(d[0] for d in out[a:a + 50]) / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
