#!/usr/bin/env python
"""Golden vectors for the hier block's frequency modes, made by EXECUTING THE REFERENCE'S OWN SOURCE LINES:
the conversion lambdas of python/FrequencyDomainChannelizer.py:70-91, get_channel / get_segment (:349-357), nextpow2 (:37-40) and
get_opt_channelparams (:322-345) are cut out of /root/reference/python/FrequencyDomainChannelizer.py as text and exec'd here
(the module itself cannot be imported: it needs GNU Radio).  `/` on integers is Python-2 floor division in the reference; the
only such expression on this path is `blocklen/2` with blocklen a power of two >= 2, where true division gives the same value.

    python tests/golden/make_freqmode_golden.py      ->  tests/golden/freqmodes.json"""
import json
import os
import re
import textwrap

import numpy

SRC = "/root/reference/python/FrequencyDomainChannelizer.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(lines, first, last):
    return textwrap.dedent("".join(lines[first - 1:last]))


def main():
    lines = open(SRC).read().splitlines(True)
    ns = {"numpy": numpy}
    exec(cut(lines, 31, 32), ns)                      # class FREQMODE
    exec(cut(lines, 37, 40), ns)                      # nextpow2
    assert "def nextpow2" in cut(lines, 37, 40) and "self.get_freq=lambda f: (f+0.5)%1.0" in lines[69]
    lam = cut(lines, 70, 91)
    assert lam.count("lambda") == 12 and "raise ValueError('Unknown Frequency mode" in lam
    gp = cut(lines, 322, 345)
    assert gp.startswith("def get_opt_channelparams(self, freq, bw):")
    gc = cut(lines, 349, 357)
    assert gc.startswith("def get_channel(self,c):") and "def get_segment(self,c):" in gc

    class Holder(object):
        pass
    exec(gp, ns); exec(gc, ns)
    Holder.get_opt_channelparams = ns["get_opt_channelparams"]
    Holder.get_channel = ns["get_channel"]; Holder.get_segment = ns["get_segment"]

    cases = []
    fs, cf = 2.4e6, 433.92e6
    grid = {
        "normalized": (1.0, 0.0, [(-0.31, 0.05), (0.12, 0.1), (0.4999, 0.02), (-0.5, 0.03), (0.0, 0.081)], [(-0.4, 0.4), (0.05, 0.3)]),
        "basebandfs": (fs, 0.0, [(-0.31 * fs, 0.05 * fs), (2.9e5, 2.4e5), (1.19e6, 5e4), (-1.2e6, 7.2e4), (0.0, 1.9e5)], [(-9.6e5, 9.6e5), (1.2e5, 7.2e5)]),
        "centerfreqfs": (fs, cf, [(cf - 0.31 * fs, 0.05 * fs), (cf + 2.9e5, 2.4e5), (cf + 1.19e6, 5e4), (cf - 1.2e6, 7.2e4), (cf, 1.9e5)],
                         [(cf - 9.6e5, cf + 9.6e5), (cf + 1.2e5, cf + 7.2e5)]),
    }
    for mode, (fs_, cf_, chans, segs) in grid.items():
        for blocksize, relinvovl in ((4096, 4), (1024, 2), (16384, 8)):
            h = Holder()
            env = dict(ns); env.update({"self": h, "freqmode": mode, "fs": fs_, "centerfrequency": cf_})
            exec(lam, env)
            h.blocksize = blocksize; h.relinvovl = relinvovl
            conv = [h.get_channel(list(c)) for c in chans]
            params = [list(h.get_opt_channelparams(c[0], c[1])) for c in conv]
            back = [[h.set_freq(c[0]), h.set_bw(c[1])] for c in conv]
            cases.append({"freqmode": mode, "fs": fs_, "centerfrequency": cf_, "blocksize": blocksize, "relinvovl": relinvovl,
                          "channels": [list(c) for c in chans], "segments": [list(s) for s in segs],
                          "normalized_channels": conv, "normalized_segments": [h.get_segment(list(s)) for s in segs],
                          "channel_params": params, "set_freq_bw_of_normalized": back, "freqmode_enum": h.freqmode})
    out = {"source": "python/FrequencyDomainChannelizer.py lines 31-32, 37-40, 70-91, 322-345, 349-357 executed by tests/golden/make_freqmode_golden.py",
           "cases": cases}
    with open(os.path.join(HERE, "freqmodes.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
