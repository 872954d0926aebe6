#!/usr/bin/env python
"""Golden vectors for the waterfall consumer's arithmetic, made by EXECUTING THE REFERENCE'S OWN LINES:
work() (python/WaterfallMsgTagging.py:247-256: mean over blocklen/1024 bins, or repetition), apply_colorscheme() (:261-262) and
cr_colorscheme() (:276-313) are cut out of /root/reference/python/WaterfallMsgTagging.py as text and exec'd on a plain object that
carries the attributes __init__ sets (:33-45); the class itself cannot be imported (PyQt4, GNU Radio).

    python tests/golden/make_waterfall_golden.py      ->  tests/golden/waterfall.npz"""
import os
import textwrap

import numpy

SRC = "/root/reference/python/WaterfallMsgTagging.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(lines, first, last):
    return textwrap.dedent("".join(lines[first - 1:last]))


def main():
    lines = open(SRC).read().splitlines(True)
    ns = {"numpy": numpy}
    work = cut(lines, 247, 256); apply_cs = cut(lines, 261, 262); cr = cut(lines, 276, 313)
    assert work.startswith("def work(self, input_items, output_items):") and "numpy.mean( in0.reshape(" in work and "numpy.kron(in0" in work
    assert apply_cs.startswith("def apply_colorscheme(self, blocks):") and "numpy.digitize" in apply_cs
    assert cr.startswith("def cr_colorscheme(self, colorscheme):") and cr.rstrip().endswith("return colorscheme_cols, colorscheme_bins, colorscheme_frame")
    for src in (work, apply_cs, cr):
        exec(src, ns)

    class Holder(object):
        work = ns["work"]; apply_colorscheme = ns["apply_colorscheme"]; cr_colorscheme = ns["cr_colorscheme"]

    out = {}
    names = []
    rng = numpy.random.default_rng(11)
    for blocklen, loginput, scheme, lo, hi in ((4096, False, 0, -60.0, 0.0), (16384, True, 1, -80.0, -10.0), (1024, False, 2, -50.0, 5.0),
                                                (256, True, 3, -70.0, -20.0), (65536, False, 1, -90.0, -30.0)):
        h = Holder()
        # the attributes of __init__ (:33-45)
        h.blocklen = int(blocklen); h.loginput = bool(loginput); h.minvaldb = float(lo); h.maxvaldb = float(hi)
        h.normwidth = 1024
        h.block_reduction = int(h.blocklen // h.normwidth); h.block_interpolation = int(h.normwidth // h.blocklen)
        h.puffer_blocks = []; h.colorscheme = scheme
        h.colorscheme_cols, h.colorscheme_bins, h.colorscheme_frame = h.cr_colorscheme(int(scheme))
        nrows = 5
        power = (10.0 ** (rng.uniform(lo - 10.0, hi + 10.0, size=(nrows, blocklen)) / 10.0)).astype(numpy.float32)
        x = (10.0 * numpy.log10(power)).astype(numpy.float32) if loginput else power
        assert h.work([x], None) == nrows
        rows = numpy.array(h.puffer_blocks)
        pix = numpy.array([h.apply_colorscheme(r) for r in rows])
        key = "b%d_log%d_s%d" % (blocklen, int(loginput), scheme)
        names.append(key)
        out[key + "_x"] = x; out[key + "_rows"] = rows; out[key + "_pixels"] = pix
        out[key + "_cols"] = h.colorscheme_cols; out[key + "_bins"] = h.colorscheme_bins; out[key + "_frame"] = h.colorscheme_frame
        out[key + "_meta"] = numpy.array([blocklen, int(loginput), scheme, lo, hi], dtype=numpy.float64)
    out["names"] = numpy.array(names)
    numpy.savez_compressed(os.path.join(HERE, "waterfall.npz"), **out)
    print("wrote", names)


if __name__ == "__main__":
    main()
