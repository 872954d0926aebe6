"""Oracle (TEST INFRASTRUCTURE ONLY): restatement of the image arithmetic of the reference's waterfall consumer,
python/WaterfallMsgTagging.py, as plain functions on a state dict, pixel by pixel where the reference uses array tricks.

PARITY: the reference class needs PyQt4 and GNU Radio, neither of which can be imported in this container, and the reference
ships no test or fixture for it.  Its ARITHMETIC is pinned: reduce_vectors and color_tables (and the colour mapping) are checked
against tests/golden/waterfall.npz, which tests/golden/make_waterfall_golden.py makes by executing the reference's own lines
(work :247-256, apply_colorscheme :261-262, cr_colorscheme :276-313).  The widget-side logic (resize, scrolling repaint, tag
frames) is restated from the source and stays UNPINNED.  What is restated, with the lines it follows:
  reduce_vectors   work()              :272-279   mean over blocklen/1024 bins, or repetition when blocklen < 1024
  color_tables     cr_colorscheme()    :256-315
  resize           renew_pixmap()      :113-126
  repaint          pxupdate()          :152-196   block decimation, scrolling, tag evaluation
  rect/hline/vline draw_*()            :199-244
  add_tag          msg_handler()       :85-110
Only tests/ may import this module."""
import math

import numpy as np

W = 1024


def reduce_vectors(x, blocklen):
    x = np.asarray(x, dtype=np.float32).reshape(-1, blocklen)
    out = np.zeros((x.shape[0], W), dtype=np.float64)
    if blocklen > W:
        red = blocklen // W
        for c in range(W):
            out[:, c] = x[:, c * red:(c + 1) * red].astype(np.float64).sum(axis=1) / red
    else:
        rep = W // blocklen
        for c in range(W):
            out[:, c] = x[:, c // rep]
    return out


def color_tables(scheme, minvaldb, maxvaldb, loginput):
    n = 1024
    edges = [float(e) for e in np.linspace(minvaldb, maxvaldb, n - 1)]      # the reference calls numpy.linspace itself (:266)
    if not loginput:
        edges = [10.0 ** (e / 10.0) for e in edges]

    def ramp(a, b, m):          # numpy.linspace(a, b, m, dtype=uint8) (:270): computed in double, truncated
        return [int(v) for v in np.linspace(a, b, m)]
    frame = (255, 255, 255)
    if scheme == 1:
        q = n // 4
        r = ramp(0, 75, q) + ramp(75, 0, q) + [0] * q + ramp(0, 255, q)
        g = [0] * q + [0] * q + ramp(0, 255, q) + [255] * q
        b = ramp(0, 130, q) + ramp(130, 255, q) + ramp(255, 0, q) + [0] * q
    elif scheme == 2:
        h = n // 2
        r = ramp(0, 255, h) + [255] * h
        g = [0] * h + ramp(0, 255, h)
        b = [0] * n
    elif scheme == 3:
        r = g = b = ramp(0, 255, n)
        frame = (0, 255, 0)
    else:
        h = n // 2
        r = [0] * n
        g = [0] * h + ramp(0, 255, h)
        b = ramp(0, 255, h) + [255] * h
    return np.array([r, g, b], dtype=np.uint8).T.copy(), np.array(edges), np.array(frame, dtype=np.uint8)


def new_state(blockdecimation, scheme, minvaldb, maxvaldb, loginput):
    cols, edges, frame = color_tables(scheme, minvaldb, maxvaldb, loginput)
    return {"dec": max(int(blockdecimation), 1), "cols": cols, "edges": edges, "frame": frame,
            "px": np.zeros((1, 3 * W), dtype=np.uint8), "min_block": -1, "max_block": 0, "rows": [], "tags": [], "h": 1}


def resize(st, height):
    old = st["px"]
    px = np.zeros((height, 3 * W), dtype=np.uint8)
    keep = min(height, old.shape[0])
    st["min_block"] += (old.shape[0] - height) * st["dec"]
    px[height - keep:] = old[old.shape[0] - keep:]
    st["px"], st["h"] = px, height


def add_tag(st, blockstart, blockend, rel_cfreq, rel_bw):
    if blockstart == -1024 or blockend == -1024 or rel_cfreq < 0.0 or rel_bw < 0.0:
        return
    st["tags"].append((blockstart, blockend, int(W * (rel_cfreq - rel_bw / 2.0)), int(math.ceil(W * (rel_cfreq + rel_bw / 2.0)))))


def _set(px, r, c, frame):
    if 0 <= c < W and -px.shape[0] <= r < px.shape[0]:
        px[r, 3 * c:3 * c + 3] = frame


def _rect(st, b0, b1, left, right):
    h, dec, px = st["h"], st["dec"], st["px"]
    begin = h - int(math.ceil(float(st["max_block"] - b0) / dec))
    end = h - int(float(st["max_block"] - b1) / dec)
    if end == h:
        end -= 1
    for r in range(begin, end):
        _set(px, r, left, st["frame"]); _set(px, r, right, st["frame"])
    for c in range(left, right):
        _set(px, begin, c, st["frame"]); _set(px, end, c, st["frame"])


def _hline(st, block, left, right):
    line = st["h"] - max(int(float(st["max_block"] - block) / st["dec"]), 1)
    for c in range(left, right):
        _set(st["px"], line, c, st["frame"])


def _vline(st, block, left, right, up, length=4):
    px = st["px"]
    line = st["h"] - int(float(st["max_block"] - block) / st["dec"])
    if up:
        if line < length:
            length = line
        rows = range(line - length, line)
    else:
        if px.shape[0] - line < length:
            length = px.shape[0] - line
        rows = range(line, line + length)
    if length <= 0:
        return
    for r in rows:
        _set(px, r, left, st["frame"]); _set(px, r, right, st["frame"])


def repaint(st):
    dec = st["dec"]
    if len(st["rows"]) < dec:
        return
    n = len(st["rows"]) - len(st["rows"]) % dec
    st["min_block"] += n; st["max_block"] += n
    lines = []
    for g in range(n // dec):
        acc = np.zeros(W, dtype=np.float64)
        for k in range(dec):
            acc += st["rows"][g * dec + k]
        lines.append(acc / dec)
    del st["rows"][:n]
    new = np.zeros((len(lines), 3 * W), dtype=np.uint8)
    for i, ln in enumerate(lines):
        idx = np.searchsorted(st["edges"], ln, side="right")          # numpy.digitize(x, bins, right=False) for increasing bins
        new[i] = st["cols"][idx].reshape(-1)
    st["px"] = np.concatenate([st["px"][len(lines):], new], axis=0)
    i = len(st["tags"]) - 1
    while i >= 0:
        b0, b1, left, right = st["tags"][i]
        if b1 <= st["min_block"]:
            del st["tags"][i]
        elif b0 >= st["max_block"]:
            pass
        elif b1 < st["max_block"] and b0 > st["min_block"]:
            _rect(st, b0, b1, left, right); del st["tags"][i]
        elif b0 <= st["min_block"]:
            _hline(st, b1, left, right); _vline(st, b1, left, right, True); del st["tags"][i]
        else:
            _hline(st, b0, left, right); _vline(st, b1, left, right, False)
        i -= 1
