"""FDC.WaterfallMsgTagging -- headless model of the reference's waterfall consumer (python/WaterfallMsgTagging.py).

The reference block is a PyQt4 widget: it takes float power vectors of `blocklen` bins, reduces each to 1024 columns,
averages `blockdecimation` consecutive vectors into one image line, maps the lines through a colour table and frames the
bursts announced by the PDUs of the activity-gated blocks (blockstart, blockend, rel_cfreq, rel_bw).  Here the same
constructor arguments and the same image arithmetic, without Qt: the image is the `pixels` array (rows x 3*1024 uint8, RGB,
newest line last) and `save_ppm()` writes it out.  New in this implementation: `work_spectrum()` /
`work_spectrum_device()` take the complex spectrum itself and do |X|^2 (and 10 log10 for a dB display) plus the reduction
to 1024 columns on the GPU (fdc_waterfall_*), so 1024 floats per block reach the host instead of blocklen.
"""
import ctypes as C

import numpy as np

NORMWIDTH = 1024            # python/WaterfallMsgTagging.py:43
TABLE = 1024                # entries of a colour table (:263)


def color_table(scheme, minvaldb, maxvaldb, loginput):
    """(colours[TABLE, 3] uint8, bin edges[TABLE - 1], frame colour[3]) of python/WaterfallMsgTagging.py:256-315.
    Schemes: 0 black-blue-cyan-white, 1 black-rainbow, 2 black-red-yellow, 3 black-white (green frames)."""
    edges = np.linspace(float(minvaldb), float(maxvaldb), TABLE - 1)
    if not loginput:
        edges = 10.0 ** (edges / 10.0)

    def ramp(a, b, n):
        return np.linspace(a, b, n, dtype=np.uint8)

    def flat(v, n):
        return np.full(n, v, dtype=np.uint8)
    frame = np.array([255, 255, 255], dtype=np.uint8)
    if scheme == 1:
        q = TABLE // 4
        r = np.concatenate([ramp(0, 75, q), ramp(75, 0, q), flat(0, q), ramp(0, 255, q)])
        g = np.concatenate([flat(0, q), flat(0, q), ramp(0, 255, q), flat(255, q)])
        b = np.concatenate([ramp(0, 130, q), ramp(130, 255, q), ramp(255, 0, q), flat(0, q)])
    elif scheme == 2:
        h = TABLE // 2
        r = np.concatenate([ramp(0, 255, h), flat(255, h)])
        g = np.concatenate([flat(0, h), ramp(0, 255, h)])
        b = flat(0, TABLE)
    elif scheme == 3:
        r = g = b = ramp(0, 255, TABLE)
        frame = np.array([0, 255, 0], dtype=np.uint8)
    else:
        h = TABLE // 2
        r = flat(0, TABLE)
        g = np.concatenate([flat(0, h), ramp(0, 255, h)])
        b = np.concatenate([ramp(0, 255, h), flat(255, h)])
    return np.stack([r, g, b], axis=1), edges, frame


class WaterfallMsgTagging(object):
    """WaterfallMsgTagging(blocklen, samp_rate, relinvovl, blockdecimation, loginput, minvaldb, maxvaldb, colorscheme, tagmode)
    -- python/WaterfallMsgTagging.py:32.  `height`: image lines (the widget derives it from its size, :129)."""

    def __init__(self, blocklen, samp_rate, relinvovl, blockdecimation, loginput, minvaldb, maxvaldb, colorscheme, tagmode=0, height=512):
        self.blocklen = int(blocklen)
        self.samp_rate = float(samp_rate)
        self.relinvovl = int(relinvovl)
        self.data_rate = self.samp_rate * (1.0 - 1.0 / float(relinvovl))
        self.blockdecimation = max(int(blockdecimation), 1)
        self.loginput = bool(loginput)
        self.minvaldb, self.maxvaldb, self.colorscheme, self.tagmode = float(minvaldb), float(maxvaldb), int(colorscheme), tagmode
        if self.blocklen < 1 or (self.blocklen & (self.blocklen - 1)):
            raise ValueError("blocklen must be a power of two")
        self._retable()
        self._img = np.zeros((1, NORMWIDTH, 3), dtype=np.uint8)           # one black line, like the widget before its first resize (:48)
        self.min_block, self.max_block = -1, 0                            # block index range the image covers (:65-66)
        self._rows = []                                                   # reduced vectors waiting for update()
        self._tags = []                                                   # (blockstart, blockend, left column, right column)
        self._gpu = None
        self.set_height(height)

    # ---- geometry / colours ------------------------------------------------------------------------------------------
    def _retable(self):
        self._colors, self._edges, self._frame = color_table(self.colorscheme, self.minvaldb, self.maxvaldb, self.loginput)

    def set_minvaldb(self, v):
        self.minvaldb = float(v); self._retable()

    def set_maxvaldb(self, v):
        self.maxvaldb = float(v); self._retable()

    def set_colorscheme(self, v):
        self.colorscheme = int(v); self._retable()

    def set_height(self, height):
        """What a resize does (:113-131): a new black image of `height` lines, the newest old lines kept at the bottom; the
        first block index of the window moves by the difference in lines."""
        height = max(int(height), 1)
        old = self._img
        keep = min(height, old.shape[0])
        self.min_block += (old.shape[0] - height) * self.blockdecimation
        img = np.zeros((height, NORMWIDTH, 3), dtype=np.uint8)
        img[height - keep:] = old[old.shape[0] - keep:]
        self._img, self.height = img, height

    @property
    def pixels(self):
        """rows x (3 * 1024) uint8, the reference's layout (:48)"""
        return self._img.reshape(self._img.shape[0], 3 * NORMWIDTH)

    # ---- input ---------------------------------------------------------------------------------------------------------
    def work(self, input_items, output_items=None):
        """float32 power vectors of blocklen bins (linear, or dB when loginput), as the reference block takes them (:272-279)"""
        x = np.asarray(input_items[0], dtype=np.float32).reshape(-1, self.blocklen)
        if self.blocklen > NORMWIDTH:
            red = x.reshape(x.shape[0], NORMWIDTH, self.blocklen // NORMWIDTH).mean(axis=2)
        else:
            red = np.repeat(x, NORMWIDTH // self.blocklen, axis=1)
        self._rows.extend(red)
        return x.shape[0]

    def _reducer(self):
        if self._gpu is None:
            from ._cabi import lib, handle
            self._gpu = handle(lib().fdc_waterfall_create(self.blocklen, NORMWIDTH, 1 if self.loginput else 0), "WaterfallMsgTagging")
        return self._gpu

    def work_spectrum(self, spectrum):
        """complex spectrum rows (host): |X|^2, 10 log10 when loginput, reduction to 1024 columns -- all on the GPU"""
        from ._cabi import lib, check
        x = np.ascontiguousarray(spectrum, dtype=np.complex64).reshape(-1, self.blocklen)
        out = np.empty((x.shape[0], NORMWIDTH), dtype=np.float32)
        check(lib().fdc_waterfall_work_host(self._reducer(), x.shape[0], x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)), "WaterfallMsgTagging")
        self._rows.extend(out)
        return x.shape[0]

    def work_spectrum_device(self, nblocks, d_rows, stream=0):
        """the same from spectrum rows that are already in device memory (the channelizer's debug spectrum)"""
        from ._cabi import lib, check
        out = np.empty((int(nblocks), NORMWIDTH), dtype=np.float32)
        check(lib().fdc_waterfall_work_device(self._reducer(), int(nblocks), C.c_void_p(d_rows), out.ctypes.data_as(C.c_void_p),
                                              C.c_void_p(stream) if stream else None), "WaterfallMsgTagging")
        self._rows.extend(out)
        return int(nblocks)

    def __del__(self):
        try:
            from . import _cabi
            L = _cabi.loaded()
            if L is not None and self._gpu:
                L.fdc_waterfall_destroy(self._gpu); self._gpu = None
        except Exception:                       # interpreter shutdown
            pass

    def msg_handler(self, m):
        """A PDU's metadata (dict with blockstart, blockend, rel_cfreq, rel_bw -- what messages() of the activity-gated blocks
        returns); incomplete ones are ignored (:85-110)."""
        meta = m[0] if isinstance(m, tuple) else m
        if not isinstance(meta, dict):
            return
        b0, b1 = meta.get("blockstart", -1024), meta.get("blockend", -1024)
        cf, bw = meta.get("rel_cfreq", -1.0), meta.get("rel_bw", -1.0)
        if b0 == -1024 or b1 == -1024 or cf < 0.0 or bw < 0.0:
            return
        self._tags.append((int(b0), int(b1), int(NORMWIDTH * (cf - bw / 2.0)), int(np.ceil(NORMWIDTH * (cf + bw / 2.0)))))

    # ---- image ---------------------------------------------------------------------------------------------------------
    def update(self):
        """What a repaint does (:152-196): average groups of blockdecimation buffered vectors into lines, scroll them in at the
        bottom, then frame every announced burst that is (partly) inside the window."""
        dec = self.blockdecimation
        n = len(self._rows) - len(self._rows) % dec
        if len(self._rows) < dec:
            return 0
        self.min_block += n; self.max_block += n
        lines = np.asarray(self._rows[:n]).reshape(n // dec, dec, NORMWIDTH).mean(axis=1)
        del self._rows[:n]
        colored = self._colors[np.digitize(lines, self._edges, False)]
        self._img = np.concatenate([self._img[lines.shape[0]:], colored], axis=0)
        keep = []
        for tag in reversed(self._tags):                  # newest first, like the reference's reversed index loop
            b0, b1, left, right = tag
            if b1 <= self.min_block:
                continue                                  # scrolled out
            if b0 >= self.max_block:
                keep.append(tag); continue                # still ahead
            if b1 < self.max_block and b0 > self.min_block:
                self._rect(b0, b1, left, right)
            elif b0 <= self.min_block:
                self._hline(b1, left, right); self._vline(b1, left, right, True)
            else:                                         # the end has not arrived yet: keep the tag
                self._hline(b0, left, right); self._vline(b1, left, right, False)
                keep.append(tag)
        self._tags = list(reversed(keep))
        return lines.shape[0]

    def _line_of(self, block, ceil=False):
        d = float(self.max_block - block) / self.blockdecimation
        return int(np.ceil(d)) if ceil else int(d)

    def _rect(self, b0, b1, left, right):                 # :199-212
        top = self.height - self._line_of(b0, True)
        bottom = self.height - self._line_of(b1)
        if bottom == self.height:
            bottom -= 1
        px = self.pixels
        px[top:bottom, 3 * left:3 * left + 3] = self._frame
        px[top:bottom, 3 * right:3 * right + 3] = self._frame
        px[top, 3 * left:3 * right] = np.tile(self._frame, right - left)
        px[bottom, 3 * left:3 * right] = np.tile(self._frame, right - left)

    def _hline(self, block, left, right):                 # :214-220
        line = self.height - max(self._line_of(block), 1)
        self.pixels[line, 3 * left:3 * right] = np.tile(self._frame, right - left)

    def _vline(self, block, left, right, up, length=4):   # :222-244
        line = self.height - self._line_of(block)
        px = self.pixels
        if up:
            length = min(length, line)
            if length <= 0:
                return
            rows = slice(line - length, line)
        else:
            length = min(length, px.shape[0] - line)
            if length <= 0:
                return
            rows = slice(line, line + length)
        px[rows, 3 * left:3 * left + 3] = self._frame
        px[rows, 3 * right:3 * right + 3] = self._frame

    def save_ppm(self, path):
        """the image as a binary PPM (P6)"""
        img = self._img
        with open(path, "wb") as fh:
            fh.write(("P6\n%d %d\n255\n" % (img.shape[1], img.shape[0])).encode())
            fh.write(np.ascontiguousarray(img).tobytes())
