"""The waterfall consumer: FDC.WaterfallMsgTagging (headless model of python/WaterfallMsgTagging.py) against
the pixel-by-pixel restatement in oracle/waterfall_numpy.py, and the GPU reduction of spectrum rows to 1024 columns against
NumPy.  The widget itself needs PyQt4 + GNU Radio and cannot run here; its ARITHMETIC (row reduction, colour tables, colour mapping:
python/WaterfallMsgTagging.py:247-256, 261-262, 276-313) is pinned by tests/golden/waterfall.npz, made by executing those lines of the
reference (tests/golden/make_waterfall_golden.py).  The scrolling / tag-frame drawing of the restatement stays unpinned."""
import os
import numpy as np
import pytest



def _drive(FDC, wf_or, blocklen, dec, scheme, loginput, height, chunks, rng, use_power=True):
    lo, hi = (-60.0, 0.0)
    model = FDC.WaterfallMsgTagging(blocklen, 1e6, 4, dec, loginput, lo, hi, scheme, 0, height=height)
    st = wf_or.new_state(dec, scheme, lo, hi, loginput)
    wf_or.resize(st, height)
    assert (model.min_block, model.max_block) == (st["min_block"], st["max_block"])
    blk = 0
    for n in chunks:
        p = rng.random((n, blocklen)).astype(np.float32) ** 4
        p[:, blocklen // 4:blocklen // 4 + max(blocklen // 16, 1)] += 0.5
        x = (10.0 * np.log10(p + 1e-9)).astype(np.float32) if loginput else p
        model.work([x])
        st["rows"].extend(wf_or.reduce_vectors(x, blocklen))
        # bursts: one inside the window, one that started before it, one whose end is still to come, one incomplete
        for (b0, b1, cf, bw) in ((blk + 1, blk + max(n // 2, 2), 0.30, 0.05), (blk - 3 * height * dec, blk + 2, 0.55, 0.02),
                                 (blk + n - 2, blk + n + 5 * dec, 0.75, 0.10), (-1024, blk, 0.5, 0.1)):
            m = {"blockstart": b0, "blockend": b1, "rel_cfreq": cf, "rel_bw": bw}
            model.msg_handler(m); wf_or.add_tag(st, b0, b1, cf, bw)
        model.update(); wf_or.repaint(st)
        blk += n
        assert (model.min_block, model.max_block) == (st["min_block"], st["max_block"])
        assert model.pixels.shape == st["px"].shape
        diff = np.count_nonzero(np.any(model.pixels.reshape(-1, 3) != st["px"].reshape(-1, 3), axis=1))
        # fp32 vs fp64 means can land on different sides of a colour-bin edge for a few pixels; frames must agree exactly
        assert diff <= 2e-3 * model.pixels.size / 3, diff
        frame = np.all(st["px"].reshape(-1, 3) == st["frame"], axis=1)
        assert np.array_equal(frame, np.all(model.pixels.reshape(-1, 3) == st["frame"], axis=1)) or scheme != 3
        assert sorted(model._tags) == sorted(st["tags"])
    return model, st


@pytest.mark.parametrize("blocklen,dec,scheme,loginput", [(4096, 1, 0, False), (1024, 3, 1, True), (256, 2, 2, False), (16384, 4, 3, True)])
def test_model_against_restatement(blocklen, dec, scheme, loginput):
    import FDC
    from oracle import waterfall_numpy as wf_or
    rng = np.random.default_rng(blocklen + dec)
    model, st = _drive(FDC, wf_or, blocklen, dec, scheme, loginput, 48, (7, 1, 30, 64, 5), rng)
    assert np.count_nonzero(np.all(model.pixels.reshape(-1, 3) == st["frame"], axis=1)) > 20      # frames were drawn


def test_color_tables_and_resize(tmp_path):
    import FDC
    from FDC import waterfall
    from oracle import waterfall_numpy as wf_or
    for scheme in range(4):
        for log in (False, True):
            a = waterfall.color_table(scheme, -80.0, -10.0, log); b = wf_or.color_tables(scheme, -80.0, -10.0, log)
            assert np.array_equal(a[0], b[0]) and np.allclose(a[1], b[1], rtol=1e-15) and np.array_equal(a[2], b[2])
    m = FDC.WaterfallMsgTagging(2048, 1e6, 4, 2, False, -60, 0, 0, 0, height=10)
    st = wf_or.new_state(2, 0, -60, 0, False); wf_or.resize(st, 10)
    x = np.random.default_rng(1).random((20, 2048)).astype(np.float32)
    m.work([x]); st["rows"].extend(wf_or.reduce_vectors(x, 2048)); m.update(); wf_or.repaint(st)
    for h in (25, 4):
        m.set_height(h); wf_or.resize(st, h)
        assert (m.min_block, m.max_block) == (st["min_block"], st["max_block"]) and m.pixels.shape == st["px"].shape
    m.save_ppm(str(tmp_path / "w.ppm"))
    assert open(str(tmp_path / "w.ppm"), "rb").read(15).startswith(b"P6\n1024 4\n255\n")


@pytest.mark.gpu
@pytest.mark.parametrize("blocklen,loginput", [(65536, False), (16384, True), (1024, False), (256, True), (4096, False)])
def test_gpu_reduction_of_spectrum_rows(blocklen, loginput):
    """|X|^2 (+ 10 log10) + the reduction to 1024 columns on the device == NumPy on the host, and the images agree"""
    import FDC
    import torch
    rng = np.random.default_rng(blocklen)
    nb = 9
    x = ((rng.standard_normal((nb, blocklen)) + 1j * rng.standard_normal((nb, blocklen))) * 0.05).astype(np.complex64)
    x[:, blocklen // 3:blocklen // 3 + max(blocklen // 20, 2)] *= 30.0
    pw = (x.real.astype(np.float64) ** 2 + x.imag.astype(np.float64) ** 2)
    pw = 10.0 * np.log10(pw) if loginput else pw
    a = FDC.WaterfallMsgTagging(blocklen, 1e6, 4, 1, loginput, -60, 10, 1, 0, height=16)
    b = FDC.WaterfallMsgTagging(blocklen, 1e6, 4, 1, loginput, -60, 10, 1, 0, height=16)
    c = FDC.WaterfallMsgTagging(blocklen, 1e6, 4, 1, loginput, -60, 10, 1, 0, height=16)
    a.work([pw.astype(np.float32)]); b.work_spectrum(x)
    d = torch.from_numpy(x.view(np.float32)).cuda()
    c.work_spectrum_device(nb, d.data_ptr())
    ra, rb, rc = np.asarray(a._rows), np.asarray(b._rows), np.asarray(c._rows)
    assert np.array_equal(rb.view(np.uint32), rc.view(np.uint32))
    tol = 2e-5 if loginput else 2e-6
    assert np.max(np.abs(rb - ra) / np.maximum(np.abs(ra), 1e-3)) < tol
    for w in (a, b):
        w.msg_handler({"blockstart": 2, "blockend": 6, "rel_cfreq": 0.34, "rel_bw": 0.06}); w.update()
    diff = np.count_nonzero(np.any(a.pixels.reshape(-1, 3) != b.pixels.reshape(-1, 3), axis=1))
    assert diff <= 2e-3 * a.pixels.size / 3


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waterfall.npz")


def test_arithmetic_against_the_reference_lines():
    """row reduction, colour tables and colour mapping of the restatement AND of the FDC model against what the reference's own
    lines return (golden file): rows to fp32 rounding, colour tables exactly, pixels equal except where a value sits within
    rounding of a colour-bin edge"""
    import FDC
    from FDC import waterfall
    from oracle import waterfall_numpy as wf_or
    g = np.load(GOLD)
    for key in [str(k) for k in g["names"]]:
        blocklen, loginput, scheme, lo, hi = g[key + "_meta"]
        blocklen, loginput, scheme = int(blocklen), bool(loginput), int(scheme)
        x, rows, pix = g[key + "_x"], g[key + "_rows"], g[key + "_pixels"]
        for cols, edges, frame in (wf_or.color_tables(scheme, lo, hi, loginput), waterfall.color_table(scheme, lo, hi, loginput)):
            assert np.array_equal(cols, g[key + "_cols"]) and np.array_equal(frame, g[key + "_frame"])
            assert np.allclose(edges, g[key + "_bins"], rtol=1e-14, atol=0)
        mine = wf_or.reduce_vectors(x, blocklen)
        assert mine.shape == rows.shape and np.allclose(mine, rows, rtol=2e-6, atol=1e-12)
        model = FDC.WaterfallMsgTagging(blocklen, 1e6, 4, 1, loginput, lo, hi, scheme, 0, height=x.shape[0])
        model.work([x]); model.update()
        got = np.asarray(model._rows) if len(model._rows) else np.asarray(model.pixels)
        assert model.pixels.shape == pix.shape
        diff = np.count_nonzero(np.any(model.pixels.reshape(-1, 3) != pix.reshape(-1, 3), axis=1))
        assert diff <= 2e-3 * pix.size / 3, (key, diff)


@pytest.mark.gpu
def test_gpu_rows_against_the_reference_lines():
    """the device reduction (|X|^2, 10 log10, mean over blocklen / 1024 bins) fed with spectra whose powers are the golden inputs:
    rows equal to the reference lines' rows to fp32 rounding"""
    import FDC
    import torch
    g = np.load(GOLD)
    for key in [str(k) for k in g["names"]]:
        blocklen, loginput, scheme, lo, hi = g[key + "_meta"]
        blocklen, loginput = int(blocklen), bool(loginput)
        x, rows = g[key + "_x"], g[key + "_rows"]
        power = 10.0 ** (x.astype(np.float64) / 10.0) if loginput else x.astype(np.float64)
        spec = (np.sqrt(power) * np.exp(2j * np.pi * np.random.default_rng(3).random(power.shape))).astype(np.complex64)
        c = FDC.WaterfallMsgTagging(blocklen, 1e6, 4, 1, loginput, lo, hi, int(scheme), 0, height=x.shape[0])
        d = torch.from_numpy(spec.view(np.float32).copy()).cuda()
        c.work_spectrum_device(x.shape[0], d.data_ptr())
        got = np.asarray(c._rows)
        if loginput:
            assert np.max(np.abs(got - rows)) < 5e-4, key               # dB values (log per bin, then the mean, as in the reference)
        else:
            assert np.max(np.abs(got - rows) / np.maximum(np.abs(rows), 1e-30)) < 5e-6, key
