"""CPU tests (no GPU): oracle pinned against the known-answer values harvested from the reference,
and the product's host-side logic -- geometry, window tables, argument errors, activity state machines -- against the
oracle through the C ABI's host-only entry points."""
import ctypes
import os
import re

import numpy as np
import pytest

import geometry
import scenarios as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def FDC():
    import FDC as m
    m._cabi.lib()
    return m


# ------------------------------------------------------------------------------------------------ C ABI surface
def test_cabi_exports_every_declared_symbol(FDC):
    hdr = open(os.path.join(ROOT, "include", "fdc_cabi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fdc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 70
    L = ctypes.CDLL(FDC.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(FDC._cabi.SIGNATURES), declared ^ set(FDC._cabi.SIGNATURES)
    assert L.fdc_api_version() == 1


def test_host_evict_keeps_the_data(FDC):
    """fdc_host_evict only drops cache lines (clflushopt): any range, aligned or not, keeps its contents; needs no GPU"""
    L = FDC._cabi.lib()
    rng = np.random.default_rng(5)
    a = rng.integers(0, 255, size=1 << 20, dtype=np.uint8)
    want = a.copy()
    for off, n in ((0, a.size), (1, 63), (64, 64), (4097, 100000), (a.size - 5, 5), (17, 0)):
        L.fdc_host_evict(a.ctypes.data + off, n)
    L.fdc_host_evict(None, 128)
    assert np.array_equal(a, want)
    a[100:200] = 7                                   # dirty lines are written back, not lost
    L.fdc_host_evict(a.ctypes.data, a.size)
    assert np.all(a[100:200] == 7) and np.array_equal(a[200:], want[200:])


def test_no_cpu_fallback(FDC):
    if FDC._cabi.lib().fdc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    for mk in (lambda: FDC.overlap_save(8, 1024, 512), lambda: FDC.vector_cut_vxx(8, 8, 1, 2),
               lambda: FDC.phase_shifting_windowing_vcc(64, 2, 1, 0.5, 0.8, 0), lambda: FDC.fft_vcc(64, True, None, True, 1),
               lambda: FDC.Channelizer(1024, 512, 2, []),
               lambda: FDC.PowerActivationChannel(256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7),
               lambda: FDC.SegmentDetection(3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, True, False, "", False, 0),
               lambda: FDC.activity_detection_channelizer_vcm(256, [[0.1, 0.9]], 10.0, 4, 4, True, False, "", False, 0.0625, 1, 0.2, 0)):
        with pytest.raises(FDC.FDCError, match="no usable CUDA device"):
            mk()


# ------------------------------------------------------------------------------------------------ oracle pinned (Appendix B)
def test_oracle_kat_b1_b3_tables(ref):
    t = ref.phase_shifting_windowing_vcc(200, 4, 3, 0.5, 0.75, 2).tables()
    w0 = np.abs(t[0])
    assert np.count_nonzero(w0 == 0) == 50 and np.all(w0[:25] == 0) and np.all(w0[-25:] == 0)
    assert np.isclose(w0[25], 0.000192307692, rtol=1e-6) and np.isclose(w0[49], 0.00480769249, rtol=1e-6)
    assert np.float32(w0[50]) == np.float32(0.005) and np.all(w0[50:150] == w0[50])
    assert t[1][100] == np.complex64(complex(3.06161698e-19, 0.00499999989))
    assert t[2][100] == np.complex64(complex(-0.00499999989, 6.12323395e-19))
    t = ref.phase_shifting_windowing_vcc(256, 4, 5, 0.55, 0.8, 1).tables()
    assert np.isclose(t[0][25].real, 8.33168724e-06, rtol=1e-6) and np.isclose(t[0][57].real, 0.00389791839, rtol=1e-6)
    assert t[0][58].real == np.float32(0.00390625) and np.count_nonzero(t[0] == 0) == 50
    b = ref.phase_shifting_windowing_vcc(64, 2, 7, 0.88, 1.0, 0)
    t = b.tables()
    assert np.all(t[0][[0, 1, 62, 63]] == 0) and np.all(t[0][2:62] == np.complex64(0.015625)) and b.state()["shift"] == 1
    # phase sequence of B.1: six all-ones blocks as work(2), work(4)
    a = ref.phase_shifting_windowing_vcc(200, 4, 3, 0.5, 0.75, 2)
    ones = np.ones(200 * 6, dtype=np.complex64)
    y = np.concatenate([a.work(ones[:400]).view(np.complex64), a.work(ones[400:]).view(np.complex64)]).reshape(6, 200)
    ph = np.round(y[:, 100] / np.float32(0.005))
    assert list(ph) == [1, -1j, -1, 1j, 1, -1j]


def test_oracle_kat_b4_b5(ref):
    x = np.arange(1, 31, dtype=np.float32)
    y = ref.overlap_save(4, 8, 2).work(x).view(np.float32)
    assert list(y[:16]) == [0, 0, 1, 2, 3, 4, 5, 6, 5, 6, 7, 8, 9, 10, 11, 12]
    b = ref.overlap_save(4, 8, 2)
    y2 = np.concatenate([b.work(x[:6]).view(np.float32), b.work(x[6:18]).view(np.float32), b.work(x[18:]).view(np.float32)])
    assert np.array_equal(y, y2)
    z = ref.vector_cut_vxx(4, 4, 1, 2).work(np.array([0, 1, 2, 3, 10, 11, 12, 13], dtype=np.float32)).view(np.float32)
    assert list(z) == [1, 2, 11, 12]


B6 = [((0.12, 0.05), (2412, 256, 192, 0.88, 1.0)), ((0.22, 0.1), (2693, 512, 384, 0.88, 1.0)),
      ((-0.14, 0.12), (963, 1024, 768, 0.528, 0.778)), ((0.0, 0.081), (1792, 512, 384, 0.7128, 1.0))]


def test_geometry_b6(FDC):
    for (fq, bw), want in B6:
        for got in (geometry.get_opt_channelparams(4096, 4, geometry.get_freq(fq), geometry.get_bw(bw)),
                    FDC.opt_channelparams(4096, 4, geometry.get_freq(fq), geometry.get_bw(bw))):
            assert got[:3] == want[:3]
            assert abs(got[3] - want[3]) < 1e-9 and abs(got[4] - want[4]) < 1e-9


def test_geometry_native_equals_python(FDC):
    rng = np.random.default_rng(0)
    for _ in range(3000):
        N = 1 << int(rng.integers(6, 19)); R = 1 << int(rng.integers(1, 4))
        fq = float(rng.uniform(-0.5, 0.5)); bw = float(rng.uniform(2.0 / N, 0.6))
        a = geometry.get_opt_channelparams(N, R, geometry.get_freq(fq), geometry.get_bw(bw))
        b = FDC.opt_channelparams(N, R, geometry.get_freq(fq), geometry.get_bw(bw))
        assert a == b, (N, R, fq, bw, a, b)


def test_oracle_kat_b7_b8_pac(ref):
    want = {(0.12, 0.05): (2412, 2668, 256, 2437, 2642, 0, 192), (0.22, 0.1): (2693, 3205, 512, 2744, 3154, 1, 384),
            (-0.14, 0.12): (1219, 1731, 512, 1229, 1720, 3, 384), (0.0, 0.081): (1792, 2304, 512, 1882, 2214, 0, 384)}
    for (fq, bw), w in want.items():
        st = ref.PowerActivationChannel(4096, geometry.get_freq(fq), geometry.get_bw(bw), 4, 4.0, 0, 0, True, False, "", 0, 0).state()
        got = (st["extract_start"], st["extract_stop"], st["extract_width"], st["measure_start"], st["measure_stop"], st["deltaphase"], st["output_len"])
        assert got == w
    b = ref.PowerActivationChannel(256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7)
    x = sc.b8_input().reshape(-1)
    for n0, n1 in ((0, 5), (5, 8), (8, 16)):
        b.work(x[n0 * 256:n1 * 256])
    msgs = b.messages()
    assert [(m["part"], m["blockstart"], m["blockend"], m["data"].size, m["finalized"]) for m in msgs] == \
        [(0, 3, 6, 72, False), (1, 3, 9, 72, False), (2, 3, 11, 48, True)]
    assert abs(msgs[0]["rel_cfreq"] - 0.44921875) < 1e-9 and msgs[0]["rel_bw"] == 0.125
    assert sc.strip_time(msgs[2]["ID"]) == "PowActChan.7.0.fin"


def test_oracle_kat_b9_segdet(ref):
    b = ref.SegmentDetection(3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, True, False, "", False, 0)
    st = b.state()
    assert (st["d_start"], st["d_stop"], st["d_width"], st["D"], st["M"]) == (24, 232, 208, 8, 26)
    x = sc.b9_input().reshape(-1)
    pos = 0
    for n in (3, 4, 2, 11):
        b.work(x[pos * 256:(pos + n) * 256]); pos += n
    got = [(sc.strip_time(m["ID"]), m["finalized"], m["part"], m["blockstart"], m["blockend"], m["vectorstart"], m["vectorend"], m["data"].size)
           for m in b.messages()]
    assert got == [("DETECTED.3.0", False, 0, 2, 6, 80, 144, 192), ("DETECTED.3.1", False, 0, 4, 8, 132, 196, 192),
                   ("DETECTED.3.0", False, 1, 2, 10, 80, 144, 192), ("DETECTED.3.0", True, 2, 3, 11, 80, 144, 0),
                   ("DETECTED.3.1", False, 1, 4, 12, 132, 196, 192), ("DETECTED.3.1", True, 2, 5, 16, 132, 196, 144)]


# ------------------------------------------------------------------------------------------------ product host logic vs oracle
@pytest.mark.parametrize("args", [(200, 4, 3, 0.5, 0.75, 2), (256, 4, 5, 0.55, 0.8, 1), (64, 2, 7, 0.88, 1.0, 0), (512, 4, 1, 0.55, 0.8, 1),
                                  (128, 4, 2, 0.55, 0.8, 2), (1024, 8, 3, 0.3, 0.9, 1), (64, 2, 0, 1.5, 1.6, 2), (32, 4, 1, 0.9, 1.7, 1)])
def test_psw_tables_bit_exact(FDC, ref, args):
    blocklen, R, shifts, pb, sb, wt = args
    want = ref.phase_shifting_windowing_vcc(*args).tables()
    got = FDC.psw_tables(blocklen, R, pb, sb, wt)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("bad", [(64, 2, 1, 0.0, 0.5, 0), (64, 2, 1, 0.5, -0.1, 0), (64, 2, 1, 0.6, 0.5, 0)])
def test_psw_argument_errors_match(FDC, ref, bad):
    with pytest.raises(ref.RefError) as e1:
        ref.phase_shifting_windowing_vcc(*bad)
    with pytest.raises(FDC.FDCError) as e2:
        FDC.psw_tables(bad[0], bad[1], bad[3], bad[4], bad[5])
    assert str(e1.value) in str(e2.value)


PAC_CASES = [(4096, 0.62, 0.05, 4), (4096, 0.72, 0.1, 4), (4096, 0.36, 0.12, 4), (4096, 0.5, 0.081, 4), (256, 0.45, 0.1, 4),
             (256, 0.97, 0.05, 4), (256, 0.90, 0.2, 4), (1024, 0.25, 0.3, 2), (16384, 0.1234, 0.01, 8), (64, 0.5, 1.0, 4)]


@pytest.mark.parametrize("case", PAC_CASES)
def test_pac_geometry_and_tables(FDC, ref, case):
    N, cf, bw, R = case
    a = ref.PowerActivationChannel(N, cf, bw, R, 6.0, 3, 1, True, False, "", 0, 7)
    b = FDC.PowerActivationChannel(N, cf, bw, R, 6.0, 3, 1, True, False, "", 0, 7, _logic=True)
    sa, sb = a.state(), b.state()
    sa.pop("count"); sb.pop("count")           # the reference leaves `count` uninitialised until the first activation
    assert sa == sb
    assert np.array_equal(a.tables().view(np.uint32), b.tables().view(np.uint32))


@pytest.mark.parametrize("bad", [(0, 0.5, 0.1, 4, 6.0), (256, 0.5, 0.1, 3, 6.0), (256, 0.02, 0.1, 4, 6.0), (256, 0.5, 0.1, 4, 0.0), (256, 0.5, 1.5, 4, 3.0)])
def test_pac_argument_errors_match(FDC, ref, bad):
    N, cf, bw, R, th = bad
    with pytest.raises(ref.RefError) as e1:
        ref.PowerActivationChannel(N, cf, bw, R, th, 3, 1, True, False, "", 0, 7)
    with pytest.raises(FDC.FDCError) as e2:
        FDC.PowerActivationChannel(N, cf, bw, R, th, 3, 1, True, False, "", 0, 7, _logic=True)
    assert str(e1.value) in str(e2.value)


SEG_CASES = [(256, 4, 0.1, 0.9, 0.0625, 0.2), (4096, 4, 0.6, 0.95, 0.002, 0.2), (16384, 4, 0.55, 0.9, 0.002, 0.2), (1024, 2, 0.9, 0.2, 0.01, 0.1),
             (256, 4, 0.02, 0.999, 0.0625, 0.2), (2048, 8, 0.3, 0.35, 0.0001, 0.0), (512, 4, 1.25, 0.75, 0.05, 0.45)]


@pytest.mark.parametrize("case", SEG_CASES)
def test_segdet_geometry_and_windows(FDC, ref, case):
    N, R, s0, s1, mcd, fl = case
    a = ref.SegmentDetection(3, N, R, s0, s1, 10.0, mcd, fl, 4, 1, True, False, "", False, 0)
    b = FDC.SegmentDetection(3, N, R, s0, s1, 10.0, mcd, fl, 4, 1, True, False, "", False, 0, _logic=True)
    sa, sb = a.state(), b.state()
    for k in ("d_start", "d_stop", "d_width", "D", "M", "blockcount", "thresh"):
        assert sa[k] == sb[k], k
    for lg in range(int(np.log2(N)) + 1):
        for ph in range(R):
            assert np.array_equal(a.window(lg, ph).view(np.uint32), b.window(lg, ph).view(np.uint32))


@pytest.mark.parametrize("bad", [dict(N=300), dict(R=3), dict(th=-1.0), dict(fl=-0.1), dict(s0=0.0, s1=1.0)])
def test_segdet_argument_errors_match(FDC, ref, bad):
    kw = dict(N=256, R=4, s0=0.1, s1=0.9, th=10.0, fl=0.2); kw.update(bad)
    args = (3, kw["N"], kw["R"], kw["s0"], kw["s1"], kw["th"], 0.0625, kw["fl"], 4, 1, True, False, "", False, 0)
    with pytest.raises(ref.RefError) as e1:
        ref.SegmentDetection(*args)
    with pytest.raises(FDC.FDCError) as e2:
        FDC.SegmentDetection(*args, _logic=True)
    assert str(e1.value) in str(e2.value)


def _feed(block, x, N, chunks):
    pos = 0
    for n in chunks:
        block.work(x[pos * N:(pos + n) * N]); pos += n
    assert pos * N == x.size


def test_pac_state_machine_b8(FDC, ref):
    x = sc.b8_input()
    a = ref.PowerActivationChannel(256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7)
    _feed(a, x.reshape(-1), 256, (5, 3, 8))
    b = FDC.PowerActivationChannel(256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7, _logic=True)
    st = b.state()
    pw = sc.band_power(x, st["measure_start"], st["measure_stop"])
    for n0, n1 in ((0, 2), (2, 9), (9, 16)):                       # a different chunking than the oracle's
        b.logic_work(n1 - n0, pw[n0:n1])
    assert a.state() == b.state()
    ma, mb = a.messages(), b.messages()
    assert [sc.meta_tuple(m) for m in ma] == [sc.meta_tuple(m) for m in mb]


@pytest.mark.parametrize("maxblocks,delay", [(4, 1), (0, 0), (-1, 2), (1, 0), (7, 3)])
def test_segdet_state_machine_vs_oracle(FDC, ref, maxblocks, delay):
    N = 1024
    x, truth = sc.bursty_spectra(N, 120, 8, seed=21 + maxblocks, widths=(16, 32, 64), mean_on=9, mean_off=12)
    args = (5, N, 4, 0.1, 0.9, 10.0, 0.0312, 0.2, maxblocks, delay, True, False, "", False, 0)
    a = ref.SegmentDetection(*args)
    _feed(a, x.reshape(-1), N, (7, 1, 50, 62))
    b = FDC.SegmentDetection(*args, _logic=True)
    st = b.state()
    P = sc.group_power(x, st["d_start"], st["D"], st["M"])
    pos = 0
    for n in (30, 30, 1, 59):
        b.logic_work(n, P[pos:pos + n]); pos += n
    ma, mb = a.messages(), b.messages()
    assert len(ma) >= 3
    assert [sc.meta_tuple(m) for m in ma] == [sc.meta_tuple(m) for m in mb]
    assert a.active_channels() == b.active_channels()
    sa, sb = a.state(), b.state()
    assert (sa["blockcount"], sa["n_active"], sa["chan_counter"]) == (sb["blockcount"], sb["n_active"], sb["chan_counter"])
    assert np.array_equal(a.power(), b.power())


@pytest.mark.parametrize("kind", ["noise", "crowded", "ties"])
def test_segdet_dense_scenes_vs_oracle(FDC, ref, kind):
    """hundreds of candidates per block: the candidate overlap test and the candidate/channel matching are done with binary
    searches here and with nested walks in the reference -- decisions, order of activation and PDUs must not differ"""
    if kind == "noise":              # D = 1 on pure noise: touching and overlapping candidates everywhere
        N = 4096
        rng = np.random.default_rng(9)
        x = (rng.standard_normal((14, N)) + 1j * rng.standard_normal((14, N))).astype(np.complex64)
        args = (0, N, 4, 0.02, 0.999, 3.0, 0.0001, 0.1, 2, 0, True, False, "", False, 0)
        chunks_a, chunks_b = (14,), (5, 9)
    elif kind == "ties":             # hundreds of rising edges with EXACTLY equal ratios per block: the order among them is whatever the
        N = 4096                     # reference's std::sort call leaves, and it decides channel ids and which overlapping candidate wins
        rng = np.random.default_rng(21)
        amp = np.ones((20, N), dtype=np.float32)
        for b_ in range(20):
            on = rng.random(N // 8) < 0.55          # 8-bin cells: 3 bins at amplitude 3 (power 9, ratio exactly 9) when the cell is on
            for c_ in np.nonzero(on)[0]:
                w_ = int(rng.integers(2, 5))
                amp[b_, 8 * c_ + 2:8 * c_ + 2 + w_] = 3.0
        x = amp.astype(np.complex64)
        args = (1, N, 4, 0.02, 0.98, 6.0, 0.0001, 0.1, 3, 1, True, False, "", False, 0)
        chunks_a, chunks_b = (20,), (7, 13)
    else:                            # several hundred narrow carriers switching on and off
        N = 32768
        x, _ = sc.bursty_spectra(N, 40, 400, seed=77, raster=64, widths=(24, 40), mean_on=3, mean_off=9, lo=0.05, hi=0.95)
        args = (4, N, 4, 0.05, 0.95, 10.0, 16.0 / N, 0.2, 3, 1, True, False, "", False, 0)
        chunks_a, chunks_b = (40,), (13, 1, 26)
    a = ref.SegmentDetection(*args)
    _feed(a, x.reshape(-1), N, chunks_a)
    b = FDC.SegmentDetection(*args, _logic=True)
    st = b.state()
    P = sc.group_power(x, st["d_start"], st["D"], st["M"])
    pos = 0
    for n in chunks_b:
        b.logic_work(n, P[pos:pos + n]); pos += n
    ma, mb = a.messages(), b.messages()
    assert len(ma) > (50 if kind == "noise" else 300)
    assert [sc.meta_tuple(m) for m in ma] == [sc.meta_tuple(m) for m in mb]
    assert a.active_channels() == b.active_channels()
    sa, sb = a.state(), b.state()
    assert (sa["blockcount"], sa["n_active"], sa["chan_counter"]) == (sb["blockcount"], sb["n_active"], sb["chan_counter"])


@pytest.mark.parametrize("seed", range(16))
def test_segdet_randomised_differential(FDC, ref, seed):
    """random geometry (raster 1 .. 32 bins), thresholds, deactivation delays, partial-emission limits and scenes (noise only,
    sparse, crowded, touching carriers): the bookkeeping must make the reference's decisions, in its order"""
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.choice([1024, 2048, 4096]))
    nblocks = int(rng.integers(12, 40))
    kind = seed % 4
    if kind == 0:
        x = (rng.standard_normal((nblocks, N)) + 1j * rng.standard_normal((nblocks, N))).astype(np.complex64)
    else:
        ncar = [0, 6, 40, 90][kind]
        w = [(16, 32, 64), (16, 32, 64), (8, 16), (4, 8)][kind]
        x, _ = sc.bursty_spectra(N, nblocks, ncar, seed=seed, widths=w, raster=[0, 0, 32, 8][kind] or None, mean_on=int(rng.integers(2, 9)),
                                 mean_off=int(rng.integers(3, 12)), snr_db=float(rng.choice([12.0, 25.0])), lo=0.05, hi=0.95)
    thresh = float(rng.choice([3.0, 6.0, 10.0]))
    mcd = float(rng.choice([1.0, 2.0, 4.0, 8.0, 16.0, 32.0])) / N
    args = (seed, N, int(rng.choice([2, 4, 8])), 0.05, 0.95, thresh, mcd, float(rng.choice([0.0, 0.2, 0.5])), int(rng.integers(-1, 6)),
            int(rng.integers(0, 4)), True, False, "", False, 0)
    a = ref.SegmentDetection(*args)
    _feed(a, x.reshape(-1), N, (nblocks,))
    b = FDC.SegmentDetection(*args, _logic=True)
    st = b.state()
    assert {k: st[k] for k in ("d_start", "d_stop", "D", "M")} == {k: a.state()[k] for k in ("d_start", "d_stop", "D", "M")}
    P = sc.group_power(x, st["d_start"], st["D"], st["M"])
    cut = int(rng.integers(1, nblocks))
    b.logic_work(cut, P[:cut]); b.logic_work(nblocks - cut, P[cut:])
    assert [sc.meta_tuple(m) for m in a.messages()] == [sc.meta_tuple(m) for m in b.messages()]
    assert a.active_channels() == b.active_channels()
    sa, sb = a.state(), b.state()
    assert (sa["blockcount"], sa["n_active"], sa["chan_counter"]) == (sb["blockcount"], sb["n_active"], sb["chan_counter"])


@pytest.mark.parametrize("threads", [False, True])
def test_actdet_state_machine_vs_oracle(FDC, ref, threads):
    N = 1024
    x, truth = sc.bursty_spectra(N, 90, 10, seed=33, widths=(16, 32), mean_on=8, mean_off=10)
    segs = [[0.1, 0.45], [0.55, 0.9]]
    args = (N, segs, 10.0, 4, 4, True, False, "", threads, 0.0312, 1, 0.2, 0)
    a = ref.activity_detection_channelizer_vcm(*args)
    _feed(a, x.reshape(-1), N, (11, 30, 49))
    b = FDC.activity_detection_channelizer_vcm(*args, _logic=True)
    geo = lambda segs_: [{k: v for k, v in s_.items() if k != "n_active"} for s_ in segs_]
    assert geo(a.segments()) == geo(b.segments())
    P = np.concatenate([sc.group_power(x, s["start"], s["D"], s["M"], mean=True) for s in b.segments()], axis=1)
    pos = 0
    for n in (45, 45):
        b.logic_work(n, P[pos:pos + n]); pos += n
    ma, mb = a.messages(), b.messages()
    assert len(ma) >= 3
    key = (lambda m: sc.meta_tuple(m))
    if threads:          # worker threads publish in completion order in the reference: compare as sets per block
        assert sorted(map(key, ma)) == sorted(map(key, mb))
    else:
        assert list(map(key, ma)) == list(map(key, mb))
    assert a.segments() == b.segments()
    for i in range(len(segs)):
        assert np.array_equal(a.power(i), b.power(i))


# ------------------------------------------------------------------------------------------------ GRC descriptors
def test_grc_yaml_descriptors_match_the_block_api():
    """GNU Radio >= 3.8 descriptors (gr-fdc_b200/grc): ids and make templates are the reference's keys / make strings
    (grc/FDC_*.xml), every make argument is a declared parameter"""
    import glob
    import yaml
    files = sorted(glob.glob(os.path.join(ROOT, "gr-fdc_b200", "grc", "*.block.yml")))
    assert len(files) == 7
    want = {"FDC_overlap_save": ("FDC.overlap_save", 3), "FDC_vector_cut_vxx": ("FDC.vector_cut_vxx", 4),
            "FDC_phase_shifting_windowing_vcc": ("FDC.phase_shifting_windowing_vcc", 6), "FDC_PowerActivationChannel": ("FDC.PowerActivationChannel", 12),
            "FDC_SegmentDetection": ("FDC.SegmentDetection", 15), "FDC_activity_detection_channelizer_vcm": ("FDC.activity_detection_channelizer_vcm", 13),
            "FDC_FrequencyDomainChannelizer": ("FDC.FrequencyDomainChannelizer", 25)}
    for f in files:
        d = yaml.safe_load(open(f))
        fn, nargs = want[d["id"]]
        make = d["templates"]["make"]
        assert make.startswith(fn + "(")
        args = re.findall(r"\$\{\s*([A-Za-z_][A-Za-z0-9_]*)", make)
        assert len(args) == nargs, (d["id"], args)
        ids = {p["id"] for p in d["parameters"]}
        assert set(args) <= ids, (d["id"], set(args) - ids)
