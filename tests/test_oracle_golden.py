"""CPU tests (no GPU): the oracle pinned against the committed golden vectors.

tests/golden/*.npz|json were produced by the reference's own code (oracle/_ref = unmodified lib/*_impl.cc) with
tests/golden/make_golden.py.  Here (a) the fp64 NumPy restatement oracle/fdc_numpy.py is checked against them,
(b) when oracle/_ref is present, the live reference build is checked against them too (drift of the shim or the
build recipe), (c) the seeded input generators are checked to still produce the stored inputs."""
import json
import os

import numpy as np
import pytest

import geometry
import scenarios as sc
import workloads
from helpers import rel_l2
from oracle import fdc_numpy as fnp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHAINS = ["chain_n1024_r4_hann", "chain_n512_r2_rect", "chain_n2048_r8_ramp"]


def load_chain(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    N, R, wintype, nblocks, seed = [int(v) for v in g["meta"]]
    cfg = workloads.cfg_example(N, R, wintype)
    return g, cfg, nblocks, seed


@pytest.mark.parametrize("name", CHAINS)
def test_numpy_restatement_matches_reference_vectors(name):
    g, cfg, nblocks, seed = load_chain(name)
    # geometry integers: bit exact
    assert [list(p[:3]) for p in cfg.params] == g["params"].tolist()
    assert np.array_equal(np.array([p[3:] for p in cfg.params]), g["bands"])
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=seed)
    assert np.array_equal(x.view(np.uint8), g["x"].view(np.uint8)), "input generator drifted"
    outs, spec = fnp.channelize(x, cfg.N, cfg.R, cfg.params, cfg.windowtype, want_spectrum=True)
    assert rel_l2(spec[0], g["spectrum_first_block"]) < 2e-7
    assert rel_l2(spec[-1], g["spectrum_last_block"]) < 2e-7
    for i, o in enumerate(outs):
        want = g["out%d" % i]
        assert o.size == want.size == nblocks * cfg.params[i][2]
        # the reference vectors are fp32 results of an fp64-accurate FFT: only the final rounding differs
        assert rel_l2(o, want) < 3e-7, (name, i)


@pytest.mark.parametrize("name", CHAINS)
def test_live_reference_build_matches_its_vectors(ref, name):
    g, cfg, nblocks, seed = load_chain(name)
    outs, spec = ref.Chain(cfg.N, cfg.R, cfg.params, cfg.windowtype).run(g["x"], nthreads=1, want_spectrum=True)
    assert np.array_equal(spec[:cfg.N].view(np.uint8), g["spectrum_first_block"].view(np.uint8))
    for i, o in enumerate(outs):
        assert np.array_equal(o.view(np.uint8), g["out%d" % i].view(np.uint8))


def test_window_tables_restatement_bit_exact():
    g = np.load(os.path.join(GOLD, "psw_tables.npz"))
    for key in ("b1", "b2", "b3", "cfg4", "r8"):
        blocklen, nphase, shifts, pb, sb, wt = g[key + "_args"]
        t = fnp.psw_tables(int(blocklen), int(nphase), pb, sb, int(wt))
        assert np.array_equal(t.view(np.uint8), g[key + "_tables"].view(np.uint8)), key
        # work(): block b multiplies by table[(b * shift) % R], VOLK generic complex multiply (fp32, no FMA)
        x = g[key + "_x"].reshape(-1, int(blocklen)); R = int(nphase); shift = ((int(shifts) % R) + R) % R
        w = t[(np.arange(x.shape[0]) * shift) % R]
        re = (x.real * w.real).astype(np.float32) - (x.imag * w.imag).astype(np.float32)
        im = (x.real * w.imag).astype(np.float32) + (x.imag * w.real).astype(np.float32)
        y = np.empty(x.shape, dtype=np.complex64); y.real = re; y.imag = im; y = y.reshape(-1)
        assert np.array_equal(y.view(np.uint8), g[key + "_y"].view(np.uint8)), key


def test_copy_blocks_restatement():
    g = np.load(os.path.join(GOLD, "copy_blocks.npz"))
    blocks, hist = fnp.overlap_save(g["ovl_x"], 8, 2)
    assert np.array_equal(blocks.reshape(-1), g["ovl_y"])
    assert list(hist) == [59.0, 60.0]
    assert list(g["ovl_y"][:16]) == [0, 0, 1, 2, 3, 4, 5, 6, 5, 6, 7, 8, 9, 10, 11, 12]
    v = fnp.vector_cut(np.array([[0, 1, 2, 3], [10, 11, 12, 13]], dtype=np.float32), 1, 2)
    assert np.array_equal(v.reshape(-1), g["cut_y"]) and list(g["cut_y"]) == [1, 2, 11, 12]      # Appendix B.5


def test_activity_kats_match_survey_appendix_b():
    k = json.load(open(os.path.join(GOLD, "activity_kat.json")))
    # B.6 geometry (f, l, lout) of examples/FDC_example.grc
    assert [p[:3] for p in k["geometry_b6"]] == [[2412, 256, 192], [2693, 512, 384], [963, 1024, 768], [1792, 512, 384]]
    assert [list(geometry.get_opt_channelparams(4096, 4, geometry.get_freq(f), geometry.get_bw(bw))) for (f, bw) in
            workloads.example_channels()] == k["geometry_b6"]
    # B.8: (part, blockstart, blockend, samples)
    assert [(m[2], m[3], m[4], m[9]) for m in k["pac_b8"]["msgs"]] == [(0, 3, 6, 72), (1, 3, 9, 72), (2, 3, 11, 48)]
    # B.9: ID suffix, finalized, part, blockstart, blockend, vectorstart, vectorend, samples
    got = [(m[0].split("DETECTED.")[1], m[1], m[2], m[3], m[4], m[5], m[6], m[9]) for m in k["segdet_b9"]["msgs"]]
    assert got == [("3.0", False, 0, 2, 6, 80, 144, 192), ("3.1", False, 0, 4, 8, 132, 196, 192),
                   ("3.0", False, 1, 2, 10, 80, 144, 192), ("3.0", True, 2, 3, 11, 80, 144, 0),
                   ("3.1", False, 1, 4, 12, 132, 196, 192), ("3.1", True, 2, 5, 16, 132, 196, 144)]
    # B.11 = B.9 with block indices + 1 and segment index 0
    got = [(m[0].split("DETECTED.")[1], m[1], m[2], m[3], m[4], m[9]) for m in k["actdet_b11"]["msgs"]]
    assert got == [("0.0", False, 0, 3, 7, 192), ("0.1", False, 0, 5, 9, 192), ("0.0", False, 1, 3, 11, 192),
                   ("0.0", True, 2, 4, 12, 0), ("0.1", False, 1, 5, 13, 192), ("0.1", True, 2, 6, 17, 144)]


def test_live_reference_activity_matches_kats(ref):
    k = json.load(open(os.path.join(GOLD, "activity_kat.json")))
    x = sc.b9_input()
    s = ref.SegmentDetection(3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, True, False, "", False, 0); msgs = []
    for a, e in ((0, 3), (3, 7), (7, 9), (9, 20)):
        s.work(x[a:e]); msgs += s.messages()
    assert [list(sc.meta_tuple(m)) for m in msgs] == k["segdet_b9"]["msgs"]


def test_frequency_modes_against_the_reference_lines():
    """normalized / basebandfs / centerfreqfs (python/FrequencyDomainChannelizer.py:70-91) and the geometry derived from the
    converted values (:322-345): tests/golden/freqmodes.json holds what the reference's own source lines return
    (tests/golden/make_freqmode_golden.py executes them); the mirror's conversions and geometry.py must reproduce it exactly"""
    import json
    import geometry
    from FDC.FrequencyDomainChannelizer import frequency_conversions
    with open(os.path.join(GOLD, "freqmodes.json")) as fh:
        cases = json.load(fh)["cases"]
    assert len(cases) == 9 and {c["freqmode"] for c in cases} == {"normalized", "basebandfs", "centerfreqfs"}
    for c in cases:
        mode, get_freq, set_freq, get_bw, set_bw = frequency_conversions(c["freqmode"], c["fs"], c["centerfrequency"])
        assert mode == c["freqmode_enum"]
        conv = [[get_freq(f), get_bw(bw)] for f, bw in c["channels"]]
        assert conv == c["normalized_channels"]                                    # same expressions: bit-identical doubles
        assert [[get_freq(a), get_freq(b)] for a, b in c["segments"]] == c["normalized_segments"]
        assert [[set_freq(f), set_bw(bw)] for f, bw in conv] == c["set_freq_bw_of_normalized"]
        got = [list(geometry.get_opt_channelparams(c["blocksize"], c["relinvovl"], f, bw)) for f, bw in conv]
        assert got == c["channel_params"], (c["freqmode"], c["blocksize"])
    # the enum spellings the GRC file passes (grc/FDC_FrequencyDomainChannelizer.xml) and the error of an unknown mode
    assert frequency_conversions(1, 2.0, 0.0)[0] == 1 and frequency_conversions(2, 2.0, 1.0)[0] == 2
    with pytest.raises(ValueError, match="Unknown Frequency mode"):
        frequency_conversions("kHz", 1.0, 0.0)
