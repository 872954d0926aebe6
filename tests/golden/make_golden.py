#!/usr/bin/env python
"""Generates the committed golden fixtures from the REFERENCE'S OWN CODE run in this container:
oracle/_ref/libfdc_ref.so = the unmodified /root/reference/lib/*_impl.cc compiled against oracle/shim
(oracle/Makefile), with the fp64-accurate FFT stand-in behind gr::fft::fft_complex (fft mode 0).

    make -C oracle && python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the vectors travel as small .npz/.json files.  Inputs are
regenerated from seeds by the tests (numpy default_rng), stored here too so a generator drift is detected."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import geometry            # noqa: E402
import scenarios as sc     # noqa: E402
import workloads           # noqa: E402
from oracle import fdc_ref as ref   # noqa: E402


def chain_case(name, N, R, wintype, nblocks, seed):
    cfg = workloads.cfg_example(N, R, wintype)
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=seed)
    outs, spec = ref.Chain(cfg.N, cfg.R, cfg.params, cfg.windowtype).run(x, nthreads=1, want_spectrum=True)
    d = {"x": x, "spectrum_first_block": spec[:N], "spectrum_last_block": spec[-N:],
         "params": np.array([p[:3] for p in cfg.params], dtype=np.int64),
         "bands": np.array([p[3:] for p in cfg.params], dtype=np.float64),
         "meta": np.array([N, R, wintype, nblocks, seed], dtype=np.int64)}
    for i, o in enumerate(outs):
        d["out%d" % i] = o
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def tables_case():
    d = {}
    for key, args in {"b1": (200, 4, 3, 0.5, 0.75, 2), "b2": (256, 4, 5, 0.55, 0.8, 1), "b3": (64, 2, 7, 0.88, 1.0, 0),
                      "cfg4": (512, 4, 0, 0.55, 0.8, 1), "r8": (512, 8, -3, 0.3, 0.9, 1)}.items():
        b = ref.phase_shifting_windowing_vcc(*args)
        d[key + "_args"] = np.array(args, dtype=np.float64)
        d[key + "_tables"] = b.tables()
        rng = np.random.default_rng(args[0])
        x = (rng.standard_normal(args[0] * 7) + 1j * rng.standard_normal(args[0] * 7)).astype(np.complex64)
        d[key + "_x"] = x
        d[key + "_y"] = np.concatenate([b.work(x[:args[0] * 3]).view(np.complex64), b.work(x[args[0] * 3:]).view(np.complex64)])
    np.savez_compressed(os.path.join(HERE, "psw_tables.npz"), **d)


def copy_blocks_case():
    x = np.arange(1, 61, dtype=np.float32)
    a = ref.overlap_save(4, 8, 2)
    y = np.concatenate([a.work(x[:30]).view(np.float32), a.work(x[30:]).view(np.float32)])
    v = ref.vector_cut_vxx(4, 4, 1, 2).work(np.array([0, 1, 2, 3, 10, 11, 12, 13], dtype=np.float32)).view(np.float32)
    np.savez_compressed(os.path.join(HERE, "copy_blocks.npz"), ovl_x=x, ovl_y=y, cut_y=v)


def activity_case():
    """PDU metadata of the activity-gated blocks on the Appendix B.8 / B.9 / B.11 scenarios (IDs without the time stamp)"""
    res = {}
    b = ref.PowerActivationChannel(256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7)
    x = sc.b8_input(); msgs = []
    for a, e in ((0, 5), (5, 8), (8, 16)):
        b.work(x[a:e]); msgs += b.messages()
    res["pac_b8"] = {"state": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in b.state().items() if k != "lastpower"},
                     "msgs": [list(sc.meta_tuple(m)) for m in msgs]}
    x = sc.b9_input()
    s = ref.SegmentDetection(3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, True, False, "", False, 0); msgs = []
    for a, e in ((0, 3), (3, 7), (7, 9), (9, 20)):
        s.work(x[a:e]); msgs += s.messages()
    st = s.state()
    res["segdet_b9"] = {"state": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in st.items()},
                        "msgs": [list(sc.meta_tuple(m)) for m in msgs]}
    a_ = ref.activity_detection_channelizer_vcm(256, [[0.1, 0.9]], 10.0, 4, 4, True, False, "", False, 0.0625, 1, 0.2, 0); msgs = []
    for a, e in ((0, 3), (3, 7), (7, 9), (9, 20)):
        a_.work(x[a:e]); msgs += a_.messages()
    res["actdet_b11"] = {"segments": a_.segments(), "msgs": [list(sc.meta_tuple(m)) for m in msgs]}
    # hier-block geometry of examples/FDC_example.grc (Appendix B.6) from the restated get_opt_channelparams
    res["geometry_b6"] = [list(geometry.get_opt_channelparams(4096, 4, geometry.get_freq(f), geometry.get_bw(bw)))
                          for (f, bw) in workloads.example_channels()]
    with open(os.path.join(HERE, "activity_kat.json"), "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("build oracle/_ref first: make -C oracle")
    ref.set_fft_mode(0)
    chain_case("chain_n1024_r4_hann", 1024, 4, workloads.HANN, 12, 21)
    chain_case("chain_n512_r2_rect", 512, 2, workloads.RECTANGULAR, 16, 22)
    chain_case("chain_n2048_r8_ramp", 2048, 8, workloads.RAMP, 6, 23)
    tables_case(); copy_blocks_case(); activity_case()
    for f in sorted(os.listdir(HERE)):
        print("%8d  %s" % (os.path.getsize(os.path.join(HERE, f)), f))
