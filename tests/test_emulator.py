"""CPU test: the tile-FFT engine and every kernel's loader / storer functors, compiled for the HOST from the same headers
the CUDA kernels are built from (csrc/fdc_hd.h) and stepped phase by phase over all thread ids of a CTA
(tests/emu/emu_engine_check.cc).  Checks the index arithmetic of every engine variant (8 / 16 / 32 points per thread,
four-step column / row tiles, channel tiles, job lists) against fp64 DFTs without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("twgen", [True, False])
def test_host_emulation_of_all_kernels(twgen):
    """twgen: pass twiddles generated from every 8th table entry (the library's default build, -DFDC_TWGEN=1) or all loaded"""
    out_dir = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "emu_check_tw" if twgen else "emu_check")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-DFDC_HOST_EMU"] + (["-DFDC_TWGEN=1"] if twgen else []) +
                       ["-I" + os.path.join(ROOT, "gr-fdc_b200", "csrc"),
                        os.path.join(ROOT, "tests", "emu", "emu_engine_check.cc"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0 and lines[-1] == "EMU OK", "\n".join(l for l in lines if not l.endswith(" ok"))
    assert len(lines) > 100


def test_host_emulation_under_address_sanitizer():
    """the same program with -fsanitize=address,undefined: every loader / storer address of every emulated kernel stays
    inside its (exactly sized) buffer, including ragged last tiles and clamped signals (compute-sanitizer is not
    available on the GPU pool, this is the bounds check of the kernels' index arithmetic)"""
    out_dir = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "emu_check_asan")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-ffp-contract=off",
                        "-DFDC_HOST_EMU", "-DFDC_TWGEN=1", "-I" + os.path.join(ROOT, "gr-fdc_b200", "csrc"),
                        os.path.join(ROOT, "tests", "emu", "emu_engine_check.cc"), "-o", exe], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ); env["ASAN_OPTIONS"] = "detect_leaks=0"
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "EMU OK", (r.stdout[-1500:] + r.stderr[-3000:])
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr
