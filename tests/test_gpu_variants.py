"""GPU parity of the kernel variants behind the run-time switches (FDC_PREFETCH, FDC_EXTRACT_E8, FDC_FWD_SPLIT, FDC_STREAMS,
FDC_PDL).  The switches are read once per process, so every variant runs the parity tests of test_gpu_chan.py in a
subprocess with its environment."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "no_prefetch_one_stream_no_pdl": {"FDC_PREFETCH": "0", "FDC_STREAMS": "1", "FDC_PDL": "0"},
    "prefetch_everywhere_four_streams": {"FDC_PREFETCH": "3", "FDC_STREAMS": "4"},
    "extract_8_points_per_thread": {"FDC_EXTRACT_E8": "1"},
    "extract_8_points_per_thread_prefetch": {"FDC_EXTRACT_E8": "1", "FDC_PREFETCH": "3"},
    "extract_16_points_per_thread": {"FDC_EXTRACT_E32": "0"},
    "extract_without_l2_prefetch": {"FDC_L2PF": "0"},
    "extract_one_block_per_tile": {"FDC_PACK": "0"},
    "forward_16_points_per_thread": {"FDC_FWD_E32": "0"},
    "forward_32_points_everywhere": {"FDC_FWD_E32": "2"},
    "four_step_from_4096": {"FDC_FWD_SPLIT": "4096"},
    "one_cta_per_sm": {"FDC_CTAS_PER_SM": "1"},
    "cluster_fused_forward": {"FDC_FUSED": "1"},
    "one_kernel_for_short_transforms": {"FDC_FUSE_SMALL": "1"},
    "one_kernel_two_thread_groups": {"FDC_FUSE_SMALL": "2"},
    "sinks_forwarded_by_the_copy_engines": {"FDC_SINK_DMA": "1"},
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_parity(name):
    env = dict(os.environ); env.update(VARIANTS[name])
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_chan.py"), "-m", "gpu", "-x", "-q",
                        "-k", "golden or chain_matches or sliding or device_path or time_sharded or fft_vcc or sinks"],
                       env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout
