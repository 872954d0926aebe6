"""Real multi-GPU runs of the sharded paths (channel-sharded sinks, instance-sharded activity sinks) under torchrun; needs at
least two GPUs on the box (skipped otherwise -- the host-side logic of the same paths runs on CPU in test_sharded_gloo.py and on
one GPU with virtual ranks in test_gpu_chan.py / test_gpu_activity.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_paths_on_real_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU only")
    n = min(n, 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_check.py")], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multirank ok" in r.stdout
