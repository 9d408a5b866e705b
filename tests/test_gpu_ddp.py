"""2-GPU NCCL parity test (-m gpu, skipped with fewer than two devices): see tests/ddp_parity_worker.py."""
import json
import os
import socket
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT, have_cuda

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_allreduced_gradient_equals_mean_of_single_gpu_gradients():
    import torch
    if not have_cuda() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "ddp.json")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "ddp_parity_worker.py"), out]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        res = json.load(open(out))
    print(res)
    assert res["ranks_agree"]
    assert res["flat"] < 1e-5 and res["bucketed"] < 1e-5, res          # fp32 path: rel 1e-5
    assert res["n_buckets"] > 1
