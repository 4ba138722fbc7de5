"""Multi-GPU correctness under pytest: with >= 2 GPUs visible, 2 ranks (torchrun, NCCL) trace shards of the SAME
injected bundle; the all-reduced image, messages, extent and spectra must equal the single-GPU result — pixel counts
bit for bit — and restricted focus searches / one-rank status bits must not hang (tests/mgpu_worker.py)."""
import json
import os
import pathlib
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_two_ranks_equal_one_rank(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    rep = tmp_path / "report.json"
    env = dict(os.environ, MGPU_REPORT=str(rep))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "tests" / "mgpu_worker.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    d = json.loads(rep.read_text())
    assert d["world"] == 2 and d["status_or"] and d["double_gauss"]["hits"] > 0
