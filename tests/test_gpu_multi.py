"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a single-GPU box): runs tests/mg_check.py under torchrun -
the hash-partitioned streaming build of ONE file split by byte range (records over NVLink peer stores, receiver-side
split, region sweep), the distributed stages 2-5 and the CLI, against the oracle on the whole file."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_mg_check_torchrun():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "mg_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    ok = [l for l in out.stdout.splitlines() if l.startswith("mg_check ok")]
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert len(ok) == 10, out.stdout[-2000:]           # per file: 2 dBG lines, 2 graph lines, 1 CLI line
