"""GPU parity of the multi-GPU encode (hsrle_b200.sliced -> hsrle_slice_compress_phase): one stream cut into
slices, one rank per slice, stream gathered and compared with the oracle bit for bit.  On a single-GPU box the
ranks share cuda:0 and exchange their messages over gloo; with >= 2 GPUs the NCCL path runs as well."""
import os
import subprocess
import sys

import pytest

from common import CODECS, ROOT

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(mode, world, codecs, which, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "sliced_worker.py"), mode, ",".join(codecs), which]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mismatches=0" in r.stdout, r.stdout[-3000:]


def test_sliced_shared_gpu_world2_all_codecs():
    _run("gpu1", 2, [c.name for c in CODECS], "all", 29621)


def test_sliced_shared_gpu_world3():
    _run("gpu1", 3, [c.name for c in CODECS][::3], "quick", 29622)


def test_sliced_nccl():
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    _run("gpu", min(ng, 4), [c.name for c in CODECS][::2], "all", 29623)
