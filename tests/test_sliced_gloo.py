"""N>1 path on CPU: hsrle_b200.sliced over gloo (world sizes 2 and 3), engine = the host-side stage simulator.
Checks the slice bookkeeping the kernels share with the simulator (hsrle_slice.cuh: boundary-run fix-up, state
exchange, placement) and the host orchestration (all-gathers, repeat-until-unchanged loop, share gathering)
against the oracle, bit for bit."""
import os
import subprocess
import sys

import pytest

from common import ROOT

HERE = os.path.dirname(os.path.abspath(__file__))
SOME = ["rle8_multi", "rle8_packed_multi", "rle8_3symlut", "rle16_sym", "rle24_byte_packed", "rle32_7symlut_sym", "rle48_byte",
        "rle64_sym_packed", "rle64_3symlut_byte"]


def _run(world, codecs, which, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "sliced_worker.py"), "sim", ",".join(codecs), which]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mismatches=0" in r.stdout, r.stdout[-3000:]


def test_sliced_world2_all_inputs():
    _run(2, SOME, "all", 29611)


def test_sliced_world3_quick():
    _run(3, SOME[:5], "quick", 29612)


def test_slice_bounds():
    sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
    try:
        from hsrle_b200.sliced import SLICE_ALIGN, frame_bounds, slice_bounds
    except Exception as e:      # the product module refuses to load without its CUDA library
        pytest.skip(str(e))
    for n in (1, 1000, SLICE_ALIGN, SLICE_ALIGN + 1, 5 * SLICE_ALIGN + 3, 88473600, (1 << 30)):
        for world in (1, 2, 3, 4, 8):
            b, active = slice_bounds(n, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and 1 <= active <= world
            assert all(b[r] <= b[r + 1] for r in range(world))
            assert all(b[r] % SLICE_ALIGN == 0 for r in range(active))
            assert all(b[r] < b[r + 1] for r in range(active)) and all(b[r] == n for r in range(active, world + 1))
    fr = frame_bounds((1 << 34) + 5)
    assert len(fr) == 17 and fr[0] == (0, 1 << 30) and fr[-1] == (1 << 34, (1 << 34) + 5)
