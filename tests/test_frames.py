"""Frame sequences (inputs above the format's u32 ceiling; BASELINE configs[3] / configs[4]): host logic on CPU --
dealing frames to ranks, the size all-gather (gloo, world sizes 2 and 3) and the self-delimiting container."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import CODEC_BY_NAME, ROOT, gen_dct, oracle_compress

HERE = os.path.dirname(os.path.abspath(__file__))


def _frames():
    sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
    try:
        from hsrle_b200 import frames
    except Exception as e:      # the product module refuses to load without its CUDA library
        pytest.skip(str(e))
    return frames


def test_deal_frames_partitions():
    fr = _frames()
    for F in (1, 4, 16, 17):
        for world in (1, 2, 3, 4, 8):
            seen = sorted(f for r in range(world) for f in fr.deal_frames(F, r, world))
            assert seen == list(range(F))
            assert max(len(fr.deal_frames(F, r, world)) for r in range(world)) - min(len(fr.deal_frames(F, r, world)) for r in range(world)) <= 1
    b = fr.frame_bounds(16 << 30)
    assert len(b) == 16 and all(e - s == 1 << 30 for s, e in b)


def test_container_is_self_delimiting():
    fr = _frames()
    data = gen_dct(300000, seed=4)
    codec = CODEC_BY_NAME["rle24_3symlut_byte"]
    parts = [oracle_compress(codec, data[a:b]) for a, b in fr.frame_bounds(len(data), 65536)]
    offs, tot = fr.concat_layout([len(p) for p in parts])
    blob = np.concatenate(parts)
    assert tot == len(blob) and offs[1] == len(parts[0])
    back = fr.split_concat(blob)
    assert len(back) == len(parts) and all(np.array_equal(x, y) for x, y in zip(back, parts))
    with pytest.raises(ValueError):
        fr.split_concat(blob[:-3])


@pytest.mark.parametrize("world", [2, 3])
def test_frame_sequence_over_gloo(world):
    _frames()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29640 + world), os.path.join(HERE, "frames_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mismatches=0" in r.stdout, r.stdout[-3000:]
