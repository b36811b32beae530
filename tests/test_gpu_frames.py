"""Frame sequences on the GPU (hsrle_b200.frames.FrameCodec): every frame is a complete stream identical to the compiled
reference's for that frame; the concatenated container splits and decodes back.  Small frames here (the logic does not
depend on the frame size); 2^30-byte frames are covered by test_gpu_parity.py and by bench.py --workload configs3."""
import numpy as np
import pytest

from common import CODEC_BY_NAME, gen_run_mixed_pieces, oracle_compress, ref_compress, ref_lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rle8_multi", "rle64_byte", "rle64_byte_packed", "rle32_7symlut_sym"])
def test_frame_codec_vs_reference(name):
    import torch
    from hsrle_b200 import frames as fr
    dev = torch.device("cuda:0")
    FB, F = 2 << 20, 5
    whole = gen_run_mixed_pieces(0, F, dev, piece_bytes=FB)[: F * FB - 12345]
    bounds = fr.frame_bounds(whole.numel(), FB)
    ins = [whole[a:b] for a, b in bounds]
    fc = fr.FrameCodec(name, [b - a for a, b in bounds], device=dev, streams=3)
    fc.encode_async(ins)
    sizes = fc.finish_encode()
    codec = CODEC_BY_NAME[name]
    host = whole.cpu().numpy()
    make = ref_compress if ref_lib() is not None else oracle_compress
    for i, (a, b) in enumerate(bounds):
        want = make(codec, host[a:b])
        assert sizes[i] == len(want) and np.array_equal(fc.stream(i).cpu().numpy(), want), (name, i)
    blob = np.concatenate([fc.stream(i).cpu().numpy() for i in range(F)])
    parts = fr.split_concat(blob)
    assert [len(p) for p in parts] == sizes
    outs = [torch.zeros(FB + 128, dtype=torch.uint8, device=dev) for _ in range(F)]
    fc.decode_async(outs)
    fc.finish_decode()
    for i, (a, b) in enumerate(bounds):
        assert torch.equal(outs[i][: b - a], ins[i])
