"""CPU test of the stretch walk of the 8-bit LUT encoders (hypersonic-rle-kit_b200/csrc/hsrle_enc_lutwalk.cuh).  tests/sim/lutwalk_check.cpp
restates k_enc_lut_stretch / k_enc_lut_walk with host loops around the header's own decision functions and descriptor layout and
compares with the exact sequential automaton (hsrle_core.cuh: enc_eval_t): the emit rule `lw_emit` on every record, the table at
every stretch boundary (must be exact) and at every super-chunk start (a guess: may differ inside the uncertain head of a stretch).
The model is a test tool; it is not part of the product."""
import os
import re
import subprocess

import numpy as np
import pytest

from common import ROOT, gen_dct, gen_fuzz, gen_short_runs

SIM_DIR = os.path.join(ROOT, "tests", "sim")
CSRC = os.path.join(ROOT, "hypersonic-rle-kit_b200", "csrc")


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lutwalk") / "lutwalk_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-I", CSRC, "-o", exe, os.path.join(SIM_DIR, "lutwalk_check.cpp")], check=True)
    return exe


def _inputs():
    rng = np.random.default_rng(17)
    noise = (np.arange(300, dtype=np.uint32) * 7 % 251 + 1).astype(np.uint8)
    piece = np.concatenate([noise, np.full(3, 0x61, dtype=np.uint8)])
    return {"dct": gen_dct(6 << 20, seed=3),                                 # one dominant symbol: the walk's own territory
            "fuzz_few_symbols": gen_fuzz(rng, 1 << 20, max_sym=3, p_run=0.7),
            "fuzz_many_symbols": gen_fuzz(rng, 1 << 20, max_sym=40, p_run=0.3),
            "short_runs": gen_short_runs(2 << 20, seed=4, W=1),
            "uncertain_stretch": np.concatenate([np.tile(piece, 500), gen_dct(1 << 18, seed=5)])}   # > LW_BACK records without a certain one


@pytest.mark.parametrize("variant", [2, 3], ids=["3symlut", "7symlut"])
def test_walk_model_matches_the_sequential_automaton(tool, tmp_path, variant):
    for label, data in _inputs().items():
        path = str(tmp_path / f"{label}.bin")
        data.tofile(path)
        r = subprocess.run([tool, path, str(variant)], capture_output=True, text=True, timeout=300)
        got = {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", r.stdout)}
        assert r.returncode == 0 and got["emit_mismatch"] == 0 and got["boundary_mismatch"] == 0, (label, r.stdout, r.stderr)
        if label == "uncertain_stretch":
            assert got["overflow"] > 0                                      # the product then skips the walk (sc.lwOk stays 0)
        else:
            assert got["overflow"] == 0 and got["sc_checked"] > 0
            assert got["sc_guess_mismatch"] * 4 <= got["sc_checked"], (label, got)   # guesses: mostly right (exactness comes from E2's verify)
        if label == "dct":
            assert got["sc_guess_mismatch"] == 0 and got["boundaries"] < 16384
