"""The C restatement replays every committed golden trajectory (recorded from the compiled reference by
tests/golden/make_golden.py) bit-exactly.  The same fixtures are replayed by the CUDA engine in tests/test_gpu_*.py."""
import glob
import os

import numpy as np
import pytest

import harness
import pyoracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_replays_golden(path):
    g = np.load(path)
    eng = pyoracle.OracleVector(int(g["n"]), int(g["H"]), int(g["W"]), tuple(int(x) for x in g["pieces"]),
                                n_players=int(g["n_players"]) if "n_players" in g else 2)
    assert harness.replay_golden(eng, g) == int(g["n"]) * int(g["T"])


def test_goldens_present():
    assert len(GOLDEN) >= 9
