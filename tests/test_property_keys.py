"""Property test (hypothesis): arbitrary key strings — any byte value, any length up to the engine's capacity, repeated hard
drops, keys after a lock — leave the game core (the headers the CUDA kernels are built from, run on the CPU by tests/hostsim)
and the oracle in identical states.  The GPU suite replays the same kind of strings through the kernels ("junk" policy)."""
import os
import sys

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import harness
import pyoracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))
import pyhostsim  # noqa: E402

key = st.one_of(st.integers(0, 10), st.integers(0, 255), st.sampled_from([7, 7, 5, 6, 8, 9, 10, 2, 4]))
strings = st.lists(st.lists(key, min_size=0, max_size=40), min_size=4, max_size=24)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seqs=strings, seed=st.integers(0, 65535), geo=st.sampled_from([(22, 10, (0, 1, 2, 3, 4, 5, 6)), (16, 8, (0, 6)), (40, 20, (4, 5))]),
       solo=st.booleans())
def test_arbitrary_key_strings(seqs, seed, geo, solo):
    H, W, pieces = geo
    a = pyoracle.OracleVector(1, H, W, pieces)
    b = pyhostsim.HostSimVector(1, H, W, pieces)
    a.reset([seed]); b.reset([seed])
    for t, s in enumerate(seqs):
        keys = np.zeros((1, 64), np.uint8)
        keys[0, : len(s)] = s
        ln = np.array([len(s)], np.uint8)
        pl = np.array([0 if solo else t & 1], np.uint8)
        fin = (t % 5) != 4                                   # every fifth call is make_action only (simulate, finalize=False)
        ra, rb = a.step(pl, keys, ln, finalize=fin), b.step(pl, keys, ln, finalize=fin)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        fa, fb = a.observe(), b.observe()
        assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[2], fb[2])
        if t:
            assert np.array_equal(fa[1], fb[1]) and np.array_equal(a.scal(), b.scal())
        ea, eb = a.enumerate(pl, kmax=256, lmax=128)[:3], b.enumerate(pl, kmax=256, lmax=128)[:3]
        for x, y in zip(ea, eb):
            assert np.array_equal(x, y)
        if ra[1][0]:
            a.reset([seed + t + 1]); b.reset([seed + t + 1])
