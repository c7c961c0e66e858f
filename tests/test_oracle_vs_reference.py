"""Pins the plain-C restatement (oracle/tetris_oracle.c) to the compiled, UNMODIFIED reference backend
(oracle/_ref/B, built from /root/reference by oracle/Makefile) — every observable and the hidden timers, after
every call, under four policies and five geometries.  The reference ships no tests of its own."""
import numpy as np
import pytest

import harness
import pyoracle
import refdrive
import refmod

pytestmark = pytest.mark.skipif(not refmod.available("B"), reason="oracle/_ref/B not built (make -C oracle ref)")

CASES = [
    # kind, H, W, pieces, games, steps, solo_period
    ("random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 12, 260, 0),
    ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 6, 420, 0),
    ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 300, 23),   # same side repeats -> lockdown / gravity paths
    ("junk", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 220, 11),
    ("sventon", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 160, 0),
    ("press_key", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 700, 0),   # action_type "press_key": gravity / lockdown do the placing
    ("press_key", 22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 500, 40),
    ("press_key", 16, 8, (0, 6), 4, 400, 0),
    ("random", 16, 8, (0, 6), 8, 200, 0),
    ("greedy", 16, 8, (0, 6), 6, 300, 0),
    ("random", 40, 20, (0, 1, 2, 3, 4, 5, 6), 3, 200, 0),
    ("greedy", 40, 20, (0, 1, 2, 3, 4, 5, 6), 2, 160, 0),
    ("greedy", 25, 12, (0, 1, 2, 3, 4, 5, 6), 4, 260, 0),    # W > 10 on a field the engine keeps in 32-bit columns
    ("junk", 25, 12, (0, 1, 2, 3, 4, 5, 6), 4, 160, 9),
    ("random", 10, 6, (2, 3), 8, 120, 0),                    # only_zs, empty placement lists
    ("greedy", 10, 6, (2, 3, 4, 5), 6, 200, 0),
    ("junk", 12, 5, (4,), 4, 120, 5),
]


@pytest.mark.parametrize("kind,H,W,pieces,n,steps,solo", CASES)
def test_restatement_matches_compiled_reference(kind, H, W, pieces, n, steps, solo):
    kmax, lmax = (256, 128) if W > 10 else (128, 64)
    lead = pyoracle.OracleVector(n, H, W, pieces)
    ref = refdrive.RefVector(n, H, W, pieces)
    stats = harness.rollout_pair(lead, ref, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, solo_period=solo),
                                 compare_hidden=True)
    assert stats["steps"] == n * steps
    if kind == "greedy" and W >= 8:
        assert stats["lines"] > 0 and stats["garbage"] > 0 and stats["combos"] > 0, stats


TIMINGS = [  # kind, time_elapsed_each_action, games, steps: gravity tiers (every > 3000 ms: -10 / -5 / -2 / -1, DropDelay.cpp:3-21),
    ("greedy", 3100, 4, 420),     # lockdown and garbage / combo timers at other cadences than the default 400 ms
    ("press_key", 3100, 6, 300),
    ("press_key", 1000, 6, 300),
    ("greedy", 50, 4, 200),
    ("junk", 7, 6, 200),
]


@pytest.mark.parametrize("kind,time_ms,n,steps", TIMINGS)
def test_other_time_steps_match_compiled_reference(kind, time_ms, n, steps):
    lead = pyoracle.OracleVector(n)
    ref = refdrive.RefVector(n)
    stats = harness.rollout_pair(lead, ref, steps, harness.Policy(kind, kmax=128, lmax=64), time_ms=time_ms, compare_hidden=True)
    assert stats["steps"] == n * steps
    if kind == "greedy" and time_ms > 3000:
        assert int(lead.hidden()[..., 1].min()) < 1000 or stats["dones"] > 0     # the drop delay did shrink


def test_rng_streams_match_reference():
    """Piece and hole streams of the restatement vs the compiled reference for edge seeds."""
    for seed in (0, 7, 1000, 32767, 32768, 40000, 65535, 66536, 123456):
        ref = refdrive.RefVector(1)
        ref.reset([seed])
        want_cur = int(np.array(ref.h[0].states[0].piece).max())
        got = pyoracle.piece_stream(seed, n=40)
        code = {1: 5, 2: 4, 3: 1, 4: 0, 5: 2, 6: 6, 7: 3}
        assert code[want_cur] == got[0]
        assert int(ref.h[0].states[0].nextpiece[0]) == got[1]
        # the idle player redraws every step too, so 38 finish_actions expose stream entries 2..39
        keys = np.zeros((1, 4), np.uint8)
        for t in range(2, 40):
            ref.step([0], keys, [1])
            if ref.env_done[0]:
                break
            assert int(ref.h[0].states[0].nextpiece[0]) == got[t]


def test_two_build_variants_agree():
    """As-shipped -O0 build (A) vs patched -O2 build (B): identical trajectories (checked in a subprocess because
    pybind11 registers the module's types once per process)."""
    import subprocess, sys, os
    if not refmod.available("A"):
        pytest.skip("oracle/_ref/A not built")
    code = r'''
import sys, numpy as np
sys.path[:0] = [%r, %r]
import harness, refdrive
n = 4
v = refdrive.RefVector(n, variant=sys.argv[1])
pol = harness.Policy("junk", kmax=128, lmax=64, solo_period=9)
gids = np.arange(n); v.reset(harness.seeds_for(gids, 0))
out = []
for t in range(150):
    pl = pol.players(n, t, gids)
    keys, ln, _ = pol.choose(v, t, gids, pl)
    r = v.step(pl, keys, ln)
    out.append(harness.digest(*r, *v.raw(), v.hidden()[..., :13]))
    d = np.nonzero(r[1])[0]
    if len(d): v.reset(harness.seeds_for(d, t), d)
print(harness.digest(np.array(out)))
''' % (os.path.join(harness.ROOT, "tests"), os.path.join(harness.ROOT, "oracle"))
    outs = [subprocess.check_output([sys.executable, "-c", code, var], text=True).strip() for var in ("A", "B")]
    assert outs[0] == outs[1]


N_PLAYERS = [  # n_players, kind, H, W, games, steps: PythonHandle with 1, 3 and 4 players (PythonHandle.cpp:5-25,124-136,149-188 — the
    (1, "random", 22, 10, 6, 200),     # float split of sent lines among the others, solo rounds that end when the one player dies)
    (1, "greedy", 22, 10, 4, 300),
    (3, "random", 22, 10, 6, 260),
    (3, "greedy", 22, 10, 4, 420),
    (3, "junk", 22, 10, 6, 200),
    (4, "greedy", 22, 10, 3, 420),
    (4, "random", 22, 10, 16, 300),    # many rounds end with a late player to move: the reward rule's Python index (states[1 - player])
    (3, "random", 16, 8, 16, 300),
    (4, "press_key", 22, 10, 4, 500),
    (3, "greedy", 16, 8, 4, 300),
]


@pytest.mark.parametrize("P,kind,H,W,n,steps", N_PLAYERS)
def test_other_player_counts_match_compiled_reference(P, kind, H, W, n, steps):
    """n_players != 2: not built in the CUDA engine, but the restatement follows the reference there too."""
    pieces = (0, 1, 2, 3, 4, 5, 6) if W >= 10 else (0, 6)
    lead = pyoracle.OracleVector(n, H, W, pieces, n_players=P)
    ref = refdrive.RefVector(n, H, W, pieces, n_players=P)
    stats = harness.rollout_pair(lead, ref, steps, harness.Policy(kind, kmax=128, lmax=64, n_players=P), compare_hidden=True)
    assert stats["steps"] == n * steps
    if kind == "greedy" and P > 1:
        assert stats["lines"] > 0 and stats["garbage"] > 0, stats
