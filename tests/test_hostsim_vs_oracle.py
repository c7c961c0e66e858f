"""The CUDA engine's game core (drl-tetris_b200/csrc/b2t_{core,enum,game,rng}.cuh), compiled for the CPU by tests/hostsim,
must agree with the oracle call by call.  This is the GPU-less half of the parity gate: it checks the bit-board rules, the
HBM word layout and the stream tables; tests/test_gpu_*.py check the kernels that wrap the same headers."""
import glob
import os

import numpy as np
import pytest

import harness
import pyoracle

import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))
import pyhostsim  # noqa: E402

CASES = [
    ("random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 32, 400, 0),
    ("random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 256, 60, 0),   # step 20 / game 90 holds a 33-key placement (> the 32 the survey saw)
    ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 12, 700, 0),
    ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 500, 23),
    ("junk", 22, 10, (0, 1, 2, 3, 4, 5, 6), 16, 400, 11),
    ("sventon", 22, 10, (0, 1, 2, 3, 4, 5, 6), 16, 200, 0),
    ("press_key", 22, 10, (0, 1, 2, 3, 4, 5, 6), 16, 700, 0),
    ("press_key", 16, 8, (0, 6), 8, 500, 30),
    ("greedy", 25, 12, (0, 1, 2, 3, 4, 5, 6), 6, 260, 0),      # tallest field on 32-bit columns, W > 10
    ("junk", 25, 12, (0, 1, 2, 3, 4, 5, 6), 8, 200, 9),
    ("random", 16, 8, (0, 6), 16, 300, 0),
    ("greedy", 16, 8, (0, 6), 8, 500, 0),
    ("random", 40, 20, (0, 1, 2, 3, 4, 5, 6), 4, 300, 0),
    ("greedy", 40, 20, (0, 1, 2, 3, 4, 5, 6), 3, 250, 0),
    ("junk", 40, 20, (0, 1, 2, 3, 4, 5, 6), 4, 200, 7),
    ("random", 10, 6, (2, 3), 16, 200, 0),
    ("greedy", 10, 6, (2, 3, 4, 5), 8, 300, 0),
    ("junk", 12, 5, (4,), 8, 200, 5),
]


@pytest.mark.parametrize("kind,H,W,pieces,n,steps,solo", CASES)
def test_core_matches_oracle(kind, H, W, pieces, n, steps, solo):
    kmax, lmax = (256, 64) if W > 10 else (96, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces)
    sim = pyhostsim.HostSimVector(n, H, W, pieces)
    stats = harness.rollout_pair(lead, sim, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, solo_period=solo))
    assert stats["steps"] == n * steps
    assert all(sim.flags(i) == 0 for i in range(n)), "garbage-queue / stream-table overflow flag raised"


TIMINGS = [  # kind, time_elapsed_each_action, games, steps: gravity tiers (every > 3000 ms: -10 / -5 / -2 / -1, DropDelay.cpp:3-21),
    ("greedy", 3100, 4, 420),     # lockdown and garbage / combo timers at other cadences than the default 400 ms
    ("press_key", 3100, 6, 300),
    ("press_key", 1000, 6, 300),
    ("greedy", 50, 4, 200),
    ("junk", 7, 6, 200),
]


@pytest.mark.parametrize("kind,time_ms,n,steps", TIMINGS)
def test_other_time_steps(kind, time_ms, n, steps):
    lead = pyoracle.OracleVector(n)
    sim = pyhostsim.HostSimVector(n)
    stats = harness.rollout_pair(lead, sim, steps, harness.Policy(kind, kmax=96, lmax=32), time_ms=time_ms)
    assert stats["steps"] == n * steps


@pytest.mark.parametrize("finalize", [True, False])
@pytest.mark.parametrize("H,W,pieces", [(22, 10, (0, 1, 2, 3, 4, 5, 6)), (16, 8, (0, 6)), (40, 20, (0, 1, 2, 3, 4, 5, 6))])
def test_world_model_matches_oracle(H, W, pieces, finalize):
    """simulate_all_actions on boards reached by greedy play (fields non-empty, garbage pending, combos running)."""
    n = 6
    kmax, lmax = (128, 64) if W > 10 else (64, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces)
    sim = pyhostsim.HostSimVector(n, H, W, pieces)

    def check(t, player, *_):
        if t % 7 != 3:
            return
        a = lead.expand(player, finalize=finalize, kmax=kmax, lmax=lmax)
        b = sim.expand(player, finalize=finalize, kmax=kmax, lmax=lmax)
        assert a["bad"] == 0 and b["bad"] == 0
        for k in ("K", "keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
            x, y = a[k], b[k]
            if k == "vec" and t < 2:
                continue
            harness.assert_same("expand." + k, t, x, y)

    harness.rollout_pair(lead, sim, 120, harness.Policy("greedy", kmax=kmax, lmax=lmax), on_step=lambda t, pl, *r: check(t + 1, harness.Policy("greedy").players(n, t + 1, np.arange(n))))


def test_bar_null_moves_false_prepends_null_action():
    lead = pyoracle.OracleVector(4)
    sim = pyhostsim.HostSimVector(4)
    s = harness.seeds_for(np.arange(4), np.zeros(4, int))
    lead.reset(s); sim.reset(s)
    pl = np.array([0, 1, 0, 1], np.uint8)
    a = lead.enumerate(pl, bar_null_moves=False)
    b = sim.enumerate(pl, bar_null_moves=False)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    assert (a[2][:, 0] == 1).all() and (a[1][:, 0, 0] == 0).all()


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_core_replays_reference_golden(path):
    g = np.load(path)
    eng = pyhostsim.HostSimVector(int(g["n"]), int(g["H"]), int(g["W"]), tuple(int(x) for x in g["pieces"]),
                                  n_players=int(g["n_players"]) if "n_players" in g else 2)
    harness.replay_golden(eng, g)


def test_stream_tables_match_oracle():
    """piece/hole stream rows (b2t_rng.cuh) against the oracle's std::mt19937 restatement, incl. sign-extended seeds."""
    import ctypes as C
    for pieces in ((0, 1, 2, 3, 4, 5, 6), (0, 6), (2, 3), (2, 3, 4, 5)):
        sim = pyhostsim.HostSimVector(1, 22, 10, pieces, piece_len=700, hole_len=64)
        for seed in (0, 1, 1000, 32767, 32768, 40000, 65535):
            sim.reset([seed])
            want = pyoracle.piece_stream(seed, pieces, 700)
            keys = np.zeros((1, 32), np.uint8)
            got = [int(sim.observe()[2][0, 0]), int(np.argmax(sim.observe()[1][0, 0, 5:12]))]
            assert got == [want[0], want[1]]


N_PLAYERS = [  # n_players, kind, H, W, pieces, games, steps — the rules gamen_kernel runs for handles with 1, 3 or 4 players
    (1, "random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 24, 300),
    (1, "greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 400),
    (3, "random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 32, 300),
    (3, "greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 600),
    (3, "junk", 22, 10, (0, 1, 2, 3, 4, 5, 6), 16, 300),
    (3, "press_key", 22, 10, (0, 1, 2, 3, 4, 5, 6), 12, 600),
    (4, "greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 6, 600),
    (4, "random", 16, 8, (0, 6), 24, 300),
    (3, "greedy", 40, 20, (0, 1, 2, 3, 4, 5, 6), 2, 200),
    (4, "junk", 25, 12, (0, 1, 2, 3, 4, 5, 6), 6, 200),
]


@pytest.mark.parametrize("P,kind,H,W,pieces,n,steps", N_PLAYERS)
def test_core_matches_oracle_for_other_player_counts(P, kind, H, W, pieces, n, steps):
    """distributeLines' float split among all the others, solo rounds, the reward rule's Python index"""
    kmax, lmax = (256, 64) if W > 10 else (96, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces, n_players=P)
    sim = pyhostsim.HostSimVector(n, H, W, pieces, n_players=P)
    stats = harness.rollout_pair(lead, sim, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, n_players=P))
    assert stats["steps"] == n * steps
    assert all(sim.flags(i) == 0 for i in range(n))
    if kind == "greedy" and P > 1 and W >= 8:
        assert stats["lines"] > 0 and stats["garbage"] > 0, stats
    player = (np.arange(n) % P).astype(np.uint8)
    for fin in (True, False):
        a = lead.expand(player, finalize=fin, kmax=kmax, lmax=lmax)
        b = sim.expand(player, finalize=fin, kmax=kmax, lmax=lmax)
        for k in ("K", "keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
            assert np.array_equal(a[k], b[k]), (fin, k)
