"""Arbitrary boards, not only those reachable by play: random fields with caves, overhangs, single-cell holes and even full rows
are installed in the oracle (cell by cell) and in the engine (b2t_set_game), and the placement lists, the world model in both
finalize modes and the following steps must agree.  This drives the under-overhang search (wall kicks, slides with rises) far
harder than self-play boards do.  CPU: the engine's game core compiled for the CPU (tests/hostsim); GPU: the CUDA engine."""
import sys
import os

import numpy as np
import pytest

import harness
import pyoracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))


def random_field(rng, H, W, style):
    f = np.zeros((H, W), np.uint8)
    top = rng.integers(3, H - 2)
    if style == 0:            # noise that gets denser with depth
        for y in range(top, H):
            f[y] = rng.random(W) < min(0.9, 0.35 + 0.5 * (y - top) / max(1, H - top))
    elif style == 1:          # solid stack with carved caves and tunnels
        f[top:] = 1
        for _ in range(rng.integers(3, 10)):
            y, x = rng.integers(top, H), rng.integers(0, W)
            for _ in range(rng.integers(2, 12)):
                f[y, x] = 0
                d = rng.integers(0, 4)
                y = int(np.clip(y + (d == 0) - (d == 1), top, H - 1)); x = int(np.clip(x + (d == 2) - (d == 3), 0, W - 1))
        f[top:, rng.integers(0, W)] = 0                      # a well down to the floor
    else:                     # overhang shelves
        for y in range(top, H, 2):
            a, b = sorted(rng.integers(0, W + 1, 2))
            f[y, a:b] = 1
        f[H - 1] = rng.random(W) < 0.8
    if rng.random() < 0.85:   # mostly without full rows (a live game never shows one), sometimes with
        for y in range(H):
            if f[y].all():
                f[y, rng.integers(0, W)] = 0
    f[:3] = 0                 # spawn area free: the piece on the board must fit
    return f


def run(make_engine, H, W, pieces, n, rounds, seed):
    rng = np.random.default_rng(seed)
    kmax, lmax = (256, 128) if W > 10 else (96, 64)
    o = pyoracle.OracleVector(n, H, W, pieces)
    e = make_engine(n, H, W, pieces)
    totals = dict(K=0, finesse=0)
    for r in range(rounds):
        seeds = rng.integers(0, 1 << 16, n).astype(np.uint32)
        o.reset(seeds); e.reset(seeds)
        for g in range(n):
            fields = np.stack([random_field(rng, H, W, int(rng.integers(0, 3))) for _ in range(2)])
            for p in range(2):
                cur = o.raw([g])[0][0, p]
                for y, x in np.argwhere((cur > 0) != (fields[p] > 0)):
                    o.poke_cell(g, p, int(y), int(x), int(fields[p][y, x]))
            raw, _ = e.get_game(g)
            e.set_game(g, raw, fields)
        for x, y in zip(o.observe(), e.observe()):
            assert np.array_equal(x, y)
        for t in range(3):
            player = ((np.arange(n) + t + r) & 1).astype(np.uint8)
            for fin in (True, False):
                a = o.expand(player, finalize=fin, kmax=kmax, lmax=lmax)
                b = e.expand(player, finalize=fin, kmax=kmax, lmax=lmax)
                assert a["bad"] == 0
                for k in ("K", "keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
                    assert np.array_equal(a[k], b[k]), (r, t, fin, k, np.argwhere(a[k] != b[k])[:4])
            K, keys, klen = a["K"], a["keys"], a["key_len"]
            totals["K"] += int(K.sum()); totals["finesse"] += int(((keys == 5) | (keys == 6)).any(-1).sum())
            sel = rng.integers(0, 1 << 30, n) % K
            ra = o.step(player, keys[np.arange(n), sel], klen[np.arange(n), sel])
            rb = e.step(player, keys[np.arange(n), sel], klen[np.arange(n), sel])
            for x, y in zip(ra, rb[:3]):
                assert np.array_equal(x, y), (r, t)
            for x, y in zip(o.observe(), e.observe()):
                assert np.array_equal(x, y), (r, t)
    return totals


CASES = [(22, 10, (0, 1, 2, 3, 4, 5, 6), 24, 4), (16, 8, (0, 6), 16, 3), (25, 12, (0, 1, 2, 3, 4, 5, 6), 8, 2), (40, 20, (0, 1, 2, 3, 4, 5, 6), 4, 2),
         (10, 6, (2, 3, 4, 5), 16, 3)]


@pytest.mark.parametrize("H,W,pieces,n,rounds", CASES)
def test_random_boards_host_core(H, W, pieces, n, rounds):
    import pyhostsim

    class V(pyhostsim.HostSimVector):
        def expand(self, player, finalize=True, kmax=64, lmax=32, **kw):
            return super().expand(player, None, True, finalize, 400, kmax, lmax)

        def step(self, player, keys, klen):
            return super().step(player, keys, klen)
    tot = run(lambda n_, H_, W_, p_: V(n_, H_, W_, p_), H, W, pieces, n, rounds, 11)
    assert tot["finesse"] > 0.02 * tot["K"]          # the boards do offer under-overhang placements


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,pieces,n,rounds", [(h, w, p, 4 * n, r) for h, w, p, n, r in CASES])
def test_random_boards_gpu(H, W, pieces, n, rounds):
    from b200vec import B200Vector

    class V(B200Vector):
        def get_game(self, g):
            return self.e.get_game(g)

        def set_game(self, g, raw, field):
            self.e.set_game(g, raw, field)
    tot = run(lambda n_, H_, W_, p_: V(n_, H_, W_, p_), H, W, pieces, n, rounds, 12)
    assert tot["finesse"] > 0.02 * tot["K"]
