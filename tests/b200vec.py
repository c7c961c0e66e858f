"""Test adapter: the CUDA engine behind the C ABI (b2t_*_host entry points, numpy in/out) with the same vector interface as
oracle/pyoracle.OracleVector, so tests/harness.py can diff them."""
import numpy as np

import harness  # noqa: F401  (sys.path)
from b200tetris import Engine


class B200Vector:
    def __init__(self, n_games, height=22, width=10, pieces=(0, 1, 2, 3, 4, 5, 6), bar_null_moves=True, n_players=2, **kw):
        self.n, self.H, self.W, self.P = n_games, height, width, n_players
        self.bar_null = bar_null_moves
        self.e = Engine(n_games, height, width, pieces, n_players=n_players, bar_null_moves=bar_null_moves, **kw)

    def reset(self, seeds, idx=None):
        self.e.reset(seeds, idx)

    def step(self, player, keys, key_len, idx=None, time_ms=400, finalize=True):
        return self.e.step(player, keys, key_len, idx, time_ms=time_ms, finalize=finalize)

    def observe(self, idx=None):
        return self.e.observe(idx)

    def scal(self, idx=None):
        return self.e.scal(idx)

    def enumerate(self, player, idx=None, bar_null_moves=True, kmax=64, lmax=32):
        assert bar_null_moves == self.bar_null
        K, keys, klen = self.e.enumerate(player, idx, kmax=kmax, lmax=lmax)
        return K, keys, klen, None, 0

    def expand(self, player, idx=None, bar_null_moves=True, finalize=True, time_ms=400, kmax=64, lmax=32, want_obs=True):
        """compact C-ABI result re-padded to the oracle's [n,kmax,...] layout"""
        assert bar_null_moves == self.bar_null
        n = self.n if idx is None else len(idx)
        o = self.e.expand(player, idx, finalize=finalize, time_ms=time_ms, lmax=lmax)
        out = dict(K=o["K"], keys=np.zeros((n, kmax, lmax), np.uint8), key_len=np.zeros((n, kmax), np.uint8),
                   reward=np.zeros((n, kmax), np.int8), done=np.zeros((n, kmax), np.uint8), lines=np.zeros((n, kmax, self.P), np.uint8),
                   field=np.zeros((n, kmax, self.P, self.H, self.W), np.uint8), vec=np.zeros((n, kmax, self.P, 19), np.uint8),
                   piece_idx=np.zeros((n, kmax, self.P), np.uint8), bad=0, total=o["total"], offsets=o["offsets"])
        assert o["offsets"][0] == 0 and np.array_equal(np.diff(o["offsets"]), o["K"])
        for i in range(n):
            a, b = o["offsets"][i], o["offsets"][i + 1]
            k = min(b - a, kmax)
            for name in ("keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
                out[name][i, :k] = o[name][a:a + k]
        return out
