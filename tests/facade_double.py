"""TEST INFRASTRUCTURE ONLY.  A CPU stand-in for b200tetris.Engine with the handful of methods the per-env facade
(b200tetris/tetris_env.py) calls, backed by tests/hostsim (the engine's own game-core headers compiled for the CPU).  It lets the
GPU-less container run the facade's Python — and, where /root/reference is present, the reference's unmodified front-end on
top of it — against the compiled reference.  The product never sees it: b200tetris.tetris_env defaults to the CUDA Engine, and
the `-m gpu` tests run the same cases on the real engine."""
import numpy as np

import harness  # noqa: F401
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import pyhostsim  # noqa: E402


class HostSimEngine:
    def __init__(self, n_games, height, width, pieces, n_players=2, bar_null_moves=True, time_ms_per_action=400, device=0, tiles=False):
        self.bar_null, self.device = bool(bar_null_moves), device
        self.v = pyhostsim.HostSimVector(n_games, height, width, pieces, tiles=tiles, n_players=n_players)
        self.n, self.H, self.W = n_games, height, width
        self.kmax, self.lmax = (256, 128) if width > 10 else (96, 64)
        self.default_ms = time_ms_per_action
        self._h = self.v.h

    def reset(self, seeds, idx=None):
        self.v.reset(seeds, idx)

    def step(self, player, keys, key_len, idx=None, time_ms=0, finalize=True, want_obs=False, want_scalars=False, **_):
        out = self.v.step(player, keys, key_len, idx, time_ms=time_ms if time_ms > 0 else self.default_ms, finalize=finalize)
        if want_obs:
            out = out + (self.v.observe(idx),)
        if want_scalars:
            out = out + (self.v.scal(idx),)
        return out

    def enumerate(self, player, idx=None, kmax=None, lmax=None, want_keys=True):
        K, keys, klen, _, bad = self.v.enumerate(player, idx, self.bar_null, kmax or self.kmax, lmax or self.lmax)
        assert bad == 0
        return K, keys, klen

    def observe(self, idx=None, want_scalars=False):
        o = self.v.observe(idx)
        return o + (self.v.scal(idx),) if want_scalars else o

    def expand(self, player, idx=None, finalize=True, time_ms=0, capacity=None, lmax=None, want_obs=True, want_scalars=False):
        """compact rows like b2t_expand_host, from the host sim's padded result"""
        kmax, lmax = self.kmax, lmax or self.lmax
        o = self.v.expand(player, idx, self.bar_null, finalize, time_ms if time_ms > 0 else self.default_ms, kmax, lmax)
        assert o["bad"] == 0
        K = o["K"]
        offs = np.concatenate([[0], np.cumsum(K)]).astype(np.int32)
        out = {"K": K, "offsets": offs, "total": int(offs[-1])}
        for name in ("keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
            out[name] = np.concatenate([o[name][i, :K[i]] for i in range(len(K))]) if len(K) else o[name][:0, 0]
        if want_scalars:   # only the `dead` column (7) is filled: it is the one the vector env reads from simulated outcomes
            dead = np.concatenate([o["dead"][i, :K[i]] for i in range(len(K))]) if len(K) else o["dead"][:0, 0]
            out["scalars"] = np.zeros((len(dead), dead.shape[1] if dead.ndim > 1 else 2, 8), np.int32)
            out["scalars"][:, :, 7] = dead
        return out

    def export_states(self, idx=None):
        return [self.export_state(i) for i in (range(self.n) if idx is None else idx)]

    def import_states(self, bufs, idx=None):
        for i, b in zip(range(self.n) if idx is None else idx, bufs):
            self.import_state(i, b)

    def export_state(self, game):
        if self.v.P != 2:
            return (self.v.get_blob(game),)
        raw, field = self.v.get_game(game)
        return (bytes(raw), field.tobytes())

    def import_state(self, game, buf):
        if self.v.P != 2:
            return self.v.set_blob(game, buf[0])
        from b200tetris import _lib as abi
        raw = abi.GameRaw.from_buffer_copy(buf[0])
        self.v.set_game(game, raw, np.frombuffer(buf[1], np.uint8).reshape(2, self.H, self.W).copy())

    def get_game(self, game, want_field=True):
        return self.v.get_game(game, want_field)

    def set_game(self, game, raw, field=None):
        self.v.set_game(game, raw, field)

    def copy_game(self, dst, src_engine, src):
        self.v.copy_game(dst, src_engine.v, src)


def install():
    """Point the facade (and the vector env) at the CPU double; returns the facade module with small arenas."""
    from b200tetris import tetris_env, vector_env
    vector_env._engine_factory = HostSimEngine
    tetris_env._engine_factory = HostSimEngine
    tetris_env._arenas.clear()
    tetris_env.set_arena_slots(64)
    return tetris_env


def uninstall():
    from b200tetris import tetris_env, vector_env
    from b200tetris.engine import Engine
    vector_env._engine_factory = Engine
    tetris_env._engine_factory = Engine
    tetris_env._arenas.clear()
    tetris_env.set_arena_slots(4096)
    tetris_env.set_tile_colours(False)
