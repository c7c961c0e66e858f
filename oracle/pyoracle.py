"""TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_ref/libtetris_oracle.so (the plain-C restatement,
oracle/tetris_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libtetris_oracle.so")
VEC_LEN = 19
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        vp, i32, u8p, i32p, u32p, i8p = C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int8)
        L.tor_create.restype = vp
        L.tor_create.argtypes = [i32, i32, i32, i32, u8p]
        L.tor_destroy.argtypes = [vp]
        L.tor_reset.argtypes = [vp, i32p, i32, u32p]
        L.tor_step.argtypes = [vp, i32p, i32, u8p, u8p, u8p, i32, i32, i32, i8p, u8p, u8p]
        L.tor_observe.argtypes = [vp, i32p, i32, u8p, u8p, u8p]
        L.tor_raw.argtypes = [vp, i32p, i32, u8p, u8p, i32p]
        L.tor_hidden.argtypes = [vp, i32p, i32, i32p]
        L.tor_enumerate.restype = i32
        L.tor_enumerate.argtypes = [vp, i32p, i32, u8p, i32, i32, i32, i32p, u8p, u8p, i32p]
        L.tor_expand.restype = i32
        L.tor_expand.argtypes = [vp, i32p, i32, u8p, i32, i32, i32, i32, i32, i32p, u8p, u8p, u8p, u8p, u8p, i8p, u8p, u8p]
        L.tor_copy_game.argtypes = [vp, i32, vp, i32]
        L.tor_poke_dead.argtypes = [vp, i32, i32, i32]
        L.tor_poke_cell.argtypes = [vp, i32, i32, i32, i32, i32]
        L.tor_piece_stream.argtypes = [C.c_uint32, u8p, i32, u8p]
        L.tor_hole_stream.argtypes = [C.c_uint32, i32, i32, u8p]
        L.tor_piece_grids.argtypes = [u8p]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def piece_map(pieces):
    """tetris_environment.generate_pieces (tetris_environment.py:191-193): (p*7)[:7]"""
    return np.array((list(pieces) * 7)[:7], dtype=np.uint8)


class OracleVector:
    """A vector of independent games stepped by the C restatement."""

    def __init__(self, n_games, height=22, width=10, pieces=(0, 1, 2, 3, 4, 5, 6), n_players=2):
        self.n, self.H, self.W, self.P = n_games, height, width, n_players
        self.map = piece_map(pieces)
        self.h = lib().tor_create(n_games, n_players, height, width, _p(self.map, C.c_uint8))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            lib().tor_destroy(self.h)
            self.h = None

    def _idx(self, idx):
        if idx is None:
            return None, self.n
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        return idx, len(idx)

    def reset(self, seeds, idx=None):
        idx, n = self._idx(idx)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert len(seeds) == n
        lib().tor_reset(self.h, _p(idx, C.c_int32), n, _p(seeds, C.c_uint32))

    def step(self, player, keys, key_len, idx=None, time_ms=400, finalize=True):
        idx, n = self._idx(idx)
        player = np.ascontiguousarray(player, dtype=np.uint8)
        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        key_len = np.ascontiguousarray(key_len, dtype=np.uint8)
        assert keys.shape[0] == n and len(player) == n and len(key_len) == n
        reward = np.zeros(n, np.int8)
        done = np.zeros(n, np.uint8)
        lines = np.zeros((n, self.P), np.uint8)
        lib().tor_step(self.h, _p(idx, C.c_int32), n, _p(player, C.c_uint8), _p(keys, C.c_uint8), _p(key_len, C.c_uint8),
                       keys.shape[1], time_ms, int(finalize), _p(reward, C.c_int8), _p(done, C.c_uint8), _p(lines, C.c_uint8))
        return reward, done, lines

    def observe(self, idx=None):
        idx, n = self._idx(idx)
        field = np.zeros((n, self.P, self.H, self.W), np.uint8)
        vec = np.zeros((n, self.P, VEC_LEN), np.uint8)
        pidx = np.zeros((n, self.P), np.uint8)
        lib().tor_observe(self.h, _p(idx, C.c_int32), n, _p(field, C.c_uint8), _p(vec, C.c_uint8), _p(pidx, C.c_uint8))
        return field, vec, pidx

    def raw(self, idx=None):
        idx, n = self._idx(idx)
        field = np.zeros((n, self.P, self.H, self.W), np.uint8)
        grid = np.zeros((n, self.P, 4, 4), np.uint8)
        scal = np.zeros((n, self.P, 8), np.int32)
        lib().tor_raw(self.h, _p(idx, C.c_int32), n, _p(field, C.c_uint8), _p(grid, C.c_uint8), _p(scal, C.c_int32))
        return field, grid, scal

    def scal(self, idx=None):
        """[n,P,8] int32: x,y,inc_lines,combo_time(raw u16),combo_count,nextpiece,reward,dead"""
        return self.raw(idx)[2]

    def hidden(self, idx=None):
        idx, n = self._idx(idx)
        out = np.zeros((n, self.P, 16), np.int32)
        lib().tor_hidden(self.h, _p(idx, C.c_int32), n, _p(out, C.c_int32))
        return out

    def enumerate(self, player, idx=None, bar_null_moves=True, kmax=64, lmax=32):
        idx, n = self._idx(idx)
        player = np.ascontiguousarray(player, dtype=np.uint8)
        K = np.zeros(n, np.int32)
        Kraw = np.zeros(n, np.int32)
        keys = np.zeros((n, kmax, lmax), np.uint8)
        klen = np.zeros((n, kmax), np.uint8)
        bad = lib().tor_enumerate(self.h, _p(idx, C.c_int32), n, _p(player, C.c_uint8), int(bar_null_moves), kmax, lmax,
                                  _p(K, C.c_int32), _p(keys, C.c_uint8), _p(klen, C.c_uint8), _p(Kraw, C.c_int32))
        return K, keys, klen, Kraw, bad

    def expand(self, player, idx=None, bar_null_moves=True, finalize=True, time_ms=400, kmax=64, lmax=32, want_obs=True):
        """simulate_all_actions; padded [n,kmax,...] outputs as a dict."""
        idx, n = self._idx(idx)
        player = np.ascontiguousarray(player, dtype=np.uint8)
        o = dict(K=np.zeros(n, np.int32), keys=np.zeros((n, kmax, lmax), np.uint8), key_len=np.zeros((n, kmax), np.uint8),
                 reward=np.zeros((n, kmax), np.int8), done=np.zeros((n, kmax), np.uint8), lines=np.zeros((n, kmax, self.P), np.uint8))
        if want_obs:
            o.update(field=np.zeros((n, kmax, self.P, self.H, self.W), np.uint8), vec=np.zeros((n, kmax, self.P, VEC_LEN), np.uint8),
                     piece_idx=np.zeros((n, kmax, self.P), np.uint8))
        o["bad"] = lib().tor_expand(self.h, _p(idx, C.c_int32), n, _p(player, C.c_uint8), int(bar_null_moves), int(finalize), time_ms,
                                    kmax, lmax, _p(o["K"], C.c_int32), _p(o["keys"], C.c_uint8), _p(o["key_len"], C.c_uint8),
                                    _p(o.get("field"), C.c_uint8), _p(o.get("vec"), C.c_uint8), _p(o.get("piece_idx"), C.c_uint8),
                                    _p(o["reward"], C.c_int8), _p(o["done"], C.c_uint8), _p(o["lines"], C.c_uint8))
        return o

    def copy_game_from(self, dst, other, src):
        lib().tor_copy_game(self.h, dst, other.h, src)

    def poke_dead(self, game, player, dead):
        lib().tor_poke_dead(self.h, game, player, int(dead))

    def poke_cell(self, game, player, y, x, value):
        lib().tor_poke_cell(self.h, game, player, y, x, value)


def piece_stream(seed, pieces=(0, 1, 2, 3, 4, 5, 6), n=64):
    out = np.zeros(n, np.uint8)
    m = piece_map(pieces)
    lib().tor_piece_stream(int(seed) & 0xFFFFFFFF, _p(m, C.c_uint8), n, _p(out, C.c_uint8))
    return out


def hole_stream(seed, width=10, n=64):
    out = np.zeros(n, np.uint8)
    lib().tor_hole_stream(int(seed) & 0xFFFFFFFF, width, n, _p(out, C.c_uint8))
    return out


def piece_grids():
    out = np.zeros((7, 4, 4, 4), np.uint8)
    lib().tor_piece_grids(_p(out, C.c_uint8))
    return out
