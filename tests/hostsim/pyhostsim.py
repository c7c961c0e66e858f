"""TEST INFRASTRUCTURE ONLY: ctypes view of tests/hostsim/libhostsim.so (the CUDA engine's game-core headers compiled for
the CPU).  Same vector interface as oracle/pyoracle.OracleVector."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "..", "..", "drl-tetris_b200", "csrc")
_LIB = os.path.join(_HERE, "libhostsim.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "hostsim.cpp")] + [os.path.join(_SRC, f) for f in os.listdir(_SRC) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) > os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
                           "-I" + _SRC, "-I" + os.path.join(_HERE, "..", "..", "include"), "-x", "c++", os.path.join(_HERE, "hostsim.cpp"), "-o", _LIB])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, i32, u32, u8p, i32p, u32p, i8p = C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int8)
        L.hs_create.restype = vp
        L.hs_create.argtypes = [i32, i32, i32, u8p, u32, u32]
        L.hs_create_players.restype = vp
        L.hs_create_players.argtypes = [i32, i32, i32, i32, u8p, u32, u32]
        L.hs_destroy.argtypes = [vp]
        L.hs_reset.argtypes = [vp, i32p, i32, u32p]
        L.hs_step.argtypes = [vp, i32p, i32, u8p, u8p, u8p, i32, i32, i32, i8p, u8p, u8p]
        L.hs_observe.argtypes = [vp, i32p, i32, u8p, u8p, u8p]
        L.hs_scal.argtypes = [vp, i32p, i32, i32p]
        L.hs_enumerate.restype = i32
        L.hs_enumerate.argtypes = [vp, i32p, i32, u8p, i32, i32, i32, i32p, u8p, u8p]
        L.hs_expand.restype = i32
        L.hs_expand.argtypes = [vp, i32p, i32, u8p, i32, i32, i32, i32, i32, i32p, u8p, u8p, u8p, u8p, u8p, i8p, u8p, u8p, u8p]
        L.hs_get_game.argtypes = [vp, i32, vp, u8p]
        L.hs_set_game.argtypes = [vp, i32, vp, u8p]
        L.hs_copy_game.argtypes = [vp, i32, vp, i32]
        L.hs_enable_tiles.argtypes = [vp]
        L.hs_blob_bytes.restype = i32
        L.hs_blob_bytes.argtypes = [vp]
        L.hs_get_blob.argtypes = [vp, i32, vp]
        L.hs_set_blob.argtypes = [vp, i32, vp]
        L.hs_flags.restype = i32
        L.hs_flags.argtypes = [vp, i32]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class HostSimVector:
    def __init__(self, n_games, height=22, width=10, pieces=(0, 1, 2, 3, 4, 5, 6), piece_len=1024, hole_len=256, tiles=False, n_players=2):
        self.n, self.H, self.W, self.P = n_games, height, width, n_players
        self.map = np.array((list(pieces) * 7)[:7], np.uint8)
        self.h = lib().hs_create_players(n_games, n_players, height, width, _p(self.map, C.c_uint8), piece_len, hole_len)
        assert self.h, "geometry %dx%d not in b2t_geoms.h" % (height, width)
        if tiles:
            lib().hs_enable_tiles(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().hs_destroy(self.h)
            self.h = None

    def _idx(self, idx):
        if idx is None:
            return None, self.n
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        return idx, len(idx)

    def reset(self, seeds, idx=None):
        idx, n = self._idx(idx)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        lib().hs_reset(self.h, _p(idx, C.c_int32), n, _p(seeds, C.c_uint32))

    def step(self, player, keys, key_len, idx=None, time_ms=400, finalize=True):
        idx, n = self._idx(idx)
        player = np.ascontiguousarray(player, dtype=np.uint8)
        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        key_len = np.ascontiguousarray(key_len, dtype=np.uint8)
        reward = np.zeros(n, np.int8); done = np.zeros(n, np.uint8); lines = np.zeros((n, self.P), np.uint8)
        lib().hs_step(self.h, _p(idx, C.c_int32), n, _p(player, C.c_uint8), _p(keys, C.c_uint8), _p(key_len, C.c_uint8), keys.shape[1],
                      time_ms, int(finalize), _p(reward, C.c_int8), _p(done, C.c_uint8), _p(lines, C.c_uint8))
        return reward, done, lines

    def observe(self, idx=None):
        idx, n = self._idx(idx)
        field = np.zeros((n, self.P, self.H, self.W), np.uint8); vec = np.zeros((n, self.P, 19), np.uint8); pidx = np.zeros((n, self.P), np.uint8)
        lib().hs_observe(self.h, _p(idx, C.c_int32), n, _p(field, C.c_uint8), _p(vec, C.c_uint8), _p(pidx, C.c_uint8))
        return field, vec, pidx

    def scal(self, idx=None):
        idx, n = self._idx(idx)
        out = np.zeros((n, self.P, 8), np.int32)
        lib().hs_scal(self.h, _p(idx, C.c_int32), n, _p(out, C.c_int32))
        return out

    def enumerate(self, player, idx=None, bar_null_moves=True, kmax=64, lmax=32):
        idx, n = self._idx(idx)
        player = np.ascontiguousarray(player, dtype=np.uint8)
        K = np.zeros(n, np.int32); keys = np.zeros((n, kmax, lmax), np.uint8); klen = np.zeros((n, kmax), np.uint8)
        bad = lib().hs_enumerate(self.h, _p(idx, C.c_int32), n, _p(player, C.c_uint8), int(bar_null_moves), kmax, lmax,
                                 _p(K, C.c_int32), _p(keys, C.c_uint8), _p(klen, C.c_uint8))
        return K, keys, klen, None, bad

    def expand(self, player, idx=None, bar_null_moves=True, finalize=True, time_ms=400, kmax=64, lmax=32, want_obs=True):
        idx, n = self._idx(idx)
        player = np.ascontiguousarray(player, dtype=np.uint8)
        o = dict(K=np.zeros(n, np.int32), keys=np.zeros((n, kmax, lmax), np.uint8), key_len=np.zeros((n, kmax), np.uint8),
                 reward=np.zeros((n, kmax), np.int8), done=np.zeros((n, kmax), np.uint8), lines=np.zeros((n, kmax, self.P), np.uint8),
                 field=np.zeros((n, kmax, self.P, self.H, self.W), np.uint8), vec=np.zeros((n, kmax, self.P, 19), np.uint8),
                 piece_idx=np.zeros((n, kmax, self.P), np.uint8), dead=np.zeros((n, kmax, self.P), np.uint8))
        o["bad"] = lib().hs_expand(self.h, _p(idx, C.c_int32), n, _p(player, C.c_uint8), int(bar_null_moves), int(finalize), time_ms, kmax, lmax,
                                   _p(o["K"], C.c_int32), _p(o["keys"], C.c_uint8), _p(o["key_len"], C.c_uint8), _p(o["field"], C.c_uint8),
                                   _p(o["vec"], C.c_uint8), _p(o["piece_idx"], C.c_uint8), _p(o["reward"], C.c_int8), _p(o["done"], C.c_uint8),
                                   _p(o["lines"], C.c_uint8), _p(o["dead"], C.c_uint8))
        return o

    def get_game(self, game, want_field=True):
        from b200tetris import _lib as abi   # struct layout only (ctypes mirror of b2t_game_raw)
        raw = abi.GameRaw()
        field = np.zeros((2, self.H, self.W), np.uint8) if want_field else None
        lib().hs_get_game(self.h, int(game), C.byref(raw), _p(field, C.c_uint8))
        return raw, field

    def set_game(self, game, raw, field=None):
        field = None if field is None else np.ascontiguousarray(field, np.uint8)
        lib().hs_set_game(self.h, int(game), C.byref(raw), _p(field, C.c_uint8))

    def copy_game(self, dst, other, src):
        lib().hs_copy_game(self.h, int(dst), other.h, int(src))

    def get_blob(self, game):
        """opaque snapshot of one game (n_players != 2 simulator)"""
        buf = C.create_string_buffer(lib().hs_blob_bytes(self.h))
        lib().hs_get_blob(self.h, int(game), buf)
        return buf.raw

    def set_blob(self, game, blob):
        lib().hs_set_blob(self.h, int(game), C.create_string_buffer(blob, len(blob)))

    def flags(self, game):
        return lib().hs_flags(self.h, game)
