"""Engine: a vector of 2-player Tetris games resident on one B200, stepped through the C ABI.

Two faces over the same handle:
  * host arrays (numpy) -> the b2t_*_host entry points (pinned staging inside the library); this is what the
    reference-shaped front-end (vector_env.py) and the parity tests use;
  * device tensors (torch, uint8/int32 on the handle's GPU) -> the asynchronous b2t_* entry points on
    torch's current stream; this is the throughput path (bench.py).
torch is only memory + streams here."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import B2TError, VEC_LEN


def piece_map(pieces):
    """tetris_environment.generate_pieces (tetris_environment.py:191-193)"""
    return np.array((list(pieces) * 7)[:7], dtype=np.uint8)


def _hp(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _np(a, dtype, n=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if n is not None and a.shape[0] != n:
        raise ValueError("expected %d rows, got %d" % (n, a.shape[0]))
    return a


def host_array(shape, dtype, write_combined=False):
    """numpy array over pinned host memory from b2t_host_alloc (freed when the array is collected).  write_combined: for inputs the
    host only writes — the device reads them faster, the host reads them very slowly."""
    lib = _lib.load()
    shape = tuple(int(x) for x in (shape if hasattr(shape, "__len__") else (shape,)))
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = lib.b2t_host_alloc(max(nbytes, 1), int(bool(write_combined)), C.byref(p))
    if rc != 0:
        raise B2TError(rc, "b2t_host_alloc failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.b2t_host_free, p)
    return arr


class Engine:
    def __init__(self, n_games, height=22, width=10, pieces=(0, 1, 2, 3, 4, 5, 6), n_players=2, bar_null_moves=True,
                 time_ms_per_action=400, piece_stream_len=0, hole_stream_len=0, device=0, tiles=False, seed=0):
        """tiles: keep the raw cell values 1..8 of states[p].field (B2T_FLAG_TILES) — get_game / set_game then carry colours"""
        self._h = None
        self.lib = _lib.load()
        self.n, self.H, self.W, self.P = int(n_games), int(height), int(width), int(n_players)
        self.device = int(device)
        cfg = _lib.Config()
        cfg.n_games, cfg.n_players, cfg.height, cfg.width = self.n, self.P, self.H, self.W
        for i, v in enumerate(piece_map(pieces)):
            cfg.piece_map[i] = int(v)
        cfg.bar_null_moves = 1 if bar_null_moves else 0
        cfg.time_ms_per_action = int(time_ms_per_action)
        cfg.piece_stream_len, cfg.hole_stream_len = int(piece_stream_len), int(hole_stream_len)
        cfg.device, cfg.flags = self.device, (_lib.FLAG_TILES if tiles else 0)
        cfg.seed = int(seed) & 0xFFFFFFFF
        h = C.c_void_p()
        rc = self.lib.b2t_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B2TError(rc, (self.lib.b2t_last_error(None) or b"").decode())
        self._h = h
        info = _lib.Info()
        self._ck(self.lib.b2t_get_info(self._h, C.byref(info)))
        self.kmax, self.lmax, self.state_bytes, self.obs_bytes = info.kmax, info.lmax, info.state_bytes, info.obs_bytes

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise B2TError(rc, (self.lib.b2t_last_error(self._h) or b"").decode())

    def close(self):
        if self._h is not None:
            self.lib.b2t_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _idx(self, idx):
        if idx is None:
            return None, self.n
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        if idx.ndim != 1:
            raise ValueError("game ids must be a 1-d list")
        if len(idx) and (int(idx.min()) < 0 or int(idx.max()) >= self.n):   # the reference's list of envs raises the same way
            raise IndexError("game id outside [0, %d)" % self.n)
        return idx, len(idx)

    # ------------------------------------------------------------------ host-array face (b2t_*_host)
    def reset(self, seeds, idx=None):
        idx, n = self._idx(idx)
        seeds = _np(seeds, np.uint32, n)
        self._ck(self.lib.b2t_reset_host(self._h, _hp(idx, C.c_int32), n, _hp(seeds, C.c_uint32)))

    def step(self, player, keys=None, key_len=None, idx=None, time_ms=0, finalize=True, placement=None, reset_seeds=None,
             want_obs=False, want_scalars=False, packed_keys=False, auto_reset=False):
        """perform_action for the listed games.  Returns (reward_base i8[n], done u8[n], lines u8[n,2]) and, with
        want_obs, also (field, vec, piece_idx) after the step; with want_scalars the raw State scalars [n,2,8]."""
        idx, n = self._idx(idx)
        player = _np(player, np.uint8, n)
        keys = _np(keys, np.uint8, n)
        key_len = _np(key_len, np.uint8, n)
        placement = _np(placement, np.int32, n)
        reset_seeds = _np(reset_seeds, np.uint32, n)
        lmax = (keys.shape[1] * (2 if packed_keys else 1)) if keys is not None else 0
        reward = np.empty(n, np.int8); done = np.empty(n, np.uint8); lines = np.empty((n, self.P), np.uint8)
        field = np.empty((n, self.P, self.H, self.W), np.uint8) if want_obs else None
        vec = np.empty((n, self.P, VEC_LEN), np.uint8) if want_obs else None
        pidx = np.empty((n, self.P), np.uint8) if want_obs else None
        scal = np.empty((n, self.P, 8), np.int32) if want_scalars else None
        self._ck(self.lib.b2t_step_host(self._h, _hp(idx, C.c_int32), n, _hp(player, C.c_uint8), _hp(keys, C.c_uint8),
                                        _hp(key_len, C.c_uint8), lmax, _hp(placement, C.c_int32), int(bool(finalize)), int(time_ms),
                                        _hp(reset_seeds, C.c_uint32), _hp(reward, C.c_int8), _hp(done, C.c_uint8), _hp(lines, C.c_uint8),
                                        _hp(field, C.c_uint8), _hp(vec, C.c_uint8), _hp(pidx, C.c_uint8), _hp(scal, C.c_int32),
                                        int(bool(packed_keys)), int(bool(auto_reset))))
        out = (reward, done, lines)
        if want_obs:
            out = out + ((field, vec, pidx),)
        if want_scalars:
            out = out + (scal,)
        return out

    def observe(self, idx=None, want_scalars=False):
        idx, n = self._idx(idx)
        field = np.empty((n, self.P, self.H, self.W), np.uint8); vec = np.empty((n, self.P, VEC_LEN), np.uint8); pidx = np.empty((n, self.P), np.uint8)
        scal = np.empty((n, self.P, 8), np.int32) if want_scalars else None
        self._ck(self.lib.b2t_observe_host(self._h, _hp(idx, C.c_int32), n, _hp(field, C.c_uint8), _hp(vec, C.c_uint8),
                                           _hp(pidx, C.c_uint8), _hp(scal, C.c_int32)))
        return (field, vec, pidx, scal) if want_scalars else (field, vec, pidx)

    def scal(self, idx=None):
        idx, n = self._idx(idx)
        scal = np.empty((n, self.P, 8), np.int32)
        self._ck(self.lib.b2t_observe_host(self._h, _hp(idx, C.c_int32), n, None, None, None, _hp(scal, C.c_int32)))
        return scal

    def enumerate(self, player, idx=None, kmax=None, lmax=None, want_keys=True):
        """get_actions: K[n] and (want_keys) keys u8[n,kmax,lmax], key_len u8[n,kmax]"""
        idx, n = self._idx(idx)
        player = _np(player, np.uint8, n)
        kmax = kmax or self.kmax
        lmax = lmax or self.lmax
        K = np.empty(n, np.int32)
        keys = np.empty((n, kmax, lmax), np.uint8) if want_keys else None
        klen = np.empty((n, kmax), np.uint8) if want_keys else None
        self._ck(self.lib.b2t_enumerate_host(self._h, _hp(idx, C.c_int32), n, _hp(player, C.c_uint8), _hp(K, C.c_int32),
                                             _hp(keys, C.c_uint8), _hp(klen, C.c_uint8), kmax, lmax))
        return K, keys, klen

    def expand(self, player, idx=None, finalize=True, time_ms=0, capacity=None, lmax=None, want_obs=True, want_scalars=False):
        """simulate_all_actions, compact rows.  Returns dict(K, offsets, total, keys, key_len, reward, done, lines, field, vec, piece_idx)
        and, with want_scalars, scalars i32[capacity,2,8] (the raw State views of every simulated outcome, incl. `dead`)."""
        idx, n = self._idx(idx)
        player = _np(player, np.uint8, n)
        lmax = lmax or self.lmax
        if capacity is None:
            capacity = n * self.kmax
        K = np.empty(n, np.int32); offs = np.empty(n + 1, np.int32); total = C.c_int32(0)
        o = dict(keys=np.zeros((capacity, lmax), np.uint8), key_len=np.zeros(capacity, np.uint8), reward=np.zeros(capacity, np.int8),
                 done=np.zeros(capacity, np.uint8), lines=np.zeros((capacity, self.P), np.uint8))
        if want_obs:
            o.update(field=np.zeros((capacity, self.P, self.H, self.W), np.uint8), vec=np.zeros((capacity, self.P, VEC_LEN), np.uint8),
                     piece_idx=np.zeros((capacity, self.P), np.uint8))
        if want_scalars:
            o["scalars"] = np.zeros((capacity, self.P, 8), np.int32)
        self._ck(self.lib.b2t_expand_host(self._h, _hp(idx, C.c_int32), n, _hp(player, C.c_uint8), int(bool(finalize)), int(time_ms), capacity,
                                          _hp(K, C.c_int32), _hp(offs, C.c_int32), C.byref(total), _hp(o["keys"], C.c_uint8),
                                          _hp(o["key_len"], C.c_uint8), lmax, _hp(o["reward"], C.c_int8), _hp(o["done"], C.c_uint8),
                                          _hp(o["lines"], C.c_uint8), _hp(o.get("field"), C.c_uint8), _hp(o.get("vec"), C.c_uint8),
                                          _hp(o.get("piece_idx"), C.c_uint8), _hp(o.get("scalars"), C.c_int32)))
        o.update(K=K, offsets=offs, total=int(total.value))
        return o

    def snapshot(self, slot=0):
        self._ck(self.lib.b2t_snapshot(self._h, slot, None))

    def restore(self, slot=0):
        self._ck(self.lib.b2t_restore(self._h, slot, None))

    def export_state(self, game):
        buf = np.empty(self.state_bytes, np.uint8)
        self._ck(self.lib.b2t_export_state(self._h, int(game), buf.ctypes.data_as(C.c_void_p), None))
        return buf

    def import_state(self, game, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        assert buf.size == self.state_bytes
        self._ck(self.lib.b2t_import_state(self._h, int(game), buf.ctypes.data_as(C.c_void_p), None))

    def export_states(self, idx=None):
        """packed states of the listed games, u8 [n, state_bytes], one call"""
        idx, n = self._idx(idx)
        buf = np.empty((n, self.state_bytes), np.uint8)
        self._ck(self.lib.b2t_export_states(self._h, _hp(idx, C.c_int32), n, buf.ctypes.data_as(C.c_void_p)))
        return buf

    def import_states(self, buf, idx=None):
        idx, n = self._idx(idx)
        buf = np.ascontiguousarray(buf, np.uint8)
        assert buf.shape == (n, self.state_bytes)
        self._ck(self.lib.b2t_import_states(self._h, _hp(idx, C.c_int32), n, buf.ctypes.data_as(C.c_void_p)))

    def get_game(self, game, want_field=True):
        """b2t_get_game: (GameRaw, field u8[2,H,W] or None) — the game in the reference's own terms"""
        raw = _lib.GameRaw()
        field = np.empty((2, self.H, self.W), np.uint8) if want_field else None
        self._ck(self.lib.b2t_get_game(self._h, int(game), C.byref(raw), _hp(field, C.c_uint8)))
        return raw, field

    def set_game(self, game, raw, field=None):
        field = None if field is None else np.ascontiguousarray(field, np.uint8)
        assert field is None or field.shape == (2, self.H, self.W)
        self._ck(self.lib.b2t_set_game(self._h, int(game), C.byref(raw), _hp(field, C.c_uint8)))

    def copy_game(self, dst, src_engine, src):
        """PythonHandle.copy()/set() of one game slot (device-side, possibly from another handle of the same field size)"""
        self._ck(self.lib.b2t_copy_game(self._h, int(dst), src_engine._h, int(src), None))

    def poke_dead(self, game, player, dead):
        self._ck(self.lib.b2t_poke_dead(self._h, int(game), int(player), int(bool(dead)), None))

    def clear_field(self, game, player):
        self._ck(self.lib.b2t_clear_field(self._h, int(game), int(player), None))

    def stats(self):
        out = np.zeros(8, np.uint64)
        self._ck(self.lib.b2t_stats(self._h, _hp(out, C.c_uint64), None))
        return dict(zip(("env_steps", "episodes", "wins_p0", "wins_p1", "both_dead", "lines", "garbage_queue_overflow",
                         "capacity_overflow"), (int(x) for x in out)))

    def debug_streams(self, seed16, n=64):
        p = np.zeros(n, np.uint8); hl = np.zeros(n, np.uint8)
        self._ck(self.lib.b2t_debug_streams(self._h, int(seed16) & 0xFFFF, n, _hp(p, C.c_uint8), _hp(hl, C.c_uint8)))
        return p, hl

    # ------------------------------------------------------------------ device-tensor face (asynchronous, torch's current stream)
    @staticmethod
    def _dp(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset_dev(self, seeds, idx=None):
        n = self.n if idx is None else idx.numel()
        self._ck(self.lib.b2t_reset(self._h, self._dp(idx), n, self._dp(seeds), self._stream()))

    def _obs(self, field=None, vec=None, piece_idx=None, scalars=None, perspective=False, me=None):
        o = _lib.Obs()
        o.d_field, o.d_vec, o.d_piece_idx, o.d_scalars = (self._dp(field), self._dp(vec), self._dp(piece_idx), self._dp(scalars))
        o.perspective, o.d_me = int(bool(perspective)), self._dp(me)
        o.dtype = self._dtype_of(field, vec)
        return o

    @staticmethod
    def _dtype_of(field, vec):
        """b2t_dtype from the torch dtype of the field / vec tensors (they must agree)"""
        import torch
        if field is not None and field.dtype == torch.int32:     # B2T_BITS: field [rows, 2, SW] int32 bit words, vec stays uint8
            if vec is not None and vec.dtype != torch.uint8:
                raise ValueError("with a bit-packed field (int32) vec must be uint8")
            return _lib.BITS
        codes = {torch.uint8: _lib.U8, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}
        kinds = {codes[t.dtype] for t in (field, vec) if t is not None}
        if len(kinds) > 1:
            raise ValueError("field and vec must have the same dtype")
        return kinds.pop() if kinds else _lib.U8

    def select_dev(self, scores, offsets, placement, capacity=None):
        """argmax of per-placement scores within every board of the last expand_dev -> placement indices for step_dev.  capacity = the
        capacity given to that expand_dev (default: len(scores)); rows it dropped are never chosen."""
        cap = scores.numel() if capacity is None else int(capacity)
        self._ck(self.lib.b2t_select(self._h, self._dp(scores), self._dp(offsets), placement.numel(), cap, self._dp(placement), self._stream()))

    def step_dev(self, player, keys=None, key_len=None, idx=None, placement=None, finalize=True, time_ms=0, reset_seeds=None,
                 reward=None, done=None, lines=None, field=None, vec=None, piece_idx=None, scalars=None, perspective=False,
                 packed_keys=False, auto_reset=False):
        a = _lib.StepArgs()
        a.d_game_idx = self._dp(idx)
        a.n = self.n if idx is None else idx.numel()
        a.d_player, a.d_keys, a.d_key_len = self._dp(player), self._dp(keys), self._dp(key_len)
        a.lmax = (keys.shape[1] * (2 if packed_keys else 1)) if keys is not None else 0
        a.key_format, a.auto_reset = int(bool(packed_keys)), int(bool(auto_reset))
        a.d_placement = self._dp(placement)
        a.finalize, a.time_ms = int(bool(finalize)), int(time_ms)
        a.d_reset_seeds = self._dp(reset_seeds)
        a.d_reward, a.d_done, a.d_lines = self._dp(reward), self._dp(done), self._dp(lines)
        a.obs = self._obs(field, vec, piece_idx, scalars, perspective)
        self._ck(self.lib.b2t_step(self._h, C.byref(a), self._stream()))

    def observe_dev(self, idx=None, field=None, vec=None, piece_idx=None, scalars=None, me=None):
        n = self.n if idx is None else idx.numel()
        o = self._obs(field, vec, piece_idx, scalars, me is not None, me)
        self._ck(self.lib.b2t_observe(self._h, self._dp(idx), n, C.byref(o), self._stream()))

    def enumerate_dev(self, player, K, idx=None, keys=None, key_len=None):
        n = self.n if idx is None else idx.numel()
        kmax = keys.shape[1] if keys is not None else (key_len.shape[1] if key_len is not None else 0)
        lmax = keys.shape[2] if keys is not None else 1
        self._ck(self.lib.b2t_enumerate(self._h, self._dp(idx), n, self._dp(player), self._dp(K), self._dp(keys), self._dp(key_len),
                                        kmax, lmax, self._stream()))

    def expand_dev(self, player, K, offsets, capacity, idx=None, finalize=True, time_ms=0, board=None, keys=None, key_len=None,
                   reward=None, done=None, lines=None, field=None, vec=None, piece_idx=None, perspective=False):
        a = _lib.ExpandArgs()
        a.d_game_idx = self._dp(idx)
        a.n = self.n if idx is None else idx.numel()
        a.d_player = self._dp(player)
        a.finalize, a.time_ms, a.capacity = int(bool(finalize)), int(time_ms), int(capacity)
        a.d_K, a.d_offsets, a.d_board = self._dp(K), self._dp(offsets), self._dp(board)
        a.d_keys, a.d_key_len = self._dp(keys), self._dp(key_len)
        a.lmax = keys.shape[1] if keys is not None else 0
        a.d_reward, a.d_done, a.d_lines = self._dp(reward), self._dp(done), self._dp(lines)
        a.obs = self._obs(field, vec, piece_idx, None, perspective)
        self._ck(self.lib.b2t_expand(self._h, C.byref(a), self._stream()))
