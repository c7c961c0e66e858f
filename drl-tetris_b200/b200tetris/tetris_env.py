"""tetris_env — per-env compatibility facade: the surface of the reference's pybind11 module of the same name
(environment/game_backend/source/PythonHandle.h:113-340) on top of the batched CUDA engine, so that the UNMODIFIED
environment/tetris_environment.py (and everything above it: tetris_environment_vector.py, scripts/eval.py, the agents'
sandboxes, the runner's pickling) runs on b200tetris.  Integration = one line, see INTEGRATION.md:

    # environment/game_backend/modules/tetris_env.py
    from b200tetris.tetris_env import *          # noqa

Every `PythonHandle` owns one game SLOT of a shared device-resident engine (one arena per field size / piece set / device),
so `copy()` / `set()` are device-side slot copies instead of the reference's 31 KB deep copies, and nothing but the few bytes
a caller actually looks at crosses PCIe.  This is the compatibility path (one small launch per call); throughput lives in
the vector API (b200tetris.Engine / tetris_environment_vector), which steps every game in one launch.

What is reproduced (reference file:line):
  set_pieces                          PythonHandle.h:116-121   process-global piece map ("last env created imposes its settings")
  PythonHandle(n_players, [H, W])     PythonHandle.cpp:5-25    construction seeds the game from the clock like reset()
  .reset()                            PythonHandle.cpp:49-71   seed = time(NULL)  (set_time_source() to control it)
  .make_action / .finish_action       PythonHandle.cpp:138-188 keys are applied lazily: make_action followed by finish_action is
                                                               ONE fused step launch; any read in between flushes it first
  .get_actions(p) / .masks[p].action  PythonHandle.cpp:190     list of key lists (already de-duplicated; action_list() is idempotent)
  .copy() / .set(other)               PythonHandle.cpp:36-42
  .states[p].field/piece/x/y/inc_lines/combo_time/combo_count/nextpiece/reward/dead   PythonHandle.h:54-82, same dtypes and shapes;
                                      writes through the views (dead[0] = 1 of data_types/state.py:9-17, field[:, :] = 0 of
                                      scripts/eval.py:156) reach the device.  Views are snapshots: re-read after stepping.
  .last_winner / .recreate_state()    PythonHandle.h:315,319
  pickling                            PythonHandle.h:283-308   (unlike the reference the RNG position survives a pickle)
Tile colours in `field` (1..7 = piece.tile, 8 = garbage) are kept after set_tile_colours(True) (B2T_FLAG_TILES); by default cells
read 1 where the reference has 1..8 — only `field > 0` is observable through state_dict.  Not offered by this facade: n_players != 2 (the engine and the vector env take 1..4).
"""
import gc
import time as _time_mod

import numpy as np

from . import _lib
from .engine import Engine

__all__ = ["set_pieces", "PythonHandle", "State", "Mask", "set_time_source", "set_device", "set_arena_slots", "set_tile_colours"]

_pieces = [0, 1, 2, 3, 4, 5, 6]   # GamePlay::randomizer_map (gamePlay.cpp:9)
_time = _time_mod.time            # what time(NULL) returns
_device = 0
_slots = 4096                     # game slots per arena chunk
_tiles = False                    # keep raw cell values 1..8 (B2T_FLAG_TILES) for handles created afterwards
_arenas = {}
_engine_factory = Engine          # tests substitute a CPU double with the same small interface; the product has no CPU path


def set_pieces(a):
    """tetris_env.set_pieces (PythonHandle.h:116-121): the 7-entry piece map of every handle created afterwards."""
    global _pieces
    a = [int(x) for x in a]
    assert len(a) == 7 and all(0 <= x <= 6 for x in a), "set_pieces takes 7 piece ids in 0..6"
    _pieces = a


def set_time_source(fn):
    """fn() -> what time(NULL) would return; the low 16 bits seed a round (PythonHandle.cpp:68-71, randomizer.cpp:34-36)."""
    global _time
    _time = fn


def set_device(device):
    global _device
    _device = int(device)


def set_tile_colours(on):
    """states[p].field then reads 1..7 (piece.tile) / 8 (garbage) like the reference's (renderer, raw-field parity) instead of 1."""
    global _tiles
    _tiles = bool(on)


def set_arena_slots(n):
    global _slots
    _slots = int(n)


class _Arena:
    """Game slots of one (field size, piece map, device): a list of engine chunks and a free list of (chunk, slot)."""

    def __init__(self, H, W, pieces, device, tiles=False):
        self.key = (H, W, tuple(pieces), device, tiles)
        self.chunks, self.free = [], []

    def alloc(self):
        if not self.free and self.chunks:
            gc.collect()     # handles and their State objects reference each other: dropped copies free their slots here
        if not self.free:
            H, W, pieces, device, tiles = self.key
            n = _slots * (1 << min(len(self.chunks), 4))
            eng = _engine_factory(n, H, W, pieces, 2, True, 400, device=device, tiles=tiles)
            self.chunks.append(eng)
            self.free.extend((eng, s) for s in range(n - 1, 0, -1))   # slot 0 stays a never-reset game: the template of new handles
        return self.free.pop()

    def release(self, eng, slot):
        self.free.append((eng, slot))


def _arena(H, W):
    key = (int(H), int(W), tuple(_pieces), _device, _tiles)
    a = _arenas.get(key)
    if a is None:
        a = _arenas[key] = _Arena(*key)
    return a


class _View(np.ndarray):
    """A State member: a numpy array of the reference's dtype/shape whose element writes go through to the device."""

    def __new__(cls, data, dtype, owner, player, name):
        obj = np.array(data, dtype=dtype).view(cls)
        obj._owner, obj._player, obj._name, obj._root = owner, player, name, None
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)
        self._player = getattr(obj, "_player", None)
        self._name = getattr(obj, "_name", None)
        root = getattr(obj, "_root", None)
        self._root = obj if (root is None and isinstance(obj, _View)) else root   # slices remember the whole member

    def __array_wrap__(self, arr, context=None, return_scalar=False):   # results of ufuncs / reductions are plain numpy
        arr = np.asarray(arr)
        return arr[()] if arr.ndim == 0 else arr

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        whole = self if self._root is None else self._root
        if self._owner is not None and np.shares_memory(self, whole):     # copies (f + 1, f.copy()) are plain arrays
            self._owner._write(self._player, self._name, np.asarray(whole))


_FIELDS = {  # State member -> (b2t_player_raw field, dtype)   PythonHandle.h:70-81
    "x": ("x", np.int8), "y": ("y", np.int8), "inc_lines": ("inc_lines", np.uint8), "combo_time": ("combo_time", np.uint16),
    "combo_count": ("combo_count", np.uint8), "nextpiece": ("nextpiece", np.uint8), "reward": ("reward", np.uint8),
    "dead": ("dead", np.uint8),
}


class State:
    """PythonHandle.states[p] (PythonHandle.h:54-82)."""

    def __init__(self, owner, player):
        self._owner, self._player = owner, player

    def __getattr__(self, name):
        if name in _FIELDS:
            raw, _ = self._owner._raw()
            key, dt = _FIELDS[name]
            return _View([getattr(raw.player[self._player], key)], dt, self._owner, self._player, name)
        raise AttributeError(name)

    @property
    def field(self):
        _, field = self._owner._raw()
        return _View(field[self._player], np.uint8, self._owner, self._player, "field")

    @property
    def piece(self):
        raw, _ = self._owner._raw()
        return _View(np.array(raw.player[self._player].piece, np.uint8).reshape(4, 4), np.uint8, None, self._player, "piece")

    def __getstate__(self):   # "pickles as (0)", PythonHandle.h:339
        return 0

    def __setstate__(self, _):
        self._owner, self._player = None, 0


class Mask:
    """PythonHandle.masks[p] (TestField.h:9-32): only `.action` is read by Python."""

    def __init__(self):
        self.action, self.mask = [], []

    def clear(self):
        self.action, self.mask = [], []


class PythonHandle:
    def __init__(self, n_players, field_size):
        assert int(n_players) == 2, ("the per-env facade exchanges games as b2t_game_raw, which describes two players (n_players=%r); other "
                                     "player counts run through b200tetris.Engine / tetris_environment_vector" % (n_players,))
        H, W = (int(v) for v in field_size)
        self._bind(_arena(H, W))
        self.masks = [Mask(), Mask()]
        self._eng.copy_game(self._slot, self._eng, 0)   # fresh players (a recycled slot still holds its last game)
        self._seed_round()               # PythonHandle::init -> restartRound + seed (PythonHandle.cpp:15-19)

    # ------------------------------------------------------------------ plumbing
    def _bind(self, arena):
        self._arena = arena
        self._eng, self._slot = arena.alloc()
        self._pending = None
        self._cache = None
        self.states = [State(self, 0), State(self, 1)]
        self.H, self.W = arena.key[0], arena.key[1]

    def __del__(self):
        try:
            if getattr(self, "_arena", None) is not None:
                self._arena.release(self._eng, self._slot)
                self._arena = None
        except Exception:
            pass

    def _seed_round(self):
        self._eng.reset(np.array([int(_time()) & 0xFFFFFFFF], np.uint32), [self._slot])
        self._cache = None

    @staticmethod
    def _acts(keys):
        return any(1 <= int(k) <= 10 for k in keys)      # every other code is a no-op (PythonHandle.cpp:109-110)

    def _step(self, player, keys, finalize, ms=0):
        """One step launch for this slot; key lists longer than the engine's sequence capacity are applied in pieces."""
        keys = [int(k) if 0 <= int(k) <= 255 else 0 for k in keys] or [0]
        cap = self._eng.lmax
        done = False
        while True:
            head, keys = keys[:cap], keys[cap:]
            last = not keys
            width = (len(head) + 15) // 16 * 16
            arr = np.zeros((1, width), np.uint8)
            arr[0, :len(head)] = head
            out = self._eng.step([player], arr, [len(head)], [self._slot], time_ms=ms, finalize=bool(finalize and last))
            done = bool(out[1][0])
            if last:
                break
        self._cache = None
        return done

    def _flush(self):
        """Apply a make_action that was not followed by finish_action yet (PythonHandle.cpp:138-147, players in index order)."""
        acts, self._pending = self._pending, None
        if acts is not None:
            for p in (0, 1):
                if self._acts(acts[p]):
                    self._step(p, acts[p], False)

    def _raw(self):
        self._flush()
        if self._cache is None:
            self._cache = self._eng.get_game(self._slot)
        return self._cache

    def _write(self, player, name, value):
        raw, field = self._raw()
        if name == "field":
            field[player] = np.asarray(value) if self._arena.key[4] else (np.asarray(value) != 0)
            self._eng.set_game(self._slot, raw, field)
        elif name in _FIELDS:
            setattr(raw.player[player], _FIELDS[name][0], int(np.asarray(value).ravel()[0]))
            self._eng.set_game(self._slot, raw, None)
        else:
            raise ValueError("states[p].%s is read-only in b200tetris" % name)
        self._cache = None

    # ------------------------------------------------------------------ the bound surface (PythonHandle.h:309-321)
    def reset(self):
        self._flush()
        self._seed_round()

    def make_action(self, actions):
        assert len(actions) == 2
        self._flush()
        self._pending = [list(actions[0]), list(actions[1])]

    def finish_action(self, time_elapsed):
        acts, self._pending = self._pending, None
        acts = acts or [[0], [0]]
        movers = [p for p in (0, 1) if self._acts(acts[p])]
        if len(movers) == 2:
            self._step(0, acts[0], False)
        player = movers[-1] if movers else 0
        return self._step(player, acts[player], True, int(time_elapsed))

    def get_actions(self, player):
        self._flush()
        K, keys, klen = self._eng.enumerate([int(player)], [self._slot])
        lst = [[int(k) for k in keys[0, j, :klen[0, j]]] for j in range(int(K[0]))]
        if lst == [[0]]:
            lst = []                       # the raw mask list holds no null action (action_list adds it)
        self.masks[int(player)].action = lst

    def copy(self):
        self._flush()
        other = PythonHandle.__new__(PythonHandle)
        other._bind(self._arena)
        other._eng.copy_game(other._slot, self._eng, self._slot)
        other.masks = [Mask(), Mask()]
        for a, b in zip(other.masks, self.masks):
            a.action = [list(x) for x in b.action]
        return other

    def set(self, other):
        assert isinstance(other, PythonHandle)
        assert (other.H, other.W) == (self.H, self.W), "set() between different field sizes"
        self._flush()
        other._flush()
        if other._arena is self._arena:
            self._eng.copy_game(self._slot, other._eng, other._slot)
        else:                                # another piece map / device: through the host codec
            raw, field = other._eng.get_game(other._slot)
            self._eng.set_game(self._slot, raw, field)
        for a, b in zip(self.masks, other.masks):
            a.action = [list(x) for x in b.action]
        self._cache = None

    def recreate_state(self):
        return None

    @property
    def last_winner(self):
        return int(self._raw()[0].last_winner)

    # ------------------------------------------------------------------ pickling (PythonHandle.h:283-308; drl_tetris/runner.py:69-88)
    def __getstate__(self):
        raw, field = self._raw()
        return {"H": self.H, "W": self.W, "pieces": list(self._arena.key[2]), "raw": bytes(raw), "field": field.tobytes(),
                "masks": [[list(x) for x in m.action] for m in self.masks]}

    def __setstate__(self, d):
        set_pieces(d["pieces"])              # the reference's unpickle restores its process globals the same way (:300-304)
        self._bind(_arena(d["H"], d["W"]))
        raw = _lib.GameRaw.from_buffer_copy(d["raw"])
        self._eng.set_game(self._slot, raw, np.frombuffer(d["field"], np.uint8).reshape(2, d["H"], d["W"]).copy())
        self.masks = [Mask(), Mask()]
        for m, a in zip(self.masks, d["masks"]):
            m.action = [list(x) for x in a]
