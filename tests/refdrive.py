"""Test helper: drives the compiled, unmodified reference backend (oracle/_ref, see oracle/refmod.py) through the
same vector-shaped interface as oracle/pyoracle.OracleVector, so the two can be diffed call by call.

Optionally routes get_actions / observations through the reference's own Python (`action_list`, `state_dict`)
imported from /root/reference when that tree is present (never at GPU-test time)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refmod  # noqa: E402

REFERENCE_TREE = "/root/reference"


def reference_python():
    """(action_list class, null_action, state_dict fn) from the reference tree, or None."""
    if not os.path.isdir(os.path.join(REFERENCE_TREE, "environment")):
        return None
    if REFERENCE_TREE not in sys.path:
        sys.path.append(REFERENCE_TREE)
    from environment.data_types.action_list import action_list
    from environment.data_types.action import null_action
    from environment.env_utils.state_processors import state_dict
    return action_list, null_action, state_dict


class RefVector:
    def __init__(self, n_games, height=22, width=10, pieces=(0, 1, 2, 3, 4, 5, 6), n_players=2, variant="B", backend=None):
        """backend: (module, set_time) of any tetris_env-shaped module (the per-env facade of b200tetris); default = the compiled
        reference"""
        self.mod, self.set_time = backend if backend is not None else refmod.load(variant)
        self.n, self.H, self.W, self.P = n_games, height, width, n_players
        self.map = (list(pieces) * 7)[:7]
        self.mod.set_pieces(self.map)
        self.set_time(0)
        self.h = [self.mod.PythonHandle(n_players, [height, width]) for _ in range(n_games)]
        self.env_done = [False] * n_games

    def _idx(self, idx):
        return range(self.n) if idx is None else [int(i) for i in idx]

    def reset(self, seeds, idx=None):
        for s, i in zip(seeds, self._idx(idx)):
            self.set_time(int(s))
            self.h[i].reset()
            self.env_done[i] = False

    def step(self, player, keys, key_len, idx=None, time_ms=400, finalize=True):
        ids = list(self._idx(idx))
        n = len(ids)
        reward = np.zeros(n, np.int8)
        done = np.zeros(n, np.uint8)
        lines = np.zeros((n, self.P), np.uint8)
        for j, i in enumerate(ids):
            h = self.h[i]
            a = [[0] for _ in range(self.P)]
            p = int(player[j])
            a[p] = [int(k) for k in keys[j, : int(key_len[j])]]
            h.make_action(a)
            if finalize:
                self.env_done[i] = bool(h.finish_action(int(time_ms)))
            base = 0
            if self.env_done[i]:
                me = int(h.states[p].dead[0]) != 0
                you = (int(h.states[1 - p].dead[0]) != 0) if self.P > 1 else False
                base = int(you) - int(me)
                if me and you:
                    base = -1
            reward[j] = base
            done[j] = self.env_done[i]
            for q in range(self.P):
                lines[j, q] = h.states[q].reward[0]
        return reward, done, lines

    def raw(self, idx=None):
        ids = list(self._idx(idx))
        n = len(ids)
        field = np.zeros((n, self.P, self.H, self.W), np.uint8)
        grid = np.zeros((n, self.P, 4, 4), np.uint8)
        scal = np.zeros((n, self.P, 8), np.int32)
        for j, i in enumerate(ids):
            for p in range(self.P):
                s = self.h[i].states[p]
                field[j, p] = np.array(s.field)
                grid[j, p] = np.array(s.piece)
                scal[j, p] = [s.x[0], s.y[0], s.inc_lines[0], s.combo_time[0], s.combo_count[0], s.nextpiece[0],
                              s.reward[0], s.dead[0]]
        return field, grid, scal

    def scal(self, idx=None):
        return self.raw(idx)[2]

    def observe(self, idx=None):
        """state_dict (state_processors.py:23-54): through the reference's own Python when its tree is present,
        else restated here from the raw views."""
        if reference_python() is not None:
            return self.observe_reference_python(idx)
        field, grid, scal = self.raw(idx)
        code = np.array([255, 5, 4, 1, 0, 2, 6, 3], np.uint8)
        kind = code[grid.reshape(grid.shape[0], self.P, 16).max(-1)]
        vec = np.zeros(field.shape[:2] + (19,), np.uint8)
        vec[..., 0] = scal[..., 0].astype(np.uint8)
        vec[..., 1] = scal[..., 1].astype(np.uint8)
        vec[..., 2] = scal[..., 2]
        ct = (scal[..., 3] + 50) & 0xFFFF
        vec[..., 3] = np.minimum(ct, 25000) // 100
        vec[..., 4] = scal[..., 4]
        for k in range(7):
            vec[..., 5 + k] = scal[..., 5] == k
            vec[..., 12 + k] = kind == k
        return (field > 0).astype(np.uint8), vec, kind.astype(np.uint8)

    def enumerate(self, player, idx=None, bar_null_moves=True, kmax=64, lmax=32):
        """get_actions (tetris_environment.py:77-85) incl. action_list (through the reference's own class when present)."""
        raw = self.enumerate_raw(player, idx)
        ref = reference_python()
        n = len(raw)
        K = np.zeros(n, np.int32)
        keys = np.zeros((n, kmax, lmax), np.uint8)
        klen = np.zeros((n, kmax), np.uint8)
        for i, acts in enumerate(raw):
            if ref is not None:
                lst = [list(a) for a in ref[0](acts, remove_null=bar_null_moves)]
            else:
                lst = [[0]]
                for a in acts:
                    if a not in lst:
                        lst.append(a)
                if bar_null_moves and len(lst) > 1:
                    lst = lst[1:]
            K[i] = len(lst)
            for k, a in enumerate(lst[:kmax]):
                keys[i, k, : len(a)] = a[:lmax]
                klen[i, k] = min(len(a), lmax)
        return K, keys, klen, np.array([len(a) for a in raw], np.int32), 0

    def hidden(self, idx=None):
        ids = list(self._idx(idx))
        out = np.zeros((len(ids), self.P, 16), np.int32)
        for j, i in enumerate(ids):
            players = self.h[i].__getstate__()[0]
            for p in range(self.P):
                g = players[p].__getstate__()
                garb, mn, _ad = g[5].__getstate__()
                dd = g[7].__getstate__()
                q = [x.__getstate__() for x in garb]
                out[j, p, :13] = [g[11], dd[0], dd[1], dd[3], dd[4], int(dd[5]), len(q), sum(c for c, _ in q) & 0xFFFF, mn,
                                  int(np.float32(g[9]) * np.float32(1000.0)), g[12], q[0][0] if q else 0, q[0][1] if q else 0]
        return out

    def enumerate_raw(self, player, idx=None):
        out = []
        for j, i in enumerate(self._idx(idx)):
            self.h[i].get_actions(int(player[j]))
            out.append([list(a) for a in self.h[i].masks[int(player[j])].action])
        return out

    def observe_reference_python(self, idx=None):
        """state_dict through the reference's own Python (state_processors.py:23-54)."""
        _, _, state_dict = reference_python()
        ids = list(self._idx(idx))
        n = len(ids)
        field = np.zeros((n, self.P, self.H, self.W), np.uint8)
        vec = np.zeros((n, self.P, 19), np.uint8)
        pidx = np.zeros((n, self.P), np.uint8)
        params = [self.map, False, True, False]  # pieces, old_state_dict, separate_piece, augment_data (presets.py:123-182)
        for j, i in enumerate(ids):
            for p in range(self.P):
                d = state_dict(self.h[i], p, params)
                field[j, p] = d["field"]
                vec[j, p] = np.concatenate([d["x"].ravel(), d["y"].ravel(), np.asarray(d["incoming_lines"]).ravel(),
                                            np.asarray(d["combo_time"]).ravel(), np.asarray(d["combo_count"]).ravel(),
                                            d["nextpiece"].ravel(), d["piece"].ravel()]).astype(np.uint8)
                pidx[j, p] = d["piece_idx"]
        return field, vec, pidx
