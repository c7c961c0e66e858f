"""Reference-shaped front-end: the surface of environment/tetris_environment_vector.py (reference file:line cited per
method) on top of one Engine handle.  Same method names, `env` / `player` broadcasting (tools/utils.py:10-31 parse_arg),
return shapes (Python lists of rewards / dones / state objects) and assert-style argument checks; the per-call Python
`for` loop over one C++ object per game is replaced by one batched C-ABI call.

Deliberate differences, all documented in INTEGRATION.md:
  * seeds: the reference reseeds every game from time(NULL) (PythonHandle.cpp:68-71), so envs reset in the same second
    share one piece stream; here seed = settings["seed"] (default: time) hashed with env id and episode number.
  * the methods the reference ships broken (get_actions / simulate_all_actions without a state argument, get_winner via
    the unbound backend.info) work here on the current states.
  * states returned by simulate_* carry observations only (they cannot be `set()` back).
"""
import time

import numpy as np

from .engine import Engine

_engine_factory = Engine   # tests substitute a CPU double with the same interface; the product has no CPU path

DEFAULTS = {  # experiments/presets.py:123-182 ("default"), the keys that reach the backend
    "game_size": [22, 10], "pieces": [0, 1, 2, 3, 4, 5, 6], "n_players": 2, "bar_null_moves": True,
    "time_elapsed_each_action": 400, "extra_rewards": False, "reward_ammount": (1.0, 0.0), "action_type": "place_block",
    "state_processor": "state_dict", "render": False,
}


class action(list):  # environment/data_types/action.py:2-7
    def __init__(self, container):
        assert type(container) in [list, action], "action created from non-list type object...: x={}, type(x)={}".format(container, type(container))
        super().__init__(container)

    def __str__(self):
        return "action(" + super().__str__() + ")"


null_action = action([0])


class action_list(list):  # environment/data_types/action_list.py:3-37 (the de-duplication itself already happened on the GPU)
    def __init__(self, container=None, remove_null=False):
        self.remove_null = remove_null
        if container is None or len(container) == 0:
            container = [null_action]
        super().__init__(action(list(x)) for x in container)
        self.container = self

    @classmethod
    def from_rows(cls, rows, remove_null):
        """rows: plain lists of ints straight from the engine's arrays (no per-action type checks: a vector env builds
        hundreds of thousands of these per call)"""
        obj = cls.__new__(cls)
        acts = []
        for r in rows or [[0]]:
            a = action.__new__(action)
            list.__init__(a, r)
            acts.append(a)
        list.__init__(obj, acts)
        obj.remove_null, obj.container = remove_null, obj
        return obj


class reward:  # environment/data_types/reward.py:64-71 maingoal_reward, the subset the env itself uses
    def __init__(self, extrinsic):
        self._extrinsic = np.array(extrinsic, dtype=float).ravel()
        self._intrinsic = np.zeros((1,))

    extrinsic = property(lambda self: self._extrinsic)
    intrinsic = property(lambda self: self._intrinsic)

    def __call__(self):
        return self._extrinsic.sum() + self._intrinsic.sum()

    def __add__(self, other):
        tmp = np.zeros_like(other._extrinsic)
        tmp[0] = other._extrinsic[0]
        return reward(self._extrinsic + tmp)

    def __repr__(self):
        return "reward<R=%s>" % self()


class state:
    """environment/data_types/state.py:1-40: `state[player]` -> the state_dict of state_processors.py:23-54."""

    def __init__(self, field, vec, piece_idx, dead, packed=None):
        self._field, self._vec, self._piece_idx = field, vec, piece_idx
        self.is_dead = [int(d) for d in dead]
        self.backend_state = packed   # packed engine state (bytes) or None for simulated states
        self.unlocked = True

    def _dict(self, p):
        v = self._vec[p]
        return {"field": self._field[p], "piece": v[12:19], "x": v[0:1], "y": v[1:2], "incoming_lines": v[2:3],
                "combo_time": v[3:4], "combo_count": v[4:5], "nextpiece": v[5:12],
                "piece_idx": int(self._piece_idx[p])}

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self._dict(p) for p in range(*idx.indices(len(self)))]
        if not hasattr(idx, "__iter__"):
            return self._dict(idx)
        return [self._dict(p) for p in idx]

    def __len__(self):
        return self._field.shape[0]


def parse_arg(entry_idx, data, fill_up=None):  # tools/utils.py:10-31
    """Indices into `data` (the list of envs, or of players) with Python's own indexing rules, which is what the reference gets from
    `self.envs[i]`: a negative index counts from the end, anything outside raises IndexError."""
    n = len(data)
    def norm(i):
        i = int(i)
        if not -n <= i < n:
            raise IndexError("index %d out of range for %d entries" % (i, n))
        return i + n if i < 0 else i
    if entry_idx is None:
        return list(range(n))
    if type(entry_idx) in [list, np.ndarray]:
        return [norm(i) for i in entry_idx]
    return [norm(entry_idx)] if fill_up is None else [norm(entry_idx)] * fill_up


class tetris_environment_vector:
    def __init__(self, n_envs, env_type=None, init_envs=None, settings=None, device=0):  # tetris_environment_vector.py:10-20
        s = dict(DEFAULTS)
        s.update({k: v for k, v in (settings or {}).items() if k != "presets"})
        self.settings = s
        assert s["action_type"] in ["place_block", "press_key"]          # tetris_environment.py:205
        # the reference's own front-end runs with 2 players, and — its reward rule reading backend.states[1 - player] as a Python
        # index — with 3 or 4 (tetris_environment.py:139); with one player that index does not exist
        assert s["n_players"] in (2, 3, 4), "tetris_environment needs an opponent: n_players must be 2, 3 or 4"
        self.n_envs = n_envs
        self.player_idxs = list(range(s["n_players"]))
        self.envs = list(range(n_envs))
        H, W = s["game_size"]
        self.engine = _engine_factory(n_envs, H, W, s["pieces"], s["n_players"], s["bar_null_moves"], s["time_elapsed_each_action"], device=device)
        self._seed0 = int(s.get("seed", time.time()))
        self._episode = np.zeros(n_envs, np.int64)
        self.done = np.zeros(n_envs, bool)
        self.rounds_played = np.zeros(n_envs, np.int64)
        self.round_reward = [[reward([0]) for _ in self.player_idxs] for _ in range(n_envs)]
        self.tot_reward = [[reward([0]) for _ in self.player_idxs] for _ in range(n_envs)]
        self._dead = np.zeros((n_envs, s["n_players"]), np.uint8)
        self.reset()   # "Upon agreement with backend, we always reset once." tetris_environment.py:40-41

    # ------------------------------------------------------------------ env interface
    def _seeds(self, ids):
        ids = np.asarray(ids, np.int64)
        z = (self._seed0 + ids * 0x9E3779B1 + self._episode[ids] * 0x85EBCA77) & 0xFFFFFFFF
        z ^= z >> 15
        return (z & 0xFFFF).astype(np.uint32)

    def reset(self, env=None):  # :55-58 ; tetris_environment.reset :66-75
        ids = parse_arg(env, self.envs)
        if not ids:
            return []
        self.engine.reset(self._seeds(ids), ids)
        for i in ids:
            self._episode[i] += 1
            self.done[i] = False
            self.rounds_played[i] += 1
            self.round_reward[i] = [reward([0]) for _ in self.player_idxs]
        self._dead[ids] = 0
        return [None for _ in ids]

    def _lists(self, K, keys, klen):
        """[n, kmax, lmax] key arrays -> one action_list per env; only the valid tokens are turned into Python objects"""
        n, kmax, lmax = keys.shape
        rn = self.settings["bar_null_moves"]
        valid = np.arange(kmax)[None, :] < K[:, None]
        lens = klen[valid]
        tok = keys[valid]
        flat = tok[np.arange(lmax)[None, :] < lens[:, None]].tolist()
        ends = np.cumsum(lens).tolist()
        out, a, pos = [], 0, 0
        for k in K.tolist():
            rows = []
            for _ in range(k):
                e = ends[a]
                rows.append(flat[pos:e])
                pos, a = e, a + 1
            out.append(action_list.from_rows(rows, rn))
        return out

    def get_actions(self, env=None, player=None):  # :60-65 ; tetris_environment.get_actions :77-85
        ids = parse_arg(env, self.envs)
        pl = parse_arg(player, self.player_idxs, fill_up=len(ids))
        assert len(ids) == len(pl), "You want to specify one player per tetris_environment that the tetris_environment_vector manages"
        if self.settings["action_type"] == "press_key":
            return [action_list([[k] for k in range(11)]) for _ in ids]
        return self._lists(*self.engine.enumerate(pl, ids))

    def get_random_action(self, env=None, player=None):  # :67-71
        return [a[np.random.randint(low=0, high=len(a))] for a in self.get_actions(env, player)]

    def perform_action(self, actions, env=None, player=None):  # :79-86 ; tetris_environment.perform_action :102-116
        ids = parse_arg(env, self.envs)
        pl = parse_arg(player, self.player_idxs, fill_up=len(ids))
        n = len(ids)
        assert len(actions) == n and len(pl) == n
        for a in actions:
            assert type(a) in (action, list), "tetris_environment.perform_action(action a,int p) was called with type(action)={}".format(type(a))
        # key codes outside a byte are no-ops like every other unknown code (PythonHandle.cpp:109-110)
        actions = [[k if 0 <= k <= 255 else 0 for k in map(int, a)] for a in actions]
        # A key list longer than the engine's sequence capacity (press_key-style callers may send anything) is applied in pieces, as
        # the per-env facade does (tetris_env._step): its leading keys go down in make_action-only launches, the tail with the finish.
        cap = self.engine.lmax
        long_ = [j for j, a in enumerate(actions) if len(a) > cap]
        while long_:
            head = np.array([actions[j][:cap] for j in long_], np.uint8)
            self.engine.step([pl[j] for j in long_], head, np.full(len(long_), cap, np.uint8), [ids[j] for j in long_], finalize=False)
            for j in long_:
                actions[j] = actions[j][cap:]
            long_ = [j for j in long_ if len(actions[j]) > cap]
        lmax = max(1, max(len(a) for a in actions))
        lmax = (lmax + 15) // 16 * 16
        klen = np.fromiter(map(len, actions), np.int64, n)
        keys = np.zeros((n, lmax), np.uint8)
        keys[np.arange(lmax)[None, :] < klen[:, None]] = np.fromiter((k for a in actions for k in a), np.uint8, int(klen.sum()))
        klen = klen.astype(np.uint8)
        base, done, _lines, scal = self.engine.step(pl, keys, klen, ids, want_scalars=True)
        rewards = []
        for j, i in enumerate(ids):
            self.done[i] = bool(done[j])
            self._dead[i] = scal[j, :, 7]
            if self.settings["extra_rewards"]:           # tetris_environment.py:146-149
                wb, wc = self.settings["reward_ammount"]
                r = reward([wb * int(base[j]), wc * int(scal[j, pl[j], 4])])
            else:
                r = reward([int(base[j])])
            self.round_reward[i][pl[j]] = self.round_reward[i][pl[j]] + r
            self.tot_reward[i][pl[j]] = self.tot_reward[i][pl[j]] + r
            rewards.append(r)
        return rewards, [bool(d) for d in done]

    def get_winner(self, env=None, player=None):  # :88-91,106-108 ; tetris_environment.get_winner :118-125
        out = []
        for i in parse_arg(env, self.envs):
            if not self.done[i]:
                out.append(None)
            else:
                alive = [p for p in self.player_idxs if not self._dead[i, p]]
                out.append(alive[0] if alive else 666)
        return out

    def get_state(self, env=None):  # :99-101 ; tetris_environment.get_state :131-133
        ids = parse_arg(env, self.envs)
        field, vec, pidx, scal = self.engine.observe(ids, want_scalars=True)
        packed = self.engine.export_states(ids)          # one batched copy: the snapshot data_types.state keeps of its backend
        return [state(field[j], vec[j], pidx[j], scal[j, :, 7], packed[j]) for j in range(len(ids))]

    def simulate_all_actions(self, env=None, player=None, finalize=True):  # :93-97 ; tetris_environment :127-129,87-100
        ids = parse_arg(env, self.envs)
        pl = parse_arg(player, self.player_idxs, fill_up=len(ids))
        o = self.engine.expand(pl, ids, finalize=finalize, want_scalars=True)
        dead = o["scalars"][:, :, 7]      # states[p].dead of every simulated outcome (data_types/state.py:9-17 is_dead)
        out = []
        for j in range(len(ids)):
            a, b = o["offsets"][j], o["offsets"][j + 1]
            out.append([state(o["field"][r], o["vec"][r], o["piece_idx"][r], dead[r]) for r in range(a, b)])
        return out

    def simulate_actions(self, actions, env=None, player=None, finalize=True):  # :73-77 ; tetris_environment.simulate_actions :87-100
        """Only lists returned by get_actions for the current states are supported (as every caller in the reference does)."""
        ids = parse_arg(env, self.envs)
        sims = self.simulate_all_actions(ids, player, finalize)
        cur = self.get_actions(ids, player)
        out = []
        for sim, have, want in zip(sims, cur, actions):
            lut = {tuple(a): s for a, s in zip(have, sim)}
            out.append([lut[tuple(a)] for a in want])
        return out

    def set(self, target, env=None):  # :113-118 ; tetris_environment.set :168-176
        ids = parse_arg(env, self.envs)
        targets = target if type(target) is list else [target] * len(ids)
        for i, t in zip(ids, targets):
            assert isinstance(t, state) and t.backend_state is not None, "set() needs a state returned by get_state()"
            self.engine.import_state(i, t.backend_state)
            self._dead[i] = t.is_dead

    def copy(self):  # :161-167
        other = tetris_environment_vector(self.n_envs, settings=self.settings, device=self.engine.device)
        other.engine.import_states(self.engine.export_states())
        other.done[:] = self.done
        other._dead[:] = self._dead
        # the bookkeeping of the rounds so far travels too: the copy continues the original's seed sequence and reward totals
        other._seed0 = self._seed0
        other._episode[:] = self._episode
        other.rounds_played[:] = self.rounds_played
        other.round_reward = [[reward(r.extrinsic) for r in rr] for rr in self.round_reward]
        other.tot_reward = [[reward(r.extrinsic) for r in rr] for rr in self.tot_reward]
        return other

    def get_info(self, env=None):  # tetris_environment.get_info :151-160
        return [{"is_dead": list(self._dead[i]), "tot_reward": self.tot_reward[i], "round_reward": self.round_reward[i],
                 "rounds_played": int(self.rounds_played[i])} for i in parse_arg(env, self.envs)]

    def render(self, env=None):  # :133-155 — pygame UI is out of scope
        return None

    def __str__(self):
        return "<tetris_vector_env n_envs=%d settings=%s>" % (self.n_envs, {k: self.settings[k] for k in DEFAULTS})
