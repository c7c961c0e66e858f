"""The per-env compatibility facade (drl-tetris_b200/b200tetris/tetris_env.py: a module shaped like the reference's pybind11
`tetris_env`, one game slot of the batched engine per PythonHandle).

  * CPU (`-m "not gpu"`): the facade's Python over tests/facade_double.py (the engine's own game core compiled for the CPU by
    tests/hostsim — test infrastructure) is driven call by call against the oracle, and, where /root/reference is present, the
    reference's UNMODIFIED front-end (environment/tetris_environment.py + tetris_environment_vector.py + data_types) is run once
    on the compiled reference backend and once on the facade and every value it produces is compared.
  * GPU (`-m gpu`): the same call-by-call cases on the real CUDA engine through the C ABI.
"""
import os
import pickle
import sys

import numpy as np
import pytest

import harness
import pyoracle
import refdrive
import refmod


class FacadeVector(refdrive.RefVector):
    """n PythonHandle objects of a tetris_env-shaped module behind the vector interface of the harness."""

    def __init__(self, n, H, W, pieces, module):
        self.clock = [0]
        module.set_time_source(lambda: self.clock[0])
        super().__init__(n, H, W, pieces, backend=(module, lambda t: self.clock.__setitem__(0, int(t))))

    def hidden(self, idx=None):
        ids = list(self._idx(idx))
        out = np.zeros((len(ids), 2, 16), np.int32)
        for j, i in enumerate(ids):
            raw = self.h[i]._raw()[0]
            for p in range(2):
                r = raw.player[p]
                n = r.garbage_n
                out[j, p] = [r.time_ms, r.drop_delay, r.drop_delay_time, r.increase_drop_delay_time, r.lockdown_time, r.lockdown, n,
                             sum(r.garbage_count[k] for k in range(n)) & 0xFFFF, r.min_remaining, r.incoming_lines * 1000, 0,
                             r.garbage_count[0] if n else 0, r.garbage_delay[0] if n else 0, r.combo_start, r.combo_time_total, r.line_count]
        return out


CASES = [
    # kind, H, W, pieces, games, steps, solo_period
    ("random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 6, 120, 0),
    ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 260, 0),
    ("junk", 22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 120, 7),
    ("sventon", 22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 80, 0),
    ("greedy", 16, 8, (0, 6), 4, 160, 0),
    ("random", 40, 20, (0, 1, 2, 3, 4, 5, 6), 2, 60, 0),
]


def run_case(module, kind, H, W, pieces, n, steps, solo, tiles=False):
    kmax, lmax = (256, 128) if W > 10 else (96, 64)
    lead = pyoracle.OracleVector(n, H, W, pieces)
    fac = FacadeVector(n, H, W, pieces, module)
    hid = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15]   # all of tor_hidden but the lifetime line total (not kept by the engine)

    def on_step(t, player, keys, klen, rb, fb, enum, other):
        a, b = lead.hidden()[..., hid], other.hidden()[..., hid]
        if t == 0:
            a[..., -2:] = 0; b[..., -2:] = 0
        harness.assert_same("hidden", t, a, b)
        f, g = lead.raw()[:2]
        for i in range(n):
            for p in range(2):
                assert np.array_equal(g[i, p], np.array(other.h[i].states[p].piece)), ("piece grid", t, i, p)
                if tiles:      # raw cell values 1..8 like the reference's states[p].field
                    assert np.array_equal(f[i, p], np.array(other.h[i].states[p].field)), ("raw field", t, i, p)
    return harness.rollout_pair(lead, fac, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, solo_period=solo), on_step=on_step)


def semantics(module):
    """copy / set / write-through views / lazy make_action / pickling / slot recycling of the facade against the oracle."""
    clock = [1234]
    module.set_time_source(lambda: clock[0])
    module.set_pieces([0, 1, 2, 3, 4, 5, 6])
    o = pyoracle.OracleVector(3)                       # game 0 = the handle, 1 = its copy, 2 = scratch
    o.reset([1234, 1234, 1234])
    h = module.PythonHandle(2, [22, 10])
    assert h.last_winner == -1

    def same(handle, game):
        f, g, s = o.raw([game])
        for p in range(2):
            st = handle.states[p]
            assert np.array_equal(np.array(st.field), (f[0, p] > 0).astype(np.uint8))
            assert np.array_equal(np.array(st.piece), g[0, p])
            got = [st.x[0], st.y[0], st.inc_lines[0], st.combo_time[0], st.combo_count[0], st.nextpiece[0], st.reward[0], st.dead[0]]
            assert [int(v) for v in got] == [int(v) for v in s[0, p]], (got, s[0, p])
        assert st.x.dtype == np.int8 and st.combo_time.dtype == np.uint16 and st.field.dtype == np.uint8 and st.field.shape == (22, 10)

    def play(handle, game, player, keys, finalize=True):
        a = [[0], [0]]
        a[player] = list(keys)
        handle.make_action(a)
        k = np.zeros((1, 64), np.uint8); k[0, :len(keys)] = keys
        r = o.step([player], k, [len(keys)], [game], finalize=finalize)
        if finalize:
            assert bool(handle.finish_action(400)) == bool(r[1][0])

    for t, keys in enumerate([[7], [1, 1, 7], [8, 3, 3, 3, 7], [2, 7], [10, 4, 7], [9, 7]]):
        play(h, 0, t & 1, keys)
    same(h, 0)
    # make_action without finish_action is visible to a read (lazy keys are flushed), and finishing later is the same game
    play(h, 0, 0, [1, 6], finalize=False)
    same(h, 0)
    k = np.zeros((1, 64), np.uint8)
    o.step([0], k, [1], [0])
    assert not h.finish_action(400)
    same(h, 0)
    # both players act in one make_action (the front-end never does, the binding allows it): applied in player order
    h.make_action([[3, 7], [1, 7]])
    k[0, :2] = [3, 7]; o.step([0], k, [2], [0], finalize=False)
    k[0, :2] = [1, 7]; o.step([1], k, [2], [0])
    h.finish_action(400)
    same(h, 0)
    # copy(): an independent game; set(): back again
    c = h.copy()
    o.copy_game_from(1, o, 0)
    play(c, 1, 0, [4, 7]); play(c, 1, 1, [7])
    same(c, 1); same(h, 0)
    h.set(c)
    o.copy_game_from(0, o, 1)
    same(h, 0)
    # write-through views: state.lock() (data_types/state.py:9-17) and the solo-eval field wipe (scripts/eval.py:156)
    for s in h.states:
        s.dead[0] = 1
    o.poke_dead(0, 0, 1); o.poke_dead(0, 1, 1)
    play(h, 0, 0, [7])
    same(h, 0)
    h.states[0].dead[0] = 0; h.states[1].dead[0] = 0
    o.poke_dead(0, 0, 0); o.poke_dead(0, 1, 0)
    c.states[1].field[:, :] = 0
    for y in range(22):
        for x in range(10):
            o.poke_cell(1, 1, y, x, 0)
    c.states[0].field[21, 3] = 1
    o.poke_cell(1, 0, 21, 3, 1)
    same(c, 1)
    play(c, 1, 1, [2, 7]); play(c, 1, 0, [8, 7])
    same(c, 1)
    # get_actions / masks
    for handle, game in ((h, 0), (c, 1)):
        for p in (0, 1):
            handle.get_actions(p)
            K, keys, klen, _, _ = o.enumerate([p], [game], kmax=96, lmax=64)
            want = [[int(v) for v in keys[0, j, :klen[0, j]]] for j in range(K[0])]
            assert handle.masks[p].action == want
    # pickling (drl_tetris/runner.py:69-88): the game continues identically, RNG position included
    d = pickle.loads(pickle.dumps(c))
    for t in range(6):
        play(d, 1, t & 1, [3] * t + [7])
    same(d, 1)
    # reset(): last_winner bookkeeping + a new round from the clock
    h.states[1].dead[0] = 1
    o.poke_dead(0, 1, 1)
    clock[0] = 77
    h.reset()
    o.reset([77], [0])
    same(h, 0)
    assert h.last_winner == 0
    # a recycled slot starts from fresh players
    slot = d._slot
    del d
    import gc
    gc.collect()     # a handle and its State objects form a cycle
    e = module.PythonHandle(2, [22, 10])
    assert e._slot == slot and e.last_winner == -1
    o.copy_game_from(2, pyoracle.OracleVector(1), 0)
    o.reset([77], [2])
    same(e, 2)


# ---------------------------------------------------------------------------------------------------- CPU
@pytest.fixture
def cpu_facade():
    import facade_double
    mod = facade_double.install()
    yield mod
    facade_double.uninstall()


@pytest.mark.parametrize("kind,H,W,pieces,n,steps,solo", CASES)
def test_facade_matches_oracle_cpu(cpu_facade, kind, H, W, pieces, n, steps, solo):
    cpu_facade.set_pieces((list(pieces) * 7)[:7])
    st = run_case(cpu_facade, kind, H, W, pieces, n, steps, solo)
    assert st["steps"] == n * steps


def test_facade_semantics_cpu(cpu_facade):
    semantics(cpu_facade)


@pytest.mark.parametrize("H,W,pieces,n,steps", [(22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 300), (40, 20, (0, 1, 2, 3, 4, 5, 6), 2, 120), (16, 8, (0, 6), 3, 200)])
def test_facade_tile_colours_cpu(cpu_facade, H, W, pieces, n, steps):
    """set_tile_colours(True): states[p].field carries the reference's raw values (1..7 piece.tile, 8 garbage), incl. the
    columns-0..9 garbage quirk at W != 10 (gamePlay.cpp:202) and the overwrite of a piece that dies on spawn."""
    cpu_facade.set_tile_colours(True)
    cpu_facade.set_pieces((list(pieces) * 7)[:7])
    st = run_case(cpu_facade, "greedy", H, W, pieces, n, steps, 0, tiles=True)
    assert st["garbage"] > 0 and st["lines"] > 0


# ---- the reference's own front-end, unmodified, on both backends
REF = refdrive.REFERENCE_TREE


def _reference_front_end():
    """import environment.tetris_environment(_vector) from /root/reference with its heavy, absent dependencies stubbed
    (SURVEY.md §8c); the backend module is injected by the caller."""
    from unittest.mock import MagicMock
    import collections
    import collections.abc
    for name in ("tensorflow", "tensorflow.compat", "tensorflow.compat.v1", "tensorflow.python", "tensorflow.python.client", "pygame",
                 "redis", "dill", "click", "docopt"):
        sys.modules.setdefault(name, MagicMock())
    for n in ("Collection", "Mapping"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    if REF in sys.path:
        sys.path.remove(REF)
    sys.path.insert(0, REF)              # ahead of site-packages: an unrelated `agents` distribution shadows the reference's
    for name in [m for m in sys.modules if m == "agents" or m.startswith("agents.")]:
        if not (getattr(sys.modules[name], "__file__", None) or REF).startswith(REF):
            del sys.modules[name]
    import types
    pkg = "environment.game_backend.modules"
    if pkg + ".tetris_env" not in sys.modules:
        sys.modules[pkg + ".tetris_env"] = types.ModuleType("tetris_env")   # placeholder, replaced per run
    import experiments.presets  # noqa: F401  (breaks the circular import tools.utils <-> experiments.presets)
    import environment.tetris_environment as te
    import environment.tetris_environment_vector as tev
    return te, tev


def _front_end_rollout(te, tev, module, set_time, n_envs, steps, settings):
    """worker-loop shaped rollout (drl_tetris/worker.py:91-118,157-160) recording everything the front-end returns"""
    te.tetris_env = module
    set_time(500)
    vec = tev.tetris_environment_vector(n_envs, te.tetris_environment, settings=settings)
    rec = []
    cur = np.arange(n_envs) & 1

    def dicts(state):
        out = []
        for p in (0, 1):
            d = state[p]
            out.append({k: np.array(v).tolist() for k, v in d.items()})
        return out

    for t in range(steps):
        cur = 1 - cur
        states = vec.get_state()
        actions = []
        for i, env in enumerate(vec.envs):
            p = int(cur[i])
            acts = env.get_actions(states[i], player=p)
            pick = int(harness.splitmix64(3, i, t) % np.uint64(len(acts)))
            row = {"t": t, "env": i, "state": dicts(states[i]), "actions": [list(a) for a in acts], "is_dead": [int(x) for x in states[i].is_dead]}
            if (t + i) % 7 == 0:     # the world model, both finalize modes (sherlock_utils.py:13-20 / sventon)
                fin = bool((t // 7) & 1)
                sims = env.simulate_actions(acts, player=p, finalize=fin)
                row["sims"] = [dicts(s) for s in sims]
            if (t + i) % 11 == 3:    # env.copy() / set(): play a move on a copy, the original is untouched
                twin = env.copy()
                twin.perform_action(acts[pick], player=p, simulate=True)   # (a copied env has no reward book-keeping, tetris_environment.py:42-48)
                row["twin"] = dicts(twin.get_state())
            actions.append(acts[pick])
            rec.append(row)
        set_time(1000 + 17 * t)
        rewards, dones = vec.perform_action(actions, player=[int(x) for x in cur])
        rec.append({"t": t, "rewards": [float(r()) for r in rewards], "dones": [bool(d) for d in dones]})
        dead = [i for i, d in enumerate(dones) if d]
        if dead:
            vec.reset(env=dead)
    return rec


@pytest.mark.skipif(not (os.path.isdir(os.path.join(REF, "environment")) and refmod.available("B")),
                    reason="needs /root/reference and oracle/_ref/B (build container only)")
@pytest.mark.parametrize("extra", [{}, {"game_size": [16, 8], "pieces": [0, 6], "bar_null_moves": False}])
def test_unmodified_reference_front_end_runs_on_the_facade(cpu_facade, extra):
    te, tev = _reference_front_end()
    settings = {"presets": ["default"], "render": False}
    settings.update(extra)
    ref_mod, ref_time = refmod.load("B")
    a = _front_end_rollout(te, tev, ref_mod, ref_time, 6, 90, settings)
    clock = [0]
    cpu_facade.set_time_source(lambda: clock[0])
    b = _front_end_rollout(te, tev, cpu_facade, lambda t: clock.__setitem__(0, int(t)), 6, 90, settings)
    assert len(a) == len(b)
    n_sims = n_done = 0
    for x, y in zip(a, b):
        if "state" in x and x["t"] == 0:        # combo.remaining is uninitialised in the reference until the first finish_action
            for q in (x, y):
                for d in q["state"] + [e for s in q.get("sims", []) for e in s] + q.get("twin", []):
                    d["combo_time"] = 0
        assert x == y, (x.get("t"), x.get("env"))
        n_sims += len(x.get("sims", []))
        n_done += sum(x.get("dones", []))
    assert n_sims > 500


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.fixture
def gpu_facade():
    from b200tetris import tetris_env
    tetris_env._arenas.clear()
    tetris_env.set_arena_slots(64)
    yield tetris_env
    tetris_env._arenas.clear()
    tetris_env.set_arena_slots(4096)
    tetris_env.set_tile_colours(False)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,H,W,pieces,n,steps,solo", CASES)
def test_facade_matches_oracle_gpu(gpu_facade, kind, H, W, pieces, n, steps, solo):
    gpu_facade.set_pieces((list(pieces) * 7)[:7])
    st = run_case(gpu_facade, kind, H, W, pieces, n, steps, solo)
    assert st["steps"] == n * steps


@pytest.mark.gpu
def test_facade_semantics_gpu(gpu_facade):
    semantics(gpu_facade)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,pieces,n,steps", [(22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 300), (40, 20, (0, 1, 2, 3, 4, 5, 6), 2, 120), (16, 8, (0, 6), 3, 200)])
def test_facade_tile_colours_gpu(gpu_facade, H, W, pieces, n, steps):
    gpu_facade.set_tile_colours(True)
    gpu_facade.set_pieces((list(pieces) * 7)[:7])
    st = run_case(gpu_facade, "greedy", H, W, pieces, n, steps, 0, tiles=True)
    assert st["garbage"] > 0 and st["lines"] > 0
