"""b200tetris.tetris_environment_vector (one batched call per method) against the reference's own
environment/tetris_environment_vector.py + tetris_environment.py running on the compiled reference backend: same settings, same
round seeds, same worker-loop call pattern (drl_tetris/worker.py:91-118,157-160) — every state dict, action list, simulated
placement, reward, done and winner.  Build container only (needs /root/reference and oracle/_ref/B); on the CPU our vector env
sits on tests/facade_double.py (the engine's game core compiled for the CPU — test infrastructure), the `-m gpu` suite drives it
on the CUDA engine against the oracle (tests/test_gpu_parity.py::test_vector_env_front_end)."""
import os

import numpy as np
import pytest

import harness
import refdrive
import refmod
from test_facade import _reference_front_end

REF = refdrive.REFERENCE_TREE
pytestmark = pytest.mark.skipif(not (os.path.isdir(os.path.join(REF, "environment")) and refmod.available("B")),
                                reason="needs /root/reference and oracle/_ref/B (build container only)")


def dicts(state):
    return [{k: np.array(v).tolist() for k, v in state[p].items()} for p in range(len(state))]


@pytest.mark.parametrize("extra", [{}, {"game_size": [16, 8], "pieces": [0, 6], "bar_null_moves": False},
                                   {"extra_rewards": True, "reward_ammount": (1.0, 0.25)},
                                   {"n_players": 3}, {"n_players": 4, "game_size": [16, 8], "pieces": [0, 6]}])
def test_vector_env_matches_the_reference_vector_env(extra):
    import facade_double
    facade_double.install()
    try:
        from b200tetris import tetris_environment_vector
        te, tev = _reference_front_end()
        ref_mod, ref_time = refmod.load("B")
        te.tetris_env = ref_mod
        n, T, P = 8, 120, extra.get("n_players", 2)
        ours = tetris_environment_vector(n, settings=dict(extra, seed=4242))
        ref_time(0)
        ref = tev.tetris_environment_vector(n, te.tetris_environment, settings=dict({"presets": ["default"], "render": False}, **extra))

        def seeds_of_current_round(ids):
            ids = np.asarray(ids)
            ours._episode[ids] -= 1
            s = ours._seeds(ids)
            ours._episode[ids] += 1
            return s

        def reset_reference(ids):          # the reference seeds a round from the clock: give each env the seed ours used
            for i, s in zip(ids, seeds_of_current_round(ids)):
                ref_time(int(s))
                ref.reset(env=[int(i)])
        reset_reference(list(range(n)))
        cur = np.arange(n) & 1            # (with more than two players only the first two can move through the reference's front-end:
                                          #  its per-player lists have two entries, tetris_environment.py:28-29)
        n_sims = n_done = 0
        for t in range(T):
            cur = 1 - cur
            players = [int(p) for p in cur]
            s_ref, s_our = ref.get_state(), ours.get_state()
            a_our = ours.get_actions(player=players)
            actions_ref, actions_our = [], []
            for i in range(n):
                x, y = dicts(s_ref[i]), dicts(s_our[i])
                if t == 0:                  # combo.remaining is uninitialised in a fresh reference handle
                    for d in x + y:
                        d["combo_time"] = 0
                assert x == y, (t, i, [(k, x[p][k], y[p][k]) for p in range(P) for k in x[p] if x[p][k] != y[p][k]])
                assert [int(v) for v in s_ref[i].is_dead] == [int(v) for v in s_our[i].is_dead]
                a_ref = ref.envs[i].get_actions(s_ref[i], player=players[i])
                assert [list(a) for a in a_ref] == [list(a) for a in a_our[i]], (t, i)
                pick = int(harness.splitmix64(9, i, t) % np.uint64(len(a_ref)))
                actions_ref.append(a_ref[pick]); actions_our.append(a_our[i][pick])
                if (t + i) % 9 == 0:        # the world model, both finalize modes
                    fin = bool((t // 9) & 1)
                    sims_ref = ref.envs[i].simulate_actions(a_ref, player=players[i], finalize=fin)
                    sims_our = ours.simulate_all_actions(env=[i], player=[players[i]], finalize=fin)[0]
                    assert len(sims_ref) == len(sims_our)
                    for u, v in zip(sims_ref, sims_our):
                        du, dv = dicts(u), dicts(v)
                        if t == 0:
                            for d in du + dv:
                                d["combo_time"] = 0
                        assert du == dv, (t, i)
                    n_sims += len(sims_ref)
            r_ref, d_ref = ref.perform_action(actions_ref, player=players)
            r_our, d_our = ours.perform_action(actions_our, player=players)
            assert [float(r()) for r in r_ref] == [float(r()) for r in r_our], t
            assert [bool(d) for d in d_ref] == [bool(d) for d in d_our], t
            for i in range(n):              # get_winner: the reference's reads an unbound attribute; its rule is :118-125
                dead = [int(ref.envs[i].backend.states[p].dead[0]) for p in range(P)]
                want = None if not d_ref[i] else ([p for p in range(P) if not dead[p]] + [666])[0]
                assert ours.get_winner(env=[i])[0] == want
            ended = [i for i, d in enumerate(d_ref) if d]
            if ended:
                ours.reset(env=ended)
                reset_reference(ended)
                n_done += len(ended)
            info = ours.get_info()
            for i in range(n):
                assert int(info[i]["rounds_played"]) == ref.envs[i].rounds_played - 1      # the reference counted its construction-time reset too
        assert n_sims > 300
    finally:
        facade_double.uninstall()
