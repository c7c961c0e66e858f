"""Host logic of b200tetris.tetris_environment_vector that needs no reference tree: index handling (Python indexing rules like the
reference's `self.envs[i]`), key lists longer than a sequence, the dead flags of simulated outcomes, copy() bookkeeping.  On the
CPU the vector env sits on tests/facade_double.py (test infrastructure); `-m gpu` repeats the cases on the CUDA engine."""
import numpy as np
import pytest

import harness
import pyoracle


def _cases(make):
    ours = make(6, dict(seed=77))
    # ---- env indices: negatives count from the end, anything else raises IndexError (tools/utils.py:10-31 -> self.envs[i])
    a_last = ours.get_actions(env=[-1], player=[0])[0]
    a_five = ours.get_actions(env=[5], player=[0])[0]
    assert [list(a) for a in a_last] == [list(a) for a in a_five]
    for bad in ([6], [-7], 6, [0, 99]):
        with pytest.raises(IndexError):
            ours.get_actions(env=bad, player=0)
        with pytest.raises(IndexError):
            ours.reset(env=bad)
    with pytest.raises(IndexError):
        ours.get_actions(env=[0], player=[2])

    # ---- a key list longer than the engine's sequence capacity is applied in pieces: same result as the oracle fed the whole list
    cap = ours.engine.lmax
    long_keys = ([1, 3] * cap)[: 2 * cap + 5] + [8, 8, 3, 7]          # > 2 * capacity, ends with turn, turn, right, hard drop
    assert cap < len(long_keys) <= 255
    H, W = ours.settings["game_size"]
    orc = pyoracle.OracleVector(1, H, W)
    ours._episode[0] -= 1
    seed = ours._seeds([0])
    ours._episode[0] += 1
    orc.reset(seed)
    orc_r, orc_d, _ = orc.step(np.array([0], np.uint8), np.array([long_keys], np.uint8), np.array([len(long_keys)], np.uint8))
    r, d = ours.perform_action([long_keys], env=[0], player=[0])
    st = ours.get_state(env=[0])[0]
    of, ov, op = orc.observe()
    for p in (0, 1):
        assert np.array_equal(st[p]["field"], of[0, p]), p
        assert [int(st[p]["x"][0]), int(st[p]["y"][0]), st[p]["piece_idx"]] == [int(ov[0, p, 0]), int(ov[0, p, 1]), int(op[0, p])]
    assert float(r[0]()) == float(orc_r[0]) and d[0] == bool(orc_d[0])

    # ---- simulated outcomes carry the dead flags: fill the acting player's board to the ceiling so that some placement kills it
    env = make(2, dict(seed=5))
    pol = harness.Policy("random", kmax=96, lmax=32)
    killed = 0
    for t in range(400):
        players = [t & 1, t & 1]
        sims = env.simulate_all_actions(player=players, finalize=True)
        acts = env.get_actions(player=players)
        for i in range(2):
            for s in sims[i]:
                if any(s.is_dead):
                    killed += 1
                    assert s.is_dead[players[i]] == 1 or s.is_dead[1 - players[i]] == 1
        pick = [a[int(harness.splitmix64(3, i, t) % np.uint64(len(a)))] for i, a in enumerate(acts)]
        _, dones = env.perform_action(pick, player=players)
        ended = [i for i, dn in enumerate(dones) if dn]
        if killed > 3:
            break
        if ended:
            env.reset(env=ended)
    assert killed > 3, "random play must reach boards where a placement kills the mover"

    # ---- copy(): the copy continues the original's rounds (seed sequence, counters, reward totals)
    c = env.copy()
    assert np.array_equal(c._episode, env._episode) and np.array_equal(c.rounds_played, env.rounds_played) and c._seed0 == env._seed0
    assert [[float(x()) for x in rr] for rr in c.tot_reward] == [[float(x()) for x in rr] for rr in env.tot_reward]
    env.reset(); c.reset()
    sa, sb = env.get_state(), c.get_state()
    for i in range(2):
        for p in (0, 1):
            assert np.array_equal(sa[i][p]["field"], sb[i][p]["field"]) and sa[i][p]["piece_idx"] == sb[i][p]["piece_idx"]
            assert np.array_equal(sa[i][p]["nextpiece"], sb[i][p]["nextpiece"])


def test_vector_env_host_logic_on_the_cpu_double():
    import facade_double
    facade_double.install()
    try:
        from b200tetris import tetris_environment_vector
        _cases(lambda n, s: tetris_environment_vector(n, settings=s))
    finally:
        facade_double.uninstall()


@pytest.mark.gpu
def test_vector_env_host_logic_on_the_engine():
    from b200tetris import tetris_environment_vector
    _cases(lambda n, s: tetris_environment_vector(n, settings=s))
