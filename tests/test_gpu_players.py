"""n_players != 2 on the CUDA engine: PythonHandle with 1, 3 and 4 players — the float split of sent lines among
all the others, solo rounds, the reward rule's Python index — through the C ABI against the oracle, which tests/
test_oracle_vs_reference.py pins to the compiled reference for the same player counts.  Every observable after every call,
placement lists and the world model included."""
import numpy as np
import pytest

import harness
import pyoracle

pytestmark = pytest.mark.gpu

ALL7 = (0, 1, 2, 3, 4, 5, 6)
CASES = [  # n_players, kind, H, W, pieces, games, steps
    (1, "random", 22, 10, ALL7, 96, 200),
    (1, "greedy", 22, 10, ALL7, 32, 300),
    (3, "random", 22, 10, ALL7, 128, 260),
    (3, "greedy", 22, 10, ALL7, 48, 420),
    (3, "junk", 22, 10, ALL7, 64, 200),
    (3, "press_key", 22, 10, ALL7, 48, 600),
    (4, "greedy", 22, 10, ALL7, 32, 420),
    (4, "random", 16, 8, (0, 6), 64, 200),
    (3, "greedy", 40, 20, ALL7, 8, 160),
    (4, "junk", 25, 12, ALL7, 16, 160),
]


def _engine(n, H, W, pieces, P, **kw):
    from b200vec import B200Vector
    return B200Vector(n, H, W, pieces, n_players=P, **kw)


@pytest.mark.parametrize("P,kind,H,W,pieces,n,steps", CASES)
def test_rollout_matches_oracle_for_other_player_counts(P, kind, H, W, pieces, n, steps):
    kmax, lmax = (256, 64) if W > 10 else (96, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces, n_players=P)
    dev = _engine(n, H, W, pieces, P)
    stats = harness.rollout_pair(lead, dev, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, n_players=P))
    assert stats["steps"] == n * steps
    st = dev.e.stats()
    assert st["garbage_queue_overflow"] == 0 and st["capacity_overflow"] == 0, st
    assert st["env_steps"] > 0
    if kind == "greedy" and P > 1 and W >= 8:
        assert stats["lines"] > 0 and stats["garbage"] > 0


@pytest.mark.parametrize("P", [1, 3, 4])
@pytest.mark.parametrize("finalize", [True, False])
def test_world_model_matches_oracle_for_other_player_counts(P, finalize):
    n, kmax, lmax = 40, 96, 32
    lead = pyoracle.OracleVector(n, 22, 10, ALL7, n_players=P)
    dev = _engine(n, 22, 10, ALL7, P)
    pol = harness.Policy("random", kmax=kmax, lmax=lmax, n_players=P)
    harness.rollout_pair(lead, dev, 30, pol)
    for t in range(4):
        player = ((np.arange(n) + t) % P).astype(np.uint8)
        a = lead.expand(player, finalize=finalize, kmax=kmax, lmax=lmax)
        b = dev.expand(player, finalize=finalize, kmax=kmax, lmax=lmax)
        for k in ("K", "keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
            assert np.array_equal(a[k], b[k]), (t, k)
        keys, klen, _ = pol.choose(lead, 100 + t, np.arange(n), player)
        ra, rb = lead.step(player, keys, klen), dev.step(player, keys, klen)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)


def test_other_player_counts_through_the_device_face_and_the_guards():
    """placement-index steps with fused engine-seeded resets, snapshots, export / import and the two-player-only calls"""
    import torch
    from b200tetris import B2TError, Engine, _lib
    dev = torch.device("cuda:0")
    n, P = 500, 3
    e = Engine(n, 22, 10, n_players=P, seed=5)
    e.reset(harness.seeds_for(np.arange(n), np.zeros(n, int)))
    K = torch.empty(n, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    done = torch.empty(n, dtype=torch.uint8, device=dev); lines = torch.empty((n, P), dtype=torch.uint8, device=dev)
    field = torch.empty((n, P, 22, 10), dtype=torch.uint8, device=dev)
    ended = 0
    for t in range(120):
        pl = torch.full((n,), t % P, dtype=torch.uint8, device=dev)
        e.enumerate_dev(pl, K)
        sel = (torch.randint(0, 1 << 30, (n,), device=dev, generator=gen) % K.to(torch.int64)).to(torch.int32)
        e.step_dev(pl, placement=sel, auto_reset=True, done=done, lines=lines, field=field)
        ended += int(done.sum().item())
    assert ended > 0 and field.sum().item() > 0
    st = e.stats()
    assert st["env_steps"] == 120 * n and st["episodes"] == ended and st["capacity_overflow"] == 0
    e.snapshot(0)
    before = e.export_states()
    assert before.shape == (n, e.state_bytes) and e.state_bytes == 16 * ((3 * (10 + 7 + 24) + 3) // 4)
    e.step(np.zeros(n, np.uint8), np.full((n, 16), 7, np.uint8), np.ones(n, np.uint8))
    assert not np.array_equal(before, e.export_states())
    e.restore(0)
    assert np.array_equal(before, e.export_states())
    e.import_states(before[::-1].copy())
    assert np.array_equal(before[::-1], e.export_states())
    for call in (lambda: e.get_game(0), lambda: Engine(4, 22, 10, n_players=3, tiles=True), lambda: Engine(4, 22, 10, n_players=5),
                 lambda: e.step_dev(torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros((n, 8), dtype=torch.uint8, device=dev), None, packed_keys=True),
                 lambda: e.observe_dev(field=torch.empty((n, P, 22, 10), dtype=torch.float16, device=dev))):
        with pytest.raises(B2TError) as err:
            call()
        assert err.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(B2TError):
        e.poke_dead(0, 3, True)
    e.poke_dead(0, 2, True)
    assert e.scal([0])[0, 2, 7] == 1
    e.close()
