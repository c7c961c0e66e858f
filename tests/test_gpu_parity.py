"""GPU parity gate: the CUDA engine, called through the C ABI (libb200tetris.so), against the oracle on the same seeded
inputs, against the committed reference goldens, and — at BASELINE.json's full sizes — through size-independent
properties.  Bit-exact everywhere (integer/byte work)."""
import glob
import os

import numpy as np
import pytest

import harness
import pyoracle

pytestmark = pytest.mark.gpu

ALL7 = (0, 1, 2, 3, 4, 5, 6)
ROLLOUTS = [
    ("random", 22, 10, ALL7, 256, 160, 0),
    ("greedy", 22, 10, ALL7, 48, 400, 0),
    ("greedy", 22, 10, ALL7, 32, 300, 23),
    ("junk", 22, 10, ALL7, 64, 300, 11),
    ("sventon", 22, 10, ALL7, 64, 150, 0),
    ("press_key", 22, 10, ALL7, 64, 700, 0),     # action_type "press_key": one key per step, gravity / lockdown place the pieces
    ("press_key", 16, 8, (0, 6), 32, 500, 30),
    ("greedy", 25, 12, ALL7, 24, 300, 0),          # tallest field on 32-bit columns, W > 10
    ("random", 25, 12, ALL7, 64, 200, 0),
    ("junk", 25, 12, ALL7, 32, 200, 9),
    ("random", 16, 8, (0, 6), 96, 200, 0),
    ("greedy", 16, 8, (0, 6), 32, 300, 0),
    ("random", 40, 20, ALL7, 40, 200, 0),
    ("greedy", 40, 20, ALL7, 8, 200, 0),
    ("junk", 40, 20, ALL7, 16, 150, 7),
    ("random", 10, 6, (2, 3), 64, 150, 0),
    ("greedy", 10, 6, (2, 3, 4, 5), 32, 200, 0),
    ("junk", 12, 5, (4,), 33, 150, 5),
]


def _engine(n, H=22, W=10, pieces=ALL7, **kw):
    from b200vec import B200Vector
    return B200Vector(n, H, W, pieces, **kw)


@pytest.mark.parametrize("kind,H,W,pieces,n,steps,solo", ROLLOUTS)
def test_rollout_matches_oracle(kind, H, W, pieces, n, steps, solo):
    kmax, lmax = (256, 64) if W > 10 else (96, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces)
    dev = _engine(n, H, W, pieces)
    stats = harness.rollout_pair(lead, dev, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, solo_period=solo))
    assert stats["steps"] == n * steps
    st = dev.e.stats()
    assert st["garbage_queue_overflow"] == 0 and st["capacity_overflow"] == 0, st
    if kind == "greedy" and W >= 8:
        assert stats["lines"] > 0 and stats["garbage"] > 0


TIMINGS = [  # kind, time_elapsed_each_action, games, steps: gravity tiers (every > 3000 ms: -10 / -5 / -2 / -1, DropDelay.cpp:3-21),
    ("greedy", 3100, 32, 420),     # lockdown and garbage / combo timers at other cadences than the default 400 ms
    ("press_key", 3100, 48, 300),
    ("press_key", 1000, 48, 300),
    ("greedy", 50, 32, 200),
    ("junk", 7, 48, 200),
]


@pytest.mark.parametrize("kind,time_ms,n,steps", TIMINGS)
def test_other_time_steps(kind, time_ms, n, steps):
    """settings["time_elapsed_each_action"] other than 400 ms, per call (b2t_step_args.time_ms)"""
    lead = pyoracle.OracleVector(n)
    dev = _engine(n)
    stats = harness.rollout_pair(lead, dev, steps, harness.Policy(kind, kmax=96, lmax=32), time_ms=time_ms)
    assert stats["steps"] == n * steps


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_replays_reference_golden(path):
    g = np.load(path)
    eng = _engine(int(g["n"]), int(g["H"]), int(g["W"]), tuple(int(x) for x in g["pieces"]),
                  n_players=int(g["n_players"]) if "n_players" in g else 2)
    harness.replay_golden(eng, g)


def test_config_c2_4096_games_bit_exact():
    """BASELINE config 2: 4096 parallel 2-player 10x22 games, random placements, diffed against the oracle every step."""
    n, steps = 4096, 96
    lead = pyoracle.OracleVector(n)
    dev = _engine(n)
    stats = harness.rollout_pair(lead, dev, steps, harness.Policy("random", kmax=96, lmax=32))
    assert stats["dones"] > 1000


@pytest.mark.parametrize("finalize", [True, False])
@pytest.mark.parametrize("H,W,pieces", [(22, 10, ALL7), (16, 8, (0, 6)), (40, 20, ALL7)])
def test_world_model_matches_oracle(H, W, pieces, finalize):
    """simulate_all_actions (b2t_expand) on boards reached by greedy play."""
    n = 24 if W <= 10 else 6
    kmax, lmax = (256, 64) if W > 10 else (96, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces)
    dev = _engine(n, H, W, pieces)
    pol = harness.Policy("greedy", kmax=kmax, lmax=lmax)
    checked = [0]

    def check(t, *_):
        if t % 9 != 4:
            return
        player = pol.players(n, t + 1, np.arange(n))
        a = lead.expand(player, finalize=finalize, kmax=kmax, lmax=lmax)
        b = dev.expand(player, finalize=finalize, kmax=kmax, lmax=lmax)
        assert a["bad"] == 0
        assert b["total"] == int(a["K"].sum())
        for k in ("K", "keys", "key_len", "reward", "done", "lines", "field", "vec", "piece_idx"):
            harness.assert_same("expand." + k, t, a[k], b[k])
        checked[0] += int(a["K"].sum())

    harness.rollout_pair(lead, dev, 100, pol, on_step=check)
    assert checked[0] > 500


def test_bar_null_moves_false():
    n = 64
    lead = pyoracle.OracleVector(n)
    dev = _engine(n, bar_null_moves=False)
    s = harness.seeds_for(np.arange(n), np.zeros(n, int))
    lead.reset(s); dev.reset(s)
    pl = (np.arange(n) & 1).astype(np.uint8)
    a = lead.enumerate(pl, bar_null_moves=False, kmax=96, lmax=32)
    b = dev.enumerate(pl, bar_null_moves=False, kmax=96, lmax=32)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)


def test_stream_tables_match_oracle():
    from b200tetris import Engine
    for pieces in (ALL7, (0, 6), (2, 3), (2, 3, 4, 5)):
        e = Engine(1, pieces=pieces, piece_stream_len=700, hole_stream_len=64)
        for seed in (0, 1, 1000, 32767, 32768, 40000, 65535):
            p, h = e.debug_streams(seed, 700)
            assert np.array_equal(p, pyoracle.piece_stream(seed, pieces, 700)), (pieces, seed)
            assert np.array_equal(h[:64], pyoracle.hole_stream(seed, 10, 64)), (pieces, seed)
        e.close()


def test_unaligned_key_width_and_subsets():
    """keys arrays whose width is not a multiple of 16 take the byte path; game subsets in arbitrary order."""
    n = 80
    lead = pyoracle.OracleVector(n)
    dev = _engine(n)
    s = harness.seeds_for(np.arange(n), np.zeros(n, int))
    lead.reset(s); dev.reset(s)
    rng = np.random.RandomState(3)
    for t in range(40):
        ids = rng.permutation(n)[: rng.randint(1, n)].astype(np.int32)
        pl = rng.randint(0, 2, len(ids)).astype(np.uint8)
        keys = rng.randint(0, 12, (len(ids), 21)).astype(np.uint8)
        ln = rng.randint(0, 22, len(ids)).astype(np.uint8)
        ra = lead.step(pl, keys, ln, ids)
        rb = dev.step(pl, keys, ln, ids)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        for x, y in zip(lead.observe(), dev.observe()):
            if t > 0:
                assert np.array_equal(x, y)
        d = ids[ra[1].astype(bool)]
        if len(d):
            s2 = rng.randint(0, 1 << 16, len(d)).astype(np.uint32)
            lead.reset(s2, d); dev.reset(s2, d)


# ------------------------------------------------------------------------------------------ full-size properties
def _digest_dev(t):
    return harness.digest(t.cpu().numpy())


@pytest.mark.parametrize("n,H,W,pieces,steps", [(1 << 20, 22, 10, ALL7, 24), (1 << 18, 16, 8, (0, 6), 16), (1 << 18, 40, 20, ALL7, 12)],
                         ids=["c4_1m_22x10", "c5_256k_16x8_p06", "c5_256k_40x20"])
def test_full_size_properties(n, H, W, pieces, steps):
    """BASELINE configs 4 and 5 at their full sizes (2^20 games 22x10; 2^18 games 16x8 with pieces [0,6] and 40x20): (a) results do
    not depend on how the batch is split into calls, (b) a step from a snapshot is reproducible, (c) stepping by placement
    index == stepping by the enumerated keys, (d) the sampled games agree with the oracle."""
    import torch
    from b200tetris import Engine
    dev = torch.device("cuda:0")
    e = Engine(n, H, W, pieces)
    gid = torch.arange(n, device=dev, dtype=torch.int64)
    seeds = torch.from_numpy(harness.seeds_for(np.arange(n), np.zeros(n, int)).astype(np.int64)).to(dev).to(torch.int32)
    e.reset_dev(seeds.view(torch.int32))
    K = torch.empty(n, dtype=torch.int32, device=dev)
    reward = torch.empty(n, dtype=torch.int8, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    lines = torch.empty((n, 2), dtype=torch.uint8, device=dev)
    field = torch.empty((n, 2, H, W), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
    sample = np.arange(0, n, 4099 if W <= 10 else 16411)[:250].astype(np.int32)
    lead = pyoracle.OracleVector(len(sample), H, W, pieces)
    lead.reset(harness.seeds_for(sample, np.zeros(len(sample), int)))
    okmax, olmax = (256, 128) if W > 10 else (96, 64)
    for t in range(steps):
        player = ((gid + t) & 1).to(torch.uint8)
        e.enumerate_dev(player, K)
        h = torch.from_numpy(harness.splitmix64(1, np.arange(n), t).astype(np.int64)).to(dev)
        sel = (h.remainder(1 << 30) % K.to(torch.int64)).to(torch.int32)
        if t == steps // 2:
            e.snapshot(0)
        e.step_dev(player, placement=sel, reward=reward, done=done, lines=lines, field=field, vec=vec)
        torch.cuda.synchronize()
        if t == steps // 2:
            ref = (_digest_dev(reward), _digest_dev(done), _digest_dev(field), _digest_dev(vec))
            # (b) reproducible from the snapshot, (a) in two half-batches given as explicit index lists
            e.restore(0)
            e.enumerate_dev(player, K)
            half = n // 2
            for lo, hi in ((0, half), (half, n)):
                idx = torch.arange(lo, hi, device=dev, dtype=torch.int32)
                e.step_dev(player[lo:hi].contiguous(), idx=idx, placement=sel[lo:hi].contiguous(), reward=reward[lo:hi], done=done[lo:hi],
                           lines=lines[lo:hi], field=field[lo:hi], vec=vec[lo:hi])
            torch.cuda.synchronize()
            assert ref == (_digest_dev(reward), _digest_dev(done), _digest_dev(field), _digest_dev(vec))
        # (e) oracle on the sample, driven by the same placement indices
        Ko, keys_o, klen_o = lead.enumerate(player[sample].cpu().numpy(), kmax=okmax, lmax=olmax)[:3]
        assert np.array_equal(Ko, K[sample].cpu().numpy())
        so = sel[sample].cpu().numpy()
        ro = lead.step(player[sample].cpu().numpy(), keys_o[np.arange(len(sample)), so], klen_o[np.arange(len(sample)), so])
        assert np.array_equal(ro[0], reward[sample].cpu().numpy()) and np.array_equal(ro[1], done[sample].cpu().numpy())
        fo = lead.observe()
        assert np.array_equal(fo[0], field[sample].cpu().numpy())
        if t > 0:
            assert np.array_equal(fo[1], vec[sample].cpu().numpy())
        d = done.bool()
        if d.any():
            ids = torch.nonzero(d).flatten().to(torch.int32)
            s2 = torch.from_numpy(harness.seeds_for(ids.cpu().numpy(), np.full(ids.numel(), t + 1)).astype(np.int64)).to(dev).to(torch.int32)
            e.reset_dev(s2, ids)
            ds = np.nonzero(ro[1])[0]
            if len(ds):
                lead.reset(harness.seeds_for(sample[ds], np.full(len(ds), t + 1)), ds)
    torch.cuda.synchronize()
    st = e.stats()
    assert st["garbage_queue_overflow"] == 0 and st["capacity_overflow"] == 0 and st["env_steps"] >= steps * n - st["episodes"] * 2
    e.close()


def test_fused_reset_equals_step_then_reset():
    n = 2048
    a = _engine(n); b = _engine(n)
    s = harness.seeds_for(np.arange(n), np.zeros(n, int))
    a.reset(s); b.reset(s)
    pol = harness.Policy("random", kmax=96, lmax=32)
    gids = np.arange(n)
    for t in range(70):
        pl = pol.players(n, t, gids)
        keys, ln, _ = pol.choose(a, t, gids, pl)
        s2 = harness.seeds_for(gids, np.full(n, t + 1))
        ra = a.e.step(pl, keys, ln, reset_seeds=s2, want_obs=True)
        rb = b.e.step(pl, keys, ln)
        d = np.nonzero(rb[1])[0]
        if len(d):
            b.reset(s2[d], d)
        ob = b.observe()
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
        assert np.array_equal(ra[3][0], ob[0]) and np.array_equal(ra[3][2], ob[2])
        if t > 0:
            assert np.array_equal(ra[3][1], ob[1])


def test_snapshot_export_import_roundtrip():
    n = 300
    a = _engine(n)
    s = harness.seeds_for(np.arange(n), np.zeros(n, int))
    a.reset(s)
    pol = harness.Policy("random", kmax=96, lmax=32)
    gids = np.arange(n)
    for t in range(10):
        pl = pol.players(n, t, gids)
        keys, ln, _ = pol.choose(a, t, gids, pl)
        a.step(pl, keys, ln)
    before = a.observe()
    a.e.snapshot(1)
    blob = a.e.export_state(17)
    assert blob.size == a.e.state_bytes
    pl = pol.players(n, 10, gids)
    keys, ln, _ = pol.choose(a, 10, gids, pl)
    r1 = a.step(pl, keys, ln); after = a.observe()
    a.e.restore(1)
    for x, y in zip(before, a.observe()):
        assert np.array_equal(x, y)
    r2 = a.step(pl, keys, ln)
    for x, y in zip(r1, r2):
        assert np.array_equal(x, y)
    for x, y in zip(after, a.observe()):
        assert np.array_equal(x, y)
    a.e.import_state(18, blob)       # PythonHandle.set(other): game 18 becomes a copy of (old) game 17
    a.e.restore(1)
    a.e.import_state(18, blob)
    o = a.observe()
    assert np.array_equal(o[0][18], before[0][17]) and np.array_equal(o[1][18], before[1][17])
    # state.lock(): dead players' keys are ignored (data_types/state.py:9-17, PythonHandle.cpp:141)
    a.e.poke_dead(5, 0, True); a.e.poke_dead(5, 1, True)
    f0 = a.observe(np.array([5]))[0]
    a.step(np.array([0], np.uint8), np.array([[7] + [0] * 15], np.uint8), np.array([1], np.uint8), np.array([5]), finalize=False)
    assert np.array_equal(f0, a.observe(np.array([5]))[0])


def test_vector_env_front_end():
    """The reference-shaped surface (tetris_environment_vector names) against the oracle for a short random rollout."""
    from b200tetris import tetris_environment_vector, action
    n = 32
    env = tetris_environment_vector(n, settings={"seed": 1234, "render": False})
    lead = pyoracle.OracleVector(n)
    lead.reset(env._seeds(np.arange(n)) if False else _front_seeds(env, n))
    rng = np.random.RandomState(0)
    cur = rng.choice([0, 1], n)
    for t in range(60):
        cur = 1 - cur
        acts = env.get_actions(player=list(cur))
        Ko, keys_o, klen_o = lead.enumerate(cur.astype(np.uint8), kmax=96, lmax=32)[:3]
        assert [len(a) for a in acts] == list(Ko)
        sel = [rng.randint(0, len(a)) for a in acts]
        chosen = [a[s] for a, s in zip(acts, sel)]
        for i in range(n):
            assert list(chosen[i]) == list(keys_o[i, sel[i], :klen_o[i, sel[i]]])
        rewards, dones = env.perform_action(chosen, player=list(cur))
        ro = lead.step(cur.astype(np.uint8), keys_o[np.arange(n), sel], klen_o[np.arange(n), sel])
        assert [int(r()) for r in rewards] == list(ro[0]) and dones == [bool(d) for d in ro[1]]
        st = env.get_state()
        fo = lead.observe()
        for i in range(n):
            for p in (0, 1):
                d = st[i][p]
                assert np.array_equal(d["field"], fo[0][i, p]) and d["piece_idx"] == fo[2][i, p]
                assert np.array_equal(d["nextpiece"], fo[1][i, p, 5:12])
        ids = [i for i, d in enumerate(dones) if d]
        if ids:
            before = env._episode[ids].copy()
            env.reset(env=ids)
            lead.reset(_front_seeds(env, n)[ids], np.array(ids, np.int32))
            assert (env._episode[ids] == before + 1).all()
    sims = env.simulate_all_actions(env=[0, 1], player=[int(1 - cur[0]), int(1 - cur[1])])
    assert len(sims) == 2 and len(sims[0]) == len(env.get_actions(env=[0], player=int(1 - cur[0]))[0])


def _front_seeds(env, n):
    """seeds the front-end used for the CURRENT episode of every env (episode counter was already advanced)"""
    ids = np.arange(n)
    env._episode[ids] -= 1
    s = env._seeds(ids)
    env._episode[ids] += 1
    return s


def test_perspective_order_and_select_in_a_cuda_graph():
    """SURVEY §8f-1 building blocks: observation tensors in the (me, opponent) order of agents/agent_utils/state_unpack.py,
    the segmented argmax over the caller's placement scores, and the whole enumerate -> expand -> select -> step chain
    captured in one CUDA graph (no host round trip) reproducing the eager result."""
    import torch
    from b200tetris import Engine
    n = 4096
    dev = torch.device("cuda:0")
    mk = lambda: Engine(n)
    a, b = mk(), mk()
    seeds = torch.from_numpy(harness.seeds_for(np.arange(n), np.zeros(n, int)).astype(np.int64)).to(dev).to(torch.int32)
    a.reset_dev(seeds); b.reset_dev(seeds)
    cap = n * 64
    K = torch.empty(n, dtype=torch.int32, device=dev); offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
    pf = torch.empty((cap, 2, 22, 10), dtype=torch.uint8, device=dev); pv = torch.empty((cap, 2, 19), dtype=torch.uint8, device=dev)
    pp = torch.empty((cap, 2), dtype=torch.uint8, device=dev)
    sel = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev); reward = torch.empty(n, dtype=torch.int8, device=dev)
    field = torch.empty((n, 2, 22, 10), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
    pidx = torch.empty((n, 2), dtype=torch.uint8, device=dev)
    player = torch.zeros(n, dtype=torch.uint8, device=dev)
    scores = torch.empty(cap, dtype=torch.float32, device=dev)

    def score_fn():   # a stand-in for the caller's value network: fewer filled cells of "me" in the top half + noise from the row index
        me = pf[:, 0, :11].sum((1, 2)).to(torch.float32)
        scores.copy_(-me + (torch.arange(cap, device=dev) % 7).to(torch.float32) * 0.01)

    def chain(e):
        e.expand_dev(player, K, offs, cap, field=pf, vec=pv, piece_idx=pp, perspective=True)
        score_fn()
        e.select_dev(scores, offs, sel)
        e.step_dev(player, placement=sel, reward=reward, done=done, field=field, vec=vec, piece_idx=pidx, perspective=True)

    # perspective order == swapped standard order
    for t in range(6):
        player.fill_(t & 1)
        chain(a)
        torch.cuda.synchronize()
        f2 = torch.empty_like(field); v2 = torch.empty_like(vec); p2 = torch.empty_like(pidx)
        a.observe_dev(field=f2, vec=v2, piece_idx=p2)
        torch.cuda.synchronize()
        if t & 1:
            f2, v2, p2 = f2.flip(1), v2.flip(1), p2.flip(1)
        assert torch.equal(f2, field) and torch.equal(v2, vec) and torch.equal(p2, pidx)
        f3 = torch.empty_like(field)
        a.observe_dev(field=f3, me=player)
        torch.cuda.synchronize()
        assert torch.equal(f3, field)
        # select == numpy argmax per segment
        o = offs.cpu().numpy(); s = scores.cpu().numpy(); want = np.array([np.argmax(s[o[i]:o[i + 1]]) for i in range(0, n, 97)])
        assert np.array_equal(want, sel.cpu().numpy()[::97])
        # eager twin
        chain(b)
    torch.cuda.synchronize()
    ref_a = a.observe()
    # graph-captured twin of two more plies
    g = torch.cuda.CUDAGraph()
    player.fill_(0)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain(a)               # warm-up on the side stream (allocations happen here, not during capture)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    a.restore(0) if False else None
    chain(b)                   # keep b in lock-step with a (same ply, eager)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        chain(a)
    torch.cuda.synchronize()
    # capture does not execute: replay once for a, run once eagerly for b
    g.replay()
    chain(b)
    torch.cuda.synchronize()
    xa, xb = a.observe(), b.observe()
    for u, v in zip(xa, xb):
        assert np.array_equal(u, v)
    a.close(); b.close()


def test_edge_cases_through_the_abi():
    """Empty and degenerate inputs, argument errors, finished / never-reset games, capacity reporting."""
    import ctypes as C
    import torch
    from b200tetris import B2TError, Engine, _lib
    # unsupported geometry / player count / bad arguments are reported, not crashed on
    with pytest.raises(B2TError) as e:
        Engine(4, 23, 11)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(B2TError) as e:
        Engine(4, n_players=5)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(B2TError) as e:
        Engine(0)
    assert e.value.code == _lib.ERR_ARG
    eng = Engine(5)
    with pytest.raises(IndexError):             # an id outside the vector (the reference's list of envs raises the same)
        eng.reset(np.zeros(6, np.uint32), np.arange(6))
    with pytest.raises(B2TError) as e:          # more rows than games
        eng.reset(np.zeros(6, np.uint32), np.array([0, 1, 2, 3, 4, 0]))
    assert e.value.code == _lib.ERR_ARG
    with pytest.raises(B2TError) as e:          # placement stepping before any enumerate
        eng.step(np.zeros(5, np.uint8), placement=np.zeros(5, np.int32))
    assert e.value.code == _lib.ERR_STATE
    # a never-reset game has no piece: its list is the null action, stepping it is harmless
    K, keys, klen = eng.enumerate(np.zeros(5, np.uint8))
    assert (K == 1).all() and (klen[:, 0] == 1).all() and (keys[:, 0, 0] == 0).all()
    # n == 0 everywhere
    z8, z32 = np.zeros(0, np.uint8), np.zeros(0, np.int32)
    eng.reset(np.zeros(0, np.uint32), z32)
    r = eng.step(z8, np.zeros((0, 16), np.uint8), z8, z32)
    assert all(x.shape[0] == 0 for x in r)
    assert eng.observe(z32)[0].shape == (0, 2, 22, 10)
    # single game, single-byte key rows (lmax = 1), zero-length key strings
    lead = pyoracle.OracleVector(5)
    s = harness.seeds_for(np.arange(5), np.zeros(5, int))
    eng.reset(s); lead.reset(s)
    for t in range(30):
        one = np.array([t % 5], np.int32)
        pl = np.array([t & 1], np.uint8)
        keys = np.array([[7 if t % 3 else 0]], np.uint8)
        ln = np.array([0 if t % 7 == 6 else 1], np.uint8)
        ra, rb = lead.step(pl, keys, ln, one), eng.step(pl, keys, ln, one)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    for x, y in zip(lead.observe(), eng.observe()):
        assert np.array_equal(x[..., :3] if x.ndim == 3 else x, y[..., :3] if y.ndim == 3 else y)
    # play one game to the end, then keep stepping it: done stays set, reward is repeated, state frozen (PythonHandle.cpp:139,150)
    g = np.array([2], np.int32)
    keys = np.array([[7] + [0] * 15], np.uint8); ln = np.array([1], np.uint8)
    for t in range(200):
        pl = np.array([t & 1], np.uint8)
        ra, rb = lead.step(pl, keys, ln, g), eng.step(pl, keys, ln, g)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
        if ra[1][0]:
            break
    assert ra[1][0] == 1
    frozen = eng.observe(g)
    for t in range(3):
        pl = np.array([t & 1], np.uint8)
        ra, rb = lead.step(pl, keys, ln, g), eng.step(pl, keys, ln, g)
        assert rb[1][0] == 1 and np.array_equal(ra[0], rb[0])
    for x, y in zip(frozen, eng.observe(g)):
        assert np.array_equal(x, y)
    # out-of-range placement index acts as the null action (action_list.py:27-33); negative too
    K, _, _ = eng.enumerate(np.zeros(5, np.uint8))
    before = eng.observe()[0].copy()
    eng.step(np.zeros(5, np.uint8), placement=np.array([K[0] + 5, -1, 10 ** 6, K[3], -7], np.int32), finalize=False)
    assert np.array_equal(before, eng.observe()[0])
    # expand reports the total even when the caller's capacity is too small, and fills what fits
    o_full = eng.expand(np.zeros(5, np.uint8))
    o_small = eng.expand(np.zeros(5, np.uint8), capacity=7)
    assert o_small["total"] == o_full["total"] > 7
    assert np.array_equal(o_small["field"][:7], o_full["field"][:7]) and np.array_equal(o_small["keys"][:7], o_full["keys"][:7])
    # statistics are plain sums
    st = eng.stats()
    assert st["garbage_queue_overflow"] == 0 and st["env_steps"] > 0 and st["episodes"] >= 1
    eng.close()
    # the largest supported field with its longest key strings, and the smallest
    for H, W, pieces in ((40, 20, ALL7), (10, 6, (0, 6)), (12, 5, (4,))):
        v = _engine(3, H, W, pieces)
        o = pyoracle.OracleVector(3, H, W, pieces)
        s = harness.seeds_for(np.arange(3), np.zeros(3, int))
        v.reset(s); o.reset(s)
        a = o.enumerate(np.zeros(3, np.uint8), kmax=256, lmax=v.e.lmax)
        b = v.enumerate(np.zeros(3, np.uint8), kmax=256, lmax=v.e.lmax)
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,pieces", [(22, 10, ALL7), (12, 5, (4,)), (40, 20, ALL7)])
def test_network_dtypes_equal_uint8_observations(H, W, pieces):
    """b2t_obs.dtype: field / vec written directly as half, bfloat16 or float — by step, observe and expand, in
    both player orders, for full and ragged last warps — hold exactly the uint8 values (cells 0/1, integers <= 255)."""
    import torch
    from b200tetris import Engine
    dev = torch.device("cuda:0")
    n = 1000 + 37                                   # not a multiple of 32: the last warp is ragged
    e = Engine(n, H, W, pieces)
    gids = np.arange(n)
    e.reset_dev(torch.from_numpy(harness.seeds_for(gids, np.zeros(n, int)).astype(np.int64)).to(dev).to(torch.int32))
    pol = harness.Policy("sventon", lmax=32)

    def tensors(rows, dt):
        return (torch.full((rows, 2, H, W), 7, dtype=dt, device=dev), torch.full((rows, 2, 19), 7, dtype=dt, device=dev))

    class L:                                       # the sventon policy only needs the width
        pass
    L.W = W
    for t in range(12):
        player = ((gids + t) & 1).astype(np.uint8)
        keys, klen, _ = pol.choose(L, t, gids, player)
        pl = torch.from_numpy(player).to(dev)
        f8, v8 = tensors(n, torch.uint8)
        e.step_dev(pl, torch.from_numpy(keys).to(dev), torch.from_numpy(klen).to(dev), field=f8, vec=v8, perspective=bool(t & 2))
        for dt in (torch.float16, torch.bfloat16, torch.float32):
            f, v = tensors(n, dt)
            if t & 2:
                e.observe_dev(field=f, vec=v, me=pl)
            else:
                e.observe_dev(field=f, vec=v)
            torch.cuda.synchronize()
            assert torch.equal(f.to(torch.float32), f8.to(torch.float32)), (t, dt)
            assert torch.equal(v.to(torch.float32), v8.to(torch.float32)), (t, dt)
            idx = torch.arange(5, n, 3, device=dev, dtype=torch.int32)      # a subset: 345 rows
            f, v = tensors(idx.numel(), dt)
            e.observe_dev(idx=idx, field=f, vec=v)
            fs, vs = tensors(idx.numel(), torch.uint8)
            e.observe_dev(idx=idx, field=fs, vec=vs)
            torch.cuda.synchronize()
            assert torch.equal(f.to(torch.float32), fs.to(torch.float32)) and torch.equal(v.to(torch.float32), vs.to(torch.float32))
    # the world model, and a step, writing the network's dtype directly
    K = torch.empty(n, dtype=torch.int32, device=dev); offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cap = n * e.kmax
    pl = torch.from_numpy((gids & 1).astype(np.uint8)).to(dev)
    f8, v8 = tensors(cap, torch.uint8)
    e.expand_dev(pl, K, offs, cap, field=f8, vec=v8, perspective=True)
    total = int(offs[-1].item())
    for dt in (torch.bfloat16, torch.float32):
        f, v = tensors(cap, dt)
        e.expand_dev(pl, K, offs, cap, field=f, vec=v, perspective=True)
        torch.cuda.synchronize()
        assert torch.equal(f[:total].to(torch.float32), f8[:total].to(torch.float32)) and torch.equal(v[:total].to(torch.float32), v8[:total].to(torch.float32))
        assert bool((f[total:] == 7).all()) and bool((v[total:] == 7).all())     # rows past the total are not touched
    e.snapshot(0)
    sel = (torch.arange(n, device=dev, dtype=torch.int32) * 7) % K
    e.step_dev(pl, placement=sel, field=f8[:n], vec=v8[:n])
    e.restore(0)
    fh, vh = tensors(n, torch.float16)
    e.step_dev(pl, placement=sel, field=fh, vec=vh)
    torch.cuda.synchronize()
    assert torch.equal(fh.to(torch.float32), f8[:n].to(torch.float32)) and torch.equal(vh.to(torch.float32), v8[:n].to(torch.float32))
    with pytest.raises(ValueError):
        e.observe_dev(field=fh, vec=v8[:n])
    e.close()


@pytest.mark.gpu
def test_tile_colour_planes():
    """B2T_FLAG_TILES: raw cell values 1..8 of states[p].field after every step vs the oracle (fused reset, garbage,
    line clears), through snapshot / restore and b2t_copy_game, with every other observable identical to an engine without planes."""
    from b200tetris import Engine
    from b200vec import B200Vector
    n = 96
    o = pyoracle.OracleVector(n)
    e = Engine(n, tiles=True)

    class V(B200Vector):
        def __init__(self, eng):
            self.n, self.H, self.W, self.P, self.bar_null, self.e = eng.n, eng.H, eng.W, 2, True, eng
    v = V(e)
    snap = {}

    def on_step(t, player, keys, klen, rb, fb, enum, other):
        if t % 25 == 0:
            f = o.raw()[0]
            for i in range(0, n, 7):
                assert np.array_equal(e.get_game(i)[1], f[i]), (t, i)
        if t == 120:
            e.snapshot(1)
            snap["f"] = o.raw()[0].copy()
    st = harness.rollout_pair(o, v, 260, harness.Policy("greedy", kmax=96, lmax=64), on_step=on_step)
    assert st["garbage"] > 0 and st["lines"] > 0 and st["dones"] > 0
    f = o.raw()[0]
    e.copy_game(3, e, 10)
    assert np.array_equal(e.get_game(3)[1], f[10])
    raw, fld = e.get_game(5)
    fld[0, 21, :] = [1, 2, 3, 4, 5, 6, 7, 8, 0, 3]
    e.set_game(5, raw, fld)
    assert np.array_equal(e.get_game(5)[1], fld)
    e.restore(1)
    for i in range(0, n, 5):
        assert np.array_equal(e.get_game(i)[1], snap["f"][i]), i
    e.close()


def test_key_formats_and_engine_chosen_seeds():
    """b2t_step_args.key_format / d_key_len == NULL / auto_reset: zero-terminated rows and 4-bit packed tokens step exactly like
    explicit-length byte keys (host and device entry points, aligned and odd row widths), and the fused reset with engine-chosen
    seeds equals a reset with auto_seed(cfg.seed, game, rounds started) computed here."""
    import torch
    from b200tetris import Engine
    n = 3000
    gids = np.arange(n)
    s0 = harness.seeds_for(gids, np.zeros(n, int))

    def auto_seed(base, game, episode):
        z = (base + game * 0x9E3779B1 + episode * 0x85EBCA77) & 0xFFFFFFFF
        z ^= z >> 15; z = (z * 0x2C1B3C6D) & 0xFFFFFFFF; z ^= z >> 12; z = (z * 0x297A2D39) & 0xFFFFFFFF; z ^= z >> 15
        return z & 0xFFFF
    for lmax in (16, 32, 24):
        a, b, c = Engine(n, seed=12345), Engine(n, seed=12345), Engine(n, seed=12345)
        for e in (a, b, c):
            e.reset(s0)
        episode = np.ones(n, np.int64)                 # every game has started one round
        pol = harness.Policy("junk", lmax=lmax, salt=5)
        dev = torch.device("cuda:0")
        for t in range(60):
            player = ((gids + t) & 1).astype(np.uint8)
            keys, klen, _ = pol.choose(None, t, gids, player)
            # the reference ignores key 0 and trailing keys equally: zero the tail so the row is its own length
            keys[np.arange(lmax)[None, :] >= klen[:, None]] = 0
            seeds = np.array([auto_seed(12345, int(g), int(ep)) for g, ep in zip(gids, episode)], np.uint32)
            ra = a.step(player, keys, klen, reset_seeds=seeds, want_obs=True)
            rb = b.step(player, keys, None, auto_reset=True, want_obs=True)
            packed = np.ascontiguousarray(keys[:, 0::2] | (keys[:, 1::2] << 4))
            f = torch.empty((n, 2, 22, 10), dtype=torch.uint8, device=dev); v = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
            rw = torch.empty(n, dtype=torch.int8, device=dev); dn = torch.empty(n, dtype=torch.uint8, device=dev)
            c.step_dev(torch.from_numpy(player).to(dev), torch.from_numpy(packed).to(dev), None, packed_keys=True, auto_reset=True,
                       reward=rw, done=dn, field=f, vec=v)
            torch.cuda.synchronize()
            for x, y in zip(ra[:3], rb[:3]):
                assert np.array_equal(x, y), (lmax, t)
            for x, y in zip(ra[3], rb[3]):
                assert np.array_equal(x, y), (lmax, t)
            assert np.array_equal(ra[0], rw.cpu().numpy()) and np.array_equal(ra[1], dn.cpu().numpy())
            assert np.array_equal(ra[3][0], f.cpu().numpy()) and np.array_equal(ra[3][1], v.cpu().numpy())
            episode += ra[1].astype(np.int64)
        assert episode.max() > 1
        for e in (a, b, c):
            e.close()


@pytest.mark.parametrize("H,W,pieces,n0,plies0", [(22, 10, ALL7, 2048, 240), (16, 8, (0, 6), 1024, 200), (40, 20, ALL7, 256, 120)])
def test_actor_loop_graph_equals_eager_and_oracle(H, W, pieces, n0, plies0):
    """b200tetris.ActorLoop: plies replayed as one CUDA graph == the same plies launched eagerly == the oracle driven
    by the chosen key strings, through line clears, garbage, combos and restarts (the engine-chosen seeds are a known counter hash)."""
    import torch
    from b200tetris import ActorLoop, Engine
    n, plies, base = int(os.environ.get("B2T_SOAK_N", n0)), int(os.environ.get("B2T_SOAK_PLIES", plies0)), 99   # env: a longer soak by hand
    okmax, olmax = (256, 128) if W > 10 else (96, 64)
    dev = torch.device("cuda:0")

    def score_fn(f, v, p, out):      # stand-in for the value network: low, flat, hole-free stacks of "me" (player axis 0 in perspective order)
        occ = f[:, 0].to(torch.float32)
        filled_above = (occ.cumsum(1) > 0).to(torch.float32)
        heights = filled_above.sum(1)                                       # [rows, W]
        holes = (filled_above - occ).sum((1, 2))
        out.copy_(-0.5 * heights.sum(1) - 0.4 * (heights[:, 1:] - heights[:, :-1]).abs().sum(1) - 5.0 * holes)

    def auto_seed(game, episode):
        z = (base + game * 0x9E3779B1 + episode * 0x85EBCA77) & 0xFFFFFFFF
        z ^= z >> 15; z = (z * 0x2C1B3C6D) & 0xFFFFFFFF; z ^= z >> 12; z = (z * 0x297A2D39) & 0xFFFFFFFF; z ^= z >> 15
        return z & 0xFFFF
    seeds = torch.from_numpy(harness.seeds_for(np.arange(n), np.zeros(n, int)).astype(np.int64)).to(dev).to(torch.int32)
    a, b = Engine(n, H, W, pieces, seed=base), Engine(n, H, W, pieces, seed=base)
    a.reset_dev(seeds); b.reset_dev(seeds)
    rpg = 40 if W <= 10 else 120
    la, lb = ActorLoop(a, score_fn, rows_per_game=rpg, use_graph=True), ActorLoop(b, score_fn, rows_per_game=rpg, use_graph=False)
    sample = np.arange(0, n, max(37, n // 96)).astype(np.int32)
    m = len(sample)
    lead = pyoracle.OracleVector(m, H, W, pieces)
    lead.reset(harness.seeds_for(sample, np.zeros(m, int)))
    episode = np.ones(m, np.int64)
    lines = garbage = restarts = 0
    for t in range(plies):
        la.step(); lb.step()
        torch.cuda.synchronize()
        assert not la.overflowed()
        if t % 8 == 0 or t < 4:
            for name in ("K", "choice", "reward", "done", "field", "vec", "piece"):
                assert torch.equal(getattr(la, name), getattr(lb, name)), (t, name)
            tot = int(la.offsets[-1].item())
            assert torch.equal(la.p_field[:tot], lb.p_field[:tot]) and torch.equal(la.scores[:tot], lb.scores[:tot])
        # the oracle, driven by the key strings the loop chose
        player = np.full(m, t & 1, np.uint8)
        Ko, keys_o, klen_o = lead.enumerate(player, kmax=okmax, lmax=olmax)[:3]
        assert np.array_equal(Ko, la.K.cpu().numpy()[sample]), t
        ch = la.choice.cpu().numpy()[sample]
        ro = lead.step(player, keys_o[np.arange(m), ch], klen_o[np.arange(m), ch])
        assert np.array_equal(ro[0], la.reward.cpu().numpy()[sample]) and np.array_equal(ro[1], la.done.cpu().numpy()[sample]), t
        ended = np.nonzero(ro[1])[0]
        if len(ended):                           # the fused reset used auto_seed(cfg.seed, game, rounds started)
            lead.reset(np.array([auto_seed(int(sample[i]), int(episode[i])) for i in ended], np.uint32), ended.astype(np.int32))
            episode[ended] += 1
            restarts += len(ended)
        fo, vo, _ = lead.observe()
        got_f, got_v = la.field.cpu().numpy()[sample], la.vec.cpu().numpy()[sample]
        if t & 1:                                # perspective order of the side that moved
            got_f, got_v = got_f[:, ::-1], got_v[:, ::-1]
        assert np.array_equal(fo, got_f), t
        if t > 0:
            assert np.array_equal(vo, got_v), t
        lines += int(ro[2].sum()); garbage += int((vo[..., 2] > 0).sum())
    assert la._graph is not None and (W > 10 or (lines > 0 and garbage > 0 and restarts > 0)), (lines, garbage, restarts)
    print("actor-loop soak: %d games x %d plies, %d sampled games vs oracle: lines %d, garbage observations %d, restarts %d" % (n, plies, m, lines, garbage, restarts))
    st = a.stats()
    assert st["env_steps"] == plies * n and st["capacity_overflow"] == 0 and st["garbage_queue_overflow"] == 0
    a.close(); b.close()


def test_sliced_host_step_with_a_game_subset():
    """b2t_step_host slices batches of >= 65 536 rows over two streams: a permuted 70 000-game subset given by index, with
    observations and scalars copied back, equals the unsliced device-pointer step of a twin engine."""
    import torch
    from b200tetris import Engine
    n, m = 1 << 17, 70000
    dev = torch.device("cuda:0")
    a, b = Engine(n), Engine(n)
    seeds = harness.seeds_for(np.arange(n), np.zeros(n, int))
    a.reset(seeds); b.reset(seeds)
    rng = np.random.default_rng(3)
    pol = harness.Policy("sventon", lmax=16)

    class L:
        W = 10
    for t in range(6):
        idx = rng.permutation(n)[:m].astype(np.int32)
        player = ((idx + t) & 1).astype(np.uint8)
        keys, klen, _ = pol.choose(L, t, idx, player)
        rs = harness.seeds_for(idx, np.full(m, t + 1))
        ra = a.step(player, keys, klen, idx, reset_seeds=rs, want_obs=True, want_scalars=True)
        f = torch.empty((m, 2, 22, 10), dtype=torch.uint8, device=dev); v = torch.empty((m, 2, 19), dtype=torch.uint8, device=dev)
        pi = torch.empty((m, 2), dtype=torch.uint8, device=dev); sc = torch.empty((m, 2, 8), dtype=torch.int32, device=dev)
        rw = torch.empty(m, dtype=torch.int8, device=dev); dn = torch.empty(m, dtype=torch.uint8, device=dev); ln = torch.empty((m, 2), dtype=torch.uint8, device=dev)
        b.step_dev(torch.from_numpy(player).to(dev), torch.from_numpy(keys).to(dev), torch.from_numpy(klen).to(dev), idx=torch.from_numpy(idx).to(dev),
                   reset_seeds=torch.from_numpy(rs.astype(np.int64)).to(dev).to(torch.int32), reward=rw, done=dn, lines=ln, field=f, vec=v, piece_idx=pi, scalars=sc)
        torch.cuda.synchronize()
        for x, y in zip(ra[:3], (rw, dn, ln)):
            assert np.array_equal(x, y.cpu().numpy()), t
        for x, y in zip(ra[3], (f, v, pi)):
            assert np.array_equal(x, y.cpu().numpy()), t
        assert np.array_equal(ra[4], sc.cpu().numpy()), t
    # the zero-copy form of the same call: every buffer pinned (b2t_host_alloc), no observation tensors -> the kernel reads the
    # keys and writes reward / done / lines through the caller's pages
    import ctypes as C
    from b200tetris import host_array
    from b200tetris.engine import _hp
    for t in range(6, 10):
        idx = rng.permutation(n)[:m].astype(np.int32)
        player = ((idx + t) & 1).astype(np.uint8)
        keys, klen, _ = pol.choose(L, t, idx, player)
        rs = harness.seeds_for(idx, np.full(m, t + 1))
        hp = {k: host_array(v.shape, v.dtype, write_combined=(k != "idx")) for k, v in (("idx", idx), ("player", player), ("keys", keys), ("klen", klen), ("rs", rs))}
        for k, v in (("idx", idx), ("player", player), ("keys", keys), ("klen", klen), ("rs", rs)):
            hp[k][...] = v
        o_rw, o_dn, o_ln = host_array(m, np.int8), host_array(m, np.uint8), host_array((m, 2), np.uint8)
        rc = a.lib.b2t_step_host(a._h, _hp(hp["idx"], C.c_int32), m, _hp(hp["player"], C.c_uint8), _hp(hp["keys"], C.c_uint8), _hp(hp["klen"], C.c_uint8),
                                 keys.shape[1], None, 1, 0, _hp(hp["rs"], C.c_uint32), _hp(o_rw, C.c_int8), _hp(o_dn, C.c_uint8), _hp(o_ln, C.c_uint8),
                                 None, None, None, None, 0, 0)
        assert rc == 0
        rw = torch.empty(m, dtype=torch.int8, device=dev); dn = torch.empty(m, dtype=torch.uint8, device=dev); ln = torch.empty((m, 2), dtype=torch.uint8, device=dev)
        b.step_dev(torch.from_numpy(player).to(dev), torch.from_numpy(keys).to(dev), torch.from_numpy(klen).to(dev), idx=torch.from_numpy(idx).to(dev),
                   reset_seeds=torch.from_numpy(rs.astype(np.int64)).to(dev).to(torch.int32), reward=rw, done=dn, lines=ln)
        torch.cuda.synchronize()
        assert np.array_equal(o_rw, rw.cpu().numpy()) and np.array_equal(o_dn, dn.cpu().numpy()) and np.array_equal(o_ln, ln.cpu().numpy()), t
    assert np.array_equal(a.export_states(), b.export_states())
    a.close(); b.close()


@pytest.mark.parametrize("H,W,pieces", [(22, 10, ALL7), (40, 20, ALL7), (12, 5, (4,))])
def test_game_codec_round_trip(H, W, pieces):
    """b2t_get_game -> b2t_set_game into another slot reproduces the packed state word for word (the raw form is complete:
    pickling and per-env copies through the host lose nothing), after play with garbage, combos, restarts."""
    from b200tetris import Engine
    n = 64
    e = Engine(2 * n, H, W, pieces, seed=5)
    gids = np.arange(n)
    e.reset(harness.seeds_for(gids, np.zeros(n, int)), gids)
    rng = np.random.default_rng(1)
    for t in range(90):
        player = ((gids + t) & 1).astype(np.uint8)
        K, keys, klen = e.enumerate(player, gids)
        sel = rng.integers(0, 1 << 30, n) % K
        e.step(player, keys[gids, sel], klen[gids, sel], gids, auto_reset=True, time_ms=int(rng.choice([400, 1000, 3100])))
        if t % 30 == 29:
            for g in range(0, n, 3):
                raw, field = e.get_game(g)
                e.set_game(n + g, raw, field)
            a, b = e.export_states(gids[::3]), e.export_states(n + gids[::3])
            assert np.array_equal(a, b), np.argwhere(a != b)[:5]
    e.close()
