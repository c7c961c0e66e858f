"""Shared differential-testing harness (test infrastructure).

Any two engines exposing the vector interface of oracle/pyoracle.OracleVector (`reset`, `step`, `observe`,
`scal`, `enumerate`) can be rolled out side by side under one of the deterministic policies below and diffed
after every call.  Policies read ONLY the first ("lead") engine, so the second is driven blind.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "drl-tetris_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

M64 = (1 << 64) - 1


def splitmix64(*words):
    """Counter hash used for seeds and action choices."""
    z = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        acc = np.zeros(np.broadcast(*[np.asarray(w) for w in words]).shape, np.uint64)
        for w in words:
            acc = acc + np.asarray(w).astype(np.uint64) + z
            acc = (acc ^ (acc >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            acc = (acc ^ (acc >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            acc = acc ^ (acc >> np.uint64(31))
    return acc


def seeds_for(game_ids, episode, salt=0xB200):
    return (splitmix64(salt, game_ids, episode) & np.uint64(0xFFFF)).astype(np.uint32)


# ----------------------------------------------------------------------------------------------- policies
def heuristic(field_after_lock):
    """Score boards [.., H, W] (acting player's field after the piece locked, lines not yet cleared)."""
    f = field_after_lock.astype(bool)
    H = f.shape[-2]
    full = f.all(-1)                                   # [.., H]
    lines = full.sum(-1)
    g = f & ~full[..., None]
    rows = np.arange(H)[:, None]
    top = np.where(g.any(-2), H - np.argmax(g, -2), 0)  # column heights (rows counted from the floor)
    filled_below = np.cumsum(g, -2) > 0
    holes = (filled_below & ~g & ~full[..., None]).sum((-1, -2))
    bump = np.abs(np.diff(top, axis=-1)).sum(-1)
    return 8.0 * lines - 5.0 * holes - 0.6 * top.sum(-1) - 0.4 * bump + 0.0 * rows.sum()


class Policy:
    """kind: 'random' (uniform over the de-duplicated placement list), 'greedy' (world-model heuristic with
    epsilon-random), 'junk' (random key strings incl. invalid keys and keys after a hard drop),
    'sventon' (the [8]*r+[2]+[3]*t+[7] strings of agents/sventon_agent/sventon_utils.py:9-13), 'press_key' (one key per step,
    the action_type="press_key" mode: pieces come down by gravity and lock by the lockdown timer)."""

    def __init__(self, kind, kmax=64, lmax=32, eps=0.05, salt=1, solo_period=0, n_players=2):
        self.kind, self.kmax, self.lmax, self.eps, self.salt, self.solo_period = kind, kmax, lmax, eps, salt, solo_period
        self.n_players = n_players

    def players(self, n, t, gids):
        if self.solo_period and (t // self.solo_period) % 2 == 1:
            return np.zeros(n, np.uint8)            # same side repeatedly: exercises lockdown
        return ((gids + t) % self.n_players).astype(np.uint8)

    def choose(self, lead, t, gids, player):
        n = len(gids)
        h = splitmix64(self.salt, gids, t)
        if self.kind == "junk":
            keys = np.zeros((n, self.lmax), np.uint8)
            ln = (h % np.uint64(min(self.lmax, 12) + 1)).astype(np.uint8)
            for j in range(self.lmax):
                keys[:, j] = (splitmix64(self.salt + 77, gids, t * 64 + j) % np.uint64(13)).astype(np.uint8)
            return keys, ln, None
        if self.kind == "press_key":   # action_type "press_key" (tetris_environment.py:84-85): ONE key 0..10 per perform_action
            keys = np.zeros((n, self.lmax), np.uint8)
            k = (h % np.uint64(16)).astype(np.uint8)
            # weighted towards moves and soft drops so pieces travel before gravity / lockdown / an occasional hard drop ends them
            keys[:, 0] = np.array([0, 1, 3, 5, 5, 8, 9, 10, 1, 3, 5, 2, 4, 6, 7, 5], np.uint8)[k]
            return keys, np.ones(n, np.uint8), None
        if self.kind == "sventon":
            W = lead.W
            r = (h % np.uint64(4)).astype(int)
            tr = ((h >> np.uint64(8)) % np.uint64(W)).astype(int)
            keys = np.zeros((n, self.lmax), np.uint8)
            ln = np.zeros(n, np.uint8)
            for i in range(n):
                s = [8] * r[i] + [2] + [3] * tr[i] + [7]
                keys[i, : len(s)] = s
                ln[i] = len(s)
            return keys, ln, None
        if self.kind == "random":
            K, keys, klen = lead.enumerate(player, bar_null_moves=True, kmax=self.kmax, lmax=self.lmax)[:3]
            sel = (h % K.astype(np.uint64)).astype(np.int64)
        else:  # greedy
            o = lead.expand(player, finalize=False, kmax=self.kmax, lmax=self.lmax)
            K, keys, klen = o["K"], o["keys"], o["key_len"]
            mine = o["field"][np.arange(n), :, player]                       # [n, kmax, H, W]
            score = heuristic(mine)
            score[np.arange(self.kmax)[None, :] >= K[:, None]] = -1e9
            sel = score.argmax(1)
            rnd = (h >> np.uint64(20)) % np.uint64(1000) < np.uint64(int(self.eps * 1000))
            sel = np.where(rnd, (h % K.astype(np.uint64)).astype(np.int64), sel)
        ar = np.arange(n)
        return keys[ar, sel], klen[ar, sel], (K, keys, klen)


def digest(*arrays):
    m = hashlib.blake2b(digest_size=8)
    for a in arrays:
        m.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(m.digest(), np.uint64)[0]


def assert_same(tag, t, a, b):
    if not np.array_equal(a, b):
        a, b = np.asarray(a), np.asarray(b)
        where = np.argwhere(a != b)[:6] if a.shape == b.shape else "shape %s vs %s" % (a.shape, b.shape)
        raise AssertionError("%s differs at step %d: first diffs at %s\nlead=%s\nother=%s" % (
            tag, t, where, a[tuple(where[0])] if not isinstance(where, str) else "", b[tuple(where[0])] if not isinstance(where, str) else ""))


def rollout_pair(lead, other, steps, policy, *, compare_enum=True, compare_scal=True, time_ms=400, episode0=0,
                 on_step=None, on_reset=None, skip_combo_time_first=True, compare_hidden=False):
    """Roll `lead` and `other` forward together; returns stats dict. Games auto-reset when done."""
    n = lead.n
    gids = np.arange(n)
    episode = np.full(n, episode0)
    seeds = seeds_for(gids, episode)
    lead.reset(seeds)
    other.reset(seeds)
    fresh = np.ones(n, bool)   # combo_time is uninitialised in the reference until the first finish_action
    stats = dict(steps=0, dones=0, lines=0, garbage=0, maxK=0, combos=0)
    for t in range(steps):
        player = policy.players(n, t, gids)
        keys, klen, enum = policy.choose(lead, t, gids, player)
        if compare_enum:
            if enum is None:
                enum = lead.enumerate(player, bar_null_moves=True, kmax=policy.kmax, lmax=policy.lmax)[:3]
            K2, keys2, klen2 = enum_other = other.enumerate(player, bar_null_moves=True, kmax=policy.kmax, lmax=policy.lmax)[:3]
            assert_same("K", t, enum[0], K2)
            assert_same("placement key_len", t, enum[2], klen2)
            assert_same("placement keys", t, enum[1], keys2)
            stats["maxK"] = max(stats["maxK"], int(enum[0].max()))
        ra = lead.step(player, keys, klen, time_ms=time_ms)
        rb = other.step(player, keys, klen, time_ms=time_ms)
        for name, a, b in zip(("reward", "done", "lines"), ra, rb):
            assert_same(name, t, a, b)
        fa, fb = lead.observe(), other.observe()
        for name, a, b in zip(("field", "vec", "piece_idx"), fa, fb):
            if name == "vec" and skip_combo_time_first and t == 0:
                a, b = a.copy(), b.copy()
                a[..., 3] = 0
                b[..., 3] = 0
            assert_same(name, t, a, b)
        if compare_scal:
            sa, sb = lead.scal().copy(), other.scal().copy()
            if skip_combo_time_first and t == 0:
                sa[..., 3] = 0
                sb[..., 3] = 0
            assert_same("scalars", t, sa, sb)
        if compare_hidden:
            assert_same("hidden", t, lead.hidden()[..., :13], other.hidden()[..., :13])
        if on_step is not None:
            on_step(t, player, keys, klen, rb, fb, enum_other if compare_enum else None, other)
        stats["steps"] += n
        stats["lines"] += int(ra[2].sum())
        stats["garbage"] += int((fa[1][..., 2] > 0).sum())
        stats["combos"] += int((fa[1][..., 4] > 0).sum())
        done = ra[1].astype(bool)
        if done.any():
            ids = np.nonzero(done)[0]
            episode[ids] += 1
            s2 = seeds_for(ids, episode[ids])
            lead.reset(s2, ids)
            other.reset(s2, ids)
            stats["dones"] += len(ids)
            if on_reset is not None:
                on_reset(t, ids, s2)
    return stats


# ----------------------------------------------------------------------------------------------- golden replay
def replay_golden(engine, g, check_enum=True):
    """Replay a tests/golden/*.npz trajectory (recorded from the compiled reference) on `engine` and compare every
    recorded output.  Returns number of game-steps checked."""
    n, T = int(g["n"]), int(g["T"])
    kmax, lmax = int(g["kmax"]), int(g["lmax"])
    engine.reset(g["seeds0"])
    rs = {int(t): (g["reset_ids_%d" % t], g["reset_seeds_%d" % t]) for t in g["reset_steps"]}
    for t in range(T):
        player = g["player"][t]
        if check_enum:
            K, keys, klen = engine.enumerate(player, bar_null_moves=True, kmax=kmax, lmax=lmax)[:3]
            assert_same("golden K", t, g["K"][t], K)
            assert digest(keys, klen) == g["enum_digest"][t], "placement lists differ from golden at step %d" % t
        r = engine.step(player, g["keys"][t], g["key_len"][t])
        assert_same("golden reward", t, g["reward"][t], r[0])
        assert_same("golden done", t, g["done"][t], r[1])
        assert_same("golden lines", t, g["lines"][t], r[2])
        f = engine.observe()
        s = engine.scal().copy()
        vec = f[1].copy()
        if t == 0:
            s[..., 3] = 0
            vec[..., 3] = 0
        assert digest(f[0], vec, f[2], s) == g["obs_digest"][t], "observation differs from golden at step %d" % t
        if t in g["snap_steps"]:
            i = list(g["snap_steps"]).index(t)
            assert_same("golden field", t, g["snap_field"][i], f[0])
            assert_same("golden vec", t, g["snap_vec"][i], vec)
        if t in rs:
            engine.reset(rs[t][1], rs[t][0])
    return n * T
