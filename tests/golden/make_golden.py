"""Generates tests/golden/*.npz from the compiled, UNMODIFIED reference backend (oracle/_ref/B) driven through the
reference's own Python `action_list` / `state_dict` (imported from /root/reference).  Run in the build container:

    make -C oracle ref port && python tests/golden/make_golden.py

The C restatement leads each rollout only to pick actions (greedy policy needs a world model); every recorded value is
read from the reference side, and the rollout aborts on the first difference between the two."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness  # noqa: E402
import pyoracle  # noqa: E402
import refdrive  # noqa: E402

CASES = {
    # name: (kind, H, W, pieces, games, steps, solo_period)
    "c2_random_22x10": ("random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 16, 300, 0),
    "greedy_22x10": ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 600, 0),
    "greedy_solo_22x10": ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 4, 400, 29),
    "junk_22x10": ("junk", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 250, 13),
    "sventon_22x10": ("sventon", 22, 10, (0, 1, 2, 3, 4, 5, 6), 8, 150, 0),
    "c5_greedy_16x8_p06": ("greedy", 16, 8, (0, 6), 8, 300, 0),
    "c5_random_40x20": ("random", 40, 20, (0, 1, 2, 3, 4, 5, 6), 4, 250, 0),
    "c5_greedy_40x20": ("greedy", 40, 20, (0, 1, 2, 3, 4, 5, 6), 2, 200, 0),
    "zs_only_10x6": ("random", 10, 6, (2, 3), 8, 120, 0),
    # other player counts (8th entry): PythonHandle with 3 / 4 / 1 players
    "p3_greedy_22x10": ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 6, 450, 0, 3),
    "p4_random_22x10": ("random", 22, 10, (0, 1, 2, 3, 4, 5, 6), 12, 300, 0, 4),
    "p1_greedy_22x10": ("greedy", 22, 10, (0, 1, 2, 3, 4, 5, 6), 6, 300, 0, 1),
}


def make(name, kind, H, W, pieces, n, T, solo, P=2):
    kmax, lmax = (128, 64) if W > 10 else (64, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces, n_players=P)
    ref = refdrive.RefVector(n, H, W, pieces, n_players=P)
    assert refdrive.reference_python() is not None, "needs /root/reference for action_list/state_dict"
    rec = dict(player=[], keys=[], key_len=[], reward=[], done=[], lines=[], K=[], enum_digest=[], obs_digest=[])
    snaps = dict(t=[], field=[], vec=[])
    resets = {}

    def on_step(t, player, keys, klen, r, f, enum, other):
        rec["player"].append(player.copy()); rec["keys"].append(keys.copy()); rec["key_len"].append(klen.copy())
        rec["reward"].append(r[0]); rec["done"].append(r[1]); rec["lines"].append(r[2])
        rec["K"].append(enum[0].copy()); rec["enum_digest"].append(harness.digest(enum[1], enum[2]))
        s = other.scal().copy(); vec = f[1].copy()
        if t == 0:
            s[..., 3] = 0; vec[..., 3] = 0
        rec["obs_digest"].append(harness.digest(f[0], vec, f[2], s))
        if t % 50 == 49 or t == T - 1:
            snaps["t"].append(t); snaps["field"].append(f[0].copy()); snaps["vec"].append(vec)

    def on_reset(t, ids, seeds):
        resets[t] = (ids.copy(), seeds.copy())

    st = harness.rollout_pair(lead, ref, T, harness.Policy(kind, kmax=kmax, lmax=lmax, solo_period=solo, n_players=P),
                              on_step=on_step, on_reset=on_reset, compare_hidden=True)
    out = dict(n=n, T=T, H=H, W=W, n_players=P, pieces=np.array(pieces), kmax=kmax, lmax=lmax, policy=kind,
               seeds0=harness.seeds_for(np.arange(n), np.zeros(n, int)),
               reset_steps=np.array(sorted(resets), np.int64), snap_steps=np.array(snaps["t"], np.int64),
               snap_field=np.array(snaps["field"]), snap_vec=np.array(snaps["vec"]))
    for k, v in rec.items():
        out[k] = np.array(v)
    for t, (ids, seeds) in resets.items():
        out["reset_ids_%d" % t] = ids.astype(np.int32)
        out["reset_seeds_%d" % t] = seeds.astype(np.uint32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, st)


if __name__ == "__main__":
    for name, args in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        make(name, *args)
