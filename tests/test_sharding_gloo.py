"""N>1 path on the CPU: two gloo ranks each drive their contiguous shard of global game ids (the oracle stands in for the
GPU engine) and must reproduce, game by game, what one rank computes for the whole vector; the episode counters are summed
with the same all_reduce bench.py uses."""
import os
import sys

import numpy as np
import pytest

import harness

WORKER = r'''
import os, sys, json
import numpy as np
sys.path[:0] = [%(tests)r, %(oracle)r, %(pkg)r]
import torch, torch.distributed as dist
import harness, pyoracle
from b200tetris import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
N, T = 96, 40
gids = sharding.global_ids(N, rank, world)
v = pyoracle.OracleVector(len(gids))
v.reset(harness.seeds_for(gids, np.zeros(len(gids), int)))
pol = harness.Policy("sventon", lmax=32)
stats = dict.fromkeys(sharding.STAT_KEYS, 0)
dig = np.zeros((T, len(gids)), np.uint64)
for t in range(T):
    pl = pol.players(len(gids), t, gids)
    keys, ln, _ = pol.choose(v, t, gids, pl)
    r, d, lines = v.step(pl, keys, ln)
    f = v.observe()
    for j in range(len(gids)):
        dig[t, j] = harness.digest(r[j], d[j], lines[j], f[0][j], f[1][j] if t else 0, f[2][j])
    stats["env_steps"] += len(gids); stats["episodes"] += int(d.sum()); stats["lines"] += int(lines.sum())
    ids = np.nonzero(d)[0]
    if len(ids):
        v.reset(harness.seeds_for(gids[ids], np.full(len(ids), t + 1)), ids)
tot = sharding.reduce_stats(stats)
out = [None] * world
dist.all_gather_object(out, (int(gids[0]), dig.tolist()))
if rank == 0:
    full = np.concatenate([np.array(o[1], dtype=np.uint64) for o in sorted(out)], axis=1)
    print(json.dumps({"digest": str(harness.digest(full)), "stats": tot, "shape": list(full.shape)}))
dist.destroy_process_group()
'''


def _run(world, tmp_path):
    import subprocess
    script = tmp_path / ("worker_%d.py" % world)
    script.write_text(WORKER % dict(tests=os.path.join(harness.ROOT, "tests"), oracle=os.path.join(harness.ROOT, "oracle"),
                                    pkg=os.path.join(harness.ROOT, "drl-tetris_b200")))
    port = 29500 + (os.getpid() % 2000) + world
    out = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                                   "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], text=True, timeout=600,
                                  stderr=subprocess.DEVNULL)
    import json
    return json.loads([l for l in out.splitlines() if l.startswith("{")][-1])


def test_shard_ranges_partition_the_games():
    from b200tetris import sharding
    for n in (1, 7, 96, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            rs = [sharding.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            for g in (0, n // 3, n - 1):
                lo, hi = rs[sharding.owner_of(g, n, world)]
                assert lo <= g < hi


def test_two_gloo_ranks_reproduce_one_rank(tmp_path):
    one = _run(1, tmp_path)
    two = _run(2, tmp_path)
    assert one["shape"] == two["shape"] == [40, 96]
    assert one["digest"] == two["digest"]
    assert one["stats"] == two["stats"] and one["stats"]["env_steps"] == 96 * 40
