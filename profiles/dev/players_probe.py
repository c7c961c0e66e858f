"""Throughput of the compatibility kernel (n_players != 2): workload S on 2^18 games with 1, 3 and 4 players, next to the two-player kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "drl-tetris_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
from b200tetris import Engine
from bench import splitmix64, sventon_keys

n = 1 << 18
dev = torch.device("cuda:0")
gids = np.arange(n, dtype=np.int64)
for P in (2, 1, 3, 4):
    e = Engine(n, 22, 10, n_players=P)
    e.reset_dev(torch.from_numpy((splitmix64(0xB200, gids, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))
    ring = []
    for t in range(4):
        k, ln = sventon_keys(gids, t)
        ring.append((torch.from_numpy(k).to(dev), torch.from_numpy(ln).to(dev),
                     torch.from_numpy((splitmix64(0xB200, gids, t + 1) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32),
                     torch.from_numpy(((gids + t) % P).astype(np.uint8)).to(dev)))
    field = torch.empty((n, P, 22, 10), dtype=torch.uint8, device=dev); vec = torch.empty((n, P, 19), dtype=torch.uint8, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev)
    def one(t):
        k, ln, sd, pl = ring[t % 4]
        e.step_dev(pl, k, ln, reset_seeds=sd, done=done, field=field, vec=vec)
    for t in range(6): one(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(20): one(6 + t)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    K = torch.empty(n, dtype=torch.int32, device=dev)
    pl = ring[0][3]
    for _ in range(2): e.enumerate_dev(pl, K)
    e0.record()
    for _ in range(5): e.enumerate_dev(pl, K)
    e1.record(); torch.cuda.synchronize()
    print("n_players=%d: step %.3f ms for 2^18 games -> %.3e env-steps/s; enumerate %.3f ms; episodes %d" % (P, ms, n / ms * 1e3, e0.elapsed_time(e1) / 5, e.stats()["episodes"]), flush=True)
    e.close()
