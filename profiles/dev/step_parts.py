"""How long is the fused step without its observation tensors, and with only them?  (potential of overlapping the writer with the game logic)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "drl-tetris_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
from b200tetris import Engine
from bench import splitmix64, sventon_keys

n = 1 << 20
dev = torch.device("cuda:0")
gids = np.arange(n, dtype=np.int64)
e = Engine(n, 22, 10)
e.reset_dev(torch.from_numpy((splitmix64(0xB200, gids, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))
ring = []
for t in range(4):
    k, ln = sventon_keys(gids, t)
    ring.append((torch.from_numpy(k).to(dev), torch.from_numpy(ln).to(dev),
                 torch.from_numpy((splitmix64(0xB200, gids, t + 1) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32),
                 torch.from_numpy(((gids + t) & 1).astype(np.uint8)).to(dev)))
field = torch.empty((n, 2, 22, 10), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
pidx = torch.empty((n, 2), dtype=torch.uint8, device=dev)
reward = torch.empty(n, dtype=torch.int8, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev); lines = torch.empty((n, 2), dtype=torch.uint8, device=dev)

def timeit(fn, reps=20):
    for t in range(5): fn(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(reps): fn(5 + t)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def full(t):
    k, ln, sd, pl = ring[t % 4]
    e.step_dev(pl, k, ln, reset_seeds=sd, reward=reward, done=done, lines=lines, field=field, vec=vec, piece_idx=pidx)
def noobs(t):
    k, ln, sd, pl = ring[t % 4]
    e.step_dev(pl, k, ln, reset_seeds=sd, reward=reward, done=done, lines=lines)
def fieldonly(t):
    k, ln, sd, pl = ring[t % 4]
    e.step_dev(pl, k, ln, reset_seeds=sd, reward=reward, done=done, lines=lines, field=field)
def obs(t):
    e.observe_dev(field=field, vec=vec, piece_idx=pidx)
print("fused step, all outputs      %.4f ms" % timeit(full))
print("step without observations    %.4f ms" % timeit(noobs))
print("step + field only            %.4f ms" % timeit(fieldonly))
print("observe alone (generic)      %.4f ms" % timeit(obs))
