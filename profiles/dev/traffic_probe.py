"""What bandwidth does the step kernel's read/write mix reach with no game logic?  (b2t_debug_traffic)"""
import os, sys
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "drl-tetris_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
from b200tetris import Engine
from bench import STEP_BYTES

n = 1 << 20
dev = torch.device("cuda:0")
e = Engine(n, 22, 10)
e.reset_dev(torch.zeros(n, dtype=torch.int32, device=dev))
keys = torch.zeros((n, 16), dtype=torch.uint8, device=dev); pl = torch.zeros(n, dtype=torch.uint8, device=dev)
kl = torch.zeros(n, dtype=torch.uint8, device=dev); seeds = torch.zeros(n, dtype=torch.int32, device=dev)
rew = torch.empty(n, dtype=torch.int8, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev); lines = torch.empty((n, 2), dtype=torch.uint8, device=dev)
field = torch.empty((n, 2, 22, 10), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev); pidx = torch.empty((n, 2), dtype=torch.uint8, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode in (1, 0):
    def go():
        assert e.lib.b2t_debug_traffic(e._h, n, P(keys), 16, P(pl), P(kl), P(seeds), P(rew), P(done), P(lines), P(field), P(vec), P(pidx), mode, st) == 0
    for _ in range(5): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print("step_like=%d: %.4f ms per launch -> %.0f GB/s of the step's algorithmic bytes (%d B/game)" % (mode, ms, STEP_BYTES * n / ms / 1e6, STEP_BYTES))
# plain copy of the same volume for reference
a = torch.empty(STEP_BYTES * n // 2, dtype=torch.uint8, device=dev); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): b.copy_(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("torch copy of %.0f MB (read+write = step volume): %.4f ms -> %.0f GB/s" % (a.numel() / 1e6, ms, 2 * a.numel() / ms / 1e6))
