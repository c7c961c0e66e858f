"""A/B of the two enumerate forms on the bench's world-model boards: identical lists, and CUDA-event timings of enumerate alone and of
the whole expand.  B2T_ENUM_SPLIT_MIN is read per call: 1<<30 = the one-kernel form, 0 = scan/search/collect."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "drl-tetris_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
from b200tetris import Engine
from bench import splitmix64

def boards(nb, H=22, W=10, pieces=(0,1,2,3,4,5,6), plies_max=40):
    dev = torch.device("cuda:0")
    w = Engine(nb, H, W, pieces, device=0)
    bg = np.arange(nb, dtype=np.int64)
    w.reset_dev(torch.from_numpy((splitmix64(0xC3, bg, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))
    K = torch.empty(nb, dtype=torch.int32, device=dev)
    plies = torch.from_numpy((8 + splitmix64(0xC3, bg, 1) % np.uint64(32)).astype(np.int64)).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    bd = torch.empty(nb, dtype=torch.uint8, device=dev)
    for t in range(plies_max):
        pl = torch.from_numpy(((bg + t) & 1).astype(np.uint8)).to(dev)
        w.enumerate_dev(pl, K)
        sel = (torch.randint(0, 1 << 30, (nb,), device=dev, generator=gen) % K.to(torch.int64)).to(torch.int32)
        act = torch.nonzero(plies > t).flatten().to(torch.int32)
        if act.numel() == 0:
            break
        w.step_dev(pl[act.long()].contiguous(), idx=act, placement=sel[act.long()].contiguous(), done=bd[: act.numel()])
        dn = act[bd[: act.numel()].bool()]
        if dn.numel():
            w.reset_dev(torch.randint(0, 1 << 16, (dn.numel(),), device=dev, generator=gen).to(torch.int32), dn)
    side = torch.from_numpy(((bg + 40) & 1).astype(np.uint8)).to(dev)
    return w, side

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    dev = torch.device("cuda:0")
    if len(sys.argv) > 1 and sys.argv[1] == "profile":      # one split enumerate of the 65 536-board case between profiler start/stop
        os.environ["B2T_ENUM_SPLIT_MIN"] = sys.argv[2] if len(sys.argv) > 2 else "0"
        nb = 1 << 16
        w, side = boards(nb)
        K = torch.empty(nb, dtype=torch.int32, device=dev)
        for _ in range(3): w.enumerate_dev(side, K)
        torch.cuda.synchronize()
        torch.cuda.profiler.start(); w.enumerate_dev(side, K); torch.cuda.synchronize(); torch.cuda.profiler.stop()
        return
    for (H, W, pieces, nb) in ((22, 10, (0,1,2,3,4,5,6), 1 << 16), (16, 8, (0, 6), 1 << 15), (40, 20, (0,1,2,3,4,5,6), 1 << 14), (22, 10, (0,1,2,3,4,5,6), 3000)):
        os.environ["B2T_ENUM_SPLIT_MIN"] = str(1 << 30)
        w, side = boards(nb, H, W, pieces)
        K = torch.empty(nb, dtype=torch.int32, device=dev)
        kmax, lmax = w.kmax, w.lmax
        sub = min(nb, 4096)
        idx = torch.arange(sub, device=dev, dtype=torch.int32)
        out = {}
        for name, v in (("one", 1 << 30), ("split", 0)):
            os.environ["B2T_ENUM_SPLIT_MIN"] = str(v)
            keys = torch.zeros((sub, kmax, lmax), dtype=torch.uint8, device=dev); kl = torch.zeros((sub, kmax), dtype=torch.uint8, device=dev)
            Ks = torch.empty(sub, dtype=torch.int32, device=dev)
            w.enumerate_dev(side[:sub].contiguous(), Ks, idx=idx, keys=keys, key_len=kl)
            w.enumerate_dev(side, K)
            torch.cuda.synchronize()
            out[name] = (K.cpu().numpy().copy(), Ks.cpu().numpy(), keys.cpu().numpy(), kl.cpu().numpy())
            ms = timeit(lambda: w.enumerate_dev(side, K))
            print("%dx%d n=%d %-5s enumerate %.1f us  meanK %.2f  stats %s" % (H, W, nb, name, ms * 1e3, float(K.float().mean()), w.stats()["capacity_overflow"]), flush=True)
        for a, b in zip(out["one"], out["split"]):
            assert np.array_equal(a, b), "lists differ"
        print("   lists identical", flush=True)
        w.close()

if __name__ == "__main__":
    main()
