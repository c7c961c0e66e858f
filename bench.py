#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the b200tetris hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N ...            # the reference's own CPU backend on the box's host cores

Metric (BASELINE.json): env-steps/s (and placements/s) for 2-player 10x22 Tetris.
A "step" = one perform_action over every game of the batch: ONE launch of the fused step kernel (acting player's keys,
finish_action for both players, fused reset of finished games, state write-back, uint8 observation tensors out).
Workload at N=1: BASELINE configs[3] sharded weakly — 2^20 concurrent games per GPU (state 352 MB + 503 MB of observation per
step, both far larger than the 126 MB L2, so no flush is needed between iterations); actions are SVENton key strings
[8]*r+[2]+[3]*t+[7] (agents/sventon_agent/sventon_utils.py:9-13) drawn from a counter hash, resident in HBM before the
timed region.  The all-placements world model (configs[2], 65 536 boards) is timed in the same run and reported under
"placements" (also with bit-packed fields and on boards of a heuristic policy), the device-resident actor loop under "actor_loop",
BASELINE configs[4] (custom field sizes) under "custom_sizes".  roofline.mix_peak is the step's own HBM traffic replayed without
game logic (b2t_debug_traffic): what that read/write mix reaches on the device.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "drl-tetris_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

H, W = 22, 10
OBS_BYTES = 2 * (H * W + 19 + 1)                 # 480   
S_BYTES = 2 * (H * ((W + 7) // 8) + 128)         # 344
STEP_BYTES = 2 * S_BYTES + OBS_BYTES + 37        # 1205 algorithmic bytes per env-step
def board_bytes(mean_k):                         # 352 + 520*K per expanded board
    return S_BYTES + 8 + mean_k * (OBS_BYTES + 32 + 8)


def geom_bytes(h, w):
    """SURVEY.md §8d for any field size: (OBS, S, bytes per env-step, bytes per expanded board as a function of mean K)"""
    obs = 2 * (h * w + 19 + 1)
    st = 2 * (h * ((w + 7) // 8) + 128)
    return obs, st, 2 * st + obs + 37, (lambda k: st + 8 + k * (obs + 32 + 8))


def splitmix64(*words):
    z = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        acc = np.zeros(np.broadcast(*[np.asarray(w) for w in words]).shape, np.uint64)
        for w in words:
            acc = acc + np.asarray(w).astype(np.uint64) + z
            acc = (acc ^ (acc >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            acc = (acc ^ (acc >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            acc = acc ^ (acc >> np.uint64(31))
    return acc


def sventon_keys(gids, t, lmax=16, width=W):
    """[8]*r + [2] + [3]*tr + [7], r in [0,4), tr in [0,W) from a counter hash of (global game id, step)."""
    h = splitmix64(7, gids, t)
    r = (h % np.uint64(4)).astype(np.int64)
    tr = ((h >> np.uint64(8)) % np.uint64(width)).astype(np.int64)
    n = len(gids)
    col = np.arange(lmax)[None, :]
    keys = np.zeros((n, lmax), np.uint8)
    keys[col < r[:, None]] = 8
    keys[col == r[:, None]] = 2
    keys[(col > r[:, None]) & (col <= (r + tr)[:, None])] = 3
    keys[col == (r + tr + 1)[:, None]] = 7
    return keys, (r + tr + 2).astype(np.uint8)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every ~0.5 ms (nvidia-smi is too slow for a region that lasts a
    few milliseconds); summary() reports the samples that fall inside [t0, t1] (the timed region)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.err = gpu, [], False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.rows.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), int(reasons_fn(h))))
                time.sleep(0.0005)
        except Exception as e:  # NVML missing: report nothing rather than fail the bench
            self.err = repr(e)

    def summary(self, t0, t1):
        self.stop_flag = True
        self.join(timeout=2)
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        if not inside:   # region shorter than a sampling period: take the samples bracketing it
            before = [r for r in self.rows if r[0] < t0][-1:]
            after = [r for r in self.rows if r[0] > t1][:1]
            inside = before + after
        sm = [r[1] for r in inside]
        mask = 0
        for r in inside:
            mask |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "max_mhz", None),
                "reasons": sorted(k for k, v in self.BAD.items() if mask & v), "samples_in_timed_region": len([r for r in self.rows if t0 <= r[0] <= t1]),
                "source": "NVML" if self.err is None else "unavailable: " + self.err}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key):
    """per-launch dram bytes of the dominant kernel from the committed ncu summary, if any"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
    except Exception:
        return None


def workload_config(n):
    """`config` of the JSON line — the same object for both arms (--impl b200 and --impl reference), so the driver can see that the
    two lines measure one workload; what differs per arm (sample sizes, cores) is under cpu_baseline / e2e."""
    return {"workload": "C4 (BASELINE configs[3], weak): %d concurrent 2-player 22x10 games per GPU, all 7 pieces, workload S "
                        "(SVENton key strings from a counter hash), fused reset of finished games, uint8 observation tensors "
                        "for both players written every step" % n,
            "games_per_gpu": n, "l2": "working set per step (352 MB state + 503 MB observations) exceeds the 126 MB L2; no flush"}


# ---------------------------------------------------------------------------------------------- CPU legs
def cpu_port_leg(n=8192, steps=600):
    """The plain-C oracle (oracle/tetris_oracle.c) on ONE host core, same workload S on a bounded sample."""
    import pyoracle
    v = pyoracle.OracleVector(n, H, W)
    gids = np.arange(n)
    v.reset((splitmix64(0xB200, gids, 0) & np.uint64(0xFFFF)).astype(np.uint32))
    pre = [sventon_keys(gids, t) for t in range(64)]      # a ring of pre-generated key strings, as on the GPU side
    t0 = time.perf_counter()
    for t in range(steps):
        keys, ln = pre[t % 64]
        _, done, _ = v.step(((gids + t) & 1).astype(np.uint8), keys, ln)
        v.observe()
        d = np.nonzero(done)[0]
        if len(d):
            v.reset((splitmix64(0xB200, d, t + 1) & np.uint64(0xFFFF)).astype(np.uint32), d)
    dt = time.perf_counter() - t0
    return n * steps / dt, "%d games x %d steps of workload S incl. state_dict, 1 thread" % (n, steps)


def _ref_worker(args):
    """One process of the compiled reference backend (oracle/_ref/B): raw PythonHandle loop of workload S."""
    wid, n, steps = args
    import refmod
    mod, set_time = refmod.load("B")
    mod.set_pieces([0, 1, 2, 3, 4, 5, 6])
    gids = np.arange(n) + wid * n
    hs = []
    for g in gids:
        set_time(int(splitmix64(0xB200, g, 0)) & 0xFFFF)
        h = mod.PythonHandle(2, [H, W])
        h.reset()
        hs.append(h)
    pre = [sventon_keys(gids, t) for t in range(min(steps, 64))]     # a ring of pre-generated key lists
    pre = [[[int(k) for k in keys[i, :ln[i]]] for i in range(n)] for keys, ln in pre]
    t0 = time.perf_counter()
    for t in range(steps):
        acts = pre[t % len(pre)]
        for i, h in enumerate(hs):
            p = (int(gids[i]) + t) & 1
            a = [[0], [0]]
            a[p] = acts[i]
            h.make_action(a)
            if h.finish_action(400):
                set_time(int(splitmix64(0xB200, gids[i], t + 1)) & 0xFFFF)
                h.reset()
    return n * steps, time.perf_counter() - t0


def _ref_world_model_worker(args):
    """simulate_all_actions on the compiled reference, raw PythonHandle calls (tetris_environment.py:87-100 without its Python
    wrappers): per board get_actions, then per placement set(anchor) + make_action + finish_action + copy() (the state object)."""
    wid, n, plies = args
    import refmod
    mod, set_time = refmod.load("B")
    mod.set_pieces([0, 1, 2, 3, 4, 5, 6])
    rng = np.random.RandomState(wid)
    hs = []
    for g in range(n):
        set_time(int(splitmix64(0xC3, g + wid * n, 0)) & 0xFFFF)
        h = mod.PythonHandle(2, [H, W])
        h.reset()
        for t in range(plies):                     # a few plies of random placements so the boards are not empty
            p = (g + t) & 1
            h.get_actions(p)
            acts = h.masks[p].action
            a = [[0], [0]]
            a[p] = list(acts[rng.randint(len(acts))]) if len(acts) else [0]
            h.make_action(a)
            if h.finish_action(400):
                h.reset()
        hs.append(h)
    placements = 0
    t0 = time.perf_counter()
    for g, h in enumerate(hs):
        p = (g + plies) & 1
        h.get_actions(p)
        seen, acts = set(), []
        for k in h.masks[p].action:                # action_list de-duplication (data_types/action_list.py:4-20)
            k = tuple(k)
            if k not in seen:
                seen.add(k); acts.append(list(k))
        anchor = h.copy()
        for k in acts:
            h.set(anchor)
            a = [[0], [0]]
            a[p] = k
            h.make_action(a)
            h.finish_action(400)
            h.copy()                               # data_types.state keeps a deep copy of the backend (state.py:5)
        h.set(anchor)
        placements += len(acts)
    return placements, time.perf_counter() - t0


def reference_leg(steps, per_proc_games=256, per_proc_steps=800):
    """The reference's own C++ backend (unmodified, compiled -O2 into oracle/_ref/B) on every host core."""
    import multiprocessing as mp
    import refmod
    cores = os.cpu_count() or 1
    if not refmod.available("B"):
        v, sample = cpu_port_leg()
        return v, 1, "port", sample
    ctx = mp.get_context("spawn")
    vals = []
    global _REF_SAMPLE_MS
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, steps)):
            res = pool.map(_ref_worker, [(w, per_proc_games, per_proc_steps) for w in range(cores)])
            vals.append(sum(r[0] for r in res) / max(r[1] for r in res))
            _REF_SAMPLE_MS.append(1e3 * max(r[1] for r in res))
    global _REF_PLACEMENTS
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_world_model_worker, [(w, 24, 16) for w in range(cores)])
    _REF_PLACEMENTS = {"value": sum(r[0] for r in res) / max(r[1] for r in res), "unit": "placements/s", "cores": cores,
                       "sample": "%d procs x 24 boards after 16 random plies: get_actions + per placement set/make_action/finish_action/copy "
                                 "on raw PythonHandle objects" % cores}
    return float(np.median(vals)), cores, "reference", "%d procs x %d games x %d steps of workload S, raw PythonHandle make_action+finish_action loop (no Python front-end, no state copies)" % (cores, per_proc_games, per_proc_steps)


_REF_PLACEMENTS = None
_REF_SAMPLE_MS = []


def c5_leg(torch, Engine, dev, local, rank, h, w, pieces, n, nb, peak, steps):
    """BASELINE configs[4]: a custom field size / piece set at 2^18 games — the step kernel on workload S and the world model on
    nb boards, each against the algorithmic bytes of ITS geometry."""
    obs, st, step_b, board_b = geom_bytes(h, w)
    gids = np.arange(n, dtype=np.int64) + rank * n
    eng = Engine(n, h, w, pieces, device=local)
    lmax = 16 if w <= 10 else 32
    ring = []
    for t in range(4):
        k, ln = sventon_keys(gids, t, lmax=lmax, width=w)
        ring.append((torch.from_numpy(k).to(dev), torch.from_numpy(ln).to(dev),
                     torch.from_numpy((splitmix64(0xB200, gids, t + 1) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32),
                     torch.from_numpy(((gids + t) & 1).astype(np.uint8)).to(dev)))
    reward = torch.empty(n, dtype=torch.int8, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    field = torch.empty((n, 2, h, w), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
    pidx = torch.empty((n, 2), dtype=torch.uint8, device=dev)
    eng.reset_dev(torch.from_numpy((splitmix64(0xB200, gids, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))

    def one(t):
        k, ln, sd, pl = ring[t % 4]
        eng.step_dev(pl, k, ln, reset_seeds=sd, reward=reward, done=done, field=field, vec=vec, piece_idx=pidx)
    for t in range(4):
        one(t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(steps):
        one(4 + t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"game_size": [h, w], "pieces": list(pieces), "games_per_gpu": n, "env_steps_per_sec": n / (ms * 1e-3), "ms_per_step": ms,
           "algorithmic_bytes_per_env_step": step_b, "roofline_frac": step_b * n / (ms * 1e-3) / 1e9 / peak}
    # world model on the first nb games after a few plies of random placements
    K = torch.empty(nb, dtype=torch.int32, device=dev); offs = torch.empty(nb + 1, dtype=torch.int32, device=dev)
    idx = torch.arange(nb, device=dev, dtype=torch.int32)
    gen = torch.Generator(device=dev); gen.manual_seed(77 + rank)
    for t in range(12):
        pl = ring[t & 1][3][:nb].contiguous()
        eng.enumerate_dev(pl, K, idx=idx)
        sel = (torch.randint(0, 1 << 30, (nb,), device=dev, generator=gen) % K.to(torch.int64)).to(torch.int32)
        eng.step_dev(pl, idx=idx, placement=sel)
    side = ring[0][3][:nb].contiguous()
    eng.enumerate_dev(side, K, idx=idx)
    torch.cuda.synchronize()
    cap = int(K.sum().item()) + 1024
    pf = torch.empty((cap, 2, h, w), dtype=torch.uint8, device=dev); pv = torch.empty((cap, 2, 19), dtype=torch.uint8, device=dev)
    pp = torch.empty((cap, 2), dtype=torch.uint8, device=dev); pk = torch.empty((cap, eng.lmax), dtype=torch.uint8, device=dev)
    pl_ = torch.empty(cap, dtype=torch.uint8, device=dev); pr = torch.empty(cap, dtype=torch.int8, device=dev); pd = torch.empty(cap, dtype=torch.uint8, device=dev)

    def expand():
        eng.expand_dev(side, K, offs, cap, idx=idx, keys=pk, key_len=pl_, reward=pr, done=pd, field=pf, vec=pv, piece_idx=pp)
    for _ in range(2):
        expand()
    torch.cuda.synchronize()
    total = int(offs[-1].item())
    e0.record()
    for _ in range(5):
        expand()
    e1.record()
    torch.cuda.synchronize()
    xms = e0.elapsed_time(e1) / 5
    mk = total / nb
    out.update({"boards": nb, "mean_K": mk, "ms_per_expand": xms, "placements_per_sec": total / (xms * 1e-3),
                "algorithmic_bytes_per_board": board_b(mk), "placements_roofline_frac": board_b(mk) * nb / (xms * 1e-3) / 1e9 / peak})
    eng.close()
    return out


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--games", type=int, default=1 << 20, help="games per GPU")
    ap.add_argument("--boards", type=int, default=1 << 16, help="boards per GPU for the world-model leg")
    ap.add_argument("--no-extras", action="store_true", help="skip placements / e2e / cpu legs (profiling runs)")
    ap.add_argument("--leg", default="all", help="all | step | placements  (profiling runs: time only one leg)")
    ap.add_argument("--cuda-profile", default="", help="step | placements: bracket ONE extra untimed iteration of that leg with "
                    "cudaProfilerStart/Stop after its timed region (for `ncu --profile-from-start off`)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    a.warmup = max(a.warmup, 3)

    if a.impl == "reference":
        if rank != 0:
            return
        v, cores, kind, sample = reference_leg(a.steps)
        # a "step" of this arm = one bounded sample of the workload (every host core stepping its share of games); ms_per_step = its wall time
        print(json.dumps({"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": a.gpus,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.median(_REF_SAMPLE_MS)) if _REF_SAMPLE_MS else 0.0,
                          "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                          "config": workload_config(a.games),
                          "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample,
                                           "placements": _REF_PLACEMENTS},
                          "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from b200tetris import Engine, sharding
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.games
    gids = sharding.global_ids(n * world, rank, world)        # GLOBAL game ids: results do not depend on the sharding
    eng = Engine(n, H, W, device=local)
    R = 4                                                     # ring of pre-generated action / seed tensors
    keys_r, len_r, seed_r = [], [], []
    for t in range(R):
        k, ln = sventon_keys(gids, t)
        keys_r.append(torch.from_numpy(k).to(dev)); len_r.append(torch.from_numpy(ln).to(dev))
        seed_r.append(torch.from_numpy((splitmix64(0xB200, gids, t + 1) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))
    player_r = [torch.from_numpy(((gids + t) & 1).astype(np.uint8)).to(dev) for t in range(2)]
    reward = torch.empty(n, dtype=torch.int8, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    lines = torch.empty((n, 2), dtype=torch.uint8, device=dev)
    field = torch.empty((n, 2, H, W), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
    pidx = torch.empty((n, 2), dtype=torch.uint8, device=dev)
    eng.reset_dev(torch.from_numpy((splitmix64(0xB200, gids, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))

    def one_step(t):
        eng.step_dev(player_r[t & 1], keys_r[t % R], len_r[t % R], reset_seeds=seed_r[t % R], reward=reward, done=done, lines=lines,
                     field=field, vec=vec, piece_idx=pidx)

    sampler = ClockSampler(local)
    sampler.start()
    for t in range(a.warmup):
        one_step(t)
    torch.cuda.synchronize()
    t_wait = time.perf_counter()
    while not sampler.rows and sampler.err is None and time.perf_counter() - t_wait < 5.0:
        time.sleep(0.01)           # NVML initialisation takes a moment: make sure sampling runs before the timed region starts
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    torch.cuda.synchronize()
    time.sleep(0.01)
    wall0 = time.perf_counter()
    ev[0].record()
    for t in range(a.steps):
        one_step(a.warmup + t)
        ev[t + 1].record()
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    if a.cuda_profile == "step":
        torch.cuda.profiler.start(); one_step(a.warmup + a.steps); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    clocks = sampler.summary(wall0, wall1)
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = n * world * a.steps / (total_ms_max * 1e-3)
    peak, peak_src = measured_peak()
    k_ms = float(np.mean(per_launch_ms))
    achieved = STEP_BYTES * n / (k_ms * 1e-3) / 1e9
    stats = sharding.reduce_stats(eng.stats(), device=dev)    # the only collective: end-of-rollout episode counters

    def traffic_probe():
        """b2t_debug_traffic: one launch with the HBM traffic of a fused step (same tiles, same coalesced accesses, same resident
        warps) and no game logic — the bandwidth this read/write mix reaches on this device, next to the plain-copy peak."""
        import ctypes as C
        P = lambda t: C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        snap = eng.export_states(np.arange(0, n, max(1, n // 64), dtype=np.int32))   # (the probe scribbles on the games)
        eng.snapshot(3)
        def go():
            rc = eng.lib.b2t_debug_traffic(eng._h, n, P(keys_r[0]), keys_r[0].shape[1], P(player_r[0]), P(len_r[0]), P(seed_r[0]), P(reward), P(done),
                                           P(lines), P(field), P(vec), P(pidx), 1, st)
            assert rc == 0
        for _ in range(3):
            go()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record()
        for _ in range(10):
            go()
        a1.record()
        torch.cuda.synchronize()
        eng.restore(3)
        torch.cuda.synchronize()
        assert np.array_equal(snap, eng.export_states(np.arange(0, n, max(1, n // 64), dtype=np.int32)))
        return a0.elapsed_time(a1) / 10

    out = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
           "data": "synthetic",
           "config": workload_config(n), "episode_stats": {"episodes_ended": stats["episodes"], "lines_cleared": stats["lines"]},
           "clocks": clocks, "gpu_launches": a.steps,
           "roofline": {"bound": "hbm", "kernel": "game_kernel<22,10> (fused step)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "peak_source": peak_src, "algorithmic_bytes_per_env_step": STEP_BYTES,
                        "launch_ms": k_ms, "traffic": ncu_traffic("step_bytes_per_launch")}}
    if not a.no_extras and a.leg in ("all", "step") and not a.cuda_profile:
        p_ms = traffic_probe()
        out["roofline"]["mix_peak"] = {"achieved": STEP_BYTES * n / (p_ms * 1e-3) / 1e9, "unit": "GB/s", "launch_ms": p_ms,
                                       "frac_of_mix_peak": p_ms / k_ms,
                                       "note": "b2t_debug_traffic: the step's own reads and writes (tiles, coalesced 128-bit accesses, 16 resident "
                                               "warps per SM) with no game logic; what is left between it and the step is game logic that does "
                                               "not overlap the memory traffic"}

    if a.leg == "step":
        a.no_extras = True
    if not a.no_extras and a.leg == "all":
        # ---- end-to-end through the C ABI with HOST buffers (b2t_step_host): H2D of the step's inputs, D2H of reward/done and
        # of the observation tensors, inside the timed region
        e2e_steps = max(2, min(a.steps, 30))
        import ctypes as C
        from b200tetris.engine import _hp
        from b200tetris import host_array

        def pinned(t):   # the step's INPUTS: pinned, write-combined host memory from b2t_host_alloc (the host only writes them; the device's reads skip the CPU cache snoop)
            src = t.cpu().numpy() if hasattr(t, "cpu") else t
            buf = host_array(src.shape, src.dtype, write_combined=True)
            buf[...] = src
            return buf
        hk = [(pinned(keys_r[t % R]), pinned(len_r[t % R]), pinned(seed_r[t % R]).view(np.uint32), pinned(player_r[t & 1])) for t in range(R)]
        h_rew = torch.empty(n, dtype=torch.int8).pin_memory().numpy(); h_done = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
        h_lines = torch.empty((n, 2), dtype=torch.uint8).pin_memory().numpy()
        h_field = torch.empty((n, 2, H, W), dtype=torch.uint8).pin_memory().numpy(); h_vec = torch.empty((n, 2, 19), dtype=torch.uint8).pin_memory().numpy()
        h_pidx = torch.empty((n, 2), dtype=torch.uint8).pin_memory().numpy()

        def pack4(k):     # two 4-bit key tokens per byte, low nibble first (B2T_KEYS_PACKED4)
            return np.ascontiguousarray(k[:, 0::2] | (k[:, 1::2] << 4))
        hp4 = [pinned(pack4(keys_r[t % R].cpu().numpy())) for t in range(R)]

        def host_step(t, mode, want_obs=False):
            """perform_action through the C ABI with HOST buffers: b2t_step_host copies this step's inputs from pinned host memory to
            the device, launches the step and copies reward/done/lines (and, on request, the observations) back.
              lean:     keys as one code per byte in zero-terminated rows + acting player; rounds that end restart with engine-chosen
                        seeds (as the reference's restart from its clock)                                              17 B/game
              explicit: + key lengths and caller-chosen reset seeds                                                    22 B/game
              packed:   lean with 4-bit key tokens                                                                      9 B/game"""
            k, ln, sd, pl = hk[t % R]
            keys, lens, seeds, fmt, auto = k, None, None, 0, 1
            if mode == "explicit":
                lens, seeds, auto = ln, sd, 0
            elif mode == "packed":
                keys, fmt = hp4[t % R], 1
            rc = eng.lib.b2t_step_host(eng._h, None, n, _hp(pl, C.c_uint8), _hp(keys, C.c_uint8), _hp(lens, C.c_uint8), k.shape[1], None, 1, 0,
                                       _hp(seeds, C.c_uint32), _hp(h_rew, C.c_int8), _hp(h_done, C.c_uint8), _hp(h_lines, C.c_uint8),
                                       _hp(h_field, C.c_uint8) if want_obs else None, _hp(h_vec, C.c_uint8) if want_obs else None,
                                       _hp(h_pidx, C.c_uint8) if want_obs else None, None, fmt, auto)
            assert rc == 0

        times = []
        for mode, want_obs in (("lean", False), ("explicit", False), ("packed", False), ("lean", True)):
            for t in range(R if not want_obs else 1):     # untimed: every buffer of the ring has crossed PCIe once (the first device
                host_step(t, mode, want_obs)              # access to a pinned page pays for its mapping)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for t in range(e2e_steps):
                host_step(t + 1, mode, want_obs)
            times.append(time.perf_counter() - t0)
        e2e = torch.tensor(times, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
        lm = hk[0][0].shape[1]
        rate = lambda j: n * world * e2e_steps / float(e2e[j])
        out["e2e"] = {"value": rate(0), "unit": "env-steps/s", "h2d_bytes_per_step": n * (lm + 1), "d2h_bytes_per_step": n * 4, "steps": e2e_steps,
                      "api": "b2t_step_host: pinned (b2t_host_alloc, write-combined) host keys (one code per byte, zero-terminated rows of %d) + acting player -> device, fused "
                             "step with engine-chosen seeds for the rounds that restart, reward/done/lines -> pinned host (the result "
                             "perform_action returns, tetris_environment_vector.py:79-86); observation tensors stay in HBM for the consumer. "
                             "With every buffer pinned the kernel reads the keys and writes the results through the caller's pages (zero-copy), so the "
                             "PCIe traffic overlaps the compute inside one launch." % lm,
                      "variants": {"explicit_key_len_and_reset_seeds": {"value": rate(1), "h2d_bytes_per_step": n * (lm + 1 + 1 + 4)},
                                   "packed_4bit_keys": {"value": rate(2), "h2d_bytes_per_step": n * (lm // 2 + 1)}},
                      "with_observation_copy": {"value": rate(3), "d2h_bytes_per_step": n * (4 + OBS_BYTES),
                                                "note": "same call also copying the 480 B/game uint8 state_dict tensors to pinned host memory"}}

    if not a.no_extras:
        # ---- world model: BASELINE configs[2], 65 536 boards reached by 8..40 plies of random placements
        nb = a.boards
        w = Engine(nb, H, W, device=local)
        bg = np.arange(nb, dtype=np.int64) + rank * nb
        w.reset_dev(torch.from_numpy((splitmix64(0xC3, bg, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))
        K = torch.empty(nb, dtype=torch.int32, device=dev); offs = torch.empty(nb + 1, dtype=torch.int32, device=dev)
        plies = torch.from_numpy((8 + splitmix64(0xC3, bg, 1) % np.uint64(32)).astype(np.int64)).to(dev)
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        bd = torch.empty(nb, dtype=torch.uint8, device=dev)
        for t in range(40):
            pl = torch.from_numpy(((bg + t) & 1).astype(np.uint8)).to(dev)
            w.enumerate_dev(pl, K)
            sel = (torch.randint(0, 1 << 30, (nb,), device=dev, generator=gen) % K.to(torch.int64)).to(torch.int32)
            act = torch.nonzero(plies > t).flatten().to(torch.int32)
            if act.numel() == 0:
                break
            w.step_dev(pl[act.long()].contiguous(), idx=act, placement=sel[act.long()].contiguous(), done=bd[: act.numel()])
            # finished games simply restart
            dn = act[bd[: act.numel()].bool()]
            if dn.numel():
                w.reset_dev(torch.randint(0, 1 << 16, (dn.numel(),), device=dev, generator=gen).to(torch.int32), dn)
        side = torch.from_numpy(((bg + 40) & 1).astype(np.uint8)).to(dev)
        cap = nb * 40
        pf = torch.empty((cap, 2, H, W), dtype=torch.uint8, device=dev); pv = torch.empty((cap, 2, 19), dtype=torch.uint8, device=dev)
        pp = torch.empty((cap, 2), dtype=torch.uint8, device=dev); pk = torch.empty((cap, 32), dtype=torch.uint8, device=dev)
        pl_ = torch.empty(cap, dtype=torch.uint8, device=dev); pr = torch.empty(cap, dtype=torch.int8, device=dev)
        pd = torch.empty(cap, dtype=torch.uint8, device=dev); pb = torch.empty(cap, dtype=torch.int32, device=dev)

        def expand(fin=True):
            w.expand_dev(side, K, offs, cap, finalize=fin, board=pb, keys=pk, key_len=pl_, reward=pr, done=pd, field=pf, vec=pv, piece_idx=pp)

        for _ in range(3):
            expand()
        torch.cuda.synchronize()
        total = int(offs[-1].item())
        assert total <= cap, "placement capacity exceeded"
        reps = max(3, min(a.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            expand()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if a.cuda_profile == "placements":
            torch.cuda.profiler.start(); expand(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
        pt = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        ms = float(pt.item())
        e0.record()
        for _ in range(reps):
            expand(False)                      # Sherlock's mode (sherlock_utils.py:14): pieces locked, nothing finished
        e1.record()
        torch.cuda.synchronize()
        ms_nf = e0.elapsed_time(e1) / reps
        # extra format (not the graded figure): fields as bits, [rows, 2, ceil(H*W/32)] words instead of [rows, 2, H, W] bytes
        pfb = torch.empty((cap, 2, (H * W + 31) // 32), dtype=torch.int32, device=dev)

        def expand_bits():
            w.expand_dev(side, K, offs, cap, board=pb, keys=pk, key_len=pl_, reward=pr, done=pd, field=pfb, vec=pv, piece_idx=pp)
        for _ in range(2):
            expand_bits()
        e0.record()
        for _ in range(reps):
            expand_bits()
        e1.record()
        torch.cuda.synchronize()
        ms_bits = e0.elapsed_time(e1) / reps
        del pfb
        mean_k = total / nb
        ach = board_bytes(mean_k) * nb / (ms * 1e-3) / 1e9
        out["placements"] = {"metric": "placements_per_sec", "value": total * world / (ms * 1e-3), "unit": "placements/s",
                             "boards_per_sec": nb * world / (ms * 1e-3), "boards_per_gpu": nb, "mean_K": mean_k, "ms_per_expand": ms,
                             "finalize": True, "gpu_launches_per_expand": 6,
                             "finalize_false": {"value": total * world / (ms_nf * 1e-3), "ms_per_expand": ms_nf},
                             "bit_packed_fields": {"value": total * world / (ms_bits * 1e-3), "ms_per_expand": ms_bits,
                                                   "note": "extra format b2t_obs.dtype = B2T_BITS (one bit per cell); single-GPU timing"},
                             "workload": "C3 (BASELINE configs[2]): %d boards after 8..40 plies of random placements; enumerate + scan + "
                                         "expand(finalize=1) with uint8 tensors, keys, reward, done for every placement" % nb,
                             "roofline": {"bound": "hbm", "kernel": "enumerate_kernel + game_kernel<22,10> (expand)", "achieved": ach, "peak": peak,
                                          "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_board": board_bytes(mean_k),
                                          "traffic": ncu_traffic("expand_bytes_per_launch")}}
        w.close()
        if a.leg == "all":
            # ---- actor loop: a ply = the world model of every game + the step by the chosen placement, one CUDA-graph launch
            from b200tetris import ActorLoop
            al = Engine(nb, H, W, device=local, seed=7 + rank)
            al.reset_dev(torch.from_numpy((splitmix64(0xA1, bg, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))
            noise = torch.rand(nb * 40, device=dev)

            def score_fn(f, v, p, scores):      # the value network is the consumer's: a fixed pseudo-random score per row stands in
                scores.copy_(noise)
            loop = ActorLoop(al, score_fn)
            for _ in range(24):                  # warm-up plies (capture included): boards fill up, games end and restart
                loop.step()
            torch.cuda.synchronize()
            assert not loop.overflowed()
            plies = max(4, min(a.steps, 20))
            rows0 = 0
            e0.record()
            for _ in range(plies):
                loop.step()
            e1.record()
            torch.cuda.synchronize()
            ams = e0.elapsed_time(e1) / plies
            at = torch.tensor([ams], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(at, op=dist.ReduceOp.MAX)
            ams = float(at.item())
            rows = int(loop.offsets[-1].item())
            out["actor_loop"] = {"env_steps_per_sec": nb * world / (ams * 1e-3), "placements_per_sec": rows * world / (ams * 1e-3),
                                 "ms_per_ply": ams, "games_per_gpu": nb, "gpu_launches_per_ply": 7, "cuda_graph": True,
                                 "note": "every ply: enumerate + expand of all placements (uint8 tensors, (me, opponent) order) -> scores "
                                         "(stand-in: a fixed random tensor; the network is the consumer's) -> select -> step by placement "
                                         "with engine-chosen reset seeds; no host data, one graph launch (b200tetris.ActorLoop). "
                                         "SURVEY.md 8d: step + uint8 world model is 14.1 KB per env-step, an HBM ceiling of 4.6e8/s per GPU."}
            al.close()
            # ---- the world model on boards of sensible play: the C3 boards above come from RANDOM placements (holes and overhangs
            # everywhere, 9 under-overhang candidates per board for the enumerator to search); a learning agent's boards are cleaner.
            # 40 untimed plies of a hand-made heuristic policy (low, flat, hole-free stacks; torch ops standing in for the network),
            # then the same expand call timed on those boards.
            sp = Engine(nb, H, W, device=local, seed=11 + rank)
            sp.reset_dev(torch.from_numpy((splitmix64(0x5E1F, bg, 0) & np.uint64(0xFFFF)).astype(np.int64)).to(dev).to(torch.int32))

            def heuristic(f, v, p, scores):
                occ = f[:, 0].to(torch.float16)
                filled_above = (occ.cumsum(1) > 0).to(torch.float16)
                heights = filled_above.sum(1)
                holes = (filled_above - occ).sum((1, 2))
                scores.copy_((-0.5 * heights.sum(1) - 0.4 * (heights[:, 1:] - heights[:, :-1]).abs().sum(1) - 5.0 * holes).float())
            lp = ActorLoop(sp, heuristic, use_graph=False)
            for _ in range(40):
                lp.step()
            torch.cuda.synchronize()
            del lp
            side2 = torch.from_numpy(((bg + 40) & 1).astype(np.uint8)).to(dev)

            def expand2():
                sp.expand_dev(side2, K, offs, cap, board=pb, keys=pk, key_len=pl_, reward=pr, done=pd, field=pf, vec=pv, piece_idx=pp)
            for _ in range(3):
                expand2()
            torch.cuda.synchronize()
            total2 = int(offs[-1].item())
            e0.record()
            for _ in range(reps):
                expand2()
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / reps
            st2 = torch.tensor([ms2], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(st2, op=dist.ReduceOp.MAX)
            ms2 = float(st2.item())
            mk2 = total2 / nb
            out["placements"]["selfplay_boards"] = {
                "value": total2 * world / (ms2 * 1e-3), "unit": "placements/s", "boards_per_sec": nb * world / (ms2 * 1e-3), "mean_K": mk2,
                "ms_per_expand": ms2, "roofline_frac": board_bytes(mk2) * nb / (ms2 * 1e-3) / 1e9 / peak,
                "workload": "the same call on %d boards after 40 plies of a heuristic policy (low, flat, hole-free stacks) instead of "
                            "random placements" % nb}
            sp.close()
        if a.leg == "all" and world == 1:
            # ---- BASELINE configs[4]: custom field sizes and piece sets at 2^18 games
            out["custom_sizes"] = [c5_leg(torch, Engine, dev, local, rank, 16, 8, (0, 6), 1 << 18, 1 << 15, peak, min(a.steps, 10)),
                                   c5_leg(torch, Engine, dev, local, rank, 40, 20, (0, 1, 2, 3, 4, 5, 6), 1 << 18, 1 << 14, peak, min(a.steps, 10))]
        if rank == 0 and a.leg == "all":
            v, sample = cpu_port_leg()
            out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
