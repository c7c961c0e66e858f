"""BASELINE configs[0]: the reference's REAL Python front-end — tetris_environment_vector(32, tetris_environment,
default preset) on its compiled C++ backend (oracle/_ref/B) — rolled out with uniformly random placements, reset on done, side
alternating (drl_tetris/worker.py:84-104).  Needs /root/reference, so it runs in the build container (its CPU, not the GPU
box's); the reference arm of bench.py times the raw backend on the GPU box instead.  Writes profiles/r1_c1_reference_frontend.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "drl-tetris_b200")]
import refmod  # noqa: E402
from test_facade import _reference_front_end  # noqa: E402

te, tev = _reference_front_end()
mod, set_time = refmod.load("B")
te.tetris_env = mod
set_time(1234)
n, steps = 32, 2000
vec = tev.tetris_environment_vector(n, te.tetris_environment, settings={"presets": ["default"], "render": False})
rng = np.random.RandomState(0)
cur = rng.choice([0, 1], n)
t_actions = t_step = t_state = 0.0
placements = 0
t0 = time.perf_counter()
for t in range(steps):
    cur = 1 - cur
    a = time.perf_counter()
    states = vec.get_state()
    b = time.perf_counter()
    acts = [e.get_actions(s, player=int(p)) for e, s, p in zip(vec.envs, states, cur)]
    placements += sum(len(x) for x in acts)
    chosen = [x[rng.randint(len(x))] for x in acts]
    c = time.perf_counter()
    rewards, dones = vec.perform_action(chosen, player=[int(p) for p in cur])
    d = time.perf_counter()
    ended = [i for i, x in enumerate(dones) if x]
    if ended:
        set_time(1234 + t)
        vec.reset(env=ended)
    t_state += b - a; t_actions += c - b; t_step += d - c
dt = time.perf_counter() - t0
out = {"config": "C1: tetris_environment_vector(32 envs), default preset (2-player 22x10, all 7 pieces), random placements, %d steps" % steps,
       "where": "build container CPU, 1 process (the Python front-end holds the GIL)", "backend": "oracle/_ref/B (reference sources, -O2)",
       "env_steps_per_sec": n * steps / dt, "share_get_state": t_state / dt, "share_get_actions": t_actions / dt, "share_perform_action": t_step / dt,
       "mean_placements_per_list": placements / (n * steps)}
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "profiles", "r1_c1_reference_frontend.json"), "w"), indent=1)
