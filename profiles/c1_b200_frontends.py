"""BASELINE configs[0] call pattern (get_state -> get_actions -> random placement -> perform_action -> reset on done) on the two
reference-shaped front-ends of b200tetris, on the GPU box: the batched vector env (one C-ABI call per method) at 32 and 4096
envs, and the per-env tetris_env facade (one game slot per PythonHandle, one small launch per call) at 32 envs.
Writes gpurun_out/c1_b200_frontends.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "drl-tetris_b200")]
from b200tetris import tetris_env, tetris_environment_vector  # noqa: E402

out = {}
rng = np.random.RandomState(0)
for n, steps in ((32, 300), (4096, 40)):
    env = tetris_environment_vector(n, settings={"seed": 1})
    cur = rng.choice([0, 1], n)
    parts = [0, 0, 0, 0]
    for warm in (True, False):
        t0 = time.perf_counter()
        for t in range(10 if warm else steps):
            cur = 1 - cur
            ta = time.perf_counter()
            env.get_state()
            tb = time.perf_counter()
            acts = env.get_actions(player=list(cur))
            tc = time.perf_counter()
            chosen = [a[rng.randint(len(a))] for a in acts]
            td = time.perf_counter()
            _, dones = env.perform_action(chosen, player=list(cur))
            te = time.perf_counter()
            parts = [p + q for p, q in zip(parts, (tb - ta, tc - tb, td - tc, te - td))] if not warm else [0, 0, 0, 0]
            ended = [i for i, d in enumerate(dones) if d]
            if ended:
                env.reset(env=ended)
        dt = time.perf_counter() - t0
    out["vector_env_%d" % n] = {"env_steps_per_sec": n * steps / dt, "ms_per_loop_iteration": 1e3 * dt / steps,
                                "ms_get_state/get_actions/choose/perform_action": [round(1e3 * p / steps, 3) for p in parts]}
# the per-env facade: raw PythonHandle objects, as environment/tetris_environment.py drives them
n, steps = 32, 60
hs = [tetris_env.PythonHandle(2, [22, 10]) for _ in range(n)]
cur = rng.choice([0, 1], n)
for warm in (True, False):
    t0 = time.perf_counter()
    for t in range(5 if warm else steps):
        cur = 1 - cur
        for i, h in enumerate(hs):
            p = int(cur[i])
            snap = h.copy()                      # get_state(): data_types.state keeps a copy of the backend
            h.get_actions(p)
            acts = h.masks[p].action or [[0]]
            a = [[0], [0]]
            a[p] = acts[rng.randint(len(acts))]
            h.make_action(a)
            if h.finish_action(400):
                h.reset()
            del snap
    dt = time.perf_counter() - t0
out["tetris_env_facade_32"] = {"env_steps_per_sec": n * steps / dt, "ms_per_env_step": 1e3 * dt / (n * steps)}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c1_b200_frontends.json"), "w"), indent=1)
