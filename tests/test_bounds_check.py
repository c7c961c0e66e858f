"""The game core compiled for the CPU with B2T_BOUNDS_CHECK: every board access of fits / collide_mask (the routines every rule,
the key interpreter, the wall kicks and the placement search go through) is checked against the padded column array — an access
outside it aborts the process.  compute-sanitizer is not available on the GPU pool; this is the bounds evidence for the index
arithmetic the kernels share with the host build (walls of 4 / 5 columns, 3 ceiling and 4 floor rows, see b2t_core.cuh Geo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "drl-tetris_b200", "csrc")
LIB = os.path.join(HERE, "hostsim", "libhostsim_bounds.so")

SCRIPT = r'''
import sys, numpy as np
sys.path[:0] = [%(tests)r, %(hostsim)r, %(oracle)r, %(pkg)r]
import pyhostsim
pyhostsim._LIB = %(lib)r
pyhostsim.build = lambda force=False: None
import harness, pyoracle
cases = [("junk", 22, 10, (0,1,2,3,4,5,6), 2, 24, 200), ("greedy", 22, 10, (0,1,2,3,4,5,6), 2, 6, 300), ("press_key", 22, 10, (0,1,2,3,4,5,6), 2, 8, 300),
         ("junk", 40, 20, (0,1,2,3,4,5,6), 2, 4, 120), ("random", 25, 12, (0,1,2,3,4,5,6), 2, 8, 120), ("junk", 12, 5, (4,), 2, 8, 150),
         ("random", 10, 6, (2, 3), 2, 8, 120), ("junk", 16, 8, (0, 6), 3, 8, 150), ("greedy", 22, 10, (0,1,2,3,4,5,6), 4, 4, 200)]
for kind, H, W, pieces, P, n, steps in cases:
    kmax, lmax = (256, 64) if W > 10 else (96, 32)
    lead = pyoracle.OracleVector(n, H, W, pieces, n_players=P)
    sim = pyhostsim.HostSimVector(n, H, W, pieces, n_players=P)
    harness.rollout_pair(lead, sim, steps, harness.Policy(kind, kmax=kmax, lmax=lmax, n_players=P))
    pl = (np.arange(n) %% P).astype(np.uint8)
    sim.expand(pl, kmax=kmax, lmax=lmax)
print("bounds ok")
'''


def test_core_stays_inside_its_boards():
    srcs = [os.path.join(HERE, "hostsim", "hostsim.cpp")] + [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-DB2T_BOUNDS_CHECK", "-Wno-unknown-pragmas",
                               "-I" + SRC, "-I" + os.path.join(HERE, "..", "include"), "-x", "c++", os.path.join(HERE, "hostsim", "hostsim.cpp"), "-o", LIB])
    code = SCRIPT % dict(tests=HERE, hostsim=os.path.join(HERE, "hostsim"), oracle=os.path.join(HERE, "..", "oracle"),
                         pkg=os.path.join(HERE, "..", "drl-tetris_b200"), lib=LIB)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "bounds ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
