"""Builds drl-tetris_b200/b200tetris/libb200tetris.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles).

    python drl-tetris_b200/build.py [--force]

One translation unit per field size (csrc/b2t_geoms.h) compiled in parallel, plus the C-ABI unit, linked with the
static CUDA runtime so the library is self-contained (it shares the primary context with torch, streams and device
pointers interoperate).
"""
import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "b200tetris", "libb200tetris.so")
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "-Wno-deprecated-gpu-targets"]


def geometries():
    txt = open(os.path.join(CSRC, "b2t_geoms.h")).read()
    return [(int(h), int(w)) for h, w in re.findall(r"X\((\d+),\s*(\d+)\)", txt)]


def _newest_source():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)) if os.path.isdir(CSRC) else 0


def _compile(args):
    src, obj, defs = args
    cmd = [NVCC] + FLAGS + defs + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = obj[:-2] + ".ptxas.log"
    open(log, "w").write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return obj


def build(force=False, verbose=True, out=None, defines=(), only=None):
    """out/defines/only: build an experimental variant (python build.py --out libx.so -DB2T_GAME_MINB=5 --only 22x10)"""
    global OUT, OBJ
    if out:
        OUT = os.path.join(HERE, "b200tetris", out)
        OBJ = os.path.join(HERE, "build", os.path.splitext(out)[0])
        force = True
    os.makedirs(OBJ, exist_ok=True)
    newest = max(_newest_source(), os.path.getmtime(os.path.join(HERE, "..", "include", "b200tetris.h")), os.path.getmtime(__file__))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > newest:
        return OUT
    jobs = [(os.path.join(CSRC, "b2t_api.cu"), os.path.join(OBJ, "b2t_api.o"), list(defines))]
    for h, w in geometries():
        jobs.append((os.path.join(CSRC, "b2t_geom_inst.cu"), os.path.join(OBJ, "b2t_geom_%d_%d.o" % (h, w)),
                     ["-DB2T_H=%d" % h, "-DB2T_W=%d" % w] + list(defines)))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        objs = list(ex.map(_compile, jobs))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr[-4000:])
    if verbose:
        print("built", OUT)
    return OUT


if __name__ == "__main__":
    argv = sys.argv[1:]
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    build(force="--force" in argv, out=out, defines=[x for x in argv if x.startswith("-D")])
