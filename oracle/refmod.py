"""TEST INFRASTRUCTURE ONLY.  Loader for the compiled, UNMODIFIED reference backend
(oracle/_ref/{A,B}/tetris_env*.so, built by oracle/Makefile from /root/reference).

The reference keeps field size and piece map in process globals (gameField.cpp:7-8,
gamePlay.cpp:9-10), so one process can drive exactly one (H, W, pieces) configuration at a
time; `load()` returns the module plus a `set_time(seed)` hook into oracle/ref_time_shim.c that
fixes what PythonHandle::seed() (PythonHandle.cpp:68-71) reads from time(NULL).
"""
import ctypes
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = {}


def available(variant="B"):
    return bool(glob.glob(os.path.join(_HERE, "_ref", variant, "tetris_env*.so")))


def load(variant="B"):
    """Return (module, set_time) for build variant 'A' (as shipped, -O0) or 'B' (patched -O2)."""
    if variant in _loaded:
        return _loaded[variant]
    hits = glob.glob(os.path.join(_HERE, "_ref", variant, "tetris_env*.so"))
    if not hits:
        raise FileNotFoundError("oracle/_ref/%s not built: run `make -C oracle ref`" % variant)
    path = hits[0]
    # distinct module names so A and B can coexist in one process
    spec = importlib.util.spec_from_file_location("tetris_env", path)
    if variant != "B" and "B" in _loaded or variant != "A" and "A" in _loaded:
        raise RuntimeError("pybind11 registers its types once per process: load only one variant")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(path)
    lib.oracle_set_time.argtypes = [ctypes.c_longlong]
    lib.oracle_set_time.restype = None

    def set_time(t):
        lib.oracle_set_time(int(t))

    _loaded[variant] = (mod, set_time)
    return _loaded[variant]
