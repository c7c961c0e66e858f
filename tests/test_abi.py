"""CPU-side checks of the drop-in boundary: the in-tree shared library loads, exports exactly what include/b200tetris.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import harness

HEADER = os.path.join(harness.ROOT, "include", "b200tetris.h")


def declared_functions():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b2t_[a-z_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from b200tetris import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libb200tetris.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"


def test_struct_sizes_match_header_layout(tmp_path):
    """The header is valid plain C, and the ctypes mirrors have the sizes the C compiler gives."""
    import subprocess
    from b200tetris import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "b200tetris.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(b2t_config),sizeof(b2t_info),sizeof(b2t_obs),sizeof(b2t_step_args),sizeof(b2t_expand_args),'
                   'sizeof(b2t_player_raw),sizeof(b2t_game_raw));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes == [C.sizeof(_lib.Config), C.sizeof(_lib.Info), C.sizeof(_lib.Obs), C.sizeof(_lib.StepArgs), C.sizeof(_lib.ExpandArgs),
                     C.sizeof(_lib.PlayerRaw), C.sizeof(_lib.GameRaw)]


def test_no_cpu_fallback():
    import torch
    from b200tetris import B2TError, Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B2TError) as e:
        Engine(4)
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_sources_do_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under drl-tetris_b200/ may import, include or link it."""
    bad = []
    for root, _, files in os.walk(os.path.join(harness.ROOT, "drl-tetris_b200")):
        if os.path.basename(root) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"pyoracle|tetris_oracle|refmod|oracle/_ref|libhostsim|pyhostsim", txt):
                    bad.append(os.path.join(root, f))
    assert not bad, bad
