"""The C-ABI shared library: loads, exports every symbol include/score_b200.h declares, and its
struct layouts agree between the C header (compiled with gcc) and the ctypes mirror.  No compute calls
(there is no GPU on the CPU test box)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "score_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(score_[a-z_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for n in ("score_create", "score_solve", "score_get_solution", "score_get_csr", "score_get_sizes",
              "score_round_so", "score_destroy", "score_last_error", "score_version"):
        assert n in names


def test_library_exports_every_declared_symbol(built_lib):
    from score_b200 import _lib

    lib = _lib.load()
    declared = _declared_functions()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/score_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (score_[a-z_0-9]+)", out))
    assert set(declared) <= exported
    assert lib.score_version().decode().startswith("score_b200")


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "--list-elf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header(tmp_path, built_lib):
    from score_b200 import _lib

    structs = ["ScoreProblemDesc", "ScoreParams", "ScoreInstanceStats", "ScoreStats", "ScoreRefineParams", "ScoreRefineStats",
               "ScoreRefineInstanceStats"]
    fields = {s: [f[0] for f in getattr(_lib, s)._fields_] for s in structs}
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s in structs:
        prog.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fields[s]:
            prog.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    prog.append("return 0;}")
    src = os.path.join(tmp_path, "layout.c")
    exe = os.path.join(tmp_path, "layout")
    open(src, "w").write("\n".join(prog))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", src, "-o", exe], check=True)  # the header is plain C
    got = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for s in structs:
        ct = getattr(_lib, s)
        assert int(got[s]) == C.sizeof(ct), s
        for f in fields[s]:
            assert int(got[f"{s}.{f}"]) == getattr(ct, f).offset, f"{s}.{f}"


def test_argument_validation_needs_no_device(built_lib):
    """Errors are codes + score_last_error(), never exceptions across the ABI."""
    from score_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    assert lib.score_create(None, 0, C.byref(h)) == _lib.SCORE_ERR_INVALID
    assert "null" in _lib.last_error()
    desc = _lib.ScoreProblemDesc()
    desc.dim = 4
    assert lib.score_create(C.byref(desc), 0, C.byref(h)) == _lib.SCORE_ERR_INVALID
    assert "not 2 or 3" in _lib.last_error()
    assert not h.value
    assert lib.score_solve(None, None, None, None) == _lib.SCORE_ERR_INVALID
    assert lib.score_get_solution(None, None, None, None, None) == _lib.SCORE_ERR_INVALID
    assert lib.score_round_so(5, 1, None, None, 0) == _lib.SCORE_ERR_INVALID
    lib.score_destroy(None)  # no-op


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from score_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(tmp_path, "libscore_b200.so"))
    with pytest.raises(_lib.ScoreLibraryError):
        _lib.load()
