"""The C-ABI library builds for sm_100a, loads, and exports every symbol that
include/gleipnir_b200.h declares (no compute calls: this runs without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from gleipnir_b200 import build
    return build.build()


def _declared():
    text = open(os.path.join(ROOT, "include", "gleipnir_b200.h")).read()
    return sorted(set(re.findall(r"GL_API[^;(]*?\b(gl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    names = _declared()
    assert len(names) >= 30
    for must in ("gl_ns_run_host", "gl_ns_advance", "gl_select_kworst", "gl_prior_draw", "gl_loglike_eval"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (gl_[a-z0-9_]+)", out))
    missing = [n for n in _declared() if n not in exported]
    assert not missing, missing


def test_ctypes_binding_covers_header(lib_path):
    from gleipnir_b200 import _lib
    lib = _lib.load()
    assert lib.gl_abi_version() == 1
    assert sorted(_lib.exported_symbols()) == _declared()


def test_library_is_sm100a_with_lineinfo(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gleipnir_b200.h"\nint main(void){gl_ns_config c; (void)c; return 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)
