"""The C ABI is usable from plain C: examples/c_abi_minimal.c compiles as C11 against
include/gleipnir_b200.h and links against the shared library (CPU check); on a GPU it runs the
eggcarton configuration end to end without Python or torch in the process."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gleipnir_b200")


def _build(tmp_path):
    exe = os.path.join(str(tmp_path), "c_abi_minimal")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_minimal.c"), "-o", exe, "-L", LIBDIR, "-lgleipnir_b200",
           "-Wl,-rpath," + LIBDIR, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_compiles_and_links_as_c11(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libgleipnir_b200.so")):
        pytest.skip("library not built")
    _build(tmp_path)


@pytest.mark.gpu
def test_c_host_runs_the_eggcarton_configuration(cuda, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"ln Z = ([-0-9.]+) \+- ([0-9.]+)", r.stdout)
    assert m, r.stdout
    assert abs(float(m.group(1)) - 235.856) < 0.15            # exact ln Z = 235.85594; reported sigma 0.025
    assert "148000 dead rows" in r.stdout                      # 138163 deaths + (10000 - 163) survivors of the last batch
