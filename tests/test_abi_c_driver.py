"""The C ABI driven by a plain C program (tests/abi_driver.c): the sequence an FFI binding performs, with no Python and
no C++ in between.  CPU: it must compile against include/ruhvro_b200.h alone and link every symbol it uses; GPU: it runs."""
import os
import subprocess

import pytest

from tests.golden import reference_datums as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pyruhvro_b200")


def _build(tmp_path):
    import pyruhvro_b200  # noqa: F401  (makes sure the library exists)
    exe = str(tmp_path / "abi_driver")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_driver.c"),
                           "-L", LIBDIR, "-lruhvro_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def test_c_driver_builds_against_the_header(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr       # no GPU work without arguments


@pytest.mark.gpu
def test_c_driver_runs(tmp_path):
    exe = _build(tmp_path)
    sj = tmp_path / "schema.json"
    sj.write_text(G.G345_SCHEMA)
    out = subprocess.run([exe, str(sj), G.G3_HEX, G.G4_HEX, G.G5_HEX], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_driver ok" in out.stdout and "batch 1: 2 rows" in out.stdout and "unexpected end of buffer (record 0)" in out.stdout
