"""The C ABI driven from plain C (tests/host/abi_smoke.c), the way the Julia shim's `ccall`s drive it:
no Python and no torch between the caller and libparticles_b200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "abi_smoke.c")
OUT = os.path.join(ROOT, "tests", "host", "_build", "abi_smoke")
LIBDIR = os.path.join(ROOT, "particles.jl_b200")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O1", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", OUT,
                           "-L", LIBDIR, "-lparticles_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"])
    return OUT


def test_abi_smoke_compiles_against_the_header():
    """CPU: the header is plain C (gcc -std=c11) and the program links against the shared object."""
    assert os.path.exists(build())


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [(0, 2000), (1, 512)])
def test_abi_smoke_runs(kind, n):
    exe = build()
    r = subprocess.run([exe, str(kind), str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-500:]
    assert "abi_smoke ok" in r.stdout
