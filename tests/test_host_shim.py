"""The C++ host shim (include/pk_world.hpp) compiles against the C ABI with plain g++ and behaves
like physkit::world for the approach scenario of the reference's tests/co/co_tests.cpp:358-382."""
import os
import subprocess

import pytest

import physkit_b200 as pk
from physkit_b200 import build as pk_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_shim_test")
COMM_EXE = os.path.join(ROOT, "tests", "cpp", "comm_test")


def _compile(src="host_shim_test.cpp", exe=EXE, extra=()):
    pk_build.build()
    lib_dir = os.path.dirname(pk.library_path())
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), *extra,
           os.path.join(ROOT, "tests", "cpp", src), "-o", exe,
           "-L", lib_dir, "-lpk_collide", f"-Wl,-rpath,{lib_dir}", "-pthread"]
    subprocess.run(cmd, check=True)


def test_host_shim_compiles_and_refuses_without_gpu():
    import torch

    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3, r.stdout + r.stderr  # PK_E_NO_DEVICE surfaced as pk::error, no fallback


@pytest.mark.gpu
def test_host_shim_on_gpu():
    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host shim ok" in r.stdout


def _compile_comm():
    _compile("comm_test.cpp", COMM_EXE)


def test_comm_test_compiles_against_the_c_abi():
    """pk_comm_* (the contact / pose all-gathers inside the library) from a C++ host; without a GPU the program reports
    that and exits 3."""
    import torch

    _compile_comm()
    if not torch.cuda.is_available():
        r = subprocess.run([COMM_EXE], capture_output=True, text=True)
        assert r.returncode == 3, r.stdout + r.stderr


@pytest.mark.gpu
def test_comm_allgathers_on_gpu():
    """One rank per visible device (one on the test box: a communicator of one rank still runs both all-gathers)."""
    _compile_comm()
    r = subprocess.run([COMM_EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "comm ok" in r.stdout
