"""The header-only C++ wrappers (wave-fenics_b200/hpp/wavefx.hpp) build against the C ABI with
the host compiler alone; on a GPU box the self-checking program also runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_wavefx_hpp.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_wavefx_hpp")


def _build(wfx):
    libdir = os.path.dirname(wfx.capi.LIB_PATH)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "wave-fenics_b200", "hpp"), SRC, "-o", EXE,
           "-L", libdir, "-l:libwavefx.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_wrappers_compile_and_link(wfx):
    _build(wfx)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_wrappers_run(wfx):
    _build(wfx)
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "hpp ok" in r.stdout, r.stdout + r.stderr


def test_dolfinx_adapter_against_mock(wfx):
    """the DOLFINx adapter header, compiled and run against stand-in classes
    (DOLFINx is not available here).  CPU only."""
    libdir = os.path.dirname(wfx.capi.LIB_PATH)
    src = os.path.join(ROOT, "tests", "cpp", "test_dolfinx_adapter.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_dolfinx_adapter")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "wave-fenics_b200", "hpp"), "-I", os.path.join(ROOT, "tests", "cpp"),
           src, "-o", exe, "-L", libdir, "-l:libwavefx.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "adapter ok" in r.stdout, r.stdout + r.stderr
