"""Link-level drop-in check: a plain C program written against the ICB prototypes (tests/c_caller/icb_caller.c: host
arrays, CPU operator, bind(c) names + gfortran-ABI names + stat_c/debug_c) is compiled with gcc, linked against
libarpack_b200.so and run.  On a GPU it must reproduce the reference's answers; without one the library must fail
loudly (info = -9990 -> exit code 3), never fall back to a CPU path."""
import os
import subprocess

import pytest

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_EXE = os.path.join(_ROOT, "tests", "_build", "icb_caller")


def build_caller():
    src = os.path.join(_ROOT, "tests", "c_caller", "icb_caller.c")
    libdir = os.path.join(_ROOT, "arpack-ng_b200", "lib")
    os.makedirs(os.path.dirname(_EXE), exist_ok=True)
    if os.path.exists(_EXE) and os.path.getmtime(_EXE) >= max(os.path.getmtime(src),
                                                              os.path.getmtime(os.path.join(libdir, "libarpack_b200.so"))):
        return
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O2", "-std=c11", "-Wall", "-I", os.path.join(_ROOT, "include"), src, "-o", _EXE,
                           "-L", libdir, "-larpack_b200", "-lm", "-Wl,-rpath," + libdir])


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_c_caller_links_and_fails_loudly_without_a_device():
    build_caller()
    p = subprocess.run([_EXE], capture_output=True, text=True, timeout=600)
    if _has_gpu():
        assert p.returncode == 0, p.stdout + p.stderr
    else:
        assert p.returncode == 3, p.stdout + p.stderr
        assert "no usable CUDA device" in p.stderr and "no CPU path" in p.stderr


@pytest.mark.gpu
def test_c_caller_reproduces_reference_answers_on_gpu():
    build_caller()
    p = subprocess.run([_EXE], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    for part in ("dsaupd_c/dseupd_c OK", "dsaupd_/dseupd_ (Fortran ABI) OK", "dnaupd_c/dneupd_c OK",
                 "argument errors OK", "znaupd_c/zneupd_c OK", "all checks passed"):
        assert part in p.stdout
