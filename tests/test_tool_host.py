"""Host-side pieces of arpackmm_b200's ILU preconditioner (the dual-threshold factorisation and the level sets of the
triangular solves), compiled from the tool's own translation unit and run on the CPU.  The GPU side is covered by
test_gpu_arpackmm.py."""
import os
import shutil
import subprocess

import pytest

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(_NVCC), reason="needs nvcc to compile the tool's translation unit")
def test_ilut_and_level_sets_on_the_host():
    lib = os.path.join(_ROOT, "arpack-ng_b200", "lib")
    assert os.path.exists(os.path.join(lib, "libarpack_b200.so")), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    out = os.path.join(_ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "ilut_check")
    src = os.path.join(_ROOT, "tests", "toolhost", "ilut_check.cu")
    tool = os.path.join(_ROOT, "arpack-ng_b200", "csrc", "arpackmm_b200.cu")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(tool)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([_NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-ccbin", cxx,
                        "-Wno-deprecated-declarations", "-Xcompiler", "-Wno-deprecated-declarations", "-o", exe, src, "-L" + lib,
                        "-larpack_b200", "-ldl", "-Xlinker", "-rpath," + lib], check=True, timeout=900)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "ILUT OK" in p.stdout, p.stdout + p.stderr
