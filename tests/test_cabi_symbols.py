"""The C-ABI library loads and exports every symbol include/arpack_b200.h declares; without a GPU the entry points
fail loudly (info = -9990, ido = 99) instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "arpack_b200.h")
SO = os.path.join(ROOT, "arpack-ng_b200", "lib", "libarpack_b200.so")


def declared_functions():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"^(?:void\*?|int|long long|const char\*)\s+\*?(\w+)\s*\(", txt, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(SO)


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for must in ["dsaupd_c", "dseupd_c", "dnaupd_c", "dneupd_c", "ssaupd_c", "sseupd_c", "snaupd_c", "sneupd_c",
                 "pdsaupd_c", "pdseupd_c", "pdnaupd_c", "pdneupd_c", "pssaupd_c", "psseupd_c", "psnaupd_c", "psneupd_c",
                 "dsaupd_", "dseupd_", "dnaupd_", "dneupd_", "debug_c", "stat_c", "sstats_c", "sstatn_c",
                 "znaupd_c", "zneupd_c", "cnaupd_c", "cneupd_c", "znaupd_", "zneupd_"]:
        assert must in names, must
    assert len(names) >= 45


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_oracle_symbols_in_the_product(lib):
    """The product must not contain or load the CPU oracle."""
    assert not hasattr(lib, "ref_dsaupd") and not hasattr(lib, "ref_ctx_new")
    import subprocess
    out = subprocess.run(["ldd", SO], capture_output=True, text=True).stdout
    assert "libref_arpack" not in out


def test_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    n, nev, ncv = 100, 3, 10
    ido, info = C.c_int(0), C.c_int(0)
    resid, v, workd = np.zeros(n), np.zeros(n * ncv), np.zeros(3 * n)
    workl = np.zeros(ncv * ncv + 8 * ncv)
    iparam = np.zeros(11, dtype=np.int32)
    iparam[[0, 2, 6]] = [1, 10, 1]
    ipntr = np.zeros(11, dtype=np.int32)
    lib.dsaupd_c.argtypes = [C.POINTER(C.c_int), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_void_p,
                             C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.POINTER(C.c_int)]
    lib.dsaupd_c(C.byref(ido), b"I", n, b"LM", nev, 0.0, resid.ctypes.data, ncv, v.ctypes.data, n, iparam.ctypes.data,
                 ipntr.ctypes.data, workd.ctypes.data, workl.ctypes.data, len(workl), C.byref(info))
    assert ido.value == 99 and info.value == -9990
    lib.ab200_device_count.restype = C.c_int
    assert lib.ab200_device_count() == 0


def test_python_mirror_raises_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import arpack_ng_b200 as ab
    with pytest.raises(Exception):
        ab.solve(lambda x, y: None, 100, 3, 10, "LM", host_buffers=False)
