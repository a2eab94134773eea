"""Tool-layer formats: the Matrix-Market reader semantics of
EXAMPLES/MATRIX_MARKET/arpackSolver.hpp:361-416 and the --restart dump files (:664-704), through the C-ABI.
Host-only code: runs without a GPU.  The GPU part (the arpackmm_b200 driver itself) is in test_gpu_arpackmm.py."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    lib = C.CDLL(os.path.join(_ROOT, "arpack-ng_b200", "lib", "libarpack_b200.so"))
    lib.ab200_mm_read_csr.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                                      C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
                                      C.POINTER(C.POINTER(C.c_double))]
    lib.ab200_mm_free.argtypes = [C.c_void_p]
    lib.ab200_restart_save_f64.argtypes = [C.c_char_p, C.c_longlong, C.c_void_p]
    lib.ab200_restart_load_f64.argtypes = [C.c_char_p, C.c_longlong, C.c_void_p, C.c_int]
    return lib


def read(L, path):
    n, m, nnz = C.c_int(), C.c_int(), C.c_longlong()
    rp, co, va = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    rc = L.ab200_mm_read_csr(str(path).encode(), C.byref(n), C.byref(m), C.byref(nnz), C.byref(rp), C.byref(co),
                             C.byref(va))
    if rc != 0:
        return rc, None
    A = sp.csr_matrix((np.ctypeslib.as_array(va, (nnz.value,)).copy(), np.ctypeslib.as_array(co, (nnz.value,)).copy(),
                       np.ctypeslib.as_array(rp, (n.value + 1,)).copy()), shape=(n.value, m.value))
    for p in (rp, co, va):
        L.ab200_mm_free(p)
    return 0, A


def write_mtx(path, A, base=1, banner=True, with_nnz=True, shuffle_seed=None):
    A = sp.coo_matrix(A)
    idx = np.arange(A.nnz)
    if shuffle_seed is not None:
        np.random.default_rng(shuffle_seed).shuffle(idx)
    with open(path, "w") as f:
        if banner:
            f.write("%%MatrixMarket matrix coordinate real general\n% a comment\n\n")
        f.write(f"{A.shape[0]} {A.shape[1]}" + (f" {A.nnz}" if with_nnz else "") + "\n")
        for k in idx:
            f.write(f"  {A.row[k] + base} {A.col[k] + base} {float(A.data[k])!r}\n")


def test_one_based_with_banner_comments_and_shuffled_entries(L, tmp_path):
    A = sp.random(40, 40, density=0.1, random_state=1, format="csr") + sp.eye(40)
    write_mtx(tmp_path / "a.mtx", A, base=1, shuffle_seed=3)
    rc, B = read(L, tmp_path / "a.mtx")
    assert rc == 0 and B.shape == (40, 40)
    assert abs(B - A).max() == 0.0
    assert B.has_sorted_indices


def test_zero_based_is_autodetected_and_nnz_is_optional(L, tmp_path):
    """arpackSolver.hpp:405-413: indices are 1-based only when max(i) == n or max(j) == m."""
    A = sp.random(25, 25, density=0.2, random_state=2, format="csr") + sp.eye(25)
    write_mtx(tmp_path / "z.mtx", A, base=0, banner=False, with_nnz=False)
    rc, B = read(L, tmp_path / "z.mtx")
    assert rc == 0 and abs(B - A).max() == 0.0


def test_duplicates_are_summed_like_setfromtriplets(L, tmp_path):
    with open(tmp_path / "d.mtx", "w") as f:
        f.write("3 3 6\n1 1 1.5\n3 3 2.0\n1 1 0.25\n2 3 -1\n2 3 -2\n3 1 4\n")
    rc, B = read(L, tmp_path / "d.mtx")
    assert rc == 0
    assert np.array_equal(B.toarray(), np.array([[1.75, 0, 0], [0, 0, -3.0], [4.0, 0, 2.0]]))


def test_rectangular_and_empty_rows(L, tmp_path):
    A = sp.csr_matrix(np.array([[0, 0, 1.0, 0], [0, 0, 0, 0], [2.0, 0, 0, 3.0]]))
    write_mtx(tmp_path / "r.mtx", A, base=1)
    rc, B = read(L, tmp_path / "r.mtx")
    assert rc == 0 and B.shape == (3, 4) and abs(B - A).max() == 0.0


@pytest.mark.parametrize("body,code", [("", 2), ("x y\n", 2), ("3 3 1\n1 one 2.0\n", 3), ("3 3 1\n1 2\n", 3),
                                       ("3 3 1\n7 1 1.0\n", 4)])
def test_malformed_files_are_errors(L, tmp_path, body, code):
    p = tmp_path / "bad.mtx"
    p.write_text(body)
    rc, _ = read(L, p)
    assert rc == code
    assert read(L, tmp_path / "missing.mtx")[0] == 1


def test_restart_files_roundtrip_and_zero_guard(L, tmp_path):
    """saveSolve/restartSolve: count line, one value per line; resid entries below 1e-6 become eps on load."""
    x = np.array([1.25, -3.5e-9, 0.0, 2.0e10, -7.125])
    p = str(tmp_path / "arpackSolver.resid.out").encode()
    assert L.ab200_restart_save_f64(p, x.size, x.ctypes.data) == 0
    lines = open(p.decode()).read().split()
    assert int(lines[0]) == 5 and len(lines) == 6
    y = np.zeros(5)
    assert L.ab200_restart_load_f64(p, 5, y.ctypes.data, 1) == 0
    assert np.array_equal(x, y)                      # 17 significant digits: exact round trip
    assert L.ab200_restart_load_f64(p, 5, y.ctypes.data, 0) == 0
    eps = np.finfo(np.float64).eps
    assert np.array_equal(y, np.array([1.25, eps, eps, 2.0e10, -7.125]))
    assert L.ab200_restart_load_f64(p, 6, y.ctypes.data, 1) == 2      # "bad dim - restart KO"
    assert L.ab200_restart_load_f64(str(tmp_path / "nope").encode(), 5, y.ctypes.data, 1) == 1


def test_complex_coordinate_files(L, tmp_path):
    """Complex files in the style of the reference's fixtures EXAMPLES/MATRIX_MARKET/Az.mtx / Bz.mtx: 0-based, no nnz,
    values written "(re, im)" with blanks inside the parentheses; also "(re)" and a bare real, which the stream
    extraction of std::complex the reference uses (arpackSolver.hpp:398-400) accepts; duplicates are summed."""
    L.ab200_mm_read_csr_z.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                                      C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
                                      C.POINTER(C.POINTER(C.c_double))]
    p = tmp_path / "Az.mtx"
    p.write_text("%% MatrixMarket matrix coordinate complex general\n% 0-based without (optional) nnz\n%\n4 4\n\n"
                 "0  0  (  1., 0.)\n1  1  (200., -3.5)\n2  2  (200.)\n3  3  7.25\n"
                 "1  0  (-100., 2.)\n0  1  ( -100.,2. )\n1  0  (0.5, 0.5)\n3  0  (0., -1.)\n")
    n, m, nnz = C.c_int(), C.c_int(), C.c_longlong()
    rp, co, va = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    assert L.ab200_mm_read_csr_z(str(p).encode(), C.byref(n), C.byref(m), C.byref(nnz), C.byref(rp), C.byref(co),
                                 C.byref(va)) == 0
    vals = np.ctypeslib.as_array(va, (2 * nnz.value,)).copy().view(np.complex128)
    A = sp.csr_matrix((vals, np.ctypeslib.as_array(co, (nnz.value,)).copy(),
                       np.ctypeslib.as_array(rp, (n.value + 1,)).copy()), shape=(n.value, m.value)).toarray()
    for q in (rp, co, va):
        L.ab200_mm_free(q)
    want = np.zeros((4, 4), dtype=complex)
    want[0, 0] = 1.0
    want[1, 1] = 200 - 3.5j
    want[2, 2] = 200.0
    want[3, 3] = 7.25
    want[1, 0] = (-100 + 2j) + (0.5 + 0.5j)
    want[0, 1] = -100 + 2j
    want[3, 0] = -1j
    assert (n.value, m.value, nnz.value) == (4, 4, 7)
    assert np.array_equal(A, want)
    bad = tmp_path / "bad.mtx"
    bad.write_text("2 2\n0 0 (1., \n")
    assert L.ab200_mm_read_csr_z(str(bad).encode(), C.byref(n), C.byref(m), C.byref(nnz), C.byref(rp), C.byref(co),
                                 C.byref(va)) == 3
