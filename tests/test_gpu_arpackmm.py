"""arpackmm_b200, the B200 twin of EXAMPLES/MATRIX_MARKET/arpackmm.cpp, driven like arpackmm.sh drives the reference:
option sweeps on small symmetric / non-symmetric / generalised problems, success = exit code 0 with the built-in
check ||A v - lambda B v|| <= sqrt(tol) (arpackSolver.hpp:297-352), plus the eigenvalues against dense solutions."""
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp

from test_mmio import write_mtx

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(_ROOT, "arpack-ng_b200", "bin", "arpackmm_b200")


def run(cwd, *args, expect=0):
    p = subprocess.run([EXE, *[str(a) for a in args]], cwd=cwd, capture_output=True, text=True, timeout=300)
    assert p.returncode == expect, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout, p.stderr


def values(out):
    return [complex(float(a), float(b)) for a, b in re.findall(r"Ritz value\s+\d+: \(([-+.\de]+),([-+.\de]+)\)", out)]


def found(out):
    m = re.search(r"OUT: mode (\d+), nb EV found (\d+), nb iterations (\d+)", out)
    return tuple(int(x) for x in m.groups())


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("mm")
    n = 60
    # symmetric positive definite (1-D Laplacian + a diagonal ramp): well separated ends of the spectrum
    As = sp.diags([-np.ones(n - 1), 2.0 + 0.05 * np.arange(n), -np.ones(n - 1)], [-1, 0, 1]).tocsr()
    # non-symmetric convection-diffusion-like band with complex pairs
    An = sp.diags([-1.3 * np.ones(n - 1), 2.0 + 0.03 * np.arange(n), -0.4 * np.ones(n - 1), 0.2 * np.ones(n - 5)],
                  [-1, 0, 1, 5]).tocsr()
    Bm = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).tocsr()   # SPD mass matrix
    write_mtx(d / "As.mtx", As, base=1, shuffle_seed=1)
    write_mtx(d / "An.mtx", An, base=0, banner=False, with_nnz=False)
    write_mtx(d / "B.mtx", Bm, base=1)
    return d, As, An, Bm


@pytest.mark.parametrize("mag,extra", [("LM", []), ("LA", []), ("SA", []), ("LM", ["--simplePrec"]), ("LA", ["--registered"]),
                                       ("SA", ["--tol", "1.e-5"]), ("LM", ["--invert"]), ("LA", ["--schur"]),
                                       ("SA", ["--registered"])])
def test_symmetric_standard(files, extra, mag):
    d, As, _, _ = files
    out, _ = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 20, "--mag", mag, "--maxIt", 500, "--verbose", 1, *extra)
    mode, nb, _ = found(out)
    assert mode == 1 and nb == 3
    ev = np.linalg.eigvalsh(As.toarray())
    want = ev[:3] if mag == "SA" else ev[-3:]
    got = np.sort([v.real for v in values(out)]) if "--schur" not in extra else None
    if got is not None:
        assert np.abs(got - want).max() < (1e-3 if "--simplePrec" in extra else 1e-6)


def test_symmetric_shift_is_undone(files):
    d, As, _, _ = files
    out, _ = run(d, "--A", "As.mtx", "--nbEV", 2, "--nbCV", 16, "--mag", "LM", "--shiftReal", 100.0, "--maxIt", 500,
                 "--verbose", 1)
    assert "backTransform yes" in out
    ev = np.linalg.eigvalsh(As.toarray())
    # largest |lambda - 100| = the smallest eigenvalues, reported un-shifted
    assert np.abs(np.sort([v.real for v in values(out)]) - ev[:2]).max() < 1e-6


@pytest.mark.parametrize("mag,extra", [("LM", []), ("LR", []), ("SR", []), ("LR", ["--simplePrec"]), ("SR", ["--registered"])])
def test_nonsymmetric_standard(files, mag, extra):
    d, _, An, _ = files
    out, _ = run(d, "--nonSymPb", "--A", "An.mtx", "--nbEV", 4, "--nbCV", 24, "--mag", mag, "--maxIt", 1000,
                 "--verbose", 1, *extra)
    mode, nb, _ = found(out)
    assert mode == 1 and nb >= 4
    ev = np.linalg.eigvals(An.toarray())
    tol = 1e-3 if "--simplePrec" in extra else 1e-6
    for v in values(out):
        assert np.abs(ev - v).min() < tol


@pytest.mark.parametrize("slv", ["CG", "BiCG"])
def test_generalised_mode2_and_mode3(files, slv):
    d, As, _, Bm = files
    gev = np.sort(sl.eigh(As.toarray(), Bm.toarray(), eigvals_only=True))
    out, _ = run(d, "--A", "As.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20, "--mag", "LM", "--maxIt", 500, "--slv", slv,
                 "--slvItrTol", "1.e-12", "--slvItrMaxIt", 500, "--verbose", 1)
    assert found(out)[0] == 2
    assert np.abs(np.sort([v.real for v in values(out)]) - gev[-3:]).max() < 1e-6 * gev[-1]
    # shift-invert around sigma = 1: eigenvalues nearest 1 (mode 3, OP = (A - sigma B)^-1 B)
    out, _ = run(d, "--A", "As.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20, "--mag", "LM", "--shiftReal", 1.0, "--maxIt",
                 500, "--slv", slv, "--slvItrTol", "1.e-12", "--slvItrMaxIt", 2000, "--verbose", 1)
    assert found(out)[0] == 3
    near = gev[np.argsort(np.abs(gev - 1.0))[:3]]
    assert np.abs(np.sort([v.real for v in values(out)]) - np.sort(near)).max() < 1e-6


def test_restart_files(files):
    """arpackmm always dumps resid/v; --restart feeds them back (info = 1) and finds the same eigenvalues."""
    d, As, _, _ = files
    out1, _ = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 20, "--mag", "LA", "--maxIt", 500, "--verbose", 1)
    n = As.shape[0]
    assert int(open(d / "arpackSolver.resid.out").readline()) == n
    assert int(open(d / "arpackSolver.v.out").readline()) == n * 20
    out2, _ = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 20, "--mag", "LA", "--maxIt", 500, "--verbose", 1, "--restart")
    assert "restart OK" in out2
    assert np.abs(np.sort([v.real for v in values(out1)]) - np.sort([v.real for v in values(out2)])).max() < 1e-8
    # a dump of another problem size is refused
    _, err = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 12, "--restart", expect=1)
    assert "restart KO" in err or "bad restart" in err


def test_max_iterations_and_bad_options(files):
    d = files[0]
    _, err = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 7, "--mag", "SA", "--maxIt", 1, "--noCheck")
    assert "maximum number of iterations taken" in err           # info = 1 is reported, not fatal (arpackSolver.hpp:791)
    for bad, msg in ((["--dense", "maybe"], "bad --dense"), (["--dense", "true"], "does not support iterative solvers"),
                     (["--genPb", "--slv", "XX"], "bad --slv"), (["--genPb", "--slvItrPC", "SSOR"], "bad --slvItrPC"),
                     (["--mag", "XX"], "bad --mag"), (["--nbEV"], "need argument"), (["--A", "missing.mtx"], "read A KO")):
        _, err = run(d, *(["--A", "As.mtx"] if "--A" not in bad else []), *bad, expect=1)
        assert msg in err
    # the reference's parser skips what it does not know (arpackmm.sh passes a bare "LA"); so does this one, loudly
    out, err = run(d, "--A", "As.mtx", "LA", "--nbEV", 2, "--nbCV", 12, "--maxIt", 500)
    assert "unknown option LA ignored" in err and found(out)[1] == 2
    out, _ = run(d, "--help")
    assert "--nbEV" in out and "--restart" in out and "--slvItrPC" in out and "--dense" in out


def write_mtx_complex(path, A, base=0):
    """the reference's complex files (Az.mtx, Bz.mtx): 'i j (re, im)' entries, no banner needed"""
    A = sp.coo_matrix(A)
    with open(path, "w") as f:
        f.write("% complex coordinate file\n\n")
        f.write(f"{A.shape[0]} {A.shape[1]}\n\n")
        for r, c, v in zip(A.row, A.col, A.data):
            f.write(f"{r + base}  {c + base}  ({float(v.real)!r}, {float(v.imag)!r})\n")


# (solver options, extra options, also run mode 3) -- a cut of arpackmm.sh's --slv x --dense x --simplePrec sweep
GEN_CASES = [
    (["--slv", "BiCG", "--slvItrPC", "ILU#1.e-06#2"], [], True),
    (["--slv", "CG", "--slvItrPC", "ILU"], [], False),          # "ILU" alone: drop tolerance 1, i.e. almost Jacobi
    (["--slv", "BiCG"], ["--simplePrec"], False),
    (["--slv", "LU"], [], True),
    (["--slv", "LU"], ["--simplePrec"], False),
    (["--slv", "QR", "--slvDrtPivot", "1.e-06"], [], False),
    (["--slv", "LLT"], [], True),
    (["--slv", "LDLT", "--slvDrtScale", "1."], [], True),
    (["--slv", "LU", "--dense", "true"], [], False),
    (["--slv", "QR", "--dense", "false"], ["--simplePrec"], False),
    (["--slv", "LDLT", "--dense", "false"], [], True),
]


@pytest.mark.parametrize("slv,prec,mode3", GEN_CASES, ids=lambda s: "-".join(x.strip("-") for x in s) if isinstance(s, list) else "")
def test_generalised_every_solver(files, slv, prec, mode3):
    """arpackmm.sh's --slv sweep on a symmetric generalised problem: mode 2 (B^-1 A, B SPD) and mode 3 (shift-invert
    around 1; A - B is indefinite there, which LLT must refuse and LDLT must survive by falling back)."""
    d, As, _, Bm = files
    gev = np.sort(sl.eigh(As.toarray(), Bm.toarray(), eigvals_only=True))
    itr = ["--slvItrTol", "1.e-7" if prec else "1.e-12", "--slvItrMaxIt", 2000] if slv[1] in ("BiCG", "CG") else []
    tol = 2e-3 if prec else 1e-6
    out, _ = run(d, "--A", "As.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20, "--mag", "LM", "--maxIt", 500, "--verbose", 1,
                 *slv, *itr, *prec)
    assert found(out)[0] == 2
    assert np.abs(np.sort([v.real for v in values(out)]) - gev[-3:]).max() < tol * gev[-1]
    assert re.search(r"inner solver %s: \d+ solves" % slv[1], out)
    if not mode3:
        return
    args = ["--A", "As.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20, "--mag", "LM", "--shiftReal", 1.0, "--maxIt", 500,
            "--verbose", 1, *slv, *itr, *prec]
    if slv[1] == "LLT":
        _, err = run(d, *args, expect=1)
        assert "not positive definite" in err
        return
    out, err = run(d, *args)
    assert found(out)[0] == 3
    if slv[1] == "LDLT":
        assert "Warning: LDLT" in err
    near = gev[np.argsort(np.abs(gev - 1.0))[:3]]
    assert np.abs(np.sort([v.real for v in values(out)]) - np.sort(near)).max() < tol


def test_cholesky_offset_and_scale(files):
    """--slvDrtOffset / --slvDrtScale change the diagonal the sparse Cholesky factorises (Eigen's setShift), so the inner
    solves belong to a modified B.  Arnoldi (dn*upd) does not need OP to be self-adjoint in the B inner product: its Ritz
    values are eigenvalues of Bmod^-1 A."""
    d, As, _, Bm = files
    Bmod = Bm.toarray().copy()
    Bmod[np.diag_indices_from(Bmod)] = 0.5 + 2.0 * np.diag(Bmod)
    out, _ = run(d, "--nonSymPb", "--A", "As.mtx", "--genPb", "--nbEV", 2, "--nbCV", 16, "--maxIt", 500, "--verbose", 1,
                 "--slv", "LLT", "--slvDrtOffset", 0.5, "--slvDrtScale", 2.0, "--noCheck")
    ev = np.linalg.eigvals(np.linalg.solve(Bmod, As.toarray()))
    got = values(out)
    assert len(got) >= 2
    for v in got:
        assert np.abs(ev - v).min() < 1e-5 * np.abs(ev).max()


@pytest.mark.parametrize("slv", [["--slv", "LU"], ["--slv", "QR", "--dense", "false"]], ids=["LU", "denseQR"])
def test_nonsymmetric_generalised_shift_invert(files, slv):
    d, _, An, Bm = files
    ev = sl.eigvals(An.toarray(), Bm.toarray())
    out, _ = run(d, "--nonSymPb", "--A", "An.mtx", "--genPb", "--nbEV", 4, "--nbCV", 24, "--shiftReal", 2.0, "--maxIt",
                 1000, "--verbose", 1, *slv)
    mode, nb, _ = found(out)
    assert mode == 3 and nb >= 4
    near = ev[np.argsort(np.abs(ev - 2.0))[:4]]
    got = values(out)
    for v in near:
        assert min(abs(v - g) for g in got) < 1e-6


@pytest.fixture(scope="module")
def zfiles(tmp_path_factory):
    d = tmp_path_factory.mktemp("mmz")
    n = 48
    k = np.arange(n)
    Az = sp.diags([(-1.0 - 0.1j) * np.ones(n - 1), (2.0 + 0.05 * k) + 0.3j * np.cos(k), (-1.0 + 0.3j) * np.ones(n - 1)],
                  [-1, 0, 1]).tocsr()
    Bz = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).astype(complex).tocsr()
    write_mtx_complex(d / "Az.mtx", Az, base=0)
    write_mtx_complex(d / "Bz.mtx", Bz, base=1)
    return d, Az, Bz


@pytest.mark.parametrize("extra", [[], ["--simplePrec"], ["--mag", "LI"], ["--shiftReal", "3.0"],
                                   ["--shiftReal", "3.0", "--shiftImag", "1.0"]],
                         ids=lambda e: "-".join(x.strip("-") for x in e) or "LM")
def test_complex_standard(zfiles, extra):
    """--cpxPb: zn[ae]upd / cn[ae]upd on a complex non-Hermitian matrix (arpackmm.sh's third eigPb)."""
    d, Az, _ = zfiles
    out, _ = run(d, "--nonSymPb", "--cpxPb", "--A", "Az.mtx", "--nbEV", 3, "--nbCV", 20, "--maxIt", 1000, "--verbose", 1, *extra)
    mode, nb, _ = found(out)
    assert mode == 1 and nb == 3
    assert ("backTransform yes" in out) == (extra == ["--shiftReal", "3.0"])   # only a purely real shift is applied
    ev = np.linalg.eigvals(Az.toarray())
    tol = 2e-3 if "--simplePrec" in extra else 1e-6
    got = values(out)
    for v in got:
        assert np.abs(ev - v).min() < tol
    if extra == ["--shiftReal", "3.0"]:
        want = ev[np.argsort(-np.abs(ev - 3.0))[:3]]
        for v in want:
            assert min(abs(v - g) for g in got) < tol


@pytest.mark.parametrize("slv", [["--slv", "BiCG", "--slvItrPC", "ILU#1.e-08#4", "--slvItrTol", "1.e-12", "--slvItrMaxIt", 500],
                                 ["--slv", "LU"], ["--slv", "QR", "--dense", "false"]], ids=["BiCG-ILU", "LU", "denseQR"])
def test_complex_generalised(zfiles, slv):
    d, Az, Bz = zfiles
    ev = sl.eigvals(Az.toarray(), Bz.toarray())
    # mode 2: OP = B^-1 A, largest magnitude
    out, _ = run(d, "--nonSymPb", "--cpxPb", "--A", "Az.mtx", "--B", "Bz.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20,
                 "--maxIt", 1000, "--verbose", 1, *slv)
    assert found(out)[0] == 2
    want = ev[np.argsort(-np.abs(ev))[:3]]
    got = values(out)
    for v in want:
        assert min(abs(v - g) for g in got) < 1e-6 * np.abs(ev).max()
    if "--dense" in slv:
        return
    # mode 3 with a complex shift: the eigenvalues nearest sigma
    sigma = 6.0 + 1.0j
    out, _ = run(d, "--nonSymPb", "--cpxPb", "--A", "Az.mtx", "--B", "Bz.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20,
                 "--shiftReal", sigma.real, "--shiftImag", sigma.imag, "--maxIt", 1000, "--verbose", 1, *slv)
    assert found(out)[0] == 3
    want = ev[np.argsort(np.abs(ev - sigma))[:3]]
    got = values(out)
    for v in want:
        assert min(abs(v - g) for g in got) < 1e-6


def test_complex_llt_and_restart(zfiles):
    d, Az, Bz = zfiles
    ev = sl.eigvals(Az.toarray(), Bz.toarray())
    args = ["--nonSymPb", "--cpxPb", "--A", "Az.mtx", "--B", "Bz.mtx", "--genPb", "--nbEV", 3, "--nbCV", 20, "--maxIt", 1000,
            "--verbose", 1, "--slv", "LLT"]
    out1, _ = run(d, *args)    # B is Hermitian positive definite: complex sparse Cholesky
    want = ev[np.argsort(-np.abs(ev))[:3]]
    for v in want:
        assert min(abs(v - g) for g in values(out1)) < 1e-6 * np.abs(ev).max()
    first = open(d / "arpackSolver.resid.out").read().split("\n")
    assert int(first[0]) == Az.shape[0] and first[1].startswith("(")        # complex dumps: "(re,im)" lines
    out2, _ = run(d, *args, "--restart")
    assert "restart OK" in out2
    for v in want:
        assert min(abs(v - g) for g in values(out2)) < 1e-6 * np.abs(ev).max()


def test_stat_lines(files):
    """the STAT block of arpackmm.cpp:1046-1060 (stat_c counters of the solve that just ran)"""
    d = files[0]
    out, _ = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 20, "--mag", "LA", "--maxIt", 500)
    m = re.search(r"STAT: total number of user OP\*x operation\s+(\d+)", out)
    assert m and int(m.group(1)) > 20
    assert re.search(r"STAT: total number of reorthogonalization steps taken\s+\d+", out)
