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


@pytest.mark.parametrize("extra", [[], ["--simplePrec"], ["--registered"], ["--tol", "1.e-5"], ["--invert"],
                                   ["--schur"]])
@pytest.mark.parametrize("mag", ["LM", "LA", "SA"])
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


@pytest.mark.parametrize("mag", ["LM", "LR", "SR"])
@pytest.mark.parametrize("extra", [[], ["--simplePrec"], ["--registered"]])
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


def test_max_iterations_and_rejected_options(files):
    d = files[0]
    _, err = run(d, "--A", "As.mtx", "--nbEV", 3, "--nbCV", 7, "--mag", "SA", "--maxIt", 1, "--noCheck")
    assert "maximum number of iterations taken" in err           # info = 1 is reported, not fatal (arpackSolver.hpp:791)
    for bad, msg in ((["--cpxPb"], "not built"), (["--dense", "true"], "not built"), (["--slv", "LU"], "not built"),
                     (["--slvItrPC", "ILU"], "not built"), (["--mag", "XX"], "bad --mag"), (["--nbEV"], "need argument"),
                     (["--A", "missing.mtx"], "read A KO")):
        _, err = run(d, *(["--A", "As.mtx"] if "--A" not in bad else []), *bad, expect=1)
        assert msg in err
    out, _ = run(d, "--help")
    assert "--nbEV" in out and "--restart" in out
