"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle is too slow there):
after a fixed restart budget on config 2 (n = 16.7M) the state the RCI leaves behind must satisfy the invariants of
an implicitly restarted Lanczos factorisation  A V_k = V_k H_k + r e_k^T:

  * V^T V = I                      (full re-orthogonalisation, dsaitr.f:569-780)
  * V^T A V = H (tridiagonal), with H read from workl(ipntr(5)) exactly as dseupd reads it (dseupd.f:533-534)
  * ||A x - theta x|| = the Ritz estimate rnorm*|last row| that dseigt reports (dseigt.f:167-169), for every Ritz pair
  * the Ritz values interlace/approach the analytic spectrum 4 - 2cos(i pi/(nx+1)) - 2cos(j pi/(nx+1)) from below
and the counts (OP*x, re-orth steps) follow from the restart budget alone."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import arpack_ng_b200 as m
    m.lib()
    return m


def _lanczos_invariants(ab, A, n, res, ncv, nev):
    import torch
    V = res.v.view(ncv, n)
    G = V @ V.T
    orth = float((G - torch.eye(ncv, dtype=torch.float64, device="cuda")).abs().max())
    AV = torch.empty_like(V)
    for k in range(ncv):
        A(V[k], AV[k])
    T = (V @ AV.T).cpu().numpy()
    ih = res.ipntr[4] - 1
    H = res.workl[ih:ih + 2 * ncv].reshape(2, ncv)  # column 1 = sub-diagonal (h(2:,1)), column 2 = diagonal
    Hd = np.diag(H[1]) + np.diag(H[0][1:], -1) + np.diag(H[0][1:], 1)
    Hd[0, 0] = H[1][0]
    return orth, T, Hd


def test_config2_full_size_invariants(ab):
    import torch
    nx, nev, ncv, restarts = 4096, 10, 40, 2
    n = nx * nx
    A = ab.CsrOperator.laplace2d(nx, nx)
    r0 = ab.hashed_start_vector(n)
    res = ab.solve(A, n, nev, ncv, "LA", tol=1e-10, mxiter=restarts, resid=r0, eupd=False)
    assert res.info == 1  # budget exhausted, as intended
    assert int(res.iparam[2]) == restarts + 1
    nopx = int(res.iparam[8])
    assert int(res.iparam[10]) == nopx - 1  # DGKS fires on every step of a Laplacian
    orth, T, Hd = _lanczos_invariants(ab, A, n, res, ncv, nev)
    assert orth < 1e-12, orth
    # h(1,1) carries rnorm for dseupd (dsaup2.f:645); the rest of H must equal V^T A V
    Tc, Hc = T.copy(), Hd.copy()
    assert np.abs(Tc - Tc.T).max() < 1e-11
    band = np.abs(np.triu(Tc, 2)).max()
    assert band < 1e-11, band
    assert np.abs(np.diag(Tc) - np.diag(Hc)).max() < 1e-11
    assert np.abs(np.diag(Tc, -1) - np.diag(Hc, -1)).max() < 1e-11
    # Ritz values approach the top of the analytic spectrum from below
    i = np.arange(1, 12)
    lam = np.sort((4 - 2 * np.cos((nx + 1 - i)[:, None] * np.pi / (nx + 1)) -
                   2 * np.cos((nx + 1 - i)[None, :] * np.pi / (nx + 1))).ravel())[-nev:]
    theta = np.sort(np.linalg.eigvalsh(Hc))[-nev:]
    assert (theta <= lam + 1e-12).all() and theta[-1] > 7.9


def test_config2_ritz_estimates_are_true_residuals(ab):
    """dseupd on an unconverged factorisation is not allowed by the reference (info=-14/-17), so the Ritz pairs are
    formed here from H and V; their residual norms must equal dseigt's estimates rnorm*|s_last|."""
    import torch
    nx, nev, ncv = 2048, 10, 40
    n = nx * nx
    A = ab.CsrOperator.laplace2d(nx, nx)
    res = ab.solve(A, n, nev, ncv, "LA", tol=1e-10, mxiter=3, resid=ab.hashed_start_vector(n), eupd=False)
    V = res.v.view(ncv, n)
    ih = res.ipntr[4] - 1
    H = res.workl[ih:ih + 2 * ncv].reshape(2, ncv)
    rnorm = H[0][0]
    Hd = np.diag(H[1]) + np.diag(H[0][1:], -1) + np.diag(H[0][1:], 1)
    th, S = np.linalg.eigh(Hd)
    resid_true_norm = float(torch.linalg.norm(res.resid))
    assert abs(resid_true_norm - rnorm) <= 1e-12 * max(1.0, rnorm)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for k in (-1, -2, -5, -10):
        x = (torch.as_tensor(S[:, k], device="cuda") @ V).contiguous()
        A(x, y)
        true = float(torch.linalg.norm(y - th[k] * x))
        est = rnorm * abs(S[-1, k])
        assert abs(true - est) <= 1e-9 * max(est, 1e-6) + 1e-12, (k, true, est)


def test_config4_nonsym_arnoldi_relation(ab):
    """BASELINE config 4 at full size (n = 4.2M, rho = 100): after a fixed budget the Arnoldi relation
    V^T A V = H (upper Hessenberg, workl(ipntr(5))) and V^T V = I must hold."""
    import torch
    nx, nev, ncv = 2048, 6, 30
    n = nx * nx
    A = ab.CsrOperator.convdiff2d(nx, 100.0)
    res = ab.solve(A, n, nev, ncv, "LR", sym=False, tol=1e-10, mxiter=2, resid=ab.hashed_start_vector(n), eupd=False)
    assert res.info == 1
    V = res.v.view(ncv, n)
    G = V @ V.T
    assert float((G - torch.eye(ncv, dtype=torch.float64, device="cuda")).abs().max()) < 1e-12
    AV = torch.empty_like(V)
    for k in range(ncv):
        A(V[k], AV[k])
    T = (V @ AV.T).cpu().numpy()  # T[i, j] = v_i^T A v_j
    ih = res.ipntr[4] - 1
    H = res.workl[ih:ih + ncv * ncv].reshape(ncv, ncv).T.copy()
    H[2, 0] = 0.0  # h(3,1) carries rnorm for dneupd (dnaup2.f:552)
    scale = np.abs(H).max()
    assert np.abs(np.tril(T, -2)).max() < 1e-11 * scale
    assert np.abs(T - H).max() < 1e-10 * scale
