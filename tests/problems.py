"""Operators shared by the tests, built with numpy/scipy on the CPU (the oracle side of every parity test)."""
import numpy as np
import scipy.sparse as sp


def dssimp_av(nx):
    """OP of EXAMPLES/SIMPLE/dssimp.f:484-540: 2-D 5-point Laplacian on nx x nx scaled by (nx+1)^2."""
    h2 = 1.0 / ((nx + 1) ** 2)

    def av(x):
        X = x.reshape(nx, nx)
        Y = 4 * X.copy()
        Y[:, 1:] -= X[:, :-1]
        Y[:, :-1] -= X[:, 1:]
        Y[1:, :] -= X[:-1, :]
        Y[:-1, :] -= X[1:, :]
        return (Y / h2).ravel()

    return av


def dssimp_exact(nx, k):
    """k largest eigenvalues of the dssimp operator, ascending."""
    i = np.arange(1, nx + 1)
    lam = ((nx + 1) ** 2) * (4 - 2 * np.cos(i[:, None] * np.pi / (nx + 1)) - 2 * np.cos(i[None, :] * np.pi / (nx + 1)))
    return np.sort(lam.ravel())[-k:]


def laplace2d(nx, ny, scale=1.0):
    """Row-major (iy, ix) ordering, ix fastest; stencil (4, -1) * scale (BASELINE config 2 uses scale = 1)."""
    Tx = sp.diags([-np.ones(nx - 1), 2 * np.ones(nx), -np.ones(nx - 1)], [-1, 0, 1])
    Ty = sp.diags([-np.ones(ny - 1), 2 * np.ones(ny), -np.ones(ny - 1)], [-1, 0, 1])
    A = sp.kron(sp.eye(ny), Tx) + sp.kron(Ty, sp.eye(nx))
    A = (scale * A).tocsr()
    A.sort_indices()
    return A


def laplace3d(nx, ny, nz):
    Tx = sp.diags([-np.ones(nx - 1), 2 * np.ones(nx), -np.ones(nx - 1)], [-1, 0, 1])
    Ty = sp.diags([-np.ones(ny - 1), 2 * np.ones(ny), -np.ones(ny - 1)], [-1, 0, 1])
    Tz = sp.diags([-np.ones(nz - 1), 2 * np.ones(nz), -np.ones(nz - 1)], [-1, 0, 1])
    A = sp.kron(sp.eye(nz), sp.kron(sp.eye(ny), Tx)) + sp.kron(sp.eye(nz), sp.kron(Ty, sp.eye(nx))) + \
        sp.kron(Tz, sp.eye(nx * ny))
    A = A.tocsr()
    A.sort_indices()
    return A


def convdiff2d(nx, rho=100.0):
    """EXAMPLES/NONSYM/dndrv1.f:453-470 with rho as in EXAMPLES/SIMPLE/dnsimp.f:570."""
    h = 1.0 / (nx + 1)
    dd, dl, du = 4.0 / h ** 2, -1.0 / h ** 2 - 0.5 * rho / h, -1.0 / h ** 2 + 0.5 * rho / h
    T = sp.diags([dl * np.ones(nx - 1), dd * np.ones(nx), du * np.ones(nx - 1)], [-1, 0, 1])
    off = sp.diags([np.ones(nx - 1), np.ones(nx - 1)], [-1, 1])
    A = (sp.kron(sp.eye(nx), T) + sp.kron(off, -sp.eye(nx) / h ** 2)).tocsr()
    A.sort_indices()
    return A


def complex_tridiag(n, rho=10.0, imag_span=50.0):
    """1-D convection-diffusion (EXAMPLES/COMPLEX/zndrv1.f:400-440 style: 2/h^2 on the diagonal, -1/h^2 -+ rho/(2h) off
    it) with an imaginary ramp added to the diagonal so that the matrix and its spectrum are genuinely complex."""
    h = 1.0 / (n + 1)
    A = sp.diags([(-1 / h ** 2 - rho / 2 / h) * np.ones(n - 1),
                  (2 / h ** 2 + 0j) * np.ones(n) + 1j * np.linspace(0, imag_span, n),
                  (-1 / h ** 2 + rho / 2 / h) * np.ones(n - 1)], [-1, 0, 1]).tocsr().astype(np.complex128)
    A.sort_indices()
    return A


def complex_convdiff2d(nx, rho=10.0, imag_span=200.0):
    """The dndrv1-style 2-D operator plus i*diag(ramp): non-Hermitian, complex spectrum."""
    A = convdiff2d(nx, rho).astype(np.complex128)
    n = nx * nx
    A = (A + sp.diags([1j * np.linspace(0, imag_span, n)], [0])).tocsr()
    A.sort_indices()
    return A
