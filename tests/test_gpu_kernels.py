"""Kernel-level GPU tests: every tall-skinny kernel against a plain PyTorch FP64 reference of the same op,
independent of the solver (the reference has no unit tests of its inner routines; these are the
per-kernel checks the build plan asks for).  Tolerances are for FP64 sums of O(n) terms with different summation
order: relative 1e-12 on reductions, 1e-13 elementwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import arpack_ng_b200 as m
    m.lib()
    return m


def _ref_orth(V, w):
    h = V.T @ w
    w2 = w @ w
    r = w - V @ h
    s = V.T @ r
    r2 = r @ r
    return h, w2, r, s, r2


@pytest.mark.parametrize("mode", ["auto", "generic"])
@pytest.mark.parametrize("n,ldv", [(1000, 1000), (4097, 4098), (4097, 4097), (300001, 300002), (1 << 20, 1 << 20)])
@pytest.mark.parametrize("j", [1, 2, 7, 8, 9, 16, 20, 24, 31, 32, 33, 40, 48, 57, 64])
def test_orth_step_kernels(ab, mode, n, ldv, j):
    """K4..K10 fused step: h = V^T w, r = w - V h, s = V^T r (speculative), conditional r -= V s.
    Covers every ring depth / box count of the TMA kernels (j = 1..64), odd sizes (tail tiles), and an odd
    leading dimension (-> generic kernels)."""
    import torch
    if n * j > 3e7 and j not in (20, 24, 32, 40, 64):
        pytest.skip("large case only for a subset of j")
    L = ab.lib()
    L.ab200_set_kernel_mode(0 if mode == "auto" else 1)
    try:
        g = torch.Generator(device="cuda").manual_seed(1234 + j)
        Vfull = torch.zeros(ldv * j, dtype=torch.float64, device="cuda")
        V = Vfull.view(j, ldv)[:, :n]  # column-major n x j with leading dimension ldv
        # nearly orthonormal columns (as in a Lanczos basis) so that the DGKS test is meaningful
        Q, _ = torch.linalg.qr(torch.randn(n, j, dtype=torch.float64, device="cuda", generator=g))
        V.copy_(Q.T)
        Vm = V.T  # n x j view
        for fire in (False, True):
            w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
            if fire:  # w almost inside span(V): ||r|| << ||w|| -> the reference's DGKS pass must run
                c = torch.randn(j, dtype=torch.float64, device="cuda", generator=g)
                w = Vm @ (c / torch.linalg.norm(c)) + 1e-3 * w / torch.linalg.norm(w)
            w = w.contiguous()
            resid = torch.zeros(n, dtype=torch.float64, device="cuda")
            out = np.zeros(2 * j + 4)
            assert L.ab200_debug_orth_f64(n, j, Vfull.data_ptr(), ldv, w.data_ptr(), resid.data_ptr(),
                                          out.ctypes.data) == 0
            h, w2, r, s, r2 = _ref_orth(Vm, w)
            scale = float(torch.sqrt(w2))
            assert np.allclose(out[:j], h.cpu().numpy(), rtol=0, atol=1e-12 * scale)
            assert abs(out[j] - float(w2)) <= 1e-12 * float(w2)
            assert np.allclose(out[j + 1:2 * j + 1], s.cpu().numpy(), rtol=0, atol=1e-12 * scale)
            assert abs(out[2 * j + 1] - float(r2)) <= 1e-9 * float(r2) + 1e-24 * float(w2)
            fired = not (np.sqrt(out[2 * j + 1]) > np.float64(np.float32(0.717)) * np.sqrt(out[j]))
            assert fired == fire
            assert out[2 * j + 3] == (1.0 if fire else 0.0)
            if fire:
                r1 = r - Vm @ s
                assert np.allclose(resid.cpu().numpy(), r1.cpu().numpy(), rtol=0, atol=1e-13 * scale)
                assert abs(out[2 * j + 2] - float(r1 @ r1)) <= 1e-9 * float(r1 @ r1) + 1e-24 * float(w2)
            else:
                assert np.allclose(resid.cpu().numpy(), r.cpu().numpy(), rtol=0, atol=1e-13 * scale)
    finally:
        L.ab200_set_kernel_mode(0)


@pytest.mark.parametrize("mode", ["auto", "generic"])
@pytest.mark.parametrize("n", [777, 4096, 200001, 1 << 19])
@pytest.mark.parametrize("kin,kout,beta_col", [(20, 5, 4), (20, 4, -1), (40, 11, 10), (40, 17, 16), (30, 9, 8),
                                               (64, 21, 20), (40, 40, -1), (64, 64, 3), (19, 1, 0)])
def test_vq_update_kernels(ab, mode, n, kin, kout, beta_col):
    """K12..K16: in place V(:,0:kout) = V(:,0:kin) Q, resid = sigma resid + beta Vnew(:,beta_col), ||resid||^2."""
    import torch
    L = ab.lib()
    L.ab200_set_kernel_mode(0 if mode == "auto" else 1)
    try:
        g = torch.Generator(device="cuda").manual_seed(99 + kin * 64 + kout)
        ldv = n + (n & 1)
        Vfull = torch.zeros(ldv * kin, dtype=torch.float64, device="cuda")
        V = Vfull.view(kin, ldv)[:, :n]
        V.copy_(torch.randn(kin, n, dtype=torch.float64, device="cuda", generator=g))
        V0 = V.clone()
        Qh = np.random.default_rng(kin + kout).standard_normal((kin, kout))
        resid = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        r0 = resid.clone()
        sigma, beta = 0.37, (0.0 if beta_col < 0 else -1.25)
        nrm2 = np.zeros(1)
        qcm = np.asfortranarray(Qh)
        assert L.ab200_debug_vq_f64(n, kin, kout, Vfull.data_ptr(), ldv, qcm.ctypes.data, sigma, beta, beta_col,
                                    resid.data_ptr(), nrm2.ctypes.data) == 0
        ref = torch.as_tensor(Qh, device="cuda").T @ V0  # kout x n
        assert torch.allclose(V[:kout], ref, rtol=0, atol=1e-12 * float(V0.abs().max()) * kin)
        assert torch.equal(V[kout:], V0[kout:])  # untouched columns stay bit-identical
        rr = sigma * r0 + (beta * ref[beta_col] if beta_col >= 0 else 0.0)
        assert torch.allclose(resid, rr, rtol=0, atol=1e-12 * float(rr.abs().max()))
        assert abs(nrm2[0] - float(rr @ rr)) <= 1e-11 * float(rr @ rr)
    finally:
        L.ab200_set_kernel_mode(0)


@pytest.mark.parametrize("kin,kout,npx", [(64, 21, 44), (40, 11, 30), (64, 31, 33), (24, 9, 2)])
def test_vq_update_banded_q_as_the_restart_produces_it(ab, kin, kout, npx):
    """After dsapps' QR sweeps column c of Q is zero below row np + c (dsapps.f:461 multiplies kplusp-i+1 entries only):
    the tensor-core kernel skips the k-steps that hold nothing but those zeros (vq_mma.cu) -- same result."""
    import torch
    L = ab.lib()
    n = 50000 + 37
    g = torch.Generator(device="cuda").manual_seed(5 + kin)
    ldv = n + (n & 1)
    Vfull = torch.zeros(ldv * kin, dtype=torch.float64, device="cuda")
    V = Vfull.view(kin, ldv)[:, :n]
    V.copy_(torch.randn(kin, n, dtype=torch.float64, device="cuda", generator=g))
    V0 = V.clone()
    Qh = np.random.default_rng(kin * kout).standard_normal((kin, kout))
    for c in range(kout):
        Qh[npx + c + 1:, c] = 0.0
    resid = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    r0 = resid.clone()
    nrm2 = np.zeros(1)
    qcm = np.asfortranarray(Qh)
    assert L.ab200_debug_vq_f64(n, kin, kout, Vfull.data_ptr(), ldv, qcm.ctypes.data, 0.5, 2.0, kout - 1,
                                resid.data_ptr(), nrm2.ctypes.data) == 0
    ref = torch.as_tensor(Qh, device="cuda").T @ V0
    assert torch.allclose(V[:kout], ref, rtol=0, atol=1e-12 * float(V0.abs().max()) * kin)
    rr = 0.5 * r0 + 2.0 * ref[kout - 1]
    assert torch.allclose(resid, rr, rtol=0, atol=1e-12 * float(rr.abs().max()))
    assert abs(nrm2[0] - float(rr @ rr)) <= 1e-11 * float(rr @ rr)


def test_reductions_are_bit_reproducible(ab):
    """Deterministic grids and trees: the same call twice gives the same bits (no floating-point atomics)."""
    import torch
    L = ab.lib()
    n, j = 123457, 23
    ldv = n + 1
    Vfull = torch.randn(ldv * j, dtype=torch.float64, device="cuda")
    w = torch.randn(n, dtype=torch.float64, device="cuda")
    outs = []
    for _ in range(3):
        resid = torch.zeros(n, dtype=torch.float64, device="cuda")
        out = np.zeros(2 * j + 4)
        assert L.ab200_debug_orth_f64(n, j, Vfull.data_ptr(), ldv, w.data_ptr(), resid.data_ptr(), out.ctypes.data) == 0
        outs.append((out.copy(), resid.cpu().numpy()))
    for o, r in outs[1:]:
        assert np.array_equal(o, outs[0][0]) and np.array_equal(r, outs[0][1])


# --------------------------------------------------------------------------------------------------
# K3: CSR SpMV variants (CSR-bulk ring fed by cp.async.bulk, CSR-stream, row per sub-warp)
# --------------------------------------------------------------------------------------------------
def _ragged_csr(nrows, seed, heavy_block=False, empty_rows=True):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    cnt = rng.integers(0 if empty_rows else 1, 8, size=nrows)
    if heavy_block and nrows > 600:
        cnt[256:512] = 23          # one 256-row block holds far more entries than one ring slot (several chunks)
        cnt[520:900:3] = 0
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(cnt, out=rowptr[1:])
    nnz = int(rowptr[-1])
    col = rng.integers(0, nrows, size=nnz).astype(np.int32)
    val = rng.uniform(-1, 1, size=nnz)
    return sp.csr_matrix((val, col, rowptr.astype(np.int32)), shape=(nrows, nrows))


@pytest.mark.parametrize("nrows,seed,heavy", [(1, 0, False), (37, 1, False), (255, 2, False), (256, 3, False),
                                               (257, 4, False), (1000, 5, True), (4099, 6, True), (65536, 7, False),
                                               (300001, 8, True)])
def test_csr_spmv_variants_ragged(ab, nrows, seed, heavy):
    import torch
    S = _ragged_csr(nrows, seed, heavy)
    if S.nnz == 0:
        S = _ragged_csr(nrows, seed, heavy, empty_rows=False)
    # scipy's canonical form would merge duplicates; keep the raw arrays
    rowptr = torch.from_numpy(S.indptr.astype(np.int32)).cuda()
    col = torch.from_numpy(S.indices.astype(np.int32)).cuda()
    val = torch.from_numpy(S.data.copy()).cuda()
    x = torch.from_numpy(np.random.default_rng(seed + 100).uniform(-1, 1, nrows)).cuda()
    # row sums in entry order, as the reference's av would form them
    ref = np.zeros(nrows)
    xs = x.cpu().numpy()
    prod = S.data * xs[S.indices]
    for r in range(min(nrows, 5000)):
        acc = 0.0
        for p in range(S.indptr[r], S.indptr[r + 1]):
            acc += prod[p]
        ref[r] = acc
    L = ab.lib()
    out = {}
    try:
        for variant in (0, 1, 2, 4):
            L.ab200_set_spmv_variant(variant)
            y = torch.full((nrows,), float("nan"), dtype=torch.float64, device="cuda")
            assert L.ab200_csr_spmv_f64(nrows, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), x.data_ptr(),
                                        y.data_ptr()) == 0
            torch.cuda.synchronize()
            out[variant] = y.cpu().numpy()
    finally:
        L.ab200_set_spmv_variant(0)
    m = min(nrows, 5000)
    assert np.array_equal(out[0][:m], ref[:m])          # bulk: in-order row sums, bit for bit
    assert np.array_equal(out[0], out[1])               # bulk (row-owner consumers) == stream everywhere
    assert np.array_equal(out[0], out[4])               # == bulk with parked products
    full = S @ xs
    assert np.allclose(out[2], full, rtol=1e-13, atol=1e-14)
    assert np.allclose(out[0], full, rtol=1e-13, atol=1e-14)


def test_csr_spmv_bulk_unaligned_arrays_fall_back(ab):
    """cp.async.bulk needs 16-byte aligned sources: arrays that are not must still give the same product."""
    import torch
    S = _ragged_csr(3000, 11)
    n = 3000
    colbuf = torch.zeros(S.nnz + 1, dtype=torch.int32, device="cuda")
    colbuf[1:] = torch.from_numpy(S.indices.astype(np.int32)).cuda()
    col = colbuf[1:]                                     # 4-byte aligned only
    rowptr = torch.from_numpy(S.indptr.astype(np.int32)).cuda()
    val = torch.from_numpy(S.data.copy()).cuda()
    x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, n)).cuda()
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    assert col.data_ptr() % 16 != 0
    assert ab.lib().ab200_csr_spmv_f64(n, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), x.data_ptr(),
                                       y.data_ptr()) == 0
    assert np.allclose(y.cpu().numpy(), S @ x.cpu().numpy(), rtol=1e-13, atol=1e-14)


def test_csr_spmv_f32_bulk(ab):
    import torch
    S = _ragged_csr(70001, 12, heavy_block=True).astype(np.float32)
    n = S.shape[0]
    rowptr = torch.from_numpy(S.indptr.astype(np.int32)).cuda()
    col = torch.from_numpy(S.indices.astype(np.int32)).cuda()
    val = torch.from_numpy(S.data.copy()).cuda()
    x = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)).cuda()
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    assert ab.lib().ab200_csr_spmv_f32(n, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), x.data_ptr(),
                                       y.data_ptr()) == 0
    assert np.allclose(y.cpu().numpy(), S @ x.cpu().numpy(), rtol=2e-5, atol=2e-6)
