"""ctypes bindings used by the tests: the CPU parity oracle (oracle/libref_arpack.so), the host-logic test
double (tests/_build/libab200_hostdouble.so) and the product library (arpack-ng_b200/lib/libarpack_b200.so)."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libref_arpack.so")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)
ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.ref_ctx_new.restype = C.c_void_p
        L.ref_ctx_free.argtypes = [C.c_void_p]
        L.ref_ctx_set_comm.argtypes = [C.c_void_p, C.c_int, C.c_int, ALLREDUCE_FN, C.c_void_p]
        L.ref_ctx_stats.argtypes = [C.c_void_p] + [c_int_p] * 5
        for p, rp, rt in (("d", c_dbl_p, C.c_double), ("s", c_flt_p, C.c_float)):
            for fam in ("s", "n"):
                if not hasattr(L, f"ref_{p}{fam}aupd"):
                    continue
                f = getattr(L, f"ref_{p}{fam}aupd")
                f.argtypes = [C.c_void_p, c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rp, rp, C.c_int, rp,
                              C.c_int, c_int_p, c_int_p, rp, rp, C.c_int, c_int_p]
                f.restype = None
            f = getattr(L, f"ref_{p}seupd")
            f.argtypes = [C.c_void_p, C.c_int, C.c_char_p, c_int_p, rp, rp, C.c_int, rt, C.c_char_p, C.c_int,
                          C.c_char_p, C.c_int, rt, rp, C.c_int, rp, C.c_int, c_int_p, c_int_p, rp, rp, C.c_int, c_int_p]
            f.restype = None
            if not hasattr(L, f"ref_{p}neupd"):
                continue
            f = getattr(L, f"ref_{p}neupd")
            f.argtypes = [C.c_void_p, C.c_int, C.c_char_p, c_int_p, rp, rp, rp, C.c_int, rt, rt, rp, C.c_char_p,
                          C.c_int, C.c_char_p, C.c_int, rt, rp, C.c_int, rp, C.c_int, c_int_p, c_int_p, rp, rp,
                          C.c_int, c_int_p]
            f.restype = None
        for p, rp, rt in (("z", c_dbl_p, C.c_double), ("c", c_flt_p, C.c_float)):
            vp = C.c_void_p
            f = getattr(L, f"ref_{p}naupd")
            f.argtypes = [C.c_void_p, c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rp, vp, C.c_int, vp, C.c_int,
                          c_int_p, c_int_p, vp, vp, C.c_int, rp, c_int_p]
            f.restype = None
            f = getattr(L, f"ref_{p}neupd_ri")
            f.argtypes = [C.c_void_p, C.c_int, C.c_char_p, c_int_p, vp, vp, C.c_int, rt, rt, vp, C.c_char_p, C.c_int,
                          C.c_char_p, C.c_int, rt, vp, C.c_int, vp, C.c_int, c_int_p, c_int_p, vp, vp, C.c_int, rp,
                          c_int_p]
            f.restype = None
        L.ref_dlarnv2.argtypes = [c_int_p, C.c_int, c_dbl_p]
        L.ref_csr_spmv.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int]
        L.ref_set_blas_threads.argtypes = [C.c_int]
        L.ref_get_blas_threads.restype = C.c_int
        L.ref_dsaupd_csr_solve.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_dbl_p, C.c_char_p, C.c_int,
                                           C.c_int, C.c_double, C.c_int, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p,
                                           c_dbl_p, c_int_p, C.c_int, c_dbl_p, c_dbl_p]
        L.ref_dsaupd_csr_solve.restype = C.c_int
        L.ref_gen_laplace2d.argtypes = [C.c_int, C.c_int, C.c_double, c_int_p, c_int_p, c_dbl_p]
        L.ref_gen_laplace2d.restype = C.c_longlong
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


class Result(dict):
    __getattr__ = dict.__getitem__


def _make_allreduce_cb(allreduce):
    def _ar(user, buf, count, is_double, op):
        arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double if is_double else C.c_float)), (count,))
        arr[:] = allreduce(arr.copy(), op)
    return ALLREDUCE_FN(_ar)


class Oracle:
    """One 'process' of the reference: SAVE'd state (e.g. the dgetv0 seed) persists across solves."""
    prefix = "ref_"

    def __init__(self, rank=None, nranks=None, allreduce=None):
        self.L = lib()
        self.ctx = C.c_void_p(self.L.ref_ctx_new())
        self._cb = None
        if rank is not None:
            self._cb = _make_allreduce_cb(allreduce)
            self.L.ref_ctx_set_comm(self.ctx, rank, nranks, self._cb, None)

    def __del__(self):
        try:
            self.L.ref_ctx_free(self.ctx)
        except Exception:
            pass

    def stats(self, sym=True, dtype=np.float64):
        v = [C.c_int() for _ in range(5)]
        self.L.ref_ctx_stats(self.ctx, *[C.byref(x) for x in v])
        return dict(zip(("nopx", "nbx", "nrorth", "nitref", "nrstrt"), [x.value for x in v]))

    def _fn(self, name):
        return getattr(self.L, self.prefix + name)

    def _ctxargs(self, p):
        return (self.ctx,)

    def solve(self, op, n, nev, ncv, which, *, sym=True, tol=0.0, mxiter=300, bmat="I", mode=1, resid=None,
              dtype=np.float64, bop=None, rvec=True, sigma=0.0, sigmai=0.0, c_abi_tol=False, ishift=1,
              eupd=True, ldv=None, howmny="A"):
        """RCI loop exactly as EXAMPLES/SIMPLE/dssimp.f:302-324 drives it.
        op(x)->y, bop(x)->y.  c_abi_tol=True mimics the *_c entry points (tol by value, see SRC/icbads.F90)."""
        L = self.L
        dt = np.dtype(dtype)
        p = "d" if dt == np.float64 else "s"
        rp = c_dbl_p if p == "d" else c_flt_p
        rt = C.c_double if p == "d" else C.c_float
        fam = "s" if sym else "n"
        ldv = ldv or n
        lworkl = ncv * ncv + 8 * ncv if sym else 3 * ncv * ncv + 6 * ncv
        v = np.zeros((ncv, ldv), dtype=dt)  # column-major (ldv, ncv)
        workd = np.zeros(3 * n, dtype=dt)
        workl = np.zeros(lworkl, dtype=dt)
        iparam = np.zeros(11, dtype=np.int32)
        ipntr = np.zeros(14, dtype=np.int32)
        iparam[0] = ishift
        iparam[2] = mxiter
        iparam[3] = 1
        iparam[6] = mode
        info = C.c_int(0)
        if resid is None:
            res = np.zeros(n, dtype=dt)
        else:
            res = np.array(resid, dtype=dt).copy()
            info.value = 1
        ido = C.c_int(0)
        tolv = rt(tol)
        aupd = self._fn(f"{p}{fam}aupd")
        nsteps = 0
        while True:
            if c_abi_tol:
                tolv = rt(tol)
            aupd(*self._ctxargs(p), C.byref(ido), bmat.encode(), n, which.encode(), nev, C.byref(tolv), _p(res, rp), ncv,
                 _p(v, rp), ldv, _p(iparam, c_int_p), _p(ipntr, c_int_p), _p(workd, rp), _p(workl, rp), lworkl,
                 C.byref(info))
            if ido.value in (-1, 1):
                x = workd[ipntr[0] - 1: ipntr[0] - 1 + n]
                if mode == 2 and ido.value == 1 and sym:
                    y, ax = op(x)  # dsaupd mode 2: op returns (M^-1 A x, A x); x is overwritten with A x
                    workd[ipntr[0] - 1: ipntr[0] - 1 + n] = ax
                    workd[ipntr[1] - 1: ipntr[1] - 1 + n] = y
                elif mode == 5:
                    # Cayley (dsaupd.f:128-140): y = inv(A - sigma M)(A + sigma M) x; at ido=1 M x is in workd(ipntr(3))
                    bx = workd[ipntr[2] - 1: ipntr[2] - 1 + n] if ido.value == 1 else None
                    workd[ipntr[1] - 1: ipntr[1] - 1 + n] = op(x, bx)
                elif mode >= 3 and ido.value == 1 and bmat == "G":
                    workd[ipntr[1] - 1: ipntr[1] - 1 + n] = op(workd[ipntr[2] - 1: ipntr[2] - 1 + n], True)
                else:
                    r = op(x)
                    workd[ipntr[1] - 1: ipntr[1] - 1 + n] = r[0] if isinstance(r, tuple) else r
                nsteps += 1
            elif ido.value == 2:
                workd[ipntr[1] - 1: ipntr[1] - 1 + n] = bop(workd[ipntr[0] - 1: ipntr[0] - 1 + n])
            else:
                break
        out = Result(info=info.value, iparam=iparam.copy(), ipntr=ipntr.copy(), workl=workl.copy(), v=v.copy(),
                     resid=res.copy(), nconv=int(iparam[4]), stats=self.stats(sym, dt), tol_eff=tolv.value,
                     nsteps=nsteps)
        if info.value < 0 or not eupd:
            return out
        select = np.zeros(ncv, dtype=np.int32)
        ierr = C.c_int(0)
        tol_e = tol if c_abi_tol else tolv.value
        if sym:
            d = np.zeros(nev, dtype=dt)
            z = np.zeros((nev, n), dtype=dt)
            L_ = self._fn(f"{p}seupd")
            L_(*self._ctxargs(p), int(rvec), howmny.encode(), _p(select, c_int_p), _p(d, rp), _p(z, rp), n, rt(sigma), bmat.encode(), n,
               which.encode(), nev, rt(tol_e), _p(res, rp), ncv, _p(v, rp), ldv, _p(iparam, c_int_p),
               _p(ipntr, c_int_p), _p(workd, rp), _p(workl, rp), lworkl, C.byref(ierr))
            out.update(d=d, z=z, ierr=ierr.value)
        else:
            dr = np.zeros(nev + 1, dtype=dt)
            di = np.zeros(nev + 1, dtype=dt)
            z = np.zeros((max(ncv, nev + 1), n), dtype=dt)  # dneupd.f:893 treats Z as n x ncv
            workev = np.zeros(3 * ncv, dtype=dt)
            L_ = self._fn(f"{p}neupd")
            L_(*self._ctxargs(p), int(rvec), howmny.encode(), _p(select, c_int_p), _p(dr, rp), _p(di, rp), _p(z, rp), n, rt(sigma),
               rt(sigmai), _p(workev, rp), bmat.encode(), n, which.encode(), nev, rt(tol_e), _p(res, rp), ncv,
               _p(v, rp), ldv, _p(iparam, c_int_p), _p(ipntr, c_int_p), _p(workd, rp), _p(workl, rp), lworkl,
               C.byref(ierr))
            out.update(dr=dr, di=di, z=z, ierr=ierr.value)
        out.update(workl_eupd=workl.copy(), v_eupd=v.copy(), ipntr_eupd=ipntr.copy(), select=select.copy())
        return out


def _solve_complex(self, op, n, nev, ncv, which, *, tol=0.0, mxiter=300, bmat="I", mode=1, resid=None,
                   dtype=np.complex128, bop=None, rvec=True, sigma=0.0, c_abi_tol=False, ishift=1, eupd=True, ldv=None,
                   shifts=None, howmny="A"):
    """RCI loop around znaupd/zneupd (cnaupd/cneupd for complex64) as EXAMPLES/COMPLEX/zndrv1.f drives it.
    op(x)->y, bop(x)->y; mode 3 with bmat='G': op(x, bx) receives workd(ipntr(3)) = B x as second argument."""
    dt = np.dtype(dtype)
    p = "z" if dt == np.complex128 else "c"
    rdt = np.float64 if p == "z" else np.float32
    rp = c_dbl_p if p == "z" else c_flt_p
    rt = C.c_double if p == "z" else C.c_float
    ldv = ldv or n
    lworkl = 3 * ncv * ncv + 5 * ncv
    v = np.zeros((ncv, ldv), dtype=dt)
    workd = np.zeros(3 * n, dtype=dt)
    workl = np.zeros(lworkl, dtype=dt)
    rwork = np.zeros(ncv, dtype=rdt)
    iparam = np.zeros(11, dtype=np.int32)
    ipntr = np.zeros(14, dtype=np.int32)
    iparam[0], iparam[2], iparam[3], iparam[6] = ishift, mxiter, 1, mode
    info = C.c_int(0)
    if resid is None:
        res = np.zeros(n, dtype=dt)
    else:
        res = np.array(resid, dtype=dt).copy()
        info.value = 1
    ido = C.c_int(0)
    tolv = rt(tol)
    aupd = self._fn(f"{p}naupd")
    nsteps = 0
    nshift_calls = 0
    vp = lambda a: a.ctypes.data  # noqa: E731
    while True:
        if c_abi_tol:
            tolv = rt(tol)
        aupd(*self._ctxargs(p), C.byref(ido), bmat.encode(), n, which.encode(), nev, C.byref(tolv), vp(res), ncv, vp(v),
             ldv, _p(iparam, c_int_p), _p(ipntr, c_int_p), vp(workd), vp(workl), lworkl, _p(rwork, rp), C.byref(info))
        if ido.value in (-1, 1):
            x = workd[ipntr[0] - 1: ipntr[0] - 1 + n]
            if mode == 3 and bmat == "G":
                bx = workd[ipntr[2] - 1: ipntr[2] - 1 + n] if ido.value == 1 else None
                workd[ipntr[1] - 1: ipntr[1] - 1 + n] = op(x, bx)
            else:
                workd[ipntr[1] - 1: ipntr[1] - 1 + n] = op(x)
            nsteps += 1
        elif ido.value == 2:
            workd[ipntr[1] - 1: ipntr[1] - 1 + n] = bop(workd[ipntr[0] - 1: ipntr[0] - 1 + n])
        elif ido.value == 3 and shifts is not None:
            # ishift = 0 (znaupd.f:168-178): iparam(8) shifts go to workl(ipntr(14)); the Ritz values of H are in
            # workl(ipntr(6)), their estimates in workl(ipntr(8))
            npsh = int(iparam[7])
            workl[ipntr[13] - 1: ipntr[13] - 1 + npsh] = shifts(workl[ipntr[5] - 1: ipntr[5] - 1 + ncv].copy(),
                                                                workl[ipntr[7] - 1: ipntr[7] - 1 + ncv].copy(), npsh)
            nshift_calls += 1
        else:
            break
    out = Result(info=info.value, iparam=iparam.copy(), ipntr=ipntr.copy(), workl=workl.copy(), v=v.copy(),
                 resid=res.copy(), nconv=int(iparam[4]), tol_eff=tolv.value, nsteps=nsteps,
                 nshift_calls=nshift_calls)
    if info.value < 0 or not eupd:
        return out
    select = np.zeros(ncv, dtype=np.int32)
    ierr = C.c_int(0)
    tol_e = tol if c_abi_tol else tolv.value
    d = np.zeros(nev + 1, dtype=dt)
    z = np.zeros((nev, n), dtype=dt)
    workev = np.zeros(2 * ncv, dtype=dt)
    sg = complex(sigma)
    self._fn(f"{p}neupd_ri")(*self._ctxargs(p), int(rvec), howmny.encode(), _p(select, c_int_p), vp(d), vp(z), n, rt(sg.real),
                             rt(sg.imag), vp(workev), bmat.encode(), n, which.encode(), nev, rt(tol_e), vp(res), ncv,
                             vp(v), ldv, _p(iparam, c_int_p), _p(ipntr, c_int_p), vp(workd), vp(workl), lworkl,
                             _p(rwork, rp), C.byref(ierr))
    out.update(d=d[:nev], z=z, ierr=ierr.value, workl_eupd=workl.copy(), v_eupd=v.copy(), ipntr_eupd=ipntr.copy(),
               select=select.copy())
    return out


Oracle.solve_complex = _solve_complex


# ----------------------------------------------------------------------------------------------------
# host-logic test double: the product's host control code over a plain-loop VecOps (no GPU, no oracle)
# ----------------------------------------------------------------------------------------------------
_HD_SO = os.path.join(_ROOT, "tests", "_build", "libab200_hostdouble.so")
_hd = None


def build_hostdouble():
    import glob
    import scipy
    os.makedirs(os.path.dirname(_HD_SO), exist_ok=True)
    src = os.path.join(_ROOT, "tests", "hostdouble", "hostdouble.cpp")
    src_z = os.path.join(_ROOT, "tests", "hostdouble", "hostdouble_cplx.cpp")
    deps = [src, src_z] + glob.glob(os.path.join(_ROOT, "arpack-ng_b200", "csrc", "*.hpp"))
    if os.path.exists(_HD_SO) and all(os.path.getmtime(_HD_SO) >= os.path.getmtime(d) for d in deps):
        return
    blas = os.path.abspath(glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                                  "libscipy_openblas*.so"))[0])
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _HD_SO, src, src_z, blas,
                           "-Wl,-rpath," + os.path.dirname(blas)])


def hostdouble_lib():
    global _hd
    if _hd is None:
        build_hostdouble()
        L = C.CDLL(_HD_SO)
        L.hd_new.restype = C.c_void_p
        L.hd_new.argtypes = [C.c_int]
        L.hd_free.argtypes = [C.c_void_p, C.c_int]
        L.hd_set_comm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ALLREDUCE_FN]
        L.hd_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, c_int_p]
        L.hd_set_registered_op.argtypes = [C.c_void_p, C.c_int, OP_FN, C.c_int, C.c_int]
        L.hd_fused_dot_maxdiff.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hd_fused_dot_maxdiff.restype = C.c_double
        L.hd_deferred_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        L.hd_set_deferral.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for p, rp, rt in (("d", c_dbl_p, C.c_double), ("s", c_flt_p, C.c_float)):
            for fam in ("s", "n"):
                f = getattr(L, f"hd_{p}{fam}aupd")
                f.argtypes = [C.c_void_p, c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rp, rp, C.c_int, rp,
                              C.c_int, c_int_p, c_int_p, rp, rp, C.c_int, c_int_p]
                f.restype = None
            f = getattr(L, f"hd_{p}seupd")
            f.argtypes = [C.c_void_p, C.c_int, C.c_char_p, c_int_p, rp, rp, C.c_int, rt, C.c_char_p, C.c_int,
                          C.c_char_p, C.c_int, rt, rp, C.c_int, rp, C.c_int, c_int_p, c_int_p, rp, rp, C.c_int, c_int_p]
            f.restype = None
            f = getattr(L, f"hd_{p}neupd")
            f.argtypes = [C.c_void_p, C.c_int, C.c_char_p, c_int_p, rp, rp, rp, C.c_int, rt, rt, rp, C.c_char_p,
                          C.c_int, C.c_char_p, C.c_int, rt, rp, C.c_int, rp, C.c_int, c_int_p, c_int_p, rp, rp,
                          C.c_int, c_int_p]
            f.restype = None
        L.hd_debug.argtypes = [c_int_p]
        L.hdz_new.restype = C.c_void_p
        L.hdz_new.argtypes = [C.c_int]
        L.hdz_free.argtypes = [C.c_void_p, C.c_int]
        L.hdz_stats.argtypes = [C.c_void_p, C.c_int, c_int_p]
        L.hdz_set_comm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ALLREDUCE_FN]
        for p, rp, rt in (("z", c_dbl_p, C.c_double), ("c", c_flt_p, C.c_float)):
            vp = C.c_void_p
            f = getattr(L, f"hd_{p}naupd")
            f.argtypes = [C.c_void_p, c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rp, vp, C.c_int, vp, C.c_int,
                          c_int_p, c_int_p, vp, vp, C.c_int, rp, c_int_p]
            f.restype = None
            f = getattr(L, f"hd_{p}neupd_ri")
            f.argtypes = [C.c_void_p, C.c_int, C.c_char_p, c_int_p, vp, vp, C.c_int, rt, rt, vp, C.c_char_p, C.c_int,
                          C.c_char_p, C.c_int, rt, vp, C.c_int, vp, C.c_int, c_int_p, c_int_p, vp, vp, C.c_int, rp,
                          c_int_p]
            f.restype = None
        _hd = L
    return _hd


OP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)


class HostDouble(Oracle):
    """The product's IrlSym/IrlNonsym driven through the plain-loop VecOps test double."""
    prefix = "hd_"

    def __init__(self, rank=None, nranks=None, allreduce=None):
        self.L = hostdouble_lib()
        self._procs = {True: C.c_void_p(self.L.hd_new(1)), False: C.c_void_p(self.L.hd_new(0))}
        self._procs_z = {True: C.c_void_p(self.L.hdz_new(1)), False: C.c_void_p(self.L.hdz_new(0))}
        self._cb = None
        if rank is not None:
            self._cb = _make_allreduce_cb(allreduce)
            for isd, pr in self._procs.items():
                self.L.hd_set_comm(pr, int(isd), rank, nranks, self._cb)
            for isd, pr in self._procs_z.items():
                self.L.hdz_set_comm(pr, int(isd), rank, nranks, self._cb)

    def __del__(self):
        try:
            for isd, pr in self._procs.items():
                self.L.hd_free(pr, int(isd))
            for isd, pr in self._procs_z.items():
                self.L.hdz_free(pr, int(isd))
        except Exception:
            pass

    def _ctxargs(self, p):
        if p in ("z", "c"):
            return (self._procs_z[p == "z"],)
        return (self._procs[p == "d"],)

    def register_op(self, op, n, dtype=np.float64, fused=False):
        """Registered-operator mode of the product's control code: OP is applied inside *aupd (no ido=+-1)."""
        dt = np.dtype(dtype)
        isd = dt == np.float64

        def cb(xp, yp, nn):
            ct = C.c_double if isd else C.c_float
            x = np.ctypeslib.as_array(C.cast(xp, C.POINTER(ct)), shape=(nn,))
            y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(ct)), shape=(nn,))
            y[:] = op(x.copy())
        self._opcb = OP_FN(cb) if op is not None else C.cast(None, OP_FN)
        self.L.hd_set_registered_op(self._procs[isd], int(isd), self._opcb, n, int(fused))

    def deferred_stats(self, sym=True, dtype=np.float64):
        """(steps run inside device-resident batches, batches cut short by a rare path, host round trips)."""
        isd = np.dtype(dtype) == np.float64
        out = (C.c_longlong * 3)()
        self.L.hd_deferred_stats(self._procs[isd], int(isd), int(sym), out)
        return int(out[0]), int(out[1]), int(out[2])

    def set_deferral(self, on, dtype=np.float64):
        isd = np.dtype(dtype) == np.float64
        self.L.hd_set_deferral(self._procs[isd], int(isd), int(on))

    def fused_dot_maxdiff(self, sym=True, dtype=np.float64):
        isd = np.dtype(dtype) == np.float64
        return self.L.hd_fused_dot_maxdiff(self._procs[isd], int(isd), int(sym))

    def stats(self, sym=True, dtype=np.float64):
        out = np.zeros(5, dtype=np.int32)
        isd = np.dtype(dtype) == np.float64
        self.L.hd_stats(self._procs[isd], int(isd), int(sym), _p(out, c_int_p))
        return dict(zip(("nopx", "nbx", "nrorth", "nitref", "nrstrt"), [int(x) for x in out]))
