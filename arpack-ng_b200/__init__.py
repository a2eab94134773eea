"""arpack-ng_b200 -- Python host-side mirror of the reference's reverse-communication interface.

The product is the C-ABI shared library ``arpack-ng_b200/lib/libarpack_b200.so`` (sm_100a CUDA kernels +
C++ host control; see ``include/arpack_b200.h``).  This module is the thin ctypes binding a Python caller
uses, with the same entry-point names and argument meaning as ``ICB/arpack.h`` / ``ICB/parpack.h`` of
arpack-ng (``dsaupd_c``, ``dseupd_c``, ``dnaupd_c``, ``dneupd_c``, ``pdsaupd_c`` ...), plus an
``arpackmm``-style convenience driver (``eigsh_csr``; EXAMPLES/MATRIX_MARKET/arpackSolver.hpp:721-880).

PyTorch is used only for device memory, streams and (multi-GPU) torch.distributed plumbing.
There is no CPU fallback: if the library is missing or no CUDA device is usable, calls raise.

The directory name contains a hyphen; import it through the ``arpack_ng_b200`` shim at the repo root.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libarpack_b200.so")

c_int_p = C.POINTER(C.c_int)
_lib = None

INFO_DEVICE_ERROR = -9990


class ArpackB200Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile libarpack_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    import subprocess
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise ArpackB200Error("building libarpack_b200.so failed")
    return LIB_PATH


def _sig_aupd(rp, rt, par):
    a = [c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rt, rp, C.c_int, rp, C.c_int, c_int_p, c_int_p, rp, rp,
         C.c_int, c_int_p]
    return ([C.c_int] if par else []) + a


def _sig_seupd(rp, rt, par):
    a = [C.c_int, C.c_char_p, c_int_p, rp, rp, C.c_int, rt, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rt, rp, C.c_int,
         rp, C.c_int, c_int_p, c_int_p, rp, rp, C.c_int, c_int_p]
    return ([C.c_int] if par else []) + a


def _sig_neupd(rp, rt, par):
    a = [C.c_int, C.c_char_p, c_int_p, rp, rp, rp, C.c_int, rt, rt, rp, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rt,
         rp, C.c_int, rp, C.c_int, c_int_p, c_int_p, rp, rp, C.c_int, c_int_p]
    return ([C.c_int] if par else []) + a


def lib():
    """Load the C-ABI library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ArpackB200Error(f"{LIB_PATH} not found: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU path.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p  # array arguments are passed as raw addresses (host or device)
    for p, rt in (("d", C.c_double), ("s", C.c_float)):
        for par in (False, True):
            pre = "p" if par else ""
            for fam in ("s", "n"):
                f = getattr(L, f"{pre}{p}{fam}aupd_c")
                f.argtypes = _sig_aupd(vp, rt, par)
                f.restype = None
            f = getattr(L, f"{pre}{p}seupd_c")
            f.argtypes = _sig_seupd(vp, rt, par)
            f.restype = None
            f = getattr(L, f"{pre}{p}neupd_c")
            f.argtypes = _sig_neupd(vp, rt, par)
            f.restype = None
    for p, rt in (("z", C.c_double), ("c", C.c_float)):
        rp = C.POINTER(rt)
        f = getattr(L, f"{p}naupd_c")   # ICB/arpack.h:10,20
        f.argtypes = [c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rt, vp, C.c_int, vp, C.c_int, c_int_p, c_int_p,
                      vp, vp, C.c_int, rp, c_int_p]
        f.restype = None
        f = getattr(L, f"ab200_{p}neupd_ri")   # [cz]neupd_c with sigma as two reals (ctypes has no C complex)
        f.argtypes = [C.c_int, C.c_char_p, c_int_p, vp, vp, C.c_int, rt, rt, vp, C.c_char_p, C.c_int, C.c_char_p,
                      C.c_int, rt, vp, C.c_int, vp, C.c_int, c_int_p, c_int_p, vp, vp, C.c_int, rp, c_int_p]
        f.restype = None
        f = getattr(L, f"p{p}naupd_c")  # ICB/parpack.h:28,32
        f.argtypes = [C.c_int, c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, rt, vp, C.c_int, vp, C.c_int, c_int_p,
                      c_int_p, vp, vp, C.c_int, rp, c_int_p]
        f.restype = None
    L.ab200_pzneupd_ri.argtypes = [C.c_int, C.c_int, C.c_char_p, c_int_p, vp, vp, C.c_int, C.c_double, C.c_double, vp,
                                   C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, vp, C.c_int, vp, C.c_int,
                                   c_int_p, c_int_p, vp, vp, C.c_int, C.POINTER(C.c_double), c_int_p]
    L.ab200_pzneupd_ri.restype = None
    L.ab200_set_stream.argtypes = [vp]
    L.ab200_get_stream.restype = vp
    L.ab200_set_kernel_mode.argtypes = [C.c_int]
    L.ab200_set_compat.argtypes = [C.c_int]
    L.ab200_release.argtypes = [vp]
    L.ab200_launch_stats.argtypes = [C.POINTER(C.c_ulonglong)]
    L.ab200_device_count.restype = C.c_int
    L.ab200_host_round_trips.restype = C.c_ulonglong
    L.ab200_register_csr_op_f64.argtypes = [vp, C.c_int, C.c_longlong, vp, vp, vp]
    L.ab200_register_csr_op_f32.argtypes = [vp, C.c_int, C.c_longlong, vp, vp, vp]
    L.ab200_register_csr_halo_op_f64.argtypes = [vp, C.c_int, C.c_int, C.c_longlong, vp, vp, vp, C.c_int, C.c_int, vp]
    L.ab200_fused_dot_maxdiff.argtypes = [vp]
    L.ab200_fused_dot_maxdiff.restype = C.c_double
    L.ab200_kernel_probe_f64.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ab200_debug_orth_f64.argtypes = [C.c_longlong, C.c_int, vp, C.c_longlong, vp, vp, vp]
    L.ab200_debug_vq_f64.argtypes = [C.c_longlong, C.c_int, C.c_int, vp, C.c_longlong, vp, C.c_double, C.c_double,
                                     C.c_int, vp, vp]
    L.ab200_debug_zorth_f64.argtypes = [C.c_longlong, C.c_int, vp, C.c_longlong, vp, vp, vp]
    L.ab200_debug_zvq_f64.argtypes = [C.c_longlong, C.c_int, C.c_int, vp, C.c_longlong, vp, vp, C.c_longlong, C.c_double,
                                      C.c_double, C.c_double, C.c_double, C.c_int, vp, C.POINTER(C.c_double)]
    L.ab200_profile_enable.argtypes = [C.c_int]
    L.ab200_profile_get.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong),
                                    C.POINTER(C.c_double)]
    L.ab200_profile_get.restype = C.c_int
    L.ab200_version.restype = C.c_char_p
    L.ab200_nccl_unique_id.argtypes = [vp]
    L.ab200_comm_create.argtypes = [vp, C.c_int, C.c_int]
    L.ab200_comm_destroy.argtypes = [C.c_int]
    L.ab200_comm_halo_buffer.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_int]
    L.ab200_comm_halo_buffer.restype = vp
    L.ab200_csr_spmv_f64.argtypes = [C.c_int, vp, vp, vp, vp, vp]
    L.ab200_csr_spmv_f32.argtypes = [C.c_int, vp, vp, vp, vp, vp]
    L.ab200_csr_spmv_hostvec_f64.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.ab200_csr_spmv_halo_f64.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    L.ab200_gen_laplace2d.argtypes = [C.c_int, C.c_int, C.c_double, vp, vp, vp]
    L.ab200_gen_laplace2d.restype = C.c_longlong
    L.ab200_gen_laplace3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, vp, vp, vp]
    L.ab200_gen_laplace3d.restype = C.c_longlong
    L.ab200_gen_convdiff2d.argtypes = [C.c_int, C.c_double, vp, vp, vp]
    L.ab200_gen_convdiff2d.restype = C.c_longlong
    L.ab200_fill_hash_f64.argtypes = [C.c_longlong, C.c_longlong, C.c_ulonglong, vp]
    L.ab200_gen_randsparse.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, vp, vp, vp]
    L.ab200_gen_randsparse.restype = C.c_longlong
    L.ab200_csr_transpose_f64.argtypes = [C.c_int, C.c_int, C.c_longlong, vp, vp, vp, vp, vp, vp]
    L.ab200_gram_create.argtypes = [C.c_int, C.c_int]
    L.ab200_gram_add_shard.argtypes = [C.c_int, C.c_int, C.c_longlong, vp, vp, vp, vp, vp, vp]
    L.ab200_gram_apply.argtypes = [C.c_int, vp, vp]
    L.ab200_gram_destroy.argtypes = [C.c_int]
    L.ab200_register_gram_op_f64.argtypes = [vp, C.c_int]
    L.ab200_residuals_f64.argtypes = [C.c_int, vp, vp, vp, C.c_int, vp, C.c_longlong, vp, vp]
    _lib = L
    return L


def launch_stats():
    out = (C.c_ulonglong * 4)()
    lib().ab200_launch_stats(out)
    return {"kernels": int(out[0]), "allreduces": int(out[1]), "tma_path": int(out[2]), "generic_path": int(out[3])}


def host_round_trips():
    """Blocking device->host mailbox reads since the library was loaded."""
    return int(lib().ab200_host_round_trips())


def profile(enable=None, reset=False):
    """Per-kernel CUDA-event timings accumulated by the library: {name: {launches, ms, bytes}}."""
    L = lib()
    if reset:
        L.ab200_profile_reset()
    if enable is not None:
        L.ab200_profile_enable(int(enable))
    out = {}
    name = C.create_string_buffer(64)
    ms, ln, by = C.c_double(), C.c_ulonglong(), C.c_double()
    cnt = L.ab200_profile_get(-1, name, C.byref(ms), C.byref(ln), C.byref(by))
    for i in range(cnt):
        L.ab200_profile_get(i, name, C.byref(ms), C.byref(ln), C.byref(by))
        out[name.value.decode()] = {"launches": int(ln.value), "ms": float(ms.value), "bytes": float(by.value)}
    return out


def _addr(a):
    """Raw address of a numpy array (host) or a torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _real(dtype):
    dt = np.dtype(dtype) if not hasattr(dtype, "is_floating_point") else None
    if dt is None:
        import torch
        return ("d", C.c_double) if dtype == torch.float64 else ("s", C.c_float)
    return ("d", C.c_double) if dt == np.float64 else ("s", C.c_float)


def _prec_of(arr):
    if isinstance(arr, np.ndarray):
        return ("d", C.c_double) if arr.dtype == np.float64 else ("s", C.c_float)
    return _real(arr.dtype)


# ---- the reference's entry points (ICB/arpack.h); ido/info are 1-element int32 numpy arrays -------
def _call_aupd(name, comm, ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, info):
    p, rt = _prec_of(workl)
    f = getattr(lib(), f"{'p' if comm is not None else ''}{p}{name}_c")
    args = [ido.ctypes.data_as(c_int_p), bmat.encode(), n, which.encode(), nev, rt(tol), _addr(resid), ncv, _addr(v),
            ldv, iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), _addr(workd), _addr(workl),
            int(workl.size), info.ctypes.data_as(c_int_p)]
    if comm is not None:
        args = [comm] + args
    f(*args)
    if info[0] == INFO_DEVICE_ERROR:
        raise ArpackB200Error(f"{name}_c: CUDA device error (see stderr); there is no CPU fallback")


def dsaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, info, comm=None):
    """SRC/icbads.F90:3-40 (dsaupd_c) / ssaupd_c by dtype of workl; comm != None -> pdsaupd_c."""
    _call_aupd("saupd", comm, ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, info)


def dnaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, info, comm=None):
    """SRC/icbadn.F90 (dnaupd_c) / snaupd_c by dtype of workl; comm != None -> pdnaupd_c."""
    _call_aupd("naupd", comm, ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, info)


def dseupd_c(rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr,
             workd, workl, info, comm=None):
    """SRC/icbads.F90:42-92 (dseupd_c)."""
    p, rt = _prec_of(workl)
    f = getattr(lib(), f"{'p' if comm is not None else ''}{p}seupd_c")
    args = [int(rvec), howmny.encode(), select.ctypes.data_as(c_int_p), _addr(d), _addr(z), ldz, rt(sigma),
            bmat.encode(), n, which.encode(), nev, rt(tol), _addr(resid), ncv, _addr(v), ldv,
            iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), _addr(workd), _addr(workl), int(workl.size),
            info.ctypes.data_as(c_int_p)]
    if comm is not None:
        args = [comm] + args
    f(*args)
    if info[0] == INFO_DEVICE_ERROR:
        raise ArpackB200Error("seupd_c: CUDA device error (see stderr)")


def dneupd_c(rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which, nev, tol, resid, ncv, v,
             ldv, iparam, ipntr, workd, workl, info, comm=None):
    """SRC/icbadn.F90 (dneupd_c)."""
    p, rt = _prec_of(workl)
    f = getattr(lib(), f"{'p' if comm is not None else ''}{p}neupd_c")
    args = [int(rvec), howmny.encode(), select.ctypes.data_as(c_int_p), _addr(dr), _addr(di), _addr(z), ldz,
            rt(sigmar), rt(sigmai), _addr(workev), bmat.encode(), n, which.encode(), nev, rt(tol), _addr(resid), ncv,
            _addr(v), ldv, iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), _addr(workd), _addr(workl),
            int(workl.size), info.ctypes.data_as(c_int_p)]
    if comm is not None:
        args = [comm] + args
    f(*args)
    if info[0] == INFO_DEVICE_ERROR:
        raise ArpackB200Error("neupd_c: CUDA device error (see stderr)")


# ---- RCI driver in the style of arpackSolver.hpp:721-880 -------------------------------------------
class Result(dict):
    __getattr__ = dict.__getitem__


def alloc_host_buffers(n, ncv, dtype=np.float64, pinned=True):
    """(V, workd, resid) as host tensors of the sizes a reference caller declares; pinned by default."""
    import torch
    t_dt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32

    def _h(cnt):
        t = torch.zeros(cnt, dtype=t_dt)
        return t.pin_memory() if pinned else t
    return _h(n * ncv), _h(3 * n), _h(n)


def alloc_device_buffers(n, ncv, dtype=np.float64, device="cuda"):
    """(V, workd, resid) in HBM, the device-resident twin of alloc_host_buffers."""
    import torch
    t_dt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    return (torch.zeros(n * ncv, dtype=t_dt, device=device), torch.zeros(3 * n, dtype=t_dt, device=device),
            torch.zeros(n, dtype=t_dt, device=device))


def solve(op, n, nev, ncv, which, *, sym=True, tol=0.0, mxiter=300, bmat="I", mode=1, resid=None, dtype=np.float64,
          bop=None, rvec=True, sigma=0.0, sigmai=0.0, device="cuda", host_buffers=False, comm=None, ishift=1,
          eupd=True, pinned=True, registered_op=None, buffers=None, howmny="A"):
    """Run a whole *aupd/*eupd solve through the C-ABI.

    device arrays (default): resid/v/workd are torch CUDA tensors; ``op(x, y)`` receives tensor views of the
    workd slots and must fill y on the library's stream (torch's current stream == legacy default stream).
    host_buffers=True: resid/v/workd are (pinned) host arrays exactly as an unmodified reference caller would
    own them; ``op(x, y)`` then receives numpy views.
    """
    import torch
    L = lib()
    np_dt = np.dtype(dtype)
    t_dt = torch.float64 if np_dt == np.float64 else torch.float32
    lworkl = ncv * ncv + 8 * ncv if sym else 3 * ncv * ncv + 6 * ncv
    workl = np.zeros(lworkl, dtype=np_dt)
    iparam = np.zeros(11, dtype=np.int32)
    ipntr = np.zeros(14, dtype=np.int32)
    iparam[0], iparam[2], iparam[3], iparam[6] = ishift, mxiter, 1, mode
    ido = np.zeros(1, dtype=np.int32)
    info = np.zeros(1, dtype=np.int32)
    ldv = n
    if host_buffers:
        # the caller's own arrays, as dssimp.f:213-219 declares them; reusable across solves via `buffers`
        v_t, workd_t, resid_t = buffers if buffers is not None else alloc_host_buffers(n, ncv, np_dt, pinned)
        v, workd, res = v_t.numpy(), workd_t.numpy(), resid_t.numpy()
        if resid is not None:
            res[:] = np.asarray(resid, dtype=np_dt)
            info[0] = 1
    elif buffers is not None:
        v, workd, res = buffers          # the caller's device arrays, reused across solves
        if resid is not None:
            res.copy_(torch.as_tensor(resid, dtype=t_dt))
            info[0] = 1
    else:
        v, workd, res = alloc_device_buffers(n, ncv, np_dt, device)
        if resid is not None:
            res.copy_(torch.as_tensor(resid, dtype=t_dt))
            info[0] = 1
    if isinstance(registered_op, GramOperator):
        # opt-in extension: the library applies the A^T A operator itself (one *aupd call per solve)
        if L.ab200_register_gram_op_f64(workl.ctypes.data, registered_op.h) != 0:
            raise ArpackB200Error("register_gram_op failed")
        if op is None:
            op = registered_op
    elif registered_op is not None:
        # opt-in extension: the library applies the CSR operator itself (one *aupd call per solve)
        if registered_op.val.dtype not in (t_dt, np_dt):
            raise ArpackB200Error("registered_op values must have the solve's dtype")
        if comm is not None:
            # row-partitioned operator: the library also runs the neighbour exchange of the halo planes
            if np_dt != np.float64:
                raise ArpackB200Error("halo registration is FP64 only")
            if not hasattr(registered_op, "halo"):
                import torch as _t
                registered_op.halo_lo = registered_op.halo_hi = 0
                registered_op.halo = _t.zeros(1, dtype=t_dt, device=device)
            rc = L.ab200_register_csr_halo_op_f64(workl.ctypes.data, comm, registered_op.n, registered_op.nnz,
                                                  registered_op.rowptr.data_ptr(), registered_op.col.data_ptr(),
                                                  registered_op.val.data_ptr(), registered_op.halo_lo,
                                                  registered_op.halo_hi, registered_op.halo_address(comm))
        else:
            reg = L.ab200_register_csr_op_f64 if np_dt == np.float64 else L.ab200_register_csr_op_f32
            # device tensors are used in place; host arrays (HostCsr) are uploaded once by the library at ido = 0
            rc = reg(workl.ctypes.data, registered_op.n, registered_op.nnz, _addr(registered_op.rowptr),
                     _addr(registered_op.col), _addr(registered_op.val))
        if rc != 0:
            raise ArpackB200Error("register_csr_op failed")
        if op is None:
            op = registered_op
    # the argument list is the same on every re-entry (the reference's loop passes the same variables each time,
    # dssimp.f:302-311): marshal it once so that the per-step host cost is one foreign call
    p_, rt = _prec_of(workl)
    f_aupd = getattr(L, f"{'p' if comm is not None else ''}{p_}{'saupd' if sym else 'naupd'}_c")
    cargs = [ido.ctypes.data_as(c_int_p), bmat.encode(), n, which.encode(), nev, rt(tol), _addr(res), ncv, _addr(v),
             ldv, iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), _addr(workd), _addr(workl),
             int(workl.size), info.ctypes.data_as(c_int_p)]
    if comm is not None:
        cargs = [comm] + cargs
    cargs = tuple(cargs)
    # operators that accept raw addresses (CsrOperator on device arrays) skip the per-step tensor views
    fast = getattr(op, "apply_ptr", None) if (not host_buffers and not (mode >= 3 and bmat == "G")) else None
    if fast is not None and comm is not None:
        fast = getattr(op, "apply_comm_ptr", None) or (op.apply_halo_ptr if hasattr(op, "halo_lo") else None)
    wbase, isz = _addr(workd), np_dt.itemsize
    nsteps = 0
    while True:
        f_aupd(*cargs)
        if info[0] == INFO_DEVICE_ERROR:
            raise ArpackB200Error("aupd_c: CUDA device error (see stderr); there is no CPU fallback")
        i = ido[0]
        if i == 1 or i == -1:
            if fast is not None:
                if comm is not None:
                    fast(comm, wbase + (int(ipntr[0]) - 1) * isz, wbase + (int(ipntr[1]) - 1) * isz)
                else:
                    fast(wbase + (int(ipntr[0]) - 1) * isz, wbase + (int(ipntr[1]) - 1) * isz)
            else:
                x = workd[ipntr[0] - 1: ipntr[0] - 1 + n]
                y = workd[ipntr[1] - 1: ipntr[1] - 1 + n]
                if mode == 5:   # Cayley: the operator needs x and, at ido = 1, the M x the library provides
                    op(x, y, workd[ipntr[2] - 1: ipntr[2] - 1 + n] if i == 1 else None)
                elif mode >= 3 and bmat == "G" and i == 1:
                    op(workd[ipntr[2] - 1: ipntr[2] - 1 + n], y, True)
                else:
                    op(x, y)
            nsteps += 1
        elif i == 2:
            bop(workd[ipntr[0] - 1: ipntr[0] - 1 + n], workd[ipntr[1] - 1: ipntr[1] - 1 + n])
        else:
            break
    out = Result(info=int(info[0]), iparam=iparam.copy(), ipntr=ipntr.copy(), workl=workl.copy(), v=v, resid=res,
                 nconv=int(iparam[4]), nsteps=nsteps, workd=workd,
                 fused_dot_maxdiff=(L.ab200_fused_dot_maxdiff(workl.ctypes.data)
                                    if registered_op is not None and not isinstance(registered_op, GramOperator)
                                    else None))
    if info[0] < 0 or not eupd:
        L.ab200_release(workl.ctypes.data)   # no *eupd will follow: drop the solve's context (and its HBM mirrors)
        return out
    select = np.zeros(ncv, dtype=np.int32)
    ierr = np.zeros(1, dtype=np.int32)
    if sym:
        d = np.zeros(nev, dtype=np_dt)
        dseupd_c(rvec, howmny, select, d, v, ldv, sigma, bmat, n, which, nev, tol, res, ncv, v, ldv, iparam, ipntr, workd,
                 workl, ierr, comm=comm)
        out.update(d=d, z=v, ierr=int(ierr[0]))
    else:
        dr = np.zeros(nev + 1, dtype=np_dt)
        di = np.zeros(nev + 1, dtype=np_dt)
        workev = np.zeros(3 * ncv, dtype=np_dt)
        dneupd_c(rvec, howmny, select, dr, di, v, ldv, sigma, sigmai, workev, bmat, n, which, nev, tol, res, ncv, v, ldv,
                 iparam, ipntr, workd, workl, ierr, comm=comm)
        out.update(dr=dr, di=di, z=v, ierr=int(ierr[0]))
    out.update(workl_eupd=workl.copy(), ipntr_eupd=ipntr.copy())
    L.ab200_release(workl.ctypes.data)
    return out


def solve_complex(op, n, nev, ncv, which, *, tol=0.0, mxiter=300, bmat="I", mode=1, resid=None, dtype=np.complex128,
                  bop=None, rvec=True, sigma=0.0, device="cuda", host_buffers=False, ishift=1, eupd=True, comm=None,
                  howmny="A"):
    """znaupd_c/zneupd_c (cnaupd_c/cneupd_c for complex64) solve through the C-ABI, the loop of
    EXAMPLES/COMPLEX/zndrv1.f.  Device arrays (default): ``op(x, y)`` gets complex CUDA tensor views of the workd slots;
    host_buffers=True: numpy views.  Mode 3 with bmat='G': ``op(x, y, bx)`` also receives workd(ipntr(3)) at ido = 1."""
    import torch
    L = lib()
    np_dt = np.dtype(dtype)
    p_, rt = ("z", C.c_double) if np_dt == np.complex128 else ("c", C.c_float)
    r_dt = np.float64 if p_ == "z" else np.float32
    t_dt = torch.complex128 if p_ == "z" else torch.complex64
    lworkl = 3 * ncv * ncv + 5 * ncv
    workl = np.zeros(lworkl, dtype=np_dt)
    rwork = np.zeros(ncv, dtype=r_dt)
    iparam = np.zeros(11, dtype=np.int32)
    ipntr = np.zeros(14, dtype=np.int32)
    iparam[0], iparam[2], iparam[3], iparam[6] = ishift, mxiter, 1, mode
    ido = np.zeros(1, dtype=np.int32)
    info = np.zeros(1, dtype=np.int32)
    ldv = n
    if host_buffers:
        v, workd, res = np.zeros(n * ncv, dtype=np_dt), np.zeros(3 * n, dtype=np_dt), np.zeros(n, dtype=np_dt)
        if resid is not None:
            res[:] = np.asarray(resid, dtype=np_dt)
            info[0] = 1
    else:
        v = torch.zeros(n * ncv, dtype=t_dt, device=device)
        workd = torch.zeros(3 * n, dtype=t_dt, device=device)
        res = torch.zeros(n, dtype=t_dt, device=device)
        if resid is not None:
            res.copy_(torch.as_tensor(np.asarray(resid, dtype=np_dt)))
            info[0] = 1
    # comm != None -> the PARPACK twins pznaupd_c / pzneupd_c (n = local rows, every rank calls in lock-step)
    f_aupd = getattr(L, f"{'p' if comm is not None else ''}{p_}naupd_c")
    rwp = rwork.ctypes.data_as(C.POINTER(rt))
    cargs = (ido.ctypes.data_as(c_int_p), bmat.encode(), n, which.encode(), nev, rt(tol), _addr(res), ncv, _addr(v), ldv,
             iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), _addr(workd), _addr(workl), lworkl, rwp,
             info.ctypes.data_as(c_int_p))
    if comm is not None:
        if p_ != "z":
            raise ArpackB200Error("solve_complex(comm=...) binds pznaupd_c/pzneupd_c only (complex128)")
        cargs = (comm,) + cargs
    nsteps = 0
    while True:
        f_aupd(*cargs)
        if info[0] == INFO_DEVICE_ERROR:
            raise ArpackB200Error("naupd_c (complex): CUDA device error (see stderr); there is no CPU fallback")
        i = ido[0]
        if i == 1 or i == -1:
            x = workd[ipntr[0] - 1: ipntr[0] - 1 + n]
            y = workd[ipntr[1] - 1: ipntr[1] - 1 + n]
            if mode == 3 and bmat == "G":
                op(x, y, workd[ipntr[2] - 1: ipntr[2] - 1 + n] if i == 1 else None)
            else:
                op(x, y)
            nsteps += 1
        elif i == 2:
            bop(workd[ipntr[0] - 1: ipntr[0] - 1 + n], workd[ipntr[1] - 1: ipntr[1] - 1 + n])
        else:
            break
    out = Result(info=int(info[0]), iparam=iparam.copy(), ipntr=ipntr.copy(), workl=workl.copy(), v=v, resid=res,
                 nconv=int(iparam[4]), nsteps=nsteps, workd=workd)
    if info[0] < 0 or not eupd:
        L.ab200_release(workl.ctypes.data)
        return out
    select = np.zeros(ncv, dtype=np.int32)
    ierr = np.zeros(1, dtype=np.int32)
    d = np.zeros(nev + 1, dtype=np_dt)
    workev = np.zeros(2 * ncv, dtype=np_dt)
    sg = complex(sigma)
    eargs = (int(rvec), howmny.encode(), select.ctypes.data_as(c_int_p), _addr(d), _addr(v), ldv, rt(sg.real), rt(sg.imag),
             _addr(workev), bmat.encode(), n, which.encode(), nev, rt(tol), _addr(res), ncv, _addr(v), ldv,
             iparam.ctypes.data_as(c_int_p), ipntr.ctypes.data_as(c_int_p), _addr(workd), _addr(workl), lworkl, rwp,
             ierr.ctypes.data_as(c_int_p))
    if comm is not None:
        L.ab200_pzneupd_ri(comm, *eargs)
    else:
        getattr(L, f"ab200_{p_}neupd_ri")(*eargs)
    if ierr[0] == INFO_DEVICE_ERROR:
        raise ArpackB200Error("neupd_c (complex): CUDA device error (see stderr)")
    out.update(d=d[:nev], z=v, ierr=int(ierr[0]), workl_eupd=workl.copy(), ipntr_eupd=ipntr.copy())
    L.ab200_release(workl.ctypes.data)
    return out


# ---- driver-side operators on the device ------------------------------------------------------------
class CsrOperator:
    """A CSR matrix resident in HBM with y = A x on the library's stream (K3)."""

    def __init__(self, n, rowptr, col, val, ncols=None):
        self.n, self.rowptr, self.col, self.val = n, rowptr, col, val
        self.ncols = ncols or n
        self.nnz = int(val.numel())
        self._spmv_fn = None
        self._spmv_args = None

    @staticmethod
    def laplace2d(nx, ny=None, scale=1.0, device="cuda"):
        import torch
        ny = ny or nx
        L = lib()
        n = nx * ny
        nnz = L.ab200_gen_laplace2d(nx, ny, scale, None, None, None)
        if nnz < 0:
            raise ArpackB200Error("laplace2d: nnz exceeds int32")
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=device)
        col = torch.empty(nnz, dtype=torch.int32, device=device)
        val = torch.empty(nnz, dtype=torch.float64, device=device)
        if L.ab200_gen_laplace2d(nx, ny, scale, rowptr.data_ptr(), col.data_ptr(), val.data_ptr()) != nnz:
            raise ArpackB200Error("laplace2d generator failed")
        return CsrOperator(n, rowptr, col, val)

    @staticmethod
    def convdiff2d(nx, rho=100.0, device="cuda"):
        import torch
        L = lib()
        n = nx * nx
        nnz = L.ab200_gen_convdiff2d(nx, rho, None, None, None)
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=device)
        col = torch.empty(nnz, dtype=torch.int32, device=device)
        val = torch.empty(nnz, dtype=torch.float64, device=device)
        if L.ab200_gen_convdiff2d(nx, rho, rowptr.data_ptr(), col.data_ptr(), val.data_ptr()) != nnz:
            raise ArpackB200Error("convdiff2d generator failed")
        return CsrOperator(n, rowptr, col, val)

    @staticmethod
    def laplace3d(nx, ny, nz, z0=0, nzloc=None, device="cuda", diag=6.0):
        import torch
        L = lib()
        nzloc = nz if nzloc is None else nzloc
        nloc = nx * ny * nzloc
        nnz = L.ab200_gen_laplace3d(nx, ny, nz, z0, nzloc, diag, None, None, None)
        if nnz < 0:
            raise ArpackB200Error("laplace3d: nnz exceeds int32")
        rowptr = torch.empty(nloc + 1, dtype=torch.int32, device=device)
        col = torch.empty(nnz, dtype=torch.int32, device=device)
        val = torch.empty(nnz, dtype=torch.float64, device=device)
        if L.ab200_gen_laplace3d(nx, ny, nz, z0, nzloc, diag, rowptr.data_ptr(), col.data_ptr(), val.data_ptr()) != nnz:
            raise ArpackB200Error("laplace3d generator failed")
        op = CsrOperator(nloc, rowptr, col, val)
        op.halo_lo = nx * ny if z0 > 0 else 0
        op.halo_hi = nx * ny if z0 + nzloc < nz else 0
        op.halo = torch.zeros(max(1, op.halo_lo + op.halo_hi), dtype=torch.float64, device=device)
        return op

    @staticmethod
    def from_scipy(A, device="cuda"):
        import torch
        A = A.tocsr()
        A.sort_indices()
        return CsrOperator(A.shape[0], torch.as_tensor(A.indptr.astype(np.int32), device=device),
                           torch.as_tensor(A.indices.astype(np.int32), device=device),
                           torch.as_tensor(A.data.astype(np.float64), device=device), ncols=A.shape[1])

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.cpu().numpy(), self.col.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=(self.n, self.ncols))

    def __call__(self, x, y, *_):
        """y = A x; x, y device tensors (views of workd) or host numpy arrays (host-buffer RCI loop)."""
        if getattr(self, "halo_lo", 0) + getattr(self, "halo_hi", 0) > 0:
            raise ArpackB200Error("row-partitioned operator: use apply_halo(comm, x, y)")
        L = lib()
        if isinstance(x, np.ndarray):
            rc = L.ab200_csr_spmv_hostvec_f64(self.n, self.ncols, self.rowptr.data_ptr(), self.col.data_ptr(),
                                              self.val.data_ptr(), x.ctypes.data, y.ctypes.data)
        else:
            import torch
            f = L.ab200_csr_spmv_f64 if self.val.dtype == torch.float64 else L.ab200_csr_spmv_f32
            rc = f(self.n, self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr(), x.data_ptr(),
                   y.data_ptr())
        if rc != 0:
            raise ArpackB200Error(f"csr_spmv failed ({rc})")

    def apply_ptr(self, xptr, yptr):
        """y = A x on raw device addresses (the lean RCI loop of solve())."""
        f = self._spmv_fn
        if f is None:
            import torch
            L = lib()
            f = self._spmv_fn = L.ab200_csr_spmv_f64 if self.val.dtype == torch.float64 else L.ab200_csr_spmv_f32
            self._spmv_args = (self.n, self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr())
        if f(*self._spmv_args, xptr, yptr) != 0:
            raise ArpackB200Error("csr_spmv failed")

    def halo_address(self, comm):
        """Where the neighbours' planes arrive.  The first use under a communicator is COLLECTIVE: the communicator is
        asked for its own halo buffer (ab200_comm_halo_buffer: the neighbours then store their planes straight into it
        over NVLink); without peer memory the operator's ordinary device buffer is used (ncclSend/ncclRecv)."""
        if getattr(self, "_halo_comm", None) != comm:
            self._halo_comm = comm
            self._halo_ptr = None
            if self.halo_lo + self.halo_hi > 0:
                p = lib().ab200_comm_halo_buffer(comm, self.halo_lo, self.halo_hi, self.val.element_size())
                self._halo_ptr = p if p else None
        return self._halo_ptr if self._halo_ptr is not None else self.halo.data_ptr()

    def apply_halo_ptr(self, comm, xptr, yptr):
        if lib().ab200_csr_spmv_halo_f64(comm, self.n, self.halo_lo, self.halo_hi, self.rowptr.data_ptr(),
                                        self.col.data_ptr(), self.val.data_ptr(), xptr, yptr,
                                        self.halo_address(comm)) != 0:
            raise ArpackB200Error("csr_spmv_halo failed")

    def apply_halo(self, comm, x, y):
        rc = lib().ab200_csr_spmv_halo_f64(comm, self.n, self.halo_lo, self.halo_hi, self.rowptr.data_ptr(),
                                          self.col.data_ptr(), self.val.data_ptr(), x.data_ptr(), y.data_ptr(),
                                          self.halo_address(comm))
        if rc != 0:
            raise ArpackB200Error(f"csr_spmv_halo failed ({rc})")

    def spmv_bytes(self, w=8):
        """Algorithmic traffic of one product: nnz*(w+4) + (n+1)*4 + 2*n*w."""
        return self.nnz * (w + 4) + (self.n + 1) * 4 + 2 * self.n * w

    def residuals(self, d, z, ldz):
        out = np.zeros(len(d), dtype=np.float64)
        dd = np.ascontiguousarray(d, dtype=np.float64)
        rc = lib().ab200_residuals_f64(self.n, self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr(), len(d),
                                       z.data_ptr(), ldz, dd.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise ArpackB200Error("residuals failed")
        return out


class HostCsr:
    """A CSR matrix in HOST memory (int32 indices), as a caller of the reference owns it (arpackSolver.hpp:361-424 reads
    the matrix into host memory).  Pass it as ``solve(None, ..., registered_op=HostCsr(...), host_buffers=True)``: the
    library uploads it once and runs the whole solve on the GPU."""

    def __init__(self, n, rowptr, col, val):
        self.n, self.rowptr, self.col, self.val = int(n), rowptr, col, val
        self.nnz = int(val.shape[0])

    @staticmethod
    def from_operator(A, pinned=True):
        """Host copy of a CsrOperator (pinned by default, like alloc_host_buffers)."""
        def _h(t):
            t = t.cpu()
            return t.pin_memory() if pinned else t
        return HostCsr(A.n, _h(A.rowptr), _h(A.col), _h(A.val))

    def nbytes(self):
        return sum(int(a.numel() * a.element_size()) if hasattr(a, "numel") else int(a.nbytes)
                   for a in (self.rowptr, self.col, self.val))


def hashed_start_vector(n, i0=0, seed=0x5EED, device="cuda"):
    """resid[i] = 2 u(i0+i) - 1, u = top 53 bits of splitmix64(seed + i) / 2^53."""
    import torch
    x = torch.empty(n, dtype=torch.float64, device=device)
    if lib().ab200_fill_hash_f64(n, i0, seed, x.data_ptr()) != 0:
        raise ArpackB200Error("fill_hash failed")
    return x


def hashed_start_vector_numpy(n, i0=0, seed=0x5EED):
    """CPU twin of hashed_start_vector (same bits), for the oracle side of parity tests."""
    with np.errstate(over="ignore"):
        x = (np.arange(i0, i0 + n, dtype=np.uint64) + np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return 2.0 * u - 1.0


def nccl_comm_from_torch_distributed():
    """Create the library's NCCL communicator for the current torch.distributed world (one rank per GPU).
    torch.distributed is only the bootstrap (it ships the 128-byte unique id); the data path is the library's own
    communicator, used for the fused length-(j+1) all-reduces and the halo exchange."""
    import torch
    import torch.distributed as dist
    L = lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = np.zeros(128, dtype=np.uint8)
    if rank == 0 and L.ab200_nccl_unique_id(buf.ctypes.data) != 0:
        raise ArpackB200Error("ncclGetUniqueId failed")
    t = torch.from_numpy(buf)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    buf = t.cpu().numpy()
    h = L.ab200_comm_create(buf.ctypes.data, rank, world)
    if h < 1:
        raise ArpackB200Error(f"ab200_comm_create failed ({h})")
    return h


def slab_partition(nlines, world, rank):
    """PARPACK-style block-row layout (dsaupd.f:331-349; icb_parpack_c.c:60-69): `nlines` grid lines/planes split
    over `world` ranks, the remainder spread over the first ranks.  Returns (first line, number of lines)."""
    base, rem = divmod(nlines, world)
    cnt = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, cnt


# ---- OP = A^T A on a row-sharded sparse A (BASELINE config 5: SVD through dsaupd, EXAMPLES/SVD/dsvd.f) ---------------
def randsparse_numpy(row0, nrows, ncols, per_row=16, seed=0x5EED):
    """CPU twin of ab200_gen_randsparse (same bits): rows [row0, row0+nrows) of the synthetic matrix of SURVEY.md 8d as a
    scipy CSR matrix (duplicate columns of a row are summed, exactly as the product does)."""
    import scipy.sparse as sp

    def mix(x):
        with np.errstate(over="ignore"):
            x = x + np.uint64(0x9E3779B97F4A7C15)
            x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return x ^ (x >> np.uint64(31))
    with np.errstate(over="ignore"):
        r = np.repeat(np.arange(row0, row0 + nrows, dtype=np.uint64), per_row)
        k = np.tile(np.arange(per_row, dtype=np.uint64), nrows)
        h1 = mix(np.uint64(seed) + np.uint64(per_row) * r + k)
    h2 = mix(h1)
    col = (h1 % np.uint64(ncols)).astype(np.int64)
    val = 2.0 * ((h2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0
    rows = (r - np.uint64(row0)).astype(np.int64)
    return sp.csr_matrix((val, (rows, col)), shape=(nrows, ncols))


class GramOperator:
    """z = A^T A x for a row-sharded sparse A resident in HBM (gram.cu): the caller's `av` + `atv` of
    EXAMPLES/SVD/dsvd.f:342-343.  comm = None: one process, vectors of ncols entries.  comm = handle: this rank owns
    `shards` row blocks of A and the slice [rank*ncols/world, ...) of the eigenproblem vectors (PARPACK's layout)."""

    def __init__(self, ncols, comm=None):
        self.ncols, self.comm = int(ncols), comm
        self.h = lib().ab200_gram_create(comm or 0, self.ncols)
        if self.h < 1:
            raise ArpackB200Error(f"ab200_gram_create failed ({self.h}): ncols must be a multiple of the rank count")
        self._keep = []   # the shards' device arrays stay alive as long as the operator
        self.nnz = 0
        self.rows = 0
        world = lib().ab200_comm_size(comm) if comm else 1
        self.n = self.ncols // world     # local length of the eigenproblem vectors

    def add_shard(self, row0, nrows, per_row=16, seed=0x5EED, device="cuda"):
        """Generate rows [row0, row0+nrows) of the synthetic matrix on the device, build the shard's transpose."""
        import torch
        L = lib()
        nnz = L.ab200_gen_randsparse(row0, nrows, self.ncols, per_row, seed, None, None, None)
        if nnz < 0:
            raise ArpackB200Error("randsparse: shard too large for int32 indices")
        i32, f64 = torch.int32, torch.float64
        rp, col, val = (torch.empty(nrows + 1, dtype=i32, device=device), torch.empty(nnz, dtype=i32, device=device),
                        torch.empty(nnz, dtype=f64, device=device))
        if L.ab200_gen_randsparse(row0, nrows, self.ncols, per_row, seed, rp.data_ptr(), col.data_ptr(),
                                  val.data_ptr()) != nnz:
            raise ArpackB200Error("randsparse generator failed")
        return self.add_csr(nrows, rp, col, val)

    def add_csr(self, nrows, rp, col, val):
        import torch
        L = lib()
        nnz = int(val.numel())
        trp = torch.empty(self.ncols + 1, dtype=torch.int32, device=val.device)
        tcol = torch.empty(nnz, dtype=torch.int32, device=val.device)
        tval = torch.empty(nnz, dtype=torch.float64, device=val.device)
        if L.ab200_csr_transpose_f64(nrows, self.ncols, nnz, rp.data_ptr(), col.data_ptr(), val.data_ptr(),
                                     trp.data_ptr(), tcol.data_ptr(), tval.data_ptr()) != 0:
            raise ArpackB200Error("csr transpose failed")
        if L.ab200_gram_add_shard(self.h, nrows, nnz, rp.data_ptr(), col.data_ptr(), val.data_ptr(), trp.data_ptr(),
                                  tcol.data_ptr(), tval.data_ptr()) != 0:
            raise ArpackB200Error("gram_add_shard failed")
        self._keep.append((rp, col, val, trp, tcol, tval))
        self.nnz += nnz
        self.rows += nrows
        return self

    @staticmethod
    def randsparse(nrows_total, ncols, per_row=16, seed=0x5EED, comm=None, shard_rows=None, device="cuda"):
        """The config-5 operator: A is nrows_total x ncols; under a communicator this rank takes its block of rows
        (slab_partition), which is further cut into shards of at most shard_rows rows (default: so that a shard's
        product vector is 20 MB, comfortably L2-resident next to x)."""
        L = lib()
        world = L.ab200_comm_size(comm) if comm else 1
        rank = L.ab200_comm_rank(comm) if comm else 0
        first, cnt = slab_partition(nrows_total, world, rank)
        shard_rows = shard_rows or 2_500_000
        G = GramOperator(ncols, comm)
        r = first
        while r < first + cnt:
            m = min(shard_rows, first + cnt - r)
            G.add_shard(r, m, per_row, seed, device)
            r += m
        return G

    def apply_ptr(self, xptr, yptr):
        if lib().ab200_gram_apply(self.h, xptr, yptr) != 0:
            raise ArpackB200Error("gram_apply failed")

    def apply_comm_ptr(self, comm, xptr, yptr):
        self.apply_ptr(xptr, yptr)

    def __call__(self, x, y, *_):
        self.apply_ptr(x.data_ptr(), y.data_ptr())

    def spmv_bytes(self):
        """Algorithmic traffic of one OP: both CSR copies streamed once, x and the shard products gathered once."""
        return 2 * self.nnz * 12 + (self.rows + len(self._keep) * (self.ncols + 2)) * 4 + \
            (self.ncols * 8 * 3 + self.rows * 16) * 1

    def close(self):
        if self.h:
            lib().ab200_gram_destroy(self.h)
            self.h = 0
            self._keep = []
