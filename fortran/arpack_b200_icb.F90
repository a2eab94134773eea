! arpack_b200_icb.F90 -- ISO_C_BINDING view of libarpack_b200.so for Fortran host code.
!
! north_star keeps the host side in Fortran, calling CUDA through a thin ISO_C_BINDING layer that extends the
! reference's ICB *_c symbols.  The reference's own shims (SRC/icbads.F90:3-92, SRC/icbadn.F90:3-97,
! SRC/icbazn.F90:3-94, PARPACK/SRC/MPI/icbpds.F90) go from C to the Fortran routines; this module is the same set of
! signatures seen from the other side: a Fortran program (EXAMPLES/SIMPLE/dssimp.f:302-324 is the model) calls the C
! symbols the GPU library exports.  Argument order, by-value scalars and array shapes are those of the reference shims.
!
! NOT COMPILED in this repository: the build image has no Fortran compiler (DESIGN.md section 0).  The module is
! shipped as the interface contract and has been reviewed by hand only.  Legacy callers need nothing from it: the
! library also exports the Fortran-77 names dsaupd_ dseupd_ dnaupd_ dneupd_ znaupd_ zneupd_ (gfortran ABI), so
! `call dsaupd(...)` resolves to the GPU library when it is linked in place of libarpack.
!
! Device memory: resid, v, workd (and z) may be host arrays (the library mirrors them in HBM) or device arrays.
! With device arrays, declare them with the CUDA Fortran `device` attribute or obtain them from cudaMalloc through
! c_ptr/c_f_pointer; the hand-off slots workd(ipntr(1)) / workd(ipntr(2)) are then device addresses and the user's
! OP kernel must be ordered on the stream returned by ab200_get_stream().
module arpack_b200_icb
  use, intrinsic :: iso_c_binding
  implicit none
  public

  interface
    ! ---- ICB/arpack.h:14-15 (SRC/icbads.F90) ----
    subroutine dsaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, &
                        info) bind(c, name="dsaupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), intent(inout) :: ido, info
      character(kind=c_char), intent(in) :: bmat(*), which(*)
      integer(c_int), value :: n, nev, ncv, ldv, lworkl
      real(c_double), value :: tol
      real(c_double) :: resid(*), v(ldv, *), workd(*), workl(*)
      integer(c_int) :: iparam(11), ipntr(11)
    end subroutine dsaupd_c

    subroutine dseupd_c(rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v, ldv, &
                        iparam, ipntr, workd, workl, lworkl, info) bind(c, name="dseupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), value :: rvec, ldz, n, nev, ncv, ldv, lworkl
      character(kind=c_char), intent(in) :: howmny(*), bmat(*), which(*)
      integer(c_int) :: select(*), iparam(11), ipntr(11)
      integer(c_int), intent(inout) :: info
      real(c_double), value :: sigma, tol
      real(c_double) :: d(*), z(ldz, *), resid(*), v(ldv, *), workd(*), workl(*)
    end subroutine dseupd_c

    ! ---- ICB/arpack.h:12-13 (SRC/icbadn.F90) ----
    subroutine dnaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, &
                        info) bind(c, name="dnaupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), intent(inout) :: ido, info
      character(kind=c_char), intent(in) :: bmat(*), which(*)
      integer(c_int), value :: n, nev, ncv, ldv, lworkl
      real(c_double), value :: tol
      real(c_double) :: resid(*), v(ldv, *), workd(*), workl(*)
      integer(c_int) :: iparam(11), ipntr(14)
    end subroutine dnaupd_c

    subroutine dneupd_c(rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which, nev, tol, &
                        resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info) bind(c, name="dneupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), value :: rvec, ldz, n, nev, ncv, ldv, lworkl
      character(kind=c_char), intent(in) :: howmny(*), bmat(*), which(*)
      integer(c_int) :: select(*), iparam(11), ipntr(14)
      integer(c_int), intent(inout) :: info
      real(c_double), value :: sigmar, sigmai, tol
      real(c_double) :: dr(*), di(*), z(ldz, *), workev(*), resid(*), v(ldv, *), workd(*), workl(*)
    end subroutine dneupd_c

    ! ---- ICB/arpack.h:20-21 (SRC/icbazn.F90) ----
    subroutine znaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, &
                        rwork, info) bind(c, name="znaupd_c")
      import :: c_int, c_double, c_double_complex, c_char
      integer(c_int), intent(inout) :: ido, info
      character(kind=c_char), intent(in) :: bmat(*), which(*)
      integer(c_int), value :: n, nev, ncv, ldv, lworkl
      real(c_double), value :: tol
      complex(c_double_complex) :: resid(*), v(ldv, *), workd(*), workl(*)
      real(c_double) :: rwork(*)
      integer(c_int) :: iparam(11), ipntr(14)
    end subroutine znaupd_c

    subroutine zneupd_c(rvec, howmny, select, d, z, ldz, sigma, workev, bmat, n, which, nev, tol, resid, ncv, v, &
                        ldv, iparam, ipntr, workd, workl, lworkl, rwork, info) bind(c, name="zneupd_c")
      import :: c_int, c_double, c_double_complex, c_char
      integer(c_int), value :: rvec, ldz, n, nev, ncv, ldv, lworkl
      character(kind=c_char), intent(in) :: howmny(*), bmat(*), which(*)
      integer(c_int) :: select(*), iparam(11), ipntr(14)
      integer(c_int), intent(inout) :: info
      complex(c_double_complex), value :: sigma
      real(c_double), value :: tol
      complex(c_double_complex) :: d(*), z(ldz, *), workev(*), resid(*), v(ldv, *), workd(*), workl(*)
      real(c_double) :: rwork(*)
    end subroutine zneupd_c

    ! ---- ICB/parpack.h:20-21 (PARPACK/SRC/MPI/icbpds.F90); comm = handle from ab200_comm_create ----
    subroutine pdsaupd_c(comm, ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, &
                         lworkl, info) bind(c, name="pdsaupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), value :: comm
      integer(c_int), intent(inout) :: ido, info
      character(kind=c_char), intent(in) :: bmat(*), which(*)
      integer(c_int), value :: n, nev, ncv, ldv, lworkl
      real(c_double), value :: tol
      real(c_double) :: resid(*), v(ldv, *), workd(*), workl(*)
      integer(c_int) :: iparam(11), ipntr(11)
    end subroutine pdsaupd_c

    subroutine pdseupd_c(comm, rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v, &
                         ldv, iparam, ipntr, workd, workl, lworkl, info) bind(c, name="pdseupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), value :: comm, rvec, ldz, n, nev, ncv, ldv, lworkl
      character(kind=c_char), intent(in) :: howmny(*), bmat(*), which(*)
      integer(c_int) :: select(*), iparam(11), ipntr(11)
      integer(c_int), intent(inout) :: info
      real(c_double), value :: sigma, tol
      real(c_double) :: d(*), z(ldz, *), resid(*), v(ldv, *), workd(*), workl(*)
    end subroutine pdseupd_c

    ! ---- ICB/parpack.h:22-23 (PARPACK/SRC/MPI/icbpdn.F90) ----
    subroutine pdnaupd_c(comm, ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, &
                         lworkl, info) bind(c, name="pdnaupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), value :: comm
      integer(c_int), intent(inout) :: ido, info
      character(kind=c_char), intent(in) :: bmat(*), which(*)
      integer(c_int), value :: n, nev, ncv, ldv, lworkl
      real(c_double), value :: tol
      real(c_double) :: resid(*), v(ldv, *), workd(*), workl(*)
      integer(c_int) :: iparam(11), ipntr(14)
    end subroutine pdnaupd_c

    subroutine pdneupd_c(comm, rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which, nev, &
                         tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info) bind(c, name="pdneupd_c")
      import :: c_int, c_double, c_char
      integer(c_int), value :: comm, rvec, ldz, n, nev, ncv, ldv, lworkl
      character(kind=c_char), intent(in) :: howmny(*), bmat(*), which(*)
      integer(c_int) :: select(*), iparam(11), ipntr(14)
      integer(c_int), intent(inout) :: info
      real(c_double), value :: sigmar, sigmai, tol
      real(c_double) :: dr(*), di(*), z(ldz, *), workev(*), resid(*), v(ldv, *), workd(*), workl(*)
    end subroutine pdneupd_c

    ! ---- extensions of the GPU library (include/arpack_b200.h) ----
    ! stream on which the library enqueues its kernels and on which the ido = -1/1/2 hand-off is ordered
    function ab200_get_stream() bind(c, name="ab200_get_stream") result(stream)
      import :: c_ptr
      type(c_ptr) :: stream
    end function ab200_get_stream

    subroutine ab200_set_stream(stream) bind(c, name="ab200_set_stream")
      import :: c_ptr
      type(c_ptr), value :: stream
    end subroutine ab200_set_stream

    ! registered CSR operator (mode 1, bmat 'I'): the library applies OP itself, one d[sn]aupd_c call runs the solve.
    ! rowptr/col/val: host or device arrays, 0-based int32 indices.
    function ab200_register_csr_op_f64(workl, nrows, nnz, rowptr, col, val) bind(c, name="ab200_register_csr_op_f64") &
        result(rc)
      import :: c_int, c_long_long, c_double
      real(c_double), intent(in) :: workl(*)
      integer(c_int), value :: nrows
      integer(c_long_long), value :: nnz
      integer(c_int), intent(in) :: rowptr(*), col(*)
      real(c_double), intent(in) :: val(*)
      integer(c_int) :: rc
    end function ab200_register_csr_op_f64

    ! y = A x on the library's stream for a CSR matrix in HBM (the driver-side OP of EXAMPLES/MATRIX_MARKET)
    function ab200_csr_spmv_f64(nrows, rowptr, col, val, x, y) bind(c, name="ab200_csr_spmv_f64") result(rc)
      import :: c_int, c_double
      integer(c_int), value :: nrows
      integer(c_int), intent(in) :: rowptr(*), col(*)
      real(c_double), intent(in) :: val(*), x(*)
      real(c_double) :: y(*)
      integer(c_int) :: rc
    end function ab200_csr_spmv_f64

    ! drop the solve context (and its HBM mirrors) keyed to this workl
    subroutine ab200_release(workl) bind(c, name="ab200_release")
      import :: c_double
      real(c_double), intent(in) :: workl(*)
    end subroutine ab200_release

    ! NCCL communicator for the p*_c entry points: rank 0 fills id128, every rank receives it (MPI_Bcast), then creates
    function ab200_nccl_unique_id(id128) bind(c, name="ab200_nccl_unique_id") result(rc)
      import :: c_int, c_signed_char
      integer(c_signed_char) :: id128(128)
      integer(c_int) :: rc
    end function ab200_nccl_unique_id

    function ab200_comm_create(id128, rank, nranks) bind(c, name="ab200_comm_create") result(comm)
      import :: c_int, c_signed_char
      integer(c_signed_char), intent(in) :: id128(128)
      integer(c_int), value :: rank, nranks
      integer(c_int) :: comm
    end function ab200_comm_create
  end interface

contains

  ! The loop of EXAMPLES/SIMPLE/dssimp.f:302-324 with a registered host CSR matrix: one call, no hand-offs.
  ! which must be two characters; returns info of dsaupd_c (0 = converged, 1 = restart budget exhausted).
  subroutine ab200_dsaupd_csr(n, nnz, rowptr, col, val, which, nev, ncv, tol, mxiter, resid, v, ldv, workd, workl, &
                              lworkl, iparam, ipntr, info)
    integer(c_int), intent(in) :: n, nev, ncv, ldv, lworkl, mxiter
    integer(c_long_long), intent(in) :: nnz
    integer(c_int), intent(in) :: rowptr(*), col(*)
    real(c_double), intent(in) :: val(*), tol
    character(len=2), intent(in) :: which
    real(c_double) :: resid(*), v(ldv, *), workd(*), workl(*)
    integer(c_int) :: iparam(11), ipntr(11)
    integer(c_int), intent(inout) :: info
    integer(c_int) :: ido, rc
    character(kind=c_char) :: cbmat(2), cwhich(3)

    cbmat(1) = 'I'
    cbmat(2) = c_null_char
    cwhich(1) = which(1:1)
    cwhich(2) = which(2:2)
    cwhich(3) = c_null_char
    iparam(1) = 1
    iparam(3) = mxiter
    iparam(4) = 1
    iparam(7) = 1
    rc = ab200_register_csr_op_f64(workl, n, nnz, rowptr, col, val)
    if (rc /= 0) then
      info = -9990
      return
    end if
    ido = 0
    do
      call dsaupd_c(ido, cbmat, n, cwhich, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info)
      if (ido == 99) exit
      ! with a registered operator ido is 99 after the first call; any other value is a protocol error
      info = -9990
      exit
    end do
  end subroutine ab200_dsaupd_csr

end module arpack_b200_icb
