// irl_complex.hpp -- implicitly restarted Arnoldi for complex matrices: the znaupd/zneupd (cnaupd/cneupd) entry points
// (SURVEY.md 8f row 4).
//
//   aupd()          <->  SRC/znaupd.f:384-664 + SRC/znaup2.f:171-801
//   start_vector()  <->  SRC/zgetv0.f:114-416
//   extend()        <->  SRC/znaitr.f:209-850
//   ritz_bounds()   <->  SRC/zneigh.f:103-257   (LAPACK zlahqr + ztrevc on the host)
//   select_wanted() <->  SRC/zngets.f:93-178 ;  the sorts are SRC/zsortc.f
//   shift_sweeps()  <->  SRC/znapps.f:144-435   (complex Givens sweeps on the ncv x ncv Hessenberg matrix, host)
//   eupd()          <->  SRC/zneupd.f:248-876
//
// Every n-length operation goes through VecOps<std::complex<R>> (device kernels, vecops_cplx.cu); its reductions are
// Hermitian: dot(x, y) = sum conj(x_i) y_i, dots(V, x) = V^H x.  Norms arrive as complex mailbox entries with a zero
// imaginary part.  Same resumable-routine structure as irl_base.hpp.
#pragma once
#include <complex>

#include "hostmath_cplx.hpp"
#include "irl_base.hpp"

namespace ab200 {

template <typename R>
class IrlComplex {
  using Z = std::complex<R>;
  using L = Lapack<R>;
  using LZ = LapackZ<R>;

 public:
  // parpack: the semantics of PARPACK/SRC/MPI/pzn*.f (per-rank seeds, the initial OP*x only for bmat = 'G', five
  // start-vector refinements, eps23 with a REAL exponent, no ncv > n test, machine constants per call); the
  // all-reduces themselves are the ops' allreduce_sum, a no-op on one rank
  IrlComplex(VecOps<Z>* ops, SeedState* seed, R* smlnum_first, bool parpack = false)
      : ops_(ops), seed_(seed), smlnum_first_(smlnum_first), par_(parpack) {}

  Counters cnt;
  const Counters& counters() const { return cnt; }
  R tol_effective = 0;

  void ensure_mailbox(int ncv) {
    if (!mb_) {
      ncv_ = ncv;
      setup_mailbox();
    }
  }

  void aupd(int* ido, char bmat, int n, const char* which, int nev, R* tol, Z* resid_dev, int ncv, Z* v_dev,
            int64_t ldv, int* iparam, int* ipntr, Z* workd_dev, Z* workl, int lworkl, R* rwork, int* info) {
    if (*ido == 0) {
      cnt = Counters();  // zstatn (znaupd.f:448)
      int ierr = 0;
      ishift_ = iparam[0];
      mxiter_ = iparam[2];
      mode_ = iparam[6];
      which_ = key_of(which);
      const bool okw = which_ == Key::LM || which_ == Key::SM || which_ == Key::LR || which_ == Key::SR ||
                       which_ == Key::LI || which_ == Key::SI;
      if (n <= 0) ierr = -1;
      else if (nev <= 0) ierr = -2;
      else if (ncv <= nev || (!par_ && ncv > n)) ierr = -3;  // znaupd.f:467 (zneupd is stricter); pznaupd.f:481
      else if (mxiter_ <= 0) ierr = -4;
      else if (!okw) ierr = -5;
      else if (bmat != 'I' && bmat != 'G') ierr = -6;
      else if (lworkl < 3 * ncv * ncv + 5 * ncv) ierr = -7;
      else if (mode_ < 1 || mode_ > 3) ierr = -10;
      else if (mode_ == 1 && bmat == 'G') ierr = -11;
      if (ierr != 0) {
        *info = ierr;
        *ido = 99;
        return;
      }
      if (*tol <= R(0)) *tol = L::lamch("E");
      if (ishift_ != 0 && ishift_ != 1 && ishift_ != 2) ishift_ = 1;
      n_ = n; ncv_ = ncv; bmat_ = bmat;
      resid_ = resid_dev; v_ = v_dev; ldv_ = ldv; workd_ = workd_dev;
      nev0_ = nev; np0_ = ncv - nev; nev_ = nev0_; np_ = np0_; kplusp_ = ncv;
      std::fill(workl, workl + 3 * (size_t)ncv * ncv + 5 * (size_t)ncv, Z(0));
      // workl partition (znaupd.f:531-548), 0-based offsets
      ldh_ = ncv; ldq_ = ncv;
      ih_ = 0; iritz_ = ih_ + ldh_ * ncv; ibounds_ = iritz_ + ncv; iq_ = ibounds_ + ncv; iw_ = iq_ + ldq_ * ncv;
      ipntr[3] = iw_ + ncv * ncv + 3 * ncv + 1;
      ipntr[4] = ih_ + 1; ipntr[5] = iritz_ + 1; ipntr[6] = iq_ + 1; ipntr[7] = ibounds_ + 1;
      ipntr[13] = iw_ + 1;
      setup_mailbox();
      eps23_ = eps23_of<R>(L::lamch("E"), par_);  // pznaup2.f:277: REAL exponent
      // machine constants of znaitr/znapps (znaitr.f:303-317): smlnum from the n of the first ever call
      unfl_ = L::lamch("S");
      R ovfl = R(1) / unfl_;
      L::labad(unfl_, ovfl);
      ulp_ = L::lamch("P");
      if (par_ || *smlnum_first_ < R(0)) *smlnum_first_ = unfl_ * (R(n) / ulp_);  // pcontext: per call under PARPACK
      smlnum_ = *smlnum_first_;
      nconv_ = 0; iter_ = 0;
      initv_ = (*info != 0);
      *info = 0;
      info_ = 0;
      pc_ = 0; gv_pc_ = 0; ai_pc_ = 0;
    }
    wl_ = workl;
    rwork_ = rwork;
    tol_ = *tol;
    tol_effective = tol_;
    const bool done = run();
    if (!done) {
      *ido = ido_;
      ipntr[0] = ipntr_[0]; ipntr[1] = ipntr_[1]; ipntr[2] = ipntr_[2];
      if (ido_ == 3) iparam[7] = np_;
      return;
    }
    *ido = 99;
    iparam[2] = mxiter_out_;
    iparam[4] = np_;
    iparam[8] = cnt.nopx; iparam[9] = cnt.nbx; iparam[10] = cnt.nrorth;
    *info = info_;
    if (*info == 2) *info = 3;
    if (*info >= 0 && trace_levels().mcaupd > 0 && ops_->rank() == 0) {  // znaupd.f:603-660
      trace::ivout1(mxiter_out_, "_naupd: Number of update iterations taken");
      trace::ivout1(np_, "_naupd: Number of wanted \"converged\" Ritz values");
      trace::zvout(np_, ritz(), "_naupd: The final Ritz values");
      trace::zvout(np_, bounds(), "_naupd: Associated Ritz estimates");
      trace::summary("Complex implicit Arnoldi update code", mxiter_out_, cnt.nopx, cnt.nbx, cnt.nrorth, cnt.nitref,
                     cnt.nrstrt);
    }
  }

  // ---------------------------------------------------------------------------------------------
  // zneupd: eigenvalues d on the host, Ritz/Schur vectors in z (device, may alias v)
  // ---------------------------------------------------------------------------------------------
  void eupd(bool rvec, char howmny, int* select, Z* d, Z* z_dev, int64_t ldz, Z sigma, Z* workev, char bmat, int n,
            const char* which, int nev, R tol, Z* resid_dev, int ncv, Z* v_dev, int64_t ldv, int* iparam, int* ipntr,
            Z* workd_dev, Z* workl, int lworkl, R* rwork, int* info) {
    (void)workd_dev;
    const int mode = iparam[6];
    int nconv = iparam[4];
    *info = 0;
    const R eps23 = eps23_of<R>(L::lamch("E"), par_);  // pzneupd.f:359
    int ierr = 0;
    const Key wk = key_of(which);
    const bool okw = wk == Key::LM || wk == Key::SM || wk == Key::LR || wk == Key::SR || wk == Key::LI || wk == Key::SI;
    if (nconv <= 0) ierr = -14;
    else if (n <= 0) ierr = -1;
    else if (nev <= 0) ierr = -2;
    else if (ncv <= nev + 1 || ncv > n) ierr = -3;
    else if (!okw) ierr = -5;
    else if (bmat != 'I' && bmat != 'G') ierr = -6;
    else if (lworkl < 3 * ncv * ncv + 4 * ncv) ierr = -7;
    else if ((howmny != 'A' && howmny != 'P' && howmny != 'S') && rvec) ierr = -13;
    else if (howmny == 'S') ierr = -12;
    bool shifti = false;
    if (mode == 1 || mode == 2) shifti = false;
    else if (mode == 3) shifti = true;
    else ierr = -10;
    if (mode == 1 && bmat == 'G') ierr = -11;
    if (ierr != 0) { *info = ierr; return; }

    // workl layout (zneupd.f:431-460), 0-based offsets
    const int ih = ipntr[4] - 1, ritz = ipntr[5] - 1, bounds = ipntr[7] - 1;
    const int ldh = ncv, ldq = ncv;
    const int iheig = bounds + ldh, ihbds = iheig + ldh, iuptri = ihbds + ldh, invsub = iuptri + ldh * ncv;
    ipntr[8] = iheig + 1; ipntr[10] = ihbds + 1; ipntr[11] = iuptri + 1; ipntr[12] = invsub + 1;
    const int irz = ipntr[13] - 1 + ncv * ncv, ibd = irz + ncv;
    Z* W = workl;
    const Z rnorm = W[ih + 2];  // smuggled by aupd (znaup2.f:541)
    W[ih + 2] = Z(0);
    if (rvec) {
      bool reord = false;
      for (int j = 0; j < ncv; ++j) { W[bounds + j] = Z(R(j + 1)); select[j] = 0; }
      select_wanted(wk, 0, nev, ncv - nev, W + irz, W + bounds);
      int numcnv = 0;
      for (int j = 1; j <= ncv; ++j) {
        const R rtemp = std::max(eps23, LZ::abs(W[irz + ncv - j]));
        const int jj = (int)W[bounds + ncv - j].real();
        if (numcnv < nconv && LZ::abs(W[ibd + jj - 1]) <= tol * rtemp) {
          select[jj - 1] = 1;
          numcnv++;
          if (jj > (par_ ? nev : nconv)) reord = true;  // pzneupd.f:547 compares with nev
        }
      }
      if (numcnv != nconv) { *info = -15; return; }
      // Schur form of H and its Schur vectors (zneupd.f:567-577)
      std::copy(W + ih, W + ih + (size_t)ldh * ncv, W + iuptri);
      for (int j = 0; j < ncv; ++j)
        for (int i = 0; i < ncv; ++i) W[invsub + (size_t)j * ldq + i] = (i == j) ? Z(1) : Z(0);
      ierr = LZ::lahqr(true, true, ncv, 1, ncv, W + iuptri, ldh, W + iheig, 1, ncv, W + invsub, ldq);
      for (int j = 0; j < ncv; ++j) W[ihbds + j] = W[invsub + (size_t)j * ldq + ncv - 1];
      if (ierr != 0) { *info = -8; return; }
      if (reord) {
        int nconv2 = 0;
        ierr = LZ::trsen_NV(select, ncv, W + iuptri, ldh, W + invsub, ldq, W + iheig, &nconv2, workev, ncv);
        if (par_) nconv = nconv2;  // pzneupd.f:611-615 lets ztrsen overwrite nconv
        else if (nconv2 < nconv) nconv = nconv2;
        if (ierr == 1) { *info = 1; return; }
      }
      for (int j = 0; j < ncv; ++j) W[ihbds + j] = W[invsub + (size_t)j * ldq + ncv - 1];
      if (!shifti) std::copy(W + iheig, W + iheig + nconv, d);
      // Orthonormal basis of the wanted invariant subspace: QR of the leading Schur vectors, V <- V*Q1
      // (zneupd.f:654-670).  The reference applies the reflectors to the n x ncv array V (zunm2r); here the small
      // unitary factor Q1 is formed on the host and applied in one device pass.
      LZ::geqr2(ncv, nconv, W + invsub, ldq, workev, workev + ncv);
      std::vector<Z> q1((size_t)ncv * ncv, Z(0)), wk2((size_t)ncv);
      for (int i = 0; i < ncv; ++i) q1[(size_t)i * ncv + i] = Z(1);
      LZ::unm2r("R", "N", ncv, ncv, nconv, W + invsub, ldq, workev, q1.data(), ncv, wk2.data());
      ops_->vq_update(n, ncv, ncv, v_dev, ldv, q1.data(), ncv, false, Z(0), Z(0), 0, nullptr, nullptr);
      for (int j = 0; j < nconv; ++j) {  // zneupd.f:672-688
        if (W[invsub + (size_t)j * ldq + j].real() < R(0)) {
          for (int k = 0; k < nconv; ++k) W[iuptri + j + (size_t)k * ldq] = -W[iuptri + j + (size_t)k * ldq];
          for (int k = 0; k < nconv; ++k) W[iuptri + (size_t)j * ldq + k] = -W[iuptri + (size_t)j * ldq + k];
        }
      }
      if (howmny == 'A') {
        for (int j = 0; j < ncv; ++j) select[j] = (j < nconv) ? 1 : 0;
        int outncv = 0;
        Z vl[1];
        ierr = LZ::trevc("R", "S", select, ncv, W + iuptri, ldq, vl, 1, W + invsub, ldq, ncv, &outncv, workev, rwork);
        if (ierr != 0) { *info = -9; return; }
        for (int j = 0; j < nconv; ++j) {  // zneupd.f:722-737
          Z* cj = W + invsub + (size_t)j * ldq;
          const R rt = R(1) / LZ::nrm2(ncv, cj, 1);
          for (int i = 0; i < ncv; ++i) cj[i] *= rt;
          Z acc(0);
          for (int i = 0; i <= j; ++i) acc += std::conj(W[ihbds + i]) * cj[i];  // zzdotc(j, ihbds, invsub(:,j))
          workev[j] = acc;
        }
        std::copy(workev, workev + nconv, W + ihbds);
        // Z = (V*Q1)(:,1:nconv) * E, E = the upper-triangular eigenvector block (ztrmm, zneupd.f:760-763)
        std::vector<Z> m((size_t)nconv * nconv, Z(0));
        for (int c = 0; c < nconv; ++c)
          for (int r = 0; r <= c; ++r) m[(size_t)c * nconv + r] = W[invsub + (size_t)c * ldq + r];
        ops_->vq_out(n, nconv, nconv, v_dev, ldv, m.data(), nconv, z_dev, ldz);
      } else {
        if (z_dev != v_dev) ops_->copy2d(n, nconv, v_dev, ldv, z_dev, ldz);
      }
    } else {
      std::copy(W + ritz, W + ritz + nconv, d);
      std::copy(W + ritz, W + ritz + nconv, W + iheig);
      std::copy(W + bounds, W + bounds + nconv, W + ihbds);
    }
    // Ritz estimates and back-transformation (zneupd.f:786-824)
    if (rvec)
      for (int k = 0; k < ncv; ++k) W[ihbds + k] *= rnorm;
    if (shifti) {
      for (int k = 0; k < ncv; ++k) {
        const Z t = W[iheig + k];
        W[ihbds + k] = W[ihbds + k] / t / t;
      }
      for (int k = 0; k < nconv; ++k) d[k] = Z(1) / W[iheig + k] + sigma;
    }
    // eigenvector purification for shift-invert (zneupd.f:845-868)
    if (rvec && howmny == 'A' && shifti) {
      for (int j = 0; j < nconv; ++j)
        if (W[iheig + j] != Z(0)) workev[j] = W[invsub + (size_t)j * ldq + ncv - 1] / W[iheig + j];
      ops_->ger(n, nconv, resid_dev, workev, z_dev, ldz);
    }
  }

 private:
  VecOps<Z>* ops_;
  SeedState* seed_;
  R* smlnum_first_;
  const bool par_;
  int n_ = 0, ncv_ = 0, mode_ = 1;
  char bmat_ = 'I';
  Z *resid_ = nullptr, *v_ = nullptr, *workd_ = nullptr;
  int64_t ldv_ = 0;
  int ido_ = 0, ipntr_[3] = {0, 0, 0};
  // mailbox: three segments of seg_ complex entries
  Z* mb_ = nullptr;
  int seg_ = 0;
  std::vector<Z> mbh_;
  Z* mbA() { return mb_; }
  Z* mbC() { return mb_ + 2 * seg_; }
  Z* hA() { return mbh_.data(); }
  Z* hC() { return mbh_.data() + 2 * seg_; }
  void setup_mailbox() {
    seg_ = ncv_ + 2;
    mb_ = ops_->mailbox((size_t)3 * seg_);
    mbh_.assign((size_t)3 * seg_, Z(0));
  }
  Z* vcol(int j1) { return v_ + (int64_t)(j1 - 1) * ldv_; }
  Z* slot(int off1) { return workd_ + (off1 - 1); }
  static constexpr int IPJ = 1;
  int irj() const { return 1 + n_; }
  int ivj() const { return 1 + 2 * n_; }

  int pc_ = 0;
  int ishift_ = 1, mxiter_ = 0, mxiter_out_ = 0;
  Key which_ = Key::NONE;
  int nev0_ = 0, np0_ = 0, nev_ = 0, np_ = 0, kplusp_ = 0, nconv_ = 0, iter_ = 0, info_ = 0;
  bool initv_ = false;
  int ldh_ = 0, ldq_ = 0, ih_ = 0, iritz_ = 0, ibounds_ = 0, iq_ = 0, iw_ = 0;
  Z* wl_ = nullptr;
  R* rwork_ = nullptr;
  R tol_ = 0, eps23_ = 0, unfl_ = 0, ulp_ = 0, smlnum_ = 0, rnorm_ = 0;

  Z& H(int i, int j) { return wl_[ih_ + (i - 1) + (size_t)(j - 1) * ldh_]; }
  Z& Q(int i, int j) { return wl_[iq_ + (i - 1) + (size_t)(j - 1) * ldq_]; }
  Z* ritz() { return wl_ + iritz_; }
  Z* bounds() { return wl_ + ibounds_; }
  Z* wrk() { return wl_ + iw_; }
  static R abs1(Z z) { return std::fabs(z.real()) + std::fabs(z.imag()); }  // zabs1 of znapps.f:199-201

  // B-norm from a Hermitian inner product in the mailbox: sqrt(dlapy2(re, im)) (zgetv0.f:300-301)
  R fetch_norm_from_dot(Z* mbslot) {
    ops_->allreduce_sum(mbslot, 1);
    ops_->fetch(hC(), mbslot, 1);
    return std::sqrt(LZ::abs(hC()[0]));
  }

  // ---------------------------------------------------------------------------------------------
  // start / restart vector (zgetv0.f)
  // ---------------------------------------------------------------------------------------------
  int gv_pc_ = 0, gv_itry_ = 1, gv_j_ = 1, gv_iter_ = 0, gv_ierr_ = 0;
  bool gv_initv_ = false;
  R gv_rnorm0_ = 0;

  bool start_vector() {
    CO_BEGIN(gv_pc_)
    if (!seed_->inited) {
      if (!par_) {  // zgetv0.f:196-202
        seed_->iseed[0] = 1; seed_->iseed[1] = 3; seed_->iseed[2] = 5; seed_->iseed[3] = 7;
      } else {  // pzgetv0.f:210-222, digits of 1000 + 2*rank + 1
        int igen = 1000 + 2 * ops_->rank() + 1;
        seed_->iseed[0] = igen / 1000; igen %= 1000;
        seed_->iseed[1] = igen / 100;  igen %= 100;
        seed_->iseed[2] = igen / 10;
        seed_->iseed[3] = igen % 10;
      }
      seed_->inited = true;
    }
    gv_ierr_ = 0;
    gv_iter_ = 0;
    if (!gv_initv_) ops_->larnv_uniform_m1_1(n_, seed_->iseed, resid_);  // zlarnv(idist = 2)
    if (par_ ? (bmat_ == 'G') : (gv_itry_ == 1)) {  // force the vector into range(OP) (zgetv0.f:238-245; pzgetv0.f:249)
      cnt.nopx++;
      ops_->copy(n_, resid_, slot(1));
      ipntr_[0] = 1; ipntr_[1] = n_ + 1;
      ido_ = -1;
      CO_YIELD(gv_pc_);
      ops_->copy(n_, slot(n_ + 1), resid_);
    } else if (!par_ && bmat_ == 'G') {
      ops_->copy(n_, resid_, slot(n_ + 1));
    }
    if (bmat_ == 'G') {
      cnt.nbx++;
      ipntr_[0] = n_ + 1; ipntr_[1] = 1;
      ido_ = 2;
      CO_YIELD(gv_pc_);
      ops_->dot(n_, resid_, slot(1), mbC());
    } else {
      ops_->dot(n_, resid_, resid_, mbC());
    }
    gv_rnorm0_ = fetch_norm_from_dot(mbC());
    rnorm_ = gv_rnorm0_;
    if (gv_j_ > 1) {
      // orthogonalise against V(:,1:j-1); ONE refinement at most (zgetv0.f:326-383)
      for (;;) {
        ops_->dots(n_, gv_j_ - 1, v_, ldv_, bmat_ == 'G' ? slot(1) : resid_, resid_, mbA());
        ops_->allreduce_sum(mbA(), (size_t)gv_j_ - 1);
        if (bmat_ == 'G') {
          ops_->update(n_, gv_j_ - 1, v_, ldv_, mbA(), resid_, resid_, nullptr);
          cnt.nbx++;
          ops_->copy(n_, resid_, slot(n_ + 1));
          ipntr_[0] = n_ + 1; ipntr_[1] = 1;
          ido_ = 2;
          CO_YIELD(gv_pc_);
          ops_->dot(n_, resid_, slot(1), mbC());
        } else {
          ops_->update(n_, gv_j_ - 1, v_, ldv_, mbA(), resid_, resid_, mbC());
        }
        rnorm_ = fetch_norm_from_dot(mbC());
        if (rnorm_ > dgks_threshold<R>() * gv_rnorm0_) break;
        gv_iter_++;
        if (gv_iter_ <= (par_ ? 5 : 1)) {  // zgetv0.f:376 ; pzgetv0.f:372
          gv_rnorm0_ = rnorm_;
        } else {
          ops_->zero(n_, resid_);
          rnorm_ = 0;
          gv_ierr_ = -1;
          break;
        }
      }
    }
    if (trace_levels().mgetv0 > 0 && ops_->rank() == 0)  // zgetv0.f:392-395
      trace::dvout1(rnorm_, "_getv0: B-norm of initial / restarted starting vector");
    CO_END(gv_pc_)
  }

  // ---------------------------------------------------------------------------------------------
  // k -> k+np step extension (znaitr.f)
  // ---------------------------------------------------------------------------------------------
  int ai_pc_ = 0, ai_k_ = 0, ai_np_ = 0, ai_j_ = 0, ai_itry_ = 0, ai_iter_ = 0, ai_info_ = 0;
  R ai_wnorm_ = 0, ai_beta_ = 0, ai_rnorm1_ = 0;

  void h_add(int j, const Z* s) {
    for (int i = 1; i <= j; ++i) H(i, j) += s[i - 1];
  }

  bool extend() {
    CO_BEGIN(ai_pc_)
    ai_info_ = 0;
    for (ai_j_ = ai_k_ + 1; ai_j_ <= ai_k_ + ai_np_; ++ai_j_) {
      ai_beta_ = rnorm_;
      if (!(rnorm_ > R(0))) {
        // invariant subspace: new vector orthogonal to the current basis (znaitr.f:396-440)
        ai_beta_ = 0;
        cnt.nrstrt++;
        if (trace_levels().mcaitr > 0 && ops_->rank() == 0) trace::ivout1(ai_j_, "_naitr: ****** RESTART AT STEP ******");  // znaitr.f:397-402
        for (ai_itry_ = 1; ai_itry_ <= 3; ++ai_itry_) {
          gv_itry_ = ai_itry_; gv_initv_ = false; gv_j_ = ai_j_;
          CO_CALL(ai_pc_, start_vector());
          if (gv_ierr_ >= 0) break;
        }
        if (gv_ierr_ < 0) {
          ai_info_ = ai_j_ - 1;
          ai_pc_ = 0;
          return true;
        }
      }
      // v_j = r/||r||, p_j = B r/||r||, x = v_j   (znaitr.f:442-481)
      {
        Z* bx = (bmat_ == 'I' && mode_ == 1) ? nullptr : slot(IPJ);
        if (rnorm_ >= unfl_) {
          ops_->start_step(n_, Z(R(1) / rnorm_), resid_, vcol(ai_j_), slot(ivj()), bx, bmat_ == 'I');
        } else {
          // zlascl fallback of the reference (znaitr.f:463-467): scale in two safe steps
          const R big = std::ldexp(R(1), sizeof(R) == 8 ? 500 : 60);
          ops_->start_step(n_, Z(R(1) / (rnorm_ * big)), resid_, vcol(ai_j_), slot(ivj()), slot(IPJ), bmat_ == 'I');
          ops_->scal(n_, Z(big), vcol(ai_j_));
          ops_->scal(n_, Z(big), slot(ivj()));
          ops_->scal(n_, Z(big), slot(IPJ));
        }
      }
      cnt.nopx++;
      ipntr_[0] = ivj(); ipntr_[1] = irj(); ipntr_[2] = IPJ;
      ido_ = 1;
      CO_YIELD(ai_pc_);
      // workd(irj) = OP*v_j ; the residual is formed from it without an intermediate copy
      if (bmat_ == 'G') {
        cnt.nbx++;
        ipntr_[0] = irj(); ipntr_[1] = IPJ;
        ido_ = 2;
        CO_YIELD(ai_pc_);
      }
      // h(1:j,j) = V_j^H (B w), <B w, w>; r = w - V_j h   (znaitr.f:529-561)
      ops_->dots(n_, ai_j_, v_, ldv_, bmat_ == 'G' ? slot(IPJ) : slot(irj()), slot(irj()), mbA());
      ops_->allreduce_sum(mbA(), (size_t)ai_j_ + 1);
      ops_->update(n_, ai_j_, v_, ldv_, mbA(), slot(irj()), resid_, bmat_ == 'I' ? mbC() : nullptr);
      if (bmat_ == 'I') {
        ops_->allreduce_sum(mbC(), 1);
        ops_->fetch(mbh_.data(), mb_, (size_t)2 * seg_ + 1);  // h, ||w||^2 and ||r||^2 in one round trip
      } else {
        ops_->fetch(hA(), mbA(), (size_t)ai_j_ + 1);
      }
      ai_wnorm_ = std::sqrt(LZ::abs(hA()[ai_j_]));
      for (int i = 1; i <= ai_j_; ++i) H(i, ai_j_) = hA()[i - 1];
      if (ai_j_ > 1) H(ai_j_, ai_j_ - 1) = Z(ai_beta_);
      if (bmat_ == 'G') {
        cnt.nbx++;
        ops_->copy(n_, resid_, slot(irj()));
        ipntr_[0] = irj(); ipntr_[1] = IPJ;
        ido_ = 2;
        CO_YIELD(ai_pc_);
        ops_->dot(n_, resid_, slot(IPJ), mbC());
        rnorm_ = fetch_norm_from_dot(mbC());
      } else {
        rnorm_ = std::sqrt(LZ::abs(hC()[0]));
      }
      if (!(rnorm_ > dgks_threshold<R>() * ai_wnorm_)) {
        // DGKS re-orthogonalisation, at most two passes (znaitr.f:619-742)
        cnt.nrorth++;
        ai_iter_ = 0;
        for (;;) {
          ops_->dots(n_, ai_j_, v_, ldv_, bmat_ == 'G' ? slot(IPJ) : resid_, resid_, mbA());
          ops_->allreduce_sum(mbA(), (size_t)ai_j_);
          ops_->update(n_, ai_j_, v_, ldv_, mbA(), resid_, resid_, bmat_ == 'I' ? mbC() : nullptr);
          if (bmat_ == 'G') {
            ops_->fetch(hA(), mbA(), (size_t)ai_j_);
            h_add(ai_j_, hA());
            cnt.nbx++;
            ops_->copy(n_, resid_, slot(irj()));
            ipntr_[0] = irj(); ipntr_[1] = IPJ;
            ido_ = 2;
            CO_YIELD(ai_pc_);
            ops_->dot(n_, resid_, slot(IPJ), mbC());
            ai_rnorm1_ = fetch_norm_from_dot(mbC());
          } else {
            ops_->allreduce_sum(mbC(), 1);
            ops_->fetch(mbh_.data(), mb_, (size_t)2 * seg_ + 1);
            h_add(ai_j_, hA());
            ai_rnorm1_ = std::sqrt(LZ::abs(hC()[0]));
          }
          if (ai_rnorm1_ > dgks_threshold<R>() * rnorm_) {
            rnorm_ = ai_rnorm1_;
            break;
          }
          cnt.nitref++;
          rnorm_ = ai_rnorm1_;
          ai_iter_++;
          if (ai_iter_ > 1) {
            ops_->zero(n_, resid_);
            rnorm_ = 0;
            break;
          }
        }
      }
    }
    // zlahqr-style deflation test on the new sub-diagonals (znaitr.f:768-782)
    {
      std::vector<R> work((size_t)(ai_k_ + ai_np_));
      for (int i = std::max(1, ai_k_); i <= ai_k_ + ai_np_ - 1; ++i) {
        R tst1 = LZ::abs(H(i, i)) + LZ::abs(H(i + 1, i + 1));
        if (tst1 == R(0)) tst1 = LZ::lanhs1(ai_k_ + ai_np_, &H(1, 1), ldh_, work.data());
        if (LZ::abs(H(i + 1, i)) <= std::max(ulp_ * tst1, smlnum_)) H(i + 1, i) = Z(0);
      }
    }
    CO_END(ai_pc_)
  }

  // zneigh.f:172-219
  int ritz_bounds() {
    const int m = kplusp_;
    Z* wl = wrk();
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) wl[(size_t)j * m + i] = H(i + 1, j + 1);
    for (int j = 1; j <= m; ++j)
      for (int i = 1; i <= m; ++i) Q(i, j) = (i == j) ? Z(1) : Z(0);
    int ierr = LZ::lahqr(true, true, m, 1, m, wl, ldh_, ritz(), 1, m, &Q(1, 1), ldq_);
    if (ierr != 0) return ierr;
    int sel[1] = {0}, mout = 0;
    Z vl[1];
    ierr = LZ::trevc("R", "B", sel, m, wl, m, vl, m, &Q(1, 1), ldq_, m, &mout, wl + (size_t)m * m, rwork_);
    if (ierr != 0) return ierr;
    for (int j = 1; j <= m; ++j) {
      const R t = R(1) / LZ::nrm2(m, &Q(1, j), 1);
      for (int r = 1; r <= m; ++r) Q(r, j) *= t;
    }
    for (int j = 1; j <= m; ++j) bounds()[j - 1] = Q(m, j) * rnorm_;
    return 0;
  }

  // zngets.f:139-151
  static void select_wanted(Key which, int ishift, int kev, int np, Z* rz, Z* bnd) {
    sort_z<R>(which, kev + np, rz, bnd);
    if (ishift == 1) sort_z<R>(Key::SM, np, bnd, rz);
  }

  // znapps.f:252-433: apply the np shifts to H with complex Givens rotations, accumulating Q
  void shift_sweeps(int kev, int np, const Z* shift) {
    const int kp = kev + np;
    std::vector<R> lw((size_t)kp);
    for (int j = 1; j <= kp; ++j)
      for (int i = 1; i <= kp; ++i) Q(i, j) = (i == j) ? Z(1) : Z(0);
    if (np == 0) return;
    for (int jj = 1; jj <= np; ++jj) {
      const Z sigma = shift[jj - 1];
      int istart = 1, iend;
      do {
        iend = kp;
        for (int i = istart; i <= kp - 1; ++i) {
          R tst1 = abs1(H(i, i)) + abs1(H(i + 1, i + 1));
          if (tst1 == R(0)) tst1 = LZ::lanhs1(kp - jj + 1, &H(1, 1), ldh_, lw.data());
          if (std::fabs(H(i + 1, i).real()) <= std::max(ulp_ * tst1, smlnum_)) {
            iend = i;
            H(i + 1, i) = Z(0);
            break;
          }
        }
        if (!(istart == iend || istart > kev)) {
          Z f = H(istart, istart) - sigma, g = H(istart + 1, istart), s, r;
          R c;
          for (int i = istart; i <= iend - 1; ++i) {
            LZ::lartg(f, g, c, s, r);
            if (i > istart) {
              H(i, i - 1) = r;
              H(i + 1, i - 1) = Z(0);
            }
            for (int j = i; j <= kp; ++j) {
              const Z t = c * H(i, j) + s * H(i + 1, j);
              H(i + 1, j) = -std::conj(s) * H(i, j) + c * H(i + 1, j);
              H(i, j) = t;
            }
            for (int j = 1; j <= std::min(i + 2, iend); ++j) {
              const Z t = c * H(j, i) + std::conj(s) * H(j, i + 1);
              H(j, i + 1) = -s * H(j, i) + c * H(j, i + 1);
              H(j, i) = t;
            }
            for (int j = 1; j <= std::min(i + jj, kp); ++j) {
              const Z t = c * Q(j, i) + std::conj(s) * Q(j, i + 1);
              Q(j, i + 1) = -s * Q(j, i) + c * Q(j, i + 1);
              Q(j, i) = t;
            }
            if (i < iend - 1) { f = H(i + 1, i); g = H(i + 2, i); }
          }
        }
        istart = iend + 1;
      } while (iend < kp);
    }
    // real non-negative sub-diagonal in the leading kev block (znapps.f:404-414)
    for (int j = 1; j <= kev; ++j) {
      if (H(j + 1, j).real() < R(0) || H(j + 1, j).imag() != R(0)) {
        const Z t = H(j + 1, j) / LZ::abs(H(j + 1, j));
        const Z tc = std::conj(t);
        for (int c2 = j; c2 <= kp; ++c2) H(j + 1, c2) *= tc;
        for (int r2 = 1; r2 <= std::min(j + 2, kp); ++r2) H(r2, j + 1) *= t;
        for (int r2 = 1; r2 <= std::min(j + np + 1, kp); ++r2) Q(r2, j + 1) *= t;
        H(j + 1, j) = Z(H(j + 1, j).real());
      }
    }
    for (int i = 1; i <= kev; ++i) {
      R tst1 = abs1(H(i, i)) + abs1(H(i + 1, i + 1));
      if (tst1 == R(0)) tst1 = LZ::lanhs1(kev, &H(1, 1), ldh_, lw.data());
      if (H(i + 1, i).real() <= std::max(ulp_ * tst1, smlnum_)) H(i + 1, i) = Z(0);
    }
  }

  bool run() {
    CO_BEGIN(pc_)
    gv_itry_ = 1; gv_initv_ = initv_; gv_j_ = 1;
    CO_CALL(pc_, start_vector());
    if (rnorm_ == R(0)) {  // znaup2.f:326-333 (through label 1100: mxiter = iter, nev = nconv)
      info_ = -9;
      mxiter_out_ = iter_;
      np_ = np0_;
      CO_END_EARLY(pc_);
    }
    ai_k_ = 0; ai_np_ = nev_;
    CO_CALL(pc_, extend());
    if (ai_info_ > 0) { fail_no_factorisation(); CO_END_EARLY(pc_); }
    for (;;) {
      iter_++;
      np_ = kplusp_ - nev_;  // znaup2.f:397
      if (trace_levels().mcaup2 > 0 && ops_->rank() == 0) {  // znaup2.f:391-409
        trace::ivout1(iter_, "_naup2: **** Start of major iteration number ****");
        if (trace_levels().mcaup2 > 1) {
          trace::ivout1(nev_, "_naup2: The length of the current Arnoldi factorization");
          trace::ivout1(np_, "_naup2: Extend the Arnoldi factorization by");
        }
      }
      ai_k_ = nev_; ai_np_ = np_;
      CO_CALL(pc_, extend());
      if (ai_info_ > 0) { fail_no_factorisation(); CO_END_EARLY(pc_); }
      if (ritz_bounds() != 0) {
        info_ = -8;
        mxiter_out_ = mxiter_;
        CO_END_EARLY(pc_);
      }
      nev_ = nev0_; np_ = np0_;
      std::copy(ritz(), ritz() + kplusp_, wrk() + kplusp_ * kplusp_);
      std::copy(bounds(), bounds() + kplusp_, wrk() + kplusp_ * kplusp_ + kplusp_);
      select_wanted(which_, ishift_, nev_, np_, ritz(), bounds());
      nconv_ = 0;  // znaup2.f:489-497
      for (int i = 0; i < nev_; ++i)
        if (LZ::abs(bounds()[np_ + i]) <= tol_ * std::max(eps23_, LZ::abs(ritz()[np_ + i]))) nconv_++;
      if (trace_levels().mcaup2 > 2 && ops_->rank() == 0) {  // znaup2.f:499-509
        const int kp[3] = {nev_, np_, nconv_};
        trace::ivout(3, kp, "_naup2: NEV, NP, NCONV are");
        trace::zvout(kplusp_, ritz(), "_naup2: The eigenvalues of H");
        trace::zvout(kplusp_, bounds(), "_naup2: Ritz estimates of the current NCV Ritz values");
      }
      {
        const int nptemp = np_;
        for (int j = 0; j < nptemp; ++j)
          if (bounds()[j] == Z(0)) { np_--; nev_++; }
      }
      if (nconv_ >= nev0_ || iter_ > mxiter_ || np_ == 0) {
        finish_sorted();
        break;
      } else if (nconv_ < nev0_ && ishift_ == 1) {
        const int nevbef = nev_;
        nev_ += std::min(nconv_, np_ / 2);
        if (nev_ == 1 && kplusp_ >= 6) nev_ = kplusp_ / 2;
        else if (nev_ == 1 && kplusp_ > 3) nev_ = 2;
        np_ = kplusp_ - nev_;
        if (nevbef < nev_) select_wanted(which_, ishift_, nev_, np_, ritz(), bounds());
      }
      if (ishift_ == 0) {
        ido_ = 3;
        CO_YIELD(pc_);
      }
      if (ishift_ != 1) std::copy(wrk(), wrk() + np_, ritz());  // znaup2.f:676-685
      // implicit restart: host sweeps, then V <- V*Q, r <- sigma_k r + beta_k v_{kev+1} and ||r|| in one device pass
      shift_sweeps(nev_, np_, ritz());
      {
        const Z sigmak = Q(kplusp_, nev_), betak = H(nev_ + 1, nev_);
        const bool has_beta = betak.real() > R(0);
        ops_->vq_update(n_, kplusp_, nev_ + (has_beta ? 1 : 0), v_, ldv_, wl_ + iq_, ldq_, true, sigmak,
                        has_beta ? betak : Z(0), has_beta ? nev_ : -1, resid_, bmat_ == 'I' ? mbC() : nullptr);
      }
      if (bmat_ == 'G') {
        cnt.nbx++;
        ops_->copy(n_, resid_, slot(n_ + 1));
        ipntr_[0] = n_ + 1; ipntr_[1] = 1;
        ido_ = 2;
        CO_YIELD(pc_);
        ops_->dot(n_, resid_, slot(1), mbC());
      }
      rnorm_ = fetch_norm_from_dot(mbC());
    }
    CO_END(pc_)
  }

  void fail_no_factorisation() {
    np_ = ai_info_;
    mxiter_out_ = iter_;
    info_ = -9999;
  }

  // exit ordering of (ritz, bounds) (znaup2.f:541-616)
  void finish_sorted() {
    H(3, 1) = Z(rnorm_);  // for eupd
    Z *rz = ritz(), *b = bounds();
    Key wp = Key::NONE;
    switch (which_) {
      case Key::LM: wp = Key::SM; break;
      case Key::SM: wp = Key::LM; break;
      case Key::LR: wp = Key::SR; break;
      case Key::SR: wp = Key::LR; break;
      case Key::LI: wp = Key::SI; break;
      case Key::SI: wp = Key::LI; break;
      default: break;
    }
    sort_z<R>(wp, kplusp_, rz, b);
    for (int j = 0; j < nev0_; ++j) b[j] = b[j] / std::max(eps23_, LZ::abs(rz[j]));
    sort_z<R>(Key::LM, nev0_, b, rz);
    for (int j = 0; j < nev0_; ++j) b[j] = b[j] * std::max(eps23_, LZ::abs(rz[j]));
    sort_z<R>(which_, nconv_, rz, b);
    if (iter_ > mxiter_ && nconv_ < nev0_) info_ = 1;
    if (np_ == 0 && nconv_ < nev0_) info_ = 2;
    np_ = nconv_;
    mxiter_out_ = iter_;
  }
};

}  // namespace ab200
