// irl_nonsym.hpp -- implicitly restarted Arnoldi: the dnaupd/dneupd (snaupd/sneupd) entry points.
//
//   aupd()         <->  SRC/dnaupd.f:406-693 + SRC/dnaup2.f:174-846  (pdnaupd.f / pdnaup2.f when parpack)
//   ritz_bounds()  <->  SRC/dneigh.f:100-318 (LAPACK xLAHQR + xTREVC on the host)
//   select()       <->  SRC/dngets.f:95-231,  count_converged() <-> SRC/dnconv.f:66-146
//   shift_sweeps() <->  SRC/dnapps.f:143-573 (Givens / double-shift Householder sweeps, host)
//   eupd()         <->  SRC/dneupd.f:302-1071
#pragma once
#include "irl_base.hpp"

namespace ab200 {

template <typename T>
class IrlNonsym : public IrlBase<T> {
  using B = IrlBase<T>;
  using B::ops_; using B::par_; using B::n_; using B::ncv_; using B::bmat_; using B::mode_; using B::resid_;
  using B::v_; using B::ldv_; using B::workd_; using B::ido_; using B::ipntr_; using B::rnorm_; using B::cnt;
  using B::mbC; using B::hC;
  using L = Lapack<T>;

 public:
  IrlNonsym(VecOps<T>* ops, bool parpack, SeedState* seed, T* smlnum_first) : B(ops, parpack) {
    this->seed_ = seed;
    smlnum_first_ = smlnum_first;
  }
  T tol_effective = 0;

  void aupd(int* ido, char bmat, int n, const char* which, int nev, T* tol, T* resid_dev, int ncv, T* v_dev,
            int64_t ldv, int* iparam, int* ipntr, T* workd_dev, T* workl, int lworkl, int* info) {
    if (*ido == 0) {
      cnt = Counters();  // dstatn (dnaupd.f:478)
      int ierr = 0;
      ishift_ = iparam[0];
      mxiter_ = iparam[2];
      mode_ = iparam[6];
      which_ = key_of(which);
      const bool okw = which_ == Key::LM || which_ == Key::SM || which_ == Key::LR || which_ == Key::SR ||
                       which_ == Key::LI || which_ == Key::SI;
      if (n <= 0) ierr = -1;
      else if (nev <= 0) ierr = -2;
      else if (ncv <= nev + 1 || (!par_ && ncv > n)) ierr = -3;
      else if (mxiter_ <= 0) ierr = -4;
      else if (!okw) ierr = -5;
      else if (bmat != 'I' && bmat != 'G') ierr = -6;
      else if (lworkl < 3 * ncv * ncv + 6 * ncv) ierr = -7;
      else if (mode_ < 1 || mode_ > 4) ierr = -10;
      else if (mode_ == 1 && bmat == 'G') ierr = -11;
      else if (ishift_ < 0 || ishift_ > 1) ierr = -12;
      if (ierr != 0) {
        *info = ierr;
        *ido = 99;
        return;
      }
      if (*tol <= T(0)) *tol = L::lamch("E");
      n_ = n; ncv_ = ncv; bmat_ = bmat;
      resid_ = resid_dev; v_ = v_dev; ldv_ = ldv; workd_ = workd_dev;
      nev0_ = nev; np0_ = ncv - nev; nev_ = nev0_; np_ = np0_; kplusp_ = ncv;
      std::fill(workl, workl + 3 * (size_t)ncv * ncv + 6 * (size_t)ncv, T(0));
      // workl partition (dnaupd.f:579-594), 0-based offsets
      ldh_ = ncv; ldq_ = ncv;
      ih_ = 0; iritzr_ = ih_ + ldh_ * ncv; iritzi_ = iritzr_ + ncv; ibounds_ = iritzi_ + ncv;
      iq_ = ibounds_ + ncv; iw_ = iq_ + ldq_ * ncv;
      ipntr[3] = iw_ + ncv * ncv + 3 * ncv + 1;
      ipntr[4] = ih_ + 1; ipntr[5] = iritzr_ + 1; ipntr[6] = iritzi_ + 1; ipntr[7] = ibounds_ + 1;
      ipntr[13] = iw_ + 1;
      this->setup_mailbox();
      eps_ = L::lamch("E");
      eps23_ = eps23_of<T>(eps_, par_);
      // machine constants as LAPACK xLAHQR sets them (dnaitr.f:297-313, dnapps.f:218-233).  The
      // reference computes smlnum once per process from the n of its FIRST call (SAVE'd `first`);
      // PARPACK recomputes it at every p*aupd call (pcontext).
      unfl_ = L::lamch("S");
      T ovfl = T(1) / unfl_;
      L::labad(unfl_, ovfl);
      ulp_ = L::lamch("P");
      if (par_ || *smlnum_first_ < T(0)) *smlnum_first_ = unfl_ * (T(n) / ulp_);
      smlnum_ = *smlnum_first_;
      nconv_ = 0; iter_ = 0;
      initv_ = (*info != 0);
      *info = 0;
      info_ = 0;
      pc_ = 0;
      this->gv_pc_ = 0; this->ai_pc_ = 0;
    }
    wl_ = workl;
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
    if (*info >= 0 && trace_levels().mnaupd > 0 && ops_->rank() == 0) {  // dnaupd.f:630-680
      trace::ivout1(mxiter_out_, "_naupd: Number of update iterations taken");
      trace::ivout1(np_, "_naupd: Number of wanted \"converged\" Ritz values");
      trace::dvout(np_, ritzr(), "_naupd: Real part of the final Ritz values");
      trace::dvout(np_, ritzi(), "_naupd: Imaginary part of the final Ritz values");
      trace::dvout(np_, bounds(), "_naupd: Associated Ritz estimates");
      trace::summary("Nonsymmetric implicit Arnoldi update code", mxiter_out_, cnt.nopx, cnt.nbx, cnt.nrorth,
                     cnt.nitref, cnt.nrstrt);
    }
  }

  // ---------------------------------------------------------------------------------------------
  // dneupd: eigenvalues (dr, di) on the host, Ritz/Schur vectors in z (device, may alias v)
  // ---------------------------------------------------------------------------------------------
  void eupd(bool rvec, char howmny, int* select, T* dr, T* di, T* z_dev, int64_t ldz, T sigmar, T sigmai,
            T* workev, char bmat, int n, const char* which, int nev, T tol, T* resid_dev, int ncv, T* v_dev,
            int64_t ldv, int* iparam, int* ipntr, T* workd_dev, T* workl, int lworkl, int* info) {
    (void)workd_dev;
    const int mode = iparam[6];
    int nconv = iparam[4];
    *info = 0;
    const T eps23 = eps23_of<T>(L::lamch("E"), par_);
    int ierr = 0;
    const Key wk = key_of(which);
    const bool okw = wk == Key::LM || wk == Key::SM || wk == Key::LR || wk == Key::SR || wk == Key::LI || wk == Key::SI;
    if (nconv <= 0) ierr = -14;
    else if (n <= 0) ierr = -1;
    else if (nev <= 0) ierr = -2;
    else if (ncv <= nev + 1 || (!par_ && ncv > n)) ierr = -3;
    else if (!okw) ierr = -5;
    else if (bmat != 'I' && bmat != 'G') ierr = -6;
    else if (lworkl < 3 * ncv * ncv + 6 * ncv) ierr = -7;
    else if ((howmny != 'A' && howmny != 'P' && howmny != 'S') && rvec) ierr = -13;
    else if (howmny == 'S') ierr = -12;
    enum { REGULR, SHIFTI, REALPT, IMAGPT } type = REGULR;
    if (mode == 1 || mode == 2) type = REGULR;
    else if (mode == 3 && sigmai == T(0)) type = SHIFTI;
    else if (mode == 3) type = REALPT;
    else if (mode == 4) type = IMAGPT;
    else ierr = -10;
    if (mode == 1 && bmat == 'G') ierr = -11;
    if (ierr != 0) { *info = ierr; return; }

    // workl layout (dneupd.f:464-521), 0-based offsets
    const int ih = ipntr[4] - 1, ritzr = ipntr[5] - 1, ritzi = ipntr[6] - 1, bounds = ipntr[7] - 1;
    const int ldh = ncv, ldq = ncv;
    const int iheigr = bounds + ldh, iheigi = iheigr + ldh, ihbds = iheigi + ldh, iuptri = ihbds + ldh,
              invsub = iuptri + ldh * ncv;
    ipntr[8] = iheigr + 1; ipntr[9] = iheigi + 1; ipntr[10] = ihbds + 1; ipntr[11] = iuptri + 1;
    ipntr[12] = invsub + 1;
    const int irr = ipntr[13] - 1 + ncv * ncv, iri = irr + ncv, ibd = iri + ncv;
    T* W = workl;
    const T rnorm = W[ih + 2];  // smuggled by aupd (dnaup2.f:552)
    W[ih + 2] = T(0);
    if (rvec) {
      bool reord = false;
      for (int j = 0; j < ncv; ++j) { W[bounds + j] = T(j + 1); select[j] = 0; }
      {
        int kev = nev, np = ncv - nev;
        select_wanted(wk, 0, kev, np, W + irr, W + iri, W + bounds);
      }
      int numcnv = 0;
      for (int j = 1; j <= ncv; ++j) {
        const T temp1 = std::max(eps23, L::lapy2(W[irr + ncv - j], W[iri + ncv - j]));
        const int jj = (int)W[bounds + ncv - j];
        if (numcnv < nconv && W[ibd + jj - 1] <= tol * temp1) {
          select[jj - 1] = 1;
          numcnv++;
          if (jj > nconv) reord = true;
        }
      }
      if (numcnv != nconv) { *info = -15; return; }
      // real Schur form of H and its Schur vectors (dneupd.f:622-637)
      std::copy(W + ih, W + ih + (size_t)ldh * ncv, W + iuptri);
      for (int j = 0; j < ncv; ++j)
        for (int i = 0; i < ncv; ++i) W[invsub + (size_t)j * ldq + i] = (i == j) ? T(1) : T(0);
      ierr = L::lahqr(true, true, ncv, 1, ncv, W + iuptri, ldh, W + iheigr, W + iheigi, 1, ncv, W + invsub, ldq);
      for (int j = 0; j < ncv; ++j) W[ihbds + j] = W[invsub + (size_t)j * ldq + ncv - 1];
      if (ierr != 0) { *info = -8; return; }
      if (reord) {
        int nconv2 = 0;
        ierr = L::trsen_NV(select, ncv, W + iuptri, ldh, W + invsub, ldq, W + iheigr, W + iheigi, &nconv2,
                           W + ihbds, ncv);
        if (nconv2 < nconv) nconv = nconv2;
        if (ierr == 1) { *info = 1; return; }
      }
      for (int j = 0; j < ncv; ++j) W[ihbds + j] = W[invsub + (size_t)j * ldq + ncv - 1];
      if (type == REGULR) {
        std::copy(W + iheigr, W + iheigr + nconv, dr);
        std::copy(W + iheigi, W + iheigi + nconv, di);
      }
      // Orthonormal basis of the wanted invariant subspace: QR of the leading Schur vectors,
      // V <- V*Q1 (dneupd.f:718-738).  Q1 is formed explicitly on the host.
      L::geqr2(ncv, nconv, W + invsub, ldq, workev, workev + ncv);
      std::vector<T> q1((size_t)ncv * ncv, T(0)), wk2((size_t)ncv);
      for (int i = 0; i < ncv; ++i) q1[(size_t)i * ncv + i] = T(1);
      L::orm2r("R", "N", ncv, ncv, nconv, W + invsub, ldq, workev, q1.data(), ncv, wk2.data());
      ops_->vq_update(n, ncv, ncv, v_dev, ldv, q1.data(), ncv, false, T(0), T(0), 0, nullptr, nullptr);
      for (int j = 0; j < nconv; ++j) {
        // keep the diagonal of R positive so that the Schur form is unchanged (dneupd.f:740-756)
        if (W[invsub + (size_t)j * ldq + j] < T(0)) {
          for (int k = 0; k < nconv; ++k) W[iuptri + j + (size_t)k * ldq] = -W[iuptri + j + (size_t)k * ldq];
          for (int k = 0; k < nconv; ++k) W[iuptri + (size_t)j * ldq + k] = -W[iuptri + (size_t)j * ldq + k];
        }
      }
      if (howmny == 'A') {
        for (int j = 0; j < ncv; ++j) select[j] = (j < nconv) ? 1 : 0;
        int outncv = 0;
        T vl[1];
        ierr = L::trevc("R", "S", select, ncv, W + iuptri, ldq, vl, 1, W + invsub, ldq, ncv, &outncv, workev);
        if (ierr != 0) { *info = -9; return; }
        // normalise the eigenvectors of the Schur block (dneupd.f:792-833)
        int iconj = 0;
        for (int j = 0; j < nconv; ++j) {
          T* cj = W + invsub + (size_t)j * ldq;
          if (W[iheigi + j] == T(0)) {
            const T t = L::nrm2(ncv, cj, 1);
            for (int i = 0; i < ncv; ++i) cj[i] *= T(1) / t;
          } else if (iconj == 0) {
            const T t = L::lapy2(L::nrm2(ncv, cj, 1), L::nrm2(ncv, cj + ldq, 1));
            for (int i = 0; i < ncv; ++i) { cj[i] *= T(1) / t; cj[ldq + i] *= T(1) / t; }
            iconj = 1;
          } else {
            iconj = 0;
          }
        }
        L::gemvT(ncv, nconv, W + invsub, ldq, W + ihbds, workev);
        iconj = 0;
        for (int j = 0; j < nconv; ++j) {
          if (W[iheigi + j] != T(0)) {
            if (iconj == 0) {
              workev[j] = L::lapy2(workev[j], workev[j + 1]);
              workev[j + 1] = workev[j];
              iconj = 1;
            } else {
              iconj = 0;
            }
          }
        }
        std::copy(workev, workev + nconv, W + ihbds);
        // Z = (V*Q1)(:,1:nconv) * E with E the (upper triangular) eigenvector block.  The reference
        // factors E = Q2*R2 and applies dorm2r + dtrmm to Z (dneupd.f:881-901); here the small
        // product M = Q2(1:nconv,1:nconv)*R2 is formed on the host and applied in one device pass.
        L::geqr2(ncv, nconv, W + invsub, ldq, workev, workev + ncv);
        std::vector<T> m((size_t)ncv * ncv, T(0));
        for (int i = 0; i < ncv; ++i) m[(size_t)i * ncv + i] = T(1);
        L::orm2r("R", "N", ncv, ncv, nconv, W + invsub, ldq, workev, m.data(), ncv, wk2.data());
        L::trmm_RUNN(ncv, nconv, W + invsub, ldq, m.data(), ncv);
        ops_->vq_out(n, nconv, nconv, v_dev, ldv, m.data(), ncv, z_dev, ldz);
      } else {
        if (z_dev != v_dev) ops_->copy2d(n, nconv, v_dev, ldv, z_dev, ldz);
      }
    } else {
      std::copy(W + ritzr, W + ritzr + nconv, dr);
      std::copy(W + ritzi, W + ritzi + nconv, di);
      std::copy(W + ritzr, W + ritzr + nconv, W + iheigr);
      std::copy(W + ritzi, W + ritzi + nconv, W + iheigi);
      std::copy(W + bounds, W + bounds + nconv, W + ihbds);
    }
    // back-transformation of Ritz values and estimates (dneupd.f:925-993)
    if (type == REGULR) {
      if (rvec)
        for (int k = 0; k < ncv; ++k) W[ihbds + k] *= rnorm;
    } else {
      if (type == SHIFTI) {
        if (rvec)
          for (int k = 0; k < ncv; ++k) W[ihbds + k] *= rnorm;
        for (int k = 0; k < ncv; ++k) {
          const T t = L::lapy2(W[iheigr + k], W[iheigi + k]);
          W[ihbds + k] = std::fabs(W[ihbds + k]) / t / t;
        }
        for (int k = 0; k < ncv; ++k) {
          const T t = L::lapy2(W[iheigr + k], W[iheigi + k]);
          W[iheigr + k] = W[iheigr + k] / t / t + sigmar;
          W[iheigi + k] = -W[iheigi + k] / t / t + sigmai;
        }
      }
      std::copy(W + iheigr, W + iheigr + nconv, dr);
      std::copy(W + iheigi, W + iheigi + nconv, di);
    }
    // eigenvector purification for shift-invert (dneupd.f:1017-1059)
    if (rvec && howmny == 'A' && type == SHIFTI) {
      int iconj = 0;
      for (int j = 0; j < nconv; ++j) {
        const T lr = W[invsub + (size_t)j * ldq + ncv - 1];
        if (W[iheigi + j] == T(0) && W[iheigr + j] != T(0)) {
          workev[j] = lr / W[iheigr + j];
        } else if (iconj == 0) {
          const T t = L::lapy2(W[iheigr + j], W[iheigi + j]);
          if (t != T(0)) {
            const T lr2 = W[invsub + (size_t)(j + 1) * ldq + ncv - 1];
            workev[j] = (lr * W[iheigr + j] + lr2 * W[iheigi + j]) / t / t;
            workev[j + 1] = (lr2 * W[iheigr + j] - lr * W[iheigi + j]) / t / t;
          }
          iconj = 1;
        } else {
          iconj = 0;
        }
      }
      ops_->ger(n, nconv, resid_dev, workev, z_dev, ldz);
    }
  }

 private:
  int pc_ = 0;
  int ishift_ = 1, mxiter_ = 0, mxiter_out_ = 0;
  Key which_ = Key::NONE;
  int nev0_ = 0, np0_ = 0, nev_ = 0, np_ = 0, kplusp_ = 0, nconv_ = 0, numcnv_ = 0, iter_ = 0, info_ = 0;
  bool initv_ = false;
  int ldh_ = 0, ldq_ = 0, ih_ = 0, iritzr_ = 0, iritzi_ = 0, ibounds_ = 0, iq_ = 0, iw_ = 0;
  T* wl_ = nullptr;
  T tol_ = 0, eps_ = 0, eps23_ = 0, unfl_ = 0, ulp_ = 0, smlnum_ = 0;
  T* smlnum_first_ = nullptr;

  T& H(int i, int j) { return wl_[ih_ + (i - 1) + (size_t)(j - 1) * ldh_]; }
  T& Q(int i, int j) { return wl_[iq_ + (i - 1) + (size_t)(j - 1) * ldq_]; }
  T* ritzr() { return wl_ + iritzr_; }
  T* ritzi() { return wl_ + iritzi_; }
  T* bounds() { return wl_ + ibounds_; }
  T* wrk() { return wl_ + iw_; }

  void h_store(int j, const T* hcol, T beta, bool) override {
    for (int i = 1; i <= j; ++i) H(i, j) = hcol[i - 1];
    if (j > 1) H(j, j - 1) = beta;
  }
  void h_add(int j, const T* scol, bool) override {
    for (int i = 1; i <= j; ++i) H(i, j) += scol[i - 1];
  }
  // dlahqr-style deflation test on the sub-diagonals of the new columns (dnaitr.f:798-811)
  void sweep_done(int k, int np) override {
    std::vector<T> work((size_t)(k + np));
    for (int i = std::max(1, k); i <= k + np - 1; ++i) {
      T tst1 = std::fabs(H(i, i)) + std::fabs(H(i + 1, i + 1));
      if (tst1 == T(0)) tst1 = L::lanhs1(k + np, &H(1, 1), ldh_, work.data());
      if (std::fabs(H(i + 1, i)) <= std::max(ulp_ * tst1, smlnum_)) H(i + 1, i) = T(0);
    }
  }
  int aitr_trace_level() const override { return trace_levels().mnaitr; }
  T tiny_norm() override { return unfl_; }

  int ritz_bounds() {
    const int m = kplusp_;
    T* wl = wrk();  // workl(1:m*m) Schur form, workl(m*m+1:..) trevc workspace
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) wl[(size_t)j * m + i] = H(i + 1, j + 1);
    T* bnd = bounds();
    for (int j = 0; j < m - 1; ++j) bnd[j] = T(0);
    bnd[m - 1] = T(1);
    int ierr = L::lahqr(true, true, m, 1, m, wl, m, ritzr(), ritzi(), 1, 1, bnd, 1);
    if (ierr != 0) return ierr;
    int sel[1] = {0}, mout = 0;
    T vl[1];
    ierr = L::trevc("R", "A", sel, m, wl, m, vl, m, &Q(1, 1), ldq_, m, &mout, wl + (size_t)m * m);
    if (ierr != 0) return ierr;
    // scale each (complex) eigenvector to unit Euclidean norm (dneigh.f:227-257)
    int iconj = 0;
    for (int i = 1; i <= m; ++i) {
      if (std::fabs(ritzi()[i - 1]) <= T(0)) {
        const T t = L::nrm2(m, &Q(1, i), 1);
        for (int r = 1; r <= m; ++r) Q(r, i) *= T(1) / t;
      } else if (iconj == 0) {
        const T t = L::lapy2(L::nrm2(m, &Q(1, i), 1), L::nrm2(m, &Q(1, i + 1), 1));
        for (int r = 1; r <= m; ++r) { Q(r, i) *= T(1) / t; Q(r, i + 1) *= T(1) / t; }
        iconj = 1;
      } else {
        iconj = 0;
      }
    }
    L::gemvT(m, m, &Q(1, 1), ldq_, bnd, wl);
    iconj = 0;
    for (int i = 1; i <= m; ++i) {
      if (std::fabs(ritzi()[i - 1]) <= T(0)) {
        bnd[i - 1] = rnorm_ * std::fabs(wl[i - 1]);
      } else if (iconj == 0) {
        bnd[i - 1] = rnorm_ * L::lapy2(wl[i - 1], wl[i]);
        bnd[i] = bnd[i - 1];
        iconj = 1;
      } else {
        iconj = 0;
      }
    }
    return 0;
  }

  // dngets: wanted values to the end; a conjugate pair is never split (kev may grow by one)
  static void select_wanted(Key which, int ishift, int& kev, int& np, T* rr, T* ri, T* bnd) {
    Key pre = Key::NONE;
    switch (which) {
      case Key::LM: pre = Key::LR; break;
      case Key::SM: pre = Key::SR; break;
      case Key::LR: pre = Key::LM; break;
      case Key::SR: pre = Key::SM; break;
      case Key::LI: pre = Key::LM; break;
      case Key::SI: pre = Key::SM; break;
      default: break;
    }
    sort_cplx(pre, kev + np, rr, ri, bnd);
    sort_cplx(which, kev + np, rr, ri, bnd);
    if (np >= 1 && (rr[np] - rr[np - 1]) == T(0) && (ri[np] + ri[np - 1]) == T(0)) {
      np--;
      kev++;
    }
    if (ishift == 1) sort_cplx(Key::SR, np, bnd, rr, ri);
  }

  int count_converged(int cnt2, const T* rr, const T* ri, const T* bnd) const {
    const T eps23 = eps23_of<T>(eps_, false);  // dnconv.f:128-129
    int nc = 0;
    for (int i = 0; i < cnt2; ++i)
      if (bnd[i] <= tol_ * std::max(eps23, L::lapy2(rr[i], ri[i]))) ++nc;
    return nc;
  }

  // Apply the np shifts to H (upper Hessenberg), accumulating Q; returns the possibly grown kev.
  int shift_sweeps(int kev, int np, const T* shiftr, const T* shifti) {
    const int kp = kev + np;
    T* wl = wrk();
    for (int j = 1; j <= kp; ++j)
      for (int i = 1; i <= kp; ++i) Q(i, j) = (i == j) ? T(1) : T(0);
    if (np == 0) return kev;
    bool cconj = false;
    for (int jj = 1; jj <= np; ++jj) {
      const T sigr = shiftr[jj - 1], sigi = shifti[jj - 1];
      // conjugate pairs are applied together or not at all (dnapps.f:283-310)
      if (cconj) { cconj = false; continue; }
      if (jj < np && std::fabs(sigi) > T(0)) cconj = true;
      else if (jj == np && std::fabs(sigi) > T(0)) { kev++; continue; }
      int istart = 1, iend;
      do {
        iend = kp;
        for (int i = istart; i <= kp - 1; ++i) {
          T tst1 = std::fabs(H(i, i)) + std::fabs(H(i + 1, i + 1));
          if (tst1 == T(0)) tst1 = L::lanhs1(kp - jj + 1, &H(1, 1), ldh_, wl);
          if (std::fabs(H(i + 1, i)) <= std::max(ulp_ * tst1, smlnum_)) {
            iend = i;
            H(i + 1, i) = T(0);
            break;
          }
        }
        const bool skip = (istart == iend) || (istart + 1 == iend && std::fabs(sigi) > T(0));
        if (!skip) {
          const T h11 = H(istart, istart), h21 = H(istart + 1, istart);
          if (std::fabs(sigi) <= T(0)) {
            // real shift: Givens rotations chase the bulge (dnapps.f:382-447)
            T f = h11 - sigr, g = h21, c, s, r;
            for (int i = istart; i <= iend - 1; ++i) {
              L::lartg(f, g, c, s, r);
              if (i > istart) {
                if (r < T(0)) { r = -r; c = -c; s = -s; }
                H(i, i - 1) = r;
                H(i + 1, i - 1) = T(0);
              }
              for (int j = i; j <= kp; ++j) {
                const T t = c * H(i, j) + s * H(i + 1, j);
                H(i + 1, j) = -s * H(i, j) + c * H(i + 1, j);
                H(i, j) = t;
              }
              for (int j = 1; j <= std::min(i + 2, iend); ++j) {
                const T t = c * H(j, i) + s * H(j, i + 1);
                H(j, i + 1) = -s * H(j, i) + c * H(j, i + 1);
                H(j, i) = t;
              }
              for (int j = 1; j <= std::min(i + jj, kp); ++j) {
                const T t = c * Q(j, i) + s * Q(j, i + 1);
                Q(j, i + 1) = -s * Q(j, i) + c * Q(j, i + 1);
                Q(j, i) = t;
              }
              if (i < iend - 1) { f = H(i + 1, i); g = H(i + 2, i); }
            }
          } else {
            // complex conjugate pair: Francis double-shift with 3x3 Householder reflectors (dnapps.f:459-523)
            const T h12 = H(istart, istart + 1), h22 = H(istart + 1, istart + 1), h32 = H(istart + 2, istart + 1);
            const T s2 = T(2.0f) * sigr, t = L::lapy2(sigr, sigi);
            T u[3], tau;
            u[0] = (h11 * (h11 - s2) + t * t) / h21 + h12;
            u[1] = h11 + h22 - s2;
            u[2] = h32;
            for (int i = istart; i <= iend - 1; ++i) {
              const int nr = std::min(3, iend - i + 1);
              L::larfg(nr, u[0], u + 1, 1, tau);
              if (i > istart) {
                H(i, i - 1) = u[0];
                H(i + 1, i - 1) = T(0);
                if (i < iend - 1) H(i + 2, i - 1) = T(0);
              }
              u[0] = T(1);
              L::larf("L", nr, kp - i + 1, u, 1, tau, &H(i, i), ldh_, wl);
              L::larf("R", std::min(i + 3, iend), nr, u, 1, tau, &H(1, i), ldh_, wl);
              L::larf("R", kp, nr, u, 1, tau, &Q(1, i), ldq_, wl);
              if (i < iend - 1) {
                u[0] = H(i + 1, i);
                u[1] = H(i + 2, i);
                if (i < iend - 2) u[2] = H(i + 3, i);
              }
            }
          }
        }
        istart = iend + 1;
      } while (iend < kp);
    }
    // non-negative sub-diagonal in the leading kev block (dnapps.f:552-558)
    for (int j = 1; j <= kev; ++j) {
      if (H(j + 1, j) < T(0)) {
        for (int c2 = j; c2 <= kp; ++c2) H(j + 1, c2) = -H(j + 1, c2);
        for (int r2 = 1; r2 <= std::min(j + 2, kp); ++r2) H(r2, j + 1) = -H(r2, j + 1);
        for (int r2 = 1; r2 <= std::min(j + np + 1, kp); ++r2) Q(r2, j + 1) = -Q(r2, j + 1);
      }
    }
    for (int i = 1; i <= kev; ++i) {
      T tst1 = std::fabs(H(i, i)) + std::fabs(H(i + 1, i + 1));
      if (tst1 == T(0)) tst1 = L::lanhs1(kev, &H(1, 1), ldh_, wl);
      if (H(i + 1, i) <= std::max(ulp_ * tst1, smlnum_)) H(i + 1, i) = T(0);
    }
    return kev;
  }

  bool run() {
    CO_BEGIN(pc_)
    this->gv_itry_ = 1; this->gv_initv_ = initv_; this->gv_j_ = 1;
    CO_CALL(pc_, this->start_vector());
    if (rnorm_ == T(0)) {  // dnaup2.f:324-331 (goes through label 1100: nev = numcnv, mxiter = iter)
      info_ = -9;
      mxiter_out_ = iter_;
      np_ = np0_;
      CO_END_EARLY(pc_);
    }
    this->ai_k_ = 0; this->ai_np_ = nev_;
    CO_CALL(pc_, this->extend());
    if (this->ai_info_ > 0) { fail_no_factorisation(); CO_END_EARLY(pc_); }
    for (;;) {
      iter_++;
      np_ = kplusp_ - nev_;  // dnaup2.f:401
      if (trace_levels().mnaup2 > 0 && ops_->rank() == 0) {  // dnaup2.f:390-408
        trace::ivout1(iter_, "_naup2: **** Start of major iteration number ****");
        if (trace_levels().mnaup2 > 1) {
          trace::ivout1(nev_, "_naup2: The length of the current Arnoldi factorization");
          trace::ivout1(np_, "_naup2: Extend the Arnoldi factorization by");
        }
      }
      this->ai_k_ = nev_; this->ai_np_ = np_;
      CO_CALL(pc_, this->extend());
      if (this->ai_info_ > 0) { fail_no_factorisation(); CO_END_EARLY(pc_); }
      if (ritz_bounds() != 0) {
        info_ = -8;
        mxiter_out_ = mxiter_;
        CO_END_EARLY(pc_);
      }
      std::copy(ritzr(), ritzr() + kplusp_, wrk() + kplusp_ * kplusp_);
      std::copy(ritzi(), ritzi() + kplusp_, wrk() + kplusp_ * kplusp_ + kplusp_);
      std::copy(bounds(), bounds() + kplusp_, wrk() + kplusp_ * kplusp_ + 2 * kplusp_);
      nev_ = nev0_; np_ = np0_;
      numcnv_ = nev_;
      select_wanted(which_, ishift_, nev_, np_, ritzr(), ritzi(), bounds());
      if (nev_ == nev0_ + 1) numcnv_ = nev0_ + 1;
      std::copy(bounds() + np_, bounds() + np_ + nev_, wrk() + 2 * np_);
      nconv_ = count_converged(nev_, ritzr() + np_, ritzi() + np_, wrk() + 2 * np_);
      if (trace_levels().mnaup2 > 2 && ops_->rank() == 0) {  // dnaup2.f:492-505
        const int kp[4] = {nev_, np_, numcnv_, nconv_};
        trace::ivout(4, kp, "_naup2: NEV, NP, NUMCNV, NCONV are");
        trace::dvout(kplusp_, ritzr(), "_naup2: Real part of the eigenvalues of H");
        trace::dvout(kplusp_, ritzi(), "_naup2: Imaginary part of the eigenvalues of H");
        trace::dvout(kplusp_, bounds(), "_naup2: Ritz estimates of the current NCV Ritz values");
      }
      {
        const int nptemp = np_;
        for (int j = 0; j < nptemp; ++j)
          if (bounds()[j] == T(0)) { np_--; nev_++; }
      }
      if (nconv_ >= numcnv_ || iter_ > mxiter_ || np_ == 0) {
        finish_sorted();
        break;
      } else if (nconv_ < numcnv_ && ishift_ == 1) {
        const int nevbef = nev_;
        nev_ += std::min(nconv_, np_ / 2);
        if (nev_ == 1 && kplusp_ >= 6) nev_ = kplusp_ / 2;
        else if (nev_ == 1 && kplusp_ > 3) nev_ = 2;
        if (!par_ && nev_ > kplusp_ - 2) nev_ = kplusp_ - 2;  // dnaup2.f:667-676; absent from pdnaup2.f
        np_ = kplusp_ - nev_;
        if (nevbef < nev_) select_wanted(which_, ishift_, nev_, np_, ritzr(), ritzi(), bounds());
      }
      if (ishift_ == 0) {
        ido_ = 3;
        CO_YIELD(pc_);
        std::copy(wrk(), wrk() + np_, ritzr());
        std::copy(wrk() + np_, wrk() + 2 * np_, ritzi());
      }
      // implicit restart; kev may grow to keep a conjugate pair together (dnaup2.f:762)
      nev_ = shift_sweeps(nev_, np_, ritzr(), ritzi());
      {
        const T sigmak = Q(kplusp_, nev_), betak = H(nev_ + 1, nev_);
        const bool has_beta = betak > T(0);
        ops_->vq_update(n_, kplusp_, nev_ + (has_beta ? 1 : 0), v_, ldv_, wl_ + iq_, ldq_, true, sigmak,
                        has_beta ? betak : T(0), has_beta ? nev_ : -1, resid_, bmat_ == 'I' ? mbC() : nullptr);
      }
      if (bmat_ == 'G') {
        cnt.nbx++;
        ops_->copy(n_, resid_, this->slot(n_ + 1));
        ipntr_[0] = n_ + 1; ipntr_[1] = 1;
        ido_ = 2;
        CO_YIELD(pc_);
        ops_->dot(n_, resid_, this->slot(1), mbC());
      }
      rnorm_ = this->fetch_norm_from_dot(mbC());
    }
    CO_END(pc_)
  }

  void fail_no_factorisation() {
    np_ = this->ai_info_;
    mxiter_out_ = iter_;
    info_ = -9999;
  }

  // exit ordering of (ritzr, ritzi, bounds) (dnaup2.f:552-650)
  void finish_sorted() {
    H(3, 1) = rnorm_;  // for eupd
    T *rr = ritzr(), *ri = ritzi(), *b = bounds();
    Key w1 = Key::NONE, w2 = Key::NONE;
    switch (which_) {
      case Key::LM: w1 = Key::SR; w2 = Key::SM; break;
      case Key::SM: w1 = Key::LR; w2 = Key::LM; break;
      case Key::LR: w1 = Key::SM; w2 = Key::SR; break;
      case Key::SR: w1 = Key::LM; w2 = Key::LR; break;
      case Key::LI: w1 = Key::SM; w2 = Key::SI; break;
      case Key::SI: w1 = Key::LM; w2 = Key::LI; break;
      default: break;
    }
    sort_cplx(w1, kplusp_, rr, ri, b);
    sort_cplx(w2, kplusp_, rr, ri, b);
    for (int j = 0; j < numcnv_; ++j) b[j] /= std::max(eps23_, L::lapy2(rr[j], ri[j]));
    sort_cplx(Key::LR, numcnv_, b, rr, ri);
    for (int j = 0; j < numcnv_; ++j) b[j] *= std::max(eps23_, L::lapy2(rr[j], ri[j]));
    sort_cplx(which_, nconv_, rr, ri, b);
    if (iter_ > mxiter_ && nconv_ < numcnv_) info_ = 1;
    if (np_ == 0 && nconv_ < numcnv_) info_ = 2;
    np_ = nconv_;
    mxiter_out_ = iter_;
    nev_ = numcnv_;
  }
};

}  // namespace ab200
