// irl_sym.hpp -- implicitly restarted Lanczos: the dsaupd/dseupd (ssaupd/sseupd) entry points.
//
//   aupd()        <->  SRC/dsaupd.f:408-690 + SRC/dsaup2.f:179-851   (pdsaupd.f / pdsaup2.f when parpack)
//   ritz_bounds() <->  SRC/dseigt.f:87-181 (+ dstqrb.f): LAPACK xSTEQR('I') and its last row
//   select()      <->  SRC/dsgets.f:93-219,  count_converged() <-> SRC/dsconv.f:59-138
//   restart()     <->  SRC/dsapps.f:131-518: QR sweeps on the host, V <- V*Q on the device
//   eupd()        <->  SRC/dseupd.f:218-867 (pdseupd.f)
#pragma once
#include "irl_base.hpp"

namespace ab200 {

template <typename T>
class IrlSym : public IrlBase<T> {
  using B = IrlBase<T>;
  using B::ops_; using B::par_; using B::n_; using B::ncv_; using B::bmat_; using B::mode_; using B::resid_;
  using B::v_; using B::ldv_; using B::workd_; using B::ido_; using B::ipntr_; using B::rnorm_; using B::cnt;
  using B::mbC; using B::hC;
  using L = Lapack<T>;

 public:
  IrlSym(VecOps<T>* ops, bool parpack, SeedState* seed) : B(ops, parpack) { this->seed_ = seed; }

  // effective tolerance of the last aupd call (the *_c entry points pass tol by value, so the
  // "tol <= 0 -> eps" substitution of dsaupd.f:550 is invisible to the caller)
  T tol_effective = 0;

  // One reverse-communication call.  Device pointers only; workl/iparam/ipntr are host arrays.
  void aupd(int* ido, char bmat, int n, const char* which, int nev, T* tol, T* resid_dev, int ncv, T* v_dev,
            int64_t ldv, int* iparam, int* ipntr, T* workd_dev, T* workl, int lworkl, int* info) {
    if (*ido == 0) {
      cnt = Counters();  // dstats (dsaupd.f:480)
      int ierr = 0;
      ishift_ = iparam[0];
      mxiter_ = iparam[2];
      mode_ = iparam[6];
      if (n <= 0) ierr = -1;
      else if (nev <= 0) ierr = -2;
      else if (ncv <= nev || (!par_ && ncv > n)) ierr = -3;
      if (mxiter_ <= 0) ierr = -4;
      which_ = key_of(which);
      be_ = (which[0] == 'B' && which[1] == 'E');
      if (!be_ && which_ != Key::LM && which_ != Key::SM && which_ != Key::LA && which_ != Key::SA) ierr = -5;
      if (bmat != 'I' && bmat != 'G') ierr = -6;
      if (lworkl < ncv * ncv + 8 * ncv) ierr = -7;
      if (mode_ < 1 || mode_ > 5) ierr = -10;
      else if (mode_ == 1 && bmat == 'G') ierr = -11;
      else if (ishift_ < 0 || ishift_ > 1) ierr = -12;
      else if (nev == 1 && be_) ierr = -13;
      if (ierr != 0) {
        *info = ierr;
        *ido = 99;
        return;
      }
      if (*tol <= T(0)) *tol = L::lamch("E");
      n_ = n; ncv_ = ncv; bmat_ = bmat;
      resid_ = resid_dev; v_ = v_dev; ldv_ = ldv; workd_ = workd_dev;
      nev0_ = nev; np0_ = ncv - nev; nev_ = nev0_; np_ = np0_; kplusp_ = ncv;
      std::fill(workl, workl + (size_t)ncv * ncv + 8 * (size_t)ncv, T(0));
      // workl partition (dsaupd.f:582-595), 0-based offsets
      ldh_ = ncv; ldq_ = ncv;
      ih_ = 0; iritz_ = ih_ + 2 * ldh_; ibounds_ = iritz_ + ncv; iq_ = ibounds_ + ncv; iw_ = iq_ + ncv * ncv;
      ipntr[3] = iw_ + 3 * ncv + 1;
      ipntr[4] = ih_ + 1; ipntr[5] = iritz_ + 1; ipntr[6] = ibounds_ + 1; ipntr[10] = iw_ + 1;
      this->setup_mailbox();
      eps_ = L::lamch("E");
      eps23_ = eps23_of<T>(eps_, par_);
      safmin_ = L::lamch("S");
      nconv_ = 0; iter_ = 0;
      initv_ = (*info != 0);  // dsaup2.f:306-316
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
    this->phase_clock.dump("dsaupd");
    iparam[2] = mxiter_out_;
    iparam[4] = np_;
    iparam[8] = cnt.nopx; iparam[9] = cnt.nbx; iparam[10] = cnt.nrorth;
    *info = info_;
    if (*info == 2) *info = 3;
    if (*info >= 0 && trace_levels().msaupd > 0 && ops_->rank() == 0) {  // dsaupd.f:630-680
      trace::ivout1(mxiter_out_, "_saupd: number of update iterations taken");
      trace::ivout1(np_, "_saupd: number of \"converged\" Ritz values");
      trace::dvout(np_, ritz(), "_saupd: final Ritz values");
      trace::dvout(np_, bounds(), "_saupd: corresponding error bounds");
      trace::summary("Symmetric implicit Arnoldi update code", mxiter_out_, cnt.nopx, cnt.nbx, cnt.nrorth, cnt.nitref,
                     cnt.nrstrt);
    }
  }

  int kplusp() const { return kplusp_; }

  // ---------------------------------------------------------------------------------------------
  // dseupd: eigenvalues in d (host), Ritz vectors in z (device, may alias v)
  // ---------------------------------------------------------------------------------------------
  void eupd(bool rvec, char howmny, int* select, T* d, T* z_dev, int64_t ldz, T sigma, char bmat, int n,
            const char* which, int nev, T tol, T* resid_dev, int ncv, T* v_dev, int64_t ldv, int* iparam,
            int* ipntr, T* workd_dev, T* workl, int lworkl, int* info) {
    const int mode = iparam[6];
    int nconv = iparam[4];
    *info = 0;
    if (nconv == 0) return;
    int ierr = 0;
    const Key wk = key_of(which);
    const bool be = (which[0] == 'B' && which[1] == 'E');
    if (nconv <= 0) ierr = -14;
    if (n <= 0) ierr = -1;
    if (nev <= 0) ierr = -2;
    if (ncv <= nev || (!par_ && ncv > n)) ierr = -3;
    if (!be && wk != Key::LM && wk != Key::SM && wk != Key::LA && wk != Key::SA) ierr = -5;
    if (bmat != 'I' && bmat != 'G') ierr = -6;
    if ((howmny != 'A' && howmny != 'P' && howmny != 'S') && rvec) ierr = -15;
    if (rvec && howmny == 'S') ierr = -16;
    if (rvec && lworkl < ncv * ncv + 8 * ncv) ierr = -7;
    enum { REGULR, SHIFTI, BUCKLE, CAYLEY } type = REGULR;
    if (mode == 1 || mode == 2) type = REGULR;
    else if (mode == 3) type = SHIFTI;
    else if (mode == 4) type = BUCKLE;
    else if (mode == 5) type = CAYLEY;
    else ierr = -10;
    if (mode == 1 && bmat == 'G') ierr = -11;
    if (nev == 1 && be) ierr = -12;
    if (ierr != 0) { *info = ierr; return; }

    // workl layout after dsaupd (dseupd.f:356-423); 0-based offsets
    const int ih = ipntr[4] - 1, ritz = ipntr[5] - 1, bounds = ipntr[6] - 1;
    const int ldh = ncv, ldq = ncv;
    const int ihd = bounds + ldh, ihb = ihd + ldh, iq = ihb + ldh, iw = iq + ldh * ncv;
    ipntr[3] = iw + 2 * ncv + 1;
    ipntr[7] = ihd + 1; ipntr[8] = ihb + 1; ipntr[9] = iq + 1;
    const int irz = ipntr[10] - 1 + ncv, ibd = irz + ncv;
    T* W = workl;
    const T eps23 = eps23_of<T>(L::lamch("E"), par_);
    const T rnorm = W[ih];  // smuggled by aupd (dsaup2.f:645)
    T bnorm2 = rnorm;
    if (bmat == 'G') {
      ops_->dot(n, workd_dev, workd_dev, this->mbC());
      ops_->allreduce_sum(this->mbC(), 1);
      ops_->fetch(this->hC(), this->mbC(), 1);
      bnorm2 = std::sqrt(this->hC()[0]);
    }
    bool reord = false;
    if (rvec) {
      // which Ritz values did aupd accept?  (dseupd.f:463-525)
      for (int j = 0; j < ncv; ++j) { W[bounds + j] = T(j + 1); select[j] = 0; }
      select_wanted(wk, be, /*ishift=*/0, nev, ncv - nev, W + irz, W + bounds, nullptr);
      int numcnv = 0;
      for (int j = 1; j <= ncv; ++j) {
        const T temp1 = std::max(eps23, std::fabs(W[irz + ncv - j]));
        const int jj = (int)W[bounds + ncv - j];
        if (numcnv < nconv && W[ibd + jj - 1] <= tol * temp1) {
          select[jj - 1] = 1;
          numcnv++;
          if (jj > (par_ ? nev : nconv)) reord = true;
        }
      }
      if (numcnv != nconv) { *info = -17; return; }
      // eigen-decomposition of the final tridiagonal (dseupd.f:533-542)
      std::copy(W + ih + 1, W + ih + ncv, W + ihb);
      std::copy(W + ih + ldh, W + ih + ldh + ncv, W + ihd);
      if (L::steqr_I(ncv, W + ihd, W + ihb, W + iq, ldq, W + iw) != 0) { *info = -8; return; }
      if (reord) {
        // two-pointer partition: selected pairs to the front (dseupd.f:563-610)
        int left = 0, right = ncv - 1;
        while (left < right) {
          if (select[left]) ++left;
          else if (!select[right]) --right;
          else {
            std::swap(W[ihd + left], W[ihd + right]);
            std::swap_ranges(W + iq + (size_t)ncv * left, W + iq + (size_t)ncv * (left + 1),
                             W + iq + (size_t)ncv * right);
            ++left; --right;
          }
        }
      }
      std::copy(W + ihd, W + ihd + nconv, d);
    } else {
      std::copy(W + ritz, W + ritz + nconv, d);
      std::copy(W + ritz, W + ritz + ncv, W + ihd);
    }
    // spectral back-transformation of the Ritz values (dseupd.f:642-714)
    if (type == REGULR) {
      if (rvec) sort_real_cols(Key::LA, nconv, d, ncv, W + iq, ldq);
      else std::copy(W + bounds, W + bounds + ncv, W + ihb);
    } else {
      std::copy(W + ihd, W + ihd + ncv, W + iw);
      for (int k = 0; k < ncv; ++k) {
        T& th = W[ihd + k];
        if (type == SHIFTI) th = T(1) / th + sigma;
        else if (type == BUCKLE) th = sigma * th / (th - T(1));
        else th = sigma * (th + T(1)) / (th - T(1));
      }
      std::copy(W + ihd, W + ihd + nconv, d);
      sort_real(Key::LA, nconv, W + ihd, W + iw);
      if (rvec) {
        sort_real_cols(Key::LA, nconv, d, ncv, W + iq, ldq);
      } else {
        std::copy(W + bounds, W + bounds + ncv, W + ihb);
        for (int k = 0; k < ncv; ++k) W[ihb + k] *= bnorm2 / rnorm;
        sort_real(Key::LA, nconv, d, W + ihb);
      }
    }
    if (rvec && howmny == 'A') {
      // QR of the wanted eigenvectors of H; V <- V*Q on the device; Z = first nconv columns
      // (dseupd.f:730-746: dgeqr2 + dorm2r + dlacpy).  Q is formed explicitly (ncv x ncv, host) so
      // that the n-length work is one tall-skinny pass.
      L::geqr2(ncv, nconv, W + iq, ldq, W + iw + ncv, W + ihb);
      std::vector<T> qfull((size_t)ncv * ncv, T(0)), wk2((size_t)ncv);
      for (int i = 0; i < ncv; ++i) qfull[(size_t)i * ncv + i] = T(1);
      L::orm2r("R", "N", ncv, ncv, nconv, W + iq, ldq, W + iw + ncv, qfull.data(), ncv, wk2.data());
      ops_->vq_update(n, ncv, ncv, v_dev, ldv, qfull.data(), ncv, false, T(0), T(0), 0, nullptr, nullptr);
      if (z_dev != v_dev) ops_->copy2d(n, nconv, v_dev, ldv, z_dev, ldz);
      // last row of the eigenvector matrix for the error bounds (dseupd.f:754-771)
      for (int j = 0; j < ncv - 1; ++j) W[ihb + j] = T(0);
      W[ihb + ncv - 1] = T(1);
      T tmp;
      L::orm2r("L", "T", ncv, 1, nconv, W + iq, ldq, W + iw + ncv, W + ihb, ncv, &tmp);
      if (!par_)
        for (int j = 0; j < nconv; ++j) W[iw + ncv + j] = W[ihb + j];
    }
    if (type == REGULR && rvec) {
      for (int j = 0; j < ncv; ++j) W[ihb + j] = rnorm * std::fabs(W[ihb + j]);
    } else if (type != REGULR && rvec) {
      for (int k = 0; k < ncv; ++k) {
        T& b = W[ihb + k];
        b *= bnorm2;
        const T th = W[iw + k];
        if (type == SHIFTI) b = std::fabs(b) / (th * th);
        else if (type == BUCKLE) b = sigma * std::fabs(b) / ((th - T(1)) * (th - T(1)));
        else b = std::fabs(b / th * (th - T(1)));
      }
    }
    // eigenvector purification (dseupd.f:840-857)
    if (rvec && (type == SHIFTI || type == CAYLEY)) {
      for (int k = 0; k < nconv; ++k) W[iw + k] = (par_ ? W[iq + k * ldq + ncv - 1] : W[iw + ncv + k]) / W[iw + k];
    } else if (rvec && type == BUCKLE) {
      for (int k = 0; k < nconv; ++k)
        W[iw + k] = (par_ ? W[iq + k * ldq + ncv - 1] : W[iw + ncv + k]) / (W[iw + k] - T(1));
    }
    // pdseupd.f:858 applies this rank-1 update unconditionally -- with rvec = 0 it scribbles uninitialised
    // coefficients over a Z the caller never asked for.  Deliberate deviation: without rvec nothing is written
    // (z is not even mapped then), exactly as the sequential dseupd.f:840-857 behaves.
    if (rvec && z_dev != nullptr && type != REGULR) ops_->ger(n, nconv, resid_dev, W + iw, z_dev, ldz);
  }

 private:
  // ---- state ----
  int pc_ = 0;
  int ishift_ = 1, mxiter_ = 0, mxiter_out_ = 0;
  Key which_ = Key::NONE;
  bool be_ = false;
  int nev0_ = 0, np0_ = 0, nev_ = 0, np_ = 0, kplusp_ = 0, nconv_ = 0, iter_ = 0, info_ = 0;
  bool initv_ = false;
  int ldh_ = 0, ldq_ = 0, ih_ = 0, iritz_ = 0, ibounds_ = 0, iq_ = 0, iw_ = 0;
  T* wl_ = nullptr;
  T tol_ = 0, eps_ = 0, eps23_ = 0, safmin_ = 0;
  T sigmak_ = 0, betak_ = 0;

  T& H(int i, int j) { return wl_[ih_ + (i - 1) + (size_t)(j - 1) * ldh_]; }  // 1-based; (.,1)=sub-diag (.,2)=diag
  T* ritz() { return wl_ + iritz_; }
  T* bounds() { return wl_ + ibounds_; }
  T& Q(int i, int j) { return wl_[iq_ + (i - 1) + (size_t)(j - 1) * ldq_]; }
  T* wrk() { return wl_ + iw_; }

  // ---- hooks of the step extension ----
  void h_store(int j, const T* hcol, T beta, bool after_restart) override {
    H(j, 2) = hcol[j - 1];
    H(j, 1) = (j == 1 || after_restart) ? T(0) : beta;
  }
  void h_add(int j, const T* scol, bool after_restart) override {
    if (j == 1 || after_restart) H(j, 1) = T(0);
    H(j, 2) += scol[j - 1];
  }
  void sweep_done(int, int) override {}
  int aitr_trace_level() const override { return trace_levels().msaitr; }
  T tiny_norm() override { return safmin_; }
  bool mode2_shortcut() const override { return mode_ == 2; }

  // Ritz values of the kplusp x kplusp tridiagonal and their error bounds rnorm*|last row|
  int ritz_bounds() {
    const int m = kplusp_;
    std::vector<T> e((size_t)std::max(1, m)), z((size_t)m * m), work((size_t)std::max(1, 2 * m - 2));
    for (int i = 0; i < m; ++i) ritz()[i] = H(i + 1, 2);
    for (int i = 0; i < m - 1; ++i) e[i] = H(i + 2, 1);
    if (m == 1) {
      bounds()[0] = rnorm_;
      return 0;
    }
    const int ierr = L::steqr_I(m, ritz(), e.data(), z.data(), m, work.data());
    if (ierr != 0) return ierr;
    for (int k = 0; k < m; ++k) bounds()[k] = rnorm_ * std::fabs(z[(size_t)k * m + (m - 1)]);
    return 0;
  }

  // wanted values to the END of ritz (dsgets.f); with exact shifts the unwanted ones are re-sorted
  // so that those with the largest bounds come first
  static void select_wanted(Key which, bool be, int ishift, int kev, int np, T* ritz, T* bounds, T* shifts) {
    if (be) {
      sort_real(Key::LA, kev + np, ritz, bounds);
      const int kevd2 = kev / 2;
      if (kev > 1) {
        const int cnt = std::min(kevd2, np), off = std::max(kevd2, np);
        std::swap_ranges(ritz, ritz + cnt, ritz + off);
        std::swap_ranges(bounds, bounds + cnt, bounds + off);
      }
    } else {
      sort_real(which, kev + np, ritz, bounds);
    }
    if (ishift == 1 && np > 0) {
      sort_real(Key::SM, np, bounds, ritz);
      if (shifts) std::copy(ritz, ritz + np, shifts);
    }
  }

  int count_converged(int cnt, const T* ritz, const T* bnd) const {
    // dsconv.f:111-125 always uses the DOUBLE exponent 2/3, also under PARPACK
    const T eps23 = eps23_of<T>(eps_, false);
    int nc = 0;
    for (int i = 0; i < cnt; ++i)
      if (bnd[i] <= tol_ * std::max(eps23, std::fabs(ritz[i]))) ++nc;
    return nc;
  }

  // implicit QR sweeps with the np shifts on the tridiagonal, accumulating Q (dsapps.f:226-442)
  void qr_sweeps(int kev, int np, const T* shift) {
    const int kp = kev + np;
    for (int j = 1; j <= kp; ++j)
      for (int i = 1; i <= kp; ++i) Q(i, j) = (i == j) ? T(1) : T(0);
    if (np == 0) return;
    int itop = 1;
    for (int jj = 1; jj <= np; ++jj) {
      int istart = itop, iend;
      do {
        // split at a negligible sub-diagonal
        iend = kp;
        for (int i = istart; i <= kp - 1; ++i) {
          const T big = std::fabs(H(i, 2)) + std::fabs(H(i + 1, 2));
          if (H(i + 1, 1) <= eps_ * big) {
            H(i + 1, 1) = T(0);
            iend = i;
            break;
          }
        }
        if (istart < iend) {
          T c, s, r;
          auto rotate = [&](int i) {
            // symmetric 2x2 similarity on rows/cols i, i+1 of the tridiagonal + columns of Q
            const T a1 = c * H(i, 2) + s * H(i + 1, 1);
            const T a2 = c * H(i + 1, 1) + s * H(i + 1, 2);
            const T a4 = c * H(i + 1, 2) - s * H(i + 1, 1);
            const T a3 = c * H(i + 1, 1) - s * H(i, 2);
            H(i, 2) = c * a1 + s * a2;
            H(i + 1, 2) = c * a4 - s * a3;
            H(i + 1, 1) = c * a3 + s * a4;
            const int jmax = std::min(i + jj, kp);
            for (int j = 1; j <= jmax; ++j) {
              const T q1 = c * Q(j, i) + s * Q(j, i + 1);
              Q(j, i + 1) = -s * Q(j, i) + c * Q(j, i + 1);
              Q(j, i) = q1;
            }
          };
          L::lartg(H(istart, 2) - shift[jj - 1], H(istart + 1, 1), c, s, r);
          rotate(istart);
          for (int i = istart + 1; i <= iend - 1; ++i) {
            // chase the bulge
            const T f = H(i, 1), g = s * H(i + 1, 1);
            H(i + 1, 1) = c * H(i + 1, 1);
            L::lartg(f, g, c, s, r);
            if (r < T(0)) { r = -r; c = -c; s = -s; }
            H(i, 1) = r;
            rotate(i);
          }
        }
        istart = iend + 1;
        if (H(iend, 1) < T(0)) {
          H(iend, 1) = -H(iend, 1);
          for (int i = 1; i <= kp; ++i) Q(i, iend) = -Q(i, iend);
        }
      } while (iend < kp);
      for (int i = itop; i <= kp - 1; ++i) {
        if (H(i + 1, 1) > T(0)) break;
        itop++;
      }
    }
    for (int i = itop; i <= kp - 1; ++i) {
      const T big = std::fabs(H(i, 2)) + std::fabs(H(i + 1, 2));
      if (H(i + 1, 1) <= eps_ * big) H(i + 1, 1) = T(0);
    }
  }

  // ---- the restart loop (dsaup2.f) as a resumable routine ----
  bool run() {
    CO_BEGIN(pc_)
    this->gv_itry_ = 1; this->gv_initv_ = initv_; this->gv_j_ = 1;
    CO_CALL(pc_, this->start_vector());
    if (rnorm_ == T(0)) {  // dsaup2.f:332-340
      info_ = -9;
      mxiter_out_ = mxiter_;
      np_ = np0_;
      CO_END_EARLY(pc_);
    }
    this->ai_k_ = 0; this->ai_np_ = nev0_;
    CO_CALL(pc_, this->extend());
    if (this->ai_info_ > 0) { fail_no_factorisation(); CO_END_EARLY(pc_); }
    for (;;) {
      iter_++;
      if (trace_levels().msaup2 > 0 && ops_->rank() == 0) {  // dsaup2.f:404-413
        trace::ivout1(iter_, "_saup2: **** Start of major iteration number ****");
        if (trace_levels().msaup2 > 1) {
          trace::ivout1(nev_, "_saup2: The length of the current Lanczos factorization");
          trace::ivout1(np_, "_saup2: Extend the Lanczos factorization by");
        }
      }
      this->ai_k_ = nev_; this->ai_np_ = np_;
      CO_CALL(pc_, this->extend());
      if (this->ai_info_ > 0) { fail_no_factorisation(); CO_END_EARLY(pc_); }
      int rb_ierr;
      {
        PhaseClock::Scope pcs(&this->phase_clock, PhaseClock::PROJECTED);
        rb_ierr = ritz_bounds();
      }
      if (rb_ierr != 0) {
        info_ = -8;
        mxiter_out_ = mxiter_;
        CO_END_EARLY(pc_);
      }
      std::copy(ritz(), ritz() + kplusp_, wrk() + kplusp_);
      std::copy(bounds(), bounds() + kplusp_, wrk() + 2 * kplusp_);
      nev_ = nev0_; np_ = np0_;
      select_wanted(which_, be_, ishift_, nev_, np_, ritz(), bounds(), wrk());
      std::copy(bounds() + np_, bounds() + np_ + nev_, wrk() + np_);
      nconv_ = count_converged(nev_, ritz() + np_, wrk() + np_);
      if (trace_levels().msaup2 > 2 && ops_->rank() == 0) {  // dsaup2.f:494-504
        const int kp[3] = {nev_, np_, nconv_};
        trace::ivout(3, kp, "_saup2: NEV, NP, NCONV are");
        trace::dvout(kplusp_, ritz(), "_saup2: The eigenvalues of H");
        trace::dvout(kplusp_, bounds(), "_saup2: Ritz estimates of the current NCV Ritz values");
      }
      {
        // shifts with a zero error bound are not applied (dsaup2.f:516-522)
        const int nptemp = np_;
        for (int j = 0; j < nptemp; ++j)
          if (bounds()[j] == T(0)) { np_--; nev_++; }
      }
      if (nconv_ >= nev0_ || iter_ > mxiter_ || np_ == 0) {
        finish_sorted();
        break;
      } else if (nconv_ < nev_ && ishift_ == 1) {
        // keep more Ritz values to avoid stagnation (dsaup2.f:677-693)
        const int nevbef = nev_;
        nev_ += std::min(nconv_, np_ / 2);
        if (nev_ == 1 && kplusp_ >= 6) nev_ = kplusp_ / 2;
        else if (nev_ == 1 && kplusp_ > 2) nev_ = 2;
        np_ = kplusp_ - nev_;
        if (nevbef < nev_) select_wanted(which_, be_, ishift_, nev_, np_, ritz(), bounds(), wrk());
      }
      if (ishift_ == 0) {  // user-supplied shifts (dsaup2.f:713-743)
        ido_ = 3;
        CO_YIELD(pc_);
        std::copy(wrk(), wrk() + np_, ritz());
      }
      // implicit restart: host QR sweeps, then V <- V*Q, r <- sigma*r + beta*v_{kev+1}, ||r|| in one pass
      {
        PhaseClock::Scope pcs(&this->phase_clock, PhaseClock::SHIFTS);
        qr_sweeps(nev_, np_, ritz());
      }
      sigmak_ = Q(kplusp_, nev_);
      betak_ = H(nev_ + 1, 1);
      {
        PhaseClock::Scope pcs(&this->phase_clock, PhaseClock::RESTART_ENQ);
        const bool has_beta = betak_ > T(0);
        ops_->vq_update(n_, kplusp_, nev_ + (has_beta ? 1 : 0), v_, ldv_, wl_ + iq_, ldq_, true, sigmak_,
                        has_beta ? betak_ : T(0), has_beta ? nev_ : -1, resid_, bmat_ == 'I' ? mbC() : nullptr);
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
    // dsaup2.f:378-390: size of the factorisation that was built goes out through iparam(5)
    np_ = this->ai_info_;
    mxiter_out_ = iter_;
    info_ = -9999;
  }

  // exit path: order Ritz values so that the wanted, converged ones lead (dsaup2.f:536-667)
  void finish_sorted() {
    T* r = ritz();
    T* b = bounds();
    if (be_) {
      sort_real(Key::SA, kplusp_, r, b);
      const int nevd2 = nev0_ / 2, nevm2 = nev0_ - nevd2;
      if (nev_ > 1) {
        if (!par_) np_ = kplusp_ - nev0_;  // dsaup2.f:552, absent from pdsaup2.f
        const int cnt2 = std::min(nevd2, np_);
        const int off = std::max(kplusp_ - nevd2 + 1, kplusp_ - np_ + 1) - 1;
        std::swap_ranges(r + nevm2, r + nevm2 + cnt2, r + off);
        std::swap_ranges(b + nevm2, b + nevm2 + cnt2, b + off);
      }
    } else {
      Key wprime = Key::NONE;
      if (which_ == Key::LM) wprime = Key::SM;
      if (which_ == Key::SM) wprime = Key::LM;
      if (which_ == Key::LA) wprime = Key::SA;
      if (which_ == Key::SA) wprime = Key::LA;
      sort_real(wprime, kplusp_, r, b);
    }
    // order the wanted ones by relative error bound, largest last
    for (int j = 0; j < nev0_; ++j) b[j] /= std::max(eps23_, std::fabs(r[j]));
    sort_real(Key::LA, nev0_, b, r);
    for (int j = 0; j < nev0_; ++j) b[j] *= std::max(eps23_, std::fabs(r[j]));
    if (be_) sort_real(Key::LA, nconv_, r, b);
    else sort_real(which_, nconv_, r, b);
    H(1, 1) = rnorm_;  // for eupd (dsaup2.f:645)
    if (iter_ > mxiter_ && nconv_ < nev_) info_ = 1;
    if (np_ == 0 && nconv_ < nev0_) info_ = 2;
    np_ = nconv_;
    mxiter_out_ = iter_;
    nev_ = nconv_;
  }
};

}  // namespace ab200
