// irl_base.hpp -- shared host control of the implicitly restarted Lanczos/Arnoldi iteration.
//
// What the reference spreads over SAVE'd flags and computed GOTOs is written
// here as resumable routines: each reverse-communication hand-off is a CO_YIELD, all state lives in
// the solver object, and every n-length operation goes through VecOps (device kernels).
//
//   start_vector()  <->  SRC/dgetv0.f:119-421, PARPACK/SRC/MPI/pdgetv0.f:131-463
//   extend()        <->  SRC/dsaitr.f:204-853 and SRC/dnaitr.f:209-840 (+ pdsaitr.f / pdnaitr.f)
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <vector>

#include "hostmath.hpp"
#include "trace.hpp"
#include "vecops.hpp"

namespace ab200 {

// resumable-routine plumbing (switch-based; no locals may live across a yield)
#define CO_BEGIN(pc) \
  switch (pc) {      \
    case 0:
#define CO_YIELD(pc)  \
  do {                \
    pc = __LINE__;    \
    return false;     \
    case __LINE__:;   \
  } while (0)
// run a nested resumable routine to completion, yielding whenever it yields
#define CO_CALL(pc, expr)      \
  do {                         \
    pc = __LINE__;             \
    case __LINE__:             \
      if (!(expr)) return false; \
  } while (0)
#define CO_END(pc) \
  }                \
  pc = 0;          \
  return true;
// leave a resumable routine for good from the middle of its body
#define CO_END_EARLY(pc) \
  do {                   \
    pc = 0;              \
    return true;         \
  } while (0)

// COMMON /timing/ counters of the reference (stat.h:11)
struct Counters {
  int nopx = 0, nbx = 0, nrorth = 0, nitref = 0, nrstrt = 0;
};

// Host-side phase clock (diagnostic, AB200_TIMING=1): where the host spends its time inside the library during a solve.
// The reference's own timers (stat.h t*) are dead in arpack-ng builds (UTIL/second_NONE.f); this is not a replacement
// for them, it only answers "is the device waiting for the host, and where".
struct PhaseClock {
  enum { FETCH_LOG, REPLAY, PROJECTED, SHIFTS, RESTART_ENQ, FETCH_NORM, STEP_ENQ, NPHASE };
  double acc[NPHASE] = {0, 0, 0, 0, 0, 0, 0};
  long long cnt[NPHASE] = {0, 0, 0, 0, 0, 0, 0};
  bool on = false;
  PhaseClock() { on = getenv("AB200_TIMING") != nullptr; }
  static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  struct Scope {
    PhaseClock* pc; int ph; double t0;
    Scope(PhaseClock* p, int phase) : pc(p->on ? p : nullptr), ph(phase), t0(pc ? now() : 0.0) {}
    ~Scope() { if (pc) { pc->acc[ph] += now() - t0; pc->cnt[ph]++; } }
  };
  void dump(const char* what) const {
    if (!on) return;
    static const char* nm[NPHASE] = {"fetch_log(blocked)", "replay", "projected_eig", "select+shifts+qr", "restart_enqueue",
                                     "fetch_norm(blocked)", "step_enqueue"};
    std::fprintf(stderr, "arpack_b200 timing [%s]:", what);
    for (int i = 0; i < NPHASE; ++i) std::fprintf(stderr, " %s=%.3fms/%lld", nm[i], 1e3 * acc[i], cnt[i]);
    std::fprintf(stderr, "\n");
  }
};

// LAPACK xLARNV seed, SAVE'd across solves in the reference (dgetv0.f:164,202-208); one per precision
struct SeedState {
  bool inited = false;
  int iseed[4] = {0, 0, 0, 0};
};

template <typename T>
class IrlBase {
 public:
  IrlBase(VecOps<T>* ops, bool parpack) : ops_(ops), par_(parpack) {}
  virtual ~IrlBase() {}

  Counters cnt;
  PhaseClock phase_clock;
  const Counters& counters() const { return cnt; }
  // largest relative disagreement between the SpMV-epilogue dots and the CGS sweep (registered-operator mode)
  T fused_dot_maxdiff = 0;
  // steps that ran inside a device-resident batch (no host round trip), and batches cut short by a rare path
  long long deferred_steps() const { return cnt_deferred_steps_; }
  long long deferred_trips() const { return cnt_deferred_trips_; }
  long long host_round_trips() const { return cnt_round_trips_; }
  // deferral needs hand-off slots the device can reach without the host (device-resident workd or a registered OP)
  void set_deferral(bool on) { defer_enabled_ = on; }
  // compatibility option: also maintain workd(ipntr(3)) = B*x = x in mode 1 (dsaitr.f:517 fills it; nothing in mode 1
  // reads it, so by default the store -- one more n-vector per step -- is skipped)
  void set_keep_bx(bool on) { keep_bx_ = on; }
  bool has_registered_op() const { return (bool)op_; }
  using FusedOp = std::function<bool(T inv, const StepGate<T>* gate, const T* resid, T* vj, T* y, T* mb_dots)>;
  void set_registered_op(std::function<void(const T*, T*)> op, FusedOp fused) {
    op_ = std::move(op);
    fused_op_ = std::move(fused);
  }

  // *eupd on a context that never ran *aupd (fresh process): make sure the mailbox exists
  void ensure_mailbox(int ncv) {
    if (!mb_) {
      ncv_ = ncv;
      setup_mailbox();
    }
  }

 protected:
  VecOps<T>* ops_;
  const bool par_;  // PARPACK semantics (pd*.f): per-rank seeds, no initial OP*x for bmat='I', ...

  // problem description (fixed at ido = 0)
  int n_ = 0, ncv_ = 0;
  char bmat_ = 'I';
  int mode_ = 1;
  // device views of the caller's arrays
  T* resid_ = nullptr;
  T* v_ = nullptr;
  int64_t ldv_ = 0;
  T* workd_ = nullptr;
  // reverse communication outputs
  int ido_ = 0;
  int ipntr_[3] = {0, 0, 0};

  // mailbox layout: slot 0 = three segments of seg_ entries each (the synchronous paths), then 4 entries for the
  // sticky stop flag of a device-resident batch, then one 3*seg_ slot per step of such a batch (at most ncv)
  T* mb_ = nullptr;
  int seg_ = 0;
  std::vector<T> mbh_;  // host copy of slot 0
  std::vector<T> logh_; // host copy of [stop | batch slots]
  T* mb_stop() { return mb_ + (size_t)3 * seg_; }
  T* slot_dev(int s) { return mb_ + (size_t)3 * seg_ + 4 + (size_t)3 * seg_ * (s - 1); }   // s = 1..ncv
  const T* slot_host(int s) const { return logh_.data() + 4 + (size_t)3 * seg_ * (s - 1); }
  T* mbA() { return mb_; }
  T* mbB() { return mb_ + seg_; }
  T* mbC() { return mb_ + 2 * seg_; }
  T* hA() { return mbh_.data(); }
  T* hB() { return mbh_.data() + seg_; }
  T* hC() { return mbh_.data() + 2 * seg_; }

  SeedState* seed_ = nullptr;

  // registered operator (opt-in extension, mode 1 / bmat='I'): the solver applies OP itself instead of
  // returning ido = +-1, so a whole solve is one *aupd call (the arpackmm-style driver loop, natively)
  std::function<void(const T* x, T* y)> op_;
  FusedOp fused_op_;
  bool ai_fused_op_ = false;
  // device-resident batch state
  bool defer_enabled_ = false;
  bool keep_bx_ = false;
  int df_first_ = 0, df_last_ = 0, df_j_ = 0;
  std::vector<char> df_fused_;
  long long cnt_deferred_steps_ = 0, cnt_deferred_trips_ = 0, cnt_round_trips_ = 0;

  T* vcol(int j1) { return v_ + (int64_t)(j1 - 1) * ldv_; }  // 1-based column
  T* slot(int off1) { return workd_ + (off1 - 1); }           // 1-based workd offset

  void setup_mailbox() {
    seg_ = ncv_ + 2;
    mb_ = ops_->mailbox((size_t)3 * seg_ * (ncv_ + 1) + 4);
    mbh_.assign((size_t)3 * seg_, T(0));
    logh_.assign((size_t)3 * seg_ * ncv_ + 4, T(0));
    df_fused_.assign((size_t)ncv_ + 2, 0);
  }

  // ---------------------------------------------------------------------------------------------
  // start / restart vector
  // ---------------------------------------------------------------------------------------------
  int gv_pc_ = 0, gv_itry_ = 1, gv_j_ = 1, gv_iter_ = 0, gv_ierr_ = 0;
  bool gv_initv_ = false;
  T gv_rnorm0_ = 0, rnorm_ = 0;

  T fetch_norm_from_dot(T* mbslot) {
    PhaseClock::Scope pcs(&phase_clock, PhaseClock::FETCH_NORM);
    cnt_round_trips_++;
    ops_->allreduce_sum(mbslot, 1);
    ops_->fetch(hC(), mbslot, 1);
    return std::sqrt(std::fabs(hC()[0]));
  }

  // returns true when finished (ierr in gv_ierr_, norm in rnorm_); false on a hand-off
  bool start_vector() {
    CO_BEGIN(gv_pc_)
    if (!seed_->inited) {
      if (!par_) {  // dgetv0.f:202-208
        seed_->iseed[0] = 1; seed_->iseed[1] = 3; seed_->iseed[2] = 5; seed_->iseed[3] = 7;
      } else {  // pdgetv0.f:233-245, digits of 1000 + 2*rank + 1
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
    if (!gv_initv_) ops_->larnv_uniform_m1_1(n_, seed_->iseed, resid_);
    // force the vector into range(OP) (dgetv0.f:245-251; PARPACK only for bmat='G', pdgetv0.f:285)
    if ((!par_ && gv_itry_ == 1) || (par_ && bmat_ == 'G')) {
      cnt.nopx++;
      ops_->copy(n_, resid_, slot(1));
      if (op_ && mode_ == 1) {
        op_(slot(1), slot(n_ + 1));
      } else {
        ipntr_[0] = 1; ipntr_[1] = n_ + 1;
        ido_ = -1;
        CO_YIELD(gv_pc_);
      }
      ops_->copy(n_, slot(n_ + 1), resid_);
    } else if (!par_ && gv_itry_ > 1 && bmat_ == 'G') {
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
    if (bmat_ == 'I' && gv_j_ == 1 && (!std::isfinite(gv_rnorm0_) || gv_rnorm0_ == T(0))) rescale_start_vector();
    rnorm_ = gv_rnorm0_;
    if (gv_j_ > 1) {
      // orthogonalise against V(:,1:j-1) with iterative refinement (dgetv0.f:326-397)
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
        if (rnorm_ > dgks_threshold<T>() * gv_rnorm0_) break;
        gv_iter_++;
        if (gv_iter_ <= (par_ ? 1 : 5)) {  // dgetv0.f:378 / pdgetv0.f
          gv_rnorm0_ = rnorm_;
        } else {
          ops_->zero(n_, resid_);
          rnorm_ = 0;
          gv_ierr_ = -1;
          break;
        }
      }
    }
    if (trace_levels().mgetv0 > 0 && ops_->rank() == 0)  // dgetv0.f:398-401
      trace::dvout1(rnorm_, "_getv0: B-norm of initial / restarted starting vector");
    CO_END(gv_pc_)
  }

  // The norms of this library are square roots of sums of squares; pdnorm2.f:72-80 (and the BLAS dnrm2 behind the
  // sequential code) divide by the largest entry first, so they survive start vectors whose squares overflow or
  // underflow.  Same result here for the only vector the caller controls: when the sum of squares of the start vector
  // is not a normal number, scale the vector by the power of two that brings its largest entry to [1, 2) -- an exact
  // operation that leaves the first Lanczos vector resid/||resid|| unchanged -- and take the norm again.
  void rescale_start_vector() {
    const int nr = ops_->nranks();
    ops_->zero((int64_t)nr, mbA());
    if (!ops_->absmax(n_, resid_, mbA() + ops_->rank())) return;
    ops_->allreduce_sum(mbA(), (size_t)nr);   // every rank filled its own slot: the sum is a gather
    ops_->fetch(hA(), mbA(), (size_t)nr);
    cnt_round_trips_++;
    T amax = T(0);
    for (int r = 0; r < nr; ++r) amax = std::max(amax, hA()[r]);
    if (!(amax > T(0)) || !std::isfinite(amax)) return;   // a genuinely zero (or non-finite) vector: as the reference
    ops_->scal(n_, std::ldexp(T(1), -std::ilogb(amax)), resid_);
    ops_->dot(n_, resid_, resid_, mbC());
    gv_rnorm0_ = fetch_norm_from_dot(mbC());
  }

  // trace level of the step routine: msaitr (symmetric) or mnaitr (nonsymmetric)
  virtual int aitr_trace_level() const { return 0; }

  // ---------------------------------------------------------------------------------------------
  // k -> k+np step extension of the factorisation
  // ---------------------------------------------------------------------------------------------
  // hooks: where the projected matrix lives differs between the symmetric (tridiagonal, 2 columns)
  // and the nonsymmetric (full Hessenberg) solver.
  virtual void h_store(int j, const T* hcol, T beta, bool after_restart) = 0;  // CGS column of step j
  virtual void h_add(int j, const T* scol, bool after_restart) = 0;            // DGKS correction
  virtual void sweep_done(int k, int np) = 0;
  virtual T tiny_norm() = 0;  // dsaitr: safmin; dnaitr: unfl after dlabad
  // dsaitr's mode-2 contract (dsaupd.f:309-313, dsaitr.f:504-515): the caller overwrote x with A*x, so B*OP*x is
  // taken from workd(ivj) and the ido = 2 hand-off is skipped.  dnaitr.f has no such shortcut (it has no mode
  // argument at all): the nonsymmetric solver always issues ido = 2 for bmat = 'G'.
  virtual bool mode2_shortcut() const { return false; }

  int ai_pc_ = 0, ai_k_ = 0, ai_np_ = 0, ai_j_ = 0, ai_itry_ = 0, ai_iter_ = 0, ai_info_ = 0;
  bool ai_rstart_ = false;
  T ai_wnorm_ = 0, ai_beta_ = 0, ai_rnorm1_ = 0;

  static constexpr int IPJ = 1;
  int irj() const { return 1 + n_; }
  int ivj() const { return 1 + 2 * n_; }

  bool can_defer() {
    static const bool off = getenv("AB200_DEFER") && std::strcmp(getenv("AB200_DEFER"), "0") == 0;
    return !off && defer_enabled_ && ops_->deferred_ok() && bmat_ == 'I' && mode_ == 1 && rnorm_ >= tiny_norm() &&
           rnorm_ > T(0);
  }

  // Host side of one orthogonalisation whose reductions are in hA()/hB()/hC() (dsaitr.f:552-780 for bmat = 'I'):
  // H column, DGKS bookkeeping and -- rare -- the third pass, which runs synchronously on the device.
  void finish_orth() {
    if (ai_fused_op_ && !par_) {
      // alpha = v_j^T OP v_j and ||OP v_j||^2 from the SpMV epilogue must agree with the CGS sweep
      const T da = std::fabs(hC()[2] - hA()[ai_j_ - 1]), dw = std::fabs(hC()[3] - hA()[ai_j_]);
      const T sc = std::sqrt(hA()[ai_j_]);
      fused_dot_maxdiff = std::max(fused_dot_maxdiff, std::max(da / (sc > T(0) ? sc : T(1)),
                                                               dw / (hA()[ai_j_] > T(0) ? hA()[ai_j_] : T(1))));
    }
    ai_wnorm_ = std::sqrt(hA()[ai_j_]);
    h_store(ai_j_, hA(), ai_beta_, ai_rstart_);
    rnorm_ = std::sqrt(hB()[ai_j_]);
    if (!(rnorm_ > dgks_threshold<T>() * ai_wnorm_)) {
      cnt.nrorth++;
      h_add(ai_j_, hB(), ai_rstart_);
      ai_rnorm1_ = std::sqrt(hC()[0]);
      if (ai_rnorm1_ > dgks_threshold<T>() * rnorm_) {
        rnorm_ = ai_rnorm1_;
      } else {
        // one more refinement pass (iter = 1), then give up (dsaitr.f:768-780)
        cnt.nitref++;
        rnorm_ = ai_rnorm1_;
        ops_->dots(n_, ai_j_, v_, ldv_, resid_, resid_, mbA());
        ops_->allreduce_sum(mbA(), (size_t)ai_j_);
        ops_->update(n_, ai_j_, v_, ldv_, mbA(), resid_, resid_, mbC());
        ops_->allreduce_sum(mbC(), 1);
        ops_->fetch(hA(), mbA(), (size_t)ai_j_);
        ops_->fetch(hC(), mbC(), 1);
        cnt_round_trips_ += 2;
        h_add(ai_j_, hA(), ai_rstart_);
        ai_rnorm1_ = std::sqrt(std::fabs(hC()[0]));
        if (ai_rnorm1_ > dgks_threshold<T>() * rnorm_) {
          rnorm_ = ai_rnorm1_;
        } else {
          cnt.nitref++;
          ops_->zero(n_, resid_);
          rnorm_ = 0;
        }
      }
    }
  }

  // returns true when finished; ai_info_ > 0 <=> no restart vector could be found (info = j-1)
  bool extend() {
    CO_BEGIN(ai_pc_)
    ai_info_ = 0;
    for (ai_j_ = ai_k_ + 1; ai_j_ <= ai_k_ + ai_np_; ++ai_j_) {
      ai_beta_ = rnorm_;
      ai_rstart_ = false;
      if (!(rnorm_ > T(0))) {
        // invariant subspace: find a new vector orthogonal to the current basis (dsaitr.f:378-427)
        ai_beta_ = 0;
        cnt.nrstrt++;
        ai_rstart_ = true;
        if (aitr_trace_level() > 0 && ops_->rank() == 0) {  // dsaitr.f:398-403
          trace::ivout1(ai_j_, "_aitr: ****** restart at step ******");
        }
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
      if (can_defer()) {
        // ---- device-resident batch (bmat = 'I', mode 1): steps ai_j_ .. k+np are enqueued back to back, the
        // host reads ONE block of mailbox slots when the batch is over and replays its bookkeeping from it ----
        df_first_ = ai_j_;
        df_last_ = ai_k_ + ai_np_;
        ops_->zero(4, mb_stop());
        ops_->set_stop_flag(mb_stop());
        for (df_j_ = df_first_; df_j_ <= df_last_; ++df_j_) {
          {
            PhaseClock::Scope pcs(&phase_clock, PhaseClock::STEP_ENQ);
            const int s = df_j_ - df_first_ + 1;
            T* sC = slot_dev(s) + 2 * seg_;
            StepGate<T> g;
            const bool first = (df_j_ == df_first_);
            if (!first) {
              g.A = slot_dev(s - 1); g.B = slot_dev(s - 1) + seg_; g.C = slot_dev(s - 1) + 2 * seg_;
              g.prev_j = df_j_ - 1; g.tiny = tiny_norm(); g.stop = mb_stop(); g.stop_code = T(df_j_ - 1);
            }
            // v_j = r/||r||, x = v_j (dsaitr.f:438-468); with a registered operator K1+K2+K3 are one kernel
            bool fused = false;
            if (fused_op_ && !keep_bx_)
              fused = fused_op_(first ? T(1) / rnorm_ : T(0), first ? nullptr : &g, resid_, vcol(df_j_), slot(irj()),
                                sC + 2);
            if (!fused) {
              T* bx = keep_bx_ ? slot(IPJ) : nullptr;
              if (first) ops_->start_step(n_, T(1) / rnorm_, resid_, vcol(df_j_), slot(ivj()), bx, true);
              else ops_->start_step_gated(n_, g, resid_, vcol(df_j_), slot(ivj()), bx);
            }
            df_fused_[s] = fused ? 1 : 0;
            if (op_ && !fused) op_(slot(ivj()), slot(irj()));
          }
          if (!op_) {
            ipntr_[0] = ivj(); ipntr_[1] = irj(); ipntr_[2] = IPJ;
            ido_ = 1;
            CO_YIELD(ai_pc_);
          }
          {
            PhaseClock::Scope pcs(&phase_clock, PhaseClock::STEP_ENQ);
            T* sA = slot_dev(df_j_ - df_first_ + 1);
            ops_->orth_step(n_, df_j_, v_, ldv_, slot(irj()), resid_, sA, sA + seg_, sA + 2 * seg_);
          }
        }
        ops_->set_stop_flag(nullptr);
        {
          PhaseClock::Scope pcs(&phase_clock, PhaseClock::FETCH_LOG);
          ops_->fetch(logh_.data(), mb_stop(), (size_t)4 + (size_t)3 * seg_ * (df_last_ - df_first_ + 1));
        }
        cnt_round_trips_++;
        // replay: the host logic of every step, in order, from the step's own mailbox slot
        for (; ai_j_ <= df_last_; ++ai_j_) {
          if (ai_j_ > df_first_) { ai_beta_ = rnorm_; ai_rstart_ = false; }
          std::copy(slot_host(ai_j_ - df_first_ + 1), slot_host(ai_j_ - df_first_ + 1) + 3 * seg_, mbh_.begin());
          ai_fused_op_ = df_fused_[ai_j_ - df_first_ + 1] != 0;
          cnt.nopx++;
          cnt_deferred_steps_++;
          {
            const int nitref0 = cnt.nitref;
            finish_orth();
            const bool rare = cnt.nitref != nitref0 || !(rnorm_ >= tiny_norm()) || !(rnorm_ > T(0));
            if (ai_j_ < df_last_) {
              // the device took the same decision in the gated start of step j+1: later kernels were early exits
              const bool stopped = (logh_[0] == T(ai_j_));
              if (rare != stopped)
                throw std::runtime_error("device-resident sweep: host and device disagree on a rare-path decision");
              if (rare) break;
            }
          }
        }
        if (ai_j_ <= df_last_) cnt_deferred_trips_++;
        // a cut-short batch leaves the loop variable on the step it stopped at: the sweep goes on from the next one
        continue;
      }
      // v_j = r/||r||, p_j = B r/||r||, x = v_j   (dsaitr.f:438-468)
      ai_fused_op_ = false;
      if (fused_op_ && !keep_bx_ && bmat_ == 'I' && mode_ == 1 && rnorm_ >= tiny_norm()) {
        // registered operator: K1+K2+K3 in one kernel (v_j written on the way, x never materialised)
        ai_fused_op_ = fused_op_(T(1) / rnorm_, nullptr, resid_, vcol(ai_j_), slot(irj()), mbC() + 2);
      }
      if (!ai_fused_op_) {
        const T tiny = tiny_norm();
        if (rnorm_ >= tiny) {
          // mode 1 / bmat 'I': the B*x slot is neither read by this code nor part of the hand-off (ipntr(3) is
          // documented for the shift-invert modes only), so the third store of K2 is skipped
          ops_->start_step(n_, T(1) / rnorm_, resid_, vcol(ai_j_), slot(ivj()),
                           (bmat_ == 'I' && mode_ == 1 && !keep_bx_) ? nullptr : slot(IPJ), bmat_ == 'I');
        } else {
          // dlascl fallback of the reference (dsaitr.f:450-453): scale in two safe steps
          const T big = std::ldexp(T(1), sizeof(T) == 8 ? 500 : 60);
          ops_->start_step(n_, T(1) / (rnorm_ * big), resid_, vcol(ai_j_), slot(ivj()), slot(IPJ),
                           bmat_ == 'I');
          ops_->scal(n_, big, vcol(ai_j_));
          ops_->scal(n_, big, slot(ivj()));
          ops_->scal(n_, big, slot(IPJ));
        }
      }
      cnt.nopx++;
      if (op_ && mode_ == 1) {
        if (!ai_fused_op_) op_(slot(ivj()), slot(irj()));  // registered operator: no hand-off
      } else {
        ipntr_[0] = ivj(); ipntr_[1] = irj(); ipntr_[2] = IPJ;
        ido_ = 1;
        CO_YIELD(ai_pc_);
      }
      // workd(irj) = OP*v_j
      if (bmat_ == 'I' && !mode2_shortcut()) {
        // ---- fused path: CGS + speculative DGKS dots, one host round trip per step (K4..K10) ----
        ops_->orth_step(n_, ai_j_, v_, ldv_, slot(irj()), resid_, mbA(), mbB(), mbC());
        ops_->fetch(mbh_.data(), mb_, (size_t)2 * seg_ + 4);
        cnt_round_trips_++;
        finish_orth();
      } else {
        // ---- generic path (bmat='G' and/or mode 2): B-inner products need hand-offs ----
        ops_->copy(n_, slot(irj()), resid_);
        if (!mode2_shortcut()) {
          // bmat == 'G' here
          cnt.nbx++;
          ipntr_[0] = irj(); ipntr_[1] = IPJ;
          ido_ = 2;
          CO_YIELD(ai_pc_);
        }
        // wnorm and the CGS coefficients use B*OP*v_j: workd(ipj), or workd(ivj) = A*v_j in mode 2
        ops_->dots(n_, ai_j_, v_, ldv_, mode2_shortcut() ? slot(ivj()) : slot(IPJ), resid_, mbA());
        ops_->allreduce_sum(mbA(), (size_t)ai_j_ + 1);
        ops_->update(n_, ai_j_, v_, ldv_, mbA(), resid_, resid_, bmat_ == 'I' ? mbC() : nullptr);
        ops_->fetch(hA(), mbA(), (size_t)ai_j_ + 1);
        ai_wnorm_ = std::sqrt(std::fabs(hA()[ai_j_]));
        h_store(ai_j_, hA(), ai_beta_, ai_rstart_);
        ai_iter_ = 0;
        if (bmat_ == 'G') {
          cnt.nbx++;
          ops_->copy(n_, resid_, slot(irj()));
          ipntr_[0] = irj(); ipntr_[1] = IPJ;
          ido_ = 2;
          CO_YIELD(ai_pc_);
          ops_->dot(n_, resid_, slot(IPJ), mbC());
        }
        rnorm_ = fetch_norm_from_dot(mbC());
        if (!(rnorm_ > dgks_threshold<T>() * ai_wnorm_)) {
          cnt.nrorth++;
          for (;;) {
            ops_->dots(n_, ai_j_, v_, ldv_, bmat_ == 'G' ? slot(IPJ) : resid_, resid_, mbA());
            ops_->allreduce_sum(mbA(), (size_t)ai_j_);
            ops_->update(n_, ai_j_, v_, ldv_, mbA(), resid_, resid_, bmat_ == 'I' ? mbC() : nullptr);
            ops_->fetch(hA(), mbA(), (size_t)ai_j_);
            h_add(ai_j_, hA(), ai_rstart_);
            if (bmat_ == 'G') {
              cnt.nbx++;
              ops_->copy(n_, resid_, slot(irj()));
              ipntr_[0] = irj(); ipntr_[1] = IPJ;
              ido_ = 2;
              CO_YIELD(ai_pc_);
              ops_->dot(n_, resid_, slot(IPJ), mbC());
            }
            ai_rnorm1_ = fetch_norm_from_dot(mbC());
            if (ai_rnorm1_ > dgks_threshold<T>() * rnorm_) {
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
    }
    sweep_done(ai_k_, ai_np_);
    CO_END(ai_pc_)
  }
};

}  // namespace ab200
