// ORACLE -- test infrastructure only. Never linked into the product library.
//
// Stage-structured QP of one RTI iteration and its interior-point solver.
//
// What it restates.  The reference hands this QP to acados' PARTIAL_CONDENSING_HPIPM with cond_N = N
// (reference controller.py:97-110,208-209), i.e. to HPIPM's OCP-QP interior-point method in BALANCE mode.
// Neither acados nor HPIPM is vendored under /root/reference (acados version unpinned, README.md:13), so
// this file restates HPIPM's *published* algorithm (G. Frison, M. Diehl, "HPIPM: a high-performance
// quadratic programming framework for model predictive control", IFAC 2020; Frison & Jorgensen, "Efficient
// implementation of the Riccati recursion for solving linear-quadratic control problems", 2013):
//   - Mehrotra predictor-corrector primal-dual IPM on the KKT system of the OCP-QP,
//   - two-sided inequality rows with (lam, t) per side, soft rows with one slack per side (L1 penalty),
//   - inequality rows condensed into the stage Hessian/gradient (Gamma = lam/t, gamma),
//   - backward Riccati factorisation (Cholesky of R + B'PB per stage), forward substitution,
//   - cold start: primal 0 shifted inside the box, t >= thr0 = 0.1, lam = mu0 / t,
//   - sigma = (mu_aff/mu)^3, step 0.995*alpha, conditional predictor-corrector (fall back to pure centering when
//     the corrected step would more than double mu_aff), exit on max-norm residuals or iter_max or alpha_min.
// PARITY UNPINNED: iteration-level details (iterative refinement, the lq_fact fallback, clipping constants) are
// not reproduced; the solution of a strictly convex QP does not depend on them beyond the exit tolerance.
//
// Variables per stage k: z_k = [du_k (nu); dx_k (nx)] for k < N, z_N = [dx_N].  Dynamics are the constant
// double integrator of reference env_model.py:63-71:  dx_{k+1} = A dx_k + B du_k + b_k.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

constexpr int QNQ = 5, QNX = 10, QNU = 5, QNZ = 15;
constexpr int QNG = 12;               // general rows: 5 torque + 6 capsule pairs + 1 viability row
constexpr int QNR = QNX + QNG;        // two-sided rows per stage (box first)
constexpr int QNC = 2 * QNR + 2;      // constraint slots per stage: lower[NR], upper[NR], slack_l, slack_u

struct QpStage {
  int nu;                    // 5, or 0 at the terminal stage
  double H[QNZ][QNZ];        // symmetric, [u;x] order (terminal stage: only the leading 10x10 = x block)
  double g[QNZ];
  double b[QNX];             // dynamics offset k -> k+1 (unused at the terminal stage)
  double blo[QNX], bhi[QNX]; // box on dx
  int ng;                    // number of general rows in use
  double C[QNG][QNZ];        // general rows, [u;x] order
  double glo[QNG], ghi[QNG];
  int soft_row;              // index into the general rows of the soft row, or -1
  double zl, zu;             // L1 penalties of the soft row's two slacks
};

struct QpSol {
  std::vector<double> z;     // (N+1)*15, stage-major, [u;x] (terminal: x in the first 10 slots)
  std::vector<double> pi;    // N*10
  std::vector<double> lam, t;  // (N+1)*QNC
  std::vector<double> s;     // (N+1)*2 slacks
  int iter = 0;
  int status = 0;            // 0 success, 1 max iter, 2 min step, 3 NaN
  double res[4] = {0, 0, 0, 0};  // max-norm residuals: stationarity, dynamics, inequality, complementarity
  double mu = 0;
  double best_ratio = 1e300; // smallest max_i(res_i / tol_i) over the iterates of the solve (margin probe of oracle.cpp)
};

struct QpOpts {
  int iter_max = 200;
  double mu0 = 1e1, tol_stat = 1e-6, tol_eq = 1e-8, tol_ineq = 1e-8, tol_comp = 1e-8, alpha_min = 1e-12;
  double reg_prim = 1e-15;
  double thr0 = 1e-1;
  double lam_min = 1e-16, t_min = 1e-16;
  int cond_pred_corr = 1;
};

class QpIpm {
 public:
  QpIpm(int N, double dt) : N_(N), dt_(dt), st_(N + 1) {
    int n = N + 1;
    sol_.z.assign(n * QNZ, 0.0); sol_.pi.assign(N * QNX, 0.0);
    sol_.lam.assign(n * QNC, 0.0); sol_.t.assign(n * QNC, 0.0); sol_.s.assign(n * 2, 0.0);
    res_g_.assign(n * QNZ, 0.0); res_gs_.assign(n * 2, 0.0); res_b_.assign(N * QNX, 0.0);
    res_d_.assign(n * QNC, 0.0); res_m_.assign(n * QNC, 0.0); rm_.assign(n * QNC, 0.0);
    dz_.assign(n * QNZ, 0.0); dpi_.assign(N * QNX, 0.0); ds_.assign(n * 2, 0.0);
    dlam_.assign(n * QNC, 0.0); dt__.assign(n * QNC, 0.0);
    dlam_aff_.assign(n * QNC, 0.0); dt_aff_.assign(n * QNC, 0.0);
    Lr_.assign(n * 25, 0.0); Ls_.assign(n * 50, 0.0); P_.assign(n * 100, 0.0); L0_.assign(100, 0.0);
    l_.assign(n * 5, 0.0); p_.assign(n * 10, 0.0);
    Hc_.assign(n * QNZ * QNZ, 0.0);
  }
  std::vector<QpStage>& stages() { return st_; }
  const QpSol& sol() const { return sol_; }
  // tolerances the residual ratios of QpSol::best_ratio refer to (default: those of the solve; the margin probe re-solves with
  // the stop test disabled and keeps the original ones here)
  void set_ratio_tolerances(double ts, double te, double ti, double tc) { ratio_tol_[0] = ts; ratio_tol_[1] = te; ratio_tol_[2] = ti; ratio_tol_[3] = tc; }
  void restore(const QpSol& s) { sol_ = s; }   // (sensitivity probe of oracle.cpp: put the un-perturbed result back)

  // true when constraint slot c of stage k exists
  bool present(int k, int c) const {
    const QpStage& S = st_[k];
    int nr = QNX + S.ng;
    if (c < QNR) return c < nr;
    if (c < 2 * QNR) return (c - QNR) < nr;
    return S.soft_row >= 0;
  }

  int solve(const QpOpts& o) {
    init(o);
    compute_residuals();
    int kk = 0;
    double alpha = 1.0;
    sol_.status = 0;
    sol_.best_ratio = 1e300;
    auto ratio = [&]() {
      const double t[4] = {ratio_tol_[0], ratio_tol_[1], ratio_tol_[2], ratio_tol_[3]};
      double r = 0.0;
      for (int i = 0; i < 4; ++i) { const double v = sol_.res[i] / t[i]; r = (v > r || !(v == v)) ? v : r; }
      return r;
    };
    for (; kk < o.iter_max; ++kk) {
      { const double r = ratio(); if (r < sol_.best_ratio) sol_.best_ratio = r; }
      if (!(sol_.res[0] > o.tol_stat || sol_.res[1] > o.tol_eq || sol_.res[2] > o.tol_ineq || sol_.res[3] > o.tol_comp)) break;
      if (!(alpha > o.alpha_min)) break;
      // ---- predictor (affine scaling direction) ----
      for (size_t i = 0; i < rm_.size(); ++i) rm_[i] = res_m_[i];
      factorize(o);
      solve_direction();
      double alpha_aff = step_length();
      double mu_aff = mu_after(alpha_aff);
      double sigma = mu_aff / sol_.mu; sigma = sigma * sigma * sigma;
      dlam_aff_ = dlam_; dt_aff_ = dt__;
      // ---- corrector ----
      for (size_t i = 0; i < rm_.size(); ++i) rm_[i] = res_m_[i] + dlam_aff_[i] * dt_aff_[i] - sigma * sol_.mu;
      solve_direction();
      alpha = step_length();
      if (o.cond_pred_corr) {
        double mu_corr = mu_after(alpha);
        if (mu_corr > 2.0 * mu_aff) {   // corrector made things worse: pure centering step instead
          for (size_t i = 0; i < rm_.size(); ++i) rm_[i] = res_m_[i] - sigma * sol_.mu;
          solve_direction();
          alpha = step_length();
        }
      }
      update(0.995 * alpha, o);
      compute_residuals();
      if (!(sol_.res[0] == sol_.res[0]) || !(sol_.res[1] == sol_.res[1]) || !(sol_.res[2] == sol_.res[2]) ||
          !(sol_.res[3] == sol_.res[3])) { sol_.status = 3; sol_.iter = kk + 1; return 3; }
    }
    { const double r = ratio(); if (r < sol_.best_ratio) sol_.best_ratio = r; }
    sol_.iter = kk;
    bool conv = !(sol_.res[0] > o.tol_stat || sol_.res[1] > o.tol_eq || sol_.res[2] > o.tol_ineq || sol_.res[3] > o.tol_comp);
    if (conv) sol_.status = 0;
    else if (kk >= o.iter_max) sol_.status = 1;
    else sol_.status = 2;
    return sol_.status;
  }

 private:
  int N_;
  double dt_;
  std::vector<QpStage> st_;
  QpSol sol_;
  std::vector<double> res_g_, res_gs_, res_b_, res_d_, res_m_, rm_, dz_, dpi_, ds_, dlam_, dt__, dlam_aff_, dt_aff_;
  std::vector<double> Lr_, Ls_, P_, L0_, l_, p_, Hc_;
  int nc_ = 0;
  double ratio_tol_[4] = {1e-6, 1e-8, 1e-8, 1e-8};

  // row product a_j . z for row j of stage k (box rows first)
  double row_dot(const QpStage& S, int j, const double* z) const {
    if (j < QNX) return z[S.nu + j];
    const double* a = S.C[j - QNX];
    double r = 0.0;
    for (int i = 0; i < S.nu + QNX; ++i) r += a[i] * z[i];
    return r;
  }
  void row_axpy(const QpStage& S, int j, double w, double* y) const {
    if (j < QNX) { y[S.nu + j] += w; return; }
    const double* a = S.C[j - QNX];
    for (int i = 0; i < S.nu + QNX; ++i) y[i] += w * a[i];
  }
  double row_lo(const QpStage& S, int j) const { return j < QNX ? S.blo[j] : S.glo[j - QNX]; }
  double row_hi(const QpStage& S, int j) const { return j < QNX ? S.bhi[j] : S.ghi[j - QNX]; }

  // y = A x + B u  (double integrator)
  void dyn(const double* z, int nu, double* y) const {
    const double* u = z; const double* x = z + nu;
    for (int i = 0; i < QNQ; ++i) {
      double ui = nu ? u[i] : 0.0;
      y[i] = x[i] + dt_ * x[QNQ + i] + 0.5 * dt_ * dt_ * ui;
      y[QNQ + i] = x[QNQ + i] + dt_ * ui;
    }
  }
  // y[15] += [B A]' v
  void dynT_add(const double* v, double* y) const {
    for (int i = 0; i < QNQ; ++i) {
      y[i] += 0.5 * dt_ * dt_ * v[i] + dt_ * v[QNQ + i];
      y[QNU + i] += v[i];
      y[QNU + QNQ + i] += dt_ * v[i] + v[QNQ + i];
    }
  }

  void init(const QpOpts& o) {
    nc_ = 0;
    std::fill(sol_.z.begin(), sol_.z.end(), 0.0);
    std::fill(sol_.pi.begin(), sol_.pi.end(), 0.0);
    std::fill(sol_.s.begin(), sol_.s.end(), 0.0);
    std::fill(sol_.lam.begin(), sol_.lam.end(), 0.0);
    std::fill(sol_.t.begin(), sol_.t.end(), 0.0);
    for (int k = 0; k <= N_; ++k) {
      const QpStage& S = st_[k];
      double* z = &sol_.z[k * QNZ];
      double* lam = &sol_.lam[k * QNC];
      double* t = &sol_.t[k * QNC];
      for (int j = 0; j < QNX; ++j) {           // box rows: move the primal inside
        double tl = z[S.nu + j] - S.blo[j], tu = S.bhi[j] - z[S.nu + j];
        if (tl < o.thr0) {
          if (tu < o.thr0) { z[S.nu + j] = 0.5 * (S.blo[j] + S.bhi[j]); tl = o.thr0; tu = o.thr0; }
          else { tl = o.thr0; z[S.nu + j] = S.blo[j] + o.thr0; }
        } else if (tu < o.thr0) { tu = o.thr0; z[S.nu + j] = S.bhi[j] - o.thr0; }
        t[j] = tl; t[QNR + j] = tu;
      }
      for (int j = QNX; j < QNX + S.ng; ++j) {  // general rows: only t is clipped
        double v = row_dot(S, j, z);
        t[j] = std::fmax(o.thr0, v - row_lo(S, j));
        t[QNR + j] = std::fmax(o.thr0, row_hi(S, j) - v);
      }
      if (S.soft_row >= 0) { t[2 * QNR] = o.thr0; t[2 * QNR + 1] = o.thr0; sol_.s[2 * k] = o.thr0; sol_.s[2 * k + 1] = o.thr0; }
      for (int c = 0; c < QNC; ++c)
        if (present(k, c)) { lam[c] = o.mu0 / t[c]; ++nc_; }
    }
  }

  void compute_residuals() {
    double ng_ = 0, nb_ = 0, nd_ = 0, nm_ = 0, mu = 0;
    for (int k = 0; k <= N_; ++k) {
      const QpStage& S = st_[k];
      const int nz = S.nu + QNX;
      const double* z = &sol_.z[k * QNZ];
      const double* lam = &sol_.lam[k * QNC];
      const double* t = &sol_.t[k * QNC];
      double* rg = &res_g_[k * QNZ];
      for (int i = 0; i < QNZ; ++i) rg[i] = 0.0;
      for (int i = 0; i < nz; ++i) {
        double r = S.g[i];
        for (int j = 0; j < nz; ++j) r += S.H[i][j] * z[j];
        rg[i] = r;
      }
      const int nr = QNX + S.ng;
      for (int j = 0; j < nr; ++j) row_axpy(S, j, lam[QNR + j] - lam[j], rg);
      if (k < N_) dynT_add(&sol_.pi[k * QNX], rg);
      if (k > 0) for (int i = 0; i < QNX; ++i) rg[S.nu + i] -= sol_.pi[(k - 1) * QNX + i];
      for (int i = 0; i < nz; ++i) ng_ = std::fmax(ng_, std::fabs(rg[i]));
      if (k < N_) {
        double y[QNX];
        dyn(z, S.nu, y);
        const double* xn = &sol_.z[(k + 1) * QNZ] + st_[k + 1].nu;
        for (int i = 0; i < QNX; ++i) {
          double r = y[i] + S.b[i] - xn[i];
          res_b_[k * QNX + i] = r;
          nb_ = std::fmax(nb_, std::fabs(r));
        }
      }
      double* rd = &res_d_[k * QNC];
      double* rm = &res_m_[k * QNC];
      for (int c = 0; c < QNC; ++c) { rd[c] = 0.0; rm[c] = 0.0; }
      for (int j = 0; j < nr; ++j) {
        double v = row_dot(S, j, z);
        bool soft = (j - QNX) == S.soft_row && S.soft_row >= 0;
        double sl = soft ? sol_.s[2 * k] : 0.0, su = soft ? sol_.s[2 * k + 1] : 0.0;
        rd[j] = t[j] - (v + sl - row_lo(S, j));
        rd[QNR + j] = t[QNR + j] - (row_hi(S, j) - v + su);
      }
      res_gs_[2 * k] = res_gs_[2 * k + 1] = 0.0;
      if (S.soft_row >= 0) {
        int j = QNX + S.soft_row;
        rd[2 * QNR] = t[2 * QNR] - sol_.s[2 * k];
        rd[2 * QNR + 1] = t[2 * QNR + 1] - sol_.s[2 * k + 1];
        res_gs_[2 * k] = S.zl - lam[j] - lam[2 * QNR];
        res_gs_[2 * k + 1] = S.zu - lam[QNR + j] - lam[2 * QNR + 1];
        ng_ = std::fmax(ng_, std::fmax(std::fabs(res_gs_[2 * k]), std::fabs(res_gs_[2 * k + 1])));
      }
      for (int c = 0; c < QNC; ++c)
        if (present(k, c)) {
          rm[c] = lam[c] * t[c];
          mu += rm[c];
          nd_ = std::fmax(nd_, std::fabs(rd[c]));
          nm_ = std::fmax(nm_, std::fabs(rm[c]));
        }
    }
    sol_.res[0] = ng_; sol_.res[1] = nb_; sol_.res[2] = nd_; sol_.res[3] = nm_;
    sol_.mu = mu / nc_;
    // NaN anywhere must surface in the norms (fmax drops NaNs)
    double chk = 0.0;
    for (double v : res_g_) chk += v;
    for (double v : res_b_) chk += v;
    for (double v : res_m_) chk += v;
    for (double v : res_d_) chk += v;
    if (!(chk == chk)) sol_.res[0] = chk;
  }

  // per-row condensation terms from the current (lam, t, res_d, rm)
  struct RowTerms { double Gam, gam; };
  void side_terms(int k, int c, double& Gam, double& cc) const {
    const double lam = sol_.lam[k * QNC + c], t = sol_.t[k * QNC + c];
    Gam = lam / t;
    cc = (rm_[k * QNC + c] - lam * res_d_[k * QNC + c]) / t;
  }
  RowTerms row_terms(int k, int j) const {
    const QpStage& S = st_[k];
    double Gl, cl, Gu, cu;
    side_terms(k, j, Gl, cl);
    side_terms(k, QNR + j, Gu, cu);
    if (S.soft_row >= 0 && j == QNX + S.soft_row) {
      double Gsl, csl, Gsu, csu;
      side_terms(k, 2 * QNR, Gsl, csl);
      side_terms(k, 2 * QNR + 1, Gsu, csu);
      double Wl = 1.0 / (Gl + Gsl), Wu = 1.0 / (Gu + Gsu);
      cl = cl - Gl * Wl * (res_gs_[2 * k] + cl + csl);
      cu = cu - Gu * Wu * (res_gs_[2 * k + 1] + cu + csu);
      Gl = Gl * Gsl * Wl;
      Gu = Gu * Gsu * Wu;
    }
    return {Gl + Gu, cl - cu};
  }

  // Cholesky of the leading n x n block of a row-major matrix with leading dimension ld (lower, in place).
  // Non-positive pivot: the column is zeroed (BLASFEO dpotrf convention) instead of producing NaNs.
  static void chol(double* M, int n, int ld) {
    for (int j = 0; j < n; ++j) {
      double d = M[j * ld + j];
      for (int k = 0; k < j; ++k) d -= M[j * ld + k] * M[j * ld + k];
      double inv = d > 0.0 ? 1.0 / std::sqrt(d) : 0.0;
      M[j * ld + j] = d > 0.0 ? std::sqrt(d) : 0.0;
      for (int i = j + 1; i < n; ++i) {
        double s = M[i * ld + j];
        for (int k = 0; k < j; ++k) s -= M[i * ld + k] * M[j * ld + k];
        M[i * ld + j] = s * inv;
      }
    }
  }

  // Backward Riccati factorisation with the inequality rows condensed in. Stores Lr, Ls, P per stage.
  void factorize(const QpOpts& o) {
    for (int k = N_; k >= 0; --k) {
      const QpStage& S = st_[k];
      const int nu = S.nu, nz = nu + QNX;
      double M[QNZ][QNZ];
      for (int i = 0; i < nz; ++i) for (int j = 0; j < nz; ++j) M[i][j] = S.H[i][j];
      for (int i = 0; i < nz; ++i) M[i][i] += o.reg_prim;
      const int nr = QNX + S.ng;
      for (int j = 0; j < nr; ++j) {
        RowTerms rt = row_terms(k, j);
        if (j < QNX) M[nu + j][nu + j] += rt.Gam;
        else {
          const double* a = S.C[j - QNX];
          for (int r = 0; r < nz; ++r) { double w = rt.Gam * a[r]; for (int c = 0; c < nz; ++c) M[r][c] += w * a[c]; }
        }
      }
      if (k < N_) {
        // M += [B A]' P_{k+1} [B A]
        const double* Pn = &P_[(k + 1) * 100];
        double W[QNX][QNZ];   // P [B A]
        for (int r = 0; r < QNX; ++r) {
          for (int c = 0; c < QNQ; ++c) {
            W[r][c] = 0.5 * dt_ * dt_ * Pn[r * 10 + c] + dt_ * Pn[r * 10 + QNQ + c];
            W[r][QNU + c] = Pn[r * 10 + c];
            W[r][QNU + QNQ + c] = dt_ * Pn[r * 10 + c] + Pn[r * 10 + QNQ + c];
          }
        }
        for (int c = 0; c < QNZ; ++c)
          for (int r = 0; r < QNQ; ++r) {
            M[r][c] += 0.5 * dt_ * dt_ * W[r][c] + dt_ * W[QNQ + r][c];
            M[QNU + r][c] += W[r][c];
            M[QNU + QNQ + r][c] += dt_ * W[r][c] + W[QNQ + r][c];
          }
      }
      // keep the condensed matrix for the vector-only re-solves
      for (int i = 0; i < QNZ; ++i) for (int j = 0; j < QNZ; ++j) Hc_[(k * QNZ + i) * QNZ + j] = (i < nz && j < nz) ? M[i][j] : 0.0;
      double* P = &P_[k * 100];
      if (nu == 0) {
        for (int i = 0; i < QNX; ++i) for (int j = 0; j < QNX; ++j) P[i * 10 + j] = M[i][j];
      } else {
        chol(&M[0][0], nu, QNZ);
        double* Lr = &Lr_[k * 25];
        double* Ls = &Ls_[k * 50];
        for (int i = 0; i < nu; ++i) for (int j = 0; j < nu; ++j) Lr[i * 5 + j] = j <= i ? M[i][j] : 0.0;
        // Ls = M_xu Lr^-T
        for (int i = 0; i < QNX; ++i)
          for (int j = 0; j < nu; ++j) {
            double s = M[nu + i][j];
            for (int c = 0; c < j; ++c) s -= Ls[i * 5 + c] * Lr[j * 5 + c];
            Ls[i * 5 + j] = Lr[j * 5 + j] > 0.0 ? s / Lr[j * 5 + j] : 0.0;
          }
        for (int i = 0; i < QNX; ++i)
          for (int j = 0; j < QNX; ++j) {
            double s = M[nu + i][nu + j];
            for (int c = 0; c < nu; ++c) s -= Ls[i * 5 + c] * Ls[j * 5 + c];
            P[i * 10 + j] = s;
          }
      }
      if (k == 0) {
        for (int i = 0; i < 100; ++i) L0_[i] = P[i];
        chol(L0_.data(), QNX, 10);
      }
    }
  }

  // Solve the Newton system for the current rm_ (uses the stored factorisation), then recover ds, dt, dlam.
  void solve_direction() {
    // backward vector recursion
    for (int k = N_; k >= 0; --k) {
      const QpStage& S = st_[k];
      const int nu = S.nu, nz = nu + QNX;
      double gv[QNZ];
      for (int i = 0; i < QNZ; ++i) gv[i] = res_g_[k * QNZ + i];
      const int nr = QNX + S.ng;
      for (int j = 0; j < nr; ++j) row_axpy(S, j, row_terms(k, j).gam, gv);
      if (k < N_) {
        const double* Pn = &P_[(k + 1) * 100];
        double v[QNX];
        for (int i = 0; i < QNX; ++i) {
          double s = p_[(k + 1) * 10 + i];
          for (int j = 0; j < QNX; ++j) s += Pn[i * 10 + j] * res_b_[k * QNX + j];
          v[i] = s;
        }
        dynT_add(v, gv);
      }
      double* p = &p_[k * 10];
      if (nu == 0) {
        for (int i = 0; i < QNX; ++i) p[i] = gv[i];
      } else {
        const double* Lr = &Lr_[k * 25];
        const double* Ls = &Ls_[k * 50];
        double* l = &l_[k * 5];
        for (int i = 0; i < nu; ++i) {
          double s = gv[i];
          for (int c = 0; c < i; ++c) s -= Lr[i * 5 + c] * l[c];
          l[i] = Lr[i * 5 + i] > 0.0 ? s / Lr[i * 5 + i] : 0.0;
        }
        for (int i = 0; i < QNX; ++i) {
          double s = gv[nu + i];
          for (int c = 0; c < nu; ++c) s -= Ls[i * 5 + c] * l[c];
          p[i] = s;
        }
      }
    }
    // stage-0 state:  P_0 dx_0 = -p_0
    {
      double y[QNX];
      for (int i = 0; i < QNX; ++i) {
        double s = -p_[i];
        for (int c = 0; c < i; ++c) s -= L0_[i * 10 + c] * y[c];
        y[i] = L0_[i * 10 + i] > 0.0 ? s / L0_[i * 10 + i] : 0.0;
      }
      double* dx = &dz_[st_[0].nu];
      for (int i = QNX - 1; i >= 0; --i) {
        double s = y[i];
        for (int c = i + 1; c < QNX; ++c) s -= L0_[c * 10 + i] * dx[c];
        dx[i] = L0_[i * 10 + i] > 0.0 ? s / L0_[i * 10 + i] : 0.0;
      }
    }
    // forward substitution
    for (int k = 0; k < N_; ++k) {
      const QpStage& S = st_[k];
      const int nu = S.nu;
      double* dz = &dz_[k * QNZ];
      const double* dx = dz + nu;
      const double* Lr = &Lr_[k * 25];
      const double* Ls = &Ls_[k * 50];
      const double* l = &l_[k * 5];
      double w[QNU];
      for (int j = 0; j < nu; ++j) {
        double s = l[j];
        for (int i = 0; i < QNX; ++i) s += Ls[i * 5 + j] * dx[i];
        w[j] = -s;
      }
      for (int i = nu - 1; i >= 0; --i) {
        double s = w[i];
        for (int c = i + 1; c < nu; ++c) s -= Lr[c * 5 + i] * dz[c];
        dz[i] = Lr[i * 5 + i] > 0.0 ? s / Lr[i * 5 + i] : 0.0;
      }
      double y[QNX];
      dyn(dz, nu, y);
      double* dxn = &dz_[(k + 1) * QNZ] + st_[k + 1].nu;
      for (int i = 0; i < QNX; ++i) dxn[i] = y[i] + res_b_[k * QNX + i];
      const double* Pn = &P_[(k + 1) * 100];
      for (int i = 0; i < QNX; ++i) {
        double s = p_[(k + 1) * 10 + i];
        for (int j = 0; j < QNX; ++j) s += Pn[i * 10 + j] * dxn[j];
        dpi_[k * QNX + i] = s;
      }
    }
    // recover slack / t / lam directions
    for (int k = 0; k <= N_; ++k) {
      const QpStage& S = st_[k];
      const double* dz = &dz_[k * QNZ];
      const double* lam = &sol_.lam[k * QNC];
      const double* t = &sol_.t[k * QNC];
      const double* rd = &res_d_[k * QNC];
      const double* rm = &rm_[k * QNC];
      double* dl = &dlam_[k * QNC];
      double* dtt = &dt__[k * QNC];
      for (int c = 0; c < QNC; ++c) { dl[c] = 0.0; dtt[c] = 0.0; }
      ds_[2 * k] = ds_[2 * k + 1] = 0.0;
      const int nr = QNX + S.ng;
      for (int j = 0; j < nr; ++j) {
        double adz = row_dot(S, j, dz);
        double dsl = 0.0, dsu = 0.0;
        if (S.soft_row >= 0 && j == QNX + S.soft_row) {
          double Gl, cl, Gu, cu, Gsl, csl, Gsu, csu;
          side_terms(k, j, Gl, cl); side_terms(k, QNR + j, Gu, cu);
          side_terms(k, 2 * QNR, Gsl, csl); side_terms(k, 2 * QNR + 1, Gsu, csu);
          dsl = -(res_gs_[2 * k] + cl + csl + Gl * adz) / (Gl + Gsl);
          dsu = -(res_gs_[2 * k + 1] + cu + csu - Gu * adz) / (Gu + Gsu);
          ds_[2 * k] = dsl; ds_[2 * k + 1] = dsu;
          dtt[2 * QNR] = dsl - rd[2 * QNR];
          dtt[2 * QNR + 1] = dsu - rd[2 * QNR + 1];
          dl[2 * QNR] = -(rm[2 * QNR] + lam[2 * QNR] * dtt[2 * QNR]) / t[2 * QNR];
          dl[2 * QNR + 1] = -(rm[2 * QNR + 1] + lam[2 * QNR + 1] * dtt[2 * QNR + 1]) / t[2 * QNR + 1];
        }
        dtt[j] = adz + dsl - rd[j];
        dtt[QNR + j] = -adz + dsu - rd[QNR + j];
        dl[j] = -(rm[j] + lam[j] * dtt[j]) / t[j];
        dl[QNR + j] = -(rm[QNR + j] + lam[QNR + j] * dtt[QNR + j]) / t[QNR + j];
      }
    }
  }

  double step_length() const {
    double alpha = 1.0;
    for (int k = 0; k <= N_; ++k)
      for (int c = 0; c < QNC; ++c)
        if (present(k, c)) {
          int i = k * QNC + c;
          if (dlam_[i] < 0.0) { double a = -sol_.lam[i] / dlam_[i]; if (a < alpha) alpha = a; }
          if (dt__[i] < 0.0) { double a = -sol_.t[i] / dt__[i]; if (a < alpha) alpha = a; }
        }
    return alpha;
  }
  double mu_after(double alpha) const {
    double mu = 0.0;
    for (int k = 0; k <= N_; ++k)
      for (int c = 0; c < QNC; ++c)
        if (present(k, c)) {
          int i = k * QNC + c;
          mu += (sol_.lam[i] + alpha * dlam_[i]) * (sol_.t[i] + alpha * dt__[i]);
        }
    return mu / nc_;
  }
  void update(double alpha, const QpOpts& o) {
    for (size_t i = 0; i < sol_.z.size(); ++i) sol_.z[i] += alpha * dz_[i];
    for (size_t i = 0; i < sol_.pi.size(); ++i) sol_.pi[i] += alpha * dpi_[i];
    for (size_t i = 0; i < sol_.s.size(); ++i) sol_.s[i] += alpha * ds_[i];
    for (int k = 0; k <= N_; ++k)
      for (int c = 0; c < QNC; ++c)
        if (present(k, c)) {
          int i = k * QNC + c;
          sol_.lam[i] = std::fmax(sol_.lam[i] + alpha * dlam_[i], o.lam_min);
          sol_.t[i] = std::fmax(sol_.t[i] + alpha * dt__[i], o.t_min);
        }
  }
};

}  // namespace orc
