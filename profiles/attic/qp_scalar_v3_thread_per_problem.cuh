// Stage-structured interior-point QP solve of one RTI iteration -- one problem per thread, 32 problems per warp.
//
// Replaces the HPIPM call inside AcadosOcpSolver.solve() (reference controller.py:158; options :97-110,208-209;
// algorithm: Frison & Diehl, HPIPM, IFAC 2020 -- Mehrotra predictor-corrector IPM, inequality rows condensed into
// the stage Hessian, backward Riccati factorisation / forward substitution).
//
// Why one thread per problem: the stage recursion of one problem is a long dependency chain of small dense
// operations (5x5 Cholesky, 10x5 triangular solve, rank-5 Schur update).  Spreading ONE problem over lanes turns every
// step of that chain into a shuffle / shared-memory round trip and leaves the FP64 pipe idle (measured: profiles/
// r01_*).  With one problem per thread all 32 lanes of a warp run the same chain on different problems, the small
// dense blocks unroll into independent FMAs (instruction-level parallelism instead of lane parallelism), and every
// global access is a fully coalesced 256-byte row of the [tile][stage][field][32 problems] layout.
//   * the constant double-integrator A, B (env_model.py:63-71) are never stored: [B A]' P [B A] is formed from the 5x5
//     blocks of P in closed form;
//   * the stage Hessian is kept as a packed lower triangle; torque rows are dense (5x15), capsule rows touch q only,
//     the viability row touches x only, box rows are diagonal;
//   * the primal-dual update of iteration i is fused into the factorisation sweep of iteration i+1 (both walk the stages
//     backwards), which removes one full pass over the data per iteration.
// The code is written against a small accessor (`A::rec(k,f)`, `A::ld(k,f)`, `A::sd(k,f,v)`) so that the same source runs
// on the device (strided [..][32] layout, qp.cu) and on the host for kernel-logic tests without a GPU (tests/emu).
#pragma once
#include "dev_model.cuh"

namespace smpc {

constexpr int QNR = 22;            // two-sided rows per stage: box 0-9, torque 10-14, capsule 15-20, viability 21
constexpr int QNS = 2 * QNR;       // constraint slots: lower[22], upper[22]

// ---- per-stage state of one problem (doubles) ----
enum {
  F_Z = 0,       // 15 primal iterate [du dq dv]
  F_PI = 15,     // 10 multipliers of the dynamics k -> k+1
  F_LAM = 25,    // 44
  F_T = 69,      // 44
  F_SLK = 113,   // 14 soft row: s_l s_u lam_sl lam_su t_sl t_su ds_l ds_u dlam_sl dlam_su dt_sl dt_su prod_sl prod_su
  F_DZ = 127,    // 15 step
  F_DPI = 142,   // 10
  F_DLAM = 152,  // 44
  F_DTT = 196,   // 44
  F_PROD = 240,  // 44 dlam_aff * dt_aff
  F_GB = 284,    // 15 res_g
  F_WV = 299,    // 10 P_{k+1} res_b_k
  F_RB = 309,    // 10 res_b_k
  F_PV = 319,    // 15 l (5) and p (10) of the current solve
  F_LR = 334,    // 15 Cholesky factor of the control block, packed lower triangle
  F_LS = 349,    // 50 Ls[i][j], i < 10, j < 5
  F_PM = 399,    // 55 Riccati matrix, packed lower triangle
  QS_ST = 454
};
constexpr int QS_L0 = 55;          // stage-0 state factor (packed lower triangle), stored after the stage blocks
SMPC_HD constexpr size_t qs_doubles_per_problem(int N) { return (size_t)(N + 1) * QS_ST + QS_L0; }

SMPC_HD int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

struct QpResult {
  int iter, status;          // status: 0 success, 1 max iter, 2 min step, 3 NaN
  double res[4], mu;
};

template <class A>
struct QpScalar {
  const smpc_problem_t& P;
  A& m;
  const double* x0;
  const int N, rrec;     // rrec: receding index (RealReceding box override)
  int nc;
  double dt, hdt2;

  // per-stage constants
  struct Stage {
    bool tau, dist, nn, soft;
    double zpen;
    double jt[75], jd[30], jn[10];
    double lo[QNR], hi[QNR];
  };

  SMPC_HD QpScalar(const smpc_problem_t& p, A& acc, const double* x0_, int r) : P(p), m(acc), x0(x0_), N(p.N), rrec(r), nc(0), dt(p.dt), hdt2(0.5 * p.dt * p.dt) {}

  SMPC_HD bool present(const Stage& s, int row) const { return row < 10 ? true : (row < 15 ? s.tau : (row < 21 ? s.dist : s.nn)); }

  SMPC_HD void load_stage(int k, Stage& s) {
    s.tau = m.rec(k, SMPC_REC_NTAU) > 0.5; s.dist = m.rec(k, SMPC_REC_NDIST) > 0.5; s.nn = m.rec(k, SMPC_REC_NNROW) > 0.5;
    s.zpen = m.rec(k, SMPC_REC_SOFT);
    s.soft = s.nn && s.zpen >= 0.0;
    for (int i = 0; i < 10; ++i) {
      const double xk = m.rec(k, SMPC_REC_X + i);
      double lo, hi;
      if (k == 0) { lo = hi = x0[i] - xk; }
      else if (k == N) { lo = P.lbx_e[i] - xk; hi = P.ubx_e[i] - xk; }
      else if (P.controller == SMPC_CTRL_REAL_RECEDING) {
        if (k == rrec) { const double c = m.rec(k + 1, SMPC_REC_X + i); lo = c - 1e-3 - xk; hi = c + 1e-3 - xk; }
        else { lo = P.x_min[i] - xk; hi = P.x_max[i] - xk; }
      } else { lo = P.lbx[i] - xk; hi = P.ubx[i] - xk; }
      s.lo[i] = lo; s.hi[i] = hi;
    }
    if (s.tau) {
      for (int i = 0; i < 75; ++i) s.jt[i] = m.rec(k, SMPC_REC_JTAU + i);
      for (int i = 0; i < 5; ++i) { const double v = m.rec(k, SMPC_REC_TAU + i); s.lo[10 + i] = P.tau_min[i] - v; s.hi[10 + i] = P.tau_max[i] - v; }
    }
    if (s.dist) {
      for (int i = 0; i < 30; ++i) s.jd[i] = m.rec(k, SMPC_REC_JDIST + i);
      for (int p = 0; p < 6; ++p) { const double v = m.rec(k, SMPC_REC_DIST + p); s.lo[15 + p] = P.pair_lo_ocp[p] - v; s.hi[15 + p] = P.pair_hi - v; }
    }
    if (s.nn) {
      for (int i = 0; i < 10; ++i) s.jn[i] = m.rec(k, SMPC_REC_JNN + i);
      const double v = m.rec(k, SMPC_REC_NN);
      s.lo[21] = 0.0 - v; s.hi[21] = 1e6 - v;
    }
  }

  // a_row . y for all rows (y in [u q v] order); absent rows give 0
  SMPC_HD void row_dots(const Stage& s, const double* y, double* out) const {
    for (int i = 0; i < 10; ++i) out[i] = y[5 + i];
    for (int i = 0; i < 5; ++i) {
      double r = 0.0;
      if (s.tau) for (int j = 0; j < 15; ++j) r += s.jt[i * 15 + j] * y[j];
      out[10 + i] = r;
    }
    for (int p = 0; p < 6; ++p) {
      double r = 0.0;
      if (s.dist) for (int j = 0; j < 5; ++j) r += s.jd[p * 5 + j] * y[5 + j];
      out[15 + p] = r;
    }
    double r = 0.0;
    if (s.nn) for (int j = 0; j < 10; ++j) r += s.jn[j] * y[5 + j];
    out[21] = r;
  }
  // y[15] += sum_rows a_row * w_row
  SMPC_HD void rows_T(const Stage& s, const double* w, double* y) const {
    for (int i = 0; i < 10; ++i) y[5 + i] += w[i];
    if (s.tau) for (int i = 0; i < 5; ++i) for (int j = 0; j < 15; ++j) y[j] += s.jt[i * 15 + j] * w[10 + i];
    if (s.dist) for (int p = 0; p < 6; ++p) for (int j = 0; j < 5; ++j) y[5 + j] += s.jd[p * 5 + j] * w[15 + p];
    if (s.nn) for (int j = 0; j < 10; ++j) y[5 + j] += s.jn[j] * w[21];
  }
  // y[15] += [B A]' v   (v: 10)
  SMPC_HD void dynT_add(const double* v, double* y) const {
    for (int i = 0; i < 5; ++i) {
      y[i] += hdt2 * v[i] + dt * v[5 + i];
      y[5 + i] += v[i];
      y[10 + i] += dt * v[i] + v[5 + i];
    }
  }

  SMPC_HD double rm_of(int mode, double lam, double t, double prod, double sigmu) const {
    return mode == 0 ? lam * t : (mode == 1 ? lam * t + prod - sigmu : lam * t - sigmu);
  }

  struct Slack { double sl, su, lsl, lsu, tsl, tsu, rsl, rsu, rgsl, rgsu; };

  // res_d of every slot: r[c] = t - (a z [+ s] - lo) / t - (hi - a z [+ s])
  SMPC_HD void res_d(const Stage& s, const double* az, const double* t, const Slack& k, double* r) const {
    for (int j = 0; j < QNR; ++j) {
      if (!present(s, j)) { r[j] = 0.0; r[QNR + j] = 0.0; continue; }
      const double sl = (j == 21 && s.soft) ? k.sl : 0.0, su = (j == 21 && s.soft) ? k.su : 0.0;
      r[j] = t[j] - (az[j] + sl - s.lo[j]);
      r[QNR + j] = t[QNR + j] - (s.hi[j] - az[j] + su);
    }
  }

  // per-row condensation terms; prod = dlam_aff*dt_aff per slot (corrector mode only), sprod: slack products
  SMPC_HD void row_terms(const Stage& s, const double* lam, const double* t, const double* r, const Slack& k, int mode, double sigmu,
                         const double* prod, double sprod_l, double sprod_u, double* Gam, double* gam, double* nu) const {
    for (int j = 0; j < QNR; ++j) {
      if (!present(s, j)) { Gam[j] = 0.0; gam[j] = 0.0; if (nu) nu[j] = 0.0; continue; }
      const double rl = rm_of(mode, lam[j], t[j], mode == 1 ? prod[j] : 0.0, sigmu);
      const double ru = rm_of(mode, lam[QNR + j], t[QNR + j], mode == 1 ? prod[QNR + j] : 0.0, sigmu);
      double cl = (rl - lam[j] * r[j]) / t[j], cu = (ru - lam[QNR + j] * r[QNR + j]) / t[QNR + j];
      double Gl = lam[j] / t[j], Gu = lam[QNR + j] / t[QNR + j];
      if (j == 21 && s.soft) {
        const double rsl = rm_of(mode, k.lsl, k.tsl, mode == 1 ? sprod_l : 0.0, sigmu);
        const double rsu = rm_of(mode, k.lsu, k.tsu, mode == 1 ? sprod_u : 0.0, sigmu);
        const double Gsl = k.lsl / k.tsl, Gsu = k.lsu / k.tsu;
        const double csl = (rsl - k.lsl * k.rsl) / k.tsl, csu = (rsu - k.lsu * k.rsu) / k.tsu;
        const double Wl = 1.0 / (Gl + Gsl), Wu = 1.0 / (Gu + Gsu);
        cl = cl - Gl * Wl * (k.rgsl + cl + csl);
        cu = cu - Gu * Wu * (k.rgsu + cu + csu);
        Gl = Gl * Gsl * Wl;
        Gu = Gu * Gsu * Wu;
      }
      Gam[j] = Gl + Gu; gam[j] = cl - cu;
      if (nu) nu[j] = lam[QNR + j] - lam[j];
    }
  }

  // ------------------------------------------------------------------------------------ S0: cold start
  SMPC_HD void init(double mu0, double thr0) {
    int cnt = 0;
    for (int k = N; k >= 0; --k) {
      Stage s;
      load_stage(k, s);
      double z[15], az[QNR];
      for (int i = 0; i < 15; ++i) z[i] = 0.0;
      double t[QNS];
      for (int i = 0; i < 10; ++i) {          // box rows: move the primal inside
        double zc = 0.0, tl = zc - s.lo[i], tu = s.hi[i] - zc;
        if (tl < thr0) {
          if (tu < thr0) { zc = 0.5 * (s.lo[i] + s.hi[i]); tl = thr0; tu = thr0; }
          else { tl = thr0; zc = s.lo[i] + thr0; }
        } else if (tu < thr0) { tu = thr0; zc = s.hi[i] - thr0; }
        z[5 + i] = zc; t[i] = tl; t[QNR + i] = tu;
      }
      row_dots(s, z, az);
      for (int j = 10; j < QNR; ++j) {
        if (present(s, j)) { t[j] = fmax(thr0, az[j] - s.lo[j]); t[QNR + j] = fmax(thr0, s.hi[j] - az[j]); }
        else { t[j] = 0.0; t[QNR + j] = 0.0; }
      }
      for (int j = 0; j < QNR; ++j) {
        const bool p = present(s, j);
        m.sd(k, F_LAM + j, p ? mu0 / t[j] : 0.0); m.sd(k, F_LAM + QNR + j, p ? mu0 / t[QNR + j] : 0.0);
        m.sd(k, F_T + j, t[j]); m.sd(k, F_T + QNR + j, t[QNR + j]);
        m.sd(k, F_DLAM + j, 0.0); m.sd(k, F_DLAM + QNR + j, 0.0); m.sd(k, F_DTT + j, 0.0); m.sd(k, F_DTT + QNR + j, 0.0);
        if (p) cnt += 2;
      }
      for (int i = 0; i < 14; ++i) {
        double v = 0.0;
        if (s.soft) v = (i == 0 || i == 1 || i == 4 || i == 5) ? thr0 : ((i == 2 || i == 3) ? mu0 / thr0 : 0.0);
        m.sd(k, F_SLK + i, v);
      }
      if (s.soft) cnt += 2;
      for (int i = 0; i < 15; ++i) { m.sd(k, F_Z + i, z[i]); m.sd(k, F_DZ + i, 0.0); }
      for (int i = 0; i < 10; ++i) { m.sd(k, F_PI + i, 0.0); m.sd(k, F_DPI + i, 0.0); }
    }
    nc = cnt;
  }

  // ------------------------------------------------- S1: (update of the previous step) + factorisation + affine gradient
  // dx0 (affine) is returned in `dx0`
  SMPC_HD void update_factorize(double a, double lam_min, double t_min, double reg, QpResult& R, double* dx0) {
    double ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, musum = 0.0, chk = 0.0;
    double Pn[55], pn[10], zxn[10];
    for (int i = 0; i < 55; ++i) Pn[i] = 0.0;
    for (int i = 0; i < 10; ++i) { pn[i] = 0.0; zxn[i] = 0.0; }
    for (int k = N; k >= 0; --k) {
      Stage s;
      load_stage(k, s);
      // ---- primal-dual update of this stage ----
      double z[15], pi[10], pim[10];
      for (int i = 0; i < 15; ++i) { z[i] = m.ld(k, F_Z + i) + a * m.ld(k, F_DZ + i); m.sd(k, F_Z + i, z[i]); }
      for (int i = 0; i < 10; ++i) {
        pi[i] = 0.0; pim[i] = 0.0;
        if (k < N) { pi[i] = m.ld(k, F_PI + i) + a * m.ld(k, F_DPI + i); m.sd(k, F_PI + i, pi[i]); }
        if (k > 0) pim[i] = m.ld(k - 1, F_PI + i) + a * m.ld(k - 1, F_DPI + i);      // updated (and stored) at stage k-1
      }
      double lam[QNS], t[QNS], r[QNS], az[QNR];
      for (int c = 0; c < QNS; ++c) {
        const int j = c < QNR ? c : c - QNR;
        lam[c] = 0.0; t[c] = 0.0;
        if (present(s, j)) {
          lam[c] = fmax(m.ld(k, F_LAM + c) + a * m.ld(k, F_DLAM + c), lam_min);
          t[c] = fmax(m.ld(k, F_T + c) + a * m.ld(k, F_DTT + c), t_min);
          m.sd(k, F_LAM + c, lam[c]); m.sd(k, F_T + c, t[c]);
          const double cc = lam[c] * t[c];
          musum += cc; chk += cc;
          nm = fmax(nm, fabs(cc));
        }
      }
      Slack sk;
      sk.sl = sk.su = sk.lsl = sk.lsu = sk.tsl = sk.tsu = sk.rsl = sk.rsu = sk.rgsl = sk.rgsu = 0.0;
      if (s.soft) {
        sk.sl = m.ld(k, F_SLK + 0) + a * m.ld(k, F_SLK + 6); sk.su = m.ld(k, F_SLK + 1) + a * m.ld(k, F_SLK + 7);
        sk.lsl = fmax(m.ld(k, F_SLK + 2) + a * m.ld(k, F_SLK + 8), lam_min); sk.lsu = fmax(m.ld(k, F_SLK + 3) + a * m.ld(k, F_SLK + 9), lam_min);
        sk.tsl = fmax(m.ld(k, F_SLK + 4) + a * m.ld(k, F_SLK + 10), t_min); sk.tsu = fmax(m.ld(k, F_SLK + 5) + a * m.ld(k, F_SLK + 11), t_min);
        m.sd(k, F_SLK + 0, sk.sl); m.sd(k, F_SLK + 1, sk.su); m.sd(k, F_SLK + 2, sk.lsl); m.sd(k, F_SLK + 3, sk.lsu);
        m.sd(k, F_SLK + 4, sk.tsl); m.sd(k, F_SLK + 5, sk.tsu);
        sk.rsl = sk.tsl - sk.sl; sk.rsu = sk.tsu - sk.su;
        sk.rgsl = s.zpen - lam[21] - sk.lsl; sk.rgsu = s.zpen - lam[QNR + 21] - sk.lsu;
        musum += sk.lsl * sk.tsl + sk.lsu * sk.tsu;
        chk += sk.lsl * sk.tsl + sk.lsu * sk.tsu + sk.rsl + sk.rsu + sk.rgsl + sk.rgsu;
        nm = fmax(nm, fmax(fabs(sk.lsl * sk.tsl), fabs(sk.lsu * sk.tsu)));
        nd = fmax(nd, fmax(fabs(sk.rsl), fabs(sk.rsu)));
        ng = fmax(ng, fmax(fabs(sk.rgsl), fabs(sk.rgsu)));
      }
      row_dots(s, z, az);
      res_d(s, az, t, sk, r);
      for (int c = 0; c < QNS; ++c) { nd = fmax(nd, fabs(r[c])); chk += r[c]; }
      // ---- condensation terms (affine rhs) ----
      double Gam[QNR], gam[QNR], nu[QNR];
      row_terms(s, lam, t, r, sk, 0, 0.0, nullptr, 0.0, 0.0, Gam, gam, nu);
      // ---- condensed stage matrix, packed lower triangle of the 15x15 ----
      double M[120];
      for (int i = 0; i < 120; ++i) M[i] = 0.0;
      const double hu = (k == N) ? 1.0 : m.rec(k, SMPC_REC_HU) + reg;
      const double hq = m.rec(k, SMPC_REC_HQ) + reg, hv = m.rec(k, SMPC_REC_HV) + reg;
      double Hqq[15];
      for (int i = 0; i < 15; ++i) Hqq[i] = m.rec(k, SMPC_REC_HQQ + i);
      for (int i = 0; i < 5; ++i) {
        M[tri(i, i)] = hu;
        for (int j = 0; j <= i; ++j) M[tri(5 + i, 5 + j)] = Hqq[tri(i, j)];
        M[tri(5 + i, 5 + i)] += hq + Gam[i];
        M[tri(10 + i, 10 + i)] = hv + Gam[5 + i];
      }
      if (s.tau)
        for (int rr = 0; rr < 5; ++rr) {
          const double* ar = s.jt + rr * 15;
          for (int i = 0; i < 15; ++i) { const double coef = Gam[10 + rr] * ar[i]; for (int j = 0; j <= i; ++j) M[tri(i, j)] += coef * ar[j]; }
        }
      if (s.dist)
        for (int p = 0; p < 6; ++p) {
          const double* ar = s.jd + p * 5;
          for (int i = 0; i < 5; ++i) { const double coef = Gam[15 + p] * ar[i]; for (int j = 0; j <= i; ++j) M[tri(5 + i, 5 + j)] += coef * ar[j]; }
        }
      if (s.nn)
        for (int i = 0; i < 10; ++i) { const double coef = Gam[21] * s.jn[i]; for (int j = 0; j <= i; ++j) M[tri(5 + i, 5 + j)] += coef * s.jn[j]; }
      // ---- stationarity residual and affine gradient ----
      double rg[15], gv[15];
      for (int i = 0; i < 5; ++i) {
        rg[i] = (k == N) ? 0.0 : m.rec(k, SMPC_REC_HU) * z[i] + m.rec(k, SMPC_REC_G + i);
        double q = m.rec(k, SMPC_REC_G + 5 + i) + m.rec(k, SMPC_REC_HQ) * z[5 + i];
        for (int j = 0; j < 5; ++j) q += Hqq[i > j ? tri(i, j) : tri(j, i)] * z[5 + j];
        rg[5 + i] = q;
        rg[10 + i] = m.rec(k, SMPC_REC_G + 10 + i) + m.rec(k, SMPC_REC_HV) * z[10 + i];
      }
      rows_T(s, nu, rg);
      if (k < N) dynT_add(pi, rg);
      if (k > 0) for (int i = 0; i < 10; ++i) rg[5 + i] -= pim[i];
      if (k == N) for (int i = 0; i < 5; ++i) rg[i] = 0.0;
      for (int i = 0; i < 15; ++i) { ng = fmax(ng, fabs(rg[i])); chk += rg[i]; m.sd(k, F_GB + i, rg[i]); gv[i] = rg[i]; }
      rows_T(s, gam, gv);
      if (k == N) for (int i = 0; i < 5; ++i) gv[i] = 0.0;
      if (k < N) {
        // res_b_k, w = P_{k+1} res_b_k, dynamics coupling
        double rb[10], y[10];
        for (int i = 0; i < 5; ++i) {
          rb[i] = z[5 + i] + dt * z[10 + i] + hdt2 * z[i] + m.rec(k, SMPC_REC_B + i) - zxn[i];
          rb[5 + i] = z[10 + i] + dt * z[i] + m.rec(k, SMPC_REC_B + 5 + i) - zxn[5 + i];
        }
        for (int i = 0; i < 10; ++i) {
          double w = 0.0;
          for (int j = 0; j < 10; ++j) w += Pn[i >= j ? tri(i, j) : tri(j, i)] * rb[j];
          nb = fmax(nb, fabs(rb[i])); chk += rb[i];
          m.sd(k, F_RB + i, rb[i]); m.sd(k, F_WV + i, w);
          y[i] = w + pn[i];
        }
        dynT_add(y, gv);
        // M += [B A]' P_{k+1} [B A] from the 5x5 blocks of P: P11 = P[q][q], P21 = P[v][q], P22 = P[v][v]
        for (int i = 0; i < 5; ++i)
          for (int j = 0; j < 5; ++j) {
            const double p11 = Pn[i >= j ? tri(i, j) : tri(j, i)];
            const double p21 = Pn[tri(5 + i, j)];       // P[v_i][q_j]
            const double p12 = Pn[tri(5 + j, i)];       // P[q_i][v_j] = P[v_j][q_i]
            const double p22 = Pn[i >= j ? tri(5 + i, 5 + j) : tri(5 + j, 5 + i)];
            if (j <= i) {
              M[tri(i, j)] += hdt2 * hdt2 * p11 + hdt2 * dt * (p12 + p21) + dt * dt * p22;          // uu
              M[tri(5 + i, 5 + j)] += p11;                                                          // qq
              M[tri(10 + i, 10 + j)] += dt * dt * p11 + dt * (p12 + p21) + p22;                     // vv
            }
            M[tri(5 + i, j)] += hdt2 * p11 + dt * p12;                                              // row q_i, col u_j: (P11 B1 + P12 B2)
            M[tri(10 + i, j)] += hdt2 * (dt * p11 + p21) + dt * (dt * p12 + p22);                   // row v_i, col u_j
            M[tri(10 + i, 5 + j)] += dt * p11 + p21;                                                // row v_i, col q_j
          }
      }
      // ---- Cholesky of the control block, Ls, Schur complement ----
      double Lr[15], Ls[50], l[5], dinv[5];
      for (int j = 0; j < 5; ++j) {
        double d = M[tri(j, j)];
        for (int c = 0; c < j; ++c) d -= Lr[tri(j, c)] * Lr[tri(j, c)];
        const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
        dinv[j] = inv;
        Lr[tri(j, j)] = d > 0.0 ? d * inv : 0.0;
        for (int i = j + 1; i < 5; ++i) {
          double v = M[tri(i, j)];
          for (int c = 0; c < j; ++c) v -= Lr[tri(i, c)] * Lr[tri(j, c)];
          Lr[tri(i, j)] = v * inv;
        }
        for (int i = 0; i < 10; ++i) {
          double v = M[tri(5 + i, j)];
          for (int c = 0; c < j; ++c) v -= Ls[i * 5 + c] * Lr[tri(j, c)];
          Ls[i * 5 + j] = v * inv;
        }
        double v = gv[j];
        for (int c = 0; c < j; ++c) v -= Lr[tri(j, c)] * l[c];
        l[j] = v * inv;
      }
      double Pk[55], pk[10];
      for (int i = 0; i < 10; ++i) {
        for (int j = 0; j <= i; ++j) {
          double v = M[tri(5 + i, 5 + j)];
          for (int c = 0; c < 5; ++c) v -= Ls[i * 5 + c] * Ls[j * 5 + c];
          Pk[tri(i, j)] = v;
        }
        double v = gv[5 + i];
        for (int c = 0; c < 5; ++c) v -= Ls[i * 5 + c] * l[c];
        pk[i] = v;
      }
      for (int i = 0; i < 15; ++i) m.sd(k, F_LR + i, Lr[i]);
      for (int i = 0; i < 50; ++i) m.sd(k, F_LS + i, Ls[i]);
      for (int i = 0; i < 5; ++i) m.sd(k, F_PV + i, l[i]);
      for (int i = 0; i < 10; ++i) m.sd(k, F_PV + 5 + i, pk[i]);
      if (k > 0) {
        for (int i = 0; i < 55; ++i) { m.sd(k, F_PM + i, Pk[i]); Pn[i] = Pk[i]; }
        for (int i = 0; i < 10; ++i) { pn[i] = pk[i]; zxn[i] = z[5 + i]; }
      } else {
        // stage 0: Cholesky of P_0, kept for the re-solves; dx_0 = -P_0^-1 p_0
        for (int j = 0; j < 10; ++j) {
          double d = Pk[tri(j, j)];
          for (int c = 0; c < j; ++c) d -= Pk[tri(j, c)] * Pk[tri(j, c)];
          const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
          Pk[tri(j, j)] = d > 0.0 ? d * inv : 0.0;
          for (int i = j + 1; i < 10; ++i) {
            double v = Pk[tri(i, j)];
            for (int c = 0; c < j; ++c) v -= Pk[tri(i, c)] * Pk[tri(j, c)];
            Pk[tri(i, j)] = v * inv;
          }
        }
        for (int i = 0; i < 55; ++i) m.sl0(i, Pk[i]);
        solve_dx0(Pk, pk, dx0);
      }
    }
    R.res[0] = (chk != chk) ? chk : ng; R.res[1] = nb; R.res[2] = nd; R.res[3] = nm;
    R.mu = musum / nc;
  }

  // dx0 = -(L L')^-1 p
  SMPC_HD void solve_dx0(const double* L, const double* p, double* dx0) const {
    double y[10];
    for (int i = 0; i < 10; ++i) {
      double v = -p[i];
      for (int c = 0; c < i; ++c) v -= L[tri(i, c)] * y[c];
      y[i] = L[tri(i, i)] > 0.0 ? v / L[tri(i, i)] : 0.0;
    }
    for (int i = 9; i >= 0; --i) {
      double v = y[i];
      for (int c = i + 1; c < 10; ++c) v -= L[tri(c, i)] * dx0[c];
      dx0[i] = L[tri(i, i)] > 0.0 ? v / L[tri(i, i)] : 0.0;
    }
  }

  // loads the iterate-dependent slot data of a stage (no update)
  SMPC_HD void load_slots(int k, const Stage& s, const double* z, double* lam, double* t, double* r, Slack& sk) {
    for (int c = 0; c < QNS; ++c) { lam[c] = m.ld(k, F_LAM + c); t[c] = m.ld(k, F_T + c); }
    sk.sl = sk.su = sk.lsl = sk.lsu = sk.tsl = sk.tsu = sk.rsl = sk.rsu = sk.rgsl = sk.rgsu = 0.0;
    if (s.soft) {
      sk.sl = m.ld(k, F_SLK + 0); sk.su = m.ld(k, F_SLK + 1); sk.lsl = m.ld(k, F_SLK + 2); sk.lsu = m.ld(k, F_SLK + 3);
      sk.tsl = m.ld(k, F_SLK + 4); sk.tsu = m.ld(k, F_SLK + 5);
      sk.rsl = sk.tsl - sk.sl; sk.rsu = sk.tsu - sk.su;
      sk.rgsl = s.zpen - lam[21] - sk.lsl; sk.rgsu = s.zpen - lam[QNR + 21] - sk.lsu;
    }
    double az[QNR];
    row_dots(s, z, az);
    res_d(s, az, t, sk, r);
  }

  // ------------------------------------------------- S3: vector-only backward recursion (corrector / centering)
  SMPC_HD void resolve_backward(int mode, double sigmu, double* dx0) {
    double pn[10];
    for (int i = 0; i < 10; ++i) pn[i] = 0.0;
    for (int k = N; k >= 0; --k) {
      Stage s;
      load_stage(k, s);
      double z[15], lam[QNS], t[QNS], r[QNS], prod[QNS];
      for (int i = 0; i < 15; ++i) z[i] = m.ld(k, F_Z + i);
      Slack sk;
      load_slots(k, s, z, lam, t, r, sk);
      for (int c = 0; c < QNS; ++c) prod[c] = mode == 1 ? m.ld(k, F_PROD + c) : 0.0;
      double Gam[QNR], gam[QNR];
      row_terms(s, lam, t, r, sk, mode, sigmu, prod, s.soft ? m.ld(k, F_SLK + 12) : 0.0, s.soft ? m.ld(k, F_SLK + 13) : 0.0, Gam, gam, nullptr);
      double gv[15];
      for (int i = 0; i < 15; ++i) gv[i] = m.ld(k, F_GB + i);
      rows_T(s, gam, gv);
      if (k == N) for (int i = 0; i < 5; ++i) gv[i] = 0.0;
      if (k < N) {
        double y[10];
        for (int i = 0; i < 10; ++i) y[i] = m.ld(k, F_WV + i) + pn[i];
        dynT_add(y, gv);
      }
      double Lr[15], l[5], pk[10];
      for (int i = 0; i < 15; ++i) Lr[i] = m.ld(k, F_LR + i);
      for (int j = 0; j < 5; ++j) {
        double v = gv[j];
        for (int c = 0; c < j; ++c) v -= Lr[tri(j, c)] * l[c];
        l[j] = Lr[tri(j, j)] > 0.0 ? v / Lr[tri(j, j)] : 0.0;
      }
      for (int i = 0; i < 10; ++i) {
        double v = gv[5 + i];
        for (int c = 0; c < 5; ++c) v -= m.ld(k, F_LS + i * 5 + c) * l[c];
        pk[i] = v;
      }
      for (int i = 0; i < 5; ++i) m.sd(k, F_PV + i, l[i]);
      for (int i = 0; i < 10; ++i) { m.sd(k, F_PV + 5 + i, pk[i]); pn[i] = pk[i]; }
      if (k == 0) {
        double L0[55];
        for (int i = 0; i < 55; ++i) L0[i] = m.ll0(i);
        solve_dx0(L0, pk, dx0);
      }
    }
  }

  // ------------------------------------------------------------------------- S2 / S4: forward substitution
  struct StepStats { double alpha, s_lin, s_quad; };
  SMPC_HD StepStats forward(int mode, double sigmu, bool store_prod, bool final, const double* dx0) {
    double alpha = 1.0, s_lin = 0.0, s_quad = 0.0;
    double dx[10];
    for (int i = 0; i < 10; ++i) dx[i] = dx0[i];
    for (int k = 0; k <= N; ++k) {
      Stage s;
      load_stage(k, s);
      // multiplier step of the dynamics k-1 -> k:  dpi_{k-1} = P_k dx_k + p_k
      if (final && k > 0) {
        for (int i = 0; i < 10; ++i) {
          double v = m.ld(k, F_PV + 5 + i);
          for (int j = 0; j < 10; ++j) v += m.ld(k, F_PM + (i >= j ? tri(i, j) : tri(j, i))) * dx[j];
          m.sd(k - 1, F_DPI + i, v);
        }
      }
      double dz[15];
      for (int i = 0; i < 5; ++i) dz[i] = 0.0;
      for (int i = 0; i < 10; ++i) dz[5 + i] = dx[i];
      if (k < N) {
        double Lr[15], w[5];
        for (int i = 0; i < 15; ++i) Lr[i] = m.ld(k, F_LR + i);
        for (int j = 0; j < 5; ++j) w[j] = m.ld(k, F_PV + j);
        for (int i = 0; i < 10; ++i) for (int j = 0; j < 5; ++j) w[j] += m.ld(k, F_LS + i * 5 + j) * dx[i];
        for (int i = 4; i >= 0; --i) {
          double v = -w[i];
          for (int c = i + 1; c < 5; ++c) v -= Lr[tri(c, i)] * dz[c];
          dz[i] = Lr[tri(i, i)] > 0.0 ? v / Lr[tri(i, i)] : 0.0;
        }
      }
      // ---- slots ----
      double z[15], lam[QNS], t[QNS], r[QNS], adz[QNR];
      for (int i = 0; i < 15; ++i) z[i] = m.ld(k, F_Z + i);
      Slack sk;
      load_slots(k, s, z, lam, t, r, sk);
      row_dots(s, dz, adz);
      double dsl = 0.0, dsu = 0.0;
      if (s.soft) {
        const double pl = mode == 1 ? m.ld(k, F_PROD + 21) : 0.0, pu = mode == 1 ? m.ld(k, F_PROD + QNR + 21) : 0.0;
        const double spl = mode == 1 ? m.ld(k, F_SLK + 12) : 0.0, spu = mode == 1 ? m.ld(k, F_SLK + 13) : 0.0;
        const double rl = rm_of(mode, lam[21], t[21], pl, sigmu), ru = rm_of(mode, lam[QNR + 21], t[QNR + 21], pu, sigmu);
        const double rsl = rm_of(mode, sk.lsl, sk.tsl, spl, sigmu), rsu = rm_of(mode, sk.lsu, sk.tsu, spu, sigmu);
        const double Gl = lam[21] / t[21], Gu = lam[QNR + 21] / t[QNR + 21], Gsl = sk.lsl / sk.tsl, Gsu = sk.lsu / sk.tsu;
        const double cl = (rl - lam[21] * r[21]) / t[21], cu = (ru - lam[QNR + 21] * r[QNR + 21]) / t[QNR + 21];
        const double csl = (rsl - sk.lsl * sk.rsl) / sk.tsl, csu = (rsu - sk.lsu * sk.rsu) / sk.tsu;
        dsl = -(sk.rgsl + cl + csl + Gl * adz[21]) / (Gl + Gsl);
        dsu = -(sk.rgsu + cu + csu - Gu * adz[21]) / (Gu + Gsu);
        const double dtsl = dsl - sk.rsl, dtsu = dsu - sk.rsu;
        const double dlsl = -(rsl + sk.lsl * dtsl) / sk.tsl, dlsu = -(rsu + sk.lsu * dtsu) / sk.tsu;
        if (dlsl < 0.0) alpha = fmin(alpha, -sk.lsl / dlsl);
        if (dlsu < 0.0) alpha = fmin(alpha, -sk.lsu / dlsu);
        if (dtsl < 0.0) alpha = fmin(alpha, -sk.tsl / dtsl);
        if (dtsu < 0.0) alpha = fmin(alpha, -sk.tsu / dtsu);
        s_lin += sk.lsl * dtsl + sk.tsl * dlsl + sk.lsu * dtsu + sk.tsu * dlsu;
        s_quad += dlsl * dtsl + dlsu * dtsu;
        if (store_prod) { m.sd(k, F_SLK + 12, dlsl * dtsl); m.sd(k, F_SLK + 13, dlsu * dtsu); }
        if (final) { m.sd(k, F_SLK + 6, dsl); m.sd(k, F_SLK + 7, dsu); m.sd(k, F_SLK + 8, dlsl); m.sd(k, F_SLK + 9, dlsu); m.sd(k, F_SLK + 10, dtsl); m.sd(k, F_SLK + 11, dtsu); }
      }
      for (int j = 0; j < QNR; ++j) {
        double dl0 = 0.0, dl1 = 0.0, dt0 = 0.0, dt1 = 0.0;
        if (present(s, j)) {
          const double sl = (j == 21) ? dsl : 0.0, su = (j == 21) ? dsu : 0.0;
          dt0 = adz[j] + sl - r[j];
          dt1 = -adz[j] + su - r[QNR + j];
          const double rm0 = rm_of(mode, lam[j], t[j], mode == 1 ? m.ld(k, F_PROD + j) : 0.0, sigmu);
          const double rm1 = rm_of(mode, lam[QNR + j], t[QNR + j], mode == 1 ? m.ld(k, F_PROD + QNR + j) : 0.0, sigmu);
          dl0 = -(rm0 + lam[j] * dt0) / t[j];
          dl1 = -(rm1 + lam[QNR + j] * dt1) / t[QNR + j];
          if (dl0 < 0.0) alpha = fmin(alpha, -lam[j] / dl0);
          if (dt0 < 0.0) alpha = fmin(alpha, -t[j] / dt0);
          if (dl1 < 0.0) alpha = fmin(alpha, -lam[QNR + j] / dl1);
          if (dt1 < 0.0) alpha = fmin(alpha, -t[QNR + j] / dt1);
          s_lin += lam[j] * dt0 + t[j] * dl0 + lam[QNR + j] * dt1 + t[QNR + j] * dl1;
          s_quad += dl0 * dt0 + dl1 * dt1;
        }
        if (store_prod) { m.sd(k, F_PROD + j, dl0 * dt0); m.sd(k, F_PROD + QNR + j, dl1 * dt1); }
        if (final) { m.sd(k, F_DLAM + j, dl0); m.sd(k, F_DLAM + QNR + j, dl1); m.sd(k, F_DTT + j, dt0); m.sd(k, F_DTT + QNR + j, dt1); }
      }
      if (final) for (int i = 0; i < 15; ++i) m.sd(k, F_DZ + i, dz[i]);
      if (k < N) {
        double dxn[10];
        for (int i = 0; i < 5; ++i) {
          dxn[i] = dx[i] + dt * dx[5 + i] + hdt2 * dz[i] + m.ld(k, F_RB + i);
          dxn[5 + i] = dx[5 + i] + dt * dz[i] + m.ld(k, F_RB + 5 + i);
        }
        for (int i = 0; i < 10; ++i) dx[i] = dxn[i];
      }
    }
    StepStats o;
    o.alpha = alpha; o.s_lin = s_lin; o.s_quad = s_quad;
    return o;
  }

  // --------------------------------------------------------------------------------------- driver
  SMPC_HD QpResult solve() {
    QpResult R;
    R.iter = 0; R.status = 0;
    const double thr0 = 1e-1, lam_min = 1e-16, t_min = 1e-16;
    init(P.qp_mu0, thr0);
    double alpha = 1.0, step = 0.0;
    double dx0[10];
    int kk = 0;
    bool nan = false;
    for (;; ++kk) {
      update_factorize(step, lam_min, t_min, P.qp_reg_prim, R, dx0);
      nan = (R.res[0] != R.res[0]) || (R.res[1] != R.res[1]) || (R.res[2] != R.res[2]) || (R.res[3] != R.res[3]);
      if (nan && kk > 0) break;
      const bool unconv = (R.res[0] > P.qp_tol_stat) || (R.res[1] > P.qp_tol_eq) || (R.res[2] > P.qp_tol_ineq) || (R.res[3] > P.qp_tol_comp);
      if (!unconv && !nan) break;
      if (kk >= P.qp_iter_max) break;
      if (!(alpha > P.qp_alpha_min)) break;
      const StepStats aff = forward(0, 0.0, true, false, dx0);
      const double mu_aff = R.mu + (aff.alpha * aff.s_lin + aff.alpha * aff.alpha * aff.s_quad) / nc;
      double sigma = mu_aff / R.mu; sigma = sigma * sigma * sigma;
      const double sigmu = sigma * R.mu;
      resolve_backward(1, sigmu, dx0);
      StepStats cor = forward(1, sigmu, false, true, dx0);
      alpha = cor.alpha;
      if (P.qp_cond_pred_corr) {
        const double mu_corr = R.mu + (alpha * cor.s_lin + alpha * alpha * cor.s_quad) / nc;
        if (mu_corr > 2.0 * mu_aff) {
          resolve_backward(2, sigmu, dx0);
          cor = forward(2, sigmu, false, true, dx0);
          alpha = cor.alpha;
        }
      }
      step = 0.995 * alpha;
    }
    R.iter = kk;
    const bool unconv = (R.res[0] > P.qp_tol_stat) || (R.res[1] > P.qp_tol_eq) || (R.res[2] > P.qp_tol_ineq) || (R.res[3] > P.qp_tol_comp);
    if (nan) R.status = 3;
    else if (!unconv) R.status = 0;
    else if (kk >= P.qp_iter_max) R.status = 1;
    else R.status = 2;
    return R;
  }
};

}  // namespace smpc
