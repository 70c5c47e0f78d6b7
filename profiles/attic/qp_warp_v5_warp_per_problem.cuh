// Stage-structured interior-point QP solve of one RTI iteration -- one problem per warp.
//
// Replaces the HPIPM call inside AcadosOcpSolver.solve() (reference controller.py:158; options :97-110,208-209;
// algorithm: Frison & Diehl, HPIPM, IFAC 2020 -- Mehrotra predictor-corrector IPM, inequality rows condensed into
// the stage Hessian, backward Riccati factorisation / forward substitution) and the full-step update / status
// mapping acados' SQP_RTI performs around it (controller.py:161-167).
//
// Work decomposition (why a warp per problem):
//   * the stage recursion of one problem is sequential, everything inside a stage is small dense algebra on the
//     15 variables z = [du(5); dq(5); dv(5)].  Lane (i, h) = (lane & 15, lane >> 4) owns ROW i of the 16x16 padded
//     stage matrix, columns 8h .. 8h+7; column 15 carries the gradient, so the backward vector recursion is the
//     same elimination as the matrix factorisation.  Operands that every lane needs (a pivot row, a constraint row)
//     are read from shared memory as broadcasts -- one wavefront feeds 16 lanes -- and all FMAs run on registers;
//   * inequality rows are owned by lanes as well: lanes 0-4 a torque row, lanes 5-14 the box row of their own
//     state variable, lanes 5-10 additionally a capsule row, lane 11 the viability row and its slacks; h selects
//     the lower / upper side.  Box multipliers therefore never leave the lane that owns the matching diagonal;
//   * control flow (iteration count, conditional corrector, early exit) is per problem = warp uniform: no divergence;
//   * the constant double-integrator A, B (env_model.py:63-71) are never stored: [B A]' P [B A] is formed in closed
//     form from the 5x5 blocks of P.
// Memory: a problem's solver state (iterate, step, factors: WS doubles per stage) lives in a per-warp workspace in
// global memory; each of the four sweeps of an IPM iteration (update+factorise backward, affine forward, corrector
// backward, corrector forward) streams one contiguous range per stage through shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier) issued one stage ahead, and writes its results back with one bulk store per stage, so the
// sequential recursion never waits on HBM latency and the kernel is bounded by HBM bandwidth, not by load latency.
// The primal-dual update of iteration i is fused into the factorisation sweep of iteration i+1 (both walk backwards).
//
// The code is written against a small warp policy W (lane id, shuffles, warp barrier, staged copies) so that the same
// source runs on the device (qp.cu) and, for kernel-logic tests without a GPU, on the host (tests/emu: 32 fibers and
// a barrier).  The product only ever instantiates the device policy.
#pragma once
#include "dev_model.cuh"

namespace smpc {

constexpr int QNR = 22;            // two-sided rows per stage: box 0-9, torque 10-14, capsule 15-20, viability 21
constexpr int QNS = 2 * QNR;       // constraint slots: lower[22], upper[22]
SMPC_HD int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // packed lower triangle, i >= j

// ---- work block of one stage in the per-warp workspace (doubles; every block starts on a 16-byte boundary) ----
enum {
  WB = 0,                                                     // B: step
  B_DZ = 0, B_DPIM = 16, B_DSLK = 26, B_DLAM = 32, B_DTT = 76,
  WA = 120,                                                   // A: iterate
  A_Z = 120, A_PIM = 136, A_SLK = 146, A_LAM = 152, A_T = 196,
  WD = 240,                                                   // D: residuals kept for the re-solves
  D_GB = 240, D_WV = 256, D_RB = 266,
  WE = 276,                                                   // E: l~ (5) and p (10) of the current solve
  WF = 292,                                                   // F: elimination multipliers T[15][5], T[j][j] = 1/d_j
  WG = 368,                                                   // G: Riccati matrix P[10][10]
  WC = 468,                                                   // C: dlam_aff * dt_aff per slot (+ the two slack slots)
  C_PROD = 468, C_SPROD = 512,
  WS = 514
};
// slack sub-layout (A_SLK / B_DSLK): [0..1] s_l s_u, [2..3] lam_sl lam_su, [4..5] t_sl t_su (resp. their steps)

constexpr int QW_IN = REC + WS;          // one staging buffer: stage record + work block (natural offsets)
constexpr int QW_OUT = WC - WA;          // largest range a sweep writes (A..G)
// scratch (doubles)
enum { S_YS = 0, S_ZS = 16, S_GG = 32, S_GM = 48, S_NU = 64, S_RB = 80, S_YV = 96, S_ROW = 112, S_L0 = 144, S_TOT = 256 };
constexpr int QW_SMEM_DOUBLES = 2 * QW_IN + 2 * QW_OUT + S_TOT;

SMPC_HD constexpr size_t qw_ws_doubles(int N) { return (size_t)(N + 1) * WS; }

struct QpResult {
  int iter, status;          // status: 0 success, 1 max iter, 2 min step, 3 NaN
  double res[4], mu;
};

template <class W>
struct QpWarp {
  W& w;
  const smpc_problem_t& P;
  const int N, lane, i, h;
  const double dt, a2, sgn;
  const double* grec;     // stage records of this problem [N+1][REC]
  double* gws;            // workspace of this warp [N+1][WS]
  double* scr;
  int rrec;               // receding index (RealReceding box override)
  int nc;
  // per-lane constants
  double x0i, lbx_i, ubx_i, lbxe_i, ubxe_i, xmin_i, xmax_i, tlo_i, thi_i, plo_i, phi;
  int gd_off, gd_cnt, gd_y;
  double ar, gr;          // row coefficients of [B A]' (.) for my row
  double b1[8], b2[8];    // column coefficients for my 8 columns
  int cq[8];              // P column index of the q-part of my columns (v-part = +5); -1: not a variable column
  double m[8];            // my row of the stage matrix, columns 8h..8h+7 (column 15 = gradient)
  double dxi;             // forward sweeps: dx_i of the current stage (lanes 5..14)

  SMPC_HD QpWarp(W& w_, const smpc_problem_t& p, const double* rec, double* ws, const double* x0, int r)
      : w(w_), P(p), N(p.N), lane(w_.lane()), i(w_.lane() & 15), h(w_.lane() >> 4), dt(p.dt), a2(0.5 * p.dt * p.dt),
        sgn((w_.lane() >> 4) ? -1.0 : 1.0), grec(rec), gws(ws), scr(w_.scratch()), rrec(r), nc(0) {
    const int xi = (i >= 5 && i < 15) ? i - 5 : 0;
    x0i = x0[xi]; lbx_i = p.lbx[xi]; ubx_i = p.ubx[xi]; lbxe_i = p.lbx_e[xi]; ubxe_i = p.ubx_e[xi]; xmin_i = p.x_min[xi]; xmax_i = p.x_max[xi];
    const int ti = i < 5 ? i : 0;
    tlo_i = p.tau_min[ti]; thi_i = p.tau_max[ti];
    const int pi_ = (i >= 5 && i <= 10) ? i - 5 : 0;
    plo_i = p.pair_lo_ocp[pi_]; phi = p.pair_hi;
    if (i < 5) { gd_off = SMPC_REC_JTAU + i * 15 + 8 * h; gd_cnt = h ? 7 : 8; gd_y = 8 * h; }
    else if (i <= 10) { gd_off = SMPC_REC_JDIST + (i - 5) * 5; gd_cnt = 5; gd_y = 5; }
    else if (i == 11) { gd_off = SMPC_REC_JNN; gd_cnt = 10; gd_y = 5; }
    else { gd_off = 0; gd_cnt = 0; gd_y = 0; }
    const int rt = i / 5;     // 0 u, 1 q, 2 v, 3 pad
    ar = rt == 0 ? a2 : (rt == 1 ? 1.0 : (rt == 2 ? dt : 0.0));
    gr = rt == 0 ? dt : (rt == 1 ? 0.0 : (rt == 2 ? 1.0 : 0.0));
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = 8 * h + c, ct = col / 5;
      b1[c] = ct == 0 ? a2 : (ct == 1 ? 1.0 : (ct == 2 ? dt : 0.0));
      b2[c] = ct == 0 ? dt : (ct == 1 ? 0.0 : (ct == 2 ? 1.0 : 0.0));
      cq[c] = col < 15 ? col % 5 : -1;
      m[c] = 0.0;
    }
    dxi = 0.0;
  }

  // ------------------------------------------------------------------------------------------- warp reductions
  SMPC_HD double wsum(double v) { for (int o = 16; o > 0; o >>= 1) v += w.shfl_xor(v, o); return v; }
  SMPC_HD double wmin(double v) { for (int o = 16; o > 0; o >>= 1) v = fmin(v, w.shfl_xor(v, o)); return v; }
  SMPC_HD double wmax(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, w.shfl_xor(v, o)); return v; }

  // ------------------------------------------------------------------------------------------- stage constants
  struct Stage {
    bool tau, dist, nn, soft, hasA, hasB;
    double zpen, bndA, bndB, lo, hi;    // lo/hi: box bounds of my variable (box lanes)
    int slotA, slotB;
  };

  SMPC_HD void stage_consts(int k, const double* rec, Stage& s) const {
    s.tau = rec[SMPC_REC_NTAU] > 0.5; s.dist = rec[SMPC_REC_NDIST] > 0.5; s.nn = rec[SMPC_REC_NNROW] > 0.5;
    s.zpen = rec[SMPC_REC_SOFT];
    s.soft = s.nn && s.zpen >= 0.0;
    s.hasA = i < 5 ? s.tau : (i < 15);
    s.hasB = (i >= 5 && i <= 10) ? s.dist : (i == 11 ? s.nn : false);
    s.slotA = h * QNR + (i < 5 ? 10 + i : i - 5);
    s.slotB = h * QNR + 10 + i;       // valid for 5 <= i <= 11: rows 15..21
    s.lo = 0.0; s.hi = 0.0; s.bndA = 0.0; s.bndB = 0.0;
    if (i < 5) {
      const double v = rec[SMPC_REC_TAU + i];
      s.bndA = h ? thi_i - v : tlo_i - v;
    } else if (i < 15) {
      const double xk = rec[SMPC_REC_X + i - 5];
      double lo, hi;
      if (k == 0) { lo = hi = x0i - xk; }
      else if (k == N) { lo = lbxe_i - xk; hi = ubxe_i - xk; }
      else if (P.controller == SMPC_CTRL_REAL_RECEDING) {
        if (k == rrec) { const double c = grec[(size_t)(k + 1) * REC + SMPC_REC_X + i - 5]; lo = c - 1e-3 - xk; hi = c + 1e-3 - xk; }
        else { lo = xmin_i - xk; hi = xmax_i - xk; }
      } else { lo = lbx_i - xk; hi = ubx_i - xk; }
      s.lo = lo; s.hi = hi;
      s.bndA = h ? hi : lo;
      if (i <= 10) { const double v = rec[SMPC_REC_DIST + i - 5]; s.bndB = h ? phi - v : plo_i - v; }
      else if (i == 11) { const double v = rec[SMPC_REC_NN]; s.bndB = h ? 1e6 - v : 0.0 - v; }
    }
  }

  // a_row . y for the general row this lane owns (torque rows are split over the two halves)
  SMPC_HD double gdot(const double* rec, const double* y) {
    double acc = 0.0;
    const double* a = rec + gd_off;
    const double* yy = y + gd_y;
#pragma unroll
    for (int c = 0; c < 10; ++c) if (c < gd_cnt) acc += a[c] * yy[c];
    const double o = w.shfl_xor(acc, 16);
    if (i < 5) acc += o;
    return acc;
  }
  SMPC_HD void gdot2(const double* rec, const double* ya, const double* yb, double& ra, double& rb) {
    double aa = 0.0, ab = 0.0;
    const double* a = rec + gd_off;
#pragma unroll
    for (int c = 0; c < 10; ++c) if (c < gd_cnt) { const double v = a[c]; aa += v * ya[gd_y + c]; ab += v * yb[gd_y + c]; }
    const double oa = w.shfl_xor(aa, 16), ob = w.shfl_xor(ab, 16);
    if (i < 5) { aa += oa; ab += ob; }
    ra = aa; rb = ab;
  }

  // (C' w)_i for my variable: w given per general row in scratch (tau 0-4, dist 5-10, nn 11); work split over the halves
  SMPC_HD double rowsT(const Stage& s, const double* rec, const double* wv) {
    double acc = 0.0;
    if (i < 15) {
      if (h == 0) {
        if (s.tau) {
#pragma unroll
          for (int r = 0; r < 5; ++r) acc += rec[SMPC_REC_JTAU + r * 15 + i] * wv[r];
        }
      } else {
        if (s.dist && i >= 5 && i < 10) {
#pragma unroll
          for (int p = 0; p < 6; ++p) acc += rec[SMPC_REC_JDIST + p * 5 + i - 5] * wv[5 + p];
        }
        if (s.nn && i >= 5) acc += rec[SMPC_REC_JNN + i - 5] * wv[11];
      }
    }
    return acc + w.shfl_xor(acc, 16);
  }
  SMPC_HD void rowsT2(const Stage& s, const double* rec, const double* wa, const double* wb, double& ra, double& rb) {
    double aa = 0.0, ab = 0.0;
    if (i < 15) {
      if (h == 0) {
        if (s.tau) {
#pragma unroll
          for (int r = 0; r < 5; ++r) { const double v = rec[SMPC_REC_JTAU + r * 15 + i]; aa += v * wa[r]; ab += v * wb[r]; }
        }
      } else {
        if (s.dist && i >= 5 && i < 10) {
#pragma unroll
          for (int p = 0; p < 6; ++p) { const double v = rec[SMPC_REC_JDIST + p * 5 + i - 5]; aa += v * wa[5 + p]; ab += v * wb[5 + p]; }
        }
        if (s.nn && i >= 5) { const double v = rec[SMPC_REC_JNN + i - 5]; aa += v * wa[11]; ab += v * wb[11]; }
      }
    }
    ra = aa + w.shfl_xor(aa, 16);
    rb = ab + w.shfl_xor(ab, 16);
  }

  // ([B A]' v)_i for a 10-vector v in shared memory
  SMPC_HD double dynT(const double* v) const {
    if (i >= 15) return 0.0;
    const int j = i % 5;
    return ar * v[j] + gr * v[5 + j];
  }

  SMPC_HD static double rm_of(int mode, double lam, double t, double prod, double sigmu) {
    return mode == 0 ? lam * t : (mode == 1 ? lam * t + prod - sigmu : lam * t - sigmu);
  }

  // condensation terms of one row side
  struct Side {
    double lam, t, r, it, G, c;                    // c = (rm - lam r)/t
    double s, ls, ts, rsl, rgs, Gs, cs, rms;       // soft row only
    double rm;
  };

  // side terms of the slot (lam, t) with row product az, bound bnd, for right-hand-side mode `mode`
  SMPC_HD void side_terms(bool present, double lam, double t, double az, double bnd, double slack, int mode, double sigmu, double prod, Side& o) const {
    o.lam = lam; o.t = t; o.s = slack;
    o.r = present ? t - (sgn * (az - bnd) + slack) : 0.0;
    o.rm = rm_of(mode, lam, t, prod, sigmu);
    o.it = present ? 1.0 / t : 0.0;
    o.G = lam * o.it;
    o.c = (o.rm - lam * o.r) * o.it;
  }
  SMPC_HD void soft_terms(double zpen, double ls, double ts, int mode, double sigmu, double sprod, Side& o, double& Gc, double& cc) const {
    o.ls = ls; o.ts = ts;
    o.rsl = ts - o.s;
    o.rgs = zpen - o.lam - ls;
    o.rms = rm_of(mode, ls, ts, sprod, sigmu);
    const double its = 1.0 / ts;
    o.Gs = ls * its;
    o.cs = (o.rms - ls * o.rsl) * its;
    const double Wl = 1.0 / (o.G + o.Gs);
    cc = o.c - o.G * Wl * (o.rgs + o.c + o.cs);
    Gc = o.G * o.Gs * Wl;
  }

  // ------------------------------------------------------------------------------------------- staging helpers
  SMPC_HD const double* grec_k(int k) const { return grec + (size_t)k * REC; }
  SMPC_HD double* gws_k(int k) const { return gws + (size_t)k * WS; }
  // fetch stage k: record + work range [lo, hi) (+ optional second range) into staging buffer k & 1
  SMPC_HD void fetch(int k, int lo, int hi, int lo2 = 0, int hi2 = 0) {
    const int buf = k & 1;
    double* dst = w.inbuf(buf);
    w.load_begin(buf, (REC + (hi - lo) + (hi2 - lo2)) * (int)sizeof(double));
    w.load(buf, dst, grec_k(k), REC);
    if (hi > lo) w.load(buf, dst + REC + lo, gws_k(k) + lo, hi - lo);
    if (hi2 > lo2) w.load(buf, dst + REC + lo2, gws_k(k) + lo2, hi2 - lo2);
  }

  // ------------------------------------------------------------------------- elimination step on column j
  // all lanes: publish row j, everybody reads the pivot, its own multiplier (by symmetry M[i][j] = M[j][i]) and the
  // pivot row entries of its columns.  LDL' form: no square roots; a non-positive pivot zeroes the column (BLASFEO
  // dpotrf convention, as the oracle).
  template <int j>
  SMPC_HD double elim_step(double* rowx) {
    double* rx = rowx + 16 * (j & 1);
    if (i == j) {
#pragma unroll
      for (int c = 0; c < 8; ++c) rx[8 * h + c] = m[c];
    }
    w.sync();
    const double d = rx[j];
    const double invd = d > 0.0 ? 1.0 / d : 0.0;
    const double tij = (i > j && i < 15) ? rx[i] * invd : 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (8 * h + c > j) m[c] -= tij * rx[8 * h + c];      // h is runtime: the compiler predicates on it
    return i == j ? invd : tij;
  }
  // vector-only elimination step (re-solves): g_i -= T[i][j] g_j
  SMPC_HD void velim_step(int j, double tij, double& g) {
    const double gj = w.shfl(g, j);
    if (i > j && i < 15) g -= tij * gj;
  }

  // back substitution of the stage-0 state: dx_i = -(invd_i g_i) - sum_{c>i} T0[c][i] dx_c, i = 14..5 (g forward-eliminated)
  SMPC_HD void solve_dx0(double g) {
    const double* T0 = scr + S_L0;       // T0[(r)*10 + c], r,c in 0..9 (variables 5..14), diagonal = 1/d
    const double invd = (i >= 5 && i < 15) ? T0[(i - 5) * 10 + (i - 5)] : 0.0;
    double acc = invd * g;
    double mine = 0.0;
    for (int c = 14; c >= 5; --c) {
      const double dxc = w.shfl(invd > 0.0 ? -acc : 0.0, c);
      if (i == c) mine = dxc;
      if (i >= 5 && i < c) acc += T0[(c - 5) * 10 + (i - 5)] * dxc;
    }
    dxi = mine;
  }

  // =====================================================================================================
  // S1: (primal-dual update of the previous step | cold start) + residuals + condensation + factorisation + affine rhs
  // =====================================================================================================
  SMPC_HD void update_factorize(bool first, double a, QpResult& R) {
    const double lam_min = 1e-16, t_min = 1e-16, thr0 = 1e-1, mu0 = P.qp_mu0, reg = P.qp_reg_prim;
    double ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, musum = 0.0, chk = 0.0;
    int cnt = 0;
    const int lo = first ? WA : WB, hi = first ? WA : WD;      // first: nothing to load from the workspace
    fetch(N, lo, hi);
    for (int k = N; k >= 0; --k) {
      const int buf = k & 1;
      w.load_wait(buf);
      if (k > 0) fetch(k - 1, lo, hi);
      w.store_wait(1);                                          // OUT[buf] (used two stages ago) has been read out
      const double* rec = w.inbuf(buf);
      const double* wk = rec + REC;
      double* out = w.outbuf(buf) - WA;                         // out[A_..], out[D_..], ... address the A..G range
      const double* outp = w.outbuf(buf ^ 1) - WA;              // outputs of stage k+1
      Stage s;
      stage_consts(k, rec, s);
      double* ys = scr + S_YS;
      // ---- iterate of this stage: z, pi ----
      double z = 0.0, pim = 0.0;
      double tA = 0.0, tB = 0.0, lamA = 0.0, lamB = 0.0;
      if (first) {
        if (i >= 5 && i < 15) {                                 // box rows: move the primal inside
          double zc = 0.0, tl = zc - s.lo, tu = s.hi - zc;
          if (tl < thr0) {
            if (tu < thr0) { zc = 0.5 * (s.lo + s.hi); tl = thr0; tu = thr0; }
            else { tl = thr0; zc = s.lo + thr0; }
          } else if (tu < thr0) { tu = thr0; zc = s.hi - thr0; }
          z = zc; tA = h ? tu : tl;
        }
      } else {
        if (i < 15) z = wk[A_Z + i] + a * wk[B_DZ + i];
        if (i >= 5 && i < 15 && k > 0) pim = wk[A_PIM + i - 5] + a * wk[B_DPIM + i - 5];
      }
      if (k == N && i < 5) z = 0.0;
      if (h == 0 && i < 15) { ys[i] = z; out[A_Z + i] = z; }
      if (h == 0 && i == 15) { ys[15] = 0.0; out[A_Z + 15] = 0.0; }
      if (h == 0 && i >= 5 && i < 15) out[A_PIM + i - 5] = pim;
      w.sync();
      // ---- row products, slots ----
      const double dotg = gdot(rec, ys);
      const double azA = i < 5 ? dotg : z;
      const double azB = dotg;
      double sl = 0.0, ls = 0.0, ts = 0.0;
      if (first) {
        if (i < 5 && s.hasA) tA = fmax(thr0, sgn * (azA - s.bndA));
        if (s.hasB) tB = fmax(thr0, sgn * (azB - s.bndB));
        lamA = s.hasA ? mu0 / tA : 0.0;
        lamB = s.hasB ? mu0 / tB : 0.0;
        if (i == 11 && s.soft) { sl = thr0; ls = mu0 / thr0; ts = thr0; }
      } else {
        if (s.hasA) { lamA = fmax(wk[A_LAM + s.slotA] + a * wk[B_DLAM + s.slotA], lam_min); tA = fmax(wk[A_T + s.slotA] + a * wk[B_DTT + s.slotA], t_min); }
        if (s.hasB) { lamB = fmax(wk[A_LAM + s.slotB] + a * wk[B_DLAM + s.slotB], lam_min); tB = fmax(wk[A_T + s.slotB] + a * wk[B_DTT + s.slotB], t_min); }
        if (i == 11 && s.soft) {
          sl = wk[A_SLK + h] + a * wk[B_DSLK + h];
          ls = fmax(wk[A_SLK + 2 + h] + a * wk[B_DSLK + 2 + h], lam_min);
          ts = fmax(wk[A_SLK + 4 + h] + a * wk[B_DSLK + 4 + h], t_min);
        }
      }
      if (i < 15) { out[A_LAM + s.slotA] = lamA; out[A_T + s.slotA] = tA; }
      if (i >= 5 && i <= 11) { out[A_LAM + s.slotB] = lamB; out[A_T + s.slotB] = tB; }
      if (i == 11) { out[A_SLK + h] = sl; out[A_SLK + 2 + h] = ls; out[A_SLK + 4 + h] = ts; }
      Side SA, SB;
      side_terms(s.hasA, lamA, tA, azA, s.bndA, 0.0, 0, 0.0, 0.0, SA);
      side_terms(s.hasB, lamB, tB, azB, s.bndB, (i == 11 && s.soft) ? sl : 0.0, 0, 0.0, 0.0, SB);
      double GB = SB.G, cB = SB.c;
      if (i == 11 && s.soft) {
        soft_terms(s.zpen, ls, ts, 0, 0.0, 0.0, SB, GB, cB);
        musum += SB.rms; chk += SB.rms + SB.rsl + SB.rgs;
        nm = fmax(nm, fabs(SB.rms)); nd = fmax(nd, fabs(SB.rsl)); ng = fmax(ng, fabs(SB.rgs));
        cnt += 1;
      }
      if (s.hasA) { musum += SA.rm; chk += SA.rm + SA.r; nm = fmax(nm, fabs(SA.rm)); nd = fmax(nd, fabs(SA.r)); cnt += 1; }
      if (s.hasB) { musum += SB.rm; chk += SB.rm + SB.r; nm = fmax(nm, fabs(SB.rm)); nd = fmax(nd, fabs(SB.r)); cnt += 1; }
      // row totals: Gam = G_l + G_u, gam = c_l - c_u, nu = lam_u - lam_l
      const double GamA = SA.G + w.shfl_xor(SA.G, 16);
      const double gamA = sgn * SA.c + w.shfl_xor(sgn * SA.c, 16);
      const double nuA = -sgn * lamA - w.shfl_xor(sgn * lamA, 16);
      const double GamB = GB + w.shfl_xor(GB, 16);
      const double gamB = sgn * cB + w.shfl_xor(sgn * cB, 16);
      const double nuB = -sgn * lamB - w.shfl_xor(sgn * lamB, 16);
      if (h == 0) {
        if (i < 5) { scr[S_GG + i] = s.hasA ? GamA : 0.0; scr[S_GM + i] = s.hasA ? gamA : 0.0; scr[S_NU + i] = s.hasA ? nuA : 0.0; }
        if (i >= 5 && i <= 11) { scr[S_GG + i] = s.hasB ? GamB : 0.0; scr[S_GM + i] = s.hasB ? gamB : 0.0; scr[S_NU + i] = s.hasB ? nuB : 0.0; }
      }
      // ---- dynamics residual of the link k -> k+1 (state lanes) ----
      double rb = 0.0;
      if (k < N && i >= 5 && i < 15) {
        const int j = i < 10 ? i - 5 : i - 10;
        const double zq = ys[5 + j], zv = ys[10 + j], zu = ys[j];
        const double xn = outp[A_Z + i];
        rb = (i < 10 ? zq + dt * zv + a2 * zu : zv + dt * zu) + rec[SMPC_REC_B + i - 5] - xn;
        if (h == 0) { scr[S_RB + i - 5] = rb; out[D_RB + i - 5] = rb; nb = fmax(nb, fabs(rb)); chk += rb; }
      }
      w.sync();
      // ---- stationarity residual and affine gradient ----
      double ctn, ctg;
      rowsT2(s, rec, scr + S_NU, scr + S_GM, ctn, ctg);
      double rg = 0.0, gv = 0.0;
      if (i < 15) {
        double hz;
        if (i < 5) hz = (k == N) ? 0.0 : rec[SMPC_REC_HU] * z + rec[SMPC_REC_G + i];
        else if (i < 10) {
          hz = rec[SMPC_REC_G + i] + rec[SMPC_REC_HQ] * z;
          const int a_ = i - 5;
#pragma unroll
          for (int j = 0; j < 5; ++j) hz += rec[SMPC_REC_HQQ + (a_ >= j ? tri(a_, j) : tri(j, a_))] * ys[5 + j];
        } else hz = rec[SMPC_REC_G + i] + rec[SMPC_REC_HV] * z;
        rg = hz + ctn + (i >= 5 ? nuA : 0.0);
        if (k < N) rg += dynT(outp + A_PIM);
        if (i >= 5) rg -= pim;
        if (k == N && i < 5) rg = 0.0;
        gv = rg + ctg + (i >= 5 ? gamA : 0.0);
        if (h == 0) { ng = fmax(ng, fabs(rg)); chk += rg; out[D_GB + i] = rg; }
      }
      // w = P_{k+1} res_b, y = w + p_{k+1}
      if (k < N) {
        double wv = 0.0;
        if (i >= 5 && i < 15) {
          const double* Pr = outp + WG + (i - 5) * 10 + 5 * h;
          const double* rbs = scr + S_RB + 5 * h;
#pragma unroll
          for (int c = 0; c < 5; ++c) wv += Pr[c] * rbs[c];
        }
        wv += w.shfl_xor(wv, 16);
        if (h == 0 && i >= 5 && i < 15) { out[D_WV + i - 5] = wv; scr[S_YV + i - 5] = wv + outp[WE + i]; }
        w.sync();
        if (i < 15) gv += dynT(scr + S_YV);
      }
      if (k == N && i < 5) gv = 0.0;
      // ---- condensed stage matrix: my row, my 8 columns ----
      {
        const double hu = (k == N) ? 1.0 : rec[SMPC_REC_HU] + reg;
        const double hq = rec[SMPC_REC_HQ] + reg, hv = rec[SMPC_REC_HV] + reg;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = 8 * h + c;
          double v = 0.0;
          if (i < 15 && col < 15) {
            if (col == i) v = i < 5 ? hu : ((i < 10 ? hq : hv) + GamA);     // box multipliers sit on my own diagonal
            if (i >= 5 && i < 10 && col >= 5 && col < 10) { const int a_ = i - 5, b_ = col - 5; v += rec[SMPC_REC_HQQ + (a_ >= b_ ? tri(a_, b_) : tri(b_, a_))]; }
          }
          m[c] = v;
        }
        if (i < 15) {
          if (s.tau) {
#pragma unroll
            for (int r = 0; r < 5; ++r) {
              const double* row = rec + SMPC_REC_JTAU + r * 15;
              const double coef = scr[S_GG + r] * row[i];
#pragma unroll
              for (int c = 0; c < 8; ++c) if (8 * h + c < 15) m[c] += coef * row[8 * h + c];
            }
          }
          if (s.dist && i >= 5 && i < 10) {
#pragma unroll
            for (int p = 0; p < 6; ++p) {
              const double* row = rec + SMPC_REC_JDIST + p * 5;
              const double coef = scr[S_GG + 5 + p] * row[i - 5];
#pragma unroll
              for (int c = 0; c < 8; ++c) { const int col = 8 * h + c; if (col >= 5 && col < 10) m[c] += coef * row[col - 5]; }
            }
          }
          if (s.nn && i >= 5) {
            const double* row = rec + SMPC_REC_JNN;
            const double coef = scr[S_GG + 11] * row[i - 5];
#pragma unroll
            for (int c = 0; c < 8; ++c) { const int col = 8 * h + c; if (col >= 5 && col < 15) m[c] += coef * row[col - 5]; }
          }
          if (k < N) {
            const double* Pq = outp + WG + (i % 5) * 10;        // row q_ii of P_{k+1}
            const double* Pv = Pq + 50;                           // row v_ii
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              if (8 * h + c < 15) {
                const int j = cq[c];
                m[c] += ar * (b1[c] * Pq[j] + b2[c] * Pq[5 + j]) + gr * (b1[c] * Pv[j] + b2[c] * Pv[5 + j]);
              }
            }
          }
        }
        if (h == 1) m[7] = (i < 15) ? gv : 0.0;
      }
      // ---- eliminate the control columns (gradient column rides along) ----
      double* rowx = scr + S_ROW;
      double* Tout = out + WF;
      double tv;
      tv = elim_step<0>(rowx); if (h == 0 && i < 15) Tout[i * 5 + 0] = tv;
      tv = elim_step<1>(rowx); if (h == 0 && i < 15) Tout[i * 5 + 1] = tv;
      tv = elim_step<2>(rowx); if (h == 0 && i < 15) Tout[i * 5 + 2] = tv;
      tv = elim_step<3>(rowx); if (h == 0 && i < 15) Tout[i * 5 + 3] = tv;
      tv = elim_step<4>(rowx); if (h == 0 && i < 15) Tout[i * 5 + 4] = tv;
      if (h == 0 && i == 15) out[WF + 75] = 0.0;
      // l~ (u lanes), p (x lanes): the gradient column
      if (h == 1) out[WE + i] = (i < 15) ? m[7] : 0.0;
      // P_k = trailing block
      if (i >= 5 && i < 15) {
#pragma unroll
        for (int c = 0; c < 8; ++c) { const int col = 8 * h + c; if (col >= 5 && col < 15) out[WG + (i - 5) * 10 + col - 5] = m[c]; }
      }
      if (k == 0) {
        // factorise P_0 as well (kept in scratch for the re-solves) and solve for dx_0
        double* T0 = scr + S_L0;
#define SMPC_E0(J) tv = elim_step<J>(rowx); if (h == 0 && i >= 5 && i < 15 && i >= J) T0[(i - 5) * 10 + (J - 5)] = tv;
        SMPC_E0(5) SMPC_E0(6) SMPC_E0(7) SMPC_E0(8) SMPC_E0(9) SMPC_E0(10) SMPC_E0(11) SMPC_E0(12) SMPC_E0(13) SMPC_E0(14)
#undef SMPC_E0
        w.sync();
        solve_dx0(w.shfl(m[7], 16 + i));
      }
      w.store(gws_k(k) + WA, w.outbuf(buf), QW_OUT);
    }
    w.store_wait(0);
    ng = wmax(ng); nb = wmax(nb); nd = wmax(nd); nm = wmax(nm); musum = wsum(musum); chk = wsum(chk);
    if (first) {
      int c = cnt;
      for (int o = 16; o > 0; o >>= 1) c += w.shfl_xor_i(c, o);
      nc = c;
    }
    R.res[0] = (chk != chk) ? chk : ng; R.res[1] = nb; R.res[2] = nd; R.res[3] = nm;
    R.mu = musum / nc;
  }

  // =====================================================================================================
  // S3: vector-only backward recursion for the corrector (mode 1) / centering (mode 2) right-hand side
  // =====================================================================================================
  SMPC_HD void resolve_backward(int mode, double sigmu) {
    fetch(N, WA, WG, mode == 1 ? WC : 0, mode == 1 ? WS : 0);
    for (int k = N; k >= 0; --k) {
      const int buf = k & 1;
      w.load_wait(buf);
      if (k > 0) fetch(k - 1, WA, WG, mode == 1 ? WC : 0, mode == 1 ? WS : 0);
      w.store_wait(1);
      const double* rec = w.inbuf(buf);
      const double* wk = rec + REC;
      double* out = w.outbuf(buf) - WE;
      const double* outp = w.outbuf(buf ^ 1) - WE;
      Stage s;
      stage_consts(k, rec, s);
      const double z = i < 15 ? wk[A_Z + i] : 0.0;
      const double dotg = gdot(rec, wk + A_Z);
      const double azA = i < 5 ? dotg : z, azB = dotg;
      Side SA, SB;
      const double lamA = s.hasA ? wk[A_LAM + s.slotA] : 0.0, tA = s.hasA ? wk[A_T + s.slotA] : 0.0;
      const double lamB = s.hasB ? wk[A_LAM + s.slotB] : 0.0, tB = s.hasB ? wk[A_T + s.slotB] : 0.0;
      const double prA = (mode == 1 && s.hasA) ? wk[C_PROD + s.slotA] : 0.0, prB = (mode == 1 && s.hasB) ? wk[C_PROD + s.slotB] : 0.0;
      const bool softl = (i == 11 && s.soft);
      side_terms(s.hasA, lamA, tA, azA, s.bndA, 0.0, mode, sigmu, prA, SA);
      side_terms(s.hasB, lamB, tB, azB, s.bndB, softl ? wk[A_SLK + h] : 0.0, mode, sigmu, prB, SB);
      double GB = SB.G, cB = SB.c;
      if (softl) soft_terms(s.zpen, wk[A_SLK + 2 + h], wk[A_SLK + 4 + h], mode, sigmu, mode == 1 ? wk[C_SPROD + h] : 0.0, SB, GB, cB);
      const double gamA = sgn * SA.c + w.shfl_xor(sgn * SA.c, 16);
      const double gamB = sgn * cB + w.shfl_xor(sgn * cB, 16);
      if (h == 0) {
        if (i < 5) scr[S_GM + i] = s.hasA ? gamA : 0.0;
        if (i >= 5 && i <= 11) scr[S_GM + i] = s.hasB ? gamB : 0.0;
        if (k < N && i >= 5 && i < 15) scr[S_YV + i - 5] = wk[D_WV + i - 5] + outp[WE + i];
      }
      w.sync();
      double g = 0.0;
      const double ctg = rowsT(s, rec, scr + S_GM);
      if (i < 15) {
        g = wk[D_GB + i] + ctg + (i >= 5 ? gamA : 0.0);
        if (k < N) g += dynT(scr + S_YV);
        if (k == N && i < 5) g = 0.0;
      }
      const double* T = wk + WF + (i < 15 ? i : 0) * 5;
#pragma unroll
      for (int j = 0; j < 5; ++j) velim_step(j, T[j], g);
      if (h == 0) out[WE + i] = (i < 15) ? g : 0.0;
      if (k == 0) {
        const double* T0 = scr + S_L0;
        for (int j = 5; j < 15; ++j) velim_step(j, (i > j && i < 15) ? T0[(i - 5) * 10 + (j - 5)] : 0.0, g);
        solve_dx0(g);
      }
      w.store(gws_k(k) + WE, w.outbuf(buf), WF - WE);
      w.sync();      // scratch (S_GM, S_YV) is rewritten by the next stage
    }
    w.store_wait(0);
  }

  // =====================================================================================================
  // S2 / S4: forward substitution; step lengths and complementarity sums of the step
  // =====================================================================================================
  struct StepStats { double alpha, s_lin, s_quad; };
  SMPC_HD StepStats forward(int mode, double sigmu, bool store_prod, bool final) {
    double alpha = 1.0, s_lin = 0.0, s_quad = 0.0;
    const int lo = WA, hi = final ? WC : WG;
    const int lo2 = mode == 1 ? WC : 0, hi2 = mode == 1 ? WS : 0;
    const int olo = final ? WB : WC, on = final ? (WA - WB) : (WS - WC);
    fetch(0, lo, hi, lo2, hi2);
    for (int k = 0; k <= N; ++k) {
      const int buf = k & 1;
      w.load_wait(buf);
      if (k < N) fetch(k + 1, lo, hi, lo2, hi2);
      w.store_wait(1);
      const double* rec = w.inbuf(buf);
      const double* wk = rec + REC;
      double* out = w.outbuf(buf) - olo;
      Stage s;
      stage_consts(k, rec, s);
      double* ys = scr + S_YS;
      if (h == 0 && i >= 5 && i < 15) ys[i] = dxi;
      w.sync();
      // multiplier step of the link k-1 -> k:  dpi = P_k dx_k + p_k
      if (final) {
        double v = 0.0;
        if (k > 0 && i >= 5 && i < 15) {
          const double* Pr = wk + WG + (i - 5) * 10 + 5 * h;
#pragma unroll
          for (int c = 0; c < 5; ++c) v += Pr[c] * ys[5 + 5 * h + c];
        }
        v += w.shfl_xor(v, 16);
        if (h == 0 && i >= 5 && i < 15) out[B_DPIM + i - 5] = k > 0 ? v + wk[WE + i] : 0.0;
      }
      // control step: du = -L^-T (D^-1 l~ + T_x' dx)
      double du = 0.0;
      if (k < N) {
        double wj = 0.0;
        if (i < 5) {
#pragma unroll
          for (int r = 0; r < 5; ++r) wj += wk[WF + (5 + 5 * h + r) * 5 + i] * ys[5 + 5 * h + r];
        }
        wj += w.shfl_xor(wj, 16);
        if (i < 5) du = -(wj + wk[WF + i * 5 + i] * wk[WE + i]);
#pragma unroll
        for (int c = 4; c >= 1; --c) {
          const double duc = w.shfl(du, c);
          if (i < c) du -= wk[WF + c * 5 + i] * duc;
        }
      }
      const double dz = i < 5 ? du : (i < 15 ? dxi : 0.0);
      if (h == 0) ys[i] = dz;
      w.sync();
      // ---- slots ----
      double az, adz;
      gdot2(rec, wk + A_Z, ys, az, adz);
      const double z = i < 15 ? wk[A_Z + i] : 0.0;
      const double azA = i < 5 ? az : z, adzA = i < 5 ? adz : dz;
      const double azB = az, adzB = adz;
      const bool softl = (i == 11 && s.soft);
      const double lamA = s.hasA ? wk[A_LAM + s.slotA] : 0.0, tA = s.hasA ? wk[A_T + s.slotA] : 0.0;
      const double lamB = s.hasB ? wk[A_LAM + s.slotB] : 0.0, tB = s.hasB ? wk[A_T + s.slotB] : 0.0;
      const double prA = (mode == 1 && s.hasA) ? wk[C_PROD + s.slotA] : 0.0, prB = (mode == 1 && s.hasB) ? wk[C_PROD + s.slotB] : 0.0;
      Side SA, SB;
      side_terms(s.hasA, lamA, tA, azA, s.bndA, 0.0, mode, sigmu, prA, SA);
      side_terms(s.hasB, lamB, tB, azB, s.bndB, softl ? wk[A_SLK + h] : 0.0, mode, sigmu, prB, SB);
      double ds = 0.0, dts = 0.0, dls = 0.0;
      if (softl) {
        double Gc, cc;
        soft_terms(s.zpen, wk[A_SLK + 2 + h], wk[A_SLK + 4 + h], mode, sigmu, mode == 1 ? wk[C_SPROD + h] : 0.0, SB, Gc, cc);
        ds = -(SB.rgs + SB.c + SB.cs + sgn * SB.G * adzB) / (SB.G + SB.Gs);
        dts = ds - SB.rsl;
        dls = -(SB.rms + SB.ls * dts) / SB.ts;
        if (dls < 0.0) alpha = fmin(alpha, -SB.ls / dls);
        if (dts < 0.0) alpha = fmin(alpha, -SB.ts / dts);
        s_lin += SB.ls * dts + SB.ts * dls;
        s_quad += dls * dts;
      }
      double dtA = 0.0, dlA = 0.0, dtB = 0.0, dlB = 0.0;
      if (s.hasA) {
        dtA = sgn * adzA - SA.r;
        dlA = -(SA.rm + lamA * dtA) * SA.it;
        if (dlA < 0.0) alpha = fmin(alpha, -lamA / dlA);
        if (dtA < 0.0) alpha = fmin(alpha, -tA / dtA);
        s_lin += lamA * dtA + tA * dlA; s_quad += dlA * dtA;
      }
      if (s.hasB) {
        dtB = sgn * adzB + ds - SB.r;
        dlB = -(SB.rm + lamB * dtB) * SB.it;
        if (dlB < 0.0) alpha = fmin(alpha, -lamB / dlB);
        if (dtB < 0.0) alpha = fmin(alpha, -tB / dtB);
        s_lin += lamB * dtB + tB * dlB; s_quad += dlB * dtB;
      }
      if (store_prod) {
        if (i < 15) out[C_PROD + s.slotA] = dlA * dtA;
        if (i >= 5 && i <= 11) out[C_PROD + s.slotB] = dlB * dtB;
        if (i == 11) out[C_SPROD + h] = dls * dts;
      }
      if (final) {
        if (i < 15) { out[B_DLAM + s.slotA] = dlA; out[B_DTT + s.slotA] = dtA; }
        if (i >= 5 && i <= 11) { out[B_DLAM + s.slotB] = dlB; out[B_DTT + s.slotB] = dtB; }
        if (i == 11) { out[B_DSLK + h] = ds; out[B_DSLK + 2 + h] = dls; out[B_DSLK + 4 + h] = dts; }
        if (h == 0) out[B_DZ + i] = dz;
      }
      // ---- next state ----
      if (k < N) {
        if (i >= 5 && i < 15) {
          const int j = i < 10 ? i - 5 : i - 10;
          dxi = (i < 10 ? ys[5 + j] + dt * ys[10 + j] + a2 * ys[j] : ys[10 + j] + dt * ys[j]) + wk[D_RB + i - 5];
        }
      }
      w.store(gws_k(k) + olo, w.outbuf(buf), on);
      w.sync();      // ys is rewritten by the next stage
    }
    w.store_wait(0);
    StepStats o;
    o.alpha = wmin(alpha); o.s_lin = wsum(s_lin); o.s_quad = wsum(s_quad);
    return o;
  }

  // =====================================================================================================
  // driver (same control flow as the oracle's QpIpm::solve)
  // =====================================================================================================
  SMPC_HD QpResult solve() {
    QpResult R;
    R.iter = 0; R.status = 0; R.mu = 0.0;
    for (int q = 0; q < 4; ++q) R.res[q] = 0.0;
    double alpha = 1.0, step = 0.0;
    int kk = 0;
    bool nan = false;
    for (;; ++kk) {
      update_factorize(kk == 0, step, R);
      nan = (R.res[0] != R.res[0]) || (R.res[1] != R.res[1]) || (R.res[2] != R.res[2]) || (R.res[3] != R.res[3]);
      if (nan && kk > 0) break;
      const bool unconv = (R.res[0] > P.qp_tol_stat) || (R.res[1] > P.qp_tol_eq) || (R.res[2] > P.qp_tol_ineq) || (R.res[3] > P.qp_tol_comp);
      if (!unconv && !nan) break;
      if (kk >= P.qp_iter_max) break;
      if (!(alpha > P.qp_alpha_min)) break;
      const StepStats aff = forward(0, 0.0, true, false);
      const double mu_aff = R.mu + (aff.alpha * aff.s_lin + aff.alpha * aff.alpha * aff.s_quad) / nc;
      double sigma = mu_aff / R.mu; sigma = sigma * sigma * sigma;
      const double sigmu = sigma * R.mu;
      resolve_backward(1, sigmu);
      StepStats cor = forward(1, sigmu, false, true);
      alpha = cor.alpha;
      if (P.qp_cond_pred_corr) {
        const double mu_corr = R.mu + (alpha * cor.s_lin + alpha * alpha * cor.s_quad) / nc;
        if (mu_corr > 2.0 * mu_aff) {
          resolve_backward(2, sigmu);
          cor = forward(2, sigmu, false, true);
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
