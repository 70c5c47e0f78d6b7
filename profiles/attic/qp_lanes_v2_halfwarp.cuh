// Stage-structured interior-point QP solve of one RTI iteration, one problem per 16-lane group (half warp).
//
// Replaces the HPIPM call inside AcadosOcpSolver.solve() (reference controller.py:158; options :97-110,208-209;
// algorithm: Frison & Diehl, HPIPM, IFAC 2020 -- Mehrotra predictor-corrector IPM, inequality rows condensed into
// the stage Hessian, backward Riccati factorisation / forward substitution).  Layout of the computation:
//   * lane c (0..14) owns column c of the 15x15 condensed stage matrix ([du(5); dq(5); dv(5)] order) plus the gradient
//     entry c; the constant double-integrator A, B (env_model.py:63-71) are never stored: [B A]' P [B A] is formed from
//     lane-to-lane shuffles of the 5x5 blocks of P;
//   * the Cholesky elimination of the 5 control columns is a sequence of shuffle broadcasts + rank-1 updates on the lane
//     columns; each lane keeps its row of the factor for the vector-only re-solves (corrector / centering);
//   * inequality rows are owned by lanes: lanes 0-4 a torque row, lanes 5-14 the box row of their state, lanes 5-10 also
//     a capsule row, lane 11 the viability row (+ its slacks);
//   * all per-stage data of a problem (iterate, multipliers, factor, step) is ONE contiguous block per stage in global
//     memory; every sweep streams the stage record + the sub-range of the block it needs into shared memory with a
//     double-buffered asynchronous bulk copy (TMA 1-D, mbarrier completion) issued one stage ahead, so the
//     sequential stage recursion never waits on HBM latency; results are stored straight from registers;
//   * the primal-dual update of iteration i is fused into the factorisation sweep of iteration i+1 (both walk the
//     stages backwards), which removes one full pass over the data per iteration.
// The code is written against a small `Lanes` policy (lane id, shuffle, group barrier, scratch, async stage fetch) so
// that the same source runs on the device (LanesDev, qp.cu) and, for kernel-logic tests without a GPU, on the host
// (tests/emu: 16 threads and a barrier).  The product only ever instantiates the device policy.
#pragma once
#include "dev_model.cuh"

namespace smpc {

constexpr int QL = 16;            // lanes per problem
constexpr int NSLOT = 4;          // constraint slots per lane: rowA lower/upper, rowB lower/upper

// ---- per-stage block in global memory (doubles); [..16] arrays are indexed by lane, [..64] by slot*16+lane ----
enum {
  O_DZ = 0,      // 16  step: primal
  O_DPI = 16,    // 16  step: multipliers of the dynamics k -> k+1 (lanes 5..14)
  O_DPIM = 32,   // 16  step: multipliers of the dynamics k-1 -> k (copy kept with stage k)
  O_DLAM = 48,   // 64
  O_DTT = 112,   // 64
  O_Z = 176,     // 16  iterate
  O_PI = 192,    // 16
  O_PIM = 208,   // 16  pi_{k-1}
  O_LAM = 224,   // 64
  O_T = 288,     // 64
  O_AUX = 352,   // 16  slacks of the soft row: s_l s_u lam_sl lam_su t_sl t_su ds_l ds_u dlam_sl dlam_su dt_sl dt_su prod_sl prod_su
  O_PROD = 368,  // 64  dlam_aff * dt_aff
  O_GB = 432,    // 16  res_g
  O_WV = 448,    // 16  P_{k+1} res_b_k
  O_RB = 464,    // 16  res_b_k
  O_PV = 480,    // 16  l (lanes 0..4) and p (lanes 5..14) of the current solve
  O_FAC = 496,   // 80  rows of [Lr; Ls]: fac[j*16 + lane] = L[lane][j]
  O_PM = 576,    // 160 Riccati matrix: Pm[r*16 + lane] = P[r][lane-5]
  QP_ST = 736
};
constexpr int QP_L0 = 320;                    // stage-0 state factor (columns, rows), after the stage blocks
constexpr int QP_STATE_MAX = QP_ST - O_Z;     // largest state range a sweep fetches (560 doubles)
constexpr int QP_BUF = REC + QP_STATE_MAX;    // one staging buffer: stage record + state range
constexpr int QP_VEC = 96;                    // replicated-vector scratch
constexpr int QP_SMEM_PER_GROUP = 2 * QP_BUF + QP_VEC;   // doubles

struct QpMem {
  const double* rec;   // [N+1][REC]
  const double* x0;    // [10]
  double* st;          // [N+1][QP_ST] + [QP_L0]
  int r;               // receding index of the problem (RealReceding box override)
};
constexpr size_t qp_doubles_per_problem(int N) { return (size_t)(N + 1) * QP_ST + QP_L0; }

struct QpResult {
  int iter, status;          // status: 0 success, 1 max iter, 2 min step, 3 NaN
  double res[4], mu;
};

template <class L>
struct QpSolver {
  L& ln;
  const smpc_problem_t& P;
  const QpMem& M;
  const int N, lane;
  double* V;          // replicated-vector scratch (QP_VEC doubles)
  const double* S;    // current stage record (in the staging buffer)
  const double* T;    // current state range, biased so that T[O_xxx + i] addresses field O_xxx
  int nc;
  double dt, hdt2;

  SMPC_HD QpSolver(L& l, const smpc_problem_t& p, const QpMem& m)
      : ln(l), P(p), M(m), N(p.N), lane(l.lane()), V(l.scratch()), S(nullptr), T(nullptr), nc(0), dt(p.dt), hdt2(0.5 * p.dt * p.dt) {}

  // ---- small group helpers ----
  SMPC_HD double gsum(double v) { for (int o = 8; o > 0; o >>= 1) v += ln.shfl_xor(v, o); return v; }
  SMPC_HD double gmin(double v) { for (int o = 8; o > 0; o >>= 1) v = fmin(v, ln.shfl_xor(v, o)); return v; }
  SMPC_HD double gmax_nan(double v) {   // max that propagates NaN
    for (int o = 8; o > 0; o >>= 1) { double w = ln.shfl_xor(v, o); v = (v != v || w != w) ? (v + w) : fmax(v, w); }
    return v;
  }
  // replicate a lane-distributed 16-vector into scratch slot `off` (V[off + c] = value of lane c)
  SMPC_HD void publish(int off, double v) { ln.sync(); V[off + lane] = v; ln.sync(); }

  // ---- stage streaming ----
  SMPC_HD double* gst(int k) const { return M.st + (size_t)k * QP_ST; }
  SMPC_HD void fetch(int k, int buf, int lo, int cnt) { ln.fetch(buf, M.rec + (size_t)k * REC, gst(k) + lo, cnt); }
  SMPC_HD void acquire(int buf, int lo) { ln.wait(buf); S = ln.buffer(buf); T = S + REC - lo; }

  // ---- row ownership ----
  SMPC_HD bool hasA() const { return lane < 5 ? (S[SMPC_REC_NTAU] > 0.5) : (lane < 15); }
  SMPC_HD bool hasB() const { return (lane >= 5 && lane <= 10) ? (S[SMPC_REC_NDIST] > 0.5) : (lane == 11 ? S[SMPC_REC_NNROW] > 0.5 : false); }
  SMPC_HD bool softB() const { return lane == 11 && S[SMPC_REC_NNROW] > 0.5 && S[SMPC_REC_SOFT] >= 0.0; }
  // canonical row id (box 0-9, tau 10-14, dist 15-20, nn 21)
  SMPC_HD int idA() const { return lane < 5 ? 10 + lane : lane - 5; }
  SMPC_HD int idB() const { return lane == 11 ? 21 : 15 + (lane - 5); }

  // a_rowA . y and a_rowB . y for a replicated 15-vector y ([u q v] order)
  SMPC_HD double dotA(const double* y) const {
    if (lane < 5) {
      double r = 0.0;
      const double* a = S + SMPC_REC_JTAU + lane * 15;
#pragma unroll
      for (int i = 0; i < 15; ++i) r += a[i] * y[i];
      return r;
    }
    return lane < 15 ? y[lane] : 0.0;
  }
  SMPC_HD double dotB(const double* y) const {
    double r = 0.0;
    if (lane >= 5 && lane <= 10) {
      const double* a = S + SMPC_REC_JDIST + (lane - 5) * 5;
#pragma unroll
      for (int i = 0; i < 5; ++i) r += a[i] * y[5 + i];
    } else if (lane == 11) {
      const double* a = S + SMPC_REC_JNN;
#pragma unroll
      for (int i = 0; i < 10; ++i) r += a[i] * y[5 + i];
    }
    return r;
  }
  SMPC_HD void boundsA(int k, double& lo, double& hi) const {
    if (lane < 5) { lo = P.tau_min[lane] - S[SMPC_REC_TAU + lane]; hi = P.tau_max[lane] - S[SMPC_REC_TAU + lane]; return; }
    if (lane == 15) { lo = -1.0; hi = 1.0; return; }
    const int i = lane - 5;
    const double xk = S[SMPC_REC_X + i];
    if (k == 0) { lo = hi = M.x0[i] - xk; return; }
    if (k == N) { lo = P.lbx_e[i] - xk; hi = P.ubx_e[i] - xk; return; }
    if (P.controller == SMPC_CTRL_REAL_RECEDING) {
      if (k == M.r) { const double c = M.rec[(size_t)(k + 1) * REC + SMPC_REC_X + i]; lo = c - 1e-3 - xk; hi = c + 1e-3 - xk; }
      else { lo = P.x_min[i] - xk; hi = P.x_max[i] - xk; }
      return;
    }
    lo = P.lbx[i] - xk; hi = P.ubx[i] - xk;
  }
  SMPC_HD void boundsB(double& lo, double& hi) const {
    if (lane >= 5 && lane <= 10) { const int p = lane - 5; lo = P.pair_lo_ocp[p] - S[SMPC_REC_DIST + p]; hi = P.pair_hi - S[SMPC_REC_DIST + p]; }
    else { lo = 0.0 - S[SMPC_REC_NN]; hi = 1e6 - S[SMPC_REC_NN]; }
  }

  // ---- per-lane slot state of one stage ----
  struct Slots {
    bool pa, pb, soft;
    double lam[NSLOT], t[NSLOT], r[NSLOT];   // r = res_d
    double sl, su, lsl, lsu, tsl, tsu, rsl, rsu, rgsl, rgsu;   // slack data (lane 11, soft only)
    double aA, aB;                            // row products with z
  };

  // res_d of the own slots for the replicated iterate zr, given lam/t (and the slack entries) already in `s`
  SMPC_HD void eval_slots(int k, const double* zr, Slots& s) {
    s.aA = dotA(zr); s.aB = dotB(zr);
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) s.r[i] = 0.0;
    if (s.soft) {
      s.rsl = s.tsl - s.sl; s.rsu = s.tsu - s.su;
      s.rgsl = S[SMPC_REC_SOFT] - s.lam[2] - s.lsl;
      s.rgsu = S[SMPC_REC_SOFT] - s.lam[3] - s.lsu;
    }
    if (s.pa) { double lo, hi; boundsA(k, lo, hi); s.r[0] = s.t[0] - (s.aA - lo); s.r[1] = s.t[1] - (hi - s.aA); }
    if (s.pb) { double lo, hi; boundsB(lo, hi); s.r[2] = s.t[2] - (s.aB + s.sl - lo); s.r[3] = s.t[3] - (hi - s.aB + s.su); }
  }
  // loads lam/t from the staged block, then eval_slots
  SMPC_HD void load_slots(int k, const double* zr, Slots& s) {
    s.pa = hasA(); s.pb = hasB(); s.soft = softB();
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) { s.lam[i] = T[O_LAM + i * 16 + lane]; s.t[i] = T[O_T + i * 16 + lane]; }
    s.sl = s.su = 0.0;
    if (s.soft) { const double* a = T + O_AUX; s.sl = a[0]; s.su = a[1]; s.lsl = a[2]; s.lsu = a[3]; s.tsl = a[4]; s.tsu = a[5]; }
    eval_slots(k, zr, s);
  }

  // complementarity right-hand side:  mode 0 affine, 1 corrector, 2 centering
  SMPC_HD double rm_of(int mode, double lam, double t, double prod, double sigmu) const {
    return mode == 0 ? lam * t : (mode == 1 ? lam * t + prod - sigmu : lam * t - sigmu);
  }

  // per-row condensation terms (Gamma, gamma) and nu = lam_hi - lam_lo of the own rows
  struct RowT { double GA, gA, nA, GB, gB, nB; };
  SMPC_HD RowT row_terms(const Slots& s, int mode, double sigmu) {
    RowT o; o.GA = o.gA = o.nA = o.GB = o.gB = o.nB = 0.0;
    const double* prod = T + O_PROD;     // only dereferenced in corrector mode (then part of the staged range)
    if (s.pa) {
      const double rl = rm_of(mode, s.lam[0], s.t[0], mode == 1 ? prod[0 * 16 + lane] : 0.0, sigmu);
      const double ru = rm_of(mode, s.lam[1], s.t[1], mode == 1 ? prod[1 * 16 + lane] : 0.0, sigmu);
      const double cl = (rl - s.lam[0] * s.r[0]) / s.t[0], cu = (ru - s.lam[1] * s.r[1]) / s.t[1];
      o.GA = s.lam[0] / s.t[0] + s.lam[1] / s.t[1];
      o.gA = cl - cu;
      o.nA = s.lam[1] - s.lam[0];
    }
    if (s.pb) {
      const double rl = rm_of(mode, s.lam[2], s.t[2], mode == 1 ? prod[2 * 16 + lane] : 0.0, sigmu);
      const double ru = rm_of(mode, s.lam[3], s.t[3], mode == 1 ? prod[3 * 16 + lane] : 0.0, sigmu);
      double cl = (rl - s.lam[2] * s.r[2]) / s.t[2], cu = (ru - s.lam[3] * s.r[3]) / s.t[3];
      double Gl = s.lam[2] / s.t[2], Gu = s.lam[3] / s.t[3];
      if (s.soft) {
        const double* a = T + O_AUX;
        const double rsl = rm_of(mode, s.lsl, s.tsl, mode == 1 ? a[12] : 0.0, sigmu);
        const double rsu = rm_of(mode, s.lsu, s.tsu, mode == 1 ? a[13] : 0.0, sigmu);
        const double Gsl = s.lsl / s.tsl, Gsu = s.lsu / s.tsu;
        const double csl = (rsl - s.lsl * s.rsl) / s.tsl, csu = (rsu - s.lsu * s.rsu) / s.tsu;
        const double Wl = 1.0 / (Gl + Gsl), Wu = 1.0 / (Gu + Gsu);
        cl = cl - Gl * Wl * (s.rgsl + cl + csl);
        cu = cu - Gu * Wu * (s.rgsu + cu + csu);
        Gl = Gl * Gsl * Wl;
        Gu = Gu * Gsu * Wu;
      }
      o.GB = Gl + Gu; o.gB = cl - cu; o.nB = s.lam[3] - s.lam[2];
    }
    return o;
  }

  // publish per-row values in canonical row order: V[off + id]
  SMPC_HD void publish_rows(int off, bool pa, double a, bool pb, double b) {
    ln.sync();
    if (lane < 15) V[off + idA()] = pa ? a : 0.0;
    if (lane >= 5 && lane <= 11) V[off + idB()] = pb ? b : 0.0;
    ln.sync();
  }

  // sum_rows a_r[c] * w_r for this lane's column c, w in canonical order at V[off..]
  SMPC_HD double rowsT(int off) const {
    if (lane == 15) return 0.0;
    const double* w = V + off;
    double r = 0.0;
    if (S[SMPC_REC_NTAU] > 0.5) {
#pragma unroll
      for (int i = 0; i < 5; ++i) r += S[SMPC_REC_JTAU + i * 15 + lane] * w[10 + i];
    }
    if (lane >= 5) {
      r += w[lane - 5];                                   // box row of this state
      if (lane < 10 && S[SMPC_REC_NDIST] > 0.5) {
#pragma unroll
        for (int p = 0; p < 6; ++p) r += S[SMPC_REC_JDIST + p * 5 + (lane - 5)] * w[15 + p];
      }
      if (S[SMPC_REC_NNROW] > 0.5) r += S[SMPC_REC_JNN + (lane - 5)] * w[21];
    }
    return r;
  }

  // ([B A]' y)[c] for a replicated 10-vector y
  SMPC_HD double dynT(const double* y) const {
    if (lane < 5) return hdt2 * y[lane] + dt * y[5 + lane];
    if (lane < 10) return y[lane - 5];
    if (lane < 15) return dt * y[lane - 10] + y[lane - 5];
    return 0.0;
  }

  // (H z + g)[c]
  SMPC_HD double cost_grad(int k, const double* zr) const {
    if (lane == 15) return 0.0;
    if (lane < 5) return k == N ? 0.0 : S[SMPC_REC_HU] * zr[lane] + S[SMPC_REC_G + lane];
    if (lane < 10) {
      const int i = lane - 5;
      double r = S[SMPC_REC_G + lane] + S[SMPC_REC_HQ] * zr[lane];
#pragma unroll
      for (int j = 0; j < 5; ++j) { const int a = i > j ? i : j, b = i > j ? j : i; r += S[SMPC_REC_HQQ + a * (a + 1) / 2 + b] * zr[5 + j]; }
      return r;
    }
    return S[SMPC_REC_G + lane] + S[SMPC_REC_HV] * zr[lane];
  }

  // ------------------------------------------------------------------------------------ S0: cold start
  // writes the iterate block (z, pi, pim, lam, t, aux) and a zero step block; counts the constraints
  SMPC_HD void init(double mu0, double thr0) {
    int cnt = 0;
    int buf = 0;
    fetch(N, buf, O_Z, 0);
    for (int k = N; k >= 0; --k) {
      if (k > 0) fetch(k - 1, buf ^ 1, O_Z, 0);
      acquire(buf, O_Z);
      // box rows: move the primal inside (lanes 5..14)
      double zc = 0.0, tl = 0.0, tu = 0.0, loA = 0, hiA = 0;
      if (lane >= 5 && lane < 15) {
        boundsA(k, loA, hiA);
        tl = zc - loA; tu = hiA - zc;
        if (tl < thr0) {
          if (tu < thr0) { zc = 0.5 * (loA + hiA); tl = thr0; tu = thr0; }
          else { tl = thr0; zc = loA + thr0; }
        } else if (tu < thr0) { tu = thr0; zc = hiA - thr0; }
      }
      publish(0, zc);
      const double* zr = V;
      const bool pa = hasA(), pb = hasB(), soft = softB();
      double t[NSLOT] = {0, 0, 0, 0};
      if (lane < 5 && pa) {
        boundsA(k, loA, hiA);
        const double v = dotA(zr);
        tl = fmax(thr0, v - loA); tu = fmax(thr0, hiA - v);
      }
      if (pa) { t[0] = tl; t[1] = tu; }
      if (pb) {
        double lo, hi; boundsB(lo, hi);
        const double v = dotB(zr);
        t[2] = fmax(thr0, v - lo); t[3] = fmax(thr0, hi - v);
      }
      double* g = gst(k);
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const bool p = i < 2 ? pa : pb;
        g[O_LAM + i * 16 + lane] = p ? mu0 / t[i] : 0.0;
        g[O_T + i * 16 + lane] = t[i];
        g[O_DLAM + i * 16 + lane] = 0.0;
        g[O_DTT + i * 16 + lane] = 0.0;
        if (p) ++cnt;
      }
      double av = 0.0;
      if (soft) { av = (lane == 0 || lane == 1 || lane == 4 || lane == 5) ? thr0 : ((lane == 2 || lane == 3) ? mu0 / thr0 : 0.0); }
      // aux is indexed by field, not by lane: lane i writes field i (soft flag is a lane-11 property -> broadcast it)
      const bool stage_soft = ln.shfl(soft ? 1.0 : 0.0, 11) > 0.5;
      av = stage_soft ? ((lane == 0 || lane == 1 || lane == 4 || lane == 5) ? thr0 : ((lane == 2 || lane == 3) ? mu0 / thr0 : 0.0)) : 0.0;
      g[O_AUX + lane] = av;
      if (stage_soft && lane == 11) cnt += 2;
      g[O_Z + lane] = zc; g[O_PI + lane] = 0.0; g[O_PIM + lane] = 0.0;
      g[O_DZ + lane] = 0.0; g[O_DPI + lane] = 0.0; g[O_DPIM + lane] = 0.0;
      ln.sync();
      buf ^= 1;
    }
    nc = (int)(gsum((double)cnt) + 0.5);
    ln.flush();
  }

  // ------------------------------------------------- S1: (update of the previous step) + factorisation + affine gradient
  // Applies  w <- w + a*dw  (a = 0 right after init), evaluates all residual norms and mu of the new iterate, factorises.
  // Leaves dx_0 (affine) replicated at V[64+5 .. 64+15).
  SMPC_HD void update_factorize(double a, double lam_min, double t_min, double reg, QpResult& R) {
    double ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, musum = 0.0;
    double pp[10];            // column (lane-5) of P_{k+1}   (lanes 5..14)
    double pcur = 0.0;        // p_{k+1}[lane-5]
    double zxn[10];           // z_x of stage k+1, replicated
#pragma unroll
    for (int i = 0; i < 10; ++i) { pp[i] = 0.0; zxn[i] = 0.0; }
    int buf = 0;
    fetch(N, buf, 0, O_PROD);
    for (int k = N; k >= 0; --k) {
      if (k > 0) fetch(k - 1, buf ^ 1, 0, O_PROD);
      acquire(buf, 0);
      double* g = gst(k);
      // ---- primal-dual update of this stage ----
      const double znew = T[O_Z + lane] + a * T[O_DZ + lane];
      const double pinew = T[O_PI + lane] + a * T[O_DPI + lane];
      const double pimnew = T[O_PIM + lane] + a * T[O_DPIM + lane];
      g[O_Z + lane] = znew; g[O_PI + lane] = pinew; g[O_PIM + lane] = pimnew;
      publish(0, znew);                                        // V[0..15) = z_k
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      Slots s;
      s.pa = hasA(); s.pb = hasB(); s.soft = softB();
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const bool p = i < 2 ? s.pa : s.pb;
        s.lam[i] = 0.0; s.t[i] = 0.0;
        if (p) {
          s.lam[i] = fmax(T[O_LAM + i * 16 + lane] + a * T[O_DLAM + i * 16 + lane], lam_min);
          s.t[i] = fmax(T[O_T + i * 16 + lane] + a * T[O_DTT + i * 16 + lane], t_min);
          g[O_LAM + i * 16 + lane] = s.lam[i]; g[O_T + i * 16 + lane] = s.t[i];
          const double c = s.lam[i] * s.t[i];
          musum += c;
          nm = (c != c) ? c : fmax(nm, fabs(c));
        }
      }
      s.sl = s.su = 0.0;
      if (s.soft) {
        const double* aw = T + O_AUX;
        s.sl = aw[0] + a * aw[6]; s.su = aw[1] + a * aw[7];
        s.lsl = fmax(aw[2] + a * aw[8], lam_min); s.lsu = fmax(aw[3] + a * aw[9], lam_min);
        s.tsl = fmax(aw[4] + a * aw[10], t_min); s.tsu = fmax(aw[5] + a * aw[11], t_min);
        g[O_AUX + 0] = s.sl; g[O_AUX + 1] = s.su; g[O_AUX + 2] = s.lsl; g[O_AUX + 3] = s.lsu; g[O_AUX + 4] = s.tsl; g[O_AUX + 5] = s.tsu;
        musum += s.lsl * s.tsl + s.lsu * s.tsu;
        nm = fmax(nm, fmax(fabs(s.lsl * s.tsl), fabs(s.lsu * s.tsu)));
        nd = fmax(nd, fmax(fabs(s.tsl - s.sl), fabs(s.tsu - s.su)));
      }
      eval_slots(k, zr, s);
      {
        const double rd = fmax(fmax(fabs(s.r[0]), fabs(s.r[1])), fmax(fabs(s.r[2]), fabs(s.r[3])));
        const double chk = s.r[0] + s.r[1] + s.r[2] + s.r[3];
        nd = (chk != chk) ? chk : fmax(nd, rd);
      }
      // ---- condensation terms (affine rhs: rm = lam * t) ----
      const RowT rt = row_terms(s, 0, 0.0);
      publish_rows(16, s.pa, rt.GA, s.pb, rt.GB);              // V[16..38) Gamma
      const double GamBox = (lane >= 5 && lane < 15) ? rt.GA : 0.0;
      // ---- column of the condensed matrix ----
      double m[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) m[i] = 0.0;
      if (lane < 5) { m[lane] = (k == N) ? 1.0 : S[SMPC_REC_HU] + reg; }
      else if (lane < 10) {
        const int i = lane - 5;
#pragma unroll
        for (int j = 0; j < 5; ++j) { const int aa = i > j ? i : j, bb = i > j ? j : i; m[5 + j] = S[SMPC_REC_HQQ + aa * (aa + 1) / 2 + bb]; }
        m[lane] += S[SMPC_REC_HQ] + reg + GamBox;
      } else if (lane < 15) { m[lane] = S[SMPC_REC_HV] + reg + GamBox; }
      if (lane < 15) {
        if (S[SMPC_REC_NTAU] > 0.5) {
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const double* ar = S + SMPC_REC_JTAU + r * 15;
            const double coef = V[16 + 10 + r] * ar[lane];
#pragma unroll
            for (int i = 0; i < 15; ++i) m[i] += coef * ar[i];
          }
        }
        if (lane >= 5 && lane < 10 && S[SMPC_REC_NDIST] > 0.5) {
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const double* ar = S + SMPC_REC_JDIST + p * 5;
            const double coef = V[16 + 15 + p] * ar[lane - 5];
#pragma unroll
            for (int i = 0; i < 5; ++i) m[5 + i] += coef * ar[i];
          }
        }
        if (lane >= 5 && S[SMPC_REC_NNROW] > 0.5) {
          const double* ar = S + SMPC_REC_JNN;
          const double coef = V[16 + 21] * ar[lane - 5];
#pragma unroll
          for (int i = 0; i < 10; ++i) m[5 + i] += coef * ar[i];
        }
      }
      // ---- stationarity residual and affine gradient ----
      publish_rows(40, s.pa, rt.nA, s.pb, rt.nB);              // V[40..62) nu = lam_hi - lam_lo
      double rg = cost_grad(k, zr) + rowsT(40);
      publish_rows(40, s.pa, rt.gA, s.pb, rt.gB);              // V[40..62) gamma
      double gv = rowsT(40);
      if (k < N) {
        publish(64, pinew);                                    // V[64+5 .. 64+15) = pi_k
        rg += dynT(V + 64 + 5);
      }
      if (k > 0 && lane >= 5 && lane < 15) rg -= pimnew;
      if (k == N && lane < 5) rg = 0.0;
      ng = (rg != rg) ? rg : fmax(ng, fabs(rg));
      g[O_GB + lane] = rg;
      gv += rg;
      if (k < N) {
        // res_b_k (replicated) and w = P_{k+1} res_b_k
        double rb[10];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          rb[i] = zr[5 + i] + dt * zr[10 + i] + hdt2 * zr[i] + S[SMPC_REC_B + i] - zxn[i];
          rb[5 + i] = zr[10 + i] + dt * zr[i] + S[SMPC_REC_B + 5 + i] - zxn[5 + i];
        }
        double w = 0.0;
        if (lane >= 5 && lane < 15) {
#pragma unroll
          for (int r = 0; r < 10; ++r) w += pp[r] * rb[r];
          const double mine = rb[0];   // placeholder, replaced below
          (void)mine;
        }
        double rbl = 0.0;
#pragma unroll
        for (int i = 0; i < 10; ++i) if (i == lane - 5) rbl = rb[i];
        if (lane >= 5 && lane < 15) { nb = (rbl != rbl) ? rbl : fmax(nb, fabs(rbl)); }
        g[O_RB + lane] = rbl;
        g[O_WV + lane] = w;
        publish(80, w + pcur);                                 // V[80+5..80+15) = P rb + p_{k+1}
        gv += dynT(V + 80 + 5);
        // M += [B A]' P_{k+1} [B A]
        const int j = lane % 5;
        double Pq[10], Pv[10];
#pragma unroll
        for (int r = 0; r < 10; ++r) { Pq[r] = ln.shfl(pp[r], 5 + j); Pv[r] = ln.shfl(pp[r], 10 + j); }
        if (lane < 15) {
          double W[10];
#pragma unroll
          for (int r = 0; r < 10; ++r) W[r] = lane < 5 ? hdt2 * Pq[r] + dt * Pv[r] : (lane < 10 ? Pq[r] : dt * Pq[r] + Pv[r]);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            m[i] += hdt2 * W[i] + dt * W[5 + i];
            m[5 + i] += W[i];
            m[10 + i] += dt * W[i] + W[5 + i];
          }
        }
      }
      m[15] = (k == N && lane < 5) ? 0.0 : gv;
      if (lane == 15) { for (int i = 0; i < 16; ++i) m[i] = 0.0; m[15] = 1.0; }
      // ---- Cholesky elimination: 5 control pivots (all 15 at stage 0) ----
      const int npiv = (k == 0) ? 15 : 5;
      double lrow[5] = {0, 0, 0, 0, 0};
      double l0row[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) l0row[i] = 0.0;
      for (int j = 0; j < npiv; ++j) {
        double d = 0.0;
#pragma unroll
        for (int i = 0; i < 15; ++i) if (i == j) d = m[i];
        d = ln.shfl(d, j);
        const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
        if (lane == j) {
#pragma unroll
          for (int i = 0; i < 16; ++i) m[i] = (i >= j) ? m[i] * inv : m[i];
        }
        double vr[16];
        double lc = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          vr[i] = ln.shfl(m[i], j);
          if (i == lane) lc = vr[i];
        }
        if (lane > j && lane < 15) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (i > j) m[i] -= vr[i] * lc;
        }
        const double lr = lane >= j ? lc : 0.0;
        if (j < 5) {
#pragma unroll
          for (int i = 0; i < 5; ++i) if (i == j) lrow[i] = lr;
        } else {
#pragma unroll
          for (int i = 0; i < 10; ++i) if (i == j - 5) l0row[i] = lr;
        }
      }
      // ---- store the factor ----
#pragma unroll
      for (int j = 0; j < 5; ++j) g[O_FAC + j * 16 + lane] = lrow[j];
      g[O_PV + lane] = m[15];
      if (k > 0) {
        if (lane >= 5 && lane < 15) {
#pragma unroll
          for (int r = 0; r < 10; ++r) pp[r] = m[5 + r];
          pcur = m[15];
        }
#pragma unroll
        for (int r = 0; r < 10; ++r) g[O_PM + r * 16 + lane] = (lane >= 5 && lane < 15) ? m[5 + r] : 0.0;
#pragma unroll
        for (int i = 0; i < 10; ++i) zxn[i] = zr[5 + i];
      } else {
        // stage 0: keep the state factor (columns and rows) and solve for dx_0
        double* L0 = M.st + (size_t)(N + 1) * QP_ST;
#pragma unroll
        for (int r = 0; r < 10; ++r) { L0[r * 16 + lane] = m[5 + r]; L0[(10 + r) * 16 + lane] = l0row[r]; }
        solve_dx0(m, m[15]);
      }
      ln.sync();
      buf ^= 1;
    }
    R.res[0] = gmax_nan(ng); R.res[1] = gmax_nan(nb); R.res[2] = gmax_nan(nd); R.res[3] = gmax_nan(nm);
    R.mu = gsum(musum) / nc;
    ln.flush();
  }

  // back substitution with the stage-0 state factor held column-wise in mcol[5..15): dx_0 -> V[64+5 .. 64+15), replicated
  SMPC_HD void solve_dx0(const double* mcol, double lx) {
    double acc = -lx;
    double mine = 0.0;
    for (int r = 14; r >= 5; --r) {
      double dr = 0.0, lrc = 0.0;
#pragma unroll
      for (int i = 5; i < 15; ++i) if (i == r) lrc = mcol[i];     // L0[r][lane] (valid for r >= lane)
      if (lane == r) dr = lrc > 0.0 ? acc / lrc : 0.0;
      dr = ln.shfl(dr, r);
      if (lane == r) mine = dr;
      if (lane >= 5 && lane < r) acc -= lrc * dr;
    }
    publish(64, mine);
  }

  // ------------------------------------------------- S3: vector-only backward recursion (corrector / centering)
  SMPC_HD void resolve_backward(int mode, double sigmu) {
    double pcur = 0.0;
    int buf = 0;
    const int cnt = O_PM - O_Z;
    fetch(N, buf, O_Z, cnt);
    for (int k = N; k >= 0; --k) {
      if (k > 0) fetch(k - 1, buf ^ 1, O_Z, cnt);
      acquire(buf, O_Z);
      double* g = gst(k);
      publish(0, T[O_Z + lane]);
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      Slots s;
      load_slots(k, zr, s);
      const RowT rt = row_terms(s, mode, sigmu);
      publish_rows(40, s.pa, rt.gA, s.pb, rt.gB);
      double gv = T[O_GB + lane] + rowsT(40);
      if (k < N) {
        const double w = (lane >= 5 && lane < 15) ? T[O_WV + lane] : 0.0;
        publish(80, w + pcur);
        gv += dynT(V + 80 + 5);
      }
      if (k == N && lane < 5) gv = 0.0;
      if (lane == 15) gv = 0.0;
      // forward elimination with the stored rows of [Lr; Ls]
      double lrow[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) lrow[j] = T[O_FAC + j * 16 + lane];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double lj = 0.0;
        if (lane == j) { lj = lrow[j] > 0.0 ? gv / lrow[j] : 0.0; gv = lj; }
        lj = ln.shfl(lj, j);
        if (lane > j && lane < 15) gv -= lrow[j] * lj;
      }
      g[O_PV + lane] = gv;
      if (k > 0) { pcur = (lane >= 5 && lane < 15) ? gv : 0.0; }
      else {
        // stage 0: forward then backward substitution with the state factor
        const double* L0 = M.st + (size_t)(N + 1) * QP_ST;
        double l0row[10], mcol[16];
#pragma unroll
        for (int r = 0; r < 10; ++r) { mcol[5 + r] = L0[r * 16 + lane]; l0row[r] = L0[(10 + r) * 16 + lane]; }
#pragma unroll
        for (int i = 0; i < 5; ++i) mcol[i] = 0.0;
        mcol[15] = 0.0;
        for (int j = 5; j < 15; ++j) {
          double lj = 0.0, ljj = 0.0;
#pragma unroll
          for (int i = 0; i < 10; ++i) if (i == j - 5) ljj = l0row[i];
          if (lane == j) { lj = ljj > 0.0 ? gv / ljj : 0.0; gv = lj; }
          lj = ln.shfl(lj, j);
          if (lane > j && lane < 15) gv -= ljj * lj;
        }
        solve_dx0(mcol, gv);
      }
      ln.sync();
      buf ^= 1;
    }
    ln.flush();
  }

  // ------------------------------------------------------------------------- S2 / S4: forward substitution
  struct StepStats { double alpha, s_lin, s_quad; };
  // mode: complementarity rhs as in rm_of; store_prod: keep dlam*dt (affine pass); final: store the step (dz, dpi, dlam, dt)
  SMPC_HD StepStats forward(int mode, double sigmu, bool store_prod, bool final) {
    double alpha = 1.0, s_lin = 0.0, s_quad = 0.0;
    double dx[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) dx[i] = V[64 + 5 + i];
    ln.sync();
    int buf = 0;
    const int cnt = (final ? QP_ST : O_PM) - O_Z;
    fetch(0, buf, O_Z, cnt);
    for (int k = 0; k <= N; ++k) {
      if (k < N) fetch(k + 1, buf ^ 1, O_Z, cnt);
      acquire(buf, O_Z);
      double* g = gst(k);
      // multiplier step of the dynamics k-1 -> k:  dpi_{k-1} = P_k dx_k + p_k   (stored with both stages)
      if (final && k > 0) {
        double dp = 0.0;
        if (lane >= 5 && lane < 15) {
          dp = T[O_PV + lane];
#pragma unroll
          for (int r = 0; r < 10; ++r) dp += T[O_PM + r * 16 + lane] * dx[r];
        }
        g[O_DPIM + lane] = dp;
        gst(k - 1)[O_DPI + lane] = dp;
      }
      if (final && k == 0) g[O_DPIM + lane] = 0.0;
      if (final && k == N) g[O_DPI + lane] = 0.0;
      // ---- du_k ----
      double du[5] = {0, 0, 0, 0, 0};
      if (k < N) {
        double lrow[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) lrow[j] = T[O_FAC + j * 16 + lane];
        const double lp = T[O_PV + lane];
        double sj[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          double term = 0.0;
          if (lane >= 5 && lane < 15) {
            double dxl = 0.0;
#pragma unroll
            for (int i = 0; i < 10; ++i) if (i == lane - 5) dxl = dx[i];
            term = lrow[j] * dxl;
          }
          if (lane == j) term = lp;
          sj[j] = gsum(term);
        }
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) if (lane == j) acc = -sj[j];
        for (int r = 4; r >= 0; --r) {
          double dr = 0.0, lrr = 0.0;
#pragma unroll
          for (int i = 0; i < 5; ++i) if (i == r) lrr = lrow[i];
          if (lane == r) dr = lrr > 0.0 ? acc / lrr : 0.0;
          dr = ln.shfl(dr, r);
#pragma unroll
          for (int i = 0; i < 5; ++i) if (i == r) du[i] = dr;
          for (int c = 0; c < r; ++c) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) if (i == c) v = lrow[i];
            v = ln.shfl(v, r);                    // L[r][c]
            if (lane == c) acc -= v * dr;
          }
        }
      }
      // replicated step of this stage
      double dzr[15];
#pragma unroll
      for (int i = 0; i < 5; ++i) { dzr[i] = du[i]; dzr[5 + i] = dx[i]; dzr[10 + i] = dx[5 + i]; }
      double mydz = 0.0;
#pragma unroll
      for (int i = 0; i < 15; ++i) if (i == lane) mydz = dzr[i];
      // ---- own slots ----
      publish(0, T[O_Z + lane]);
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      Slots s;
      load_slots(k, zr, s);
      const double adA = dotA(dzr), adB = dotB(dzr);
      const double* prod = T + O_PROD;
      double dlam[NSLOT] = {0, 0, 0, 0}, dtt[NSLOT] = {0, 0, 0, 0};
      double dsl = 0.0, dsu = 0.0;
      if (s.soft) {
        const double* aw = T + O_AUX;
        const double rl = rm_of(mode, s.lam[2], s.t[2], mode == 1 ? prod[2 * 16 + lane] : 0.0, sigmu);
        const double ru = rm_of(mode, s.lam[3], s.t[3], mode == 1 ? prod[3 * 16 + lane] : 0.0, sigmu);
        const double rsl = rm_of(mode, s.lsl, s.tsl, mode == 1 ? aw[12] : 0.0, sigmu);
        const double rsu = rm_of(mode, s.lsu, s.tsu, mode == 1 ? aw[13] : 0.0, sigmu);
        const double Gl = s.lam[2] / s.t[2], Gu = s.lam[3] / s.t[3], Gsl = s.lsl / s.tsl, Gsu = s.lsu / s.tsu;
        const double cl = (rl - s.lam[2] * s.r[2]) / s.t[2], cu = (ru - s.lam[3] * s.r[3]) / s.t[3];
        const double csl = (rsl - s.lsl * s.rsl) / s.tsl, csu = (rsu - s.lsu * s.rsu) / s.tsu;
        dsl = -(s.rgsl + cl + csl + Gl * adB) / (Gl + Gsl);
        dsu = -(s.rgsu + cu + csu - Gu * adB) / (Gu + Gsu);
        const double dtsl = dsl - s.rsl, dtsu = dsu - s.rsu;
        const double dlsl = -(rsl + s.lsl * dtsl) / s.tsl, dlsu = -(rsu + s.lsu * dtsu) / s.tsu;
        if (dlsl < 0.0) alpha = fmin(alpha, -s.lsl / dlsl);
        if (dlsu < 0.0) alpha = fmin(alpha, -s.lsu / dlsu);
        if (dtsl < 0.0) alpha = fmin(alpha, -s.tsl / dtsl);
        if (dtsu < 0.0) alpha = fmin(alpha, -s.tsu / dtsu);
        s_lin += s.lsl * dtsl + s.tsl * dlsl + s.lsu * dtsu + s.tsu * dlsu;
        s_quad += dlsl * dtsl + dlsu * dtsu;
        if (store_prod) { g[O_AUX + 12] = dlsl * dtsl; g[O_AUX + 13] = dlsu * dtsu; }
        if (final) { g[O_AUX + 6] = dsl; g[O_AUX + 7] = dsu; g[O_AUX + 8] = dlsl; g[O_AUX + 9] = dlsu; g[O_AUX + 10] = dtsl; g[O_AUX + 11] = dtsu; }
      }
      if (s.pa) { dtt[0] = adA - s.r[0]; dtt[1] = -adA - s.r[1]; }
      if (s.pb) { dtt[2] = adB + dsl - s.r[2]; dtt[3] = -adB + dsu - s.r[3]; }
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const bool p = i < 2 ? s.pa : s.pb;
        if (p) {
          const double rm = rm_of(mode, s.lam[i], s.t[i], mode == 1 ? prod[i * 16 + lane] : 0.0, sigmu);
          dlam[i] = -(rm + s.lam[i] * dtt[i]) / s.t[i];
          if (dlam[i] < 0.0) alpha = fmin(alpha, -s.lam[i] / dlam[i]);
          if (dtt[i] < 0.0) alpha = fmin(alpha, -s.t[i] / dtt[i]);
          s_lin += s.lam[i] * dtt[i] + s.t[i] * dlam[i];
          s_quad += dlam[i] * dtt[i];
        }
      }
      if (store_prod) {
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) g[O_PROD + i * 16 + lane] = dlam[i] * dtt[i];
      }
      if (final) {
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) { g[O_DLAM + i * 16 + lane] = dlam[i]; g[O_DTT + i * 16 + lane] = dtt[i]; }
        g[O_DZ + lane] = mydz;
      }
      // ---- next state ----
      if (k < N) {
        double dxn[10];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          dxn[i] = dx[i] + dt * dx[5 + i] + hdt2 * du[i] + T[O_RB + 5 + i];
          dxn[5 + i] = dx[5 + i] + dt * du[i] + T[O_RB + 10 + i];
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) dx[i] = dxn[i];
      }
      ln.sync();
      buf ^= 1;
    }
    StepStats o;
    o.alpha = gmin(alpha);
    o.s_lin = gsum(s_lin);
    o.s_quad = gsum(s_quad);
    ln.flush();
    return o;
  }

  // --------------------------------------------------------------------------------------- driver
  SMPC_HD QpResult solve() {
    QpResult R;
    R.iter = 0; R.status = 0;
    const double thr0 = 1e-1, lam_min = 1e-16, t_min = 1e-16;
    init(P.qp_mu0, thr0);
    double alpha = 1.0;
    double step = 0.0;      // step length still to be applied to the stored direction
    int kk = 0;
    bool nan = false;
    for (;; ++kk) {
      update_factorize(step, lam_min, t_min, P.qp_reg_prim, R);
      nan = (R.res[0] != R.res[0]) || (R.res[1] != R.res[1]) || (R.res[2] != R.res[2]) || (R.res[3] != R.res[3]);
      if (nan && kk > 0) break;
      const bool unconv = (R.res[0] > P.qp_tol_stat) || (R.res[1] > P.qp_tol_eq) || (R.res[2] > P.qp_tol_ineq) || (R.res[3] > P.qp_tol_comp);
      if (!unconv && !nan) break;
      if (kk >= P.qp_iter_max) break;
      if (!(alpha > P.qp_alpha_min)) break;
      // predictor
      const StepStats aff = forward(0, 0.0, true, false);
      const double mu_aff = R.mu + (aff.alpha * aff.s_lin + aff.alpha * aff.alpha * aff.s_quad) / nc;
      double sigma = mu_aff / R.mu; sigma = sigma * sigma * sigma;
      const double sigmu = sigma * R.mu;
      // corrector
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
