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
//     a capsule row, lane 11 the viability row (+ its slacks);  lam/t live in global memory as [slot][lane];
//   * the stage record (linearisation, 192 doubles) is staged through shared memory once per sweep and stage.
// The code is written against a small `Lanes` policy (lane id, shuffle, group barrier, scratch pointer) so that the
// same source runs on the device (LanesDev, qp.cu) and, for kernel-logic tests without a GPU, on the host
// (tests/emu: 16 threads and a barrier).  The product only ever instantiates the device policy.
#pragma once
#include "dev_model.cuh"

namespace smpc {

constexpr int QL = 16;            // lanes per problem
constexpr int NSLOT = 4;          // constraint slots per lane: rowA lower/upper, rowB lower/upper
constexpr int QP_SCRATCH = REC + 96;   // doubles of group scratch: stage record + replicated vectors

// per-problem global-memory views (all double; [..][16] arrays are indexed by lane)
struct QpMem {
  const double* rec;   // [N+1][REC]
  const double* x0;    // [10]
  double* z;           // [N+1][16]   primal iterate  (lane c: z_c)
  double* pi;          // [N+1][16]   multipliers of the dynamics k -> k+1 (lanes 5..14)
  double* lam;         // [N+1][4][16]
  double* t;           // [N+1][4][16]
  double* aux;         // [N+1][16]   slack data of the soft row: s_l s_u lam_sl lam_su t_sl t_su ds_l ds_u dlam_sl dlam_su dt_sl dt_su prod_sl prod_su
  double* fac;         // [N+1][5][16]  rows of the factor [Lr; Ls]: fac[j][lane] = L[lane][j]
  double* Pm;          // [N+1][10][16] Riccati matrix: Pm[r][lane] = P[r][lane-5]
  double* pv;          // [N+1][16]   l (lanes 0..4) and p (lanes 5..14) of the current solve
  double* wv;          // [N+1][16]   P_{k+1} res_b_k
  double* rb;          // [N+1][16]   res_b_k
  double* gb;          // [N+1][16]   res_g_k
  double* prod;        // [N+1][4][16] dlam_aff * dt_aff
  double* dz;          // [N+1][16]
  double* dpi;         // [N+1][16]
  double* dlam;        // [N+1][4][16]
  double* dtt;         // [N+1][4][16]
  double* L0;          // [2][10][16]  stage-0 state factor: columns then rows
  int r;               // receding index of the problem (RealReceding box override)
};

constexpr size_t qp_doubles_per_stage() { return 16 * (1 + 1 + 4 + 4 + 1 + 5 + 10 + 1 + 1 + 1 + 1 + 4 + 1 + 1 + 4 + 4); }
constexpr size_t qp_doubles_fixed() { return 2 * 10 * 16; }

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
  double* S;          // scratch: [0, REC) stage record, then replicated vectors
  double* V;          // S + REC: 96 doubles
  int nc;
  double dt, hdt2;

  SMPC_HD QpSolver(L& l, const smpc_problem_t& p, const QpMem& m) : ln(l), P(p), M(m), N(p.N), lane(l.lane()), S(l.scratch()), V(l.scratch() + REC), nc(0), dt(p.dt), hdt2(0.5 * p.dt * p.dt) {}

  // ---- small group helpers ----
  SMPC_HD double gsum(double v) { for (int o = 8; o > 0; o >>= 1) v += ln.shfl_xor(v, o); return v; }
  SMPC_HD double gmin(double v) { for (int o = 8; o > 0; o >>= 1) v = fmin(v, ln.shfl_xor(v, o)); return v; }
  SMPC_HD double gmax_nan(double v) {   // max that propagates NaN
    for (int o = 8; o > 0; o >>= 1) { double w = ln.shfl_xor(v, o); v = (v != v || w != w) ? (v + w) : fmax(v, w); }
    return v;
  }
  // replicate a lane-distributed 16-vector into scratch slot `off` (V[off + c] = value of lane c)
  SMPC_HD void publish(int off, double v) { ln.sync(); V[off + lane] = v; ln.sync(); }

  SMPC_HD void load_rec(int k) {
    ln.sync();
    const double* src = M.rec + (size_t)k * REC;
    for (int i = lane; i < REC; i += QL) S[i] = src[i];
    ln.sync();
  }

  // ---- row ownership ----
  SMPC_HD bool hasA(int k) const { return lane < 5 ? (S[SMPC_REC_NTAU] > 0.5) : (lane < 15); }
  SMPC_HD bool hasB(int k) const { return (lane >= 5 && lane <= 10) ? (S[SMPC_REC_NDIST] > 0.5) : (lane == 11 ? S[SMPC_REC_NNROW] > 0.5 : false); }
  SMPC_HD bool softB() const { return lane == 11 && S[SMPC_REC_NNROW] > 0.5 && S[SMPC_REC_SOFT] >= 0.0; }
  // canonical row id (box 0-9, tau 10-14, dist 15-20, nn 21)
  SMPC_HD int idA() const { return lane < 5 ? 10 + lane : lane - 5; }
  SMPC_HD int idB() const { return lane == 11 ? 21 : 15 + (lane - 5); }

  // a_rowA . y and a_rowB . y for a replicated 15-vector y (scratch offset `yo`, [u q v] order)
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
    double loA, hiA, loB, hiB;
  };

  // loads lam/t, evaluates res_d of the own slots for the replicated iterate zr
  SMPC_HD void load_slots(int k, const double* zr, Slots& s) {
    s.pa = hasA(k); s.pb = hasB(k); s.soft = softB();
    const double* lam = M.lam + (size_t)k * 64;
    const double* t = M.t + (size_t)k * 64;
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) { s.lam[i] = lam[i * 16 + lane]; s.t[i] = t[i * 16 + lane]; s.r[i] = 0.0; }
    s.aA = dotA(zr); s.aB = dotB(zr);
    s.loA = s.hiA = s.loB = s.hiB = 0.0;
    s.sl = s.su = 0.0;
    if (s.soft) {
      const double* a = M.aux + (size_t)k * 16;
      s.sl = a[0]; s.su = a[1]; s.lsl = a[2]; s.lsu = a[3]; s.tsl = a[4]; s.tsu = a[5];
      s.rsl = s.tsl - s.sl; s.rsu = s.tsu - s.su;
      s.rgsl = S[SMPC_REC_SOFT] - s.lam[2] - s.lsl;
      s.rgsu = S[SMPC_REC_SOFT] - s.lam[3] - s.lsu;
    }
    if (s.pa) { boundsA(k, s.loA, s.hiA); s.r[0] = s.t[0] - (s.aA - s.loA); s.r[1] = s.t[1] - (s.hiA - s.aA); }
    if (s.pb) { boundsB(s.loB, s.hiB); s.r[2] = s.t[2] - (s.aB + s.sl - s.loB); s.r[3] = s.t[3] - (s.hiB - s.aB + s.su); }
  }

  // complementarity right-hand side of slot i:  mode 0 affine, 1 corrector, 2 centering
  SMPC_HD double rm_of(int mode, double lam, double t, double prod, double sigmu) const {
    return mode == 0 ? lam * t : (mode == 1 ? lam * t + prod - sigmu : lam * t - sigmu);
  }

  // per-row condensation terms (Gamma, gamma) and nu = lam_hi - lam_lo of the own rows
  struct RowT { double GA, gA, nA, GB, gB, nB; };
  SMPC_HD RowT row_terms(int k, const Slots& s, int mode, double sigmu) {
    RowT o; o.GA = o.gA = o.nA = o.GB = o.gB = o.nB = 0.0;
    const double* prod = M.prod + (size_t)k * 64;
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
        const double* a = M.aux + (size_t)k * 16;
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
  SMPC_HD void init(QpResult& R, double mu0, double thr0) {
    double nb = 0.0, nd = 0.0, nm = 0.0, musum = 0.0;
    int cnt = 0;
    double zxn = 0.0;   // z of stage k+1, own lane
    for (int k = N; k >= 0; --k) {
      load_rec(k);
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
      const bool pa = hasA(k), pb = hasB(k), soft = softB();
      double lam[NSLOT] = {0, 0, 0, 0}, t[NSLOT] = {0, 0, 0, 0};
      if (lane < 5 && pa) {
        boundsA(k, loA, hiA);
        const double v = dotA(zr);
        tl = fmax(thr0, v - loA); tu = fmax(thr0, hiA - v);
      }
      if (pa) { t[0] = tl; t[1] = tu; }
      double rdB0 = 0.0, rdB1 = 0.0;
      if (pb) {
        double lo, hi; boundsB(lo, hi);
        const double v = dotB(zr);
        t[2] = fmax(thr0, v - lo); t[3] = fmax(thr0, hi - v);
        const double sl = soft ? thr0 : 0.0;
        rdB0 = t[2] - (v + sl - lo); rdB1 = t[3] - (hi - v + sl);
      }
      double rdA0 = 0.0, rdA1 = 0.0;
      if (pa) { const double v = dotA(zr); rdA0 = t[0] - (v - loA); rdA1 = t[1] - (hiA - v); }
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const bool p = i < 2 ? pa : pb;
        lam[i] = p ? mu0 / t[i] : 0.0;
        M.lam[(size_t)k * 64 + i * 16 + lane] = lam[i];
        M.t[(size_t)k * 64 + i * 16 + lane] = t[i];
        if (p) { ++cnt; musum += lam[i] * t[i]; nm = fmax(nm, fabs(lam[i] * t[i])); }
      }
      nd = fmax(nd, fmax(fmax(fabs(rdA0), fabs(rdA1)), fmax(fabs(rdB0), fabs(rdB1))));
      if (lane == 11) {
        double* a = M.aux + (size_t)k * 16;
        for (int i = 0; i < 16; ++i) a[i] = 0.0;
        if (soft) {
          a[0] = a[1] = thr0; a[4] = a[5] = thr0; a[2] = a[3] = mu0 / thr0;
          cnt += 2; musum += 2.0 * mu0; nm = fmax(nm, mu0);
        }
      }
      M.z[(size_t)k * 16 + lane] = zc;
      M.pi[(size_t)k * 16 + lane] = 0.0;
      // res_b_k = A zx_k + B zu_k + b_k - zx_{k+1}   (zu = 0)
      if (k < N && lane >= 5 && lane < 15) {
        const int i = lane - 5;
        const double y = i < 5 ? zr[5 + i] + dt * zr[10 + i] : zr[5 + i];
        nb = fmax(nb, fabs(y + S[SMPC_REC_B + i] - zxn));
      }
      zxn = zc;
    }
    nc = (int)(gsum((double)cnt) + 0.5);
    R.res[1] = gmax_nan(nb); R.res[2] = gmax_nan(nd); R.res[3] = gmax_nan(nm);
    R.mu = gsum(musum) / nc;
    R.res[0] = 0.0;
  }

  // -------------------------------------------------------------- S1: factorisation + affine gradient
  // returns the max-norm of the stationarity residual; leaves dx_0 (affine) replicated at V[64..74)
  SMPC_HD double factorize(double reg) {
    double ng = 0.0;
    double pp[10];            // column (lane-5) of P_{k+1}   (lanes 5..14)
    double pcur = 0.0;        // p_{k+1}[lane-5]
    double zxn[10];           // z_x of stage k+1, replicated
#pragma unroll
    for (int i = 0; i < 10; ++i) { pp[i] = 0.0; zxn[i] = 0.0; }
    for (int k = N; k >= 0; --k) {
      load_rec(k);
      publish(0, M.z[(size_t)k * 16 + lane]);                 // V[0..15) = z_k
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      Slots s;
      load_slots(k, zr, s);
      const RowT rt = row_terms(k, s, 0, 0.0);
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
        for (int j = 0; j < 5; ++j) { const int a = i > j ? i : j, b = i > j ? j : i; m[5 + j] = S[SMPC_REC_HQQ + a * (a + 1) / 2 + b]; }
        m[lane] += S[SMPC_REC_HQ] + reg + GamBox;
      } else if (lane < 15) { m[lane] = S[SMPC_REC_HV] + reg + GamBox; }
      if (lane < 15) {
        if (S[SMPC_REC_NTAU] > 0.5) {
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const double* a = S + SMPC_REC_JTAU + r * 15;
            const double coef = V[16 + 10 + r] * a[lane];
#pragma unroll
            for (int i = 0; i < 15; ++i) m[i] += coef * a[i];
          }
        }
        if (lane >= 5 && lane < 10 && S[SMPC_REC_NDIST] > 0.5) {
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const double* a = S + SMPC_REC_JDIST + p * 5;
            const double coef = V[16 + 15 + p] * a[lane - 5];
#pragma unroll
            for (int i = 0; i < 5; ++i) m[5 + i] += coef * a[i];
          }
        }
        if (lane >= 5 && S[SMPC_REC_NNROW] > 0.5) {
          const double* a = S + SMPC_REC_JNN;
          const double coef = V[16 + 21] * a[lane - 5];
#pragma unroll
          for (int i = 0; i < 10; ++i) m[5 + i] += coef * a[i];
        }
      }
      // ---- stationarity residual and affine gradient ----
      publish_rows(40, s.pa, rt.nA, s.pb, rt.nB);              // V[40..62) nu = lam_hi - lam_lo
      double rg = cost_grad(k, zr) + rowsT(40);
      publish_rows(40, s.pa, rt.gA, s.pb, rt.gB);              // V[40..62) gamma
      double gv = rowsT(40);
      // dynamics terms
      if (k < N) {
        publish(64, M.pi[(size_t)k * 16 + lane]);              // V[64+5 .. 64+15) = pi_k
        rg += dynT(V + 64 + 5);
      }
      if (k > 0) { if (lane >= 5 && lane < 15) rg -= M.pi[(size_t)(k - 1) * 16 + lane]; }
      if (k == N && lane < 5) rg = 0.0;
      ng = (rg != rg) ? rg : fmax(ng, fabs(rg));
      M.gb[(size_t)k * 16 + lane] = rg;
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
          M.rb[(size_t)k * 16 + lane] = rb[lane - 5];
          M.wv[(size_t)k * 16 + lane] = w;
        }
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
      for (int j = 0; j < 5; ++j) M.fac[((size_t)k * 5 + j) * 16 + lane] = lrow[j];
      M.pv[(size_t)k * 16 + lane] = m[15];
      if (k > 0) {
        if (lane >= 5 && lane < 15) {
#pragma unroll
          for (int r = 0; r < 10; ++r) { pp[r] = m[5 + r]; M.Pm[((size_t)k * 10 + r) * 16 + lane] = pp[r]; }
          pcur = m[15];
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) zxn[i] = zr[5 + i];
      } else {
        // stage 0: keep the state factor (columns and rows) and solve for dx_0
#pragma unroll
        for (int r = 0; r < 10; ++r) { M.L0[(size_t)r * 16 + lane] = m[5 + r]; M.L0[(size_t)(10 + r) * 16 + lane] = l0row[r]; }
        solve_dx0(m, m[15]);
      }
    }
    return gmax_nan(ng);
  }

  // back substitution with the stage-0 state factor held column-wise in mcol[5..15): dx_0 -> V[64+5 .. 64+15), replicated
  SMPC_HD void solve_dx0(const double* mcol, double lx) {
    double acc = -lx;
    double mine = 0.0;
    for (int r = 14; r >= 5; --r) {
      double dr = 0.0, lrr = 0.0, lrc = 0.0;
#pragma unroll
      for (int i = 5; i < 15; ++i) if (i == r) lrc = mcol[i];     // L0[r][lane] (valid for r >= lane)
      lrr = lrc;
      if (lane == r) dr = lrr > 0.0 ? acc / lrr : 0.0;
      dr = ln.shfl(dr, r);
      if (lane == r) mine = dr;
      if (lane >= 5 && lane < r) acc -= lrc * dr;
    }
    publish(64, mine);
  }

  // ------------------------------------------------- S3: vector-only backward recursion (corrector / centering)
  SMPC_HD void resolve_backward(int mode, double sigmu) {
    double pcur = 0.0;
    for (int k = N; k >= 0; --k) {
      load_rec(k);
      publish(0, M.z[(size_t)k * 16 + lane]);
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      Slots s;
      load_slots(k, zr, s);
      const RowT rt = row_terms(k, s, mode, sigmu);
      publish_rows(40, s.pa, rt.gA, s.pb, rt.gB);
      double gv = M.gb[(size_t)k * 16 + lane] + rowsT(40);
      if (k < N) {
        const double w = (lane >= 5 && lane < 15) ? M.wv[(size_t)k * 16 + lane] : 0.0;
        publish(80, w + pcur);
        gv += dynT(V + 80 + 5);
      }
      if (k == N && lane < 5) gv = 0.0;
      if (lane == 15) gv = 0.0;
      // forward elimination with the stored rows of [Lr; Ls]
      double lrow[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) lrow[j] = M.fac[((size_t)k * 5 + j) * 16 + lane];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double lj = 0.0;
        if (lane == j) { lj = lrow[j] > 0.0 ? gv / lrow[j] : 0.0; gv = lj; }
        lj = ln.shfl(lj, j);
        if (lane > j && lane < 15) gv -= lrow[j] * lj;
      }
      M.pv[(size_t)k * 16 + lane] = gv;
      if (k > 0) { pcur = (lane >= 5 && lane < 15) ? gv : 0.0; }
      else {
        // stage 0: forward then backward substitution with the state factor
        double l0row[10], mcol[16];
#pragma unroll
        for (int r = 0; r < 10; ++r) { mcol[5 + r] = M.L0[(size_t)r * 16 + lane]; l0row[r] = M.L0[(size_t)(10 + r) * 16 + lane]; }
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
    }
  }

  // ------------------------------------------------------------------------- S2 / S4: forward substitution
  struct StepStats { double alpha, s_lin, s_quad; };
  // mode: complementarity rhs as in rm_of; store_prod: keep dlam*dt (affine pass); final: store dlam, dt, dpi
  SMPC_HD StepStats forward(int mode, double sigmu, bool store_prod, bool final) {
    double alpha = 1.0, s_lin = 0.0, s_quad = 0.0;
    double dx[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) dx[i] = V[64 + 5 + i];
    ln.sync();
    for (int k = 0; k <= N; ++k) {
      load_rec(k);
      // ---- du_k ----
      double du[5] = {0, 0, 0, 0, 0};
      double lrow[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) lrow[j] = M.fac[((size_t)k * 5 + j) * 16 + lane];
      if (k < N) {
        const double lp = M.pv[(size_t)k * 16 + lane];
        double sj[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          double term = 0.0;
          if (lane >= 5 && lane < 15) term = lrow[j] * dx[lane - 5];
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
      publish(0, M.z[(size_t)k * 16 + lane]);
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      Slots s;
      load_slots(k, zr, s);
      const double adA = dotA(dzr), adB = dotB(dzr);
      const double* prod = M.prod + (size_t)k * 64;
      double dlam[NSLOT] = {0, 0, 0, 0}, dtt[NSLOT] = {0, 0, 0, 0};
      double dsl = 0.0, dsu = 0.0;
      if (s.soft) {
        const double* a = M.aux + (size_t)k * 16;
        const double rl = rm_of(mode, s.lam[2], s.t[2], mode == 1 ? prod[2 * 16 + lane] : 0.0, sigmu);
        const double ru = rm_of(mode, s.lam[3], s.t[3], mode == 1 ? prod[3 * 16 + lane] : 0.0, sigmu);
        const double rsl = rm_of(mode, s.lsl, s.tsl, mode == 1 ? a[12] : 0.0, sigmu);
        const double rsu = rm_of(mode, s.lsu, s.tsu, mode == 1 ? a[13] : 0.0, sigmu);
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
        double* aw = M.aux + (size_t)k * 16;
        if (store_prod) { aw[12] = dlsl * dtsl; aw[13] = dlsu * dtsu; }
        if (final) { aw[6] = dsl; aw[7] = dsu; aw[8] = dlsl; aw[9] = dlsu; aw[10] = dtsl; aw[11] = dtsu; }
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
        for (int i = 0; i < NSLOT; ++i) M.prod[(size_t)k * 64 + i * 16 + lane] = dlam[i] * dtt[i];
      }
      if (final) {
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) { M.dlam[(size_t)k * 64 + i * 16 + lane] = dlam[i]; M.dtt[(size_t)k * 64 + i * 16 + lane] = dtt[i]; }
        M.dz[(size_t)k * 16 + lane] = mydz;
      }
      // ---- next state ----
      if (k < N) {
        double dxn[10];
        ln.sync();
        V[80 + lane] = M.rb[(size_t)k * 16 + lane];
        ln.sync();
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          dxn[i] = dx[i] + dt * dx[5 + i] + hdt2 * du[i] + V[80 + 5 + i];
          dxn[5 + i] = dx[5 + i] + dt * du[i] + V[80 + 10 + i];
        }
        if (final) {
          double dp = 0.0;
          if (lane >= 5 && lane < 15) {
            dp = M.pv[(size_t)(k + 1) * 16 + lane];
#pragma unroll
            for (int r = 0; r < 10; ++r) dp += M.Pm[((size_t)(k + 1) * 10 + r) * 16 + lane] * dxn[r];
          }
          M.dpi[(size_t)k * 16 + lane] = dp;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) dx[i] = dxn[i];
      }
    }
    StepStats o;
    o.alpha = gmin(alpha);
    o.s_lin = gsum(s_lin);
    o.s_quad = gsum(s_quad);
    return o;
  }

  // --------------------------------------------------------------------------------------- S5: update
  SMPC_HD void update(double a, double lam_min, double t_min, QpResult& R) {
    double nb = 0.0, nd = 0.0, nm = 0.0, musum = 0.0;
    double zprev[15];   // updated z of stage k-1 (replicated)
#pragma unroll
    for (int i = 0; i < 15; ++i) zprev[i] = 0.0;
    double bprev = 0.0;  // b_{k-1}[lane-5]
    for (int k = 0; k <= N; ++k) {
      load_rec(k);
      const double znew = M.z[(size_t)k * 16 + lane] + a * M.dz[(size_t)k * 16 + lane];
      M.z[(size_t)k * 16 + lane] = znew;
      if (k < N) M.pi[(size_t)k * 16 + lane] += a * M.dpi[(size_t)k * 16 + lane];
      publish(0, znew);
      double zr[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) zr[i] = V[i];
      // res_b_{k-1}
      if (k > 0 && lane >= 5 && lane < 15) {
        const int i = lane - 5;
        const double y = i < 5 ? zprev[5 + i] + dt * zprev[10 + i] + hdt2 * zprev[i] : zprev[5 + i] + dt * zprev[i - 5];
        const double rbv = y + bprev - znew;
        nb = (rbv != rbv) ? rbv : fmax(nb, fabs(rbv));
      }
      bprev = (lane >= 5 && lane < 15) ? S[SMPC_REC_B + lane - 5] : 0.0;
#pragma unroll
      for (int i = 0; i < 15; ++i) zprev[i] = zr[i];
      // slots
      const bool pa = hasA(k), pb = hasB(k), soft = softB();
      double sl = 0.0, su = 0.0;
      if (soft) {
        double* aw = M.aux + (size_t)k * 16;
        sl = aw[0] + a * aw[6]; su = aw[1] + a * aw[7];
        const double lsl = fmax(aw[2] + a * aw[8], lam_min), lsu = fmax(aw[3] + a * aw[9], lam_min);
        const double tsl = fmax(aw[4] + a * aw[10], t_min), tsu = fmax(aw[5] + a * aw[11], t_min);
        aw[0] = sl; aw[1] = su; aw[2] = lsl; aw[3] = lsu; aw[4] = tsl; aw[5] = tsu;
        musum += lsl * tsl + lsu * tsu;
        nm = fmax(nm, fmax(fabs(lsl * tsl), fabs(lsu * tsu)));
        nd = fmax(nd, fmax(fabs(tsl - sl), fabs(tsu - su)));
      }
      double lam[NSLOT], t[NSLOT];
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const bool p = i < 2 ? pa : pb;
        const size_t o = (size_t)k * 64 + i * 16 + lane;
        lam[i] = 0.0; t[i] = 0.0;
        if (p) {
          lam[i] = fmax(M.lam[o] + a * M.dlam[o], lam_min);
          t[i] = fmax(M.t[o] + a * M.dtt[o], t_min);
          M.lam[o] = lam[i]; M.t[o] = t[i];
          const double c = lam[i] * t[i];
          musum += c;
          nm = (c != c) ? c : fmax(nm, fabs(c));
        }
      }
      double rd = 0.0;
      if (pa) { double lo, hi; boundsA(k, lo, hi); const double v = dotA(zr); rd = fmax(fabs(t[0] - (v - lo)), fabs(t[1] - (hi - v))); }
      if (pb) { double lo, hi; boundsB(lo, hi); const double v = dotB(zr); rd = fmax(rd, fmax(fabs(t[2] - (v + sl - lo)), fabs(t[3] - (hi - v + su)))); }
      nd = (rd != rd) ? rd : fmax(nd, rd);
    }
    R.res[1] = gmax_nan(nb); R.res[2] = gmax_nan(nd); R.res[3] = gmax_nan(nm);
    R.mu = gsum(musum) / nc;
  }

  // --------------------------------------------------------------------------------------- driver
  SMPC_HD QpResult solve() {
    QpResult R;
    R.iter = 0; R.status = 0;
    const double thr0 = 1e-1, lam_min = 1e-16, t_min = 1e-16;
    init(R, P.qp_mu0, thr0);
    double alpha = 1.0;
    int kk = 0;
    bool nan = false;
    for (;; ++kk) {
      R.res[0] = factorize(P.qp_reg_prim);
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
      update(0.995 * alpha, lam_min, t_min, R);
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
