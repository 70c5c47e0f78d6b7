// ORACLE -- test infrastructure only (see oracle.h).  CPU restatement of the reference hot path:
//   linearisation of one RTI iteration (reference controller.py:136-167 + the CasADi expression graphs of
//   env_model.py:58-101,130-165,246-319, cost_definition.py:61-100, safe_set.py:71-104, utils.py:94-113,207-210),
//   the stage QP + interior-point solve (qp.hpp), the controller state machines (controller.py:169-184,226-231,
//   251-689), the plant step (env_model.py:192-206) and the closed loop (scripts/mpc.py:102-291).
// Derivatives come from forward-mode AD of a plain restatement of the expressions (dual.hpp), as they come from
// CasADi's AD upstream -- deliberately NOT from the hand-derived recursions the CUDA kernels use.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#include "model.hpp"
#include "qp.hpp"

using namespace orc;

#ifdef ORC_COUNT_FLOPS
namespace orc { thread_local unsigned long long g_flops = 0; }
extern "C" void orc_flops_reset() { orc::g_flops = 0; }
extern "C" unsigned long long orc_flops_get() { return orc::g_flops; }
#endif

namespace {
constexpr int NQ = ORC_NQ, NX = ORC_NX, NU = ORC_NU, REC = SMPC_REC;

struct Workspace {
  std::unique_ptr<QpIpm> qp;
};

// dynamic-schedule parallel loop over [0, n) on `threads` std::threads; fn(index, thread_id)
template <class Fn>
void parallel_for(int n, int threads, Fn fn) {
  if (threads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) fn(i, 0); return; }
  std::atomic<int> next(0);
  auto body = [&](int tid) { for (;;) { int i = next.fetch_add(1); if (i >= n) break; fn(i, tid); } };
  std::vector<std::thread> pool;
  const int nt = std::min(threads, n);
  for (int t = 1; t < nt; ++t) pool.emplace_back(body, t);
  body(0);
  for (auto& th : pool) th.join();
}
}  // namespace

struct orc_handle {
  orc_problem_t P;
  std::vector<float> w;
  int B = 0, N = 0, threads = 1;
  int mlp_fp32 = 0;
  // tests only: the QP of every solve handed to an external solver (tests/emu: the engine's kernel sources on the host), see oracle.h
  orc_qp_hook_t qp_hook = nullptr;
  // tests only: every solve repeated on perturbed data; per problem, the number of solves whose status depends on the perturbation
  double probe_eps = 0.0;
  std::vector<int32_t> probe_flips;
  std::vector<double> xg, ug, xt, ut, lin, plant_inertial, tau_noise, x_viable;
  std::vector<double> ee_traj;   // [n][3] end-effector reference per control step (cost.traj of the reference); empty: P.ee_ref
  std::vector<double> qp_z, qp_pi, qp_lam, qp_t, qp_res;
  std::vector<int32_t> fails, r, status, qp_iter, qp_status, cur_step, cand;
  std::vector<Workspace> ws;   // one per thread
};

struct orc_sim {
  orc_handle* c;
  orc_handle* bk;
  int n_steps = 0, j = 0;
  std::vector<double> x, xlog, ulog, x_abort, u_abort, xv_first;
  std::vector<int32_t> mode, ja, outcome;
  int64_t counters[4] = {0, 0, 0, 0};
  // scripted solve outcomes of the NEXT step (tests only, orc_sim_set_script): main controller and backup OCP
  const int32_t *scr_status = nullptr, *scr_bk_status = nullptr;
  const double *scr_xt = nullptr, *scr_ut = nullptr, *scr_bk_xt = nullptr, *scr_bk_ut = nullptr;
};

namespace {

double nn_c(const orc_handle& h, const double* x, double* grad) {
  return h.mlp_fp32 ? nn_constraint<float>(h.P, x, h.P.alpha, grad) : nn_constraint<double>(h.P, x, h.P.alpha, grad);
}

void f_disc(double dt, const double* x, const double* u, double* xn) {
  // reference env_model.py:63-71
  for (int i = 0; i < NQ; ++i) {
    xn[i] = x[i] + dt * x[NQ + i] + 0.5 * dt * dt * u[i];
    xn[NQ + i] = x[NQ + i] + dt * u[i];
  }
}

void distances(const orc_problem_t& P, const double* q, double* ee, double* dist) {
  double pts[ORC_MAX_POINTS][3];
  fk_points<double>(P, q, pts);
  if (ee) for (int k = 0; k < 3; ++k) ee[k] = pts[0][k];
  if (dist)
    for (int p = 0; p < ORC_NPAIR; ++p) dist[p] = segment_dist<double>(pts[P.pair_pa[p]], pts[P.pair_pb[p]], P.pair_C[p], P.pair_D[p]);
}

// checkStateBounds (env_model.py:175-177) on one state
bool state_in_bounds(const orc_problem_t& P, const double* x) {
  for (int i = 0; i < NX; ++i)
    if (!(x[i] >= P.x_min[i] - P.tol_x && x[i] <= P.x_max[i] + P.tol_x)) return false;
  return true;
}
// checkCollision (env_model.py:236-243) on one state
bool collision_free(const orc_problem_t& P, const double* x) {
  double d[ORC_NPAIR];
  distances(P, x, nullptr, d);
  for (int p = 0; p < ORC_NPAIR; ++p)
    if (!(P.pair_lo_chk[p] <= d[p] && d[p] <= P.pair_hi + P.tol_obs)) return false;
  return true;
}

bool stage_has_nn(const orc_problem_t& P, int k) {
  switch (P.nn_rows) {
    case SMPC_NN_TERMINAL: return k == P.N;
    case SMPC_NN_RECEDING: case SMPC_NN_EVERYWHERE: case SMPC_NN_PARALLEL: return k >= 1;
    default: return false;
  }
}

// One stage of the linearisation -> stage record (layout in include/safe_mpc_b200.h)
// ee_ref: p[0:3] of this stage = cost.traj[:, current_step + k] (controller.py:153-156)
void linearize_stage(const orc_handle& h, int k, const double* x, const double* u, const double* xnext, bool gate_on, const double* ee_ref, double* rec) {
  const orc_problem_t& P = h.P;
  const int N = P.N;
  const bool term = (k == N);
  const double s = term ? 1.0 : P.dt;   // acados scales stage costs by the time step, the terminal one by 1
  for (int i = 0; i < REC; ++i) rec[i] = 0.0;
  for (int i = 0; i < NX; ++i) rec[SMPC_REC_X + i] = x[i];
  if (!term) for (int i = 0; i < NU; ++i) rec[SMPC_REC_U + i] = u[i];
  const double* q = x;
  const double* v = x + NQ;
  // ---- cost (cost_definition.py:61-100) ----
  double hu = 0.0;
  if (P.cost_type == SMPC_COST_EXT) {
    Dual2<NQ> qd[NQ], pts[ORC_MAX_POINTS][3];
    for (int i = 0; i < NQ; ++i) qd[i] = Dual2<NQ>::var(q[i], i);
    fk_points<Dual2<NQ>>(P, qd, pts);
    Dual2<NQ> c(0.0);
    for (int d = 0; d < 3; ++d) { Dual2<NQ> e = pts[0][d] - ee_ref[d]; c = c + e * e; }
    c = c * P.q_weight;
    for (int i = 0; i < NQ; ++i) rec[SMPC_REC_G + NU + i] = s * c.g[i];
    int o = 0;
    for (int i = 0; i < NQ; ++i) for (int j = 0; j <= i; ++j) rec[SMPC_REC_HQQ + o++] = s * c.h[Dual2<NQ>::idx(i, j)];
    if (!term) { for (int i = 0; i < NU; ++i) rec[SMPC_REC_G + i] = s * 2.0 * P.r_weight * u[i]; hu = s * 2.0 * P.r_weight; }
  } else if (P.cost_type == SMPC_COST_NLS) {
    Dual<NQ> qd[NQ], pts[ORC_MAX_POINTS][3];
    for (int i = 0; i < NQ; ++i) qd[i] = Dual<NQ>::var(q[i], i);
    fk_points<Dual<NQ>>(P, qd, pts);
    for (int i = 0; i < NQ; ++i) {
      double g = 0.0;
      for (int d = 0; d < 3; ++d) g += pts[0][d].d[i] * (pts[0][d].v - ee_ref[d]);
      rec[SMPC_REC_G + NU + i] = s * P.q_weight * g;
    }
    int o = 0;
    for (int i = 0; i < NQ; ++i) for (int j = 0; j <= i; ++j) {
      double a = 0.0;
      for (int d = 0; d < 3; ++d) a += pts[0][d].d[i] * pts[0][d].d[j];
      rec[SMPC_REC_HQQ + o++] = s * P.q_weight * a;
    }
    if (!term) { for (int i = 0; i < NU; ++i) rec[SMPC_REC_G + i] = s * P.r_weight * u[i]; hu = s * P.r_weight; }
  }
  const double lmk = P.lm * ((P.lm_scale_dt && !term) ? P.dt : 1.0);
  rec[SMPC_REC_HU] = term ? 0.0 : hu + lmk;
  rec[SMPC_REC_HV] = lmk;
  rec[SMPC_REC_HQ] = lmk;
  // ---- torque rows (env_model.py:80-83,258-259) ----
  if (!term) {
    Dual<15> qd[NQ], vd[NQ], ud[NQ], tau[NQ];
    for (int i = 0; i < NQ; ++i) { ud[i] = Dual<15>::var(u[i], i); qd[i] = Dual<15>::var(q[i], NU + i); vd[i] = Dual<15>::var(v[i], NU + NQ + i); }
    rnea<Dual<15>>(P, P.inertial, qd, vd, ud, tau);
    for (int i = 0; i < NU; ++i) {
      rec[SMPC_REC_TAU + i] = tau[i].v;
      for (int j = 0; j < 15; ++j) rec[SMPC_REC_JTAU + i * 15 + j] = tau[i].d[j];
    }
    rec[SMPC_REC_NTAU] = NU;
  }
  // ---- capsule rows (env_model.py:263-271, utils.py:94-113) ----
  if (k > 0 || P.stage0_collision_rows) {
    Dual<NQ> qd[NQ], pts[ORC_MAX_POINTS][3];
    for (int i = 0; i < NQ; ++i) qd[i] = Dual<NQ>::var(q[i], i);
    fk_points<Dual<NQ>>(P, qd, pts);
    for (int p = 0; p < ORC_NPAIR; ++p) {
      Dual<NQ> d = segment_dist<Dual<NQ>>(pts[P.pair_pa[p]], pts[P.pair_pb[p]], P.pair_C[p], P.pair_D[p]);
      rec[SMPC_REC_DIST + p] = d.v;
      for (int j = 0; j < NQ; ++j) rec[SMPC_REC_JDIST + p * NQ + j] = d.d[j];
    }
    rec[SMPC_REC_NDIST] = ORC_NPAIR;
  }
  // ---- viability row (safe_set.py:82-104, utils.py:207-210) ----
  rec[SMPC_REC_SOFT] = -1.0;
  if (stage_has_nn(P, k)) {
    rec[SMPC_REC_NNROW] = 1.0;
    if (gate_on) rec[SMPC_REC_NN] = nn_c(h, x, &rec[SMPC_REC_JNN]);
    else rec[SMPC_REC_NN] = 5e5;   // (0 + 1e6)/2, zero Jacobian
    if (term && P.nn_terminal_soft) rec[SMPC_REC_SOFT] = P.slack_penalty_e;
  }
  // ---- dynamics offset ----
  if (!term) {
    double xn[NX];
    f_disc(P.dt, x, u, xn);
    for (int i = 0; i < NX; ++i) rec[SMPC_REC_B + i] = xn[i] - xnext[i];
  }
}

// stage record (+ box) -> dense QP stage
void assemble_stage(const orc_problem_t& P, int k, const double* rec, const double* blo, const double* bhi, QpStage& S) {
  const bool term = (k == P.N);
  std::memset(&S, 0, sizeof(S));
  S.nu = term ? 0 : NU;
  const int nu = S.nu;
  if (!term) for (int i = 0; i < NU; ++i) { S.H[i][i] = rec[SMPC_REC_HU]; S.g[i] = rec[SMPC_REC_G + i]; }
  int o = 0;
  for (int i = 0; i < NQ; ++i) for (int j = 0; j <= i; ++j) { S.H[nu + i][nu + j] = S.H[nu + j][nu + i] = rec[SMPC_REC_HQQ + o++]; }
  for (int i = 0; i < NQ; ++i) {
    S.H[nu + i][nu + i] += rec[SMPC_REC_HQ];
    S.H[nu + NQ + i][nu + NQ + i] = rec[SMPC_REC_HV];
    S.g[nu + i] = rec[SMPC_REC_G + NU + i];
    S.g[nu + NQ + i] = rec[SMPC_REC_G + NU + NQ + i];
  }
  for (int i = 0; i < NX; ++i) { S.b[i] = rec[SMPC_REC_B + i]; S.blo[i] = blo[i]; S.bhi[i] = bhi[i]; }
  // general rows in canonical order; absent families are simply not emitted, but keep canonical slots by
  // remembering the slot of each emitted row
  S.ng = 0;
  S.soft_row = -1;
  const int ntau = (int)rec[SMPC_REC_NTAU], ndist = (int)rec[SMPC_REC_NDIST];
  for (int i = 0; i < ntau; ++i) {
    double* a = S.C[S.ng];
    for (int j = 0; j < 15; ++j) a[j] = rec[SMPC_REC_JTAU + i * 15 + j];
    S.glo[S.ng] = P.tau_min[i] - rec[SMPC_REC_TAU + i];
    S.ghi[S.ng] = P.tau_max[i] - rec[SMPC_REC_TAU + i];
    ++S.ng;
  }
  for (int p = 0; p < ndist; ++p) {
    double* a = S.C[S.ng];
    for (int j = 0; j < NQ; ++j) a[nu + j] = rec[SMPC_REC_JDIST + p * NQ + j];
    S.glo[S.ng] = P.pair_lo_ocp[p] - rec[SMPC_REC_DIST + p];
    S.ghi[S.ng] = P.pair_hi - rec[SMPC_REC_DIST + p];
    ++S.ng;
  }
  if (rec[SMPC_REC_NNROW] > 0.5) {
    double* a = S.C[S.ng];
    for (int j = 0; j < NX; ++j) a[nu + j] = rec[SMPC_REC_JNN + j];
    S.glo[S.ng] = 0.0 - rec[SMPC_REC_NN];
    S.ghi[S.ng] = 1e6 - rec[SMPC_REC_NN];
    if (rec[SMPC_REC_SOFT] >= 0.0) { S.soft_row = S.ng; S.zl = S.zu = rec[SMPC_REC_SOFT]; }
    ++S.ng;
  }
}

// canonical slot of general row `g` (index among emitted rows) -> row index in [box 10 | tau 5 | dist 6 | nn 1]
int canonical_row(const double* rec, int g) {
  const int ntau = (int)rec[SMPC_REC_NTAU], ndist = (int)rec[SMPC_REC_NDIST];
  if (g < ntau) return 10 + g;
  g -= ntau;
  if (g < ndist) return 15 + g;
  return 21;
}

void stage_box(const orc_handle& h, int b, int k, const double* x0, double* lo, double* hi) {
  const orc_problem_t& P = h.P;
  const int N = P.N;
  const double* xg = &h.xg[(size_t)b * (N + 1) * NX];
  const double* xk = xg + k * NX;
  if (k == 0) {
    for (int i = 0; i < NX; ++i) lo[i] = hi[i] = x0[i] - xk[i];     // lbx_0 = ubx_0 = x0 (controller.py:144-145)
    return;
  }
  if (k == N) { for (int i = 0; i < NX; ++i) { lo[i] = P.lbx_e[i] - xk[i]; hi[i] = P.ubx_e[i] - xk[i]; } return; }
  if (P.controller == SMPC_CTRL_REAL_RECEDING) {                     // controller.py:530-536
    const int r = h.r[b];
    if (k == r) { for (int i = 0; i < NX; ++i) { lo[i] = xg[(r + 1) * NX + i] - 1e-3 - xk[i]; hi[i] = xg[(r + 1) * NX + i] + 1e-3 - xk[i]; } }
    else { for (int i = 0; i < NX; ++i) { lo[i] = P.x_min[i] - xk[i]; hi[i] = P.x_max[i] - xk[i]; } }
    return;
  }
  for (int i = 0; i < NX; ++i) { lo[i] = P.lbx[i] - xk[i]; hi[i] = P.ubx[i] - xk[i]; }
}

bool gate_on(const orc_handle& h, int b, int k) {
  const orc_problem_t& P = h.P;
  if (P.nn_rows == SMPC_NN_RECEDING && k < P.N) return k == h.r[b];   // controller.py:452-469
  if (P.nn_rows == SMPC_NN_PARALLEL) return k == h.cand[b];           // controller.py:578-588 (constrain_n)
  return true;
}

// AbstractController.solve(x0) for problem b (controller.py:136-167)
int rti_solve_one(orc_handle& h, int b, const double* x0, Workspace& W) {
  const orc_problem_t& P = h.P;
  const int N = P.N;
  double* xg = &h.xg[(size_t)b * (N + 1) * NX];
  double* ug = &h.ug[(size_t)b * N * NU];
  double* lin = &h.lin[(size_t)b * (N + 1) * REC];
  if (!W.qp) W.qp.reset(new QpIpm(N, P.dt));
  QpIpm& qp = *W.qp;
  for (int k = 0; k <= N; ++k) {
    double* rec = lin + (size_t)k * REC;
    const double* eer = P.ee_ref;
    if (!h.ee_traj.empty()) {
      const int n = (int)(h.ee_traj.size() / 3), c = h.cur_step[b] + k;
      eer = &h.ee_traj[3 * (size_t)(c < n - 1 ? c : n - 1)];
    }
    linearize_stage(h, k, xg + k * NX, k < N ? ug + k * NU : nullptr, k < N ? xg + (k + 1) * NX : nullptr, gate_on(h, b, k), eer, rec);
    double lo[NX], hi[NX];
    stage_box(h, b, k, x0, lo, hi);
    assemble_stage(P, k, rec, lo, hi, qp.stages()[k]);
  }
  QpOpts o;
  o.iter_max = P.qp_iter_max; o.mu0 = P.qp_mu0; o.tol_stat = P.qp_tol_stat; o.tol_eq = P.qp_tol_eq;
  o.tol_ineq = P.qp_tol_ineq; o.tol_comp = P.qp_tol_comp; o.alpha_min = P.qp_alpha_min; o.reg_prim = P.qp_reg_prim;
  o.cond_pred_corr = P.qp_cond_pred_corr;
  if (h.qp_hook) {
    // the QP of this solve goes to the external solver; it returns what the engine's qs_ctl / qs_final would (status mapping included)
    double* xt = &h.xt[(size_t)b * (N + 1) * NX];
    double* ut = &h.ut[(size_t)b * N * NU];
    int32_t status = 4, it = 0, qst = 0;
    h.qp_hook(&P, lin, x0, h.P.nn_rows == SMPC_NN_PARALLEL ? h.cand[b] : h.r[b], xt, ut, &status, &it, &qst, &h.qp_res[(size_t)b * 5]);
    h.qp_iter[b] = it; h.qp_status[b] = qst; h.status[b] = status;
    return status;
  }
  qp.set_ratio_tolerances(o.tol_stat, o.tol_eq, o.tol_ineq, o.tol_comp);
  int qs = qp.solve(o);
  // acados status of a QP exit: solved, or max-iter within qp_maxiter_accept x the tolerances (include/safe_mpc_b200.h)
  auto accepted = [&](int q, const QpSol& sl) {
    const double F = P.qp_maxiter_accept;
    const bool sane = !(F > 0.0) || (sl.res[0] <= F * P.qp_tol_stat && sl.res[1] <= F * P.qp_tol_eq && sl.res[2] <= F * P.qp_tol_ineq && sl.res[3] <= F * P.qp_tol_comp);
    return q == 0 || (q == 1 && sane);
  };
  if (h.probe_eps > 0.0) {
    // margin probe (tests only): is the accept / fail status of this solve decided by rounding?
    //   accepted: the stop test is disabled and the same iteration runs three iterations further; a robust solve stays within the
    //             tolerances, a knife-edge one (a QP at the boundary of feasibility whose multipliers are about to diverge) has left them
    //   failed:   knife-edge when some iterate came within qp_maxiter_accept x the tolerances (another implementation may have stopped there)
    const QpSol keep = qp.sol();
    bool knife;
    if (accepted(qs, keep)) {
      QpOpts o2 = o;
      o2.tol_stat = o2.tol_eq = o2.tol_ineq = o2.tol_comp = -1.0;
      o2.iter_max = keep.iter + 3;
      qp.solve(o2);
      const QpSol& s2 = qp.sol();
      knife = !(s2.res[0] <= P.qp_tol_stat && s2.res[1] <= P.qp_tol_eq && s2.res[2] <= P.qp_tol_ineq && s2.res[3] <= P.qp_tol_comp);
      if (qs == 1) knife = true;
    } else {
      const double F = P.qp_maxiter_accept > 0.0 ? P.qp_maxiter_accept : 1e3;
      knife = keep.best_ratio <= F;
    }
    if (knife) h.probe_flips[b] += 1;
    qp.restore(keep);
  }
  const QpSol& sol = qp.sol();
  h.qp_iter[b] = sol.iter;
  h.qp_status[b] = qs;
  for (int i = 0; i < 4; ++i) h.qp_res[(size_t)b * 5 + i] = sol.res[i];
  h.qp_res[(size_t)b * 5 + 4] = sol.mu;
  // canonical dump of the QP solution
  {
    double* z = &h.qp_z[(size_t)b * (N + 1) * 15];
    double* pi = &h.qp_pi[(size_t)b * N * NX];
    double* lam = &h.qp_lam[(size_t)b * (N + 1) * SMPC_QP_NC];
    double* t = &h.qp_t[(size_t)b * (N + 1) * SMPC_QP_NC];
    std::copy(sol.z.begin(), sol.z.end(), z);
    std::copy(sol.pi.begin(), sol.pi.end(), pi);
    for (int k = 0; k <= N; ++k) {
      const QpStage& S = qp.stages()[k];
      const double* rec = lin + (size_t)k * REC;
      for (int c = 0; c < SMPC_QP_NC; ++c) lam[k * SMPC_QP_NC + c] = t[k * SMPC_QP_NC + c] = 0.0;
      for (int j = 0; j < NX + S.ng; ++j) {
        int row = j < NX ? j : canonical_row(rec, j - NX);
        for (int side = 0; side < 2; ++side) {
          lam[k * SMPC_QP_NC + side * SMPC_QP_NR + row] = sol.lam[k * QNC + side * QNR + j];
          t[k * SMPC_QP_NC + side * SMPC_QP_NR + row] = sol.t[k * QNC + side * QNR + j];
        }
      }
      if (S.soft_row >= 0)
        for (int side = 0; side < 2; ++side) {
          lam[k * SMPC_QP_NC + 2 * SMPC_QP_NR + side] = sol.lam[k * QNC + 2 * QNR + side];
          t[k * SMPC_QP_NC + 2 * SMPC_QP_NR + side] = sol.t[k * QNC + 2 * QNR + side];
        }
    }
  }
  double* xt = &h.xt[(size_t)b * (N + 1) * NX];
  double* ut = &h.ut[(size_t)b * N * NU];
  int status;
  if (accepted(qs, sol)) {   // success or (sane) max-iter: acados SQP_RTI takes the full step
    status = 0;
    bool nan = false;
    for (int k = 0; k <= N; ++k) {
      const int nu = k < N ? NU : 0;
      const double* z = &sol.z[k * QNZ];
      for (int i = 0; i < NX; ++i) { xt[k * NX + i] = xg[k * NX + i] + z[nu + i]; nan |= !(z[nu + i] == z[nu + i]); }
      if (k < N) for (int i = 0; i < NU; ++i) { ut[k * NU + i] = ug[k * NU + i] + z[i]; nan |= !(z[i] == z[i]); }
    }
    if (nan) status = 1;
  } else {                    // min-step / NaN in the QP: QP failure, iterate left at the guess
    status = 4;
    std::copy(xg, xg + (N + 1) * NX, xt);
    std::copy(ug, ug + N * NU, ut);
  }
  h.status[b] = status;
  return status;
}

// provideControl (controller.py:169-184)
void provide_control(orc_handle& h, int b, double* u) {
  const int N = h.P.N;
  double* xg = &h.xg[(size_t)b * (N + 1) * NX];
  double* ug = &h.ug[(size_t)b * N * NU];
  const double* xt = &h.xt[(size_t)b * (N + 1) * NX];
  const double* ut = &h.ut[(size_t)b * N * NU];
  if (h.fails[b] > 0) {
    for (int i = 0; i < NU; ++i) u[i] = ug[i];
    std::memmove(xg, xg + NX, sizeof(double) * N * NX);
    std::memmove(ug, ug + NU, sizeof(double) * (N - 1) * NU);
  } else {
    for (int i = 0; i < NU; ++i) u[i] = ut[i];
    std::memcpy(xg, xt + NX, sizeof(double) * N * NX);
    std::memcpy(ug, ut + NU, sizeof(double) * (N - 1) * NU);
  }
  for (int i = 0; i < NX; ++i) xg[N * NX + i] = xg[(N - 1) * NX + i];
  for (int i = 0; i < NU; ++i) ug[(N - 1) * NU + i] = ug[(N - 2) * NU + i];
}

// checkStateConstraints(x_temp): bounds on every row, collision on row 0 only (env_model.py:170-173,236-243)
bool check_state_constraints_traj(const orc_handle& h, int b) {
  const int N = h.P.N;
  const double* xt = &h.xt[(size_t)b * (N + 1) * NX];
  for (int k = 0; k <= N; ++k) if (!state_in_bounds(h.P, xt + k * NX)) return false;
  return collision_free(h.P, xt);
}

bool check_safe(const orc_handle& h, const double* x) {     // safe_set.py:61-68
  double c = nn_c(h, x, nullptr);
  return (0.0 - h.P.tol_safe <= c) && (c <= 1e6 + h.P.tol_safe);
}

// ParallelController.step (controller.py:614-640): one solve per candidate node n = N .. 1 (sing_step :596-612), best node kept
bool parallel_step_one(orc_handle& h, int b, const double* x, double* u, Workspace& W) {
  const orc_problem_t& P = h.P;
  const int N = P.N;
  double* xg = &h.xg[(size_t)b * (N + 1) * NX];
  double* ug = &h.ug[(size_t)b * N * NU];
  double* xt = &h.xt[(size_t)b * (N + 1) * NX];
  double* ut = &h.ut[(size_t)b * N * NU];
  for (int k = 0; k < N; ++k) f_disc(P.dt, xg + k * NX, ug + k * NU, xg + (k + 1) * NX);      // guessCorrection
  int node_success = 0;
  std::vector<double> bx, bu;
  for (int n = N; n >= 1; --n) {
    h.cand[b] = n;                                                   // constrain_n(n)
    const int status = rti_solve_one(h, b, x, W);
    int checked_r = 0;                                               // check_safe_n (:590-595)
    for (int i = h.r[b]; i <= N; ++i) if (check_safe(h, xt + i * NX)) checked_r = i;
    int result = 0;
    if (status == 0) {
      const int constr_ver = checked_r >= h.r[b] ? checked_r : std::min(n, (int)h.r[b]);
      if (constr_ver - h.r[b] >= 0 && check_state_constraints_traj(h, b)) result = constr_ver;
    }
    if (result > node_success) {
      node_success = result;
      bx.assign(xt, xt + (size_t)(N + 1) * NX); bu.assign(ut, ut + (size_t)N * NU);
      if (result == N) break;
    }
  }
  if (node_success > 1) {
    h.r[b] = node_success;
    std::copy(bx.begin(), bx.end(), xt); std::copy(bu.begin(), bu.end(), ut);
    h.fails[b] = 0;
  } else {
    h.fails[b] += 1;
    if (h.r[b] == 1) {
      for (int i = 0; i < NX; ++i) h.x_viable[(size_t)b * NX + i] = xg[NX + i];
      h.r[b] = N;
      for (int i = 0; i < NU; ++i) u[i] = ug[i];
      return true;
    }
  }
  h.r[b] -= 1;
  h.cur_step[b] += 1;
  provide_control(h, b, u);
  return false;
}

// controller.step(x) for problem b; returns abort flag
// `scripted`: the solve is replaced by a given outcome (status, x_temp, u_temp already stored in the handle) -- used to drive the state
// machines with the very sequences the reference's controller classes were driven with (tests/golden/make_ref_controllers.py)
bool controller_step_one(orc_handle& h, int b, const double* x, double* u, Workspace& W, const int32_t* scripted = nullptr) {
  if (h.P.controller == SMPC_CTRL_PARALLEL) return parallel_step_one(h, b, x, u, W);
  const orc_problem_t& P = h.P;
  const int N = P.N;
  double* xg = &h.xg[(size_t)b * (N + 1) * NX];
  double* ug = &h.ug[(size_t)b * N * NU];
  const int ctrl = P.controller;
  if (ctrl != SMPC_CTRL_REAL_RECEDING)   // guessCorrection (controller.py:226-231)
    for (int k = 0; k < N; ++k) f_disc(P.dt, xg + k * NX, ug + k * NU, xg + (k + 1) * NX);
  const int status = scripted ? (h.status[b] = scripted[b]) : rti_solve_one(h, b, x, W);
  switch (ctrl) {
    case SMPC_CTRL_NAIVE: case SMPC_CTRL_ZEROVEL: case SMPC_CTRL_ST: case SMPC_CTRL_BACKUP:
      if (status == 0) h.fails[b] = 0; else h.fails[b] += 1;
      break;
    case SMPC_CTRL_EVERYWHERE:
      if (status == 0 && check_state_constraints_traj(h, b)) h.fails[b] = 0; else h.fails[b] += 1;
      break;
    case SMPC_CTRL_STWA: case SMPC_CTRL_HTWA:
      if (status == 0 && check_state_constraints_traj(h, b)) h.fails[b] = 0;
      else {
        if (h.fails[b] == 0) for (int i = 0; i < NX; ++i) h.x_viable[(size_t)b * NX + i] = xg[(N - 1) * NX + i];
        if (h.fails[b] == N - 1) { for (int i = 0; i < NU; ++i) u[i] = ug[i]; return true; }
        h.fails[b] += 1;
      }
      break;
    case SMPC_CTRL_RECEDING: case SMPC_CTRL_REAL_RECEDING: {
      if (P.abort_flag) h.r[b] -= 1; else if (h.r[b] > 0) h.r[b] -= 1;
      if (h.r[b] == 0 && P.abort_flag) {
        for (int i = 0; i < NX; ++i) h.x_viable[(size_t)b * NX + i] = xg[NX + i];
        h.r[b] = N;
        for (int i = 0; i < NU; ++i) u[i] = ug[i];
        return true;
      }
      if (status == 0 && check_state_constraints_traj(h, b)) {
        h.fails[b] = 0;
        const double* xt = &h.xt[(size_t)b * (N + 1) * NX];
        for (int i = h.r[b] + 2; i <= N; ++i) if (check_safe(h, xt + i * NX)) h.r[b] = i - 1;
      } else h.fails[b] += 1;
      break;
    }
    default: break;
  }
  h.cur_step[b] += 1;
  provide_control(h, b, u);
  return false;
}

void mass_bias(const orc_problem_t& P, const double inertial[][10], const double* x, double* M, double* bias) {
  double zero[NQ] = {0, 0, 0, 0, 0}, g0[NQ];
  rnea<double>(P, inertial, x, x + NQ, zero, bias);
  rnea<double>(P, inertial, x, zero, zero, g0);
  for (int j = 0; j < NQ; ++j) {
    double e[NQ] = {0, 0, 0, 0, 0}, col[NQ];
    e[j] = 1.0;
    rnea<double>(P, inertial, x, zero, e, col);
    for (int i = 0; i < NQ; ++i) M[i * NQ + j] = col[i] - g0[i];
  }
}

// Torque-input dynamics x' = [v; M(q)^-1 (tau - h(q, v))] on the controller model, one explicit RK4 step, sensitivities by
// forward-mode AD (Dual<15>: 10 state + 5 torque directions) of the plain restatement -- independent of the hand-derived
// inverse-dynamics tangents the CUDA path uses.  Extension row (f)4 of SURVEY.md section 8: the reference has no torque-input
// transcription (env_model.py:58-71 is the double integrator); M and h are its mass_matrix_fun / bias_force_fun (env_model.py:42-43).
template <class T>
void forward_dynamics(const orc_problem_t& P, const T* x, const T* tau, T* a) {
  T zero[NQ], g0[NQ], bias[NQ], M[NQ * NQ], L[NQ * NQ], y[NQ];
  for (int i = 0; i < NQ; ++i) zero[i] = T(0.0);
  rnea<T>(P, P.inertial, x, x + NQ, zero, bias);
  rnea<T>(P, P.inertial, x, zero, zero, g0);
  for (int j = 0; j < NQ; ++j) {
    T e[NQ], col[NQ];
    for (int i = 0; i < NQ; ++i) e[i] = T(i == j ? 1.0 : 0.0);
    rnea<T>(P, P.inertial, x, zero, e, col);
    for (int i = 0; i < NQ; ++i) M[i * NQ + j] = col[i] - g0[i];
  }
  for (int i = 0; i < NQ * NQ; ++i) L[i] = T(0.0);
  for (int j = 0; j < NQ; ++j) {
    T d = M[j * NQ + j];
    for (int k = 0; k < j; ++k) d = d - L[j * NQ + k] * L[j * NQ + k];
    L[j * NQ + j] = sqrt(d);
    for (int i = j + 1; i < NQ; ++i) {
      T s = M[i * NQ + j];
      for (int k = 0; k < j; ++k) s = s - L[i * NQ + k] * L[j * NQ + k];
      L[i * NQ + j] = s / L[j * NQ + j];
    }
  }
  for (int i = 0; i < NQ; ++i) { T s = tau[i] - bias[i]; for (int k = 0; k < i; ++k) s = s - L[i * NQ + k] * y[k]; y[i] = s / L[i * NQ + i]; }
  for (int i = NQ - 1; i >= 0; --i) { T s = y[i]; for (int k = i + 1; k < NQ; ++k) s = s - L[k * NQ + i] * a[k]; a[i] = s / L[i * NQ + i]; }
}
template <class T>
void rk4_step(const orc_problem_t& P, double dt, const T* x, const T* tau, T* xn) {
  T k[4][NX], xi[NX], a[NQ];
  const double c[4] = {0.0, 0.5, 0.5, 1.0};
  for (int s = 0; s < 4; ++s) {
    for (int i = 0; i < NX; ++i) xi[i] = s == 0 ? x[i] : x[i] + k[s - 1][i] * (c[s] * dt);
    forward_dynamics<T>(P, xi, tau, a);
    for (int i = 0; i < NQ; ++i) { k[s][i] = xi[NQ + i]; k[s][NQ + i] = a[i]; }
  }
  for (int i = 0; i < NX; ++i) xn[i] = x[i] + (k[0][i] + k[1][i] * 2.0 + k[2][i] * 2.0 + k[3][i]) * (dt / 6.0);
}

// AdamModel.integrate (env_model.py:192-206)
void plant_step_one(const orc_handle& h, int b, const double* x, const double* u, double* xn, double* a) {
  const orc_problem_t& P = h.P;
  double tau[NQ];
  rnea<double>(P, P.inertial, x, x + NQ, u, tau);           // tau_noisy_fun is built from the NOMINAL model (quirk 3)
  for (int i = 0; i < NQ; ++i) {
    tau[i] += h.tau_noise[(size_t)b * NU + i];
    tau[i] = std::fmin(std::fmax(tau[i], P.tau_min[i]), P.tau_max[i]);
  }
  const double(*I)[10] = reinterpret_cast<const double(*)[10]>(&h.plant_inertial[(size_t)b * NQ * 10]);
  double M[NQ * NQ], bias[NQ], rhs[NQ], L[NQ * NQ];
  mass_bias(P, I, x, M, bias);
  for (int i = 0; i < NQ; ++i) rhs[i] = tau[i] - bias[i];
  // 5x5 SPD solve by Cholesky (the reference uses np.linalg.solve)
  for (int i = 0; i < NQ * NQ; ++i) L[i] = 0.0;
  for (int j = 0; j < NQ; ++j) {
    double d = M[j * NQ + j];
    for (int k = 0; k < j; ++k) d -= L[j * NQ + k] * L[j * NQ + k];
    L[j * NQ + j] = std::sqrt(d);
    for (int i = j + 1; i < NQ; ++i) {
      double s = M[i * NQ + j];
      for (int k = 0; k < j; ++k) s -= L[i * NQ + k] * L[j * NQ + k];
      L[i * NQ + j] = s / L[j * NQ + j];
    }
  }
  double y[NQ];
  for (int i = 0; i < NQ; ++i) { double s = rhs[i]; for (int k = 0; k < i; ++k) s -= L[i * NQ + k] * y[k]; y[i] = s / L[i * NQ + i]; }
  for (int i = NQ - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < NQ; ++k) s -= L[k * NQ + i] * a[k]; a[i] = s / L[i * NQ + i]; }
  f_disc(P.dt, x, a, xn);
}

}  // namespace

// =====================================================================================================
extern "C" {

int orc_create(const orc_problem_t* prob, int32_t batch, int32_t threads, orc_handle_t** out) {
  if (!prob || !out || batch <= 0) return SMPC_ERR_ARG;
  if (prob->nq != NQ || prob->n_pairs != ORC_NPAIR || prob->N < 2 || prob->N > SMPC_MAX_N || prob->n_points > ORC_MAX_POINTS) return SMPC_ERR_UNSUPPORTED;
  if (prob->nn_rows != SMPC_NN_NONE && !prob->nn_weights) return SMPC_ERR_ARG;
  orc_handle* h = new orc_handle;
  h->P = *prob;
  if (prob->nn_weights) { h->w.assign(prob->nn_weights, prob->nn_weights + SMPC_NN_NPARAM); h->P.nn_weights = h->w.data(); }
  const int B = batch, N = prob->N;
  h->B = B; h->N = N;
  h->threads = threads > 0 ? threads : std::max(1u, std::thread::hardware_concurrency());
  h->xg.assign((size_t)B * (N + 1) * NX, 0.0); h->ug.assign((size_t)B * N * NU, 0.0);
  h->xt = h->xg; h->ut = h->ug;
  h->lin.assign((size_t)B * (N + 1) * REC, 0.0);
  h->plant_inertial.resize((size_t)B * NQ * 10);
  for (int b = 0; b < B; ++b) std::memcpy(&h->plant_inertial[(size_t)b * NQ * 10], prob->inertial, sizeof(double) * NQ * 10);
  h->tau_noise.assign((size_t)B * NU, 0.0);
  h->x_viable.assign((size_t)B * NX, 0.0);
  h->qp_z.assign((size_t)B * (N + 1) * 15, 0.0); h->qp_pi.assign((size_t)B * N * NX, 0.0);
  h->qp_lam.assign((size_t)B * (N + 1) * SMPC_QP_NC, 0.0); h->qp_t = h->qp_lam;
  h->qp_res.assign((size_t)B * 5, 0.0);
  h->fails.assign(B, 0); h->r.assign(B, N); h->cand.assign(B, N); h->status.assign(B, 4); h->qp_iter.assign(B, 0); h->qp_status.assign(B, 0); h->cur_step.assign(B, 0);
  h->ws.resize(h->threads);
  *out = h;
  return SMPC_OK;
}
void orc_destroy(orc_handle_t* h) { delete h; }
int orc_num_threads(const orc_handle_t* h) { return h->threads; }
int orc_set_mlp_fp32(orc_handle_t* h, int32_t on) { h->mlp_fp32 = on; return SMPC_OK; }
int orc_set_qp_hook(orc_handle_t* h, orc_qp_hook_t fn) { h->qp_hook = fn; return SMPC_OK; }
int orc_set_probe(orc_handle_t* h, double eps) { h->probe_eps = eps; h->probe_flips.assign(h->B, 0); return SMPC_OK; }
int orc_get_probe_flips(orc_handle_t* h, int32_t* out) {
  if (h->probe_flips.size() != (size_t)h->B) h->probe_flips.assign(h->B, 0);
  std::copy(h->probe_flips.begin(), h->probe_flips.end(), out);
  return SMPC_OK;
}

int orc_set_plant_inertial(orc_handle_t* h, const double* v) { std::copy(v, v + h->plant_inertial.size(), h->plant_inertial.begin()); return SMPC_OK; }
int orc_set_ee_trajectory(orc_handle_t* h, const double* traj, int32_t n) {
  if (n < 0 || (n > 0 && !traj)) return SMPC_ERR_ARG;
  h->ee_traj.assign(traj, traj + 3 * (size_t)n);
  return SMPC_OK;
}
int orc_set_torque_noise(orc_handle_t* h, const double* v) { std::copy(v, v + h->tau_noise.size(), h->tau_noise.begin()); return SMPC_OK; }
int orc_set_guess(orc_handle_t* h, const double* xg, const double* ug) {
  std::copy(xg, xg + h->xg.size(), h->xg.begin());
  std::copy(ug, ug + h->ug.size(), h->ug.begin());
  const int N = h->N;
  for (int b = 0; b < h->B; ++b)   // STWAController.setGuess (controller.py:390-393)
    for (int i = 0; i < NX; ++i) h->x_viable[(size_t)b * NX + i] = xg[((size_t)b * (N + 1) + N) * NX + i];
  return SMPC_OK;
}
int orc_get_guess(orc_handle_t* h, double* xg, double* ug) {
  if (xg) std::copy(h->xg.begin(), h->xg.end(), xg);
  if (ug) std::copy(h->ug.begin(), h->ug.end(), ug);
  return SMPC_OK;
}
int orc_get_temp(orc_handle_t* h, double* xt, double* ut) {
  if (xt) std::copy(h->xt.begin(), h->xt.end(), xt);
  if (ut) std::copy(h->ut.begin(), h->ut.end(), ut);
  return SMPC_OK;
}
int orc_reset_controller(orc_handle_t* h) {
  std::fill(h->fails.begin(), h->fails.end(), 0);
  std::fill(h->r.begin(), h->r.end(), h->N);
  std::fill(h->cur_step.begin(), h->cur_step.end(), 0);
  return SMPC_OK;
}

int orc_rti_solve(orc_handle_t* h, const double* x0, const uint8_t* active, int32_t* status) {
  parallel_for(h->B, h->threads, [&](int b, int tid) {
    if (active && !active[b]) return;
    rti_solve_one(*h, b, x0 + (size_t)b * NX, h->ws[tid]);
  });
  if (status) std::copy(h->status.begin(), h->status.end(), status);
  return SMPC_OK;
}

int orc_controller_step(orc_handle_t* h, const double* x, const uint8_t* active, double* u, uint8_t* abort_flag) {
  parallel_for(h->B, h->threads, [&](int b, int tid) {
    if (active && !active[b]) return;
    bool ab = controller_step_one(*h, b, x + (size_t)b * NX, u + (size_t)b * NU, h->ws[tid]);
    if (abort_flag) abort_flag[b] = ab ? 1 : 0;
  });
  return SMPC_OK;
}

// controller.step(x) with the solve replaced by a scripted outcome: status[B], x_temp[B][N+1][NX], u_temp[B][N][NU]
int orc_controller_step_scripted(orc_handle_t* h, const double* x, const int32_t* status, const double* xt, const double* ut, double* u,
                                 uint8_t* abort_flag) {
  if (h->P.controller == SMPC_CTRL_PARALLEL) return SMPC_ERR_UNSUPPORTED;
  std::copy(xt, xt + h->xt.size(), h->xt.begin());
  std::copy(ut, ut + h->ut.size(), h->ut.begin());
  parallel_for(h->B, h->threads, [&](int b, int tid) {
    bool ab = controller_step_one(*h, b, x + (size_t)b * NX, u + (size_t)b * NU, h->ws[tid], status);
    if (abort_flag) abort_flag[b] = ab ? 1 : 0;
  });
  return SMPC_OK;
}

int orc_plant_step(orc_handle_t* h, const double* x, const double* u, double* xn, double* a) {
  parallel_for(h->B, h->threads, [&](int b, int) {
    double aa[NQ];
    plant_step_one(*h, b, x + (size_t)b * NX, u + (size_t)b * NU, xn + (size_t)b * NX, aa);
    if (a) for (int i = 0; i < NQ; ++i) a[(size_t)b * NU + i] = aa[i];
  });
  return SMPC_OK;
}

int orc_tau(orc_handle_t* h, int32_t n, const double* x, const double* u, double* tau) {
  for (int i = 0; i < n; ++i) rnea<double>(h->P, h->P.inertial, x + (size_t)i * NX, x + (size_t)i * NX + NQ, u + (size_t)i * NU, tau + (size_t)i * NU);
  return SMPC_OK;
}
int orc_rk4_sens(orc_handle_t* h, int32_t n, const double* x, const double* tau, double dt, double* xn, double* A, double* B) {
  using D = orc::Dual<NX + NU>;
  parallel_for(n, h->threads, [&](int i, int) {
    D xd[NX], td[NU], out[NX];
    for (int j = 0; j < NX; ++j) xd[j] = D::var(x[(size_t)i * NX + j], j);
    for (int j = 0; j < NU; ++j) td[j] = D::var(tau[(size_t)i * NU + j], NX + j);
    rk4_step<D>(h->P, dt, xd, td, out);
    for (int r = 0; r < NX; ++r) {
      xn[(size_t)i * NX + r] = out[r].v;
      if (A) for (int j = 0; j < NX; ++j) A[((size_t)i * NX + r) * NX + j] = out[r].d[j];
      if (B) for (int j = 0; j < NU; ++j) B[((size_t)i * NX + r) * NU + j] = out[r].d[NX + j];
    }
  });
  return SMPC_OK;
}
int orc_kinematics(orc_handle_t* h, int32_t n, const double* x, double* ee, double* dist) {
  for (int i = 0; i < n; ++i) distances(h->P, x + (size_t)i * NX, ee ? ee + (size_t)i * 3 : nullptr, dist ? dist + (size_t)i * ORC_NPAIR : nullptr);
  return SMPC_OK;
}
int orc_nn_constraint(orc_handle_t* h, int32_t n, const double* x, double* c, double* grad) {
  if (!h->P.nn_weights) return SMPC_ERR_ARG;
  parallel_for(n, h->threads, [&](int i, int) { c[i] = nn_c(*h, x + (size_t)i * NX, grad ? grad + (size_t)i * NX : nullptr); });
  return SMPC_OK;
}
int orc_get_lin(orc_handle_t* h, double* lin) { std::copy(h->lin.begin(), h->lin.end(), lin); return SMPC_OK; }
int orc_get_qp(orc_handle_t* h, double* dz, double* pi, double* lam, double* t) {
  if (dz) std::copy(h->qp_z.begin(), h->qp_z.end(), dz);
  if (pi) std::copy(h->qp_pi.begin(), h->qp_pi.end(), pi);
  if (lam) std::copy(h->qp_lam.begin(), h->qp_lam.end(), lam);
  if (t) std::copy(h->qp_t.begin(), h->qp_t.end(), t);
  return SMPC_OK;
}
static std::vector<int32_t>* state_field(orc_handle_t* h, int32_t f) {
  switch (f) {
    case SMPC_STATE_FAILS: return &h->fails;
    case SMPC_STATE_R: return &h->r;
    case SMPC_STATE_STATUS: return &h->status;
    case SMPC_STATE_QP_ITER: return &h->qp_iter;
    case SMPC_STATE_QP_STATUS: return &h->qp_status;
  }
  return nullptr;
}
int orc_get_state_i32(orc_handle_t* h, int32_t f, int32_t* out) {
  auto* v = state_field(h, f);
  if (!v) return SMPC_ERR_ARG;
  std::copy(v->begin(), v->end(), out);
  return SMPC_OK;
}
int orc_set_state_i32(orc_handle_t* h, int32_t f, const int32_t* in) {
  auto* v = state_field(h, f);
  if (!v) return SMPC_ERR_ARG;
  std::copy(in, in + h->B, v->begin());
  return SMPC_OK;
}
int orc_get_qp_residuals(orc_handle_t* h, double* res5) { std::copy(h->qp_res.begin(), h->qp_res.end(), res5); return SMPC_OK; }
int orc_get_x_viable(orc_handle_t* h, double* xv) { std::copy(h->x_viable.begin(), h->x_viable.end(), xv); return SMPC_OK; }

int orc_mass_bias(orc_handle_t* h, int32_t b, int32_t nominal, const double* x, double* M, double* bias) {
  const double(*I)[10] = nominal ? h->P.inertial : reinterpret_cast<const double(*)[10]>(&h->plant_inertial[(size_t)b * NQ * 10]);
  mass_bias(h->P, I, x, M, bias);
  return SMPC_OK;
}
int orc_qp_info(orc_handle_t* h, int32_t b, double* res4, double* mu, int32_t* iter, int32_t* status) {
  for (int i = 0; i < 4; ++i) res4[i] = h->qp_res[(size_t)b * 5 + i];
  *mu = h->qp_res[(size_t)b * 5 + 4];
  *iter = h->qp_iter[b];
  *status = h->qp_status[b];
  return SMPC_OK;
}

// ------------------------------------------------------------------------------------------ closed loop
int orc_sim_create(orc_handle_t* c, orc_handle_t* bk, int32_t n_steps, orc_sim_t** out) {
  if (!c || !bk || c->B != bk->B || n_steps <= 0) return SMPC_ERR_ARG;
  orc_sim* s = new orc_sim;
  s->c = c; s->bk = bk; s->n_steps = n_steps;
  const int B = c->B, Nb = bk->N;
  s->x.assign((size_t)B * NX, 0.0);
  s->xlog.assign((size_t)B * (n_steps + 1) * NX, 0.0);
  s->ulog.assign((size_t)B * n_steps * NU, 0.0);
  s->x_abort.assign((size_t)B * (Nb + 1) * NX, 0.0);
  s->u_abort.assign((size_t)B * Nb * NU, 0.0);
  s->xv_first.assign((size_t)B * NX, 0.0);
  s->mode.assign(B, 0); s->ja.assign(B, 0); s->outcome.assign(B, 0);
  *out = s;
  return SMPC_OK;
}
void orc_sim_destroy(orc_sim_t* s) { delete s; }

int orc_sim_reset(orc_sim_t* s, const double* x_init) {
  const int B = s->c->B;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::fill(s->xlog.begin(), s->xlog.end(), nan);
  std::fill(s->ulog.begin(), s->ulog.end(), nan);
  std::fill(s->xv_first.begin(), s->xv_first.end(), nan);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < NX; ++i) { s->x[(size_t)b * NX + i] = x_init[(size_t)b * NX + i]; s->xlog[(size_t)b * (s->n_steps + 1) * NX + i] = x_init[(size_t)b * NX + i]; }
  std::fill(s->mode.begin(), s->mode.end(), 0);
  std::fill(s->ja.begin(), s->ja.end(), 0);
  std::fill(s->outcome.begin(), s->outcome.end(), 0);
  s->j = 0;
  for (int i = 0; i < 4; ++i) s->counters[i] = 0;
  return SMPC_OK;
}

// one control step of every live problem (scripts/mpc.py:125-264)
int orc_sim_step(orc_sim_t* s) {
  if (s->j >= s->n_steps) return SMPC_ERR_ARG;
  orc_handle& C = *s->c;
  orc_handle& K = *s->bk;
  const int B = C.B, Nb = K.N, j = s->j;
  const double kp = 1.0, kd = 1e2;
  std::atomic<int64_t> n_rti(0), n_bk(0), n_ps(0), n_ipm(0);
  if ((int)K.ws.size() < C.threads) K.ws.resize(C.threads);
  const int32_t* scr = s->scr_status;
  if (scr) { std::copy(s->scr_xt, s->scr_xt + C.xt.size(), C.xt.begin()); std::copy(s->scr_ut, s->scr_ut + C.ut.size(), C.ut.begin()); }
  parallel_for(B, C.threads, [&](int b, int tid) {
    if (s->mode[b] == 2) return;
    Workspace& W = C.ws[tid];
    Workspace& WK = K.ws[tid];
    const double* x = &s->x[(size_t)b * NX];
    double u[NU];
    const double* xa = &s->x_abort[(size_t)b * (Nb + 1) * NX];
    const double* ua = &s->u_abort[(size_t)b * Nb * NU];
    bool done = false;
    if (s->mode[b] == 1) {               // following the safe-abort trajectory (mpc.py:130-146)
      const int ja = s->ja[b];
      if (ja < Nb) {
        for (int i = 0; i < NQ; ++i) u[i] = ua[ja * NU + i] - (kp * (x[i] - xa[ja * NX + i]) + kd * (x[NQ + i] - xa[ja * NX + NQ + i]));
      } else {
        bool slow = true;
        for (int i = 0; i < NQ; ++i) slow &= (x[NQ + i] < 5e-3);   // no abs(), as upstream
        if (slow) {
          bool sa = controller_step_one(C, b, x, u, W, scr);
          ++n_rti; n_ipm += C.qp_iter[b];
          s->mode[b] = sa ? 1 : 0;        // a repeated abort here does NOT re-solve the backup OCP (mpc.py:138-141)
        } else {
          for (int i = 0; i < NQ; ++i) u[i] = -(kp * (x[i] - xa[Nb * NX + i]) + 3e2 * (x[NQ + i] - xa[Nb * NX + NQ + i]));
        }
      }
      s->ja[b] = ja + 1;
    } else {
      bool sa = controller_step_one(C, b, x, u, W, scr);
      ++n_rti; n_ipm += C.qp_iter[b];
      if (sa) {                           // mpc.py:161-190
        const double* xv = &C.x_viable[(size_t)b * NX];
        if (!(s->outcome[b] & SMPC_OUT_ABORTED)) for (int i = 0; i < NX; ++i) s->xv_first[(size_t)b * NX + i] = xv[i];
        double* xg = &K.xg[(size_t)b * (Nb + 1) * NX];
        double* ug = &K.ug[(size_t)b * Nb * NU];
        for (int k = 0; k <= Nb; ++k) for (int i = 0; i < NX; ++i) xg[k * NX + i] = xv[i];
        for (int i = 0; i < Nb * NU; ++i) ug[i] = 0.0;
        int st;
        if (s->scr_bk_status) {
          st = s->scr_bk_status[b];
          std::copy(s->scr_bk_xt + (size_t)b * (Nb + 1) * NX, s->scr_bk_xt + (size_t)(b + 1) * (Nb + 1) * NX, &K.xt[(size_t)b * (Nb + 1) * NX]);
          std::copy(s->scr_bk_ut + (size_t)b * Nb * NU, s->scr_bk_ut + (size_t)(b + 1) * Nb * NU, &K.ut[(size_t)b * Nb * NU]);
        } else st = rti_solve_one(K, b, xv, WK);
        ++n_bk; n_ipm += K.qp_iter[b];
        if (st != 0) { s->outcome[b] |= SMPC_OUT_COLLIDED; done = true; }
        else {
          s->ja[b] = 0;
          s->outcome[b] |= SMPC_OUT_ABORTED;
          s->mode[b] = 1;
          std::copy(&K.xt[(size_t)b * (Nb + 1) * NX], &K.xt[(size_t)(b + 1) * (Nb + 1) * NX], &s->x_abort[(size_t)b * (Nb + 1) * NX]);
          std::copy(&K.ut[(size_t)b * Nb * NU], &K.ut[(size_t)(b + 1) * Nb * NU], &s->u_abort[(size_t)b * Nb * NU]);
        }
      }
    }
    for (int i = 0; i < NU; ++i) s->ulog[((size_t)b * s->n_steps + j) * NU + i] = u[i];
    if (done) { s->mode[b] = 2; return; }
    double xn[NX], a[NQ];
    plant_step_one(C, b, x, u, xn, a);
    ++n_ps;
    for (int i = 0; i < NX; ++i) s->xlog[((size_t)b * (s->n_steps + 1) + j + 1) * NX + i] = xn[i];
    if (!state_in_bounds(C.P, xn) || !collision_free(C.P, xn)) { s->outcome[b] |= SMPC_OUT_COLLIDED; s->mode[b] = 2; }
    for (int i = 0; i < NX; ++i) s->x[(size_t)b * NX + i] = xn[i];
  });
  s->counters[0] += n_rti; s->counters[1] += n_bk; s->counters[2] += n_ps; s->counters[3] += n_ipm;
  s->j += 1;
  return SMPC_OK;
}
// tests only: the solves of the next orc_sim_step calls are replaced by these outcomes (NULL status: real solves again)
int orc_sim_set_script(orc_sim_t* s, const int32_t* status, const double* xt, const double* ut, const int32_t* bk_status, const double* bk_xt,
                       const double* bk_ut) {
  s->scr_status = status; s->scr_xt = xt; s->scr_ut = ut; s->scr_bk_status = bk_status; s->scr_bk_xt = bk_xt; s->scr_bk_ut = bk_ut;
  return SMPC_OK;
}
int orc_sim_run(orc_sim_t* s, int32_t n) {
  for (int i = 0; i < n; ++i) { int rc = orc_sim_step(s); if (rc) return rc; }
  return SMPC_OK;
}
int orc_sim_get_outcome(orc_sim_t* s, int32_t* out) {
  const orc_problem_t& P = s->c->P;
  for (int b = 0; b < s->c->B; ++b) {
    const double* xl = &s->xlog[((size_t)b * (s->n_steps + 1) + s->n_steps) * NX];
    double ee[3], d2 = 0.0;
    distances(P, xl, ee, nullptr);
    for (int k = 0; k < 3; ++k) d2 += (ee[k] - P.ee_ref[k]) * (ee[k] - P.ee_ref[k]);
    int o = s->outcome[b] & ~SMPC_OUT_CONVERGED;
    if (std::sqrt(d2) < P.tol_conv) o |= SMPC_OUT_CONVERGED;   // NaN (early exit) compares false
    out[b] = o;
  }
  return SMPC_OK;
}
int orc_sim_get_log(orc_sim_t* s, double* x, double* u) {
  if (x) std::copy(s->xlog.begin(), s->xlog.end(), x);
  if (u) std::copy(s->ulog.begin(), s->ulog.end(), u);
  return SMPC_OK;
}
int orc_sim_get_x_viable(orc_sim_t* s, double* xv) { std::copy(s->xv_first.begin(), s->xv_first.end(), xv); return SMPC_OK; }
int orc_sim_get_counters(orc_sim_t* s, int64_t* out4) { for (int i = 0; i < 4; ++i) out4[i] = s->counters[i]; return SMPC_OK; }

}  // extern "C"
