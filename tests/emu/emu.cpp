// TEST HARNESS -- host compilation of the engine's kernel sources (safe_mpc_b200/csrc/*.cuh) so that the
// kernel logic can be checked against the oracle on a machine without a GPU.  Not part of the product: nothing in
// safe_mpc_b200/ loads this library, and it is not a fallback -- the product path fails without the CUDA library.
#include <cstring>
#include <vector>

#include "../../safe_mpc_b200/csrc/dev_model.cuh"
#ifdef EMU_QP
#include <cstdio>
#include <cstdlib>

#include "../../safe_mpc_b200/csrc/qp_split.cuh"
#endif

using namespace smpc;

extern "C" {

int emu_linearize(const smpc_problem_t* P, int n, const int* k, const double* x, const double* u, const double* xnext,
                  const int* gate, const double* nn11, double* rec) {
  for (int i = 0; i < n; ++i)
    linearize_stage(*P, k[i], x + i * NX, u + i * NU, xnext + i * NX, stage_has_nn(*P, k[i]), gate[i] != 0, nn11 + i * 11, rec + (size_t)i * REC);
  return 0;
}

// the same with a per-item end-effector reference (smpc_set_ee_trajectory: row current_step + k of the trajectory), ee_ref [n][3]
int emu_linearize_ref(const smpc_problem_t* P, int n, const int* k, const double* x, const double* u, const double* xnext,
                      const int* gate, const double* nn11, const double* ee_ref, double* rec) {
  for (int i = 0; i < n; ++i)
    linearize_stage(*P, k[i], x + i * NX, u + i * NU, xnext + i * NX, stage_has_nn(*P, k[i]), gate[i] != 0, nn11 + i * 11, rec + (size_t)i * REC, 1,
                    ee_ref + (size_t)i * 3);
  return 0;
}

int emu_plant_step(const smpc_problem_t* P, int n, const double* inertial, const double* noise, const double* x, const double* u,
                   double* xn, double* a) {
  for (int i = 0; i < n; ++i)
    plant_step(*P, reinterpret_cast<const double(*)[10]>(inertial + (size_t)i * NQ * 10), noise + i * NU, x + i * NX, u + i * NU, xn + i * NX, a + i * NU);
  return 0;
}

int emu_rk4_sens(const smpc_problem_t* P, int n, const double* x, const double* tau, double dt, double* xn, double* A, double* B) {
  for (int i = 0; i < n; ++i) {
    if (A && B) rk4_sens<true>(*P, P->inertial, dt, x + i * NX, tau + i * NU, xn + i * NX, A + (size_t)i * NX * NX, B + (size_t)i * NX * NU);
    else rk4_sens<false>(*P, P->inertial, dt, x + i * NX, tau + i * NU, xn + i * NX, nullptr, nullptr);
  }
  return 0;
}

int emu_checks(const smpc_problem_t* P, int n, const double* x, int* in_bounds, int* coll_free, double* ee, double* dist) {
  for (int i = 0; i < n; ++i) {
    in_bounds[i] = state_in_bounds(*P, x + i * NX);
    coll_free[i] = collision_free(*P, x + i * NX);
    distances(*P, x + i * NX, ee + i * 3, dist + i * NPAIR);
  }
  return 0;
}

}  // extern "C"

#ifdef EMU_QP
namespace {
// ----------------------------------------------------------------------------------------------------------------
// Host stand-in of the kernel launches of qp.cu: every phase of the split IPM (qp_split.cuh) is run as plain loops
// over (tile, stage, lane) in a caller-chosen order (ascending / descending), so that a phase that reads what
// another work item of the same phase writes shows up as an order-dependent result in the parity test.
// ----------------------------------------------------------------------------------------------------------------
// host stand-in of the staged copies of the Riccati sweeps: one lane at a time; lazy = 1 performs a copy only when it
// is waited for (a fetch issued before its source has been written then shows up as wrong numbers)
struct HostStage {
  int ln, nfb, lazy, nb = 2;
  int nbuf() const { return nb; }
  std::vector<qs_real> mem;
  struct Copy { qs_real* dst; const qs_real* src; size_t n; };
  std::vector<Copy> pend[2];
  HostStage(int lane_, int nfb_, int lazy_) : ln(lane_), nfb(nfb_), lazy(lazy_), mem((size_t)2 * nfb_ * TL, (qs_real)0) {}
  int lane() const { return ln; }
  bool any(bool v) const { return v; }
  void sync() const {}
  qs_real* buf(int b) { return mem.data() + (size_t)b * nfb * TL + ln; }
  void fetch_begin(int, int) {}
  void fetch(int b, int dst_field, const qs_real* gblock, int src_field, int nfields) {
    Copy c{mem.data() + ((size_t)b * nfb + dst_field) * TL, gblock + (size_t)src_field * TL, (size_t)nfields * TL};
    if (dst_field + nfields > nfb) { fprintf(stderr, "emu: staging overflow\n"); abort(); }
    if (lazy) pend[b].push_back(c); else std::memcpy(c.dst, c.src, c.n * sizeof(qs_real));
  }
  void wait(int b) { for (auto& c : pend[b]) std::memcpy(c.dst, c.src, c.n * sizeof(qs_real)); pend[b].clear(); }
  void publish() const {}
  void prefetch(const qs_real*, int, int) const {}
};

struct HostBackend {
  const smpc_problem_t& P;
  QsBufs q;
  int B, T, N, order;
  const double* x0; const int32_t* r; const uint8_t* act;
  double *xt, *ut; int32_t *status, *qp_iter, *qp_status; double* qp_res;
  std::vector<double> psm;
  int n_active = 0;
  template <class F> void each_stage(F f) {
    for (int t = 0; t < T; ++t)
      for (int kx = 0; kx <= N; ++kx) {
        const int k = order ? N - kx : kx;
        for (int lx = 0; lx < TL; ++lx) { const int l = order ? TL - 1 - lx : lx; if (t * TL + l < B + lane_pad) f(t, l, k); }
      }
  }
  template <class F> void each_problem(F f) {
    for (int t = 0; t < T; ++t) for (int lx = 0; lx < TL; ++lx) { const int l = order ? TL - 1 - lx : lx; if (t * TL + l < B + lane_pad) f(t, l); }
  }
  int lane_pad = TL;            // padding lanes past the batch are visited too (they must stay silent); 0 after init() when skip_pad is set
  bool skip_pad = false;
  void init() { lane_pad = TL; each_problem([&](int t, int l) { qs_init(q, t, l, B, x0, r, act); }); if (skip_pad) lane_pad = 0; }
  void prep(int kk) {
    std::vector<double> jsm((size_t)PREP_SCRATCH * TL, 0.0);
    each_stage([&](int t, int l, int k) {
      if (kk == 0) qs_prep<true>(P, q, t, l, k, kk, jsm.data() + l); else qs_prep<false>(P, q, t, l, k, kk, jsm.data() + l);
    });
  }
  void ctl(int kk) {
    n_active = 0;
    each_problem([&](int t, int l) { if (qs_ctl(P, q, t, l, kk, status, qp_iter, qp_status, qp_res)) ++n_active; });
  }
  void ric1() { each_problem([&](int t, int l) { HostStage w(l, RIC1_STAGE_FIELDS, order); w.nb = 2 - order; qs_ric1(P, q, t, w, psm.data() + l); }); }
  void ric2() { each_problem([&](int t, int l) { HostStage w(l, RIC2_STAGE_FIELDS, order); w.nb = 2 - order; qs_ric2(P, q, t, w); }); }
  void step(int kk, int mode) { each_stage([&](int t, int l, int k) { qs_step(P, q, t, l, k, kk, mode); }); }
  void final() { each_stage([&](int t, int l, int k) { const int b = qs_final(q, t, l, k, act, B, status, xt, ut); if (b >= 0) status[b] = 1; }); }
  // compaction of the slots (qp_split.cuh): in the test harness after every iteration that leaves a hole, when enabled
  int compact_mode = 0, n_compactions = 0, n_moves = 0;
  void compact(int kk) {
    if (!compact_mode) return;
    final();                                                   // results of the finished problems, before their slots are reused
    std::vector<int32_t> mv((size_t)T * TL, 0);
    int na = 0;
    const int nm = qs_compact_plan(q, T, mv.data(), &na);
    const int half = T * TL / 2;
    for (int i = 0; i < nm; ++i) {
      for (int k = 0; k <= N; ++k) for (int f = 0; f < CMP_FIELDS; ++f) qs_compact_move_field(q, mv[i], mv[half + i], k, kk, f);
      qs_compact_move_scalars(q, mv[i], mv[half + i]);
    }
    if (nm) { ++n_compactions; n_moves += nm; }
  }
  bool solo(int) { return false; }
  void red(bool after) { each_problem([&](int t, int l) { qs_red(P, q, t, l, after); }); }
  int redo_total = 0;
  void request_counters(int) {}
  bool wait_counters(int, bool, int& na, int& nr) {
    na = n_active; nr = 0;
    each_problem([&](int t, int l) { const int32_t* pi = q.pi + qs_pb(t, NPI, l); if (QF(pi, J_ACT) && QF(pi, J_REDO)) ++nr; });
    redo_total += nr;
    return true;
  }
};
}  // namespace

// loop options of the following solves (tests): run-ahead depth of QsLoop, compaction after every iteration that leaves a hole
static int g_depth = 0, g_compact = 0;
extern "C" void emu_set_options(int depth, int compact) { g_depth = depth; g_compact = compact; }
static int g_last_compactions = 0, g_last_moves = 0;
extern "C" void emu_last_compactions(int* n, int* moves) { *n = g_last_compactions; *moves = g_last_moves; }

// Batched QP solve with the engine's kernel sources.  rec: [B][N+1][REC] (caller layout), x0: [B][10], r: [B].
// Outputs in the layout of smpc_get_qp: z [B][N+1][15] ([du;dx], terminal stage: dx first), pi [B][N][10],
// lam / t [B][N+1][SMPC_QP_NC]; plus x_temp / u_temp / status as smpc_rti_solve produces them.
extern "C" int emu_qp_solve(const smpc_problem_t* P, int B, const double* rec, const double* x0, const int32_t* r, const uint8_t* act,
                            int order, double* z, double* pi, double* lam, double* t, double* xt, double* ut, int32_t* status,
                            int32_t* qp_iter, int32_t* qp_status, double* qp_res, int* n_redo_total) {
  const int N = P->N, T = (B + TL - 1) / TL;
  const size_t S = (size_t)T * (N + 1) * TL;
  std::vector<qs_real> vrec(S * REC, (qs_real)0), st(S * NIT, (qs_real)0), st2(S * NS2, (qs_real)0), prod(S * NPROD, (qs_real)0), sb(S * NSB, (qs_real)0);
  std::vector<double> it0(S * NIT, 0.0), it1(S * NIT, 0.0), res(S * NRES, 0.0), stp(S * NSTP, 0.0), pd((size_t)T * NPD * TL, 0.0);
  std::vector<int32_t> pi32((size_t)T * NPI * TL, 0);
  for (int b = 0; b < B; ++b)
    for (int k = 0; k <= N; ++k)
      for (int f = 0; f < REC; ++f)
        vrec[qs_blk(b / TL, N, k, REC, b % TL) + (size_t)f * TL] = (qs_real)rec[((size_t)b * (N + 1) + k) * REC + f];
  QsBufs q{vrec.data(), {it0.data(), it1.data()}, st.data(), st2.data(), sb.data(), prod.data(), res.data(), stp.data(),
           pd.data(), pi32.data(), N, 0};
  HostBackend bk{*P, q, B, T, N, order, x0, r, act, xt, ut, status, qp_iter, qp_status, qp_res, std::vector<double>(65 * TL, 0.0)};
  // two groups of tiles when there is more than one tile (exercises the group offsets), driven round-robin
  const int G = T > 1 ? 2 : 1;
  std::vector<HostBackend> gb;
  for (int g = 0; g < G; ++g) {
    const int t0 = g == 0 ? 0 : T / 2, t1 = (g == G - 1) ? T : T / 2;
    HostBackend b2 = bk;
    const size_t so = (size_t)t0 * (N + 1) * TL;
    b2.q.rec = q.rec + so * REC; b2.q.it[0] = q.it[0] + so * NIT; b2.q.it[1] = q.it[1] + so * NIT; b2.q.st = q.st + so * NIT; b2.q.st2 = q.st2 + so * NS2;
    b2.q.sb = q.sb + so * NSB; b2.q.prod = q.prod + so * NPROD; b2.q.res = q.res + so * NRES; b2.q.stp = q.stp + so * NSTP;
    b2.q.pd = q.pd + (size_t)t0 * NPD * TL; b2.q.pi = q.pi + (size_t)t0 * NPI * TL; b2.q.tile0 = t0;
    b2.T = t1 - t0;
    b2.compact_mode = g_compact;
    gb.push_back(b2);
  }
  qs_drive(gb.data(), G, g_depth);
  if (n_redo_total) { *n_redo_total = 0; for (auto& b2 : gb) *n_redo_total += b2.redo_total; }
  g_last_compactions = 0; g_last_moves = 0;
  for (auto& b2 : gb) { g_last_compactions += b2.n_compactions; g_last_moves += b2.n_moves; }
  // the canonical dump reads every problem at its original slot: only meaningful when no slot was reused
  for (int b = 0; b < B && g_last_compactions == 0; ++b) {
    const int tile = b / TL, lane = b % TL;
    const int buf = QF(pi32.data() + qs_pb(tile, NPI, lane), J_ITBUF);
    for (int k = 0; k <= N; ++k) {
      const double* it = q.it[buf] + qs_blk(tile, N, k, NIT, lane);
      const StageFlags F = qs_flags(*P, k);
      double* zo = z + ((size_t)b * (N + 1) + k) * 15;
      if (k < N) for (int i = 0; i < 15; ++i) zo[i] = QF(it, I_Z + i);
      else { for (int i = 0; i < 10; ++i) zo[i] = QF(it, I_Z + 5 + i); for (int i = 10; i < 15; ++i) zo[i] = 0.0; }
      if (k > 0) for (int i = 0; i < 10; ++i) pi[((size_t)b * N + k - 1) * 10 + i] = QF(it, I_PIM + i);
      double* ol = lam + ((size_t)b * (N + 1) + k) * SMPC_QP_NC;
      double* ot = t + ((size_t)b * (N + 1) + k) * SMPC_QP_NC;
      for (int j = 0; j < QNR; ++j) {
        const bool p = j < 10 ? true : (j < 15 ? F.tau : (j < 21 ? F.dist : F.nn));
        for (int s = 0; s < 2; ++s) { ol[s * QNR + j] = p ? QF(it, I_LAM + s * QNR + j) : 0.0; ot[s * QNR + j] = p ? QF(it, I_T + s * QNR + j) : 0.0; }
      }
      ol[2 * QNR] = F.soft ? QF(it, I_SLK + 2) : 0.0; ol[2 * QNR + 1] = F.soft ? QF(it, I_SLK + 3) : 0.0;
      ot[2 * QNR] = F.soft ? QF(it, I_SLK + 4) : 0.0; ot[2 * QNR + 1] = F.soft ? QF(it, I_SLK + 5) : 0.0;
    }
  }
  return 0;
}

// One problem at a time with a per-thread workspace: the signature of oracle.h's orc_qp_hook_t, so that the oracle's closed loop can
// run on "kernel arithmetic" (tests/test_closed_loop_emulated.py, scripts/closed_loop_probe.py).  Lane 0 of one tile carries the problem.
extern "C" int emu_qp_solve1(const smpc_problem_t* P, const double* rec, const double* x0, int32_t r, double* xt, double* ut, int32_t* status,
                             int32_t* qp_iter, int32_t* qp_status, double* qp_res) {
  struct Ws {
    int N = -1;
    std::vector<qs_real> vrec, st, st2, prod, sb;
    std::vector<double> it0, it1, res, stp, pd, psm;
    std::vector<int32_t> pi32;
  };
  thread_local Ws w;
  const int N = P->N;
  const size_t S = (size_t)(N + 1) * TL;
  if (w.N != N) {
    w.N = N;
    w.vrec.assign(S * REC, (qs_real)0); w.it0.assign(S * NIT, 0.0); w.it1.assign(S * NIT, 0.0); w.st.assign(S * NIT, (qs_real)0); w.st2.assign(S * NS2, (qs_real)0);
    w.sb.assign(S * NSB, (qs_real)0); w.prod.assign(S * NPROD, (qs_real)0); w.res.assign(S * NRES, 0.0); w.stp.assign(S * NSTP, 0.0);
    w.pd.assign((size_t)NPD * TL, 0.0); w.pi32.assign((size_t)NPI * TL, 0); w.psm.assign(65 * TL, 0.0);
  }
  for (int k = 0; k <= N; ++k)
    for (int f = 0; f < REC; ++f) w.vrec[qs_blk(0, N, k, REC, 0) + (size_t)f * TL] = (qs_real)rec[(size_t)k * REC + f];
  QsBufs q{w.vrec.data(), {w.it0.data(), w.it1.data()}, w.st.data(), w.st2.data(), w.sb.data(), w.prod.data(), w.res.data(), w.stp.data(),
           w.pd.data(), w.pi32.data(), N, 0};
  const int32_t rr = r;
  double res5[5] = {0, 0, 0, 0, 0};
  int32_t st1 = 4, it1 = 0, qst1 = 0;
  // B = 1: lanes 1..31 of the tile are padding (qs_init marks them inactive), the outputs are indexed by problem 0
  HostBackend bk{*P, q, 1, 1, N, 0, x0, &rr, nullptr, xt, ut, &st1, &it1, &qst1, res5, std::move(w.psm)};
  bk.skip_pad = true;
  qs_drive(&bk, 1, g_depth);
  w.psm = std::move(bk.psm);
  *status = st1; *qp_iter = it1; *qp_status = qst1;
  for (int c = 0; c < 5; ++c) qp_res[c] = res5[c];
  return 0;
}
#endif
