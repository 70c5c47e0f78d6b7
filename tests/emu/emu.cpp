// TEST HARNESS -- host compilation of the engine's kernel sources (safe_mpc_b200/csrc/*.cuh) so that the
// kernel logic can be checked against the oracle on a machine without a GPU.  Not part of the product: nothing in
// safe_mpc_b200/ loads this library, and it is not a fallback -- the product path fails without the CUDA library.
#include <cstring>
#include <vector>

#include "../../safe_mpc_b200/csrc/dev_model.cuh"
#ifdef EMU_QP
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <deque>

#include "../../safe_mpc_b200/csrc/qp_warp.cuh"
#endif

using namespace smpc;

extern "C" {

int emu_linearize(const smpc_problem_t* P, int n, const int* k, const double* x, const double* u, const double* xnext,
                  const int* gate, const double* nn11, double* rec) {
  for (int i = 0; i < n; ++i)
    linearize_stage(*P, k[i], x + i * NX, u + i * NU, xnext + i * NX, stage_has_nn(*P, k[i]), gate[i] != 0, nn11 + i * 11, rec + (size_t)i * REC);
  return 0;
}

int emu_plant_step(const smpc_problem_t* P, int n, const double* inertial, const double* noise, const double* x, const double* u,
                   double* xn, double* a) {
  for (int i = 0; i < n; ++i)
    plant_step(*P, reinterpret_cast<const double(*)[10]>(inertial + (size_t)i * NQ * 10), noise + i * NU, x + i * NX, u + i * NU, xn + i * NX, a + i * NU);
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
// Host stand-in of a warp: 32 cooperative fibers (ucontext) and a barrier.  Every warp primitive of the device policy
// (shuffles, __syncwarp, staged bulk copies) is emulated with exchanges through a shared array between barriers.
// Copy timing is adversarial on purpose: with lazy = 1 a staged load lands only when it is waited for and a staged
// store leaves only when it is retired, with lazy = 0 loads land at issue -- so both premature reads of a staging
// buffer and premature reuse of a buffer show up as wrong numbers in the parity test.
// ----------------------------------------------------------------------------------------------------------------
struct FiberWarp;
struct Shared {
  ucontext_t main_ctx, ctx[32];
  std::vector<char> stacks;
  int cur = 0, arrived = 0, done = 0;
  unsigned gen = 0;
  double xd[32];
  int xi[32];
  std::vector<double> smem;
  struct Copy { double* dst; const double* src; int n; };
  std::vector<Copy> pend_load[2];
  std::deque<Copy> pend_store;
  int lazy = 0;
  bool finished[32];
  void (*body)(FiberWarp&) = nullptr;
  void* user = nullptr;
};

struct FiberWarp {
  Shared* sh;
  int ln;
  int lane() const { return ln; }
  double* scratch() { return sh->smem.data() + 2 * QW_IN + 2 * QW_OUT; }
  double* inbuf(int b) { return sh->smem.data() + b * QW_IN; }
  double* outbuf(int b) { return sh->smem.data() + 2 * QW_IN + b * QW_OUT; }
  void yield() {
    int nxt = sh->cur;
    for (int n = 0; n < 32; ++n) { nxt = (nxt + 1) & 31; if (!sh->finished[nxt]) break; }
    if (nxt == sh->cur) return;
    const int me = sh->cur;
    sh->cur = nxt;
    swapcontext(&sh->ctx[me], &sh->ctx[nxt]);
  }
  void sync() {
    const unsigned g = sh->gen;
    if (++sh->arrived == 32 - sh->done) { sh->arrived = 0; ++sh->gen; return; }
    long spins = 0;
    while (sh->gen == g) { yield(); if (++spins > 100000000L) { fprintf(stderr, "emu: barrier deadlock (divergent barrier)\n"); abort(); } }
  }
  double shfl(double v, int src) { sh->xd[ln] = v; sync(); const double r = sh->xd[src & 31]; sync(); return r; }
  double shfl_xor(double v, int mask) { return shfl(v, ln ^ mask); }
  int shfl_xor_i(int v, int mask) { sh->xi[ln] = v; sync(); const int r = sh->xi[(ln ^ mask) & 31]; sync(); return r; }
  void load_begin(int buf, int bytes) { (void)buf; (void)bytes; }
  void load(int buf, double* dst, const double* src, int n) {
    if (ln != 0) return;
    if (sh->lazy) sh->pend_load[buf].push_back({dst, src, n});
    else std::memcpy(dst, src, sizeof(double) * n);
  }
  void load_wait(int buf) {
    sync();
    if (ln == 0) { for (auto& c : sh->pend_load[buf]) std::memcpy(c.dst, c.src, sizeof(double) * c.n); sh->pend_load[buf].clear(); }
    sync();
  }
  void store(double* gdst, const double* ssrc, int n) {
    sync();
    if (ln == 0) {
      if (sh->lazy) sh->pend_store.push_back({gdst, ssrc, n});
      else std::memcpy(gdst, ssrc, sizeof(double) * n);
    }
  }
  void store_wait(int keep) {
    if (ln == 0)
      while ((int)sh->pend_store.size() > keep) { auto c = sh->pend_store.front(); sh->pend_store.pop_front(); std::memcpy(c.dst, c.src, sizeof(double) * c.n); }
    sync();
  }
};

Shared* g_sh = nullptr;
void fiber_entry(int ln) {
  Shared* sh = g_sh;
  FiberWarp w{sh, ln};
  sh->body(w);
  sh->finished[ln] = true;
  ++sh->done;
  // a finished lane must not be waited for any more
  if (sh->arrived == 32 - sh->done && sh->done < 32) { sh->arrived = 0; ++sh->gen; }
  if (sh->done == 32) { swapcontext(&sh->ctx[ln], &sh->main_ctx); return; }
  int nxt = ln;
  for (int n = 0; n < 32; ++n) { nxt = (nxt + 1) & 31; if (!sh->finished[nxt]) break; }
  sh->cur = nxt;
  swapcontext(&sh->ctx[ln], &sh->ctx[nxt]);
}

void run_warp(Shared& sh) {
  const size_t stk = 1 << 20;
  sh.stacks.assign(32 * stk, 0);
  g_sh = &sh;
  for (int l = 0; l < 32; ++l) {
    sh.finished[l] = false;
    getcontext(&sh.ctx[l]);
    sh.ctx[l].uc_stack.ss_sp = sh.stacks.data() + l * stk;
    sh.ctx[l].uc_stack.ss_size = stk;
    sh.ctx[l].uc_link = &sh.main_ctx;
    makecontext(&sh.ctx[l], (void (*)())fiber_entry, 1, l);
  }
  sh.cur = 0; sh.arrived = 0; sh.done = 0; sh.gen = 0;
  swapcontext(&sh.main_ctx, &sh.ctx[0]);
}

struct QpJob {
  const smpc_problem_t* P; const double* rec; const double* x0; int r; double* ws;
  double* xt; double* ut; QpResult R;
};

void qp_body(FiberWarp& w) {
  QpJob* J = (QpJob*)w.sh->user;
  QpWarp<FiberWarp> solver(w, *J->P, J->rec, J->ws, J->x0, J->r);
  const QpResult R = solver.solve();
  if (w.lane() == 0) J->R = R;
}
}  // namespace

// z: [N+1][15], pi: [N][10] (multiplier of the link k -> k+1), lam/t: [N+1][44]
extern "C" int emu_qp_solve(const smpc_problem_t* P, const double* rec, const double* x0, int r, int lazy, double* z, double* pi,
                            double* lam, double* t, int* iter, int* status, double* res5) {
  const int N = P->N;
  std::vector<double> ws(qw_ws_doubles(N), 0.0);
  Shared sh;
  sh.smem.assign(QW_SMEM_DOUBLES, 0.0);
  sh.lazy = lazy;
  QpJob job{P, rec, x0, r, ws.data(), nullptr, nullptr, {}};
  sh.user = &job;
  sh.body = qp_body;
  run_warp(sh);
  for (int k = 0; k <= N; ++k) {
    const double* b = ws.data() + (size_t)k * WS;
    std::memcpy(z + k * 15, b + A_Z, 15 * sizeof(double));
    if (k > 0) std::memcpy(pi + (k - 1) * 10, b + A_PIM, 10 * sizeof(double));
    std::memcpy(lam + k * 44, b + A_LAM, 44 * sizeof(double));
    std::memcpy(t + k * 44, b + A_T, 44 * sizeof(double));
  }
  *iter = job.R.iter; *status = job.R.status;
  for (int i = 0; i < 4; ++i) res5[i] = job.R.res[i];
  res5[4] = job.R.mu;
  return 0;
}
#endif
