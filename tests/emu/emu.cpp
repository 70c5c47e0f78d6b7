// TEST HARNESS -- host compilation of the engine's kernel sources (safe_mpc_b200/csrc/*.cuh) so that the
// kernel logic can be checked against the oracle on a machine without a GPU.  Not part of the product: nothing in
// safe_mpc_b200/ loads this library, and it is not a fallback -- the product path fails without the CUDA library.
#include <cstring>
#include <vector>

#include "../../safe_mpc_b200/csrc/dev_model.cuh"
#ifdef EMU_QP
#include "../../safe_mpc_b200/csrc/qp_lanes.cuh"
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
#include <barrier>
#include <thread>

namespace {
struct Group {
  std::barrier<> bar{QL};
  double x[QL];
  double scratch[QP_SCRATCH];
};
struct LanesHost {
  int lane_;
  Group* g;
  int lane() const { return lane_; }
  double shfl(double v, int src) { g->x[lane_] = v; g->bar.arrive_and_wait(); double r = g->x[src & 15]; g->bar.arrive_and_wait(); return r; }
  double shfl_xor(double v, int o) { return shfl(v, lane_ ^ o); }
  void sync() { g->bar.arrive_and_wait(); }
  double* scratch() { return g->scratch; }
};
}  // namespace

extern "C" int emu_qp_solve(const smpc_problem_t* P, const double* rec, const double* x0, int r, double* z16, double* pi16,
                            double* lam64, double* t64, int* iter, int* status, double* res5) {
  const int N = P->N;
  std::vector<double> buf((size_t)(N + 1) * qp_doubles_per_stage() + qp_doubles_fixed(), 0.0);
  QpMem M;
  double* p = buf.data();
  auto take = [&](size_t per_stage) { double* q = p; p += per_stage * (N + 1); return q; };
  M.rec = rec; M.x0 = x0; M.r = r;
  M.z = z16; M.pi = pi16; M.lam = lam64; M.t = t64;
  take(16); take(16); take(64); take(64);   // (caller-provided above; keep the arithmetic of the size formula)
  M.aux = take(16); M.fac = take(80); M.Pm = take(160); M.pv = take(16); M.wv = take(16); M.rb = take(16); M.gb = take(16);
  M.prod = take(64); M.dz = take(16); M.dpi = take(16); M.dlam = take(64); M.dtt = take(64);
  M.L0 = p;
  Group g;
  QpResult R[QL];
  std::vector<std::thread> th;
  for (int l = 0; l < QL; ++l)
    th.emplace_back([&, l]() {
      LanesHost ln{l, &g};
      QpSolver<LanesHost> s(ln, *P, M);
      R[l] = s.solve();
    });
  for (auto& t : th) t.join();
  *iter = R[0].iter; *status = R[0].status;
  for (int i = 0; i < 4; ++i) res5[i] = R[0].res[i];
  res5[4] = R[0].mu;
  return 0;
}
#endif
