// TEST HARNESS -- host compilation of the engine's kernel sources (safe_mpc_b200/csrc/*.cuh) so that the
// kernel logic can be checked against the oracle on a machine without a GPU.  Not part of the product: nothing in
// safe_mpc_b200/ loads this library, and it is not a fallback -- the product path fails without the CUDA library.
#include <cstring>
#include <vector>

#include "../../safe_mpc_b200/csrc/dev_model.cuh"
#ifdef EMU_QP
#include "../../safe_mpc_b200/csrc/qp_scalar.cuh"
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
// host stand-in of the device accessor: plain (stride-1) arrays of one problem
struct AccHost {
  const double* recb;
  double* stb;
  double* l0b;
  double rec(int k, int f) const { return recb[(size_t)k * REC + f]; }
  double ld(int k, int f) const { return stb[(size_t)k * QS_ST + f]; }
  void sd(int k, int f, double v) { stb[(size_t)k * QS_ST + f] = v; }
  double ll0(int i) const { return l0b[i]; }
  void sl0(int i, double v) { l0b[i] = v; }
};
}  // namespace

// z: [N+1][15], pi: [N][10], lam/t: [N+1][44]
extern "C" int emu_qp_solve(const smpc_problem_t* P, const double* rec, const double* x0, int r, double* z, double* pi,
                            double* lam, double* t, int* iter, int* status, double* res5) {
  const int N = P->N;
  std::vector<double> st(qs_doubles_per_problem(N), 0.0);
  AccHost acc{rec, st.data(), st.data() + (size_t)(N + 1) * QS_ST};
  QpScalar<AccHost> solver(*P, acc, x0, r);
  const QpResult R = solver.solve();
  for (int k = 0; k <= N; ++k) {
    const double* b = st.data() + (size_t)k * QS_ST;
    std::memcpy(z + k * 15, b + F_Z, 15 * sizeof(double));
    if (k < N) std::memcpy(pi + k * 10, b + F_PI, 10 * sizeof(double));
    std::memcpy(lam + k * 44, b + F_LAM, 44 * sizeof(double));
    std::memcpy(t + k * 44, b + F_T, 44 * sizeof(double));
  }
  *iter = R.iter; *status = R.status;
  for (int i = 0; i < 4; ++i) res5[i] = R.res[i];
  res5[4] = R.mu;
  return 0;
}
#endif
