// QP kernel: one RTI quadratic program per thread, 32 problems per warp (see qp_scalar.cuh for the algorithm).
// Replaces the HPIPM solve inside AcadosOcpSolver.solve() (reference controller.py:158) and the full-step update /
// status mapping that acados' SQP_RTI performs around it (controller.py:161-167).
//
// Global layout (both the stage records written by the linearisation kernel and the solver state):
//   [tile = problem / 32][stage][field][problem % 32]  -> a warp's access to one field is one 256-byte row.
#include "engine.cuh"

namespace smpc {

struct AccDev {
  const double* recb;   // stage records of this problem's tile, offset by the lane
  double* stb;          // solver state of this problem's tile, offset by the lane
  double* l0b;
  __device__ __forceinline__ double rec(int k, int f) const { return __ldg(recb + ((size_t)k * REC + f) * 32); }
  __device__ __forceinline__ double ld(int k, int f) const { return stb[((size_t)k * QS_ST + f) * 32]; }
  __device__ __forceinline__ void sd(int k, int f, double v) { stb[((size_t)k * QS_ST + f) * 32] = v; }
  __device__ __forceinline__ double ll0(int i) const { return l0b[(size_t)i * 32]; }
  __device__ __forceinline__ void sl0(int i, double v) { l0b[(size_t)i * 32] = v; }
};

size_t qp_stride_doubles(int N) { return qs_doubles_per_problem(N); }

constexpr int QP_THREADS = 32;

__global__ void __launch_bounds__(QP_THREADS)
qp_kernel(const smpc_problem_t* __restrict__ dP, int B, int N, const double* __restrict__ lin, const double* __restrict__ x0,
          const int32_t* __restrict__ r, const uint8_t* __restrict__ act, double* qpbuf, double* xt, double* ut,
          int32_t* status, int32_t* qp_iter, int32_t* qp_status, double* qp_res) {
  const int b = blockIdx.x * QP_THREADS + threadIdx.x;
  if (b >= B) return;
  if (act && !act[b]) return;
  const smpc_problem_t& P = *dP;
  const int tile = b >> 5, lane = b & 31;
  AccDev acc;
  acc.recb = lin + (size_t)tile * (N + 1) * REC * 32 + lane;
  acc.stb = qpbuf + (size_t)tile * qs_doubles_per_problem(N) * 32 + lane;
  acc.l0b = acc.stb + (size_t)(N + 1) * QS_ST * 32;
  double x0l[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x0l[i] = x0[(size_t)b * NX + i];
  QpScalar<AccDev> solver(P, acc, x0l, r[b]);
  const QpResult R = solver.solve();
  // ---- full step and status mapping (acados SQP_RTI: QP success / max-iter -> step taken, else QP failure) ----
  const bool ok = (R.status == 0 || R.status == 1);
  bool nan = false;
  double* xtb = xt + (size_t)b * (N + 1) * NX;
  double* utb = ut + (size_t)b * N * NU;
  for (int k = 0; k <= N; ++k) {
    if (k < N)
      for (int i = 0; i < NU; ++i) { const double z = ok ? acc.ld(k, F_Z + i) : 0.0; utb[k * NU + i] = acc.rec(k, SMPC_REC_U + i) + z; nan |= (z != z); }
    for (int i = 0; i < NX; ++i) { const double z = ok ? acc.ld(k, F_Z + 5 + i) : 0.0; xtb[k * NX + i] = acc.rec(k, SMPC_REC_X + i) + z; nan |= (z != z); }
  }
  status[b] = ok ? (nan ? 1 : 0) : 4;
  qp_iter[b] = R.iter;
  qp_status[b] = R.status;
  for (int i = 0; i < 4; ++i) qp_res[(size_t)b * 5 + i] = R.res[i];
  qp_res[(size_t)b * 5 + 4] = R.mu;
}

void launch_qp(const LaunchCtx& c, const smpc_problem_t* dP, int B, int N, const double* lin, const double* x0, const int32_t* r,
               const uint8_t* act, double* qpbuf, double* xt, double* ut, int32_t* status, int32_t* qp_iter, int32_t* qp_status,
               double* qp_res) {
  const int grid = (B + QP_THREADS - 1) / QP_THREADS;
  qp_kernel<<<grid, QP_THREADS, 0, c.stream>>>(dP, B, N, lin, x0, r, act, qpbuf, xt, ut, status, qp_iter, qp_status, qp_res);
  ++*c.launches;
}

// stage records [tile][stage][field][32] -> caller layout [B][N+1][REC]   (smpc_get_lin)
__global__ void dump_lin_kernel(int B, int N, const double* lin, double* out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * (N + 1) * REC;
  if (idx >= total) return;
  const int f = idx % REC;
  const int k = (idx / REC) % (N + 1);
  const int b = idx / ((size_t)REC * (N + 1));
  out[idx] = lin[(((size_t)(b >> 5) * (N + 1) + k) * REC + f) * 32 + (b & 31)];
}
void launch_dump_lin(const LaunchCtx& c, int B, int N, const double* lin, double* out) {
  const size_t total = (size_t)B * (N + 1) * REC;
  dump_lin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(B, N, lin, out);
  ++*c.launches;
}

// canonical dump of the QP solution for parity tests (layout of smpc_get_qp)
__global__ void dump_qp_kernel(int B, int N, const double* qpbuf, const double* lin, double* dz, double* pi, double* lam, double* t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (N + 1)) return;
  const int b = idx / (N + 1), k = idx % (N + 1);
  const double* st = qpbuf + (size_t)(b >> 5) * qs_doubles_per_problem(N) * 32 + (b & 31) + (size_t)k * QS_ST * 32;
  const double* rec = lin + ((size_t)(b >> 5) * (N + 1) + k) * REC * 32 + (b & 31);
  auto S = [&](int f) { return st[(size_t)f * 32]; };
  auto R = [&](int f) { return rec[(size_t)f * 32]; };
  if (dz) {
    double* o = dz + (size_t)idx * 15;
    if (k < N) for (int i = 0; i < 15; ++i) o[i] = S(F_Z + i);
    else { for (int i = 0; i < 10; ++i) o[i] = S(F_Z + 5 + i); for (int i = 10; i < 15; ++i) o[i] = 0.0; }
  }
  if (pi && k < N) for (int i = 0; i < 10; ++i) pi[((size_t)b * N + k) * 10 + i] = S(F_PI + i);
  if (lam && t) {
    double* ol = lam + (size_t)idx * SMPC_QP_NC;
    double* ot = t + (size_t)idx * SMPC_QP_NC;
    const bool ntau = R(SMPC_REC_NTAU) > 0.5, ndist = R(SMPC_REC_NDIST) > 0.5, nn = R(SMPC_REC_NNROW) > 0.5;
    for (int j = 0; j < QNR; ++j) {
      const bool p = j < 10 ? true : (j < 15 ? ntau : (j < 21 ? ndist : nn));
      for (int s = 0; s < 2; ++s) { ol[s * QNR + j] = p ? S(F_LAM + s * QNR + j) : 0.0; ot[s * QNR + j] = p ? S(F_T + s * QNR + j) : 0.0; }
    }
    const bool soft = nn && R(SMPC_REC_SOFT) >= 0.0;
    ol[2 * QNR] = soft ? S(F_SLK + 2) : 0.0; ol[2 * QNR + 1] = soft ? S(F_SLK + 3) : 0.0;
    ot[2 * QNR] = soft ? S(F_SLK + 4) : 0.0; ot[2 * QNR + 1] = soft ? S(F_SLK + 5) : 0.0;
  }
}

void launch_dump_qp(const LaunchCtx& c, int B, int N, const double* qpbuf, size_t stride, const double* lin, double* dz, double* pi, double* lam, double* t) {
  const int n = B * (N + 1);
  dump_qp_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(B, N, qpbuf, lin, dz, pi, lam, t);
  ++*c.launches;
}

}  // namespace smpc
