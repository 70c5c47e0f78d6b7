// QP kernel: one RTI quadratic program per half warp (see qp_lanes.cuh for the algorithm and data layout).
// Replaces the HPIPM solve inside AcadosOcpSolver.solve() (reference controller.py:158) and the full-step update /
// status mapping that acados' SQP_RTI performs around it (controller.py:161-167).
#include "engine.cuh"

namespace smpc {

struct LanesDev {
  int lane_;
  int base_;
  unsigned mask_;
  double* scr_;
  __device__ __forceinline__ int lane() const { return lane_; }
  __device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(mask_, v, base_ + (src & 15)); }
  __device__ __forceinline__ double shfl_xor(double v, int o) { return __shfl_xor_sync(mask_, v, o); }
  __device__ __forceinline__ void sync() { __syncwarp(mask_); }
  __device__ __forceinline__ double* scratch() { return scr_; }
};

size_t qp_stride_doubles(int N) { return (size_t)(N + 1) * qp_doubles_per_stage() + qp_doubles_fixed(); }

__device__ __forceinline__ QpMem qp_views(double* base, int N) {
  QpMem M;
  double* p = base;
  const size_t n1 = (size_t)(N + 1);
  M.z = p; p += 16 * n1;
  M.pi = p; p += 16 * n1;
  M.lam = p; p += 64 * n1;
  M.t = p; p += 64 * n1;
  M.aux = p; p += 16 * n1;
  M.fac = p; p += 80 * n1;
  M.Pm = p; p += 160 * n1;
  M.pv = p; p += 16 * n1;
  M.wv = p; p += 16 * n1;
  M.rb = p; p += 16 * n1;
  M.gb = p; p += 16 * n1;
  M.prod = p; p += 64 * n1;
  M.dz = p; p += 16 * n1;
  M.dpi = p; p += 16 * n1;
  M.dlam = p; p += 64 * n1;
  M.dtt = p; p += 64 * n1;
  M.L0 = p;
  return M;
}

constexpr int QP_THREADS = 128;                 // 4 warps = 8 problems per CTA
constexpr int QP_PROBLEMS_PER_CTA = QP_THREADS / QL;

__global__ void __launch_bounds__(QP_THREADS)
qp_kernel(const smpc_problem_t* __restrict__ dP, int B, int N, const double* __restrict__ lin, const double* __restrict__ x0,
          const int32_t* __restrict__ r, const uint8_t* __restrict__ act, double* qpbuf, size_t stride, double* xt, double* ut,
          int32_t* status, int32_t* qp_iter, int32_t* qp_status, double* qp_res) {
  extern __shared__ double smem[];
  const int g = threadIdx.x / QL;                       // group within the CTA
  const int b = blockIdx.x * QP_PROBLEMS_PER_CTA + g;   // problem
  if (b >= B) return;
  if (act && !act[b]) return;
  const smpc_problem_t& P = *dP;
  LanesDev ln;
  ln.lane_ = threadIdx.x & 15;
  ln.base_ = (threadIdx.x & 31) & 16;
  ln.mask_ = 0xffffu << ln.base_;
  ln.scr_ = smem + (size_t)g * QP_SCRATCH;
  QpMem M = qp_views(qpbuf + (size_t)b * stride, N);
  M.rec = lin + (size_t)b * (N + 1) * REC;
  M.x0 = x0 + (size_t)b * NX;
  M.r = r[b];
  QpSolver<LanesDev> solver(ln, P, M);
  const QpResult R = solver.solve();
  // ---- full step and status mapping (acados SQP_RTI: QP success / max-iter -> step taken, else QP failure) ----
  const int lane = ln.lane_;
  const bool ok = (R.status == 0 || R.status == 1);
  bool nan = false;
  double* xtb = xt + (size_t)b * (N + 1) * NX;
  double* utb = ut + (size_t)b * N * NU;
  for (int k = 0; k <= N; ++k) {
    const double* rec = M.rec + (size_t)k * REC;
    const double z = ok ? M.z[(size_t)k * 16 + lane] : 0.0;
    if (lane < 5) { if (k < N) { utb[k * NU + lane] = rec[SMPC_REC_U + lane] + z; nan |= (z != z); } }
    else if (lane < 15) { xtb[k * NX + lane - 5] = rec[SMPC_REC_X + lane - 5] + z; nan |= (z != z); }
  }
  const unsigned any_nan = __ballot_sync(ln.mask_, nan) & ln.mask_;
  if (lane == 0) {
    status[b] = ok ? (any_nan ? 1 : 0) : 4;
    qp_iter[b] = R.iter;
    qp_status[b] = R.status;
    for (int i = 0; i < 4; ++i) qp_res[(size_t)b * 5 + i] = R.res[i];
    qp_res[(size_t)b * 5 + 4] = R.mu;
  }
}

void launch_qp(const LaunchCtx& c, const smpc_problem_t* dP, int B, int N, const double* lin, const double* x0, const int32_t* r,
               const uint8_t* act, double* qpbuf, double* xt, double* ut, int32_t* status, int32_t* qp_iter, int32_t* qp_status,
               double* qp_res) {
  const int grid = (B + QP_PROBLEMS_PER_CTA - 1) / QP_PROBLEMS_PER_CTA;
  const size_t smem = (size_t)QP_PROBLEMS_PER_CTA * QP_SCRATCH * sizeof(double);
  qp_kernel<<<grid, QP_THREADS, smem, c.stream>>>(dP, B, N, lin, x0, r, act, qpbuf, qp_stride_doubles(N), xt, ut, status, qp_iter,
                                                  qp_status, qp_res);
  ++*c.launches;
}

// canonical dump of the QP solution for parity tests (layout of smpc_get_qp)
__global__ void dump_qp_kernel(int B, int N, const double* qpbuf, size_t stride, const double* lin, double* dz, double* pi, double* lam, double* t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (N + 1)) return;
  const int b = idx / (N + 1), k = idx % (N + 1);
  QpMem M = qp_views(const_cast<double*>(qpbuf) + (size_t)b * stride, N);
  const double* rec = lin + ((size_t)b * (N + 1) + k) * REC;
  if (dz) {
    double* o = dz + (size_t)idx * 15;
    if (k < N) for (int i = 0; i < 15; ++i) o[i] = M.z[(size_t)k * 16 + i];
    else { for (int i = 0; i < 10; ++i) o[i] = M.z[(size_t)k * 16 + 5 + i]; for (int i = 10; i < 15; ++i) o[i] = 0.0; }
  }
  if (pi && k < N) for (int i = 0; i < 10; ++i) pi[((size_t)b * N + k) * 10 + i] = M.pi[(size_t)k * 16 + 5 + i];
  if (lam && t) {
    double* ol = lam + (size_t)idx * SMPC_QP_NC;
    double* ot = t + (size_t)idx * SMPC_QP_NC;
    for (int i = 0; i < SMPC_QP_NC; ++i) { ol[i] = 0.0; ot[i] = 0.0; }
    const bool ntau = rec[SMPC_REC_NTAU] > 0.5, ndist = rec[SMPC_REC_NDIST] > 0.5, nn = rec[SMPC_REC_NNROW] > 0.5;
    for (int lane = 0; lane < 15; ++lane) {
      const bool pa = lane < 5 ? ntau : true;
      const int ida = lane < 5 ? 10 + lane : lane - 5;
      if (pa) for (int s = 0; s < 2; ++s) {
        ol[s * SMPC_QP_NR + ida] = M.lam[(size_t)k * 64 + s * 16 + lane];
        ot[s * SMPC_QP_NR + ida] = M.t[(size_t)k * 64 + s * 16 + lane];
      }
      const bool pb = (lane >= 5 && lane <= 10) ? ndist : (lane == 11 ? nn : false);
      const int idb = lane == 11 ? 21 : 15 + lane - 5;
      if (pb) for (int s = 0; s < 2; ++s) {
        ol[s * SMPC_QP_NR + idb] = M.lam[(size_t)k * 64 + (2 + s) * 16 + lane];
        ot[s * SMPC_QP_NR + idb] = M.t[(size_t)k * 64 + (2 + s) * 16 + lane];
      }
    }
    if (nn && rec[SMPC_REC_SOFT] >= 0.0) {
      const double* a = M.aux + (size_t)k * 16;
      ol[2 * SMPC_QP_NR] = a[2]; ol[2 * SMPC_QP_NR + 1] = a[3];
      ot[2 * SMPC_QP_NR] = a[4]; ot[2 * SMPC_QP_NR + 1] = a[5];
    }
  }
}

void launch_dump_qp(const LaunchCtx& c, int B, int N, const double* qpbuf, size_t stride, const double* lin, double* dz, double* pi, double* lam, double* t) {
  const int n = B * (N + 1);
  dump_qp_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(B, N, qpbuf, stride, lin, dz, pi, lam, t);
  ++*c.launches;
}

}  // namespace smpc
