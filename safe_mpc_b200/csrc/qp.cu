// QP kernel: one RTI quadratic program per warp (see qp_warp.cuh for the algorithm and the lane mapping).
// Replaces the HPIPM solve inside AcadosOcpSolver.solve() (reference controller.py:158) and the full-step update /
// status mapping that acados' SQP_RTI performs around it (controller.py:161-167).
//
// Launch shape: persistent one-warp CTAs, QW_SMEM bytes of shared memory each (two staging buffers, two output buffers,
// scratch) -> 12 resident warps per SM; every warp owns one workspace slot in global memory and pulls problems from an
// atomic queue until the batch is done, so problems with more IPM iterations do not hold up a whole wave.
// Data movement: TMA 1-D bulk copies (cp.async.bulk global->shared with mbarrier completion, shared->global bulk
// groups); the warp's lanes only ever touch shared memory and registers inside a sweep.
#include "engine.cuh"

namespace smpc {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Device warp policy of QpWarp.
struct WarpDev {
  double* sm;              // shared memory of this warp
  uint64_t* bar;           // two mbarriers (one per staging buffer)
  uint32_t phase[2];
  int ln;

  __device__ __forceinline__ int lane() const { return ln; }
  __device__ __forceinline__ double* inbuf(int b) { return sm + b * QW_IN; }
  __device__ __forceinline__ double* outbuf(int b) { return sm + 2 * QW_IN + b * QW_OUT; }
  __device__ __forceinline__ double* scratch() { return sm + 2 * QW_IN + 2 * QW_OUT; }
  __device__ __forceinline__ void sync() { __syncwarp(); }
  __device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  __device__ __forceinline__ double shfl_xor(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
  __device__ __forceinline__ int shfl_xor_i(int v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

  __device__ __forceinline__ void init() {
    phase[0] = phase[1] = 0;
    if (ln == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 0)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  // staged load: lane 0 arms the barrier with the byte count, then issues the bulk copies
  __device__ __forceinline__ void load_begin(int buf, int bytes) {
    if (ln == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + buf)), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void load(int buf, double* dst, const double* src, int n) {
    if (ln == 0)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                   "r"(n * 8), "r"(smem_u32(bar + buf))
                   : "memory");
  }
  __device__ __forceinline__ void load_wait(int buf) {
    const uint32_t addr = smem_u32(bar + buf), par = phase[buf];
    uint32_t ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(addr), "r"(par) : "memory");
    } while (!ok);
    phase[buf] = par ^ 1u;
  }
  // staged store: make the lanes' shared-memory writes visible to the async proxy, then one bulk copy
  __device__ __forceinline__ void store(double* gdst, const double* ssrc, int n) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (ln == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(n * 8) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  // keep == 1: the buffer written two stages ago may be reused (its source has been read);
  // keep == 0: every store of the sweep has completed (the next sweep reads them back)
  __device__ __forceinline__ void store_wait(int keep) {
    if (ln == 0) {
      if (keep) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  }
};

}  // namespace

size_t qp_ws_doubles(int N) { return qw_ws_doubles(N); }
size_t qp_smem_bytes() { return (size_t)QW_SMEM_DOUBLES * sizeof(double) + 16; }

#ifndef QP_MINB
#define QP_MINB 11        // resident one-warp CTAs per SM the register allocation is sized for (shared memory allows 11)
#endif
__global__ void __launch_bounds__(32, QP_MINB)
qp_kernel(const smpc_problem_t* __restrict__ dP, int B, int N, const double* __restrict__ lin, const double* __restrict__ x0,
          const int32_t* __restrict__ r, const uint8_t* __restrict__ act, double* ws, int* queue, double* xt, double* ut,
          int32_t* status, int32_t* qp_iter, int32_t* qp_status, double* qp_res) {
  extern __shared__ __align__(16) double smem[];
  WarpDev w;
  w.sm = smem;
  w.bar = reinterpret_cast<uint64_t*>(smem + QW_SMEM_DOUBLES);
  w.ln = threadIdx.x;
  w.init();
  const smpc_problem_t& P = *dP;
  double* myws = ws + (size_t)blockIdx.x * qw_ws_doubles(N);
  for (bool once = true;; once = false) {
    int b = blockIdx.x;
    if (queue) {
      if (w.ln == 0) b = atomicAdd(queue, 1);
      b = __shfl_sync(0xffffffffu, b, 0);
    } else if (!once) break;
    if (b >= B) break;
    if (act && !act[b]) continue;
    const double* rec = lin + (size_t)b * (N + 1) * REC;
    QpWarp<WarpDev> solver(w, P, rec, myws, x0 + (size_t)b * NX, r[b]);
    const QpResult R = solver.solve();
    // ---- full step and status mapping (acados SQP_RTI: QP success / max-iter -> step taken, else QP failure) ----
    const bool ok = (R.status == 0 || R.status == 1);
    bool nan = false;
    double* xtb = xt + (size_t)b * (N + 1) * NX;
    double* utb = ut + (size_t)b * N * NU;
    __syncwarp();
    for (int e = w.ln; e < (N + 1) * NZ; e += 32) {
      const int k = e / NZ, j = e - k * NZ;
      if (j < NU && k == N) continue;
      const double z = ok ? __ldcg(myws + (size_t)k * WS + A_Z + j) : 0.0;      // written by the async proxy: bypass L1
      nan |= (z != z);
      if (j < NU) utb[k * NU + j] = rec[(size_t)k * REC + SMPC_REC_U + j] + z;
      else xtb[k * NX + j - NU] = rec[(size_t)k * REC + SMPC_REC_X + j - NU] + z;
    }
    nan = __any_sync(0xffffffffu, nan);
    if (w.ln == 0) {
      status[b] = ok ? (nan ? 1 : 0) : 4;
      qp_iter[b] = R.iter;
      qp_status[b] = R.status;
      for (int q = 0; q < 4; ++q) qp_res[(size_t)b * 5 + q] = R.res[q];
      qp_res[(size_t)b * 5 + 4] = R.mu;
    }
    __syncwarp();
  }
}

// number of persistent one-warp CTAs (= workspace slots) that are co-resident on the current device
int qp_grid(int B) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qp_smem_bytes());
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_kernel, 32, qp_smem_bytes());
  const int g = sms * (per_sm > 0 ? per_sm : 1);
  return B < g ? B : g;
}

void launch_qp(const LaunchCtx& c, const smpc_problem_t* dP, int B, int N, int grid, const double* lin, const double* x0, const int32_t* r,
               const uint8_t* act, double* ws, int* queue, double* xt, double* ut, int32_t* status, int32_t* qp_iter, int32_t* qp_status,
               double* qp_res) {
  if (queue) cudaMemsetAsync(queue, 0, sizeof(int), c.stream);
  qp_kernel<<<grid, 32, qp_smem_bytes(), c.stream>>>(dP, B, N, lin, x0, r, act, ws, queue, xt, ut, status, qp_iter, qp_status, qp_res);
  ++*c.launches;
}

// stage records [B][N+1][REC] are already in the caller layout (smpc_get_lin is a plain copy)

// canonical dump of the QP solution for parity tests (layout of smpc_get_qp); slot s holds problem s (grid == B)
__global__ void dump_qp_kernel(int B, int N, const double* ws, const double* lin, double* dz, double* pi, double* lam, double* t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (N + 1)) return;
  const int b = idx / (N + 1), k = idx % (N + 1);
  const double* st = ws + (size_t)b * qw_ws_doubles(N) + (size_t)k * WS;
  const double* rec = lin + (size_t)idx * REC;
  if (dz) {
    double* o = dz + (size_t)idx * 15;
    if (k < N) for (int i = 0; i < 15; ++i) o[i] = st[A_Z + i];
    else { for (int i = 0; i < 10; ++i) o[i] = st[A_Z + 5 + i]; for (int i = 10; i < 15; ++i) o[i] = 0.0; }
  }
  if (pi && k > 0) for (int i = 0; i < 10; ++i) pi[((size_t)b * N + k - 1) * 10 + i] = st[A_PIM + i];
  if (lam && t) {
    double* ol = lam + (size_t)idx * SMPC_QP_NC;
    double* ot = t + (size_t)idx * SMPC_QP_NC;
    const bool ntau = rec[SMPC_REC_NTAU] > 0.5, ndist = rec[SMPC_REC_NDIST] > 0.5, nn = rec[SMPC_REC_NNROW] > 0.5;
    for (int j = 0; j < QNR; ++j) {
      const bool p = j < 10 ? true : (j < 15 ? ntau : (j < 21 ? ndist : nn));
      for (int s = 0; s < 2; ++s) { ol[s * QNR + j] = p ? st[A_LAM + s * QNR + j] : 0.0; ot[s * QNR + j] = p ? st[A_T + s * QNR + j] : 0.0; }
    }
    const bool soft = nn && rec[SMPC_REC_SOFT] >= 0.0;
    ol[2 * QNR] = soft ? st[A_SLK + 2] : 0.0; ol[2 * QNR + 1] = soft ? st[A_SLK + 3] : 0.0;
    ot[2 * QNR] = soft ? st[A_SLK + 4] : 0.0; ot[2 * QNR + 1] = soft ? st[A_SLK + 5] : 0.0;
  }
}

void launch_dump_qp(const LaunchCtx& c, int B, int N, const double* ws, const double* lin, double* dz, double* pi, double* lam, double* t) {
  const int n = B * (N + 1);
  dump_qp_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(B, N, ws, lin, dz, pi, lam, t);
  ++*c.launches;
}

}  // namespace smpc
