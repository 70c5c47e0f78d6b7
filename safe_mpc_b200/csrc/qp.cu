// QP kernels: the phases of the split interior-point solver of qp_split.cuh, one small kernel each, and the host loop
// that sequences them.  Replaces the HPIPM solve inside AcadosOcpSolver.solve() (reference controller.py:158) and the
// full-step update / status mapping that acados' SQP_RTI performs around it (controller.py:161-167).
//
// Launch shapes (B problems, T = ceil(B/32) tiles, N+1 stages):
//   prep / step   one warp per (tile, stage): T (N+1) warps, 4 warps per CTA; lane = problem of the tile.  Streaming
//                 kernels: every global access of a warp is one contiguous 256-byte segment of the tile-interleaved arrays.
//   ric1 / ric2   one warp per tile (lane = problem), walking the stages; ric1 keeps P_{k+1}, p_{k+1} of its 32 problems
//                 in shared memory (33 KB per warp).
//   ctl / red     one thread per problem.
// The host reads two counters (problems still active, problems that asked for the centering re-solve) once per IPM
// iteration; everything else is asynchronous on the handle's stream.
#include "engine.cuh"

namespace smpc {

namespace {

constexpr int SP_WARPS = 4;      // warps per CTA of the stage-parallel kernels
#ifndef QS_RIC_NBUF
#define QS_RIC_NBUF 1            // staging buffers of the Riccati sweeps (1: more resident warps per SM, no fetch overlap)
#endif
#ifndef QS_PREP_MINB
#define QS_PREP_MINB 4
#endif
#ifndef QS_SP_MINB
#define QS_SP_MINB 3             // resident CTAs per SM the register allocation of prep / step is sized for
#endif

__global__ void __launch_bounds__(32) qs_init_kernel(QsBufs q, int B, const double* __restrict__ x0, const int32_t* __restrict__ r,
                                                      const uint8_t* __restrict__ act) {
  qs_init(q, blockIdx.x, threadIdx.x, B, x0, r, act);
}

constexpr size_t PREP_SMEM = sizeof(double) * 2 * PREP_SCRATCH * TL;
constexpr int PREP_WARPS = 2;    // warps per CTA of prep: 2 x 26.9 KB of lane-private Jacobian scratch, 4 CTAs per SM
__global__ void __launch_bounds__(32 * PREP_WARPS, QS_PREP_MINB) qs_prep_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q, int T, int kk) {
  extern __shared__ __align__(128) double jsm_all[];
  const int wi = threadIdx.x >> 5;
  const int w = blockIdx.x * PREP_WARPS + wi;
  if (w >= T * (q.N + 1)) return;
  qs_prep(*dP, q, w / (q.N + 1), threadIdx.x & 31, w % (q.N + 1), kk, jsm_all + (size_t)wi * PREP_SCRATCH * TL + (threadIdx.x & 31));
}

template <int MODE>
__global__ void __launch_bounds__(32 * SP_WARPS, QS_SP_MINB) qs_step_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q, int T, int kk) {
  const int w = blockIdx.x * SP_WARPS + (threadIdx.x >> 5);
  if (w >= T * (q.N + 1)) return;
  qs_step(*dP, q, w / (q.N + 1), threadIdx.x & 31, w % (q.N + 1), kk, MODE);
}

__global__ void __launch_bounds__(32) qs_ctl_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q, int kk, int32_t* status, int32_t* qp_iter,
                                                     int32_t* qp_status, double* qp_res, int* counters) {
  const bool on = qs_ctl(*dP, q, blockIdx.x, threadIdx.x, kk, status, qp_iter, qp_status, qp_res);
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if (threadIdx.x == 0 && m) atomicAdd(&counters[2 * kk], __popc(m));
}

__global__ void __launch_bounds__(32 * SP_WARPS) qs_final_kernel(QsBufs q, int T, int B, const uint8_t* __restrict__ act, int32_t* status,
                                                                 double* xt, double* ut) {
  const int w = blockIdx.x * SP_WARPS + (threadIdx.x >> 5);
  if (w >= T * (q.N + 1)) return;
  const int tile = w / (q.N + 1), lane = threadIdx.x & 31;
  if (qs_final(q, tile, lane, w % (q.N + 1), act, B, status, xt, ut)) atomicExch(&status[tile * TL + lane], 1);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Device warp policy of the Riccati sweeps: two staging buffers of `nfb` fields x 32 lanes in shared memory, filled by
// TMA 1-D bulk copies (cp.async.bulk global -> shared, mbarrier completion) that lane 0 issues one stage ahead.
struct TmaStage {
  double* sm;
  uint64_t* bar;
  uint32_t phase[2];
  int ln, nfb;
  __device__ __forceinline__ int lane() const { return ln; }
  __device__ __forceinline__ int nbuf() const { return QS_RIC_NBUF; }
  __device__ __forceinline__ bool any(bool v) const { return __any_sync(0xffffffffu, v) != 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ double* buf(int b) const { return sm + (size_t)b * nfb * TL + ln; }
  __device__ __forceinline__ void init() {
    phase[0] = phase[1] = 0;
    if (ln == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 0)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  __device__ __forceinline__ void fetch_begin(int b, int nfields) {
    if (ln == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + b)), "r"(nfields * TL * 8) : "memory");
  }
  __device__ __forceinline__ void fetch(int b, int dst_field, const double* gblock, int src_field, int nfields) {
    if (ln == 0)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(sm + ((size_t)b * nfb + dst_field) * TL)),
                   "l"(gblock + (size_t)src_field * TL), "r"(nfields * TL * 8), "r"(smem_u32(bar + b))
                   : "memory");
  }
  __device__ __forceinline__ void wait(int b) {
    const uint32_t addr = smem_u32(bar + b), par = phase[b];
    uint32_t ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(addr), "r"(par) : "memory");
    } while (!ok);
    phase[b] = par ^ 1u;
  }
  // this lane's global stores become visible to bulk copies issued after the next warp barrier
  __device__ __forceinline__ void publish() const { asm volatile("fence.proxy.async;" ::: "memory"); }
};

constexpr size_t RIC1_SMEM = sizeof(double) * (QS_RIC_NBUF * RIC1_STAGE_FIELDS + 65) * TL + 16;
constexpr size_t RIC2_SMEM = sizeof(double) * (QS_RIC_NBUF * RIC2_STAGE_FIELDS) * TL + 16;

__global__ void __launch_bounds__(32) qs_ric1_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q) {
  extern __shared__ __align__(128) double smem[];
  TmaStage w;
  w.sm = smem; w.nfb = RIC1_STAGE_FIELDS; w.ln = threadIdx.x;
  double* psm = smem + (size_t)QS_RIC_NBUF * RIC1_STAGE_FIELDS * TL;
  w.bar = reinterpret_cast<uint64_t*>(psm + 65 * TL);
  w.init();
  qs_ric1(*dP, q, blockIdx.x, w, psm + threadIdx.x);
}

template <int MODE>
__global__ void __launch_bounds__(32) qs_ric2_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q) {
  extern __shared__ __align__(128) double smem[];
  TmaStage w;
  w.sm = smem; w.nfb = RIC2_STAGE_FIELDS; w.ln = threadIdx.x;
  w.bar = reinterpret_cast<uint64_t*>(smem + (size_t)QS_RIC_NBUF * RIC2_STAGE_FIELDS * TL);
  w.init();
  qs_ric2(*dP, q, blockIdx.x, w, MODE);
}

__global__ void __launch_bounds__(32) qs_red_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q, int kk, int after_redo, int* counters) {
  qs_red(*dP, q, blockIdx.x, threadIdx.x, after_redo != 0);
  if (!after_redo) {
    const int32_t* pi = q.pi + qs_pb(blockIdx.x, NPI, threadIdx.x);
    const bool redo = QF(pi, J_ACT) && QF(pi, J_REDO);
    const unsigned m = __ballot_sync(0xffffffffu, redo);
    if (threadIdx.x == 0 && m) atomicAdd(&counters[2 * kk + 1], __popc(m));
  }
}

// stage records, tile-interleaved -> caller layout [B][N+1][REC] (smpc_get_lin)
__global__ void rec_untile_kernel(int B, int N, const double* __restrict__ rec, double* out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * (N + 1) * REC) return;
  const int f = idx % REC;
  const size_t bk = idx / REC;
  const int k = bk % (N + 1), b = bk / (N + 1);
  out[idx] = QF(rec + qs_blk(b / TL, N, k, REC, b % TL), f);
}

// canonical dump of the QP solution for parity tests (layout of smpc_get_qp)
__global__ void dump_qp_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q, int B, double* dz, double* pi, double* lam, double* t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = q.N;
  if (idx >= B * (N + 1)) return;
  const int b = idx / (N + 1), k = idx % (N + 1);
  const int tile = b / TL, lane = b % TL;
  const int buf = QF(q.pi + qs_pb(tile, NPI, lane), J_ITBUF);
  const double* it = q.it[buf] + qs_blk(tile, N, k, NIT, lane);
  const StageFlags F = qs_flags(*dP, k);
  if (dz) {
    double* o = dz + (size_t)idx * 15;
    if (k < N) for (int i = 0; i < 15; ++i) o[i] = QF(it, I_Z + i);
    else { for (int i = 0; i < 10; ++i) o[i] = QF(it, I_Z + 5 + i); for (int i = 10; i < 15; ++i) o[i] = 0.0; }
  }
  if (pi && k > 0) for (int i = 0; i < 10; ++i) pi[((size_t)b * N + k - 1) * 10 + i] = QF(it, I_PIM + i);
  if (lam && t) {
    double* ol = lam + (size_t)idx * SMPC_QP_NC;
    double* ot = t + (size_t)idx * SMPC_QP_NC;
    for (int j = 0; j < QNR; ++j) {
      const bool p = j < 10 ? true : (j < 15 ? F.tau : (j < 21 ? F.dist : F.nn));
      for (int s = 0; s < 2; ++s) { ol[s * QNR + j] = p ? QF(it, I_LAM + s * QNR + j) : 0.0; ot[s * QNR + j] = p ? QF(it, I_T + s * QNR + j) : 0.0; }
    }
    ol[2 * QNR] = F.soft ? QF(it, I_SLK + 2) : 0.0; ol[2 * QNR + 1] = F.soft ? QF(it, I_SLK + 3) : 0.0;
    ot[2 * QNR] = F.soft ? QF(it, I_SLK + 4) : 0.0; ot[2 * QNR + 1] = F.soft ? QF(it, I_SLK + 5) : 0.0;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ solver object
struct QpSolver {
  int B = 0, N = 0, T = 0, iter_max = 0;
  QsBufs q{};
  double* block = nullptr;      // one allocation for all double arrays
  int32_t* pi = nullptr;
  int* counters = nullptr;      // [2 (iter_max + 2)]: active / redo per IPM iteration
  int* h_counters = nullptr;    // pinned
  int last_iters = 0;
};

size_t qp_bytes(int B, int N) {
  const size_t T = (B + TL - 1) / TL, S = T * (N + 1) * TL;
  return sizeof(double) * (S * (REC + 3 * NIT + NSB + NPROD + NRES + NSTP) + T * NPD * TL);
}

QpSolver* qp_create(int B, int N, int iter_max, cudaStream_t stream, cudaError_t* err) {
  QpSolver* s = new QpSolver;
  s->B = B; s->N = N; s->T = (B + TL - 1) / TL; s->iter_max = iter_max;
  const size_t T = s->T, S = T * (N + 1) * TL;
  cudaError_t e = cudaMalloc((void**)&s->block, qp_bytes(B, N));
  if (e == cudaSuccess) e = cudaMemsetAsync(s->block, 0, qp_bytes(B, N), stream);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->pi, sizeof(int32_t) * T * NPI * TL);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->pi, 0, sizeof(int32_t) * T * NPI * TL, stream);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->counters, sizeof(int) * 2 * (iter_max + 2));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_counters, sizeof(int) * 2);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PREP_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_ric1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RIC1_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_ric2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RIC2_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_ric2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RIC2_SMEM);
  if (e != cudaSuccess) { *err = e; qp_destroy(s); return nullptr; }
  double* p = s->block;
  double* rec = p; p += S * REC;
  s->q.rec = rec;
  s->q.it[0] = p; p += S * NIT;
  s->q.it[1] = p; p += S * NIT;
  s->q.st = p; p += S * NIT;
  s->q.sb = p; p += S * NSB;
  s->q.prod = p; p += S * NPROD;
  s->q.res = p; p += S * NRES;
  s->q.stp = p; p += S * NSTP;
  s->q.pd = p;
  s->q.pi = s->pi;
  s->q.N = N;
  *err = cudaSuccess;
  return s;
}

void qp_destroy(QpSolver* s) {
  if (!s) return;
  if (s->block) cudaFree(s->block);
  if (s->pi) cudaFree(s->pi);
  if (s->counters) cudaFree(s->counters);
  if (s->h_counters) cudaFreeHost(s->h_counters);
  delete s;
}

double* qp_rec(QpSolver* s) { return const_cast<double*>(s->q.rec); }
int qp_last_iterations(const QpSolver* s) { return s->last_iters; }

namespace {
struct DeviceBackend {
  const LaunchCtx& c;
  const smpc_problem_t* dP;
  QpSolver* s;
  const double* x0; const int32_t* r; const uint8_t* act;
  double *xt, *ut; int32_t *status, *qp_iter, *qp_status; double* qp_res;
  int kk_last = 0;
  cudaError_t err = cudaSuccess;
  int sp_grid() const { return (s->T * (s->N + 1) + SP_WARPS - 1) / SP_WARPS; }
  void count(int n = 1) { *c.launches += n; }
  void init() {
    cudaMemsetAsync(s->counters, 0, sizeof(int) * 2 * (s->iter_max + 2), c.stream);
    qs_init_kernel<<<s->T, 32, 0, c.stream>>>(s->q, s->B, x0, r, act);
    count();
  }
  void prep(int kk) {
    qs_prep_kernel<<<(s->T * (s->N + 1) + PREP_WARPS - 1) / PREP_WARPS, 32 * PREP_WARPS, PREP_SMEM, c.stream>>>(dP, s->q, s->T, kk);
    count();
  }
  void ctl(int kk) {
    kk_last = kk;
    qs_ctl_kernel<<<s->T, 32, 0, c.stream>>>(dP, s->q, kk, status, qp_iter, qp_status, qp_res, s->counters);
    count();
  }
  void ric1() { qs_ric1_kernel<<<s->T, 32, RIC1_SMEM, c.stream>>>(dP, s->q); count(); }
  void ric2(int mode) {
    if (mode == 1) qs_ric2_kernel<1><<<s->T, 32, RIC2_SMEM, c.stream>>>(dP, s->q);
    else qs_ric2_kernel<2><<<s->T, 32, RIC2_SMEM, c.stream>>>(dP, s->q);
    count();
  }
  void step(int kk, int mode) {
    if (mode == 0) qs_step_kernel<0><<<sp_grid(), 32 * SP_WARPS, 0, c.stream>>>(dP, s->q, s->T, kk);
    else if (mode == 1) qs_step_kernel<1><<<sp_grid(), 32 * SP_WARPS, 0, c.stream>>>(dP, s->q, s->T, kk);
    else qs_step_kernel<2><<<sp_grid(), 32 * SP_WARPS, 0, c.stream>>>(dP, s->q, s->T, kk);
    count();
  }
  void final() { qs_final_kernel<<<sp_grid(), 32 * SP_WARPS, 0, c.stream>>>(s->q, s->T, s->B, act, status, xt, ut); count(); }
  void red(bool after) { qs_red_kernel<<<s->T, 32, 0, c.stream>>>(dP, s->q, kk_last, after ? 1 : 0, s->counters); count(); }
  void sync(int& na, int& nr) {
    cudaError_t e = cudaMemcpyAsync(s->h_counters, s->counters + 2 * kk_last, 2 * sizeof(int), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) { err = e; na = 0; nr = 0; return; }     // stop iterating; the caller reports the error
    na = s->h_counters[0]; nr = s->h_counters[1];
  }
};
}  // namespace

cudaError_t launch_qp_solve(const LaunchCtx& c, const smpc_problem_t* dP, QpSolver* s, const double* x0, const int32_t* r, const uint8_t* act,
                            double* xt, double* ut, int32_t* status, int32_t* qp_iter, int32_t* qp_status, double* qp_res) {
  DeviceBackend bk{c, dP, s, x0, r, act, xt, ut, status, qp_iter, qp_status, qp_res};
  s->last_iters = qs_drive(bk);
  return bk.err;
}

void launch_rec_untile(const LaunchCtx& c, QpSolver* s, double* out) {
  const size_t n = (size_t)s->B * (s->N + 1) * REC;
  rec_untile_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(s->B, s->N, s->q.rec, out);
  ++*c.launches;
}

void launch_dump_qp(const LaunchCtx& c, const smpc_problem_t* dP, QpSolver* s, double* dz, double* pi, double* lam, double* t) {
  const int n = s->B * (s->N + 1);
  dump_qp_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(dP, s->q, s->B, dz, pi, lam, t);
  ++*c.launches;
}

}  // namespace smpc
