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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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

constexpr size_t PREP_SMEM = sizeof(double) * 2 * PREP_SCRATCH * TL + 512;   // a retiring prep CTA leaves room for one Riccati CTA
constexpr int PREP_WARPS = 2;    // warps per CTA of prep: 2 x 26.9 KB of lane-private Jacobian scratch, 4 CTAs per SM
template <bool FIRST>
__global__ void __launch_bounds__(32 * PREP_WARPS, QS_PREP_MINB) qs_prep_kernel(const smpc_problem_t* __restrict__ dP, QsBufs q, int T, int kk) {
  extern __shared__ __align__(128) double jsm_all[];
  const int wi = threadIdx.x >> 5;
  const int w = blockIdx.x * PREP_WARPS + wi;
  if (w >= T * (q.N + 1)) return;
  qs_prep<FIRST>(*dP, q, w / (q.N + 1), threadIdx.x & 31, w % (q.N + 1), kk, jsm_all + (size_t)wi * PREP_SCRATCH * TL + (threadIdx.x & 31));
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
  // hint: start moving a field range of a stage block towards L2
  __device__ __forceinline__ void prefetch(const double* gblock, int src_field, int nfields) const {
    if (ln == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gblock + (size_t)src_field * TL), "r"(nfields * TL * 8) : "memory");
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
#ifndef QS_GROUPS
#define QS_GROUPS 3              // tile groups solved concurrently on their own streams (see QsLoop)
#endif
constexpr int MAX_GROUPS = 8;

struct QpGroup {
  QsBufs q{};
  int T = 0;                    // tiles of this group
  cudaStream_t stream = nullptr;   // stage-parallel kernels (bandwidth-bound)
  cudaStream_t hi = nullptr;       // Riccati sweeps (latency-bound, few warps): higher priority, so that their CTAs take the
                                   // slots that retiring stage-parallel CTAs of the other groups free
  cudaEvent_t ev = nullptr;     // counters of the current iteration have landed in h_counters
  cudaEvent_t evx = nullptr;    // hand-over between the two streams
  int* counters = nullptr;      // [2 (iter_max + 2)]: active / redo per IPM iteration
  int* h_counters = nullptr;    // pinned, 2 ints
};

struct QpSolver {
  int B = 0, N = 0, T = 0, iter_max = 0, G = 1;
  QsBufs q{};                   // whole batch (group 0 .. G-1 are sub-ranges of its tiles)
  QpGroup grp[MAX_GROUPS];
  double* block = nullptr;      // one allocation for all double arrays
  int32_t* pi = nullptr;
  int* counters = nullptr;
  int* h_counters = nullptr;
  cudaEvent_t ev_in = nullptr;  // inputs (records, x0) are ready on the caller's stream
  int last_iters = 0;
  bool profile = false;         // record one event pair per kernel of the next solves (smpc_set_profiling)
  double prof_ms[SMPC_PROF_N] = {0};
  int32_t prof_n[SMPC_PROF_N] = {0};
  double prof_span_ms = 0.0;
};

size_t qp_bytes(int B, int N) {
  const size_t T = (B + TL - 1) / TL, S = T * (N + 1) * TL;
  return sizeof(double) * (S * (REC + 3 * NIT + NSB + NPROD + NRES + NSTP) + T * NPD * TL);
}

QpSolver* qp_create(int B, int N, int iter_max, cudaStream_t stream, cudaError_t* err) {
  QpSolver* s = new QpSolver;
  s->B = B; s->N = N; s->T = (B + TL - 1) / TL; s->iter_max = iter_max;
  const size_t T = s->T, S = T * (N + 1) * TL;
  // groups: at least ~64 tiles each so that the stage-parallel kernels of one group still fill the GPU
  int G = QS_GROUPS;
  if (const char* e = getenv("SMPC_QP_GROUPS")) G = atoi(e);
  while (G > 1 && s->T / G < 64) --G;
  if (G < 1) G = 1;
  if (G > MAX_GROUPS) G = MAX_GROUPS;
  s->G = G;
  const size_t ncnt = (size_t)2 * (iter_max + 2);
  cudaError_t e = cudaMalloc((void**)&s->block, qp_bytes(B, N));
  if (e == cudaSuccess) e = cudaMemsetAsync(s->block, 0, qp_bytes(B, N), stream);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->pi, sizeof(int32_t) * T * NPI * TL);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->pi, 0, sizeof(int32_t) * T * NPI * TL, stream);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->counters, sizeof(int) * ncnt * G);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_counters, sizeof(int) * 2 * G);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_prep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PREP_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_prep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PREP_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_ric1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RIC1_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_ric2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RIC2_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(qs_ric2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RIC2_SMEM);
  for (int g = 0; g < G && e == cudaSuccess; ++g) {
    int plo = 0, phi = 0;
    cudaDeviceGetStreamPriorityRange(&plo, &phi);
    e = cudaStreamCreateWithPriority(&s->grp[g].stream, cudaStreamNonBlocking, plo);
    if (e == cudaSuccess && G > 1 && !getenv("SMPC_QP_NOPRIO")) e = cudaStreamCreateWithPriority(&s->grp[g].hi, cudaStreamNonBlocking, phi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->grp[g].ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->grp[g].evx, cudaEventDisableTiming);
  }
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
  s->q.tile0 = 0;
  for (int g = 0; g < G; ++g) {
    QpGroup& gr = s->grp[g];
    const int t0 = (int)((long long)s->T * g / G), t1 = (int)((long long)s->T * (g + 1) / G);
    const size_t so = (size_t)t0 * (N + 1) * TL;
    gr.T = t1 - t0;
    gr.q = s->q;
    gr.q.rec = s->q.rec + so * REC; gr.q.it[0] = s->q.it[0] + so * NIT; gr.q.it[1] = s->q.it[1] + so * NIT; gr.q.st = s->q.st + so * NIT;
    gr.q.sb = s->q.sb + so * NSB; gr.q.prod = s->q.prod + so * NPROD; gr.q.res = s->q.res + so * NRES; gr.q.stp = s->q.stp + so * NSTP;
    gr.q.pd = s->q.pd + (size_t)t0 * NPD * TL; gr.q.pi = s->q.pi + (size_t)t0 * NPI * TL; gr.q.tile0 = t0;
    gr.counters = s->counters + ncnt * g;
    gr.h_counters = s->h_counters + 2 * g;
  }
  *err = cudaSuccess;
  return s;
}

void qp_destroy(QpSolver* s) {
  if (!s) return;
  for (int g = 0; g < MAX_GROUPS; ++g) {
    if (s->grp[g].stream) { cudaStreamSynchronize(s->grp[g].stream); cudaStreamDestroy(s->grp[g].stream); }
    if (s->grp[g].hi) { cudaStreamSynchronize(s->grp[g].hi); cudaStreamDestroy(s->grp[g].hi); }
    if (s->grp[g].ev) cudaEventDestroy(s->grp[g].ev);
    if (s->grp[g].evx) cudaEventDestroy(s->grp[g].evx);
  }
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  if (s->block) cudaFree(s->block);
  if (s->pi) cudaFree(s->pi);
  if (s->counters) cudaFree(s->counters);
  if (s->h_counters) cudaFreeHost(s->h_counters);
  delete s;
}

double* qp_rec(QpSolver* s) { return const_cast<double*>(s->q.rec); }
int qp_last_iterations(const QpSolver* s) { return s->last_iters; }
int qp_groups(const QpSolver* s) { return s->G; }
void qp_set_profiling(QpSolver* s, bool on) { s->profile = on; }
void qp_get_profile(const QpSolver* s, double* ms, int32_t* n, double* span_ms) {
  for (int i = 0; i < SMPC_PROF_N; ++i) { ms[i] = s->prof_ms[i]; n[i] = s->prof_n[i]; }
  *span_ms = s->prof_span_ms;
}

namespace {
// optional timeline (SMPC_QP_TRACE=1): one event pair per kernel, printed to stderr after the solve
struct TraceRec { int g; const char* name; int kk; cudaEvent_t a, b; };
static std::vector<TraceRec> g_trace;
static bool g_profile = false;
static bool trace_print() { static int v = -1; if (v < 0) v = getenv("SMPC_QP_TRACE") ? 1 : 0; return v == 1; }
static bool trace_on() { return g_profile || trace_print(); }
static int prof_slot(const char* n) {
  static const char* names[SMPC_PROF_N] = {"qs_init_kernel", "qs_prep_kernel", "qs_ctl_kernel", "qs_ric1_kernel", "qs_step_kernel<0>", "qs_ric2_kernel<1>",
                                           "qs_step_kernel<1>", "qs_red_kernel", "qs_ric2_kernel<2>", "qs_step_kernel<2>", "qs_final_kernel"};
  for (int i = 0; i < SMPC_PROF_N; ++i) if (names[i] && !strcmp(names[i], n)) return i;
  return 0;
}

// kernel launches of one tile group on its own stream
struct DeviceBackend {
  int64_t* launches;
  const smpc_problem_t* dP;
  QpSolver* s;
  QpGroup* g;
  const double* x0; const int32_t* r; const uint8_t* act;
  double *xt, *ut; int32_t *status, *qp_iter, *qp_status; double* qp_res;
  int kk_last = 0;
  cudaError_t err = cudaSuccess;
  bool on_hi = false;
  // stream for the next kernel; a change of stream is ordered after everything queued on the other one
  cudaStream_t st(bool hi) {
    if (!g->hi) return g->stream;
    if (hi != on_hi) {
      cudaEventRecord(g->evx, on_hi ? g->hi : g->stream);
      cudaStreamWaitEvent(hi ? g->hi : g->stream, g->evx, 0);
      on_hi = hi;
    }
    return hi ? g->hi : g->stream;
  }
  int sp_grid() const { return (g->T * (s->N + 1) + SP_WARPS - 1) / SP_WARPS; }
  void count(int n = 1) { *launches += n; }
  int gi() const { return (int)(g - s->grp); }
  void tr0(const char* name, cudaStream_t stm) {
    if (!trace_on()) return;
    TraceRec r{gi(), name, kk_last, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stm);
    g_trace.push_back(r);
  }
  void tr1(cudaStream_t stm) { if (trace_on()) cudaEventRecord(g_trace.back().b, stm); }
  void init() {
    cudaMemsetAsync(g->counters, 0, sizeof(int) * 2 * (s->iter_max + 2), st(false));
    { cudaStream_t stm_ = st(false); tr0("qs_init_kernel", stm_); qs_init_kernel<<<g->T, 32, 0, stm_>>>(g->q, s->B, x0, r, act); tr1(stm_); }
    count();
  }
  void prep(int kk) {
    {
      cudaStream_t stm_ = st(false);
      const int grid = (g->T * (s->N + 1) + PREP_WARPS - 1) / PREP_WARPS;
      tr0("qs_prep_kernel", stm_);
      if (kk == 0) qs_prep_kernel<true><<<grid, 32 * PREP_WARPS, PREP_SMEM, stm_>>>(dP, g->q, g->T, kk);
      else qs_prep_kernel<false><<<grid, 32 * PREP_WARPS, PREP_SMEM, stm_>>>(dP, g->q, g->T, kk);
      tr1(stm_);
    }
    count();
  }
  void ctl(int kk) {
    kk_last = kk;
    { cudaStream_t stm_ = st(false); tr0("qs_ctl_kernel", stm_); qs_ctl_kernel<<<g->T, 32, 0, stm_>>>(dP, g->q, kk, status, qp_iter, qp_status, qp_res, g->counters); tr1(stm_); }
    count();
  }
  void ric1() { { cudaStream_t stm_ = st(true); tr0("qs_ric1_kernel", stm_); qs_ric1_kernel<<<g->T, 32, RIC1_SMEM, stm_>>>(dP, g->q); tr1(stm_); } count(); }
  void ric2(int mode) {
    if (mode == 1) { cudaStream_t stm_ = st(true); tr0("qs_ric2_kernel<1>", stm_); qs_ric2_kernel<1><<<g->T, 32, RIC2_SMEM, stm_>>>(dP, g->q); tr1(stm_); }
    else { cudaStream_t stm_ = st(true); tr0("qs_ric2_kernel<2>", stm_); qs_ric2_kernel<2><<<g->T, 32, RIC2_SMEM, stm_>>>(dP, g->q); tr1(stm_); }
    count();
  }
  void step(int kk, int mode) {
    if (mode == 0) { cudaStream_t stm_ = st(false); tr0("qs_step_kernel<0>", stm_); qs_step_kernel<0><<<sp_grid(), 32 * SP_WARPS, 0, stm_>>>(dP, g->q, g->T, kk); tr1(stm_); }
    else if (mode == 1) { cudaStream_t stm_ = st(false); tr0("qs_step_kernel<1>", stm_); qs_step_kernel<1><<<sp_grid(), 32 * SP_WARPS, 0, stm_>>>(dP, g->q, g->T, kk); tr1(stm_); }
    else { cudaStream_t stm_ = st(false); tr0("qs_step_kernel<2>", stm_); qs_step_kernel<2><<<sp_grid(), 32 * SP_WARPS, 0, stm_>>>(dP, g->q, g->T, kk); tr1(stm_); }
    count();
  }
  void final() { { cudaStream_t stm_ = st(false); tr0("qs_final_kernel", stm_); qs_final_kernel<<<sp_grid(), 32 * SP_WARPS, 0, stm_>>>(g->q, g->T, s->B, act, status, xt, ut); tr1(stm_); } count(); }
  void red(bool after) { { cudaStream_t stm_ = st(false); tr0("qs_red_kernel", stm_); qs_red_kernel<<<g->T, 32, 0, stm_>>>(dP, g->q, kk_last, after ? 1 : 0, g->counters); tr1(stm_); } count(); }
  void request_counters() {
    cudaError_t e = cudaMemcpyAsync(g->h_counters, g->counters + 2 * kk_last, 2 * sizeof(int), cudaMemcpyDeviceToHost, st(false));
    if (e == cudaSuccess) e = cudaEventRecord(g->ev, st(false));
    if (e != cudaSuccess) err = e;
  }
  void wait_counters(int& na, int& nr) {
    cudaError_t e = err == cudaSuccess ? cudaEventSynchronize(g->ev) : err;
    if (e != cudaSuccess) { err = e; na = 0; nr = 0; return; }     // stop iterating; the caller reports the error
    na = g->h_counters[0]; nr = g->h_counters[1];
    if (trace_print()) fprintf(stderr, "QPCOUNT g=%d kk=%d active=%d redo=%d\n", gi(), kk_last, na, nr);
  }
};
}  // namespace

cudaError_t launch_qp_solve(const LaunchCtx& c, const smpc_problem_t* dP, QpSolver* s, const double* x0, const int32_t* r, const uint8_t* act,
                            double* xt, double* ut, int32_t* status, int32_t* qp_iter, int32_t* qp_status, double* qp_res) {
  g_profile = s->profile;
  // the group streams start after everything queued on the caller's stream (linearisation, input copies) ...
  cudaError_t e = cudaEventRecord(s->ev_in, c.stream);
  if (e != cudaSuccess) return e;
  DeviceBackend bk[MAX_GROUPS];
  for (int g = 0; g < s->G; ++g) {
    e = cudaStreamWaitEvent(s->grp[g].stream, s->ev_in, 0);
    if (e != cudaSuccess) return e;
    bk[g] = DeviceBackend{c.launches, dP, s, &s->grp[g], x0, r, act, xt, ut, status, qp_iter, qp_status, qp_res};
  }
  s->last_iters = qs_drive(bk, s->G);
  if (trace_on()) {
    for (int g = 0; g < s->G; ++g) cudaStreamSynchronize(bk[g].st(false));
    if (!g_trace.empty()) {
      cudaEvent_t t0 = g_trace[0].a;
      for (int i = 0; i < SMPC_PROF_N; ++i) { s->prof_ms[i] = 0.0; s->prof_n[i] = 0; }
      float t_end = 0;
      for (auto& r : g_trace) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, t0, r.a); cudaEventElapsedTime(&b, t0, r.b);
        if (trace_print()) fprintf(stderr, "QPTRACE g=%d kk=%d %-12s start=%9.3f end=%9.3f dur=%8.3f\n", r.g, r.kk, r.name, a, b, b - a);
        const int sl = prof_slot(r.name);
        s->prof_ms[sl] += b - a; s->prof_n[sl] += 1;
        t_end = b > t_end ? b : t_end;
      }
      s->prof_span_ms = t_end;
      for (auto& r : g_trace) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
      g_trace.clear();
    }
  }
  // ... and the caller's stream continues after the last kernel of every group
  for (int g = 0; g < s->G; ++g) {
    if (bk[g].err != cudaSuccess) return bk[g].err;
    e = cudaEventRecord(s->grp[g].ev, bk[g].st(false));
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, s->grp[g].ev, 0);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
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
