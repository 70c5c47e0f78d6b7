// Viability network on the 5th-generation tensor cores (tcgen05 + TMEM), "fp32 mode".
//
// NeuralNetwork / NetSafeSet of reference safe_set.py:26-43,71-104: c(x) = NN(psi(x)) (100 - alpha)/100 - |v| and dc/dx for
// every (problem, stage) row that carries the viability constraint.  The reference evaluates the network in fp32 inside
// libtorch (L4CasADi) and differentiates it with jacrev; this kernel reproduces that precision class: fp32 storage, fp32
// accumulation, and the two 256 x 256 layers (forward and reverse sweep) as 3xTF32 split products on the tensor cores
// (a = a_hi + a_lo, w = w_hi + w_lo, a w ~ a_hi w_hi + a_lo w_hi + a_hi w_lo: 2^-21 per product, i.e. fp32-class).
// The strict fp64-accumulate kernel in kernels.cu stays the default (smpc_problem_t::nn_precision = 0).
//
// One CTA works on tiles of R = 64 rows.  Per tile
//   layer 1 (10 -> 256), the output layer (256 -> 1) and the last reverse layer (256 -> 10) run on the CUDA cores (2 % of the flops);
//   the four 256 x 256 contractions (forward 2, 3; reverse 3, 2) run as tcgen05.mma.kind::tf32 with
//     D[unit (M = 128 per half)][row (N = 64)] += W[unit][k] * X[row][k]          (both operands K-major in shared memory)
//   so that TMEM lane = hidden unit, TMEM column = row of the tile: the epilogue thread of lane j owns unit j (bias, GELU and
//   its derivative are per-thread scalars/elementwise) and writes the next layer's operand X[row][k = j] straight back to
//   shared memory in the canonical K-major core-matrix layout.  GELU derivatives of layers 1 and 2 stay in TMEM
//   (tcgen05.st) until the reverse sweep multiplies them in.  TMEM: accumulator 128 columns, D1 128, D2 128.
//   Weights are pre-split (hi / lo) and pre-packed on the host into the exact shared-memory image of a K-stage, and streamed
//   L2 -> shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier) by a producer thread; one elected thread issues
//   the MMAs; tcgen05.commit releases weight stages and signals the epilogue.
// Warp roles: 0-7 epilogue / CUDA-core layers (warp w: TMEM lanes 32 (w % 4) .. + 31, rows 32 (w / 4) .. + 31 of the tile; two warps per
// scheduler because the epilogue is a long dependent instruction stream), 8 weight producer, 9 MMA issuer.
#include <cstring>
#include <vector>

#include "engine.cuh"
#include "mlp_tc_common.cuh"

namespace smpc {

namespace {

constexpr int R = 64;                       // rows per tile = MMA N
constexpr int HID = SMPC_HID;               // 256
constexpr int KC = 16;                      // k-values per weight stage (two K = 8 MMA steps)
constexpr int NSTG = HID / KC;              // stages per layer
constexpr int NSLOT = 2;                    // weight ring slots
constexpr int XPITCH = R * 16 + 16;         // bytes between consecutive 16-byte k-chunks of X (+16: conflict-free epilogue stores)
constexpr int XBYTES = (HID / 4) * XPITCH;  // one of X_hi / X_lo
constexpr int WROWB = 128 * 16;             // bytes of one k-chunk of one half of the weight operand: 128 units x 16 B
constexpr int WHALF = (KC / 4) * WROWB;     // one (hi|lo, half) block of a stage
constexpr int WSTAGE = 4 * WHALF;           // [hl][half][chunk][unit][4 floats] = 32 KB
constexpr int NCW = 8;                       // epilogue / CUDA-core warps: warp w owns TMEM lanes 32 (w % 4) .., columns 32 (w / 4) ..
constexpr int TC_THREADS = 32 * (NCW + 2);
constexpr int D1COL = 128, D2COL = 256, TMEM_COLS = 512;

constexpr int OFF_XH = 0;
constexpr int OFF_XL = OFF_XH + XBYTES;
constexpr int OFF_W = OFF_XL + XBYTES;
constexpr int OFF_W1 = OFF_W + NSLOT * WSTAGE;          // float [256][10]
constexpr int OFF_VEC = OFF_W1 + HID * NX * 4;          // float b1[256] b2[256] b3[256] W4[256]
constexpr int OFF_INF = OFF_VEC + 4 * HID * 4;          // float [R][10]
constexpr int OFF_GIN = OFF_INF + R * NX * 4;           // float [4][R][10]  (partial sums over a quarter of the hidden units)
constexpr int OFF_YP = OFF_GIN + 4 * R * NX * 4;        // float [4][R]
constexpr int OFF_META = OFF_YP + 4 * R * 4;            // int rowb[R], rowk[R], valid[R], vote[4]
constexpr int OFF_BAR = OFF_META + (3 * R + 4) * 4;     // uint64 full[NSLOT], empty[NSLOT], acc_full, x_ready
constexpr int OFF_TMEM = OFF_BAR + (2 * NSLOT + 2) * 8;
constexpr int TC_SMEM = OFF_TMEM + 16;
static_assert(OFF_W % 128 == 0 && OFF_BAR % 8 == 0, "alignment");
static_assert(TC_SMEM <= 232448, "shared memory budget");

// instruction descriptor of tcgen05.mma.kind::tf32: D fp32, A/B tf32, both K-major, M = 128, N = R
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(R >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

using namespace tcg;

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NCW) : "memory"); }
// byte offset of X[row n][k] inside X_hi / X_lo
__device__ __forceinline__ int xoff(int n, int k) { return (k >> 2) * XPITCH + n * 16 + (k & 3) * 4; }

__device__ __forceinline__ void x_store(unsigned char* sm, int n, int k, float a) {
  float hi, lo;
  split_tf32(a, hi, lo);
  const int o = xoff(n, k);
  *reinterpret_cast<float*>(sm + OFF_XH + o) = hi;
  *reinterpret_cast<float*>(sm + OFF_XL + o) = lo;
}

// does tile `tile` hold at least one row that has to be evaluated?  (whole warp; same answer in every role)
__device__ __forceinline__ bool tile_any(int tile, int rows_mode, int n_rows, int B, int N, const int32_t* r, const uint8_t* act,
                                         const uint8_t* need, int lane) {
  bool v = false;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = tile * R + h * 32 + lane;
    int b, k;
    v = v || (i < n_rows && mlp_row(rows_mode, i, B, N, r, act, need, b, k));
  }
  return __any_sync(0xffffffffu, v) != 0;
}

enum { L_FWD2 = 0, L_FWD3 = 1, L_BWD3 = 2, L_BWD2 = 3 };

// epilogue of one tensor-core layer for the thread that owns TMEM lane j (units j and 128 + j)
template <int L>
__device__ __forceinline__ void epilogue(unsigned char* sm, uint32_t tlane, int j, int c, int warp, int lane, bool want_grad) {
  const float* vec = reinterpret_cast<const float*>(sm + OFF_VEC);
  float* YP = reinterpret_cast<float*>(sm + OFF_YP);
  {
    float yp[32];
    if (L == L_FWD3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) yp[i] = 0.0f;
    }
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int unit = h * 128 + j;
      uint32_t v[32], dv[32];
      tmem_ld32(tlane + h * R + c * 32, v);
      if (L == L_BWD3) tmem_ld32(tlane + D2COL + h * R + c * 32, dv);
      if (L == L_BWD2) tmem_ld32(tlane + D1COL + h * R + c * 32, dv);
      tmem_ld_wait();
      const float bias = L == L_FWD2 ? vec[HID + unit] : (L == L_FWD3 ? vec[2 * HID + unit] : 0.0f);
      const float w4 = vec[3 * HID + unit];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = c * 32 + i;
        const float x = __uint_as_float(v[i]);
        if (L == L_FWD2) {
          float d;
          const float a = gelu_f32(x + bias, d);
          dv[i] = __float_as_uint(d);
          x_store(sm, n, unit, a);
        } else if (L == L_FWD3) {
          float d;
          const float a = gelu_f32(x + bias, d);
          yp[i] = fmaf(w4, a, yp[i]);
          if (want_grad) x_store(sm, n, unit, w4 * d);
        } else {
          x_store(sm, n, unit, x * __uint_as_float(dv[i]));
        }
      }
      if (L == L_FWD2) tmem_st32(tlane + D2COL + h * R + c * 32, dv);
    }
    if (L == L_FWD3) {
      // sum over the 32 lanes of the warp, 32 rows at once: after the butterfly lane l holds the total of row c * 32 + l
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const bool up = (lane & s) != 0;
          const float keep = up ? yp[i + s] : yp[i];
          const float send = up ? yp[i] : yp[i + s];
          yp[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      YP[(warp & 3) * R + c * 32 + lane] = yp[0];
    }
  }
  if (L == L_FWD2) tmem_st_wait();
}

__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_kernel(const smpc_problem_t* __restrict__ dP, MlpTcWeights w, int B, int N, int rows_mode, int n_rows, const double* __restrict__ xsrc,
              const int32_t* __restrict__ r, const uint8_t* __restrict__ act, const uint8_t* __restrict__ need, double* out11, int want_grad) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* bar_empty = bar_full + NSLOT;
  uint64_t* bar_acc = bar_empty + NSLOT;
  uint64_t* bar_x = bar_acc + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
  const int n_tiles = (n_rows + R - 1) / R;
  const int n_layers = want_grad ? 4 : 2;

  // ---- one-time setup: small weights to shared memory, barriers, TMEM ----
  {
    float* W1s = reinterpret_cast<float*>(sm + OFF_W1);
    float* vec = reinterpret_cast<float*>(sm + OFF_VEC);
    for (int i = tid; i < HID * NX; i += TC_THREADS) W1s[i] = w.W1[i];
    for (int i = tid; i < HID; i += TC_THREADS) { vec[i] = w.b1[i]; vec[HID + i] = w.b2[i]; vec[2 * HID + i] = w.b3[i]; vec[3 * HID + i] = w.W4[i]; }
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_x, 32 * NCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == NCW) {
    // =========================== weight producer ===========================
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (!tile_any(tile, rows_mode, n_rows, B, N, r, act, need, lane)) continue;
      if (lane == 0) {
        for (int l = 0; l < n_layers; ++l)
          for (int s = 0; s < NSTG; ++s, ++cnt) {
            const int slot = cnt % NSLOT;
            mbar_wait(bar_empty + slot, ((cnt / NSLOT) & 1) ^ 1);
            mbar_expect(bar_full + slot, WSTAGE);
            bulk_g2s(sm + OFF_W + slot * WSTAGE, reinterpret_cast<const unsigned char*>(w.packed) + ((size_t)l * NSTG + s) * WSTAGE, WSTAGE,
                     bar_full + slot);
          }
      }
      __syncwarp();
    }
  } else if (warp == NCW + 1) {
    // =========================== MMA issuer ===========================
    uint32_t cnt = 0, xph = 0;
    const uint32_t xh = s32(sm + OFF_XH), xl = s32(sm + OFF_XL), wb = s32(sm + OFF_W);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (!tile_any(tile, rows_mode, n_rows, B, N, r, act, need, lane)) continue;
      if (lane == 0) {
        for (int l = 0; l < n_layers; ++l) {
          mbar_wait(bar_x, xph); xph ^= 1;                  // operand X of this layer is in shared memory, accumulator drained
          fence_after();
          for (int s = 0; s < NSTG; ++s, ++cnt) {
            const int slot = cnt % NSLOT;
            mbar_wait(bar_full + slot, (cnt / NSLOT) & 1);
            fence_after();
            const uint32_t ws = wb + slot * WSTAGE;
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
              const uint32_t xo = (uint32_t)((s * (KC / 4) + ks * 2) * XPITCH);
              const uint64_t bh = smem_desc(xh + xo, XPITCH, 128), bl = smem_desc(xl + xo, XPITCH, 128);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t ah = smem_desc(ws + h * WHALF + ks * 2 * WROWB, WROWB, 128);
                const uint64_t al = smem_desc(ws + 2 * WHALF + h * WHALF + ks * 2 * WROWB, WROWB, 128);
                const uint32_t d = tmem + h * R;
                umma_tf32(d, al, bh, IDESC, (s | ks) != 0);       // small terms first
                umma_tf32(d, ah, bl, IDESC, 1);
                umma_tf32(d, ah, bh, IDESC, 1);
              }
            }
            umma_commit(bar_empty + slot);                  // the stage may be refilled once these MMAs have read it
          }
          umma_commit(bar_acc);                             // accumulator of the layer complete
        }
      }
      __syncwarp();
    }
  } else {
    // =========================== CUDA-core layers + epilogues (256 threads) ===========================
    const smpc_problem_t& P = *dP;
    const int j = tid & 127;                                // TMEM lane = hidden unit (and unit 128 + j)
    const int cg = tid >> 7;                                // which 32 rows (TMEM columns) of the tile this thread handles
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* INF = reinterpret_cast<float*>(sm + OFF_INF);
    float* GIN = reinterpret_cast<float*>(sm + OFF_GIN);
    float* YP = reinterpret_cast<float*>(sm + OFF_YP);
    int* rowb = reinterpret_cast<int*>(sm + OFF_META);
    int* rowk = rowb + R;
    int* valid = rowk + R;
    int* vote = valid + R;
    const float* W1s = reinterpret_cast<const float*>(sm + OFF_W1);
    const float* vec = reinterpret_cast<const float*>(sm + OFF_VEC);
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // ---- gather the rows of the tile: psi(x) in fp64, fp32 copy for the network ----
      if (tid < R) {
        int b = 0, k = 0;
        const int i = tile * R + tid;
        const bool v = i < n_rows && mlp_row(rows_mode, i, B, N, r, act, need, b, k);
        valid[tid] = v; rowb[tid] = b; rowk[tid] = k;
        double in[NX], nrm = 1.0;
        if (v) {
          const double* x = (rows_mode == ROWS_FLAT) ? xsrc + (size_t)b * NX : xsrc + ((size_t)b * (N + 1) + k) * NX;
          nn_input(P, x, in, &nrm);
        } else {
#pragma unroll
          for (int q = 0; q < NX; ++q) in[q] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < NX; ++q) INF[tid * NX + q] = (float)in[q];
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (lane == 0) vote[warp] = m != 0;
      }
      bar_compute();
      const bool any = vote[0] | vote[1];
      if (!any) { bar_compute(); continue; }               // (second barrier: vote[] is rewritten by the next tile)

      // ---- layer 1 on the CUDA cores: a1 -> X, GELU' -> TMEM D1 ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int unit = h * 128 + j;
        float w1[NX];
#pragma unroll
        for (int q = 0; q < NX; ++q) w1[q] = W1s[unit * NX + q];
        const float bias = vec[unit];
        {
          const int c = cg;
          uint32_t dv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int n = c * 32 + i;
            float acc = bias;
#pragma unroll
            for (int q = 0; q < NX; ++q) acc = fmaf(w1[q], INF[n * NX + q], acc);
            float d;
            const float a = gelu_f32(acc, d);
            dv[i] = __float_as_uint(d);
            x_store(sm, n, unit, a);
          }
          tmem_st32(tlane + D1COL + h * R + c * 32, dv);
        }
      }
      tmem_st_wait();
      fence_async_smem(); fence_before(); mbar_arrive(bar_x);

      // ---- tensor-core layers ----
      mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
      epilogue<L_FWD2>(sm, tlane, j, cg, warp, lane, want_grad);
      fence_async_smem(); fence_before(); mbar_arrive(bar_x);

      mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
      epilogue<L_FWD3>(sm, tlane, j, cg, warp, lane, want_grad);
      if (want_grad) {
        fence_async_smem(); fence_before(); mbar_arrive(bar_x);
        mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
        epilogue<L_BWD3>(sm, tlane, j, cg, warp, lane, true);
        fence_async_smem(); fence_before(); mbar_arrive(bar_x);
        mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
        epilogue<L_BWD2>(sm, tlane, j, cg, warp, lane, true);   // X = g1 (hi + lo)
      }
      fence_before();
      bar_compute();

      // ---- last reverse layer on the CUDA cores: gin[n][i] = sum_k W1[k][i] g1[k][n] ----
      if (want_grad) {
        const int n = tid & (R - 1), part = tid >> 6;       // part: quarter of the hidden units
        float g[NX];
#pragma unroll
        for (int q = 0; q < NX; ++q) g[q] = 0.0f;
#pragma unroll 4
        for (int ch = part * (HID / 16); ch < (part + 1) * (HID / 16); ++ch) {
          const float4 hi = *reinterpret_cast<const float4*>(sm + OFF_XH + ch * XPITCH + n * 16);
          const float4 lo = *reinterpret_cast<const float4*>(sm + OFF_XL + ch * XPITCH + n * 16);
          const float gv[4] = {hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int q = 0; q < NX; ++q) g[q] = fmaf(W1s[(ch * 4 + e) * NX + q], gv[e], g[q]);
        }
#pragma unroll
        for (int q = 0; q < NX; ++q) GIN[(part * R + n) * NX + q] = g[q];
      }
      bar_compute();
      // ---- c(x), dc/dx in fp64 from the fp32 network output (safe_set.py:100-104) ----
      if (tid < R && valid[tid]) {
        const double y = (double)w.b4[0] + ((double)YP[tid] + (double)YP[R + tid] + (double)YP[2 * R + tid] + (double)YP[3 * R + tid]);
        double gin[NX], grad[NX];
#pragma unroll
        for (int q = 0; q < NX; ++q)
          gin[q] = want_grad ? (double)((GIN[tid * NX + q] + GIN[(R + tid) * NX + q]) + (GIN[(2 * R + tid) * NX + q] + GIN[(3 * R + tid) * NX + q])) : 0.0;
        double in[NX], nrm;                                // psi(x) again in fp64 (cheaper than keeping it in shared memory)
        nn_input(P, (rows_mode == ROWS_FLAT) ? xsrc + (size_t)rowb[tid] * NX : xsrc + ((size_t)rowb[tid] * (N + 1) + rowk[tid]) * NX, in, &nrm);
        const double cval = nn_output(P, in, nrm, y, want_grad ? gin : nullptr, want_grad ? grad : nullptr);
        double* o = (rows_mode == ROWS_FLAT) ? out11 + (size_t)rowb[tid] * NN_OUT : out11 + ((size_t)rowb[tid] * (N + 1) + rowk[tid]) * NN_OUT;
        o[0] = cval;
        if (want_grad) {
#pragma unroll
          for (int q = 0; q < NX; ++q) o[1 + q] = grad[q];
        }
      }
      bar_compute();                                       // tile buffers (meta, IN, YP, GIN, X) are free again
    }
  }

  // ---- teardown ----
  fence_before();
  __syncthreads();
  if (warp == 0) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace

// Host: split the four 256 x 256 operands into tf32 hi / lo and lay them out as the shared-memory images of the K-stages.
// Operand of layer l is A_l[m][k] (m = output unit of the contraction, k = contraction index):
//   forward 2: W2[m][k], forward 3: W3[m][k], reverse 3: W3[k][m], reverse 2: W2[k][m]      (nn.Linear: y = W x + b, W[out][in])
// image of stage s: [hi|lo][half = m / 128][chunk = (k % KC) / 4][row = m % 128][e = k % 4]
size_t mlp_tc_packed_floats() { return (size_t)4 * NSTG * (WSTAGE / 4); }

void mlp_tc_pack(const float* W2, const float* W3, float* out) {
  for (int l = 0; l < 4; ++l) {
    const float* W = (l == 0 || l == 3) ? W2 : W3;
    const bool transposed = l >= 2;
    for (int m = 0; m < HID; ++m)
      for (int k = 0; k < HID; ++k) {
        const float a = transposed ? W[(size_t)k * HID + m] : W[(size_t)m * HID + k];
        float hi, lo;
        split_tf32(a, hi, lo);
        const int s = k / KC, chunk = (k % KC) / 4, e = k % 4, half = m / 128, row = m % 128;
        const size_t base = ((size_t)l * NSTG + s) * (WSTAGE / 4);
        const size_t o = (size_t)half * (WHALF / 4) + (size_t)chunk * (WROWB / 4) + (size_t)row * 4 + e;
        out[base + o] = hi;
        out[base + 2 * (WHALF / 4) + o] = lo;
      }
  }
}

cudaError_t mlp_tc_prepare() {
  // per device: called by smpc_create after cudaSetDevice (the attribute belongs to the device, not to the process)
  cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  return e;
}

void launch_mlp_tc(const LaunchCtx& c, const smpc_problem_t* dP, const MlpTcWeights& w, int n_sm, int B, int N, int rows_mode, int n_flat,
                   const double* xsrc, const int32_t* r, const uint8_t* act, const uint8_t* need, double* out11, bool want_grad) {
  int n_rows = (rows_mode == ROWS_TERMINAL || rows_mode == ROWS_CAND) ? B : rows_mode == ROWS_ALL ? B * N : rows_mode == ROWS_RECEDING ? 2 * B : n_flat;
  if (rows_mode == ROWS_FLAT) B = n_flat;
  if (n_rows <= 0) return;
  const int n_tiles = (n_rows + R - 1) / R;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  mlp_tc_kernel<<<grid, TC_THREADS, TC_SMEM, c.stream>>>(dP, w, B, N, rows_mode, n_rows, xsrc, r, act, need, out11, want_grad ? 1 : 0);
  ++*c.launches;
}

}  // namespace smpc
