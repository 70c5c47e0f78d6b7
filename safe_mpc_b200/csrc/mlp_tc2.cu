// Viability network on the tensor cores, CTA-pair form (tcgen05 cta_group::2): the kernel of mlp_tc.cu with the weight ingest per SM
// halved.  mlp_tc.cu is bound by the weight stream -- every 64-row tile re-streams the 2 MB of split weights through a 64 KB ring
// (profiles/r01_mlp_tc.md).  Here two CTAs of a cluster work on 128 rows together:
//   D[unit (M = 256 = 2 x 128)][row (N = 128 = 2 x 64)] += W[unit][k] * X[row][k]        tcgen05.mma.cta_group::2.kind::tf32
// CTA r holds the weights (A operand) and the accumulators (TMEM lanes) of the units 128 r .. 128 r + 127, and the operand X (B
// operand) of the rows 64 r .. 64 r + 63 of the pair tile: each CTA streams only ITS half of the weights (1 MB per 128 rows), one
// instruction does the work of four of the single-CTA kernel, and the leader CTA issues all of them.
// The epilogue thread of TMEM lane j in CTA r owns unit 128 r + j for all 128 rows; it writes the next operand X[row][k = unit] into the
// shared memory of the CTA that owns the row (rows 64 .. 127: CTA 1) -- local or through distributed shared memory
// (st.shared::cluster).  Barriers: operand-ready = one mbarrier in the leader with 512 arrivals (256 epilogue threads of each CTA,
// the peer's through mapa + mbarrier.arrive.release.cluster); accumulator-ready and stage-free = tcgen05.commit multicast to both
// CTAs; stage-full of the peer is relayed to the leader by a relay thread.  Arithmetic, split, GELU, layouts, packing per (unit, k):
// as in mlp_tc.cu (3xTF32, fp32 class).
//
// Status (round 1): parity-green (tests/test_gpu_mlp_tc.py runs both kernels), selected with SMPC_MLP_TC=pair, NOT the default:
// 3.37 ms per 460k rows against 2.99 ms for mlp_tc.cu.  Phase timing of one pair tile (clock64, profiles/r01_mlp_tc.md): the MMA phase of
// a layer dropped from ~13 us to 7.7 us as intended (half the weights per CTA), but it is still 2.5x the tensor time (96 MMAs x 64
// cycles): 64 KB of weight ring per CTA turn over once per ~2 us (TMA latency + MMA completion + commit + relay), i.e. 33 GB/s per SM
// whatever the stage size; and the CUDA-core phases of a CTA (layer 1, four epilogues, cluster syncs, DSMEM stores) add 29 us that
// nothing overlaps.  Next: release a weight stage when it has been copied to TMEM (tcgen05.cp, A operand from TMEM) instead of when
// its MMAs have completed, and stagger the epilogues of the two unit halves under the MMAs.
#include <cstring>
#include <vector>

#include "engine.cuh"
#include "mlp_tc_common.cuh"

namespace smpc {

namespace {

using namespace tcg;

constexpr int RP = 128;                     // rows per pair tile
constexpr int RH = 64;                      // rows per CTA (its half of the B operand)
constexpr int HID = SMPC_HID;               // 256
constexpr int UH = 128;                     // units per CTA (its half of the A operand / of the accumulator rows)
constexpr int KC = 16;                      // k-values per weight stage
constexpr int NSTG = HID / KC;              // 16 stages per layer
constexpr int NSLOT = 4;
constexpr int XPITCH = RH * 16 + 16;
constexpr int XBYTES = (HID / 4) * XPITCH;
constexpr int WROWB = UH * 16;              // one k-chunk of the CTA's 128 units
constexpr int WHALF = (KC / 4) * WROWB;     // hi (or lo) block of a stage: 8 KB
constexpr int WSTAGE = 2 * WHALF;           // 16 KB per CTA and stage
constexpr int NCW = 16;                     // epilogue warps: warp w owns TMEM lanes 32 (w % 4) .., rows 32 (w / 4) .. of the pair tile (four warps per
                                            // scheduler: the epilogue is a long dependent fp32 instruction stream)
constexpr int CW = 16;                      // rows per TMEM access of an epilogue thread
constexpr int W_PROD = NCW, W_MMA = NCW + 1, W_RELAY = NCW + 2;
constexpr int TC2_THREADS = 32 * (NCW + 3);
constexpr int ACCCOL = 0, D1COL = 128, D2COL = 256, TMEM_COLS = 512;

constexpr int OFF_XH = 0;
constexpr int OFF_XL = OFF_XH + XBYTES;
constexpr int OFF_W = OFF_XL + XBYTES;
constexpr int OFF_W1 = OFF_W + NSLOT * WSTAGE;          // float [256][10]   (all units: the last reverse layer contracts over all of them)
constexpr int OFF_VEC = OFF_W1 + HID * NX * 4;          // float b1[128] b2[128] b3[128] W4[128] of this CTA's units
constexpr int OFF_INF = OFF_VEC + 4 * UH * 4;           // float [RP][10]    psi(x) of the 128 rows of the pair tile
constexpr int OFF_GIN = OFF_INF + RP * NX * 4;          // float [4][RH][10]
constexpr int OFF_YP = OFF_GIN + 4 * RH * NX * 4;       // float [4][RP]     partial network outputs over this CTA's units
constexpr int OFF_YQ = OFF_YP + 4 * RP * 4;             // float [4][RH]     the peer's partials for this CTA's rows
constexpr int OFF_META = OFF_YQ + 4 * RH * 4;           // int rowb[RH], rowk[RH], valid[RH], vote[4]
constexpr int OFF_BAR = OFF_META + (3 * RH + 4) * 4;    // uint64 full[4], empty[4], peer_full[4], acc, x, cbar
constexpr int NBAR = 3 * NSLOT + 3;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int TC2_SMEM = OFF_TMEM + 16;
static_assert(OFF_W % 128 == 0 && OFF_BAR % 8 == 0, "alignment");
static_assert(TC2_SMEM <= 232448, "shared memory budget");

// M = 256 (pair), N = 128, tf32 x tf32 -> fp32, both operands K-major
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(RP >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NCW) : "memory"); }
__device__ __forceinline__ int xoff(int n, int k) { return (k >> 2) * XPITCH + n * 16 + (k & 3) * 4; }

// X[row n of CTA `dst`][k] = a (hi and lo); xh / xl: cluster addresses of X_hi / X_lo of the destination CTA
__device__ __forceinline__ void x_store2(uint32_t xh, uint32_t xl, int n, int k, float a) {
  float hi, lo;
  split_tf32(a, hi, lo);
  const uint32_t o = (uint32_t)xoff(n, k);
  st_cluster_f32(xh + o, hi);
  st_cluster_f32(xl + o, lo);
}

// does the pair tile hold at least one row that has to be evaluated?  (whole warp; same answer in every role of both CTAs)
__device__ __forceinline__ bool pair_any(int tile, int rows_mode, int n_rows, int B, int N, const int32_t* r, const uint8_t* act,
                                         const uint8_t* need, int lane) {
  bool v = false;
#pragma unroll
  for (int h = 0; h < RP / 32; ++h) {
    const int i = tile * RP + h * 32 + lane;
    int b, k;
    v = v || (i < n_rows && mlp_row(rows_mode, i, B, N, r, act, need, b, k));
  }
  return __any_sync(0xffffffffu, v) != 0;
}

enum { L_FWD2 = 0, L_FWD3 = 1, L_BWD3 = 2, L_BWD2 = 3 };

// epilogue of one tensor-core layer: the thread owns unit 128 rank + j (TMEM lane j) and the 64 rows of CTA `cg` of the pair tile
template <int L>
__device__ __forceinline__ void epilogue2(unsigned char* sm, uint32_t tlane, uint32_t xh, uint32_t xl, int unit, int j, int cq, int warp, int lane,
                                          bool want_grad) {
  const float* vec = reinterpret_cast<const float*>(sm + OFF_VEC);
  float* YP = reinterpret_cast<float*>(sm + OFF_YP);
  const float bias = L == L_FWD2 ? vec[UH + j] : (L == L_FWD3 ? vec[2 * UH + j] : 0.0f);
  const float w4 = vec[3 * UH + j];
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int col = cq * 32 + c * CW;                    // first row (TMEM column) of this chunk in the pair tile
    uint32_t v[CW], dv[CW];
    tmem_ld16(tlane + ACCCOL + col, v);
    if (L == L_BWD3) tmem_ld16(tlane + D2COL + col, dv);
    if (L == L_BWD2) tmem_ld16(tlane + D1COL + col, dv);
    tmem_ld_wait();
    float yp[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      const int n = (cq & 1) * 32 + c * CW + i;          // row inside the destination CTA
      const float x = __uint_as_float(v[i]);
      if (L == L_FWD2) {
        float d;
        const float a = gelu_f32(x + bias, d);
        dv[i] = __float_as_uint(d);
        x_store2(xh, xl, n, unit, a);
      } else if (L == L_FWD3) {
        float d;
        const float a = gelu_f32(x + bias, d);
        yp[i] = w4 * a;
        if (want_grad) x_store2(xh, xl, n, unit, w4 * d);
      } else {
        x_store2(xh, xl, n, unit, x * __uint_as_float(dv[i]));
      }
    }
    if (L == L_FWD2) tmem_st16(tlane + D2COL + col, dv);
    if (L == L_FWD3) {
      // sum over the 32 lanes (units) of the warp, 16 rows at once: each butterfly step halves the rows a lane keeps; after four
      // steps the lanes 2 m and 2 m + 1 hold the two halves of row col + m
#pragma unroll
      for (int s = 16, cnt = CW / 2; s >= 2; s >>= 1, cnt >>= 1) {
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
          const bool up = (lane & s) != 0;
          const float keep = up ? yp[i + cnt] : yp[i];
          const float send = up ? yp[i] : yp[i + cnt];
          yp[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      yp[0] += __shfl_xor_sync(0xffffffffu, yp[0], 1);
      if ((lane & 1) == 0) YP[(warp & 3) * RP + col + (lane >> 1)] = yp[0];
    }
  }
  if (L == L_FWD2) tmem_st_wait();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
mlp_tc2_kernel(const smpc_problem_t* __restrict__ dP, MlpTcWeights w, int B, int N, int rows_mode, int n_rows, const double* __restrict__ xsrc,
               const int32_t* __restrict__ r, const uint8_t* __restrict__ act, const uint8_t* __restrict__ need, double* out11, int want_grad) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int n_clusters = gridDim.x / 2, cid = (int)cluster_id_x();
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* bar_empty = bar_full + NSLOT;
  uint64_t* bar_pfull = bar_empty + NSLOT;                  // (leader) the peer's stage has landed
  uint64_t* bar_acc = bar_pfull + NSLOT;
  uint64_t* bar_x = bar_acc + 1;                            // (leader) operand X of both CTAs written, accumulators drained
  uint64_t* bar_c = bar_x + 1;                              // compute threads of both CTAs have passed a cluster-wide point
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
  const int n_tiles = (n_rows + RP - 1) / RP;
  const int n_layers = want_grad ? 4 : 2;

  // ---- one-time setup ----
  {
    float* W1s = reinterpret_cast<float*>(sm + OFF_W1);
    float* vec = reinterpret_cast<float*>(sm + OFF_VEC);
    for (int i = tid; i < HID * NX; i += TC2_THREADS) W1s[i] = w.W1[i];
    for (int i = tid; i < UH; i += TC2_THREADS) {
      const int u = UH * rank + i;
      vec[i] = w.b1[u]; vec[UH + i] = w.b2[u]; vec[2 * UH + i] = w.b3[u]; vec[3 * UH + i] = w.W4[u];
    }
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); mbar_init(bar_pfull + s, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_x, 2 * 32 * NCW);
    mbar_init(bar_c, 2 * 32 * NCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_PROD) {
    // =========================== weight producer (this CTA's half of every stage) ===========================
    uint32_t cnt = 0;
    for (int tile = cid; tile < n_tiles; tile += n_clusters) {
      if (!pair_any(tile, rows_mode, n_rows, B, N, r, act, need, lane)) continue;
      if (lane == 0) {
        for (int l = 0; l < n_layers; ++l)
          for (int s = 0; s < NSTG; ++s, ++cnt) {
            const int slot = cnt % NSLOT;
            mbar_wait(bar_empty + slot, ((cnt / NSLOT) & 1) ^ 1);
            mbar_expect(bar_full + slot, WSTAGE);
            bulk_g2s(sm + OFF_W + slot * WSTAGE,
                     reinterpret_cast<const unsigned char*>(w.packed2) + (((size_t)l * NSTG + s) * 2 + rank) * WSTAGE, WSTAGE, bar_full + slot);
          }
      }
      __syncwarp();
    }
  } else if (warp == W_RELAY) {
    // =========================== (peer) tell the leader that this CTA's stage has landed ===========================
    if (rank == 1) {
      uint32_t cnt = 0;
      const uint32_t pf0 = map_cluster(s32(bar_pfull), 0);
      for (int tile = cid; tile < n_tiles; tile += n_clusters) {
        if (!pair_any(tile, rows_mode, n_rows, B, N, r, act, need, lane)) continue;
        if (lane == 0) {
          for (int l = 0; l < n_layers; ++l)
            for (int s = 0; s < NSTG; ++s, ++cnt) {
              const int slot = cnt % NSLOT;
              mbar_wait(bar_full + slot, (cnt / NSLOT) & 1);
              mbar_arrive_cluster(pf0 + slot * 8);
            }
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (rank == 0) {
      uint32_t cnt = 0, xph = 0;
      const uint32_t xh = s32(sm + OFF_XH), xl = s32(sm + OFF_XL), wb = s32(sm + OFF_W);
      for (int tile = cid; tile < n_tiles; tile += n_clusters) {
        if (!pair_any(tile, rows_mode, n_rows, B, N, r, act, need, lane)) continue;
        if (lane == 0) {
          for (int l = 0; l < n_layers; ++l) {
            mbar_wait_cluster(bar_x, xph); xph ^= 1;
            fence_after();
            for (int s = 0; s < NSTG; ++s, ++cnt) {
              const int slot = cnt % NSLOT;
              mbar_wait(bar_full + slot, (cnt / NSLOT) & 1);
              mbar_wait_cluster(bar_pfull + slot, (cnt / NSLOT) & 1);
              fence_after();
              const uint32_t ws = wb + slot * WSTAGE;
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {
                const uint32_t xo = (uint32_t)((s * (KC / 4) + ks * 2) * XPITCH);
                const uint64_t bh = smem_desc(xh + xo, XPITCH, 128), bl = smem_desc(xl + xo, XPITCH, 128);
                const uint64_t ah = smem_desc(ws + ks * 2 * WROWB, WROWB, 128);
                const uint64_t al = smem_desc(ws + WHALF + ks * 2 * WROWB, WROWB, 128);
                umma2_tf32(tmem + ACCCOL, al, bh, IDESC2, (s | ks) != 0);     // small terms first
                umma2_tf32(tmem + ACCCOL, ah, bl, IDESC2, 1);
                umma2_tf32(tmem + ACCCOL, ah, bh, IDESC2, 1);
              }
              umma2_commit(bar_empty + slot);                 // both CTAs may refill the stage once these MMAs have read it
            }
            umma2_commit(bar_acc);                            // accumulators of the layer complete (both CTAs)
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== CUDA-core layers + epilogues (256 threads per CTA) ===========================
    const smpc_problem_t& P = *dP;
    const int j = tid & 127;                                // TMEM lane = unit 128 rank + j
    const int cq = tid >> 7;                                // rows 32 cq .. 32 cq + 31 of the pair tile
    const int cg = cq >> 1;                                 // ... which belong to CTA cg
    const int unit = UH * (int)rank + j;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t xh_dst = map_cluster(s32(sm + OFF_XH), (uint32_t)cg), xl_dst = map_cluster(s32(sm + OFF_XL), (uint32_t)cg);
    const uint32_t peer = rank ^ 1u;
    const uint32_t bx_leader = map_cluster(s32(bar_x), 0);
    const uint32_t bc_self = map_cluster(s32(bar_c), rank), bc_peer = map_cluster(s32(bar_c), peer);
    float* INF = reinterpret_cast<float*>(sm + OFF_INF);
    float* GIN = reinterpret_cast<float*>(sm + OFF_GIN);
    float* YP = reinterpret_cast<float*>(sm + OFF_YP);
    float* YQ = reinterpret_cast<float*>(sm + OFF_YQ);
    int* rowb = reinterpret_cast<int*>(sm + OFF_META);
    int* rowk = rowb + RH;
    int* valid = rowk + RH;
    int* vote = valid + RH;
    const float* W1s = reinterpret_cast<const float*>(sm + OFF_W1);
    const float* vec = reinterpret_cast<const float*>(sm + OFF_VEC);
    uint32_t aph = 0, cph = 0;
    // every compute thread of the pair arrives on the barrier of both CTAs, then waits on its own
    auto cluster_sync_compute = [&]() {
      mbar_arrive_cluster(bc_self);
      mbar_arrive_cluster(bc_peer);
      mbar_wait_cluster(bar_c, cph); cph ^= 1;
    };
    auto x_ready = [&]() { asm volatile("fence.proxy.async;" ::: "memory"); fence_before(); mbar_arrive_cluster(bx_leader); };
    for (int tile = cid; tile < n_tiles; tile += n_clusters) {
      // ---- gather this CTA's 64 rows: psi(x) in fp64, fp32 copy into the INF of both CTAs ----
      if (tid < RH) {
        int b = 0, k = 0;
        const int i = tile * RP + RH * (int)rank + tid;
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
        const uint32_t inf_peer = map_cluster(s32(INF), peer);
#pragma unroll
        for (int q = 0; q < NX; ++q) {
          const int o = (RH * (int)rank + tid) * NX + q;
          INF[o] = (float)in[q];
          st_cluster_f32(inf_peer + o * 4, (float)in[q]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (lane == 0) {
          vote[2 * rank + warp] = m != 0;
          asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(map_cluster(s32(vote + 2 * rank + warp), peer)), "r"((uint32_t)(m != 0)) : "memory");
        }
      }
      cluster_sync_compute();
      const bool any = (vote[0] | vote[1] | vote[2] | vote[3]) != 0;
      if (!any) { cluster_sync_compute(); continue; }       // (vote[] / INF are rewritten by the next tile)

      // ---- layer 1 on the CUDA cores: this CTA's 128 units for the 64 rows of CTA cg ----
      {
        float w1[NX];
#pragma unroll
        for (int q = 0; q < NX; ++q) w1[q] = W1s[unit * NX + q];
        const float bias = vec[j];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t dv[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const int n = (cq & 1) * 32 + c * CW + i;
            const float* in = INF + (cg * RH + n) * NX;
            float acc = bias;
#pragma unroll
            for (int q = 0; q < NX; ++q) acc = fmaf(w1[q], in[q], acc);
            float d;
            const float a = gelu_f32(acc, d);
            dv[i] = __float_as_uint(d);
            x_store2(xh_dst, xl_dst, n, unit, a);
          }
          tmem_st16(tlane + D1COL + cq * 32 + c * CW, dv);
        }
        tmem_st_wait();
      }
      x_ready();

      // ---- tensor-core layers ----
      mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
      epilogue2<L_FWD2>(sm, tlane, xh_dst, xl_dst, unit, j, cq, warp, lane, want_grad);
      x_ready();

      mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
      epilogue2<L_FWD3>(sm, tlane, xh_dst, xl_dst, unit, j, cq, warp, lane, want_grad);
      if (want_grad) {
        x_ready();
        mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
        epilogue2<L_BWD3>(sm, tlane, xh_dst, xl_dst, unit, j, cq, warp, lane, true);
        x_ready();
        mbar_wait(bar_acc, aph); aph ^= 1; fence_after();
        epilogue2<L_BWD2>(sm, tlane, xh_dst, xl_dst, unit, j, cq, warp, lane, true);      // X = g1 (hi + lo) in the row owner
      }
      fence_before();
      bar_compute();                                        // YP of this CTA complete
      // partial network outputs of the peer's rows -> the peer
      if (tid < 4 * RH) {
        const int wq = tid >> 6, n = tid & 63;              // 4 lane groups x 64 rows
        st_cluster_f32(map_cluster(s32(YQ + wq * RH + n), peer), YP[wq * RP + RH * (int)peer + n]);
      }
      cluster_sync_compute();                               // X (g1), YQ of both CTAs complete

      // ---- last reverse layer on the CUDA cores, own rows: gin[n][i] = sum_k W1[k][i] g1[k][n] ----
      if (want_grad && tid < 4 * RH) {
        const int n = tid & (RH - 1), part = tid >> 6;
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
        for (int q = 0; q < NX; ++q) GIN[(part * RH + n) * NX + q] = g[q];
      }
      bar_compute();
      if (tid < RH && valid[tid]) {
        const int pr = RH * (int)rank + tid;                // row in the pair tile
        const double y = (double)w.b4[0] + (((double)YP[pr] + (double)YP[RP + pr]) + ((double)YP[2 * RP + pr] + (double)YP[3 * RP + pr])) +
                         (((double)YQ[tid] + (double)YQ[RH + tid]) + ((double)YQ[2 * RH + tid] + (double)YQ[3 * RH + tid]));
        double gin[NX], grad[NX];
#pragma unroll
        for (int q = 0; q < NX; ++q)
          gin[q] = want_grad ? (double)((GIN[tid * NX + q] + GIN[(RH + tid) * NX + q]) + (GIN[(2 * RH + tid) * NX + q] + GIN[(3 * RH + tid) * NX + q])) : 0.0;
        double in[NX], nrm;
        nn_input(P, (rows_mode == ROWS_FLAT) ? xsrc + (size_t)rowb[tid] * NX : xsrc + ((size_t)rowb[tid] * (N + 1) + rowk[tid]) * NX, in, &nrm);
        const double cval = nn_output(P, in, nrm, y, want_grad ? gin : nullptr, want_grad ? grad : nullptr);
        double* o = (rows_mode == ROWS_FLAT) ? out11 + (size_t)rowb[tid] * NN_OUT : out11 + ((size_t)rowb[tid] * (N + 1) + rowk[tid]) * NN_OUT;
        o[0] = cval;
        if (want_grad) {
#pragma unroll
          for (int q = 0; q < NX; ++q) o[1 + q] = grad[q];
        }
      }
      cluster_sync_compute();                               // tile buffers of both CTAs are free again
    }
  }

  // ---- teardown ----
  fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

// Host: the operands of mlp_tc_pack, laid out per CTA of the pair: stage (l, s) of CTA r = [hi|lo][chunk = (k % KC) / 4][row = m % 128][e = k % 4]
// for the units m in [128 r, 128 r + 128) and the k in [KC s, KC s + KC).
size_t mlp_tc2_packed_floats() { return (size_t)4 * NSTG * 2 * (WSTAGE / 4); }

void mlp_tc2_pack(const float* W2, const float* W3, float* out) {
  for (int l = 0; l < 4; ++l) {
    const float* W = (l == 0 || l == 3) ? W2 : W3;
    const bool transposed = l >= 2;
    for (int m = 0; m < HID; ++m)
      for (int k = 0; k < HID; ++k) {
        const float a = transposed ? W[(size_t)k * HID + m] : W[(size_t)m * HID + k];
        float hi, lo;
        split_tf32(a, hi, lo);
        const int s = k / KC, chunk = (k % KC) / 4, e = k % 4, rk = m / UH, row = m % UH;
        const size_t base = (((size_t)l * NSTG + s) * 2 + rk) * (WSTAGE / 4);
        const size_t o = (size_t)chunk * (WROWB / 4) + (size_t)row * 4 + e;
        out[base + o] = hi;
        out[base + (WHALF / 4) + o] = lo;
      }
  }
}

cudaError_t mlp_tc2_prepare() {
  // per device: called by smpc_create after cudaSetDevice (the attribute belongs to the device, not to the process)
  cudaError_t e = cudaFuncSetAttribute(mlp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM);
  return e;
}

void launch_mlp_tc2(const LaunchCtx& c, const smpc_problem_t* dP, const MlpTcWeights& w, int n_sm, int B, int N, int rows_mode, int n_flat,
                    const double* xsrc, const int32_t* r, const uint8_t* act, const uint8_t* need, double* out11, bool want_grad) {
  int n_rows = (rows_mode == ROWS_TERMINAL || rows_mode == ROWS_CAND) ? B : rows_mode == ROWS_ALL ? B * N : rows_mode == ROWS_RECEDING ? 2 * B : n_flat;
  if (rows_mode == ROWS_FLAT) B = n_flat;
  if (n_rows <= 0) return;
  const int n_tiles = (n_rows + RP - 1) / RP;
  int pairs = n_sm / 2;
  if (n_tiles < pairs) pairs = n_tiles;
  mlp_tc2_kernel<<<2 * pairs, TC2_THREADS, TC2_SMEM, c.stream>>>(dP, w, B, N, rows_mode, n_rows, xsrc, r, act, need, out11, want_grad ? 1 : 0);
  ++*c.launches;
}

}  // namespace smpc
