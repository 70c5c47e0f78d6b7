// Warp-per-problem Riccati sweeps (device only; included by qp.cu inside its anonymous namespace, once per storage flavour).
//
// The lane-per-problem sweeps of qp.cu take the same time for 1 or 10 000 active problems (46 dependent stages of a ~2 000-
// instruction stream); in the last iterations of a solve a handful of problems are left -- and a problem that runs into
// qp_max_iter holds the whole batch for hundreds of iterations.  When few problems iterate, the lanes of a warp own the rows /
// entries of ONE problem instead.  Every entry is computed with the expressions and the summation order of qs_ric1 / qs_ric2
// (qp_split.cuh), so a problem gets bit-identical results whichever kernel serves it.
//
// What bounds a sweep here is the dependent chain of one problem, so the code is written for a short instruction stream:
//   * a stage block of one problem is ~150 words 256 B apart in the tile-interleaved layout; they are brought in by per-lane
//     cp.async copies into a ring of RT_DEPTH stages in shared memory, three stages ahead of their use: no lane ever waits for
//     global memory inside a stage (the first version fetched one stage ahead into registers: a stage of the vector sweep is
//     shorter than an L2 round trip);
//   * everything that depends only on the lane -- which row, which entries of the packed state block, which entries of P_{k+1}
//     they combine and with which of the double-integrator coefficients -- is worked out once in front of the stage loop;
//   * the LDL' elimination of the five control columns exchanges pivots and column entries by shuffles (no shared-memory
//     round trips and warp barriers inside the pivot loop); every lane forms the reciprocal of the pivot itself.
// Measured on B200 (profiles/r02_tail.md).
#pragma once

constexpr int RT_DEPTH = 4;                    // stages in flight per warp (ring slots)
constexpr int RT_S = 184;                      // staged fields of one stage (largest range: ric2 forward, 180)
constexpr int RT_NR = 6;                       // copies per lane and stage (32 x 6 >= RT_S)
constexpr int RT_WORK = 224;                   // fp64 work area per warp: ric1 P / p double buffer [2][65], multipliers [75], exchange [16];
                                               // ric2 y / pn / dx / du vectors [70], factor of P_0 [55], step statistics [96]
constexpr int RT_WARPS = 8;                    // warps (problems) per CTA of the tail kernels
constexpr size_t RT_WARP_BYTES = sizeof(double) * RT_WORK + sizeof(qs_real) * RT_DEPTH * RT_S;
constexpr size_t RT_SMEM = RT_WARP_BYTES * RT_WARPS;
static_assert(RT_WARP_BYTES % 16 == 0, "per-warp shared memory keeps 16-byte alignment");

__device__ __forceinline__ void rt_cp(qs_real* dst, const qs_real* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"((int)sizeof(qs_real)) : "memory");
}
__device__ __forceinline__ void rt_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void rt_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }
// fields [f0, f0 + nf) of one stage block of one problem (gblk: lane offset of the problem applied) -> slot[0 .. nf)
__device__ __forceinline__ void rt_fetch(qs_real* slot, const qs_real* gblk, int f0, int nf, int lane) {
#pragma unroll
  for (int u = 0; u < RT_NR; ++u) {
    const int f = lane + 32 * u;
    if (f < nf) rt_cp(slot + f, gblk + (size_t)(f0 + f) * TL);
  }
}
__device__ __forceinline__ int rt_trs(int i, int ti, int j, int tj) { return i >= j ? ti + j : tj + i; }   // ti = tri(i, 0), tj = tri(j, 0)

// ---------------------------------------------------------------------------------------------------------------
// ric1: backward factorisation with the affine gradient, stage-0 solve, forward substitution of the affine direction.
// one warp, one problem: tile / pl = slot of the problem, W = RT_WORK doubles, ring = RT_DEPTH x RT_S words of this warp
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rt_ric1(const smpc_problem_t& P, const QsBufs& q, int tile, int pl, int lane, double* W, qs_real* ring) {
  constexpr unsigned FULL = 0xffffffffu;
  const int N = q.N;
  double* Pb = W;                                            // [2][65]: P (55) and p (10) of stage k + 1 / k
  double* TX = W + 130;                                      // [15][5] elimination multipliers of the running stage
  double* X = W + 205;                                       // [16] y = P rb + p; later dx of the forward sweep
  double* pd = q.pd + qs_pb(tile, NPD, pl);
  const double dt = P.dt, a2 = 0.5 * P.dt * P.dt;
  const qs_real* gsb = q.sb + qs_blk(tile, N, 0, NSB, pl);   // stage blocks, lane offset of this problem applied
  const size_t sstride = (size_t)NSB * TL;

  // ---- lane roles, fixed for the whole sweep ----
  // (a) row `lane` (< 15) of the 15 x 5 panel: variable group gi (u, q, v), index ii inside the group; coefficients of q+ / v+
  const bool rowon = lane < 15;
  const int gi = lane < 5 ? 0 : (lane < 10 ? 1 : 2);
  const int ii = rowon ? lane - 5 * gi : 0;
  const double aq = gi == 0 ? a2 : (gi == 1 ? 1.0 : dt), av = gi == 0 ? dt : (gi == 1 ? 0.0 : 1.0);
  const double cA = aq * a2, cB = aq * dt, cC = av * a2, cD = av * dt;   // (aqi * aqc) ... (avi * avc) of qs_y for a control column
  const bool hasv = gi != 1;
  const int tii = tri(ii, 0), t5ii = tri(5 + ii, 0), tlane = tri(lane < 15 ? lane : 0, 0);
  // (b) entries e = lane, lane + 32 (< 55) of the packed state block: row r, column c (0..9), the four P_{k+1} entries of qs_y
  int eM[2], eR[2], eC[2], eA[2], eB[2], eCi[2], eD[2];
  double kA[2], kB[2], kC[2], kD[2];
  bool eon[2], fB[2], fC[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = lane + 32 * h;
    eon[h] = e < 55;
    int r = 0;
    while (tri(r + 1, 0) <= e && r < 9) ++r;
    const int c = eon[h] ? e - tri(r, 0) : 0;
    const int gr = r >= 5 ? 2 : 1, gc = c >= 5 ? 2 : 1, ir = r - 5 * (gr - 1), ic = c - 5 * (gc - 1);
    const double aqi = gr == 1 ? 1.0 : dt, avi = gr == 1 ? 0.0 : 1.0, aqc = gc == 1 ? 1.0 : dt, avc = gc == 1 ? 0.0 : 1.0;
    kA[h] = aqi * aqc; kB[h] = aqi * avc; kC[h] = avi * aqc; kD[h] = avi * avc;
    fB[h] = gc != 1; fC[h] = gr != 1;
    eA[h] = trs(ir, ic); eB[h] = tri(5 + ic, ir); eCi[h] = tri(5 + ir, ic); eD[h] = trs(5 + ir, 5 + ic);
    eM[h] = H_M + tri(5 + r, 5 + c); eR[h] = (5 + r) * 5; eC[h] = (5 + c) * 5;
  }

  int cur = 0;                                               // Pb[cur] = (P_{k+1}, p_{k+1})
#pragma unroll
  for (int d = 0; d < RT_DEPTH - 1; ++d) {
    const int kf = N - d;
    if (kf >= 0) rt_fetch(ring + (size_t)(kf % RT_DEPTH) * RT_S, gsb + (size_t)kf * sstride, B_M, B_LP - B_M, lane);
    rt_commit();
  }
  for (int k = N; k >= 0; --k) {
    {
      const int kf = k - (RT_DEPTH - 1);                     // its slot was that of stage k + 1: every lane is done with it (barrier below)
      if (kf >= 0) rt_fetch(ring + (size_t)(kf % RT_DEPTH) * RT_S, gsb + (size_t)kf * sstride, B_M, B_LP - B_M, lane);
      rt_commit();
    }
    rt_wait<RT_DEPTH - 1>();
    __syncwarp();
    const qs_real* S = ring + (size_t)(k % RT_DEPTH) * RT_S;  // M GA RB at S[field]
    const double* pc = Pb + cur * 65;
    double* pnw = Pb + (cur ^ 1) * 65;
    qs_real* fac = q.sb + qs_blk(tile, N, k, NSB, pl);
    // y = P rb + p  (lanes 0-9)
    if (k < N && lane < 10) {
      double s_ = 0.0;
#pragma unroll
      for (int j = 0; j < 10; ++j) s_ += pc[rt_trs(lane, tlane, j, tri(j, 0))] * S[H_RB + j];
      QF(fac, F_WV + lane) = s_;
      X[lane] = s_ + pc[55 + lane];
    }
    __syncwarp();
    // gradient and panel row of lane i < 15
    double g = 0.0, pan[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (rowon) {
      g = S[H_GA + lane];
      if (k < N) {
        if (lane < 5) g += a2 * X[lane] + dt * X[5 + lane];
        else if (lane < 10) g += X[lane - 5];
        else g += dt * X[lane - 10] + X[lane - 5];
      } else if (lane < 5) g = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (j <= lane) {
          double y = 0.0;
          if (k < N) {
            y += cA * pc[rt_trs(ii, tii, j, tri(j, 0))];
            y += cB * pc[tri(5 + j, 0) + ii];
            if (hasv) y += cC * pc[t5ii + j];
            if (hasv) y += cD * pc[(ii >= j ? t5ii + 5 + j : tri(5 + j, 0) + 5 + ii)];
          }
          pan[j] = S[H_M + tlane + j] + y;
        }
    }
    // LDL' elimination of the control columns: lane i owns row i; pivot, pivot-row gradient and the not yet scaled column entries
    // of the rows j+1..4 travel by shuffle
    double dd[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const double d = __shfl_sync(FULL, pan[j], j);
      const double gj = __shfl_sync(FULL, g, j);
      double col[5];
#pragma unroll
      for (int c = j + 1; c < 5; ++c) col[c] = __shfl_sync(FULL, pan[j], c);
      const double invd = d > 0.0 ? qs_rcp(d) : 0.0;
      dd[j] = d > 0.0 ? d : 0.0;
      if (lane > j && rowon) {
        const double t = pan[j] * invd;
        g -= t * gj;
#pragma unroll
        for (int c = j + 1; c < 5; ++c)
          if (c <= lane) pan[c] -= t * col[c];
        pan[j] = t;
      }
      if (lane == j) pan[j] = invd;
    }
    if (rowon) {
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (j <= lane) { TX[lane * 5 + j] = pan[j]; QF(fac, F_T + lane * 5 + j) = pan[j]; }
      QF(fac, F_LP + lane) = g;
      if (lane >= 5) pnw[55 + lane - 5] = g;                               // p_k
    }
    __syncwarp();
    // P_k = trailing block - T_x D T_x': two entries of the packed state block per lane
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (eon[h]) {
        double y = 0.0;
        if (k < N) {
          y += kA[h] * pc[eA[h]];
          if (fB[h]) y += kB[h] * pc[eB[h]];
          if (fC[h]) y += kC[h] * pc[eCi[h]];
          if (fC[h] && fB[h]) y += kD[h] * pc[eD[h]];
        }
        double v = S[eM[h]] + y;
        const double* Tr = TX + eR[h];
        const double* Tc = TX + eC[h];
#pragma unroll
        for (int j = 0; j < 5; ++j) v -= (Tr[j] * dd[j]) * Tc[j];
        pnw[lane + 32 * h] = v;
        QF(fac, F_P + lane + 32 * h) = v;
      }
    }
    __syncwarp();
    cur ^= 1;
  }
  rt_wait<0>();
  // stage 0: factorise P_0 (kept for ric2) and solve P_0 dx_0 = -p_0   (one lane; once per sweep)
  double* DX = X;                                                          // [10]
  if (lane == 0) {
    const double* p0 = Pb + cur * 65;
    double m[10][10], gg[10], dx[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      gg[i] = p0[55 + i];
#pragma unroll
      for (int c = 0; c <= i; ++c) m[i][c] = p0[tri(i, c)];
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const double d = m[j][j];
      const double invd = d > 0.0 ? qs_rcp(d) : 0.0;
      m[j][j] = invd;
#pragma unroll
      for (int i = 9; i > j; --i) {
        const double t = m[i][j] * invd;
        gg[i] -= t * gg[j];
#pragma unroll
        for (int c = j + 1; c <= i; ++c) m[i][c] -= t * m[c][j];
        m[i][j] = t;
      }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
      for (int c = 0; c <= i; ++c) QF(pd, D_T0 + tri(i, c)) = m[i][c];
#pragma unroll
    for (int i = 9; i >= 0; --i) {
      double acc = m[i][i] * gg[i];
#pragma unroll
      for (int c = i + 1; c < 10; ++c) acc += m[c][i] * dx[c];
      dx[i] = m[i][i] > 0.0 ? -acc : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) DX[i] = dx[i];
  }
  __threadfence_block();
  __syncwarp();
  // forward substitution (affine direction): lanes 0-4 own du, lanes 0-9 own dx; stages fetch RB LP T = [B_RB, B_WV)
  double* ZZ = TX;                                                         // [15] dz of the running stage
  double* DU = TX + 16;                                                    // [5]
#pragma unroll
  for (int d = 0; d < RT_DEPTH - 1; ++d) {
    if (d <= N) rt_fetch(ring + (size_t)(d % RT_DEPTH) * RT_S, gsb + (size_t)d * sstride, B_RB, B_WV - B_RB, lane);
    rt_commit();
  }
  for (int k = 0; k <= N; ++k) {
    {
      const int kf = k + RT_DEPTH - 1;
      if (kf <= N) rt_fetch(ring + (size_t)(kf % RT_DEPTH) * RT_S, gsb + (size_t)kf * sstride, B_RB, B_WV - B_RB, lane);
      rt_commit();
    }
    rt_wait<RT_DEPTH - 1>();
    __syncwarp();
    const qs_real* sb = ring + (size_t)(k % RT_DEPTH) * RT_S - B_RB;       // sb[f] valid for B_RB <= f < B_WV
    double du = 0.0;
    if (k < N && lane < 5) {
      double ws = (double)sb[F_T + lane * 5 + lane] * (double)sb[F_LP + lane];
#pragma unroll
      for (int r = 0; r < 10; ++r) ws += sb[F_T + (5 + r) * 5 + lane] * DX[r];
      du = -ws;
    }
    if (k < N) {
      // du[i] -= T[c][i] du[c] for c = 4 .. 1, i < c (the order of qs_ric1)
#pragma unroll
      for (int c = 4; c >= 1; --c) {
        const double duc = __shfl_sync(FULL, du, c);
        if (lane < c) du -= sb[F_T + c * 5 + lane] * duc;
      }
    }
    if (lane < 5) { ZZ[lane] = du; DU[lane] = du; }
    if (lane < 10) ZZ[5 + lane] = DX[lane];
    __syncwarp();
    double nx = 0.0;
    if (k < N && lane < 10) {
      if (lane < 5) nx = DX[lane] + dt * DX[5 + lane] + a2 * DU[lane] + sb[H_RB + lane];
      else nx = DX[lane] + dt * DU[lane - 5] + sb[H_RB + lane];
    }
    if (lane < 15) QF(q.st + qs_blk(tile, N, k, NIT, pl), I_Z + lane) = ZZ[lane];
    __syncwarp();
    if (k < N && lane < 10) DX[lane] = nx;
    __syncwarp();
  }
  rt_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------
// ric2: vector-only sweeps of the corrector direction (lanes 0-15) and of the pure-centering direction (lanes 16-31); row /
// entry index = lane & 15.  Expressions and summation order of qs_ric2.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rt_ric2(const smpc_problem_t& P, const QsBufs& q, int tile, int pl, int lane, double* W, qs_real* ring) {
  constexpr unsigned FULL = 0xffffffffu;
  const int N = q.N;
  const int32_t* pi = q.pi + qs_pb(tile, NPI, pl);
  const int v = lane >> 4, i = lane & 15;                   // direction, row
  const unsigned hb = lane & 16;                             // first lane of this half warp
  double* YV = W;                                            // [2][10] y
  double* PN = W + 20;                                       // [2][10]
  double* DXV = W + 40;                                      // [2][10]
  double* DUV = W + 60;                                      // [2][5]
  double* T0 = W + 70;                                       // [55] factor of P_0
  double* RS = W + 128;                                      // [3][32] step statistics of 32 stages
  double* pd = q.pd + qs_pb(tile, NPD, pl);
  const double dt = P.dt, a2 = 0.5 * P.dt * P.dt;
  const qs_real* gsb = q.sb + qs_blk(tile, N, 0, NSB, pl);
  const size_t sstride = (size_t)NSB * TL;
  const int n1 = B_P - B_GA, n2 = NSB - B_V1;
  // backward stage fields: [B_GA, B_P) at slot[f - B_GA], [B_V1, NSB) at slot[n1 + f - B_V1]
  auto fetch_b = [&](int k) {
    qs_real* slot = ring + (size_t)(k % RT_DEPTH) * RT_S;
    const qs_real* blk = gsb + (size_t)k * sstride;
#pragma unroll
    for (int u = 0; u < RT_NR; ++u) {
      const int f = lane + 32 * u;
      if (f < n1) rt_cp(slot + f, blk + (size_t)(B_GA + f) * TL);
      else if (f < n1 + n2) rt_cp(slot + f, blk + (size_t)(B_V1 + f - n1) * TL);
    }
  };
#pragma unroll
  for (int d = 0; d < RT_DEPTH - 1; ++d) {
    if (N - d >= 0) fetch_b(N - d);
    rt_commit();
  }

  // ---- sigma from the affine step statistics (sums over the stages in stage order, like qs_reduce_step) ----
  double sigmu;
  {
    double alpha = 1.0, s_lin = 0.0, s_quad = 0.0;
    for (int k0 = 0; k0 <= N; k0 += 32) {
      const int k = k0 + lane;
      if (k <= N) {
        const double* stp = q.stp + qs_blk(tile, N, k, NSTP, pl);
        RS[lane] = QF(stp, S_ALPHA); RS[32 + lane] = QF(stp, S_LIN); RS[64 + lane] = QF(stp, S_QUAD);
      }
      __syncwarp();
      if (lane == 0) {
        const int n = N + 1 - k0 < 32 ? N + 1 - k0 : 32;
        for (int j = 0; j < n; ++j) { alpha = fmin(alpha, RS[j]); s_lin += RS[32 + j]; s_quad += RS[64 + j]; }
      }
      __syncwarp();
    }
    double sm_ = 0.0;
    if (lane == 0) {
      const double mu = QF(pd, D_MU);
      const double mu_aff = mu + (alpha * s_lin + alpha * alpha * s_quad) / QF(pi, J_NC);
      double sigma = mu_aff / mu; sigma = sigma * sigma * sigma;
      sm_ = sigma * mu;
      QF(pd, D_MUAFF) = mu_aff; QF(pd, D_SIGMU) = sm_;
    }
    sigmu = __shfl_sync(FULL, sm_, 0);
  }
  if (i < 10) { PN[v * 10 + i] = 0.0; DXV[v * 10 + i] = 0.0; }
  for (int e = lane; e < 55; e += 32) T0[e] = QF(pd, D_T0 + e);
  for (int k = N; k >= 0; --k) {
    if (k - (RT_DEPTH - 1) >= 0) fetch_b(k - (RT_DEPTH - 1));
    rt_commit();
    rt_wait<RT_DEPTH - 1>();
    __syncwarp();
    const qs_real* S = ring + (size_t)(k % RT_DEPTH) * RT_S;
    const qs_real* sb = S - B_GA;                             // sb[f] valid for B_GA <= f < B_P
    const qs_real* vv = S + n1 - B_V1;                        // vv[f] valid for B_V1 <= f < NSB
    qs_real* fac = q.sb + qs_blk(tile, N, k, NSB, pl);
    if (k < N && i < 10) YV[v * 10 + i] = sb[F_WV + i] + PN[v * 10 + i];
    __syncwarp();
    double g = 0.0;
    if (i < 15) {
      const double ga = sb[H_GA + i], s2 = sigmu * vv[V_2 + i];
      g = v == 0 ? ga + vv[V_1 + i] - s2 : ga + 0.0 - s2;
      if (k < N) {
        const double* y = YV + v * 10;
        if (i < 5) g += a2 * y[i] + dt * y[5 + i];
        else if (i < 10) g += y[i - 5];
        else g += dt * y[i - 10] + y[i - 5];
      } else if (i < 5) g = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const double gj = __shfl_sync(FULL, g, hb | j);
      if (i > j && i < 15) g -= sb[F_T + i * 5 + j] * gj;
    }
    if (i < 15) QF(fac, (v == 0 ? F_LP : F_LP2) + i) = g;
    __syncwarp();                                             // y consumed by every lane
    if (i >= 5 && i < 15) PN[v * 10 + i - 5] = g;
    __syncwarp();
  }
  rt_wait<0>();
  // stage 0: P_0 dx_0 = -p_0 with the factor kept by ric1 (one lane per direction)
  if (i == 0) {
    double pn[10], dx[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) pn[r] = PN[v * 10 + r];
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
      for (int r = j + 1; r < 10; ++r) pn[r] -= T0[tri(r, j)] * pn[j];
#pragma unroll
    for (int r = 9; r >= 0; --r) {
      const double invd = T0[tri(r, r)];
      double acc = invd * pn[r];
#pragma unroll
      for (int c = r + 1; c < 10; ++c) acc += T0[tri(c, r)] * dx[c];
      dx[r] = invd > 0.0 ? -acc : 0.0;
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) DXV[v * 10 + r] = dx[r];
  }
  __threadfence_block();
  __syncwarp();
  // forward: stages fetch RB LP T WV P LP2 = [B_RB, B_V1) at slot[f - B_RB]
#pragma unroll
  for (int d = 0; d < RT_DEPTH - 1; ++d) {
    if (d <= N) rt_fetch(ring + (size_t)(d % RT_DEPTH) * RT_S, gsb + (size_t)d * sstride, B_RB, B_V1 - B_RB, lane);
    rt_commit();
  }
  for (int k = 0; k <= N; ++k) {
    {
      const int kf = k + RT_DEPTH - 1;
      if (kf <= N) rt_fetch(ring + (size_t)(kf % RT_DEPTH) * RT_S, gsb + (size_t)kf * sstride, B_RB, B_V1 - B_RB, lane);
      rt_commit();
    }
    rt_wait<RT_DEPTH - 1>();
    __syncwarp();
    const qs_real* sb = ring + (size_t)(k % RT_DEPTH) * RT_S - B_RB;
    const double* dx = DXV + v * 10;
    const int lp = v == 0 ? F_LP : F_LP2;
    qs_real* sto = v == 0 ? q.st + qs_blk(tile, N, k, NIT, pl) : q.st2 + qs_blk(tile, N, k, NS2, pl);
    // multiplier step of the link k-1 -> k:  dpi = P_k dx_k + p_k
    if (i < 10) {
      double s_ = 0.0;
      if (k > 0) {
        s_ = sb[lp + 5 + i];
#pragma unroll
        for (int j = 0; j < 10; ++j) s_ += sb[F_P + trs(i, j)] * dx[j];
      }
      QF(sto, I_PIM + i) = s_;
    }
    double du = 0.0;
    if (k < N && i < 5) {
      double ws = (double)sb[F_T + i * 5 + i] * (double)sb[lp + i];
#pragma unroll
      for (int r = 0; r < 10; ++r) ws += sb[F_T + (5 + r) * 5 + i] * dx[r];
      du = -ws;
    }
    if (k < N) {
#pragma unroll
      for (int c = 4; c >= 1; --c) {
        const double duc = __shfl_sync(FULL, du, hb | c);
        if (i < c) du -= sb[F_T + c * 5 + i] * duc;
      }
    }
    if (i < 5) { DUV[v * 5 + i] = du; QF(sto, I_Z + i) = du; }
    if (i < 10) QF(sto, I_Z + 5 + i) = dx[i];
    __syncwarp();
    double nx = 0.0;
    if (k < N && i < 10) {
      const double* duv = DUV + v * 5;
      if (i < 5) nx = dx[i] + dt * dx[5 + i] + a2 * duv[i] + sb[H_RB + i];
      else nx = dx[i] + dt * duv[i - 5] + sb[H_RB + i];
    }
    __syncwarp();
    if (k < N && i < 10) DXV[v * 10 + i] = nx;
    __syncwarp();
  }
  rt_wait<0>();
}
