// ORACLE -- test infrastructure only. Restatement of the reference's model functions, generic in the
// scalar type so that derivatives come from forward-mode AD (dual.hpp) as they do from CasADi upstream.
#pragma once
#include <cmath>
#include "dual.hpp"
#include "oracle.h"

namespace orc {

using std::sin; using std::cos; using std::sqrt;

template <class T> inline void cross3(const T* a, const T* b, T* c) {
  T c0 = a[1] * b[2] - a[2] * b[1];
  T c1 = a[2] * b[0] - a[0] * b[2];
  T c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}
template <class T> inline void crossd(const T* a, const double* b, T* c) {
  T c0 = a[1] * b[2] - a[2] * b[1];
  T c1 = a[2] * b[0] - a[0] * b[2];
  T c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}
template <class T> inline void dcross(const double* a, const T* b, T* c) {
  T c0 = a[1] * b[2] - a[2] * b[1];
  T c1 = a[2] * b[0] - a[0] * b[2];
  T c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}
// y = R x, yT = R^T x  (R row-major 3x3 of T)
template <class T> inline void rot(const T* R, const T* x, T* y) {
  T y0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  T y1 = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  T y2 = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
  y[0] = y0; y[1] = y1; y[2] = y2;
}
template <class T> inline void rotT(const T* R, const T* x, T* y) {
  T y0 = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
  T y1 = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
  T y2 = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
  y[0] = y0; y[1] = y1; y[2] = y2;
}

// rotation parent-body <- body i:  R = joint_R[i] * exp([axis]x q)   (URDF revolute joint)
template <class T> inline void joint_rot(const orc_problem_t& P, int i, const T& q, T* R) {
  const double* a = P.joint_axis[i];
  T s = sin(q), c = cos(q), oc = T(1.0) - c;
  T E[9];
  E[0] = c + oc * (a[0] * a[0]);        E[1] = oc * (a[0] * a[1]) - s * a[2]; E[2] = oc * (a[0] * a[2]) + s * a[1];
  E[3] = oc * (a[1] * a[0]) + s * a[2]; E[4] = c + oc * (a[1] * a[1]);        E[5] = oc * (a[1] * a[2]) - s * a[0];
  E[6] = oc * (a[2] * a[0]) - s * a[1]; E[7] = oc * (a[2] * a[1]) + s * a[0]; E[8] = c + oc * (a[2] * a[2]);
  const double* F = P.joint_R[i];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k)
      R[3 * r + k] = F[3 * r + 0] * E[0 + k] + F[3 * r + 1] * E[3 + k] + F[3 * r + 2] * E[6 + k];
}

// Recursive Newton-Euler inverse dynamics of the fixed-base chain:
//   tau = M(q) a + h(q, v)   with h = Coriolis/centrifugal + gravity
// This is what the reference evaluates as mass(H_b,q)[6:,6:] @ u + bias(H_b,q,0,v)[6:] with H_b = I
// (reference env_model.py:80-83, adam CRBA/RNEA on a floating-base model whose base is at rest).
template <class T>
void rnea(const orc_problem_t& P, const double inertial[][10], const T* q, const T* v, const T* a, T* tau) {
  constexpr int n = ORC_NQ;
  T R[n][9], F[n][3], Nn[n][3];
  T w[3] = {T(0.0), T(0.0), T(0.0)}, wd[3] = {T(0.0), T(0.0), T(0.0)};
  T vd[3] = {T(-P.gravity[0]), T(-P.gravity[1]), T(-P.gravity[2])};
  for (int i = 0; i < n; ++i) {
    joint_rot(P, i, q[i], R[i]);
    const double* ax = P.joint_axis[i];
    const double* p = P.joint_p[i];
    // linear acceleration of the new frame origin, parent quantities
    T t1[3], t2[3], t3[3];
    crossd(wd, p, t1);
    crossd(w, p, t2);
    cross3(w, t2, t3);
    T acc[3] = {vd[0] + t1[0] + t3[0], vd[1] + t1[1] + t3[1], vd[2] + t1[2] + t3[2]};
    rotT(R[i], acc, vd);
    T wl[3], wdl[3];
    rotT(R[i], w, wl);
    rotT(R[i], wd, wdl);
    T av[3] = {ax[0] * v[i], ax[1] * v[i], ax[2] * v[i]};
    for (int k = 0; k < 3; ++k) w[k] = wl[k] + av[k];
    T wxav[3];
    cross3(w, av, wxav);
    for (int k = 0; k < 3; ++k) wd[k] = wdl[k] + ax[k] * a[i] + wxav[k];
    // body wrench about the CoM
    const double m = inertial[i][0];
    const double* c = &inertial[i][1];
    const double Ixx = inertial[i][4], Iyy = inertial[i][5], Izz = inertial[i][6];
    const double Ixy = inertial[i][7], Iyz = inertial[i][8], Ixz = inertial[i][9];
    T wc[3], wwc[3], wdc[3];
    crossd(w, c, wc);
    cross3(w, wc, wwc);
    crossd(wd, c, wdc);
    for (int k = 0; k < 3; ++k) F[i][k] = m * (vd[k] + wdc[k] + wwc[k]);
    T Iw[3] = {Ixx * w[0] + Ixy * w[1] + Ixz * w[2], Ixy * w[0] + Iyy * w[1] + Iyz * w[2], Ixz * w[0] + Iyz * w[1] + Izz * w[2]};
    T Iwd[3] = {Ixx * wd[0] + Ixy * wd[1] + Ixz * wd[2], Ixy * wd[0] + Iyy * wd[1] + Iyz * wd[2], Ixz * wd[0] + Iyz * wd[1] + Izz * wd[2]};
    T wIw[3], cF[3];
    cross3(w, Iw, wIw);
    dcross(c, F[i], cF);
    for (int k = 0; k < 3; ++k) Nn[i][k] = Iwd[k] + wIw[k] + cF[k];
  }
  T f[3] = {T(0.0), T(0.0), T(0.0)}, nn[3] = {T(0.0), T(0.0), T(0.0)};
  for (int i = n - 1; i >= 0; --i) {
    T fi[3], ni[3];
    for (int k = 0; k < 3; ++k) { fi[k] = F[i][k] + f[k]; ni[k] = Nn[i][k] + nn[k]; }
    const double* ax = P.joint_axis[i];
    tau[i] = ax[0] * ni[0] + ax[1] * ni[1] + ax[2] * ni[2];
    T fp[3], np_[3], pf[3];
    rot(R[i], fi, fp);
    rot(R[i], ni, np_);
    dcross(P.joint_p[i], fp, pf);
    for (int k = 0; k < 3; ++k) { f[k] = fp[k]; nn[k] = np_[k] + pf[k]; }
  }
}

// world positions of the moving points (point 0 = end effector):  T_link(q)[:3,3] + T_link(q)[:3,:3] @ local
// (reference env_model.py:91-95 for the EE, :144-147 for capsule end points)
template <class T>
void fk_points(const orc_problem_t& P, const T* q, T pts[][3]) {
  constexpr int n = ORC_NQ;
  T Rw[n][9], ow[n][3];
  T Rc[9] = {T(1.0), T(0.0), T(0.0), T(0.0), T(1.0), T(0.0), T(0.0), T(0.0), T(1.0)};
  T oc[3] = {T(0.0), T(0.0), T(0.0)};
  for (int i = 0; i < n; ++i) {
    T pj[3] = {T(P.joint_p[i][0]), T(P.joint_p[i][1]), T(P.joint_p[i][2])}, rp[3];
    rot(Rc, pj, rp);
    for (int k = 0; k < 3; ++k) oc[k] = oc[k] + rp[k];
    T Rl[9], Rn[9];
    joint_rot(P, i, q[i], Rl);
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k)
        Rn[3 * r + k] = Rc[3 * r + 0] * Rl[0 + k] + Rc[3 * r + 1] * Rl[3 + k] + Rc[3 * r + 2] * Rl[6 + k];
    for (int k = 0; k < 9; ++k) { Rc[k] = Rn[k]; Rw[i][k] = Rn[k]; }
    for (int k = 0; k < 3; ++k) ow[i][k] = oc[k];
  }
  for (int p = 0; p < P.n_points; ++p) {
    int b = P.point_body[p];
    const double* l = P.point_local[p];
    if (b < 0) { for (int k = 0; k < 3; ++k) pts[p][k] = T(l[k]); continue; }
    T lp[3] = {T(l[0]), T(l[1]), T(l[2])}, rl[3];
    rot(Rw[b], lp, rl);
    for (int k = 0; k < 3; ++k) pts[p][k] = ow[b][k] + rl[k];
  }
}

// squared distance between segments AB (moving) and CD (fixed), reference utils.py:94-113 verbatim in
// structure: clamped closest-point parameters with the 1e-5 regulariser in the denominator.
template <class T>
T segment_dist(const T* A, const T* B, const double* C, const double* D) {
  T ab[3], ca[3];
  double dc[3];
  for (int k = 0; k < 3; ++k) { ab[k] = B[k] - A[k]; dc[k] = D[k] - C[k]; ca[k] = C[k] - A[k]; }
  T R = ab[0] * dc[0] + ab[1] * dc[1] + ab[2] * dc[2];
  T S1 = ab[0] * ca[0] + ab[1] * ca[1] + ab[2] * ca[2];
  T D1 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  T S2 = ca[0] * dc[0] + ca[1] * dc[1] + ca[2] * dc[2];
  double D2 = dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2];
  T t = (S1 * D2 - S2 * R) / (D1 * D2 - (R * R + 1e-5));
  t = fmax(fmin(t, 1.0), 0.0);
  T u = (t * R - S2) / D2;
  u = fmax(fmin(u, 1.0), 0.0);
  t = (u * R + S1) / D1;
  t = fmax(fmin(t, 1.0), 0.0);
  T r[3];
  for (int k = 0; k < 3; ++k) r[k] = ab[k] * t - u * dc[k] - ca[k];
  return r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
}

// ---- viability network (reference safe_set.py:26-43): 2nq -> H -> H -> H -> 1, GELU(tanh) ----
struct MlpView {
  const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4;
  explicit MlpView(const float* w) {
    const int H = ORC_HID, I = ORC_NX;
    W1 = w; b1 = W1 + H * I; W2 = b1 + H; b2 = W2 + H * H; W3 = b2 + H; b3 = W3 + H * H; W4 = b3 + H; b4 = W4 + H;
  }
};
template <class F>
inline F gelu_tanh(F x, F* dgelu) {
  const F k0 = (F)0.7978845608028654, k1 = (F)0.044715;
  F x2 = x * x;
  F inner = k0 * (x + k1 * x * x2);
  F t = std::tanh(inner);
  *dgelu = (F)0.5 * ((F)1.0 + t) + (F)0.5 * x * ((F)1.0 - t * t) * k0 * ((F)1.0 + (F)3.0 * k1 * x2);
  return (F)0.5 * x * ((F)1.0 + t);
}
// y = NN(in) and dy/din.  F = float: everything in fp32 like the libtorch evaluation behind L4CasADi
// (SURVEY Appendix C); F = double: the same fp32 weights, fp64 arithmetic (deterministic across summation orders;
// differs from the fp32 evaluation by less than the fp32 rounding of the reference itself).
template <class F>
inline F mlp_value_grad(const MlpView& W, const F* in, F* grad) {
  const int H = ORC_HID, I = ORC_NX;
  F a1[ORC_HID], a2[ORC_HID], a3[ORC_HID], d1[ORC_HID], d2[ORC_HID], d3[ORC_HID];
  for (int j = 0; j < H; ++j) {
    F s = W.b1[j];
    for (int k = 0; k < I; ++k) s += (F)W.W1[j * I + k] * in[k];
    a1[j] = gelu_tanh<F>(s, &d1[j]);
  }
  for (int j = 0; j < H; ++j) {
    F s = W.b2[j];
    for (int k = 0; k < H; ++k) s += (F)W.W2[j * H + k] * a1[k];
    a2[j] = gelu_tanh<F>(s, &d2[j]);
  }
  for (int j = 0; j < H; ++j) {
    F s = W.b3[j];
    for (int k = 0; k < H; ++k) s += (F)W.W3[j * H + k] * a2[k];
    a3[j] = gelu_tanh<F>(s, &d3[j]);
  }
  F y = W.b4[0];
  for (int k = 0; k < H; ++k) y += (F)W.W4[k] * a3[k];
  if (grad) {
    F g3[ORC_HID], g2[ORC_HID], g1[ORC_HID];
    for (int k = 0; k < H; ++k) g3[k] = (F)W.W4[k] * d3[k];
    for (int k = 0; k < H; ++k) {
      F s = 0;
      for (int j = 0; j < H; ++j) s += (F)W.W3[j * H + k] * g3[j];
      g2[k] = s * d2[k];
    }
    for (int k = 0; k < H; ++k) {
      F s = 0;
      for (int j = 0; j < H; ++j) s += (F)W.W2[j * H + k] * g2[j];
      g1[k] = s * d1[k];
    }
    for (int k = 0; k < I; ++k) {
      F s = 0;
      for (int j = 0; j < H; ++j) s += (F)W.W1[j * I + k] * g1[j];
      grad[k] = s;
    }
  }
  return y;
}

// c(x) = NN(psi(x)) * (100 - alpha)/100 - ||v||, psi = [(q-mean)/std ; v/||v||], v = qdot with v[0] += eps
// (reference safe_set.py:82-94).  grad = dc/dx (may be null).
template <class F>
inline double nn_constraint(const orc_problem_t& P, const double* x, double alpha, double* grad) {
  constexpr int n = ORC_NQ;
  double v[n], nrm2 = 0.0;
  for (int i = 0; i < n; ++i) { v[i] = x[n + i]; }
  v[0] += P.eps;
  for (int i = 0; i < n; ++i) nrm2 += v[i] * v[i];
  double nrm = std::sqrt(nrm2);
  F in[ORC_NX], g[ORC_NX];
  double dir[n];
  for (int i = 0; i < n; ++i) {
    in[i] = (F)((x[i] - P.nn_mean[i]) / P.nn_std[i]);
    dir[i] = v[i] / nrm;
    in[n + i] = (F)dir[i];
  }
  MlpView W(P.nn_weights);
  F y = mlp_value_grad<F>(W, in, grad ? g : nullptr);
  double s = (100.0 - alpha) / 100.0;
  if (grad) {
    double gd = 0.0;
    for (int i = 0; i < n; ++i) gd += (double)g[n + i] * dir[i];
    for (int i = 0; i < n; ++i) {
      grad[i] = s * (double)g[i] / P.nn_std[i];
      grad[n + i] = s * ((double)g[n + i] - gd * dir[i]) / nrm - dir[i];
    }
  }
  return (double)y * s - nrm;
}

}  // namespace orc
