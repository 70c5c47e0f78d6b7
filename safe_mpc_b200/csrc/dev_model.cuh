// Rigid-body / geometry device functions of the batched RTI engine (sm_100a).
//
// Hand-derived recursions for the quantities the reference obtains from CasADi's AD of the adam expression
// graphs (reference env_model.py:80-95,131-151; utils.py:94-113; cost_definition.py:69-96):
//   * recursive Newton-Euler inverse dynamics of the fixed-base serial chain, tau = M(q) u + h(q, v),
//   * its analytic first derivatives d tau / d(q, v, u) by sparse tangent recursions (one direction at a time:
//     forward from the perturbed joint to the tip, backward to the base),
//   * world kinematics of points attached to bodies with geometric Jacobians z_j x (P - o_j) and the
//     second-order terms z_j x (z_k x (P - o_k)) needed by the exact cost Hessian,
//   * squared capsule segment distance (clamped closest-point parameters, utils.py:94-113) with a
//     hand-written reverse sweep for its gradient.
// Everything is SMPC_HD so that the same source can be compiled for the host by the emulation harness in
// tests/emu (kernel-logic tests without a GPU); the product only ever runs the device instantiation.
#pragma once
#include <math.h>

#include "../../include/safe_mpc_b200.h"

#if defined(__CUDACC__)
#define SMPC_HD __host__ __device__ __forceinline__
#else
#define SMPC_HD inline
#endif

namespace smpc {

constexpr int NQ = SMPC_NQ, NX = SMPC_NX, NU = SMPC_NU, NZ = SMPC_NZ, NPAIR = SMPC_NPAIR, REC = SMPC_REC;

struct V3 {
  double x, y, z;
};
SMPC_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
SMPC_HD V3 v3(const double* p) { return v3(p[0], p[1], p[2]); }
SMPC_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SMPC_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SMPC_HD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
SMPC_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SMPC_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct M3 {
  double m[9];   // row-major
};
SMPC_HD V3 mul(const M3& R, V3 a) {
  return v3(R.m[0] * a.x + R.m[1] * a.y + R.m[2] * a.z, R.m[3] * a.x + R.m[4] * a.y + R.m[5] * a.z,
            R.m[6] * a.x + R.m[7] * a.y + R.m[8] * a.z);
}
SMPC_HD V3 mulT(const M3& R, V3 a) {
  return v3(R.m[0] * a.x + R.m[3] * a.y + R.m[6] * a.z, R.m[1] * a.x + R.m[4] * a.y + R.m[7] * a.z,
            R.m[2] * a.x + R.m[5] * a.y + R.m[8] * a.z);
}
SMPC_HD M3 matmul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C.m[3 * r + c] = A.m[3 * r] * B.m[c] + A.m[3 * r + 1] * B.m[3 + c] + A.m[3 * r + 2] * B.m[6 + c];
  return C;
}

// rotation parent-body <- body i:  joint_R[i] * exp([axis]x q)
SMPC_HD M3 joint_rot(const smpc_problem_t& P, int i, double q) {
  const double* a = P.joint_axis[i];
  double s, c;
  sincos(q, &s, &c);
  const double oc = 1.0 - c;
  M3 E;
  E.m[0] = c + oc * (a[0] * a[0]);        E.m[1] = oc * (a[0] * a[1]) - s * a[2]; E.m[2] = oc * (a[0] * a[2]) + s * a[1];
  E.m[3] = oc * (a[1] * a[0]) + s * a[2]; E.m[4] = c + oc * (a[1] * a[1]);        E.m[5] = oc * (a[1] * a[2]) - s * a[0];
  E.m[6] = oc * (a[2] * a[0]) - s * a[1]; E.m[7] = oc * (a[2] * a[1]) + s * a[0]; E.m[8] = c + oc * (a[2] * a[2]);
  M3 F;
#pragma unroll
  for (int k = 0; k < 9; ++k) F.m[k] = P.joint_R[i][k];
  return matmul(F, E);
}

SMPC_HD V3 inertia_mul(const double* I10, V3 w) {   // I10 = m, c[3], Ixx,Iyy,Izz,Ixy,Iyz,Ixz
  const double Ixx = I10[4], Iyy = I10[5], Izz = I10[6], Ixy = I10[7], Iyz = I10[8], Ixz = I10[9];
  return v3(Ixx * w.x + Ixy * w.y + Ixz * w.z, Ixy * w.x + Iyy * w.y + Iyz * w.z, Ixz * w.x + Iyz * w.y + Izz * w.z);
}

// ------------------------------------------------------------------------------------------------ RNEA
// Results of the nominal pass that the tangent passes read.  Two homes: thread-private (`Rnea`: registers / local memory -- host
// emulation, plant kernel, thread-per-stage linearisation) or shared memory (`RneaSm`: field f of this lane at p[32 f] -- cooperative
// linearisation kernel, where several warps read the state of the same 32 problems and index it with run-time joint numbers).
struct Rnea {
  M3 R[NQ];                 // parent <- body
  V3 w[NQ], wd[NQ], vd[NQ]; // body angular velocity / acceleration, linear acceleration of the frame origin
  V3 f[NQ], n[NQ];          // accumulated wrench transmitted through joint i, body frame i
};
SMPC_HD M3 get_R(const Rnea& S, int i) { return S.R[i]; }
SMPC_HD V3 get_w(const Rnea& S, int i) { return S.w[i]; }
SMPC_HD V3 get_wd(const Rnea& S, int i) { return S.wd[i]; }
SMPC_HD V3 get_vd(const Rnea& S, int i) { return S.vd[i]; }
SMPC_HD V3 get_f(const Rnea& S, int i) { return S.f[i]; }
SMPC_HD V3 get_n(const Rnea& S, int i) { return S.n[i]; }
SMPC_HD void set_R(Rnea& S, int i, const M3& r) { S.R[i] = r; }
SMPC_HD void set_w(Rnea& S, int i, V3 a) { S.w[i] = a; }
SMPC_HD void set_wd(Rnea& S, int i, V3 a) { S.wd[i] = a; }
SMPC_HD void set_vd(Rnea& S, int i, V3 a) { S.vd[i] = a; }
SMPC_HD void set_f(Rnea& S, int i, V3 a) { S.f[i] = a; }
SMPC_HD void set_n(Rnea& S, int i, V3 a) { S.n[i] = a; }

enum { SM_LANES = 32 };
enum { RS_R = 0, RS_W = 9 * NQ, RS_WD = RS_W + 3 * NQ, RS_VD = RS_WD + 3 * NQ, RS_F = RS_VD + 3 * NQ, RS_N = RS_F + 3 * NQ, RS_SIZE = RS_N + 3 * NQ };
struct RneaSm {
  double* p;                // shared memory, lane offset applied
};
SMPC_HD V3 sm_ld3(const double* p, int f) { return v3(p[(f)*SM_LANES], p[(f + 1) * SM_LANES], p[(f + 2) * SM_LANES]); }
SMPC_HD void sm_st3(double* p, int f, V3 a) { p[(f)*SM_LANES] = a.x; p[(f + 1) * SM_LANES] = a.y; p[(f + 2) * SM_LANES] = a.z; }
SMPC_HD M3 sm_ld9(const double* p, int f) {
  M3 r;
#pragma unroll
  for (int e = 0; e < 9; ++e) r.m[e] = p[(f + e) * SM_LANES];
  return r;
}
SMPC_HD void sm_st9(double* p, int f, const M3& r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) p[(f + e) * SM_LANES] = r.m[e];
}
SMPC_HD M3 get_R(const RneaSm& S, int i) { return sm_ld9(S.p, RS_R + 9 * i); }
SMPC_HD V3 get_w(const RneaSm& S, int i) { return sm_ld3(S.p, RS_W + 3 * i); }
SMPC_HD V3 get_wd(const RneaSm& S, int i) { return sm_ld3(S.p, RS_WD + 3 * i); }
SMPC_HD V3 get_vd(const RneaSm& S, int i) { return sm_ld3(S.p, RS_VD + 3 * i); }
SMPC_HD V3 get_f(const RneaSm& S, int i) { return sm_ld3(S.p, RS_F + 3 * i); }
SMPC_HD V3 get_n(const RneaSm& S, int i) { return sm_ld3(S.p, RS_N + 3 * i); }
SMPC_HD void set_R(RneaSm& S, int i, const M3& r) { sm_st9(S.p, RS_R + 9 * i, r); }
SMPC_HD void set_w(RneaSm& S, int i, V3 a) { sm_st3(S.p, RS_W + 3 * i, a); }
SMPC_HD void set_wd(RneaSm& S, int i, V3 a) { sm_st3(S.p, RS_WD + 3 * i, a); }
SMPC_HD void set_vd(RneaSm& S, int i, V3 a) { sm_st3(S.p, RS_VD + 3 * i, a); }
SMPC_HD void set_f(RneaSm& S, int i, V3 a) { sm_st3(S.p, RS_F + 3 * i, a); }
SMPC_HD void set_n(RneaSm& S, int i, V3 a) { sm_st3(S.p, RS_N + 3 * i, a); }

// tau = M(q) a + h(q, v) of the chain with inertial parameters `I` ([NQ][10]); fills `S` for the tangent passes
template <class RS>
SMPC_HD void rnea(const smpc_problem_t& P, const double (*I)[10], const double* q, const double* v, const double* a,
                  RS& S, double* tau) {
  V3 w = v3(0, 0, 0), wd = v3(0, 0, 0), vd = v3(-P.gravity[0], -P.gravity[1], -P.gravity[2]);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const M3 Ri = joint_rot(P, i, q[i]);
    set_R(S, i, Ri);
    const V3 ax = v3(P.joint_axis[i]), p = v3(P.joint_p[i]);
    const V3 acc = vd + cross(wd, p) + cross(w, cross(w, p));
    vd = mulT(Ri, acc);
    const V3 av = v[i] * ax;
    w = mulT(Ri, w) + av;
    wd = mulT(Ri, wd) + a[i] * ax + cross(w, av);
    set_w(S, i, w); set_wd(S, i, wd); set_vd(S, i, vd);
    const V3 c = v3(&I[i][1]);
    const V3 F = I[i][0] * (vd + cross(wd, c) + cross(w, cross(w, c)));
    const V3 Nn = inertia_mul(I[i], wd) + cross(w, inertia_mul(I[i], w)) + cross(c, F);
    set_f(S, i, F); set_n(S, i, Nn);
  }
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {
    const V3 ni = get_n(S, i);
    tau[i] = dot(v3(P.joint_axis[i]), ni);
    if (i > 0) {
      const M3 Ri = get_R(S, i);
      const V3 fp = mul(Ri, get_f(S, i));
      set_f(S, i - 1, get_f(S, i - 1) + fp);
      set_n(S, i - 1, get_n(S, i - 1) + mul(Ri, ni) + cross(v3(P.joint_p[i]), fp));
    }
  }
}

enum { TAN_Q = 0, TAN_V = 1, TAN_U = 2 };

// d tau / d (q_j | v_j | u_j) given the nominal pass `S` (v: joint velocities, u: joint accelerations)
template <int MODE, class RS>
SMPC_HD void rnea_tangent(const smpc_problem_t& P, const double (*I)[10], const RS& S, const double* v, const double* u,
                          int j, double* dtau) {
  V3 dF[NQ], dN[NQ];
  V3 dw, dwd, dvd;
  const V3 aj = v3(P.joint_axis[j]);
  if (MODE == TAN_Q) {
    const V3 wj = get_w(S, j);
    dw = cross(wj, aj);
    const V3 rwd = get_wd(S, j) - u[j] * aj - cross(wj, v[j] * aj);   // R_j^T wd_{j-1}
    dwd = cross(rwd, aj) + cross(dw, v[j] * aj);
    dvd = cross(get_vd(S, j), aj);
  } else if (MODE == TAN_V) {
    dw = aj; dwd = cross(get_w(S, j), aj); dvd = v3(0, 0, 0);
  } else {
    dw = v3(0, 0, 0); dwd = aj; dvd = v3(0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    if (i < j) { dF[i] = v3(0, 0, 0); dN[i] = v3(0, 0, 0); continue; }
    if (i > j) {
      const V3 p = v3(P.joint_p[i]);
      const V3 wp = get_w(S, i - 1);
      const M3 Ri = get_R(S, i);
      V3 t = dvd + cross(dwd, p);
      if (MODE != TAN_U) t = t + cross(dw, cross(wp, p)) + cross(wp, cross(dw, p));
      dvd = mulT(Ri, t);
      dwd = mulT(Ri, dwd);
      if (MODE != TAN_U) { dw = mulT(Ri, dw); dwd = dwd + cross(dw, v[i] * v3(P.joint_axis[i])); }
    }
    const V3 c = v3(&I[i][1]);
    const V3 w = get_w(S, i);
    V3 t = dvd + cross(dwd, c);
    if (MODE != TAN_U) t = t + cross(dw, cross(w, c)) + cross(w, cross(dw, c));
    dF[i] = I[i][0] * t;
    V3 nn = inertia_mul(I[i], dwd) + cross(c, dF[i]);
    if (MODE != TAN_U) nn = nn + cross(dw, inertia_mul(I[i], w)) + cross(w, inertia_mul(I[i], dw));
    dN[i] = nn;
  }
  V3 cf = v3(0, 0, 0), cn = v3(0, 0, 0);
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {
    V3 fi = dF[i] + cf, ni = dN[i] + cn;
    dtau[i] = dot(v3(P.joint_axis[i]), ni);
    if (i > 0) {
      if (MODE == TAN_Q && i == j) {
        // nominal accumulated wrench of body j is S.f[j], S.n[j]
        fi = fi + cross(aj, get_f(S, j));
        ni = ni + cross(aj, get_n(S, j));
      }
      const M3 Ri = get_R(S, i);
      cf = mul(Ri, fi);
      cn = mul(Ri, ni) + cross(v3(P.joint_p[i]), cf);
    }
  }
}

// ------------------------------------------------------------------------------------------ kinematics
struct Fk {
  M3 Rw[NQ];      // world <- body
  V3 o[NQ];       // body-frame origins (= joint origins) in the world
  V3 z[NQ];       // joint axes in the world
};
enum { FK_RW = 0, FK_O = 9 * NQ, FK_Z = FK_O + 3 * NQ, FK_SIZE = FK_Z + 3 * NQ };
struct FkSm {
  double* p;      // shared memory, lane offset applied (same layout convention as RneaSm)
};
SMPC_HD V3 get_o(const Fk& K, int i) { return K.o[i]; }
SMPC_HD V3 get_z(const Fk& K, int i) { return K.z[i]; }
SMPC_HD void set_fk(Fk& K, int i, const M3& Rc, V3 oc, V3 z) { K.Rw[i] = Rc; K.o[i] = oc; K.z[i] = z; }
SMPC_HD V3 get_o(const FkSm& K, int i) { return sm_ld3(K.p, FK_O + 3 * i); }
SMPC_HD V3 get_z(const FkSm& K, int i) { return sm_ld3(K.p, FK_Z + 3 * i); }
SMPC_HD void set_fk(FkSm& K, int i, const M3& Rc, V3 oc, V3 z) { sm_st9(K.p, FK_RW + 9 * i, Rc); sm_st3(K.p, FK_O + 3 * i, oc); sm_st3(K.p, FK_Z + 3 * i, z); }

template <class KS>
SMPC_HD void fk(const smpc_problem_t& P, const double* q, KS& K) {
  M3 Rc;
  Rc.m[0] = 1; Rc.m[1] = 0; Rc.m[2] = 0; Rc.m[3] = 0; Rc.m[4] = 1; Rc.m[5] = 0; Rc.m[6] = 0; Rc.m[7] = 0; Rc.m[8] = 1;
  V3 oc = v3(0, 0, 0);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    oc = oc + mul(Rc, v3(P.joint_p[i]));
    Rc = matmul(Rc, joint_rot(P, i, q[i]));
    set_fk(K, i, Rc, oc, mul(Rc, v3(P.joint_axis[i])));
  }
}

// world position of point p: thread-private state is selected with compile-time indices (a run-time index would send the whole
// struct to local memory), the shared-memory state is addressed directly
SMPC_HD V3 point_world(const smpc_problem_t& P, const Fk& K, int p) {
  const int b = P.point_body[p];
  const V3 l = v3(P.point_local[p]);
  if (b < 0) return l;
  V3 r = l;
#pragma unroll
  for (int i = 0; i < NQ; ++i) if (i == b) r = K.o[i] + mul(K.Rw[i], l);
  return r;
}
SMPC_HD V3 point_world(const smpc_problem_t& P, const FkSm& K, int p) {
  const int b = P.point_body[p];
  const V3 l = v3(P.point_local[p]);
  if (b < 0) return l;
  return get_o(K, b) + mul(sm_ld9(K.p, FK_RW + 9 * b), l);
}

// J[:, j] = z_j x (Pw - o_j) for j <= body, else 0
template <class KS>
SMPC_HD void point_jacobian(const smpc_problem_t& P, const KS& K, int p, V3 Pw, V3* J) {
  const int b = P.point_body[p];
#pragma unroll
  for (int j = 0; j < NQ; ++j) J[j] = (j <= b) ? cross(get_z(K, j), Pw - get_o(K, j)) : v3(0, 0, 0);
}

// --------------------------------------------------------------------------- capsule segment distance
// d = |(B-A) t - (D-C) u - (C-A)|^2 with the reference's clamped closest-point parameters (utils.py:94-113);
// gA, gB = d(d)/dA, d(d)/dB by a reverse sweep through the three clamps.
SMPC_HD double clamp01(double y, bool& free) {
  if (y > 1.0) { free = false; return 1.0; }      // fmin(y,1) then fmax(.,0)
  if (y < 0.0) { free = false; return 0.0; }
  free = true;
  return y;
}
SMPC_HD double segment_dist_grad(V3 A, V3 B, V3 C, V3 D, V3* gA, V3* gB) {
  const V3 ab = B - A, dc = D - C, ca = C - A;
  const double R = dot(ab, dc), S1 = dot(ab, ca), D1 = dot(ab, ab), S2 = dot(ca, dc), D2 = dot(dc, dc);
  const double den = D1 * D2 - (R * R + 1e-5);
  const double num = S1 * D2 - S2 * R;
  const double y1 = num / den;
  bool f1, f2, f3;
  const double t1 = clamp01(y1, f1);
  const double y2 = (t1 * R - S2) / D2;
  const double u = clamp01(y2, f2);
  const double y3 = (u * R + S1) / D1;
  const double t = clamp01(y3, f3);
  const V3 r = t * ab - u * dc - ca;
  const double d = dot(r, r);
  if (gA) {
    const V3 rb = 2.0 * r;
    V3 abb = t * rb;
    V3 cab = -1.0 * rb;
    const double tb = dot(rb, ab);
    double ub = -dot(rb, dc);
    double Rb = 0.0, S1b = 0.0, D1b = 0.0, S2b = 0.0;
    if (f3) { const double yb = tb; ub += yb * R / D1; Rb += yb * u / D1; S1b += yb / D1; D1b -= yb * y3 / D1; }
    double t1b = 0.0;
    if (f2) { const double yb = ub; t1b = yb * R / D2; Rb += yb * t1 / D2; S2b -= yb / D2; }
    if (f1) {
      const double yb = t1b;
      const double numb = yb / den, denb = -yb * y1 / den;
      S1b += numb * D2; S2b -= numb * R; Rb -= numb * S2;
      D1b += denb * D2; Rb -= denb * 2.0 * R;
    }
    abb = abb + Rb * dc + S1b * ca + (2.0 * D1b) * ab;
    cab = cab + S1b * ab + S2b * dc;
    *gB = abb;
    *gA = -1.0 * (abb + cab);
  }
  return d;
}

// ----------------------------------------------------------------------------------- viability network
// psi(x) = [(q - mean)/std ; v/|v|] with v = qdot, v[0] += eps (safe_set.py:82-87)
SMPC_HD void nn_input(const smpc_problem_t& P, const double* x, double* in, double* nrm_out) {
  double v[NQ], n2 = 0.0;
#pragma unroll
  for (int i = 0; i < NQ; ++i) v[i] = x[NQ + i];
  v[0] += P.eps;
#pragma unroll
  for (int i = 0; i < NQ; ++i) n2 += v[i] * v[i];
  const double nrm = sqrt(n2);
#pragma unroll
  for (int i = 0; i < NQ; ++i) { in[i] = (x[i] - P.nn_mean[i]) / P.nn_std[i]; in[NQ + i] = v[i] / nrm; }
  *nrm_out = nrm;
}
// c = y (100 - alpha)/100 - |v| and dc/dx from the network output y and its input gradient g
SMPC_HD double nn_output(const smpc_problem_t& P, const double* in, double nrm, double y, const double* g, double* grad) {
  const double s = (100.0 - P.alpha) / 100.0;
  if (grad) {
    double gd = 0.0;
#pragma unroll
    for (int i = 0; i < NQ; ++i) gd += g[NQ + i] * in[NQ + i];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      grad[i] = s * g[i] / P.nn_std[i];
      grad[NQ + i] = s * (g[NQ + i] - gd * in[NQ + i]) / nrm - in[NQ + i];
    }
  }
  return y * s - nrm;
}

SMPC_HD double gelu_tanh(double x, double* d) {
  const double k0 = 0.7978845608028654, k1 = 0.044715;
  const double x2 = x * x;
  const double t = tanh(k0 * (x + k1 * x * x2));
  *d = 0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * k0 * (1.0 + 3.0 * k1 * x2);
  return 0.5 * x * (1.0 + t);
}

// ------------------------------------------------------------------------------------------ dynamics
SMPC_HD void f_disc(double dt, const double* x, const double* u, double* xn) {   // env_model.py:63-71
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    xn[i] = x[i] + dt * x[NQ + i] + 0.5 * dt * dt * u[i];
    xn[NQ + i] = x[NQ + i] + dt * u[i];
  }
}

SMPC_HD bool state_in_bounds(const smpc_problem_t& P, const double* x) {        // env_model.py:175-177
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NX; ++i) ok = ok && (x[i] >= P.x_min[i] - P.tol_x) && (x[i] <= P.x_max[i] + P.tol_x);
  return ok;
}

SMPC_HD void distances(const smpc_problem_t& P, const double* q, double* ee, double* dist) {
  Fk K;
  fk(P, q, K);
  if (ee) { const V3 e = point_world(P, K, 0); ee[0] = e.x; ee[1] = e.y; ee[2] = e.z; }
  if (dist)
    for (int p = 0; p < NPAIR; ++p)
      dist[p] = segment_dist_grad(point_world(P, K, P.pair_pa[p]), point_world(P, K, P.pair_pb[p]), v3(P.pair_C[p]), v3(P.pair_D[p]), nullptr, nullptr);
}

SMPC_HD bool collision_free(const smpc_problem_t& P, const double* x) {         // env_model.py:236-243
  double d[NPAIR];
  distances(P, x, nullptr, d);
  bool ok = true;
  for (int p = 0; p < NPAIR; ++p) ok = ok && (P.pair_lo_chk[p] <= d[p]) && (d[p] <= P.pair_hi + P.tol_obs);
  return ok;
}

// one stage of the linearisation -> stage record (viability row value/gradient are supplied by the MLP kernel);
// `rs` = stride between consecutive record fields (32 in the device layout [tile][stage][field][32 problems])
// R: element type of the record array (double, or float for the fp32-storage flavour of the QP solver: values are rounded on store)
// The pieces below are shared by the thread-per-stage form (linearize_stage: host emulation, SMPC_LIN=thread) and the cooperative
// kernel (kernels.cu: linearize_coop_kernel), which runs them on different warps of a CTA.
// -- cost rows: gradient, exact Hessian of the end-effector term (cost_definition.py:83-100), control weight, LM scalars
// ee_ref: end-effector reference of this stage (smpc_problem_t::ee_ref, or a row of the trajectory of smpc_set_ee_trajectory)
template <class R, class KS>
SMPC_HD void lin_cost(const smpc_problem_t& P, int k, const KS& K, const double* u, const double* ee_ref, R* rec, int rs) {
  const bool term = (k == P.N);
  const double s = term ? 1.0 : P.dt;
  double hu = 0.0;
  if (P.cost_type != SMPC_COST_ZERO) {
    const V3 Pw = point_world(P, K, 0);
    V3 J[NQ];
    point_jacobian(P, K, 0, Pw, J);
    const V3 e = Pw - v3(ee_ref);
    const bool ext = P.cost_type == SMPC_COST_EXT;
    const double wq = (ext ? 2.0 : 1.0) * P.q_weight * s;
    const int be = P.point_body[0];
    int o = 0;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      rec[(size_t)rs * (SMPC_REC_G + NU + i)] = wq * dot(J[i], e);
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double h = dot(J[i], J[j]);
        if (ext && i <= be) h += dot(e, cross(get_z(K, j), cross(get_z(K, i), Pw - get_o(K, i))));   // j <= i: d2P/dq_j dq_i
        rec[(size_t)rs * (SMPC_REC_HQQ + o++)] = wq * h;
      }
    }
    if (!term) {
      const double wr = (ext ? 2.0 : 1.0) * P.r_weight * s;
#pragma unroll
      for (int i = 0; i < NU; ++i) rec[(size_t)rs * (SMPC_REC_G + i)] = wr * u[i];
      hu = wr;
    }
  }
  const double lmk = P.lm * ((P.lm_scale_dt && !term) ? P.dt : 1.0);
  rec[(size_t)rs * (SMPC_REC_HU)] = term ? 0.0 : hu + lmk;
  rec[(size_t)rs * (SMPC_REC_HV)] = lmk;
  rec[(size_t)rs * (SMPC_REC_HQ)] = lmk;
}
// -- column `col` (0..14: u_j, q_j, v_j) of the torque Jacobian
template <int MODE, class R, class RS>
SMPC_HD void lin_tau_col(const smpc_problem_t& P, const RS& S, const double* v, const double* u, int j, R* rec, int rs) {
  double d[NQ];
  rnea_tangent<MODE>(P, P.inertial, S, v, u, j, d);
  const int col = MODE == TAN_U ? j : (MODE == TAN_Q ? NU + j : NU + NQ + j);
#pragma unroll
  for (int i = 0; i < NU; ++i) rec[(size_t)rs * (SMPC_REC_JTAU + i * 15 + col)] = d[i];
}
// -- capsule pair p: squared distance and its gradient
template <class R, class KS>
SMPC_HD void lin_pair(const smpc_problem_t& P, const KS& K, int p, R* rec, int rs) {
  const V3 A = point_world(P, K, P.pair_pa[p]), Bp = point_world(P, K, P.pair_pb[p]);
  V3 gA, gB, JA[NQ], JB[NQ];
  const double d = segment_dist_grad(A, Bp, v3(P.pair_C[p]), v3(P.pair_D[p]), &gA, &gB);
  point_jacobian(P, K, P.pair_pa[p], A, JA);
  point_jacobian(P, K, P.pair_pb[p], Bp, JB);
  rec[(size_t)rs * (SMPC_REC_DIST + p)] = d;
#pragma unroll
  for (int j = 0; j < NQ; ++j) rec[(size_t)rs * (SMPC_REC_JDIST + p * NQ + j)] = dot(gA, JA[j]) + dot(gB, JB[j]);
}
// -- guess, row counts, viability row, dynamics offset
template <class R>
SMPC_HD void lin_misc(const smpc_problem_t& P, int k, const double* x, const double* u, const double* xnext, bool has_nn, bool gate_on,
                      const double* nn11, R* rec, int rs) {
  const bool term = (k == P.N);
#pragma unroll
  for (int i = 0; i < NX; ++i) rec[(size_t)rs * (SMPC_REC_X + i)] = x[i];
  if (!term) {
#pragma unroll
    for (int i = 0; i < NU; ++i) rec[(size_t)rs * (SMPC_REC_U + i)] = u[i];
    rec[(size_t)rs * (SMPC_REC_NTAU)] = NU;
  }
  if (k > 0 || P.stage0_collision_rows) rec[(size_t)rs * (SMPC_REC_NDIST)] = NPAIR;
  rec[(size_t)rs * (SMPC_REC_SOFT)] = -1.0;
  if (has_nn) {
    rec[(size_t)rs * (SMPC_REC_NNROW)] = 1.0;
    if (gate_on) {
      rec[(size_t)rs * (SMPC_REC_NN)] = nn11[0];
#pragma unroll
      for (int i = 0; i < NX; ++i) rec[(size_t)rs * (SMPC_REC_JNN + i)] = nn11[1 + i];
    } else {
      rec[(size_t)rs * (SMPC_REC_NN)] = 5e5;
    }
    if (term && P.nn_terminal_soft) rec[(size_t)rs * (SMPC_REC_SOFT)] = P.slack_penalty_e;
  }
  if (!term) {
    double xn[NX];
    f_disc(P.dt, x, u, xn);
#pragma unroll
    for (int i = 0; i < NX; ++i) rec[(size_t)rs * (SMPC_REC_B + i)] = xn[i] - xnext[i];
  }
}

template <class R>
SMPC_HD void linearize_stage(const smpc_problem_t& P, int k, const double* x, const double* u, const double* xnext,
                             bool has_nn, bool gate_on, const double* nn11, R* rec, int rs = 1, const double* ee_ref = nullptr) {
  const int N = P.N;
  const bool term = (k == N);
  for (int i = 0; i < REC; ++i) rec[(size_t)rs * (i)] = 0.0;
  const double* q = x;
  const double* v = x + NQ;
  Fk K;
  fk(P, q, K);
  lin_cost(P, k, K, u, ee_ref ? ee_ref : P.ee_ref, rec, rs);
  // ---- torque rows ----
  if (!term) {
    Rnea S;
    double tau[NQ];
    rnea(P, P.inertial, q, v, u, S, tau);
#pragma unroll
    for (int i = 0; i < NU; ++i) rec[(size_t)rs * (SMPC_REC_TAU + i)] = tau[i];
    for (int j = 0; j < NQ; ++j) {
      lin_tau_col<TAN_U>(P, S, v, u, j, rec, rs);
      lin_tau_col<TAN_Q>(P, S, v, u, j, rec, rs);
      lin_tau_col<TAN_V>(P, S, v, u, j, rec, rs);
    }
  }
  // ---- capsule rows ----
  if (k > 0 || P.stage0_collision_rows)
    for (int p = 0; p < NPAIR; ++p) lin_pair(P, K, p, rec, rs);
  lin_misc(P, k, x, u, xnext, has_nn, gate_on, nn11, rec, rs);
}

SMPC_HD bool stage_has_nn(const smpc_problem_t& P, int k) {
  return (P.nn_rows == SMPC_NN_TERMINAL && k == P.N) || ((P.nn_rows == SMPC_NN_RECEDING || P.nn_rows == SMPC_NN_EVERYWHERE || P.nn_rows == SMPC_NN_PARALLEL) && k >= 1);
}

// M(q) (row-major 5x5) and h(q, v) of the chain with inertial parameters I
SMPC_HD void mass_bias(const smpc_problem_t& P, const double (*I)[10], const double* x, double* M, double* h) {
  Rnea S;
  const double zero[NQ] = {0, 0, 0, 0, 0};
  rnea(P, I, x, x + NQ, zero, S, h);
  double d[NQ];
  for (int j = 0; j < NQ; ++j) {
    rnea_tangent<TAN_U>(P, I, S, x + NQ, zero, j, d);
    for (int i = 0; i < NQ; ++i) M[i * NQ + j] = d[i];
  }
}

// AdamModel.integrate (env_model.py:192-206): nominal torque + noise, clip, forward dynamics on the perturbed plant
SMPC_HD void plant_step(const smpc_problem_t& P, const double (*Iplant)[10], const double* noise, const double* x,
                        const double* u, double* xn, double* a) {
  Rnea S;
  double tau[NQ];
  rnea(P, P.inertial, x, x + NQ, u, S, tau);
  double M[NQ * NQ], h[NQ], L[NQ * NQ], y[NQ];
  mass_bias(P, Iplant, x, M, h);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    double t = tau[i] + noise[i];
    t = fmin(fmax(t, P.tau_min[i]), P.tau_max[i]);
    tau[i] = t - h[i];
  }
  for (int i = 0; i < NQ * NQ; ++i) L[i] = 0.0;
  for (int j = 0; j < NQ; ++j) {
    double d = M[j * NQ + j];
    for (int k = 0; k < j; ++k) d -= L[j * NQ + k] * L[j * NQ + k];
    L[j * NQ + j] = sqrt(d);
    for (int i = j + 1; i < NQ; ++i) {
      double s = M[i * NQ + j];
      for (int k = 0; k < j; ++k) s -= L[i * NQ + k] * L[j * NQ + k];
      L[i * NQ + j] = s / L[j * NQ + j];
    }
  }
  for (int i = 0; i < NQ; ++i) { double s = tau[i]; for (int k = 0; k < i; ++k) s -= L[i * NQ + k] * y[k]; y[i] = s / L[i * NQ + i]; }
  for (int i = NQ - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < NQ; ++k) s -= L[k * NQ + i] * a[k]; a[i] = s / L[i * NQ + i]; }
  f_disc(P.dt, x, a, xn);
}

// ------------------------------------------------------------------------ torque-input dynamics with sensitivities
// Extension row (f)4 of SURVEY.md section 8 (north_star: "RNEA/ABA inside an explicit RK4 integrator that emits state/control
// sensitivities").  The reference's OCP dynamics are the constant double integrator (env_model.py:58-71); its torque model
// tau = M(q) u + h(q, v) (env_model.py:42-43,80-83) is inverted here, on the controller model:
//     x' = f(x, tau) = [ v ; a ],   a = M(q)^-1 (tau - h(q, v))
// First derivatives from the identity  d ID / d(q, v) + M d a / d(q, v) = 0  at a = FD(q, v, tau)  (the inverse-dynamics tangents
// are the ones the linearisation already uses: rnea_tangent):
//     da/dtau = M^-1,   da/dq = -M^-1 dID/dq |_(q, v, a),   da/dv = -M^-1 dID/dv |_(q, v, a)
// All matrices row-major 5x5.
SMPC_HD void chol5(const double* M, double* L) {
  for (int i = 0; i < NQ * NQ; ++i) L[i] = 0.0;
  for (int j = 0; j < NQ; ++j) {
    double d = M[j * NQ + j];
    for (int k = 0; k < j; ++k) d -= L[j * NQ + k] * L[j * NQ + k];
    L[j * NQ + j] = sqrt(d);
    for (int i = j + 1; i < NQ; ++i) {
      double s = M[i * NQ + j];
      for (int k = 0; k < j; ++k) s -= L[i * NQ + k] * L[j * NQ + k];
      L[i * NQ + j] = s / L[j * NQ + j];
    }
  }
}
SMPC_HD void chol5_solve(const double* L, const double* b, double* y) {
  double t[NQ];
  for (int i = 0; i < NQ; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * NQ + k] * t[k]; t[i] = s / L[i * NQ + i]; }
  for (int i = NQ - 1; i >= 0; --i) { double s = t[i]; for (int k = i + 1; k < NQ; ++k) s -= L[k * NQ + i] * y[k]; y[i] = s / L[i * NQ + i]; }
}

template <bool SENS>
SMPC_HD void fd_sens(const smpc_problem_t& P, const double (*I)[10], const double* x, const double* tau, double* a, double* Aq,
                     double* Av, double* Mi) {
  Rnea S;
  double M[NQ * NQ], L[NQ * NQ], h[NQ], d[NQ], y[NQ];
  const double zero[NQ] = {0, 0, 0, 0, 0};
  rnea(P, I, x, x + NQ, zero, S, h);
  for (int j = 0; j < NQ; ++j) {
    rnea_tangent<TAN_U>(P, I, S, x + NQ, zero, j, d);
    for (int i = 0; i < NQ; ++i) M[i * NQ + j] = d[i];
  }
  chol5(M, L);
  for (int i = 0; i < NQ; ++i) d[i] = tau[i] - h[i];
  chol5_solve(L, d, a);
  if (!SENS) return;
  rnea(P, I, x, x + NQ, a, S, h);                      // nominal pass at the solved acceleration (h: scratch, = tau)
  for (int j = 0; j < NQ; ++j) {
    rnea_tangent<TAN_Q>(P, I, S, x + NQ, a, j, d);
    chol5_solve(L, d, y);
    for (int i = 0; i < NQ; ++i) Aq[i * NQ + j] = -y[i];
    rnea_tangent<TAN_V>(P, I, S, x + NQ, a, j, d);
    chol5_solve(L, d, y);
    for (int i = 0; i < NQ; ++i) Av[i * NQ + j] = -y[i];
    for (int i = 0; i < NQ; ++i) d[i] = i == j ? 1.0 : 0.0;
    chol5_solve(L, d, y);
    for (int i = 0; i < NQ; ++i) Mi[i * NQ + j] = y[i];
  }
}

// One explicit RK4 step of length dt of x' = f(x, tau) with the discrete-time sensitivities
//     A = d x_next / d x  [10][10],   B = d x_next / d tau  [10][5]      (row-major)
// propagated through the four stages: with S_i = d x_i / d(x, tau) = [I 0] + c_i dt dK_{i-1} (c = 0, 1/2, 1/2, 1) the stage slope
// K_i = f(x_i, tau) has  dK_i = [ S_i(v rows) ; Aq_i S_i(q rows) + Av_i S_i(v rows) + [0 | Minv_i] ],  and
//     [A B] = [I 0] + dt/6 (dK_1 + 2 dK_2 + 2 dK_3 + dK_4).   SENS = false: x_next only (no tangent passes).
template <bool SENS>
SMPC_HD void rk4_sens(const smpc_problem_t& P, const double (*I)[10], double dt, const double* x, const double* tau, double* xn,
                      double* A, double* B) {
  constexpr int NC = NX + NU;                           // sensitivity columns: x (10), tau (5)
  double dK[SENS ? NX * NC : 1], K[NX], ka[NX], xi[NX];
  double a[NQ], Aq[SENS ? NQ * NQ : 1], Av[SENS ? NQ * NQ : 1], Mi[SENS ? NQ * NQ : 1];
  if (SENS) {
    for (int i = 0; i < NX * NC; ++i) dK[i] = 0.0;
    for (int i = 0; i < NX * NX; ++i) A[i] = 0.0;
    for (int i = 0; i < NX * NU; ++i) B[i] = 0.0;
  }
  for (int i = 0; i < NX; ++i) { K[i] = 0.0; ka[i] = 0.0; }
  for (int st = 0; st < 4; ++st) {
    const double c = st == 0 ? 0.0 : (st == 3 ? 1.0 : 0.5), wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
    for (int i = 0; i < NX; ++i) xi[i] = x[i] + c * dt * K[i];
    fd_sens<SENS>(P, I, xi, tau, a, Aq, Av, Mi);
    for (int i = 0; i < NQ; ++i) { K[i] = xi[NQ + i]; K[NQ + i] = a[i]; }
    for (int i = 0; i < NX; ++i) ka[i] += wgt * K[i];
    if (SENS)
      for (int j = 0; j < NC; ++j) {                    // column j of dK_{st-1} -> column j of S -> column j of dK_st, in place
        double Sc[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) Sc[i] = (i == j ? 1.0 : 0.0) + c * dt * dK[i * NC + j];
        double* out = j < NX ? A + j : B + (j - NX);    // accumulated column j of [A B]
        const int ld = j < NX ? NX : NU;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          double s = j >= NX ? Mi[i * NQ + (j - NX)] : 0.0;
#pragma unroll
          for (int k = 0; k < NQ; ++k) s += Aq[i * NQ + k] * Sc[k] + Av[i * NQ + k] * Sc[NQ + k];
          dK[i * NC + j] = Sc[NQ + i];
          dK[(NQ + i) * NC + j] = s;
          out[i * ld] += wgt * Sc[NQ + i];
          out[(NQ + i) * ld] += wgt * s;
        }
      }
  }
  const double w6 = dt / 6.0;
  for (int i = 0; i < NX; ++i) {
    xn[i] = x[i] + w6 * ka[i];
    if (SENS) {
      for (int j = 0; j < NX; ++j) A[i * NX + j] = (i == j ? 1.0 : 0.0) + w6 * A[i * NX + j];
      for (int j = 0; j < NU; ++j) B[i * NU + j] = w6 * B[i * NU + j];
    }
  }
}

}  // namespace smpc
