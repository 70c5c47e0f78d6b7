// ORACLE -- test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// Forward-mode automatic-differentiation scalars. The reference obtains every derivative of the
// hot path from CasADi's algorithmic differentiation of the adam / utils.py expression graphs
// (reference env_model.py:80-95,144-150; utils.py:94-113; cost_definition.py:91-96); the oracle
// does the same thing -- AD of a plain restatement of those expressions -- so that it is
// independent of the hand-derived analytic derivatives in the CUDA kernels.
#pragma once
#include <cmath>

// Flop tally (tests only): with -DORC_COUNT_FLOPS every operation on an AD scalar adds its floating-point operation count
// (value part + derivative parts; a transcendental counts as one) to a thread-local counter, read through orc_flops_get().
#ifdef ORC_COUNT_FLOPS
namespace orc { extern thread_local unsigned long long g_flops; }
#define ORC_FLOP(n) (::orc::g_flops += (unsigned long long)(n))
#else
#define ORC_FLOP(n) ((void)0)
#endif

namespace orc {

// first-order dual number with N directions
template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  Dual(double c) : v(c) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  static Dual var(double c, int k) { Dual r(c); r.d[k] = 1.0; return r; }
};

template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; ORC_FLOP(1 + N); r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; ORC_FLOP(1 + N); r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; ORC_FLOP(1 + 3 * N); r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; ORC_FLOP(2 + 3 * N); double inv = 1.0 / b.v; r.v = a.v * inv;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; ORC_FLOP(1); r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; ORC_FLOP(1); r.v += b; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; ORC_FLOP(1); r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { return Dual<N>(b) - a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) {
  Dual<N> r; ORC_FLOP(1 + N); r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N> sin(const Dual<N>& a) {
  Dual<N> r; ORC_FLOP(2 + N); r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) {
  Dual<N> r; ORC_FLOP(2 + N); r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) {
  Dual<N> r; ORC_FLOP(2 + N); r.v = std::sqrt(a.v); double k = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = k * a.d[i]; return r; }
// CasADi fmin/fmax: value of the selected branch, derivative of the selected branch
template <int N> inline Dual<N> fmin(const Dual<N>& a, double b) { return (a.v <= b) ? a : Dual<N>(b); }
template <int N> inline Dual<N> fmax(const Dual<N>& a, double b) { return (a.v >= b) ? a : Dual<N>(b); }
inline double fmin(double a, double b) { return a <= b ? a : b; }
inline double fmax(double a, double b) { return a >= b ? a : b; }
inline double value(double a) { return a; }
template <int N> inline double value(const Dual<N>& a) { return a.v; }

// second-order forward-mode scalar in N variables: value, gradient, symmetric Hessian (upper, row-major)
template <int N>
struct Dual2 {
  static constexpr int NH = N * (N + 1) / 2;
  double v;
  double g[N];
  double h[NH];
  Dual2() : v(0.0) { for (int i = 0; i < N; ++i) g[i] = 0.0; for (int i = 0; i < NH; ++i) h[i] = 0.0; }
  Dual2(double c) : Dual2() { v = c; }
  static Dual2 var(double c, int k) { Dual2 r(c); r.g[k] = 1.0; return r; }
  static int idx(int i, int j) { if (i > j) { int t = i; i = j; j = t; } return i * N - i * (i - 1) / 2 + (j - i); }
};
template <int N> inline Dual2<N> operator+(const Dual2<N>& a, const Dual2<N>& b) {
  Dual2<N> r; ORC_FLOP(1 + N + Dual2<N>::NH); r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < Dual2<N>::NH; ++i) r.h[i] = a.h[i] + b.h[i]; return r; }
template <int N> inline Dual2<N> operator-(const Dual2<N>& a, const Dual2<N>& b) {
  Dual2<N> r; ORC_FLOP(1 + N + Dual2<N>::NH); r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] - b.g[i];
  for (int i = 0; i < Dual2<N>::NH; ++i) r.h[i] = a.h[i] - b.h[i]; return r; }
template <int N> inline Dual2<N> operator*(const Dual2<N>& a, const Dual2<N>& b) {
  Dual2<N> r; ORC_FLOP(1 + 3 * N + 7 * Dual2<N>::NH); r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
  for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) {
    int k = Dual2<N>::idx(i, j);
    r.h[k] = a.h[k] * b.v + a.v * b.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  }
  return r; }
template <int N> inline Dual2<N> operator*(const Dual2<N>& a, double b) {
  Dual2<N> r; ORC_FLOP(1 + N + Dual2<N>::NH); r.v = a.v * b; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b;
  for (int i = 0; i < Dual2<N>::NH; ++i) r.h[i] = a.h[i] * b; return r; }
template <int N> inline Dual2<N> operator*(double b, const Dual2<N>& a) { return a * b; }
template <int N> inline Dual2<N> operator+(const Dual2<N>& a, double b) { Dual2<N> r = a; ORC_FLOP(1); r.v += b; return r; }
template <int N> inline Dual2<N> operator-(const Dual2<N>& a, double b) { Dual2<N> r = a; ORC_FLOP(1); r.v -= b; return r; }
template <int N> inline Dual2<N> sin(const Dual2<N>& a) {
  Dual2<N> r; ORC_FLOP(2 + N + 4 * Dual2<N>::NH); double s = std::sin(a.v), c = std::cos(a.v); r.v = s;
  for (int i = 0; i < N; ++i) r.g[i] = c * a.g[i];
  for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) { int k = Dual2<N>::idx(i, j); r.h[k] = c * a.h[k] - s * a.g[i] * a.g[j]; }
  return r; }
template <int N> inline Dual2<N> cos(const Dual2<N>& a) {
  Dual2<N> r; ORC_FLOP(2 + N + 4 * Dual2<N>::NH); double s = std::sin(a.v), c = std::cos(a.v); r.v = c;
  for (int i = 0; i < N; ++i) r.g[i] = -s * a.g[i];
  for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) { int k = Dual2<N>::idx(i, j); r.h[k] = -s * a.h[k] - c * a.g[i] * a.g[j]; }
  return r; }

}  // namespace orc
