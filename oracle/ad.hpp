// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product.
// Second-order forward-mode AD scalar with NV independent variables
// (value, gradient, packed upper-triangular Hessian).  The oracle gets every
// derivative of the OCP functions from this type applied to *value* code that
// restates the reference formulas, so its derivatives are independent of the
// hand-derived ones in the CUDA kernels.
#pragma once
#include <cmath>

template <int NV>
struct D2 {
  static constexpr int NH = NV * (NV + 1) / 2;
  double v;
  double g[NV];
  double h[NH];
  static int hidx(int i, int j) {  // i <= j
    return i * NV - i * (i - 1) / 2 + (j - i);
  }
  D2() : v(0) {
    for (int i = 0; i < NV; i++) g[i] = 0;
    for (int i = 0; i < NH; i++) h[i] = 0;
  }
  D2(double c) : v(c) {
    for (int i = 0; i < NV; i++) g[i] = 0;
    for (int i = 0; i < NH; i++) h[i] = 0;
  }
  static D2 var(double val, int idx) {
    D2 r(val);
    r.g[idx] = 1.0;
    return r;
  }
  double hess(int i, int j) const { return i <= j ? h[hidx(i, j)] : h[hidx(j, i)]; }
};

template <int NV>
inline D2<NV> operator+(const D2<NV>& a, const D2<NV>& b) {
  D2<NV> r;
  r.v = a.v + b.v;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < D2<NV>::NH; i++) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <int NV>
inline D2<NV> operator-(const D2<NV>& a, const D2<NV>& b) {
  D2<NV> r;
  r.v = a.v - b.v;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] - b.g[i];
  for (int i = 0; i < D2<NV>::NH; i++) r.h[i] = a.h[i] - b.h[i];
  return r;
}
template <int NV>
inline D2<NV> operator-(const D2<NV>& a) {
  D2<NV> r;
  r.v = -a.v;
  for (int i = 0; i < NV; i++) r.g[i] = -a.g[i];
  for (int i = 0; i < D2<NV>::NH; i++) r.h[i] = -a.h[i];
  return r;
}
template <int NV>
inline D2<NV> operator*(const D2<NV>& a, const D2<NV>& b) {
  D2<NV> r;
  r.v = a.v * b.v;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
  int k = 0;
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++, k++)
      r.h[k] = a.h[k] * b.v + a.v * b.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}
// f(a) with first/second derivative f1, f2 at a.v
template <int NV>
inline D2<NV> chain(const D2<NV>& a, double f0, double f1, double f2) {
  D2<NV> r;
  r.v = f0;
  for (int i = 0; i < NV; i++) r.g[i] = f1 * a.g[i];
  int k = 0;
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++, k++) r.h[k] = f1 * a.h[k] + f2 * a.g[i] * a.g[j];
  return r;
}
template <int NV>
inline D2<NV> operator/(const D2<NV>& a, const D2<NV>& b) {
  double iv = 1.0 / b.v;
  return a * chain(b, iv, -iv * iv, 2 * iv * iv * iv);
}
template <int NV> inline D2<NV> operator+(const D2<NV>& a, double b) { D2<NV> r = a; r.v += b; return r; }
template <int NV> inline D2<NV> operator+(double b, const D2<NV>& a) { D2<NV> r = a; r.v += b; return r; }
template <int NV> inline D2<NV> operator-(const D2<NV>& a, double b) { D2<NV> r = a; r.v -= b; return r; }
template <int NV> inline D2<NV> operator-(double b, const D2<NV>& a) { D2<NV> r = -a; r.v += b; return r; }
template <int NV>
inline D2<NV> operator*(const D2<NV>& a, double b) {
  D2<NV> r;
  r.v = a.v * b;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] * b;
  for (int i = 0; i < D2<NV>::NH; i++) r.h[i] = a.h[i] * b;
  return r;
}
template <int NV> inline D2<NV> operator*(double b, const D2<NV>& a) { return a * b; }
template <int NV> inline D2<NV> operator/(const D2<NV>& a, double b) { return a * (1.0 / b); }
template <int NV> inline D2<NV> operator/(double a, const D2<NV>& b) { return D2<NV>(a) / b; }
template <int NV> inline D2<NV> sin(const D2<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return chain(a, s, c, -s); }
template <int NV> inline D2<NV> cos(const D2<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return chain(a, c, -s, -c); }
template <int NV> inline D2<NV> exp(const D2<NV>& a) { double e = std::exp(a.v); return chain(a, e, e, e); }

inline double valof(double a) { return a; }
template <int NV> inline double valof(const D2<NV>& a) { return a.v; }
