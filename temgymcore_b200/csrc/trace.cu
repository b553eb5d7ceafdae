// K1: batched paraxial ray propagation through a model with in-register forward-mode
// duals (the 5x5 ABCD block or the full 7x7 Jacobian), K5: metres->pixels, into_image.
//
// Replaces, for whole batches in one launch:
//   run_to_end / run_iter            reference src/temgym_core/run.py:47-116
//   FreeSpaceParaxial.propagate      reference src/temgym_core/propagator.py:52-72
//   component __call__s              reference src/temgym_core/components.py
//   vmap(jacobian(run_to_end)) + custom_jacobian_matrix
//                                    reference gaussian.py:234-239, utils.py:7-43
//
// Compiled with -fmad=false: every + - * / is a separate IEEE fp64 operation in the
// order the reference writes it, so non-transcendental models reproduce the numpy
// oracle bit for bit.  The kernel is HBM-bound (312 B/ray algorithmic), so the lost
// FMA contraction costs nothing.
//
// Data layout: rays are SoA (seven fp64 arrays); one ray per thread, coalesced 8 B
// loads/stores per field.  The (n,5,5)/(n,7,7) Jacobian is AoS in global memory as the
// API returns it; a block stages its rows in shared memory (odd row pitch -> no bank
// conflicts) and emits them as ONE contiguous bulk async copy through the TMA unit
// (cp.async.bulk.global.shared::cta), so HBM sees full lines only.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "tg_common.cuh"

namespace {

constexpr int kTraceThreads = 128;

// ------------------------------------------------------------------ dual numbers
template <int N>
struct Dual {
  double v;
  double t[N > 0 ? N : 1];
};
template <int N, int M>
struct DMax {
  static_assert(N == M || N == 0 || M == 0, "mixed tangent widths");
  static constexpr int value = N > M ? N : M;
};
template <int N, int M>
using DR = Dual<DMax<N, M>::value>;

template <int N>
__device__ __forceinline__ Dual<N> dconst(double v) {
  Dual<N> r;
  r.v = v;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = 0.0;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dseed(double v, int col) {
  Dual<N> r = dconst<N>(v);
  if (N > 0 && col >= 0) {
#pragma unroll
    for (int k = 0; k < (N > 0 ? N : 1); ++k)
      if (k == col) r.t[k] = 1.0;
  }
  return r;
}
template <int N, int M>
__device__ __forceinline__ DR<N, M> operator+(const Dual<N> &a, const Dual<M> &b) {
  DR<N, M> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < DMax<N, M>::value; ++k) {
    if constexpr (N > 0 && M > 0) r.t[k] = a.t[k] + b.t[k];
    else if constexpr (N > 0) r.t[k] = a.t[k];
    else r.t[k] = b.t[k];
  }
  return r;
}
template <int N, int M>
__device__ __forceinline__ DR<N, M> operator-(const Dual<N> &a, const Dual<M> &b) {
  DR<N, M> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < DMax<N, M>::value; ++k) {
    if constexpr (N > 0 && M > 0) r.t[k] = a.t[k] - b.t[k];
    else if constexpr (N > 0) r.t[k] = a.t[k];
    else r.t[k] = -b.t[k];
  }
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N> &a) {
  Dual<N> r;
  r.v = -a.v;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = -a.t[k];
  return r;
}
template <int N, int M>
__device__ __forceinline__ DR<N, M> operator*(const Dual<N> &a, const Dual<M> &b) {
  DR<N, M> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < DMax<N, M>::value; ++k) {
    if constexpr (N > 0 && M > 0) r.t[k] = a.t[k] * b.v + b.t[k] * a.v;
    else if constexpr (N > 0) r.t[k] = a.t[k] * b.v;
    else r.t[k] = b.t[k] * a.v;
  }
  return r;
}
template <int N, int M>
__device__ __forceinline__ DR<N, M> operator/(const Dual<N> &a, const Dual<M> &b) {
  DR<N, M> r;
  const double q = a.v / b.v;
  r.v = q;
#pragma unroll
  for (int k = 0; k < DMax<N, M>::value; ++k) {
    if constexpr (N > 0 && M > 0) r.t[k] = (a.t[k] - b.t[k] * q) / b.v;
    else if constexpr (N > 0) r.t[k] = a.t[k] / b.v;
    else r.t[k] = (0.0 - b.t[k] * q) / b.v;
  }
  return r;
}
// scalar mixes
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N> &a, double b) {
  Dual<N> r = a;
  r.v = a.v + b;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(double b, const Dual<N> &a) {
  return a + b;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N> &a, double b) {
  Dual<N> r = a;
  r.v = a.v - b;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(double b, const Dual<N> &a) {
  Dual<N> r;
  r.v = b - a.v;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = -a.t[k];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N> &a, double b) {
  Dual<N> r;
  r.v = a.v * b;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = a.t[k] * b;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(double b, const Dual<N> &a) {
  return a * b;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N> &a, double b) {
  Dual<N> r;
  r.v = a.v / b;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = a.t[k] / b;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(double b, const Dual<N> &a) {
  Dual<N> r;
  const double q = b / a.v;
  r.v = q;
  const double s = q / a.v;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = -a.t[k] * s;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dsign(const Dual<N> &a) {
  // jnp.sign: sign(0) = 0, zero derivative (components.py:556)
  return dconst<N>(a.v > 0.0 ? 1.0 : (a.v < 0.0 ? -1.0 : (a.v == 0.0 ? 0.0 : a.v)));
}
template <int N>
__device__ __forceinline__ void dsincos(const Dual<N> &a, Dual<N> &s, Dual<N> &c) {
  double sv, cv;
  sincos(a.v, &sv, &cv);
  s.v = sv;
  c.v = cv;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) {
    s.t[k] = cv * a.t[k];
    c.t[k] = (-sv) * a.t[k];
  }
}
template <int N>
__device__ __forceinline__ Dual<N> dhypot(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  const double h = hypot(a.v, b.v);
  r.v = h;
  const double ca = a.v / h, cb = b.v / h;  // NaN at (0,0) like jnp.hypot's gradient
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = ca * a.t[k] + cb * b.t[k];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> datan2(const Dual<N> &y, const Dual<N> &x) {
  Dual<N> r;
  r.v = atan2(y.v, x.v);
  const double r2 = x.v * x.v + y.v * y.v;
  const double cy = x.v / r2, cx = y.v / r2;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = cy * y.t[k] - cx * x.t[k];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dwhere_zero(const Dual<N> &a, double eps) {
  return a.v == 0.0 ? dconst<N>(eps) : a;
}
// 1 / a with ONE fp64 division (value and tangents share the reciprocal)
template <int N>
__device__ __forceinline__ Dual<N> drecip(const Dual<N> &a) {
  Dual<N> r;
  const double q = 1.0 / a.v;
  r.v = q;
  const double s = -(q * q);
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = a.t[k] * s;
  return r;
}
template <int M, int N>
__device__ __forceinline__ Dual<M> dnarrow(const Dual<N> &a) {
  // keep the value; keep tangents only if the destination tracks them
  if constexpr (M == N) {
    return a;
  } else {
    static_assert(M == 0, "narrowing only to value-only");
    Dual<M> r;
    r.v = a.v;
    r.t[0] = 0.0;
    return r;
  }
}

// ------------------------------------------------------------------ Krivanek aberrations
// aberrations.py:42-108; p = comp.p + 1 (25 coefficients in KrivanekCoeffs field order).
enum {
  K_C10 = 0, K_C12, K_PHI12, K_C21, K_PHI21, K_C23, K_PHI23, K_C30, K_C32, K_PHI32, K_C34,
  K_PHI34, K_C41, K_PHI41, K_C43, K_PHI43, K_C45, K_PHI45, K_C50, K_C52, K_PHI52, K_C54,
  K_PHI54, K_C56, K_PHI56
};

// The Krivanek lens makes the ray kernel compute-bound (fp64 pipe), so its arithmetic is written
// with EXPLICIT fused multiply-adds (this TU is built with -fmad=false; fma() calls are honoured):
// ~30 % fewer fp64 instructions than separate multiplies and adds.  The helpers are used by every
// instantiation (value-only, 5- and 7-column, gradient kernel) with the same value arithmetic, so
// run_to_end and run_to_end_abcd still agree bit for bit.  (The algebraic harmonic evaluation is not
// bit-comparable to the reference's hypot / arctan2 / cos chain anyway; parity is 1e-12 relative.)
template <int N>
__device__ __forceinline__ Dual<N> kmul(const Dual<N> &a, const Dual<N> &b) {           // a b
  Dual<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = N > 0 ? fma(a.t[k], b.v, a.v * b.t[k]) : 0.0;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> kmul2(const Dual<N> &a, const Dual<N> &b, const Dual<N> &c,
                                         const Dual<N> &d, double sgn) {                   // a b + sgn c d
  Dual<N> r;
  r.v = fma(a.v, b.v, sgn * (c.v * d.v));
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k)
    r.t[k] = N > 0 ? fma(a.t[k], b.v, fma(a.v, b.t[k], sgn * fma(c.t[k], d.v, c.v * d.t[k]))) : 0.0;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> klin(const Dual<N> &a, double s, const Dual<N> &b, double t) {  // s a + t b
  Dual<N> r;
  r.v = fma(a.v, s, b.v * t);
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = N > 0 ? fma(a.t[k], s, b.t[k] * t) : 0.0;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> kaxpy(const Dual<N> &a, double s, const Dual<N> &b) {  // s a + b
  Dual<N> r;
  r.v = fma(a.v, s, b.v);
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = N > 0 ? fma(a.t[k], s, b.t[k]) : 0.0;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> kfma(const Dual<N> &a, const Dual<N> &b, const Dual<N> &c) {  // a b + c
  Dual<N> r;
  r.v = fma(a.v, b.v, c.v);
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = N > 0 ? fma(a.t[k], b.v, fma(a.v, b.t[k], c.t[k])) : 0.0;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> kscale(const Dual<N> &a, double s) {
  Dual<N> r;
  r.v = a.v * s;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) r.t[k] = N > 0 ? a.t[k] * s : 0.0;
  return r;
}

// ---- polynomial form of the Krivanek aberration function ---------------------------------------
// Every term of aberrations.py:42-60 is a POLYNOMIAL in the slope components: with w = ax + i ay,
// r2 = |w|^2, b = (n + 1 - m) / 2 and z_nm = C_nm / (n + 1) * exp(-i m phi_nm),
//     C_nm alpha^(n+1) / (n+1) cos(m (phi - phi_nm)) = r2^b Re(z_nm w^m),
// so   W = Re[ F0(w) + r2 F1(w) + r2^2 F2(w) + r2^3 F3 ]   with the holomorphic polynomials
//     F0 = z12 w^2 + z23 w^3 + z34 w^4 + z45 w^5 + z56 w^6      F2 = z30 + z41 w + z52 w^2
//     F1 = z10 + z21 w + z32 w^2 + z43 w^3 + z54 w^4            F3 = z50.
// For holomorphic F: grad Re F = conj(F') (as gx + i gy) and Hess Re F = [[Re F'', -Im F''], [-Im F'', -Re F'']],
// and the product rule with r2^b gives grad W (aberrations.py:63-108) and its Jacobian in closed form:
// no hypot / atan2 / division / normalised phasors, ~4x fewer fp64 instructions than differentiating
// the polar form, and nothing singular at the origin (where the reference's hypot / arctan2 gradients
// are NaN; that is reproduced explicitly).  Terms with C == 0 contribute exactly zero in the reference
// as well and are skipped (uniform branches on kernel-parameter constants).
struct Cx {
  double re, im;
};
__device__ __forceinline__ Cx cx_mul(const Cx &a, const Cx &b) {
  Cx r;
  r.re = fma(a.re, b.re, -(a.im * b.im));
  r.im = fma(a.re, b.im, a.im * b.re);
  return r;
}
struct KrivF {
  Cx F, F1, F2;  // F_b(w), F_b'(w), F_b''(w)
};
// accumulate the term z w^K, z = kappa C (c0 - i s0); wkm2 = w^(K-2) (only read when K >= 3)
template <int K, bool HESS>
__device__ __forceinline__ void kriv_poly_term(double C, double kappa, double c0, double s0, const Cx &w,
                                               const Cx &wkm2, KrivF &A) {
  if (C == 0.0) return;
  const double s = C * kappa;
  Cx z;
  z.re = s * c0;
  z.im = -(s * s0);
  if constexpr (K == 1) {
    const Cx t = cx_mul(z, w);
    A.F.re += t.re;
    A.F.im += t.im;
    A.F1.re += z.re;
    A.F1.im += z.im;
  } else {
    const Cx T = (K == 2) ? z : cx_mul(z, wkm2);  // z w^(K-2)
    if constexpr (HESS) {
      A.F2.re = fma((double)(K * (K - 1)), T.re, A.F2.re);
      A.F2.im = fma((double)(K * (K - 1)), T.im, A.F2.im);
    }
    const Cx T1 = cx_mul(T, w);
    A.F1.re = fma((double)K, T1.re, A.F1.re);
    A.F1.im = fma((double)K, T1.im, A.F1.im);
    const Cx T2 = cx_mul(T1, w);
    A.F.re += T2.re;
    A.F.im += T2.im;
  }
}
struct KrivOut {
  double W, Gx, Gy, Hxx, Hxy, Hyy;
};
// add the group r2^B Re F_B(w): rho = r2^B, rho1 = B r2^(B-1), rho2 = B (B-1) r2^(B-2)
template <int B, bool HESS>
__device__ __forceinline__ void kriv_poly_combine(const KrivF &A, double u, double v, double rho, double rho1,
                                                  double rho2, KrivOut &o) {
  const double f = A.F.re, gx = A.F1.re, gy = -A.F1.im;
  if constexpr (B == 0) {
    o.W += f;
    o.Gx += gx;
    o.Gy += gy;
    if constexpr (HESS) {
      o.Hxx += A.F2.re;
      o.Hxy -= A.F2.im;
      o.Hyy -= A.F2.re;
    }
  } else {
    const double t1 = 2.0 * rho1, t1f = t1 * f;      // grad rho = t1 (u, v)
    o.W = fma(rho, f, o.W);
    o.Gx = fma(rho, gx, fma(t1f, u, o.Gx));
    o.Gy = fma(rho, gy, fma(t1f, v, o.Gy));
    if constexpr (HESS) {
      // f Hess rho + grad rho grad f^T + grad f grad rho^T + rho Hess f
      const double t2f = 4.0 * rho2 * f;
      const double pu = t1 * u, pv = t1 * v;
      o.Hxx = fma(rho, A.F2.re, fma(2.0 * pu, gx, fma(t2f * u, u, o.Hxx + t1f)));
      o.Hyy = fma(-rho, A.F2.re, fma(2.0 * pv, gy, fma(t2f * v, v, o.Hyy + t1f)));
      o.Hxy = fma(-rho, A.F2.im, fma(pu, gy, fma(pv, gx, fma(t2f * u, v, o.Hxy))));
    }
  }
}
// p = comp.p + 1: 25 coefficients, then 11 (cos, sin)(m phi0) pairs
template <bool HESS>
__device__ __forceinline__ KrivOut krivanek_poly(const double *p, double u, double v) {
  const double *g = p + 25;
  Cx w;
  w.re = u;
  w.im = v;
  const double r2 = fma(u, u, v * v);
  Cx w2, w3, w4;
  w2.re = w2.im = w3.re = w3.im = w4.re = w4.im = 0.0;
  const bool n4 = p[K_C56] != 0.0, n3 = p[K_C45] != 0.0;
  const bool n2 = p[K_C34] != 0.0 || p[K_C54] != 0.0 || n3 || n4;
  if (n2) w2 = cx_mul(w, w);
  if (n3) w3 = cx_mul(w2, w);
  if (n4) w4 = cx_mul(w2, w2);
  KrivOut o;
  o.W = o.Gx = o.Gy = o.Hxx = o.Hxy = o.Hyy = 0.0;
  KrivF A;
  // b = 0: m = n + 1
  if (p[K_C12] != 0.0 || p[K_C23] != 0.0 || p[K_C34] != 0.0 || p[K_C45] != 0.0 || p[K_C56] != 0.0) {
    A.F.re = A.F.im = A.F1.re = A.F1.im = A.F2.re = A.F2.im = 0.0;
    kriv_poly_term<2, HESS>(p[K_C12], 0.5, g[0], g[1], w, w, A);
    kriv_poly_term<3, HESS>(p[K_C23], 1.0 / 3.0, g[4], g[5], w, w, A);
    kriv_poly_term<4, HESS>(p[K_C34], 0.25, g[8], g[9], w, w2, A);
    kriv_poly_term<5, HESS>(p[K_C45], 0.2, g[14], g[15], w, w3, A);
    kriv_poly_term<6, HESS>(p[K_C56], 1.0 / 6.0, g[20], g[21], w, w4, A);
    kriv_poly_combine<0, HESS>(A, u, v, 1.0, 0.0, 0.0, o);
  }
  // b = 1: m = n - 1
  if (p[K_C10] != 0.0 || p[K_C21] != 0.0 || p[K_C32] != 0.0 || p[K_C43] != 0.0 || p[K_C54] != 0.0) {
    A.F.re = 0.5 * p[K_C10];
    A.F.im = A.F1.re = A.F1.im = A.F2.re = A.F2.im = 0.0;
    kriv_poly_term<1, HESS>(p[K_C21], 1.0 / 3.0, g[2], g[3], w, w, A);
    kriv_poly_term<2, HESS>(p[K_C32], 0.25, g[6], g[7], w, w, A);
    kriv_poly_term<3, HESS>(p[K_C43], 0.2, g[12], g[13], w, w, A);
    kriv_poly_term<4, HESS>(p[K_C54], 1.0 / 6.0, g[18], g[19], w, w2, A);
    kriv_poly_combine<1, HESS>(A, u, v, r2, 1.0, 0.0, o);
  }
  // b = 2: m = n - 3
  if (p[K_C30] != 0.0 || p[K_C41] != 0.0 || p[K_C52] != 0.0) {
    A.F.re = 0.25 * p[K_C30];
    A.F.im = A.F1.re = A.F1.im = A.F2.re = A.F2.im = 0.0;
    kriv_poly_term<1, HESS>(p[K_C41], 0.2, g[10], g[11], w, w, A);
    kriv_poly_term<2, HESS>(p[K_C52], 1.0 / 6.0, g[16], g[17], w, w, A);
    kriv_poly_combine<2, HESS>(A, u, v, r2 * r2, 2.0 * r2, 2.0, o);
  }
  // b = 3: C50 alpha^6 / 6
  if (p[K_C50] != 0.0) {
    A.F.re = p[K_C50] * (1.0 / 6.0);
    A.F.im = A.F1.re = A.F1.im = A.F2.re = A.F2.im = 0.0;
    const double r4 = r2 * r2;
    kriv_poly_combine<3, HESS>(A, u, v, r4 * r2, 3.0 * r4, 6.0 * r2, o);
  }
  return o;
}

// Partial derivatives of (W, dW/dax, dW/day) w.r.t. ONE KrivanekCoeffs field (0..24 in field order), for the
// parameter tangents of run_with_grads: the polynomial form is linear in every C_nm, and a phase phi_nm only
// enters through z_nm = kappa C_nm (cos m phi - i sin m phi), whose derivative swaps the (cos, sin) pair for
// (-m sin, m cos).  So the derivative is the same polynomial evaluated on a one-term coefficient set.
__device__ __noinline__ KrivOut krivanek_poly_dfield(const double *p, int field, double u, double v) {
  // harmonic pair index (order C12 C21 C23 C32 C34 C41 C43 C45 C52 C54 C56) and m per field; -1: no phase
  const signed char pair[25] = {-1, 0, 0, 1, 1, 2, 2, -1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, -1, 8, 8, 9, 9, 10, 10};
  const signed char mval[25] = {0, 2, 2, 1, 1, 3, 3, 0, 2, 2, 4, 4, 1, 1, 3, 3, 5, 5, 0, 2, 2, 4, 4, 6, 6};
  const bool is_phi[25] = {false, false, true, false, true, false, true, false, false, true, false, true, false,
                           true, false, true, false, true, false, false, true, false, true, false, true};
  double q[47];
  for (int k = 0; k < 47; ++k) q[k] = 0.0;
  const int t = pair[field];
  const double m = (double)mval[field];
  if (!is_phi[field]) {
    q[field] = 1.0;
    if (t >= 0) {
      q[25 + 2 * t] = p[25 + 2 * t];
      q[26 + 2 * t] = p[26 + 2 * t];
    }
  } else {
    q[field - 1] = p[field - 1];
    q[25 + 2 * t] = -m * p[26 + 2 * t];
    q[26 + 2 * t] = m * p[25 + 2 * t];
  }
  return krivanek_poly<false>(q, u, v);
}

// (dW/dax, dW/day, W) of the Krivanek aberration function (aberrations.py:42-108).  N = 0: values only;
// N = 2: the arguments are the seeds (ax, e0), (ay, e1), the tangents returned are the derivatives
// w.r.t. (ax, ay) -- the symmetric Hessian of W for dWx / dWy, the gradient for W -- and the caller
// chains them to the ray tangents, which keeps the live register set small.
template <int N>
__device__ __forceinline__ void krivanek(const double *p, const Dual<N> &ax, const Dual<N> &ay,
                                         Dual<N> &dWx, Dual<N> &dWy, Dual<N> &W) {
  static_assert(N == 0 || N == 2, "krivanek: value-only or tangents w.r.t. its own two arguments");
  const KrivOut o = krivanek_poly<(N == 2)>(p, ax.v, ay.v);
  dWx.v = o.Gx;
  dWy.v = o.Gy;
  W.v = o.W;
  if constexpr (N == 2) {
    // jnp.hypot / arctan2 have NaN gradients at the origin (aberrations.py:66-67): every derivative
    // through the lens is NaN there in the reference
    const bool origin = ax.v == 0.0 && ay.v == 0.0;
    const double qn = nan("");
    dWx.t[0] = origin ? qn : o.Hxx;
    dWx.t[1] = origin ? qn : o.Hxy;
    dWy.t[0] = origin ? qn : o.Hxy;
    dWy.t[1] = origin ? qn : o.Hyy;
    W.t[0] = origin ? qn : o.Gx;
    W.t[1] = origin ? qn : o.Gy;
  } else {
    dWx.t[0] = dWy.t[0] = W.t[0] = 0.0;
  }
}

// chain rule: r = f(ax, ay) given as Dual<2> (tangents w.r.t. ax, ay) -> tangents of width N
template <int N>
__device__ __forceinline__ Dual<N> dchain(const Dual<2> &r, const Dual<N> &ax, const Dual<N> &ay) {
  Dual<N> o;
  o.v = r.v;
#pragma unroll
  for (int k = 0; k < (N > 0 ? N : 1); ++k) o.t[k] = N > 0 ? fma(r.t[0], ax.t[k], r.t[1] * ay.t[k]) : 0.0;
  return o;
}

// ------------------------------------------------------------------ the kernel
struct TraceOut {
  double *ptr[7];
};

// Kernel-parameter model descriptors.  Components other than the Krivanek lens use at most four
// parameters, so models without one travel as a compact 1.2 KB block (the full tg_model is
// 9.6 KB; a large parameter block measurably slows block launch in this HBM-bound kernel).
struct CompLite {
  int32_t op, flags;
  double z;
  double p[4];
};
struct ModelLite {
  int32_t n_comp, reserved;
  CompLite comp[TG_MAX_COMPS];
};
template <bool KRIV>
struct ModelFor {
  using type = ModelLite;
};
template <>
struct ModelFor<true> {
  using type = tg_model;
};
inline void to_lite(const tg_model &m, ModelLite &l) {
  l.n_comp = m.n_comp;
  l.reserved = 0;
  for (int c = 0; c < m.n_comp; ++c) {
    l.comp[c].op = m.comp[c].op;
    l.comp[c].flags = m.comp[c].flags;
    l.comp[c].z = m.comp[c].z;
    for (int k = 0; k < 4; ++k) l.comp[c].p[k] = m.comp[c].p[k];
  }
}

// NC = number of tangent columns (0 none, 5 = [x,y,dx,dy,_one], 7 = all Ray leaves).
// In the 5-column mode z and pathlength carry no tangents: d z_out / d {x,y,dx,dy,_one}
// is identically zero for every component on the path (z only ever receives component
// constants), and pathlength never feeds back into x,y,dx,dy.
// PERRAY: Scanner / Descanner components whose parameters are ARRAYS over the ray batch (what the reference gets
// from jax.vmap over scan positions, components.py:252-372): their four offsets come from per-ray SoA arrays
// (tg_perray) instead of the constant-bank descriptor.  A separate instantiation, so the common scalar-parameter
// kernels keep their register count and parameter block.
template <int NC, bool KRIV, bool PERRAY>
__global__ void __launch_bounds__(kTraceThreads, KRIV ? (NC == 5 ? 4 : 2) : (NC == 7 ? 3 : 6))
    trace_kernel(const __grid_constant__ typename ModelFor<KRIV>::type model, const tg_ray_in in,
                 const long long n, const TraceOut out, double *__restrict__ jac, const tg_perray pr) {
  constexpr bool FULL = (NC == 7);
  constexpr int NZ = FULL ? 7 : 0;  // tangent width of z and pathlength
  constexpr int ROWS = (NC == 7) ? 7 : 5;
  constexpr int JW = ROWS * (NC > 0 ? NC : 1);  // doubles per ray in the Jacobian
  extern __shared__ __align__(16) double s_jac[];

  // The launcher may bound the grid (TG_TRACE_PERSIST): a CTA then walks several blocks of 128 rays; its staging
  // buffer is free again once thread 0 has seen the bulk store read it (wait below) and the CTA has met at the barrier.
  const long long nblk = (n + kTraceThreads - 1) / kTraceThreads;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
  if (NC > 0 && blk != (long long)blockIdx.x) __syncthreads();
  const long long block_first = blk * kTraceThreads;
  const long long i = block_first + threadIdx.x;
  const bool active = i < n;

  if (active) {
    // column ids: NC==5 -> x0 y1 dx2 dy3 one4 ; NC==7 -> x0 y1 dx2 dy3 z4 pl5 one6
    constexpr int COL_ONE = FULL ? 6 : 4;
    auto ld = [&](int f) -> double { return in.ptr[f] ? __ldg(in.ptr[f] + i) : in.value[f]; };
    Dual<NC> x = dseed<NC>(ld(0), 0);
    Dual<NC> y = dseed<NC>(ld(1), 1);
    Dual<NC> dx = dseed<NC>(ld(2), 2);
    Dual<NC> dy = dseed<NC>(ld(3), 3);
    Dual<NZ> z = dseed<NZ>(ld(4), FULL ? 4 : -1);
    Dual<NZ> pl = dseed<NZ>(ld(5), FULL ? 5 : -1);
    Dual<NC> one = dseed<NC>(ld(6), NC > 0 ? COL_ONE : -1);

    const int nc = model.n_comp;
    for (int c = 0; c < nc; ++c) {
      const auto &cm = model.comp[c];
      if (!(cm.flags & TG_F_NOPROP)) {
        // distance = component.z - ray.z (run.py:77); FreeSpaceParaxial (propagator.py:67-72)
        const Dual<NZ> d = (cm.flags & TG_F_DIST) ? dconst<NZ>(cm.z) : cm.z - z;
        if constexpr (KRIV && !FULL) {  // compute-bound instantiation: fused multiply-adds (see kmul)
          x = kaxpy(dx, d.v, x);
          y = kaxpy(dy, d.v, y);
        } else {
          x = x + dx * d;
          y = y + dy * d;
        }
        z = z + d;
        pl = pl + d;
      }
      switch (cm.op) {
        case TG_OP_PLANE:
          break;
        case TG_OP_LENS:
        case TG_OP_THICKLENS: {  // components.py:161-174, 431-447
          const double f = cm.p[0];
          if constexpr (KRIV) {
            // compute-bound instantiation (a Krivanek lens is in the model): one reciprocal per
            // lens instead of 13 fp64 divisions, <= 1 ulp from the exact quotient
            const double inv_f = 1.0 / f;
            // (separate multiply and add: x.t * (1/f) rounds to exactly 1 where the reference's x.t / f does,
            // e.g. a lens one focal length behind a parallel beam, and the zero stays an exact zero)
            const Dual<NC> ndx = (-x) * inv_f + dx;
            const Dual<NC> ndy = (-y) * inv_f + dy;
            pl = pl - dnarrow<NZ>((x * x + y * y) * (0.5 * inv_f));
            dx = ndx;
            dy = ndy;
          } else {
            const Dual<NC> ndx = (-x) / f + dx;
            const Dual<NC> ndy = (-y) / f + dy;
            pl = pl - dnarrow<NZ>((x * x + y * y) / (2.0 * f));
            dx = ndx;
            dy = ndy;
          }
          one = one * 1.0;
          if (cm.op == TG_OP_THICKLENS) z = z - cm.p[1];
        } break;
        case TG_OP_DEFLECTOR: {  // components.py:476-482
          pl = pl + dnarrow<NZ>(dx * x) + dnarrow<NZ>(dy * y);
          if constexpr (KRIV) {
            dx = kaxpy(one, cm.p[0], dx);
            dy = kaxpy(one, cm.p[1], dy);
          } else {
            dx = dx + cm.p[0] * one;
            dy = dy + cm.p[1] * one;
          }
        } break;
        case TG_OP_BIPRISM: {  // components.py:553-559
          pl = pl + dnarrow<NZ>(dx * x) + dnarrow<NZ>(dy * y);
          if constexpr (KRIV) dx = kaxpy(one, cm.p[0] * dsign(x).v, dx);
          else dx = dx + cm.p[0] * one * dsign(x);
        } break;
        case TG_OP_OFFSET: {  // Scanner / Descanner, components.py:279-285, 343-372
          double o0 = cm.p[0], o1 = cm.p[1], o2 = cm.p[2], o3 = cm.p[3];
          if constexpr (PERRAY) {
            for (int k = 0; k < pr.n; ++k)
              if (pr.comp[k] == c) {
                if (pr.ptr[k][0]) o0 = __ldg(pr.ptr[k][0] + i);
                if (pr.ptr[k][1]) o1 = __ldg(pr.ptr[k][1] + i);
                if (pr.ptr[k][2]) o2 = __ldg(pr.ptr[k][2] + i);
                if (pr.ptr[k][3]) o3 = __ldg(pr.ptr[k][3] + i);
              }
          }
          x = x + o0 * one;
          y = y + o1 * one;
          dx = dx + o2 * one;
          dy = dy + o3 * one;
        } break;
        case TG_OP_ROTATOR: {  // components.py:503-523
          const double cs = cm.p[0], sn = cm.p[1];
          const Dual<NC> nx = x * cs - y * sn;
          const Dual<NC> ny = x * sn + y * cs;
          const Dual<NC> ndx = dx * cs - dy * sn;
          const Dual<NC> ndy = dx * sn + dy * cs;
          x = nx;
          y = ny;
          dx = ndx;
          dy = ndy;
        } break;
        case TG_OP_KRIVANEK: {  // components.py:192-215
          if constexpr (KRIV) {
            // compute-bound op: one reciprocal of f instead of ~25 fp64 divisions (<= 1 ulp each)
            const double inv_f = 1.0 / cm.p[0];
            const Dual<NC> idx = (-x) * inv_f + dx;   // unfused on purpose, see TG_OP_LENS
            const Dual<NC> idy = (-y) * inv_f + dy;
            Dual<NC> W;
            if constexpr (NC == 0) {
              Dual<0> dWx, dWy;
              krivanek<0>(cm.p + 1, idx, idy, dWx, dWy, W);
              dx.v = fma(dWx.v, -inv_f, idx.v);
              dy.v = fma(dWy.v, -inv_f, idy.v);
            } else {
              Dual<2> gx, gy, w2;
              krivanek<2>(cm.p + 1, dseed<2>(idx.v, 0), dseed<2>(idy.v, 1), gx, gy, w2);
              W = dchain<NC>(w2, idx, idy);
              // (dx, dy) = (idx, idy) - grad W / f: tangents through the 2x2 matrix I - Hess W / f
              const double m00 = fma(gx.t[0], -inv_f, 1.0), m01 = gx.t[1] * -inv_f;
              const double m10 = gy.t[0] * -inv_f, m11 = fma(gy.t[1], -inv_f, 1.0);
              dx.v = fma(gx.v, -inv_f, idx.v);
              dy.v = fma(gy.v, -inv_f, idy.v);
#pragma unroll
              for (int k = 0; k < NC; ++k) {
                dx.t[k] = fma(m00, idx.t[k], m01 * idy.t[k]);
                dy.t[k] = fma(m10, idx.t[k], m11 * idy.t[k]);
              }
            }
            pl = pl - dnarrow<NZ>((x * x + y * y) * (0.5 * inv_f)) + dnarrow<NZ>(W * inv_f);
            one = one * 1.0;
          }
        } break;
        default:
          break;
      }
    }

    if (out.ptr[0]) out.ptr[0][i] = x.v;
    if (out.ptr[1]) out.ptr[1][i] = y.v;
    if (out.ptr[2]) out.ptr[2][i] = dx.v;
    if (out.ptr[3]) out.ptr[3][i] = dy.v;
    if (out.ptr[4]) out.ptr[4][i] = z.v;
    if (out.ptr[5]) out.ptr[5][i] = pl.v;
    if (out.ptr[6]) out.ptr[6][i] = one.v;

    if constexpr (NC > 0) {
      double *row = s_jac + (size_t)threadIdx.x * JW;
      if constexpr (!FULL) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          row[0 * 5 + k] = x.t[k];
          row[1 * 5 + k] = y.t[k];
          row[2 * 5 + k] = dx.t[k];
          row[3 * 5 + k] = dy.t[k];
          row[4 * 5 + k] = one.t[k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          row[0 * 7 + k] = x.t[k];
          row[1 * 7 + k] = y.t[k];
          row[2 * 7 + k] = dx.t[k];
          row[3 * 7 + k] = dy.t[k];
          row[4 * 7 + k] = z.t[k];
          row[5 * 7 + k] = pl.t[k];
          row[6 * 7 + k] = one.t[k];
        }
      }
    }
  }

  if constexpr (NC > 0) {
    // Emit the block's Jacobian rows: contiguous in global memory.
    const long long rem = n - block_first;
    const int nvalid = rem >= kTraceThreads ? kTraceThreads : (int)rem;
    double *gdst = jac + block_first * JW;
    const uint32_t bytes = (uint32_t)nvalid * JW * 8u;
    const bool bulk_ok = ((bytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(gdst) & 15u) == 0);
    if (bulk_ok) {
      tg_fence_proxy_async();  // make the generic-proxy smem writes visible to the async proxy
      __syncthreads();
      if (threadIdx.x == 0) {
        tg_bulk_s2g(gdst, s_jac, bytes);
        tg_bulk_commit();
        tg_bulk_wait_read0();  // smem must stay valid until the TMA unit has read it
      }
    } else {
      __syncthreads();
      const int total = nvalid * JW;
      for (int k = threadIdx.x; k < total; k += kTraceThreads) gdst[k] = s_jac[k];
    }
  }
  }   // blocks of this CTA
}

// TG_TRACE_PERSIST=<CTAs per SM> (experiment knob, default 0 = one CTA per block of 128 rays): bounded grid whose CTAs
// walk the blocks with a stride.  MEASURED on B200 (RayTracePlan replays, L2 flushed): slower at every setting -- C1 at
// 1e6 rays 0.0594 ms with one CTA per block against 0.0614-0.080 ms with 4 / 6 / 8 / 12 CTAs per SM, 1e7 rays 0.494
// against 0.508-0.684 ms, C4 0.602 against 0.61-0.72 ms: the hardware's block scheduler refills an SM the moment a
// CTA's bulk store has been read, a persistent CTA serialises its own load -> compute -> store chain.
inline int trace_persist() {
  static const int v = [] {
    const char *e = getenv("TG_TRACE_PERSIST");
    const int x = e ? atoi(e) : 0;
    return (x >= 0 && x <= 32) ? x : 0;
  }();
  return v;
}

template <int NC, bool KRIV, bool PERRAY>
int launch_trace_k(const tg_model *m, int64_t n, const tg_ray_in *in, double *const out[7],
                 double *jac, cudaStream_t st, const tg_perray &pr) {
  TraceOut o;
  for (int f = 0; f < 7; ++f) o.ptr[f] = out ? out[f] : nullptr;
  constexpr int ROWS = (NC == 7) ? 7 : 5;
  const size_t smem = NC > 0 ? (size_t)kTraceThreads * ROWS * NC * sizeof(double) : 0;
  if (smem > 48 * 1024) {
    TG_CUDA(cudaFuncSetAttribute(trace_kernel<NC, KRIV, PERRAY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  }
  long long blocks = (n + kTraceThreads - 1) / kTraceThreads;
  TG_REQUIRE(blocks <= 0x7fffffffLL, "too many rays for one launch");
  if (trace_persist() > 0) {
    int dev = 0, sms = 148;
    TG_CUDA(cudaGetDevice(&dev));
    TG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long cap = (long long)sms * trace_persist();
    if (blocks > cap) blocks = cap;
  }
  if constexpr (KRIV) {
    trace_kernel<NC, true, PERRAY><<<(unsigned)blocks, kTraceThreads, smem, st>>>(*m, *in, (long long)n, o, jac, pr);
  } else {
    ModelLite lite;
    to_lite(*m, lite);
    trace_kernel<NC, false, PERRAY><<<(unsigned)blocks, kTraceThreads, smem, st>>>(lite, *in, (long long)n, o, jac, pr);
  }
  return tg_launch_check("trace_kernel");
}

template <int NC>
int launch_trace(const tg_model *m, int64_t n, const tg_ray_in *in, double *const out[7],
                 double *jac, cudaStream_t st, const tg_perray *perray) {
  bool kriv = false;
  for (int c = 0; c < m->n_comp; ++c) kriv |= (m->comp[c].op == TG_OP_KRIVANEK);
  tg_perray none;
  memset(&none, 0, sizeof(none));
  if (perray && perray->n > 0)
    return kriv ? launch_trace_k<NC, true, true>(m, n, in, out, jac, st, *perray)
                : launch_trace_k<NC, false, true>(m, n, in, out, jac, st, *perray);
  return kriv ? launch_trace_k<NC, true, false>(m, n, in, out, jac, st, none)
              : launch_trace_k<NC, false, false>(m, n, in, out, jac, st, none);
}

// ------------------------------------------------------------------ parameter tangents
// run_with_grads (reference run.py:182-267): Jacobians of the output ray w.r.t. selected input
// ray fields AND component parameters.  Same forward-mode duals, but every state variable
// carries TG_GRAD_LANES tangent lanes and component parameters are duals too: a seed
// (comp, slot, lane, weight) adds `weight` to lane `lane` of parameter `slot` (0 = z,
// k+1 = p[k]) of component `comp`; ray field f is seeded on lane ray_lane[f] (or -1).
struct GradSeeds {
  int n;
  int ray_lane[7];
  tg_seed s[TG_MAX_SEEDS];
};
constexpr int NT = TG_GRAD_LANES;

__device__ __forceinline__ Dual<NT> gparam(const tg_model &m, const GradSeeds &gs, int c, int slot) {
  Dual<NT> r = dconst<NT>(slot == 0 ? m.comp[c].z : m.comp[c].p[slot - 1]);
  for (int i = 0; i < gs.n; ++i) {
    if (gs.s[i].comp == c && gs.s[i].slot == slot) {
#pragma unroll
      for (int k = 0; k < NT; ++k) r.t[k] += (gs.s[i].lane == k) ? gs.s[i].weight : 0.0;
    }
  }
  return r;
}

template <bool KRIV>
__global__ void __launch_bounds__(kTraceThreads)
    trace_grad_kernel(const __grid_constant__ tg_model model, const __grid_constant__ GradSeeds gs,
                      const tg_ray_in in, const long long n, const TraceOut out, double *__restrict__ jac) {
  const long long i = (long long)blockIdx.x * kTraceThreads + threadIdx.x;
  if (i >= n) return;
  auto ld = [&](int f) -> double { return in.ptr[f] ? __ldg(in.ptr[f] + i) : in.value[f]; };
  using D = Dual<NT>;
  D x = dseed<NT>(ld(0), gs.ray_lane[0]), y = dseed<NT>(ld(1), gs.ray_lane[1]);
  D dx = dseed<NT>(ld(2), gs.ray_lane[2]), dy = dseed<NT>(ld(3), gs.ray_lane[3]);
  D z = dseed<NT>(ld(4), gs.ray_lane[4]), pl = dseed<NT>(ld(5), gs.ray_lane[5]);
  D one = dseed<NT>(ld(6), gs.ray_lane[6]);
  const int nc = model.n_comp;
  for (int c = 0; c < nc; ++c) {
    const tg_comp &cm = model.comp[c];
    if (!(cm.flags & TG_F_NOPROP)) {
      const D zc = gparam(model, gs, c, 0);
      const D d = (cm.flags & TG_F_DIST) ? zc : zc - z;
      x = x + dx * d;
      y = y + dy * d;
      z = z + d;
      pl = pl + d;
    }
    switch (cm.op) {
      case TG_OP_LENS:
      case TG_OP_THICKLENS: {
        const D f = gparam(model, gs, c, 1);
        const D ndx = (-x) / f + dx;
        const D ndy = (-y) / f + dy;
        pl = pl - (x * x + y * y) / (2.0 * f);
        one = one * 1.0;
        dx = ndx;
        dy = ndy;
        if (cm.op == TG_OP_THICKLENS) z = z - gparam(model, gs, c, 2);
      } break;
      case TG_OP_DEFLECTOR: {
        pl = pl + dx * x + dy * y;
        dx = dx + gparam(model, gs, c, 1) * one;
        dy = dy + gparam(model, gs, c, 2) * one;
      } break;
      case TG_OP_BIPRISM: {
        pl = pl + dx * x + dy * y;
        dx = dx + gparam(model, gs, c, 1) * one * dsign(x);
      } break;
      case TG_OP_OFFSET: {
        x = x + gparam(model, gs, c, 1) * one;
        y = y + gparam(model, gs, c, 2) * one;
        dx = dx + gparam(model, gs, c, 3) * one;
        dy = dy + gparam(model, gs, c, 4) * one;
      } break;
      case TG_OP_ROTATOR: {
        const D cs = gparam(model, gs, c, 1), sn = gparam(model, gs, c, 2);
        const D nx = x * cs - y * sn, ny = x * sn + y * cs;
        const D ndx = dx * cs - dy * sn, ndy = dx * sn + dy * cs;
        x = nx;
        y = ny;
        dx = ndx;
        dy = ndy;
      } break;
      case TG_OP_KRIVANEK: {
        if constexpr (KRIV) {
          const D f = gparam(model, gs, c, 1);
          const D idx = (-x) / f + dx;
          const D idy = (-y) / f + dy;
          Dual<2> gx, gy, w2;
          krivanek<2>(cm.p + 1, dseed<2>(idx.v, 0), dseed<2>(idy.v, 1), gx, gy, w2);
          D dWx = dchain<NT>(gx, idx, idy), dWy = dchain<NT>(gy, idx, idy), W = dchain<NT>(w2, idx, idy);
          // tangents w.r.t. the aberration coefficients (slots 2..26 = KrivanekCoeffs fields 0..24)
          for (int sI = 0; sI < gs.n; ++sI) {
            if (gs.s[sI].comp != c || gs.s[sI].slot < 2 || gs.s[sI].slot > 26) continue;
            const KrivOut dp = krivanek_poly_dfield(cm.p + 1, gs.s[sI].slot - 2, idx.v, idy.v);
            const double wgt = gs.s[sI].weight;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
              if (gs.s[sI].lane == k) {
                dWx.t[k] += wgt * dp.Gx;
                dWy.t[k] += wgt * dp.Gy;
                W.t[k] += wgt * dp.W;
              }
            }
          }
          dx = idx + (-dWx) / f;
          dy = idy + (-dWy) / f;
          pl = pl - (x * x + y * y) / (2.0 * f) + W / f;
          one = one * 1.0;
        }
      } break;
      default:
        break;
    }
  }
  if (out.ptr[0]) out.ptr[0][i] = x.v;
  if (out.ptr[1]) out.ptr[1][i] = y.v;
  if (out.ptr[2]) out.ptr[2][i] = dx.v;
  if (out.ptr[3]) out.ptr[3][i] = dy.v;
  if (out.ptr[4]) out.ptr[4][i] = z.v;
  if (out.ptr[5]) out.ptr[5][i] = pl.v;
  if (out.ptr[6]) out.ptr[6][i] = one.v;
  double *row = jac + i * (7 * NT);
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    row[0 * NT + k] = x.t[k];
    row[1 * NT + k] = y.t[k];
    row[2 * NT + k] = dx.t[k];
    row[3 * NT + k] = dy.t[k];
    row[4 * NT + k] = z.t[k];
    row[5 * NT + k] = pl.t[k];
    row[6 * NT + k] = one.t[k];
  }
}

// ------------------------------------------------------------------ transfer_rays
// transfer_rays (reference transfer.py:6-54): out[n, m, i] = sum_j T[m, i, j] * rays[n, j] for
// M <= TG_MAX_TRANSFER cumulative 5x5 matrices held in kernel-parameter constant memory.
struct TransferMats {
  int m;
  double t[TG_MAX_TRANSFER][25];
};
__global__ void __launch_bounds__(128)
    transfer_rays_kernel(const __grid_constant__ TransferMats tm, long long n,
                         const double *__restrict__ rays, double *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) r[j] = rays[i * 5 + j];
  double *o = out + i * (long long)tm.m * 5;
  for (int m = 0; m < tm.m; ++m) {
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      double acc = tm.t[m][a * 5] * r[0];
#pragma unroll
      for (int j = 1; j < 5; ++j) acc = acc + tm.t[m][a * 5 + j] * r[j];
      o[m * 5 + a] = acc;
    }
  }
}

// ------------------------------------------------------------------ Krivanek aberration function
// W_krivanek / grad_W_krivanek (aberrations.py:51-108) on arrays of slopes: the same value arithmetic
// the ray kernel uses for AberratedLensKrivanek.  kp = 25 coefficients + 11 (cos, sin)(m phi0) pairs.
struct KrivParams {
  double p[47];
};
__global__ void __launch_bounds__(256)
    krivanek_kernel(const __grid_constant__ KrivParams kp, long long n, const double *__restrict__ ax,
                    const double *__restrict__ ay, double *__restrict__ W, double *__restrict__ dWx,
                    double *__restrict__ dWy) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Dual<0> gx, gy, w;
  krivanek<0>(kp.p, dconst<0>(ax[i]), dconst<0>(ay[i]), gx, gy, w);
  if (W) W[i] = w.v;
  if (dWx) dWx[i] = gx.v;
  if (dWy) dWy[i] = gy.v;
}

// ------------------------------------------------------------------ K5
__global__ void __launch_bounds__(256)
    m2p_kernel(long long n, const double *__restrict__ x, const double *__restrict__ y,
               double m00, double m01, double m02, double m10, double m11, double m12,
               void *__restrict__ py, void *__restrict__ px, int as_float) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xv = x[i], yv = y[i];
  // apply_transformation (coordinate_transforms.py:137-140): T @ [y, x, 1], left to right,
  // no FMA contraction (this TU is built with -fmad=false; the intrinsics make it explicit)
  const double ty = __dadd_rn(__dadd_rn(__dmul_rn(m00, yv), __dmul_rn(m01, xv)), m02);
  const double tx = __dadd_rn(__dadd_rn(__dmul_rn(m10, yv), __dmul_rn(m11, xv)), m12);
  if (as_float) {
    static_cast<double *>(py)[i] = ty;
    static_cast<double *>(px)[i] = tx;
  } else {
    // jnp.round(...).astype(int32) (grid.py:147-149): half-to-even, saturating, NaN -> 0
    static_cast<int32_t *>(py)[i] = isnan(ty) ? 0 : __double2int_rn(ty);
    static_cast<int32_t *>(px)[i] = isnan(tx) ? 0 : __double2int_rn(tx);
  }
}

__global__ void __launch_bounds__(256)
    into_image_kernel(long long n, const int32_t *__restrict__ py, const int32_t *__restrict__ px,
                      int H, int W, unsigned long long *__restrict__ image) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = py[i], c = px[i];
  if (r >= 0 && r < H && c >= 0 && c < W) atomicAdd(image + (size_t)r * W + c, 1ULL);
}

}  // namespace

extern "C" int tg_trace_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                            double *const out[7], double *jac, int jac_layout, void *stream) {
  return tg_trace_perray_f64(model_host, n, in, nullptr, out, jac, jac_layout, stream);
}

extern "C" int tg_trace_perray_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                                   const tg_perray *perray, double *const out[7], double *jac, int jac_layout,
                                   void *stream) {
  TG_REQUIRE(model_host && in, "null model or input");
  if (perray) {
    TG_REQUIRE(perray->n >= 0 && perray->n <= TG_MAX_PERRAY, "bad per-ray component count");
    for (int k = 0; k < perray->n; ++k) {
      const int c = perray->comp[k];
      TG_REQUIRE(c >= 0 && c < model_host->n_comp && model_host->comp[c].op == TG_OP_OFFSET,
                 "per-ray parameters are supported for Scanner / Descanner components (TG_OP_OFFSET)");
    }
  }
  TG_REQUIRE(model_host->n_comp >= 0 && model_host->n_comp <= TG_MAX_COMPS, "bad n_comp");
  TG_REQUIRE(n >= 0, "negative n");
  TG_REQUIRE(jac_layout == TG_JAC_NONE || jac != nullptr, "jac requested but pointer is null");
  if (n == 0) return TG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (jac_layout) {
    case TG_JAC_NONE:
      return launch_trace<0>(model_host, n, in, out, nullptr, st, perray);
    case TG_JAC_ABCD5:
      return launch_trace<5>(model_host, n, in, out, jac, st, perray);
    case TG_JAC_FULL7:
      return launch_trace<7>(model_host, n, in, out, jac, st, perray);
    default:
      tg_set_error("tg_trace_f64: unknown jac_layout %d", jac_layout);
      return TG_EINVAL;
  }
}

extern "C" int tg_metres_to_pixels(int64_t n, const double *x, const double *y,
                                   const double m2px[9], void *py, void *px, int as_float,
                                   void *stream) {
  TG_REQUIRE(n >= 0 && m2px, "bad arguments");
  if (n == 0) return TG_OK;
  TG_REQUIRE(x && y && py && px, "null pointer");
  const long long blocks = (n + 255) / 256;
  m2p_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, x, y, m2px[0], m2px[1], m2px[2], m2px[3], m2px[4], m2px[5], py, px, as_float);
  return tg_launch_check("m2p_kernel");
}

extern "C" int tg_into_image_i64(int64_t n, const int32_t *py, const int32_t *px, int H, int W,
                                 long long *image, void *stream) {
  TG_REQUIRE(n >= 0 && H > 0 && W > 0, "bad arguments");
  if (n == 0) return TG_OK;
  TG_REQUIRE(py && px && image, "null pointer");
  const long long blocks = (n + 255) / 256;
  into_image_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, py, px, H, W, reinterpret_cast<unsigned long long *>(image));
  return tg_launch_check("into_image_kernel");
}

extern "C" int tg_trace_grad_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                                 const int32_t ray_lane[7], const tg_seed *seeds, int n_seeds,
                                 double *const out[7], double *jac, void *stream) {
  TG_REQUIRE(model_host && in && ray_lane && jac, "null pointer");
  TG_REQUIRE(model_host->n_comp >= 0 && model_host->n_comp <= TG_MAX_COMPS, "bad n_comp");
  TG_REQUIRE(n >= 0 && n_seeds >= 0 && n_seeds <= TG_MAX_SEEDS, "bad sizes");
  TG_REQUIRE(n_seeds == 0 || seeds, "null seeds");
  if (n == 0) return TG_OK;
  GradSeeds gs;
  gs.n = n_seeds;
  for (int f = 0; f < 7; ++f) {
    TG_REQUIRE(ray_lane[f] >= -1 && ray_lane[f] < TG_GRAD_LANES, "ray lane out of range");
    gs.ray_lane[f] = ray_lane[f];
  }
  bool kriv = false;
  for (int c = 0; c < model_host->n_comp; ++c) kriv |= (model_host->comp[c].op == TG_OP_KRIVANEK);
  for (int i = 0; i < n_seeds; ++i) {
    const tg_seed &sd = seeds[i];
    TG_REQUIRE(sd.comp >= 0 && sd.comp < model_host->n_comp, "seed component out of range");
    TG_REQUIRE(sd.slot >= 0 && sd.slot <= TG_NPARAM, "seed slot out of range");
    TG_REQUIRE(sd.lane >= 0 && sd.lane < TG_GRAD_LANES, "seed lane out of range");
    if (model_host->comp[sd.comp].op == TG_OP_KRIVANEK && sd.slot > 26) {
      // slots 2..26 are the 25 KrivanekCoeffs fields; 27.. hold derived (cos, sin)(m phi0) pairs
      tg_set_error("Krivanek slots above 26 are derived constants: seed the phi field (slots 2..26) instead");
      return TG_EUNSUPPORTED;
    }
    gs.s[i] = sd;
  }
  TraceOut o;
  for (int f = 0; f < 7; ++f) o.ptr[f] = out ? out[f] : nullptr;
  const long long blocks = (n + kTraceThreads - 1) / kTraceThreads;
  TG_REQUIRE(blocks <= 0x7fffffffLL, "too many rays for one launch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kriv)
    trace_grad_kernel<true><<<(unsigned)blocks, kTraceThreads, 0, st>>>(*model_host, gs, *in, (long long)n, o, jac);
  else
    trace_grad_kernel<false><<<(unsigned)blocks, kTraceThreads, 0, st>>>(*model_host, gs, *in, (long long)n, o, jac);
  return tg_launch_check("trace_grad_kernel");
}

extern "C" int tg_krivanek_f64(int64_t n, const double *alpha_x, const double *alpha_y, const double coeffs[47],
                               double *W, double *dWx, double *dWy, void *stream) {
  TG_REQUIRE(n >= 0 && coeffs, "bad arguments");
  if (n == 0) return TG_OK;
  TG_REQUIRE(alpha_x && alpha_y, "null pointer");
  KrivParams kp;
  for (int k = 0; k < 47; ++k) kp.p[k] = coeffs[k];
  krivanek_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(kp, n, alpha_x, alpha_y,
                                                                                           W, dWx, dWy);
  return tg_launch_check("krivanek_kernel");
}

extern "C" int tg_transfer_rays_f64(int64_t n, const double *rays, int m, const double *matrices_host,
                                    double *out, void *stream) {
  TG_REQUIRE(n >= 0 && m >= 0 && m <= TG_MAX_TRANSFER, "bad sizes (m <= TG_MAX_TRANSFER)");
  if (n == 0 || m == 0) return TG_OK;
  TG_REQUIRE(rays && matrices_host && out, "null pointer");
  TransferMats tm;
  tm.m = m;
  for (int k = 0; k < m; ++k)
    for (int j = 0; j < 25; ++j) tm.t[k][j] = matrices_host[k * 25 + j];
  const long long blocks = (n + 127) / 128;
  transfer_rays_kernel<<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(tm, n, rays, out);
  return tg_launch_check("transfer_rays_kernel");
}
