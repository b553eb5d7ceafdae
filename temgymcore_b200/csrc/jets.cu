// Higher-order derivatives of run_to_end w.r.t. the input ray (reference run.py:119-147,
// `calculate_derivatives` = repeated jax.jacfwd; SURVEY.md section 8f rank 3: the "Taylor / jet ray
// polynomial" of BASELINE config C4).
//
// The reference nests forward-mode Jacobians and materialises Ray-of-Ray-of-Ray pytrees: the order-k
// entry d^k out_f / d in_a d in_b ... for all 7^(k+1) index combinations.  Here every thread
// evaluates the model ONCE in hyper-dual arithmetic with P nilpotent units (eps_u^2 = 0): seeding
// in_a += eps_1, in_b += eps_2, in_c += eps_3 makes the coefficient of eps_1 eps_2 eps_3 exactly
// d^3 out / da db dc (and the eps_1 eps_2 coefficient d^2 out / da db, ...) -- the same numbers nested
// jacfwd produces, without the 7^k redundancy: one thread per SORTED index tuple a <= b <= c over the
// six live inputs {x, y, dx, dy, z, _one} (56 tuples at order 3, 21 at order 2), 2^P doubles per
// state variable held in registers.  The derivative tensors are symmetric; each thread writes its
// entry to all index permutations of the dense (7,...,7) output the reference returns.
// pathlength as an INPUT is handled analytically (zero-initialised tensors): d pl_out / d pl_in = 1
// and every other derivative involving pl_in vanishes (components only ever add to pathlength,
// components.py:161-559, propagator.py:67-72).
//
// Built with -fmad=false like the ray kernel; the Krivanek lens enters through a per-ray table of the partial
// derivatives of its aberration function (same polynomial form as trace.cu), composed per thread.
#include <math.h>
#include "tg_common.cuh"

namespace {

// ---- hyper-dual numbers with P units: c[mask] multiplies prod_{u in mask} eps_u -------------
template <int P>
struct HD {
  static constexpr int M = 1 << P;
  double c[M];
};
template <int P>
__device__ __forceinline__ HD<P> hconst(double v) {
  HD<P> r;
  r.c[0] = v;
#pragma unroll
  for (int m = 1; m < HD<P>::M; ++m) r.c[m] = 0.0;
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator+(const HD<P> &a, const HD<P> &b) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) r.c[m] = a.c[m] + b.c[m];
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator-(const HD<P> &a, const HD<P> &b) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) r.c[m] = a.c[m] - b.c[m];
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator-(const HD<P> &a) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) r.c[m] = -a.c[m];
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator+(const HD<P> &a, double b) {
  HD<P> r = a;
  r.c[0] = a.c[0] + b;
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator-(double b, const HD<P> &a) {
  HD<P> r = -a;
  r.c[0] = b - a.c[0];
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator*(const HD<P> &a, double b) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) r.c[m] = a.c[m] * b;
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> operator*(double b, const HD<P> &a) {
  return a * b;
}
// product: subset convolution, r[m] = sum_{s subset m} a[s] b[m \ s]   (3^P terms).  The value c[0] is one
// rounded multiplication, like the ray kernel's (bit-faithful fp64 values, this TU is built with -fmad=false);
// the derivative coefficients are sums of products and use explicit FMAs: 27 fp64 instructions per product at
// order 3 instead of 46 (the parity tolerance of the derivative tensors is 1e-10, an FMA only rounds less)
// Written subset-outer so that consecutive FMAs belong to different coefficients: 2^P independent chains in
// flight instead of one (ncu, order 3 on the C4 column: fixed-latency dependency waits were 36 % of all stalls
// at 3.5 warps per scheduler).
template <int P>
__device__ __forceinline__ HD<P> operator*(const HD<P> &a, const HD<P> &b) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) r.c[m] = a.c[0] * b.c[m];
#pragma unroll
  for (int s = 1; s < HD<P>::M; ++s)
#pragma unroll
    for (int m = 1; m < HD<P>::M; ++m)
      if ((s & m) == s) r.c[m] = fma(a.c[s], b.c[m ^ s], r.c[m]);
  return r;
}
// c + a * b (a propagation step x + dx d): the value is mul-then-add like the ray kernel's, the derivative
// coefficients start their FMA chains from c
template <int P>
__device__ __forceinline__ HD<P> hmuladd(const HD<P> &a, const HD<P> &b, const HD<P> &c) {
  HD<P> r;
  r.c[0] = c.c[0] + a.c[0] * b.c[0];
#pragma unroll
  for (int m = 1; m < HD<P>::M; ++m) r.c[m] = fma(a.c[0], b.c[m], c.c[m]);
#pragma unroll
  for (int s = 1; s < HD<P>::M; ++s)
#pragma unroll
    for (int m = 1; m < HD<P>::M; ++m)
      if ((s & m) == s) r.c[m] = fma(a.c[s], b.c[m ^ s], r.c[m]);
  return r;
}
// a * a: each unordered pair {s, m \ s} once (14 multiplications at order 3 instead of 27); c[0] = a0 a0 and
// 2 (a0 a_m) are the same doubles the general product gives
template <int P>
__device__ __forceinline__ HD<P> hsqr(const HD<P> &a) {
  HD<P> r;
  r.c[0] = a.c[0] * a.c[0];
#pragma unroll
  for (int m = 1; m < HD<P>::M; ++m) {
    double acc = a.c[0] * a.c[m];
#pragma unroll
    for (int s = 1; s < HD<P>::M; ++s)
      if ((s & m) == s && s < (m ^ s)) acc = fma(a.c[s], a.c[m ^ s], acc);
    r.c[m] = 2.0 * acc;
  }
  return r;
}
// g(a) for a smooth scalar function with Taylor coefficients g0, g1, g2/2!, g3/3! at a.c[0]:
// g(a0 + n) = g0 + g1 n + (g2/2) n^2 + (g3/6) n^3, n nilpotent (n^(P+1) = 0)
template <int P>
__device__ __forceinline__ HD<P> hcompose(const HD<P> &a, double g0, double g1, double g2h, double g3s) {
  HD<P> n = a;
  n.c[0] = 0.0;
  HD<P> r = n * g1;
  r.c[0] = g0;
  if constexpr (P >= 2) {
    const HD<P> n2 = n * n;
    r = r + n2 * g2h;
    if constexpr (P >= 3) r = r + (n2 * n) * g3s;
  }
  return r;
}
template <int P>
__device__ __forceinline__ HD<P> hrecip(const HD<P> &a) {
  const double i = 1.0 / a.c[0], i2 = i * i;
  return hcompose(a, i, -i2, i2 * i, -(i2 * i2));
}
template <int P>
__device__ __forceinline__ HD<P> hsqrt(const HD<P> &a) {
  const double s = sqrt(a.c[0]), i = 1.0 / a.c[0];
  const double g1 = 0.5 * s * i;                     // 1 / (2 sqrt a)
  return hcompose(a, s, g1, -0.25 * g1 * i, 0.125 * g1 * i * i);
}
template <int P>
__device__ __forceinline__ HD<P> operator/(const HD<P> &a, double b) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) r.c[m] = a.c[m] / b;
  return r;
}

// ---- Krivanek aberration function (aberrations.py:42-108) -----------
enum {
  K_C10 = 0, K_C12, K_PHI12, K_C21, K_PHI21, K_C23, K_PHI23, K_C30, K_C32, K_PHI32, K_C34,
  K_PHI34, K_C41, K_PHI41, K_C43, K_PHI43, K_C45, K_PHI45, K_C50, K_C52, K_PHI52, K_C54,
  K_PHI54, K_C56, K_PHI56
};
// Polynomial form (see trace.cu): with w = u + i v, r2 = |w|^2 and z_nm = C_nm / (n + 1) exp(-i m phi_nm)
//     W = Re[ F0(w) + r2 F1(w) + r2^2 F2(w) + r2^3 F3 ],   F_b = sum of z_nm w^m over the terms with (n + 1 - m) / 2 = b
// (no hypot / arctan2 / reciprocal compositions).  This TU is built with -fmad=false; fma() is honoured.

// ---- the Krivanek lens inside the jet kernel: one table of partials per ray, composed per thread ---------
// hkrivanek above carries 2^P coefficients of every intermediate through ~13 hyper-dual products in EVERY
// thread of a ray (56 at order 3): 255 registers, spills, 3.3 ms per 2e5 rays.  But the aberration function
// depends on two variables only, (u, v) = the ray slope at the lens.  The jet kernel therefore evaluates the
// partial derivatives W_ij = d^(i+j) W / du^i dv^j, i + j <= P + 1, ONCE per ray (15 numbers at order 3),
// cooperatively, in scalar complex arithmetic, and every thread composes its hyper-dual slopes with that
// table: F(u0 + nx, v0 + ny) = sum_ij F_ij nx^i ny^j / (i! j!) for F = W, dW/du, dW/dv, with nx, ny the
// nilpotent parts -- three sparse products and four cubes instead of the whole polynomial in 8-coefficient
// arithmetic.
//
// The table comes from Wirtinger calculus.  With w = u + i v every term of the polynomial form is
// z w^a conj(w)^b (a = m + b, b = (n + 1 - m) / 2, z = C_nm / (n + 1) exp(-i m phi_nm)), so
//     G_pq = d_w^p d_wbar^q G = sum_terms z ff(a, p) ff(b, q) w^(a-p) conj(w)^(b-q),    ff = falling factorial,
// and, since d/du = d_w + d_wbar and d/dv = i (d_w - d_wbar) are real operators and W = Re G,
//     W_ij = Re[ i^j sum_{k<=i, l<=j} C(i,k) C(j,l) (-1)^(j-l) G_{k+l, (i-k)+(j-l)} ]
// (tests/test_krivanek_polynomial_cpu.py checks this restatement against sympy derivatives).
struct KTerm {
  int c, g, n, m;      // coefficient index, index of its (cos, sin)(m phi) pair (-1: m = 0), radial / azimuthal order
};
__constant__ KTerm kTerms[14] = {
    {K_C10, -1, 1, 0}, {K_C12, 0, 1, 2},  {K_C21, 2, 2, 1},  {K_C23, 4, 2, 3},  {K_C30, -1, 3, 0},
    {K_C32, 6, 3, 2},  {K_C34, 8, 3, 4},  {K_C41, 10, 4, 1}, {K_C43, 12, 4, 3}, {K_C45, 14, 4, 5},
    {K_C50, -1, 5, 0}, {K_C52, 16, 5, 2}, {K_C54, 18, 5, 4}, {K_C56, 20, 5, 6}};
__constant__ double kKappa[6] = {1.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2};     // 1 / n (index n + 1 is used below)
__constant__ double kKappa6 = 1.0 / 6.0;
__constant__ int kBinom[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};

// shared scratch of one ray: powers of w, the Wirtinger derivatives, the table of partials (5 x 5, i + j <= 4 used)
struct KrivRay {
  double wr[7], wi[7];
  double gr[25], gi[25];
  double t[25];
};
// shared by the CTA's rays: z ff(a, p) ff(b, q) per (term, job) -- depends on the lens only
struct KrivCoef {
  double cr[14 * 15], ci[14 * 15];
};
template <int P>
__host__ __device__ constexpr int kriv_jobs() { return (P + 2) * (P + 3) / 2; }
__host__ __device__ constexpr int kterm_a(int t) {        // exponents of w and conj(w) of term t (kTerms order)
  constexpr int n[14] = {1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5}, m[14] = {0, 2, 1, 3, 0, 2, 4, 1, 3, 5, 0, 2, 4, 6};
  return m[t] + (n[t] + 1 - m[t]) / 2;
}
__host__ __device__ constexpr int kterm_b(int t) {
  constexpr int n[14] = {1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5}, m[14] = {0, 2, 1, 3, 0, 2, 4, 1, 3, 5, 0, 2, 4, 6};
  return (n[t] + 1 - m[t]) / 2;
}
template <int P>
__device__ __forceinline__ void kriv_job_pq(int job, int &p, int &q) {      // job -> (p, q), p + q <= P + 1
  p = 0;
  q = job;
  while (p <= P + 1 && q > P + 1 - p) {
    q -= P + 2 - p;
    ++p;
  }
}

// Called by ALL threads of the CTA (barriers inside); `job` = this thread's index within its ray.
// Three short stages: (1) all threads fill the lens's coefficient table, one thread per ray the powers of w;
// (2) job (p, q) sums G_pq over the 14 terms, branch-free (absent terms have zero coefficients);
// (3) job (i, j) combines the G into W_ij.
template <int P, int NTHREADS>
__device__ __forceinline__ void kriv_table(const double *prm, double u0, double v0, int job, KrivRay &K, KrivCoef &C) {
  constexpr int NJ = kriv_jobs<P>();
  const double *g = prm + 25;
  for (int e = threadIdx.x; e < 14 * NJ; e += NTHREADS) {
    const int t = e / NJ;
    int p, q;
    kriv_job_pq<P>(e - t * NJ, p, q);
    const KTerm T = kTerms[t];
    const int b = (T.n + 1 - T.m) >> 1, a = T.m + b;
    double zr = 0.0, zi = 0.0;
    if (a >= p && b >= q) {
      int f = 1;
      for (int k = 0; k < p; ++k) f *= a - k;
      for (int k = 0; k < q; ++k) f *= b - k;
      const double s = prm[T.c] * (T.n == 5 ? kKappa6 : kKappa[T.n + 1]) * (double)f;
      zr = T.g >= 0 ? s * g[T.g] : s;
      zi = T.g >= 0 ? -(s * g[T.g + 1]) : 0.0;
    }
    C.cr[e] = zr;
    C.ci[e] = zi;
  }
  if (job == 0) {
    double r = 1.0, i = 0.0;
    K.wr[0] = 1.0;
    K.wi[0] = 0.0;
#pragma unroll
    for (int k = 1; k < 7; ++k) {
      const double nr = fma(r, u0, -(i * v0)), ni = fma(r, v0, i * u0);
      r = nr;
      i = ni;
      K.wr[k] = r;
      K.wi[k] = i;
    }
  }
  __syncthreads();
  int pi, qj;
  kriv_job_pq<P>(job, pi, qj);
  const bool mine = job < NJ;
  if (mine) {
    double ar = 0.0, ai = 0.0;
#pragma unroll
    for (int t = 0; t < 14; ++t) {
      const int ea = kterm_a(t) - pi, eb = kterm_b(t) - qj;          // negative: the coefficient is zero
      const int ia = ea < 0 ? 0 : ea, ib = eb < 0 ? 0 : eb;
      const double zr = C.cr[t * NJ + job], zi = C.ci[t * NJ + job];
      const double xr = K.wr[ia], xi = K.wi[ia], yr = K.wr[ib], yi = -K.wi[ib];
      const double mr = fma(xr, yr, -(xi * yi)), mi = fma(xr, yi, xi * yr);
      ar = fma(zr, mr, fma(-zi, mi, ar));
      ai = fma(zr, mi, fma(zi, mr, ai));
    }
    K.gr[pi * 5 + qj] = ar;
    K.gi[pi * 5 + qj] = ai;
  }
  __syncthreads();
  if (mine) {
    const int i = pi, j = qj;
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k <= i; ++k)
      for (int l = 0; l <= j; ++l) {
        const double c = (double)(kBinom[i][k] * kBinom[j][l]) * (((j - l) & 1) ? -1.0 : 1.0);
        const int e = (k + l) * 5 + (i - k) + (j - l);
        ar = fma(c, K.gr[e], ar);
        ai = fma(c, K.gi[e], ai);
      }
    const int r = j & 3;                                  // Re[i^j (ar + i ai)]
    K.t[i * 5 + j] = r == 0 ? ar : (r == 1 ? -ai : (r == 2 ? -ar : ai));
  }
  __syncthreads();
}

__host__ __device__ constexpr int hd_popc(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1); }
// product of two nilpotent hyper-duals whose coefficients vanish below DA / DB units; only masks with
// >= DA + DB units are produced (the others are never read)
template <int P, int DA, int DB>
__device__ __forceinline__ void nmul(const double (&a)[1 << P], const double (&b)[1 << P], double (&r)[1 << P],
                                     double scale) {
#pragma unroll
  for (int m = 1; m < (1 << P); ++m) {
    if (hd_popc(m) < DA + DB) continue;
    double acc = 0.0;
    bool first = true;
#pragma unroll
    for (int s = 1; s < (1 << P); ++s) {
      if ((s & m) != s || hd_popc(s) < DA || hd_popc(m ^ s) < DB) continue;
      acc = first ? a[s] * b[m ^ s] : fma(a[s], b[m ^ s], acc);
      first = false;
    }
    r[m] = acc * scale;
  }
}
// dx, dy hold the hyper-dual slopes (u, v) entering the aberration function; on return
// dx = u - dW/du / f, dy = v - dW/dv / f, pl += W / f  (components.py:192-215), T = the ray's table W_ij
template <int P>
__device__ __forceinline__ void kriv_compose(const double *T, double inv_f, HD<P> &dx, HD<P> &dy, HD<P> &pl) {
  constexpr int M = 1 << P;
  auto t = [&](int i, int j) -> double { return T[i * 5 + j]; };
  const bool on_axis = dx.c[0] == 0.0 && dy.c[0] == 0.0;
  double xx[M] = {0.0}, xy[M] = {0.0}, yy[M] = {0.0};                           // nx^2 / 2, nx ny, ny^2 / 2   (masks with >= 2 units)
  double x3 = 0.0, x2y = 0.0, xy2 = 0.0, y3 = 0.0;      // nx^3 / 6, nx^2 ny / 2, nx ny^2 / 2, ny^3 / 6   (full mask)
  if constexpr (P >= 2) {
    nmul<P, 1, 1>(dx.c, dx.c, xx, 0.5);
    nmul<P, 1, 1>(dx.c, dy.c, xy, 1.0);
    nmul<P, 1, 1>(dy.c, dy.c, yy, 0.5);
  }
  if constexpr (P >= 3) {
    double c[M];
    nmul<P, 2, 1>(xx, dx.c, c, 1.0 / 3.0);
    x3 = c[M - 1];
    nmul<P, 2, 1>(xx, dy.c, c, 1.0);
    x2y = c[M - 1];
    nmul<P, 2, 1>(yy, dx.c, c, 1.0);
    xy2 = c[M - 1];
    nmul<P, 2, 1>(yy, dy.c, c, 1.0 / 3.0);
    y3 = c[M - 1];
  }
  // F = sum F_ij (monomial ij); for W the F_ij are t(i, j), for dW/du t(i + 1, j), for dW/dv t(i, j + 1)
  auto series = [&](int oi, int oj, int m) -> double {
    double v = fma(t(1 + oi, oj), dx.c[m], t(oi, 1 + oj) * dy.c[m]);
    if (P >= 2 && hd_popc(m) >= 2)
      v = fma(t(2 + oi, oj), xx[m], fma(t(1 + oi, 1 + oj), xy[m], fma(t(oi, 2 + oj), yy[m], v)));
    if (P >= 3 && hd_popc(m) >= 3)
      v = fma(t(3 + oi, oj), x3, fma(t(2 + oi, 1 + oj), x2y, fma(t(1 + oi, 2 + oj), xy2, fma(t(oi, 3 + oj), y3, v))));
    return v;
  };
  const double qnan = nan("");
  pl.c[0] = pl.c[0] + t(0, 0) * inv_f;
  const double u0 = dx.c[0], v0 = dy.c[0];
  double ndx[M], ndy[M];
#pragma unroll
  for (int m = 1; m < M; ++m) {
    // the reference's hypot / arctan2 have NaN derivatives at the origin (aberrations.py:66-67): values stay finite
    const double w = on_axis ? qnan : series(0, 0, m), gx = on_axis ? qnan : series(1, 0, m),
                 gy = on_axis ? qnan : series(0, 1, m);
    pl.c[m] = pl.c[m] + w * inv_f;
    ndx[m] = dx.c[m] - gx * inv_f;
    ndy[m] = dy.c[m] - gy * inv_f;
  }
#pragma unroll
  for (int m = 1; m < M; ++m) {
    dx.c[m] = ndx[m];
    dy.c[m] = ndy[m];
  }
  dx.c[0] = u0 - t(1, 0) * inv_f;
  dy.c[0] = v0 - t(0, 1) * inv_f;
}

// ---- index tuples ------------------------------------------------------------------------
// live input variables in Ray field order (pathlength = 5 is analytic, see the header)
__constant__ int kVar[6] = {0, 1, 2, 3, 4, 6};

template <int P>
struct Tuples;
template <>
struct Tuples<1> {
  static constexpr int N = 6;
  __device__ static void decode(int t, int v[1]) { v[0] = t; }
};
template <>
struct Tuples<2> {
  static constexpr int N = 21;
  __device__ static void decode(int t, int v[2]) {
    int a = 0;
    while (t >= 6 - a) { t -= 6 - a; ++a; }
    v[0] = a;
    v[1] = a + t;
  }
};
template <>
struct Tuples<3> {
  static constexpr int N = 56;
  __device__ static void decode(int t, int v[3]) {
    int a = 0;
    for (;;) {                                         // tuples starting with a: C(6 - a + 1, 2)
      const int na = (6 - a) * (7 - a) / 2;
      if (t < na) break;
      t -= na;
      ++a;
    }
    int b = a;
    while (t >= 6 - b) { t -= 6 - b; ++b; }
    v[0] = a;
    v[1] = b;
    v[2] = b + t;
  }
};

struct JetOut {
  double *ray[7];
  double *d[3];       // d[k-1]: (n, 7, 7^k) dense derivative tensors, fully written by the kernel
};

// A CTA owns kRays<P> consecutive rays (one thread per ray x tuple).  The dense tensors of those rays
// (22.3 KB per ray at order 3) are assembled in shared memory -- zero fill, scattered symmetric
// writes -- and leave as one contiguous, fully coalesced stream per order: the kernel's HBM traffic is
// its algorithmic output (no memset pass, no 8-byte scatter to global).
static_assert((49 * 8 * 8) % 16 == 0 && (4 * 49 * 8) % 16 == 0 && (4 * 343 * 8) % 16 == 0 && (2 * 49 * 8) % 16 == 0 &&
                  (2 * 343 * 8) % 16 == 0 && (2 * 2401 * 8) % 16 == 0,
              "a full CTA's blocks are whole 16-byte units (bulk copies)");
template <int P>
struct JetCfg;
template <>
struct JetCfg<1> { static constexpr int kRays = 8; };
template <>
struct JetCfg<2> { static constexpr int kRays = 4; };
template <>
struct JetCfg<3> { static constexpr int kRays = 2; };
template <int P>
__host__ __device__ constexpr int jet_doubles_per_ray() { return 49 + (P >= 2 ? 343 : 0) + (P >= 3 ? 2401 : 0); }

// KRIV = the model holds an AberratedLensKrivanek (its table of partials is built in shared memory, see above).
template <int P, bool KRIV>
__global__ void __launch_bounds__(Tuples<P>::N * JetCfg<P>::kRays, P == 3 ? 4 : 5)
    jets_kernel(const __grid_constant__ tg_model model, const tg_ray_in in, const long long n, const JetOut out) {
  using S = HD<P>;
  constexpr int NT = Tuples<P>::N, R = JetCfg<P>::kRays, NTHREADS = NT * R;
  extern __shared__ __align__(16) double s_t[];        // [order][ray][7^(order+1)]
  double *s_d1 = s_t, *s_d2 = s_d1 + R * 49, *s_d3 = s_d2 + (P >= 2 ? R * 343 : 0);
  const int rl = threadIdx.x / NT;
  const int tup = threadIdx.x - rl * NT;
  const long long i0 = (long long)blockIdx.x * R;
  const long long i = i0 + rl;
  const bool active = i < n;
  const long long il = active ? i : n - 1;             // threads past the end shadow the last ray (barriers below)
  __shared__ KrivRay s_kriv[KRIV ? R : 1];
  __shared__ KrivCoef s_kcoef;
  int var[P];
  S st[7];
  for (int k = threadIdx.x; k < R * jet_doubles_per_ray<P>(); k += NTHREADS) s_t[k] = 0.0;   // (barrier below)

  {
    Tuples<P>::decode(tup, var);
    auto ld = [&](int f) -> double { return in.ptr[f] ? __ldg(in.ptr[f] + il) : in.value[f]; };
#pragma unroll
    for (int f = 0; f < 7; ++f) st[f] = hconst<P>(ld(f));
#pragma unroll
    for (int u = 0; u < P; ++u) {
      const int f = kVar[var[u]];
#pragma unroll
      for (int k = 0; k < 7; ++k)
        if (k == f) st[k].c[1 << u] += 1.0;
    }
    S &x = st[0], &y = st[1], &dx = st[2], &dy = st[3], &z = st[4], &pl = st[5], &one = st[6];

    const int nc = model.n_comp;
    for (int c = 0; c < nc; ++c) {
      const tg_comp &cm = model.comp[c];
      if (!(cm.flags & TG_F_NOPROP)) {   // run.py:77, propagator.py:67-72
        const S d = (cm.flags & TG_F_DIST) ? hconst<P>(cm.z) : cm.z - z;
        x = hmuladd(dx, d, x);
        y = hmuladd(dy, d, y);
        z = z + d;
        pl = pl + d;
      }
      switch (cm.op) {
        case TG_OP_LENS:
        case TG_OP_THICKLENS: {   // components.py:161-174, 431-447
          // one reciprocal instead of 24 fp64 divisions (<= 1 ulp per coefficient; parity is 1e-10)
          const double inv_f = 1.0 / cm.p[0];
          const S ndx = dx - x * inv_f, ndy = dy - y * inv_f;
          pl = pl - (hsqr(x) + hsqr(y)) * (0.5 * inv_f);
          dx = ndx;
          dy = ndy;
          one = one * 1.0;
          if (cm.op == TG_OP_THICKLENS) z = z + (-cm.p[1]);
        } break;
        case TG_OP_DEFLECTOR: {   // components.py:476-482
          pl = pl + dx * x + dy * y;
          dx = dx + one * cm.p[0];
          dy = dy + one * cm.p[1];
        } break;
        case TG_OP_BIPRISM: {     // components.py:553-559; jnp.sign has zero derivative, sign(0) = 0
          pl = pl + dx * x + dy * y;
          const double xv = x.c[0];
          const double sg = xv > 0.0 ? 1.0 : (xv < 0.0 ? -1.0 : (xv == 0.0 ? 0.0 : xv));
          dx = dx + one * (cm.p[0] * sg);
        } break;
        case TG_OP_OFFSET: {      // Scanner / Descanner, components.py:279-285, 343-372
          x = x + one * cm.p[0];
          y = y + one * cm.p[1];
          dx = dx + one * cm.p[2];
          dy = dy + one * cm.p[3];
        } break;
        case TG_OP_ROTATOR: {     // components.py:503-523
          const double cs = cm.p[0], sn = cm.p[1];
          const S nx = x * cs - y * sn, ny = x * sn + y * cs;
          const S ndx = dx * cs - dy * sn, ndy = dx * sn + dy * cs;
          x = nx;
          y = ny;
          dx = ndx;
          dy = ndy;
        } break;
        case TG_OP_KRIVANEK: if constexpr (KRIV) {    // components.py:192-215
          const double inv_f = 1.0 / cm.p[0];
          const S idx = (-x) * inv_f + dx, idy = (-y) * inv_f + dy;
          pl = pl - (hsqr(x) + hsqr(y)) * (0.5 * inv_f);
          dx = idx;
          dy = idy;
          kriv_table<P, NTHREADS>(cm.p + 1, dx.c[0], dy.c[0], tup, s_kriv[rl], s_kcoef);      // barriers: every thread gets here
          kriv_compose<P>(s_kriv[rl].t, inv_f, dx, dy, pl);
          one = one * 1.0;
        } break;
        default:
          break;
      }
    }

  }
  __syncthreads();
  if (active) {
    // ---- outputs.  Sorted tuple (v0 <= v1 <= v2) in live-variable numbering; w_u = Ray field index.
    int w[P];
#pragma unroll
    for (int u = 0; u < P; ++u) w[u] = kVar[var[u]];
    if (tup == 0) {
#pragma unroll
      for (int f = 0; f < 7; ++f)
        if (out.ray[f]) out.ray[f][i] = st[f].c[0];
      s_d1[rl * 49 + 5 * 7 + 5] = 1.0;                    // d pl_out / d pl_in
    }
    // order P: the full-mask coefficient, written to every permutation of the indices
    {
      double *T = (P == 1 ? s_d1 + rl * 49 : (P == 2 ? s_d2 + rl * 343 : s_d3 + rl * 2401));
      constexpr int FS = P == 1 ? 7 : (P == 2 ? 49 : 343);
#pragma unroll
      for (int f = 0; f < 7; ++f) {
        const double v = st[f].c[HD<P>::M - 1];
        if constexpr (P == 1) {
          T[f * FS + w[0]] = v;
        } else if constexpr (P == 2) {
          T[f * FS + w[0] * 7 + w[1]] = v;
          T[f * FS + w[1] * 7 + w[0]] = v;
        } else {
          const int a = w[0], b = w[1], cc = w[2];
          T[f * FS + (a * 7 + b) * 7 + cc] = v;
          T[f * FS + (a * 7 + cc) * 7 + b] = v;
          T[f * FS + (b * 7 + a) * 7 + cc] = v;
          T[f * FS + (b * 7 + cc) * 7 + a] = v;
          T[f * FS + (cc * 7 + a) * 7 + b] = v;
          T[f * FS + (cc * 7 + b) * 7 + a] = v;
        }
      }
    }
    // lower orders come for free: the tuple (a, b, b) owns d2/da db (eps_1 eps_2), the tuple
    // (a, a[, a]) owns d/da (eps_1)
    if constexpr (P == 3) {
      if (var[1] == var[2]) {
        double *T = s_d2 + rl * 343;
#pragma unroll
        for (int f = 0; f < 7; ++f) {
          T[f * 49 + w[0] * 7 + w[1]] = st[f].c[3];
          T[f * 49 + w[1] * 7 + w[0]] = st[f].c[3];
        }
      }
    }
    if constexpr (P >= 2) {
      bool diag = true;
#pragma unroll
      for (int u = 1; u < P; ++u) diag = diag && (var[u] == var[0]);
      if (diag) {
        double *T = s_d1 + rl * 49;
#pragma unroll
        for (int f = 0; f < 7; ++f) T[f * 7 + w[0]] = st[f].c[1];
      }
    }
  }
  // contiguous emission of this CTA's rays, one stream per order: a full CTA (R rays: every size and offset is a
  // multiple of 16 bytes) hands the three blocks to the TMA unit as bulk shared -> global copies, so no thread
  // spends issue slots on the 5600 load / store pairs per CTA of order 3; a ragged last CTA or an output pointer
  // that is only 8-byte aligned takes the thread loop
  const long long rem = n - i0;
  const int nr = rem >= R ? R : (int)rem;
  bool bulk_ok = nr == R;
#pragma unroll
  for (int k = 0; k < P; ++k) bulk_ok = bulk_ok && ((reinterpret_cast<uintptr_t>(out.d[k]) & 15u) == 0);
  if (bulk_ok) {
    tg_fence_proxy_async();            // generic-proxy writes of this thread -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      tg_bulk_s2g(out.d[0] + i0 * 49, s_d1, (uint32_t)(R * 49 * sizeof(double)));
      if constexpr (P >= 2) tg_bulk_s2g(out.d[1] + i0 * 343, s_d2, (uint32_t)(R * 343 * sizeof(double)));
      if constexpr (P >= 3) tg_bulk_s2g(out.d[2] + i0 * 2401, s_d3, (uint32_t)(R * 2401 * sizeof(double)));
      tg_bulk_commit();
      tg_bulk_wait_read0();            // shared memory must stay valid until the TMA unit has read it
    }
    return;
  }
  __syncthreads();
  {
    double *g = out.d[0] + i0 * 49;
    for (int k = threadIdx.x; k < nr * 49; k += NTHREADS) g[k] = s_d1[k];
  }
  if constexpr (P >= 2) {
    double *g = out.d[1] + i0 * 343;
    for (int k = threadIdx.x; k < nr * 343; k += NTHREADS) g[k] = s_d2[k];
  }
  if constexpr (P >= 3) {
    double *g = out.d[2] + i0 * 2401;
    for (int k = threadIdx.x; k < nr * 2401; k += NTHREADS) g[k] = s_d3[k];
  }
}

template <int P, bool KRIV>
int launch_jets_k(const tg_model *m, int64_t n, const tg_ray_in *in, const JetOut &o, cudaStream_t st) {
  constexpr int R = JetCfg<P>::kRays;
  const long long blocks = (n + R - 1) / R;
  TG_REQUIRE(blocks <= 0x7fffffffLL, "too many rays for one launch");
  const size_t smem = (size_t)R * jet_doubles_per_ray<P>() * sizeof(double);
  static bool attr_set = false;
  if ((KRIV || smem > 40 * 1024) && !attr_set) {      // dynamic + static (Krivanek tables) can pass 48 KB
    TG_CUDA(cudaFuncSetAttribute(jets_kernel<P, KRIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  jets_kernel<P, KRIV><<<(unsigned)blocks, Tuples<P>::N * R, smem, st>>>(*m, *in, (long long)n, o);
  return tg_launch_check("jets_kernel");
}
template <int P>
int launch_jets(const tg_model *m, int64_t n, const tg_ray_in *in, const JetOut &o, cudaStream_t st) {
  bool kriv = false;
  for (int c = 0; c < m->n_comp; ++c) kriv |= (m->comp[c].op == TG_OP_KRIVANEK);
  return kriv ? launch_jets_k<P, true>(m, n, in, o, st) : launch_jets_k<P, false>(m, n, in, o, st);
}

}  // namespace

extern "C" int tg_trace_jets_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in, int order,
                                 double *const out[7], double *d1, double *d2, double *d3, void *stream) {
  TG_REQUIRE(model_host && in, "null model or input");
  TG_REQUIRE(model_host->n_comp >= 0 && model_host->n_comp <= TG_MAX_COMPS, "bad n_comp");
  TG_REQUIRE(n >= 0, "negative n");
  if (order < 1 || order > 3) {
    tg_set_error("tg_trace_jets_f64: derivative order %d is not implemented (1..3)", order);
    return TG_EUNSUPPORTED;
  }
  TG_REQUIRE(d1 && (order < 2 || d2) && (order < 3 || d3), "derivative tensors up to `order` must be given");
  if (n == 0) return TG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  JetOut o;
  for (int f = 0; f < 7; ++f) o.ray[f] = out ? out[f] : nullptr;
  o.d[0] = d1;
  o.d[1] = order >= 2 ? d2 : nullptr;
  o.d[2] = order >= 3 ? d3 : nullptr;
  switch (order) {
    case 1: return launch_jets<1>(model_host, n, in, o, st);
    case 2: return launch_jets<2>(model_host, n, in, o, st);
    default: return launch_jets<3>(model_host, n, in, o, st);
  }
}
