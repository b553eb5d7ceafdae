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
// Built with -fmad=false like the ray kernel; the Krivanek lens uses the same algebraic harmonic
// evaluation (powers of the unit phasor) as trace.cu.
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
// product: subset convolution, r[m] = sum_{s subset m} a[s] b[m \ s]   (3^P terms)
template <int P>
__device__ __forceinline__ HD<P> operator*(const HD<P> &a, const HD<P> &b) {
  HD<P> r;
#pragma unroll
  for (int m = 0; m < HD<P>::M; ++m) {
    double acc = a.c[0] * b.c[m];
#pragma unroll
    for (int s = 1; s < HD<P>::M; ++s)
      if ((s & m) == s) acc = acc + a.c[s] * b.c[m ^ s];
    r.c[m] = acc;
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

// ---- Krivanek aberration function in hyper-dual arithmetic (aberrations.py:42-108) -----------
enum {
  K_C10 = 0, K_C12, K_PHI12, K_C21, K_PHI21, K_C23, K_PHI23, K_C30, K_C32, K_PHI32, K_C34,
  K_PHI34, K_C41, K_PHI41, K_C43, K_PHI43, K_C45, K_PHI45, K_C50, K_C52, K_PHI52, K_C54,
  K_PHI54, K_C56, K_PHI56
};
template <int P>
__device__ __forceinline__ void hterm(double C, double m, const HD<P> &cm, const HD<P> &sm, double c0,
                                      double s0, HD<P> &B, HD<P> &T) {
  if (C == 0.0) return;
  B = B + (cm * c0 + sm * s0) * C;
  T = T + (sm * c0 - cm * s0) * (-m * C);
}
template <int P>
__device__ __forceinline__ void hkrivanek(const double *p, const HD<P> &ax, const HD<P> &ay, HD<P> &dWx,
                                          HD<P> &dWy, HD<P> &W) {
  using S = HD<P>;
  const S a = hsqrt(ax * ax + ay * ay);               // jnp.hypot; derivatives are NaN at the origin like JAX's
  S ia, c1, s1;                                       // unit phasor e^{i phi} = (ax + i ay) / |a|
  if (a.c[0] == 0.0) {
    // on-axis ray: phi = arctan2(0, 0) = 0 and alpha_safe = 1e-30 in the reference (aberrations.py:66,
    // 100-102), with NaN derivatives through hypot / arctan2 -- values stay finite, derivatives are NaN
    ia = hconst<P>(1e30);
    c1 = hconst<P>(1.0);
    s1 = hconst<P>(0.0);
#pragma unroll
    for (int m = 1; m < HD<P>::M; ++m) c1.c[m] = s1.c[m] = nan("");
  } else {
    ia = hrecip(a);
    c1 = ax * ia;
    s1 = ay * ia;
  }
  const bool u6 = p[K_C56] != 0.0, u5 = p[K_C45] != 0.0;
  const bool u4 = p[K_C34] != 0.0 || p[K_C54] != 0.0 || u5;
  const bool u3 = p[K_C23] != 0.0 || p[K_C43] != 0.0 || u6 || u5;
  const bool u2 = p[K_C12] != 0.0 || p[K_C32] != 0.0 || p[K_C52] != 0.0 || u3 || u4;
  S c2 = hconst<P>(0.0), s2 = c2, c3 = c2, s3 = c2, c4 = c2, s4 = c2, c5 = c2, s5 = c2, c6 = c2, s6 = c2;
  if (u2) { c2 = c1 * c1 - s1 * s1; s2 = (c1 * s1) * 2.0; }
  if (u3) { c3 = c2 * c1 - s2 * s1; s3 = s2 * c1 + c2 * s1; }
  if (u4) { c4 = c2 * c2 - s2 * s2; s4 = (c2 * s2) * 2.0; }
  if (u5) { c5 = c4 * c1 - s4 * s1; s5 = s4 * c1 + c4 * s1; }
  if (u6) { c6 = c3 * c3 - s3 * s3; s6 = (c3 * s3) * 2.0; }
  const double *g = p + 25;                           // (cos, sin)(m phi0) pairs
  // Radial orders n = 2..6 are folded into W, dW/dalpha and dW/dphi one at a time (aberrations.py:51-98)
  // so that a bracket B_n / T_n dies as soon as it is used: the live set stays at three accumulators,
  // the phasor powers and one power of alpha (order 3 needs 8 doubles per quantity).
  S an = a;                                           // alpha^(n-1)
  S dWa = hconst<P>(0.0), q = dWa;
  W = dWa;
  auto fold = [&](const S &Bn, const S &Tn, double inv_n) {
    dWa = dWa + an * Bn;                              // alpha^(n-1) B_n
    const S wn = (an * a) * inv_n;                    // alpha^n / n
    W = W + wn * Bn;
    q = q + wn * Tn;
    an = an * a;
  };
  {
    S B = hconst<P>(p[K_C10]), T = hconst<P>(0.0);
    hterm(p[K_C12], 2.0, c2, s2, g[0], g[1], B, T);
    fold(B, T, 0.5);
  }
  {
    S B = hconst<P>(0.0), T = B;
    hterm(p[K_C21], 1.0, c1, s1, g[2], g[3], B, T);
    hterm(p[K_C23], 3.0, c3, s3, g[4], g[5], B, T);
    fold(B, T, 1.0 / 3.0);
  }
  {
    S B = hconst<P>(p[K_C30]), T = hconst<P>(0.0);
    hterm(p[K_C32], 2.0, c2, s2, g[6], g[7], B, T);
    hterm(p[K_C34], 4.0, c4, s4, g[8], g[9], B, T);
    fold(B, T, 0.25);
  }
  {
    S B = hconst<P>(0.0), T = B;
    hterm(p[K_C41], 1.0, c1, s1, g[10], g[11], B, T);
    hterm(p[K_C43], 3.0, c3, s3, g[12], g[13], B, T);
    hterm(p[K_C45], 5.0, c5, s5, g[14], g[15], B, T);
    fold(B, T, 0.2);
  }
  {
    S B = hconst<P>(p[K_C50]), T = hconst<P>(0.0);
    hterm(p[K_C52], 2.0, c2, s2, g[16], g[17], B, T);
    hterm(p[K_C54], 4.0, c4, s4, g[18], g[19], B, T);
    hterm(p[K_C56], 6.0, c6, s6, g[20], g[21], B, T);
    fold(B, T, 1.0 / 6.0);
  }
  q = q * ia;
  dWx = dWa * c1 - q * s1;
  dWy = dWa * s1 + q * c1;
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

// KRIV = the model holds an AberratedLensKrivanek: that instantiation needs ~250 registers (and
// spills at order 3); every other model runs the lean instantiation at 4+ CTAs per SM.
template <int P, bool KRIV>
__global__ void __launch_bounds__(Tuples<P>::N * JetCfg<P>::kRays, KRIV ? 2 : (P == 3 ? 4 : 5))
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
  int var[P];
  S st[7];

  if (active) {
    Tuples<P>::decode(tup, var);
    auto ld = [&](int f) -> double { return in.ptr[f] ? __ldg(in.ptr[f] + i) : in.value[f]; };
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
        x = x + dx * d;
        y = y + dy * d;
        z = z + d;
        pl = pl + d;
      }
      switch (cm.op) {
        case TG_OP_LENS:
        case TG_OP_THICKLENS: {   // components.py:161-174, 431-447
          // one reciprocal instead of 24 fp64 divisions (<= 1 ulp per coefficient; parity is 1e-10)
          const double inv_f = 1.0 / cm.p[0];
          const S ndx = dx - x * inv_f, ndy = dy - y * inv_f;
          pl = pl - (x * x + y * y) * (0.5 * inv_f);
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
          pl = pl - (x * x + y * y) * (0.5 * inv_f);
          S dWx, dWy, W;
          hkrivanek<P>(cm.p + 1, idx, idy, dWx, dWy, W);
          dx = idx - dWx * inv_f;
          dy = idy - dWy * inv_f;
          pl = pl + W * inv_f;
          one = one * 1.0;
        } break;
        default:
          break;
      }
    }

  }
  for (int k = threadIdx.x; k < R * jet_doubles_per_ray<P>(); k += NTHREADS) s_t[k] = 0.0;
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
  __syncthreads();
  // contiguous, coalesced emission of this CTA's rays, one stream per order
  const long long rem = n - i0;
  const int nr = rem >= R ? R : (int)rem;
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
  if (smem > 48 * 1024 && !attr_set) {
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
