// K2: per-beamlet coefficient builders (one beamlet per thread, fp64 / complex fp64).
//
// Collapses the per-beamlet part of the reference's `_beam_field`
// (src/temgym_core/gaussian.py:276-316) and `Qinv_ABCD` (gaussian.py:92-96) into six
// complex coefficients of a quadratic in the observation point (x, y) [metres]:
//     field_n(x, y) = exp(i P_n(x, y)),
//     P_n = c0 + c1 x + c2 y + c3 x^2 + c4 x y + c5 y^2.
// Derivation with rt = r2 - e:
//     phase = (k/2) (rt^T Q2i rt + phi1 - phi2 + 2 rt.f) + p0,   field = pref exp(i phase)
//     phi1  = r1m^T (A Binv) r1m - 2 r1m^T Binv rt
//     phi2  = r2m^T Mq r2m      - 2 r2m^T Mq rt ,  Mq = Q1 (B (A Q1 + B))^-1
// so P(rt) = C + L.rt + rt^T Qs rt with
//     Qs = (k/2) Q2i,  L = (k/2)(-2 Binv^T r1m + 2 Mq^T r2m + 2 f),
//     C  = (k/2)(r1m^T A Binv r1m - r2m^T Mq r2m) + p0 - i log(pref)
// and the shift rt = r - e is expanded into the c_j.
//
// The 2x2 solves follow LU with partial pivoting (what jnp.linalg.solve / LAPACK gesv
// do), complex division follows numpy's Smith form, and this TU is compiled with
// -fmad=false, so every intermediate matches the numpy oracle to the last few ulps.
// Reference quirks kept: Binv and Q1 are nan_to_num'ed (gaussian.py:283-287), the
// inverse inside Mq is not (gaussian.py:305) -> B == 0 gives a NaN field.
#include <math.h>
#include "tg_common.cuh"

namespace {

struct cd {
  double re, im;
};
__device__ __forceinline__ cd mk(double re, double im = 0.0) { return cd{re, im}; }
__device__ __forceinline__ cd operator+(cd a, cd b) { return cd{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cd operator-(cd a, cd b) { return cd{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cd operator-(cd a) { return cd{-a.re, -a.im}; }
__device__ __forceinline__ cd operator*(cd a, cd b) {
  return cd{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ cd operator*(double a, cd b) { return cd{a * b.re, a * b.im}; }
__device__ __forceinline__ cd operator*(cd b, double a) { return cd{a * b.re, a * b.im}; }
__device__ __forceinline__ cd operator/(cd a, cd b) {
  // numpy complex division (Smith's method)
  if (fabs(b.re) >= fabs(b.im)) {
    if (b.re == 0.0 && b.im == 0.0) return cd{a.re / fabs(b.re), a.im / fabs(b.im)};
    const double rat = b.im / b.re;
    const double scl = 1.0 / (b.re + b.im * rat);
    return cd{(a.re + a.im * rat) * scl, (a.im - a.re * rat) * scl};
  } else {
    const double rat = b.re / b.im;
    const double scl = 1.0 / (b.im + b.re * rat);
    return cd{(a.re * rat + a.im) * scl, (a.im * rat - a.re) * scl};
  }
}
__device__ __forceinline__ double cabs_(cd a) { return hypot(a.re, a.im); }
__device__ __forceinline__ double cabs_(double a) { return fabs(a); }
__device__ __forceinline__ cd csqrt_(cd z) {
  // principal branch (gaussian.py:295 jnp.sqrt of a complex determinant)
  const double a = z.re, b = z.im;
  if (a == 0.0 && b == 0.0) return cd{0.0, b};
  if (isinf(b)) return cd{INFINITY, b};
  if (isnan(a)) return cd{a, a};
  const double h = hypot(a, b);
  if (a >= 0.0) {
    const double t = sqrt((a + h) * 0.5);
    return cd{t, b / (2.0 * t)};
  }
  const double t = sqrt((-a + h) * 0.5);
  return cd{fabs(b) / (2.0 * t), copysign(t, b)};
}
__device__ __forceinline__ double fin0(double v) { return isfinite(v) ? v : 0.0; }
__device__ __forceinline__ cd fin0(cd v) { return cd{fin0(v.re), fin0(v.im)}; }

template <typename T>
struct M2 {
  T a00, a01, a10, a11;
};
__device__ __forceinline__ cd tocd(double v) { return mk(v); }
__device__ __forceinline__ cd tocd(cd v) { return v; }

template <typename TA, typename TB>
__device__ __forceinline__ auto mm(const M2<TA> &a, const M2<TB> &b) {
  using TR = decltype(a.a00 * b.a00);
  M2<TR> r;
  r.a00 = a.a00 * b.a00 + a.a01 * b.a10;
  r.a01 = a.a00 * b.a01 + a.a01 * b.a11;
  r.a10 = a.a10 * b.a00 + a.a11 * b.a10;
  r.a11 = a.a10 * b.a01 + a.a11 * b.a11;
  return r;
}
__device__ __forceinline__ M2<cd> addm(const M2<double> &a, const M2<cd> &b) {
  return M2<cd>{mk(a.a00) + b.a00, mk(a.a01) + b.a01, mk(a.a10) + b.a10, mk(a.a11) + b.a11};
}
__device__ __forceinline__ M2<cd> addm(const M2<cd> &a, const M2<double> &b) { return addm(b, a); }

// solve(A, B) for 2x2 systems: LU with partial pivoting, singular -> inf/nan like JAX
template <typename T>
__device__ __forceinline__ M2<T> solve2(M2<T> A, M2<T> B) {
  if (cabs_(A.a10) > cabs_(A.a00)) {
    T t;
    t = A.a00; A.a00 = A.a10; A.a10 = t;
    t = A.a01; A.a01 = A.a11; A.a11 = t;
    t = B.a00; B.a00 = B.a10; B.a10 = t;
    t = B.a01; B.a01 = B.a11; B.a11 = t;
  }
  const T l10 = A.a10 / A.a00;
  const T u11 = A.a11 - l10 * A.a01;
  const T y10 = B.a10 - l10 * B.a00;
  const T y11 = B.a11 - l10 * B.a01;
  M2<T> X;
  X.a10 = y10 / u11;
  X.a11 = y11 / u11;
  X.a00 = (B.a00 - A.a01 * X.a10) / A.a00;
  X.a01 = (B.a01 - A.a01 * X.a11) / A.a00;
  return X;
}
__device__ __forceinline__ cd det2(M2<cd> A) {
  bool swap = cabs_(A.a10) > cabs_(A.a00);
  if (swap) {
    cd t;
    t = A.a00; A.a00 = A.a10; A.a10 = t;
    t = A.a01; A.a01 = A.a11; A.a11 = t;
  }
  const cd u11 = A.a11 - (A.a10 / A.a00) * A.a01;
  const cd d = A.a00 * u11;
  return swap ? -d : d;
}

__device__ __forceinline__ M2<double> ldm(const double *p) { return M2<double>{p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ M2<cd> ldmc(const double *p) {
  return M2<cd>{cd{p[0], p[1]}, cd{p[2], p[3]}, cd{p[4], p[5]}, cd{p[6], p[7]}};
}

// write c0..c5 for P(rt) = C + Lx X + Ly Y + q00 X^2 + qxy X Y + q11 Y^2, rt = r - e
__device__ __forceinline__ void emit_poly(double *out, cd C, cd Lx, cd Ly, cd q00, cd qxy, cd q11,
                                          double ex, double ey) {
  const cd c3 = q00, c4 = qxy, c5 = q11;
  const cd c1 = Lx - (2.0 * ex) * q00 - ey * qxy;
  const cd c2 = Ly - (2.0 * ey) * q11 - ex * qxy;
  const cd c0 = C - ex * Lx - ey * Ly + (ex * ex) * q00 + (ex * ey) * qxy + (ey * ey) * q11;
  out[0] = c0.re; out[1] = c0.im;
  out[2] = c1.re; out[3] = c1.im;
  out[4] = c2.re; out[5] = c2.im;
  out[6] = c3.re; out[7] = c3.im;
  out[8] = c4.re; out[9] = c4.im;
  out[10] = c5.re; out[11] = c5.im;
}

__device__ __forceinline__ void beamlet_poly(double amp, double p0, const M2<cd> &Q1i,
                                             const M2<double> &A, const M2<double> &B,
                                             const M2<double> &C, const M2<double> &D, double ex,
                                             double ey, double fx, double fy, double r1x,
                                             double r1y, double tx, double ty, double k,
                                             double *out) {
  const M2<double> I{1.0, 0.0, 0.0, 1.0};
  const M2<cd> Ic{mk(1.0), mk(0.0), mk(0.0), mk(1.0)};
  // safe inverses (gaussian.py:283-287)
  M2<double> Binv = solve2(B, I);
  Binv = M2<double>{fin0(Binv.a00), fin0(Binv.a01), fin0(Binv.a10), fin0(Binv.a11)};
  M2<cd> Q1 = solve2(Q1i, Ic);
  Q1 = M2<cd>{fin0(Q1.a00), fin0(Q1.a01), fin0(Q1.a10), fin0(Q1.a11)};
  // central ray at the output plane (gaussian.py:291)
  const double r2x = (A.a00 * r1x + A.a01 * r1y) + (B.a00 * tx + B.a01 * ty);
  const double r2y = (A.a10 * r1x + A.a11 * r1y) + (B.a10 * tx + B.a11 * ty);
  // amplitude prefactor (gaussian.py:294-295)
  const M2<cd> denom = addm(A, mm(B, Q1i));
  const cd pref = mk(amp) / csqrt_(det2(denom));
  // phi1 (gaussian.py:298-301)
  const M2<double> ABinv = mm(A, Binv);
  const double c1c = r1x * (ABinv.a00 * r1x + ABinv.a01 * r1y) + r1y * (ABinv.a10 * r1x + ABinv.a11 * r1y);
  const double l1x = r1x * Binv.a00 + r1y * Binv.a10;
  const double l1y = r1x * Binv.a01 + r1y * Binv.a11;
  // phi2 (gaussian.py:304-309)
  const M2<cd> AQ1 = mm(A, Q1);
  const M2<cd> inner = mm(B, addm(AQ1, B));
  const M2<cd> BoA = solve2(inner, Ic);  // NOT nan_to_num'ed in the reference
  const M2<cd> Mq = mm(Q1, BoA);
  const cd c2c = r2x * (Mq.a00 * r2x + Mq.a01 * r2y) + r2y * (Mq.a10 * r2x + Mq.a11 * r2y);
  const cd l2x = r2x * Mq.a00 + r2y * Mq.a10;
  const cd l2y = r2x * Mq.a01 + r2y * Mq.a11;
  // Q2^-1 (gaussian.py:92-96)
  const M2<cd> Q2i = solve2(addm(A, mm(B, Q1i)), addm(C, mm(D, Q1i)));
  const double hk = k / 2.0;
  const cd q00 = hk * Q2i.a00;
  const cd qxy = hk * (Q2i.a01 + Q2i.a10);
  const cd q11 = hk * Q2i.a11;
  const cd Lx = hk * (mk(-2.0 * l1x) + 2.0 * l2x + mk(2.0 * fx));
  const cd Ly = hk * (mk(-2.0 * l1y) + 2.0 * l2y + mk(2.0 * fy));
  // -i log(pref): Re += arg(pref), Im -= ln|pref|
  const cd Cc = hk * (mk(c1c) - c2c) + mk(p0) + cd{atan2(pref.im, pref.re), -log(cabs_(pref))};
  emit_poly(out, Cc, Lx, Ly, q00, qxy, q11, ex, ey);
}

__global__ void __launch_bounds__(128)
    coeffs_kernel(long long nb, const double *__restrict__ amp, const double *__restrict__ p0,
                  const double *__restrict__ Q1i, const double *__restrict__ A,
                  const double *__restrict__ B, const double *__restrict__ C,
                  const double *__restrict__ D, const double *__restrict__ e,
                  const double *__restrict__ f, const double *__restrict__ r1m,
                  const double *__restrict__ th, const double *__restrict__ k,
                  double *__restrict__ poly) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  beamlet_poly(amp[i], p0[i], ldmc(Q1i + i * 8), ldm(A + i * 4), ldm(B + i * 4), ldm(C + i * 4),
               ldm(D + i * 4), e[i * 2], e[i * 2 + 1], f[i * 2], f[i * 2 + 1], r1m[i * 2],
               r1m[i * 2 + 1], th[i * 2], th[i * 2 + 1], k[i], poly + i * 12);
}

__global__ void __launch_bounds__(128)
    coeffs_abcd_kernel(long long nb, const double *__restrict__ amp, const double *__restrict__ p0,
                       const double *__restrict__ Q1i, const double *__restrict__ abcd,
                       const double *__restrict__ r1x, const double *__restrict__ r1y,
                       const double *__restrict__ thx, const double *__restrict__ thy,
                       const double *__restrict__ k, double *__restrict__ poly) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const double *m = abcd + i * 25;  // (5,5) row-major: gaussian.py:244-249
  const M2<double> A{m[0], m[1], m[5], m[6]};
  const M2<double> B{m[2], m[3], m[7], m[8]};
  const M2<double> C{m[10], m[11], m[15], m[16]};
  const M2<double> D{m[12], m[13], m[17], m[18]};
  beamlet_poly(amp[i], p0[i], ldmc(Q1i + i * 8), A, B, C, D, m[4], m[9], m[14], m[19], r1x[i],
               r1y[i], thx[i], thy[i], k[i], poly + i * 12);
}

// _input_beam_field (gaussian.py:402-407): a exp(i k/2 (d^T Q1i d + 2 d.t)) exp(i p), d = r - r1m
__global__ void __launch_bounds__(128)
    input_coeffs_kernel(long long nb, const double *__restrict__ amp, const double *__restrict__ p0,
                        const double *__restrict__ Q1i, const double *__restrict__ r1m,
                        const double *__restrict__ th, const double *__restrict__ k,
                        double *__restrict__ poly) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const M2<cd> Q = ldmc(Q1i + i * 8);
  const double hk = k[i] / 2.0;
  const double a = amp[i];
  const cd Cc = mk(p0[i]) + cd{a < 0.0 ? 3.14159265358979323846 : 0.0, -log(fabs(a))};
  emit_poly(poly + i * 12, Cc, mk(hk * (2.0 * th[i * 2])), mk(hk * (2.0 * th[i * 2 + 1])),
            hk * Q.a00, hk * (Q.a01 + Q.a10), hk * Q.a11, r1m[i * 2], r1m[i * 2 + 1]);
}

// GaussianRay.q_inv / Q_inv (gaussian.py:138-177)
__device__ __forceinline__ M2<cd> qinv_of(double wx, double wy, double Rx, double Ry, double lam, double theta) {
  const double pi = 3.141592653589793;
  const cd qx = isinf(Rx) ? cd{0.0, lam / (pi * (wx * wx))} : cd{-1.0 / Rx, lam / (pi * (wx * wx))};
  const cd qy = isinf(Ry) ? cd{0.0, lam / (pi * (wy * wy))} : cd{-1.0 / Ry, lam / (pi * (wy * wy))};
  double s, c;
  sincos(theta, &s, &c);
  // einsum("nij,njk,npk->nip", R, diag, R), R = [[c,-s],[s,c]]
  M2<cd> Q;
  Q.a00 = (c * qx) * c + ((-s) * qy) * (-s);
  Q.a01 = (c * qx) * s + ((-s) * qy) * c;
  Q.a10 = (s * qx) * c + (c * qy) * (-s);
  Q.a11 = (s * qx) * s + (c * qy) * c;
  return Q;
}
__global__ void __launch_bounds__(128)
    qinv_kernel(long long nb, const double *__restrict__ waist, const double *__restrict__ radii,
                const double *__restrict__ wl, const double *__restrict__ theta,
                double *__restrict__ Qi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const M2<cd> Q = qinv_of(waist[i * 2], waist[i * 2 + 1], radii[i * 2], radii[i * 2 + 1], wl[i], theta[i]);
  double *o = Qi + i * 8;
  o[0] = Q.a00.re; o[1] = Q.a00.im; o[2] = Q.a01.re; o[3] = Q.a01.im;
  o[4] = Q.a10.re; o[5] = Q.a10.im; o[6] = Q.a11.re; o[7] = Q.a11.im;
}

// make_gaussian_image's per-beamlet chain in ONE kernel (gaussian.py:240-262): Q_inv from the beam
// parameters, k = 2 pi / lambda, p0 = k * pathlength, then the coefficients from the traced ABCD.
// Same device functions, same operation order as qinv_kernel -> wave_kernel -> coeffs_abcd_kernel.
__global__ void __launch_bounds__(128)
    coeffs_from_beam_kernel(long long nb, const double *__restrict__ amp, const double *__restrict__ pl,
                            const double *__restrict__ waist, const double *__restrict__ radii,
                            const double *__restrict__ wl, const double *__restrict__ theta,
                            const double *__restrict__ abcd, const double *__restrict__ r1x,
                            const double *__restrict__ r1y, const double *__restrict__ thx,
                            const double *__restrict__ thy, double *__restrict__ poly) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const M2<cd> Q = qinv_of(waist[i * 2], waist[i * 2 + 1], radii[i * 2], radii[i * 2 + 1], wl[i], theta[i]);
  const double kk = (2.0 * 3.141592653589793) / wl[i];  // gaussian.py:254
  const double p0 = kk * pl[i];                          // gaussian.py:255
  const double *m = abcd + i * 25;
  const M2<double> A{m[0], m[1], m[5], m[6]};
  const M2<double> B{m[2], m[3], m[7], m[8]};
  const M2<double> C{m[10], m[11], m[15], m[16]};
  const M2<double> D{m[12], m[13], m[17], m[18]};
  beamlet_poly(amp[i], p0, Q, A, B, C, D, m[4], m[9], m[14], m[19], r1x[i], r1y[i], thx[i], thy[i], kk,
               poly + i * 12);
}

__global__ void __launch_bounds__(128)
    wave_kernel(long long nb, const double *__restrict__ wl, const double *__restrict__ pl,
                double *__restrict__ k, double *__restrict__ p0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const double kk = (2.0 * 3.141592653589793) / wl[i];  // gaussian.py:254
  k[i] = kk;
  p0[i] = kk * pl[i];                                    // gaussian.py:255
}

inline unsigned nblocks(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

int tg_wave_numbers(int64_t nb, const double *wavelength, const double *pathlength, double *k,
                    double *p0, cudaStream_t st) {
  if (nb <= 0) return TG_OK;
  wave_kernel<<<nblocks(nb, 128), 128, 0, st>>>(nb, wavelength, pathlength, k, p0);
  return tg_launch_check("wave_kernel");
}

// fibonacci_spiral (reference utils.py:297-325): beamlet centres on a disc, generated where they are
// consumed (no host staging for 1e6-beamlet runs).  Same fp64 formula as the host sampler.
namespace {
__global__ void __launch_bounds__(256)
    fibonacci_kernel(long long n, double radius, double np_boundary, double *__restrict__ x, double *__restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ga = 3.14159265358979323846 * (3.0 - sqrt(5.0));
  const double fi = (double)i, fn = (double)n;
  double rr = (fi > fn - (np_boundary + 1.0)) ? radius
                                              : radius * sqrt((fi + 0.5) / (fn - 0.5 * (np_boundary + 1.0)));
  if (i == 0) rr = 0.0;
  double sn, cs;
  sincos(fi * ga, &sn, &cs);
  x[i] = rr * cs;
  y[i] = rr * sn;
}
}  // namespace

extern "C" int tg_fibonacci_spiral_f64(int64_t n, double radius, double alpha, double *x, double *y, void *stream) {
  TG_REQUIRE(n >= 0, "negative n");
  if (n == 0) return TG_OK;
  TG_REQUIRE(x && y, "null pointer");
  const double np_boundary = nearbyint(alpha * sqrt((double)n));    // np.round: half to even
  fibonacci_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, radius, np_boundary, x, y);
  return tg_launch_check("fibonacci_kernel");
}

extern "C" int tg_wave_numbers_f64(int64_t nb, const double *wavelength, const double *pathlength,
                                   double *k, double *phase_offset, void *stream) {
  TG_REQUIRE(nb >= 0, "negative nb");
  if (nb == 0) return TG_OK;
  TG_REQUIRE(wavelength && pathlength && k && phase_offset, "null pointer");
  return tg_wave_numbers(nb, wavelength, pathlength, k, phase_offset, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_beamlet_coeffs_f64(int64_t nb, const double *amp, const double *phase_offset,
                                     const double *Q1_inv, const double *A, const double *B,
                                     const double *C, const double *D, const double *e,
                                     const double *f, const double *r1m, const double *theta1m,
                                     const double *k, double *poly, void *stream) {
  TG_REQUIRE(nb >= 0, "negative nb");
  if (nb == 0) return TG_OK;
  TG_REQUIRE(amp && phase_offset && Q1_inv && A && B && C && D && e && f && r1m && theta1m && k && poly,
             "null pointer");
  coeffs_kernel<<<nblocks(nb, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      nb, amp, phase_offset, Q1_inv, A, B, C, D, e, f, r1m, theta1m, k, poly);
  return tg_launch_check("coeffs_kernel");
}

extern "C" int tg_beamlet_coeffs_abcd_f64(int64_t nb, const double *amp,
                                          const double *phase_offset, const double *Q1_inv,
                                          const double *abcd, const double *r1m_x,
                                          const double *r1m_y, const double *th_x,
                                          const double *th_y, const double *k, double *poly,
                                          void *stream) {
  TG_REQUIRE(nb >= 0, "negative nb");
  if (nb == 0) return TG_OK;
  TG_REQUIRE(amp && phase_offset && Q1_inv && abcd && r1m_x && r1m_y && th_x && th_y && k && poly,
             "null pointer");
  coeffs_abcd_kernel<<<nblocks(nb, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      nb, amp, phase_offset, Q1_inv, abcd, r1m_x, r1m_y, th_x, th_y, k, poly);
  return tg_launch_check("coeffs_abcd_kernel");
}

// internal (host_api.cu): the fused chain of tg_make_gaussian_image_f64
int tg_coeffs_from_beam(int64_t nb, const double *amp, const double *pathlength, const double *waist_xy,
                        const double *radii_xy, const double *wavelength, const double *theta,
                        const double *abcd, const double *r1x, const double *r1y, const double *thx,
                        const double *thy, double *poly, cudaStream_t st) {
  if (nb == 0) return TG_OK;
  coeffs_from_beam_kernel<<<nblocks(nb, 128), 128, 0, st>>>(nb, amp, pathlength, waist_xy, radii_xy, wavelength,
                                                           theta, abcd, r1x, r1y, thx, thy, poly);
  return tg_launch_check("coeffs_from_beam_kernel");
}

extern "C" int tg_input_coeffs_f64(int64_t nb, const double *amp, const double *phase_offset,
                                   const double *Q1_inv, const double *r1m, const double *theta1m,
                                   const double *k, double *poly, void *stream) {
  TG_REQUIRE(nb >= 0, "negative nb");
  if (nb == 0) return TG_OK;
  TG_REQUIRE(amp && phase_offset && Q1_inv && r1m && theta1m && k && poly, "null pointer");
  input_coeffs_kernel<<<nblocks(nb, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      nb, amp, phase_offset, Q1_inv, r1m, theta1m, k, poly);
  return tg_launch_check("input_coeffs_kernel");
}

extern "C" int tg_gaussian_qinv_f64(int64_t nb, const double *waist_xy, const double *radii_xy,
                                    const double *wavelength, const double *theta, double *Q_inv,
                                    void *stream) {
  TG_REQUIRE(nb >= 0, "negative nb");
  if (nb == 0) return TG_OK;
  TG_REQUIRE(waist_xy && radii_xy && wavelength && theta && Q_inv, "null pointer");
  qinv_kernel<<<nblocks(nb, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      nb, waist_xy, radii_xy, wavelength, theta, Q_inv);
  return tg_launch_check("qinv_kernel");
}
