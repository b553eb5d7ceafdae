// Device-side input generation and beam-parameter recovery: the callers either side of the
// hot path that would otherwise stage 1e6-beamlet runs through host memory.
//   * concentric_rings (reference utils.py:117-175) -> ParallelBeam / PointSource.make_rays (source.py:58-188)
//   * decompose_Q_inv  (reference gaussian.py:35-89)
// -fmad=false like the other fp64 "definition" kernels: every + - * / is its own IEEE operation in the
// reference's order, so ring radii and the running angle sums reproduce numpy bit for bit (sin / cos of the
// angles are CUDA's fp64 libm, within 1-2 ulp of numpy's).
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "tg_common.cuh"

namespace {

struct RingPlan {
  std::vector<double> radii, div_angle;     // per ring
  std::vector<long long> ring_start;        // first point of ring k (size n_rings + 1)
  std::vector<long long> seg_start;         // first point of cumulative-sum segment s (size n_seg + 1)
  long long n = 0;
};

// The ring layout of concentric_rings, with numpy's arithmetic: floor / round-half-even / linspace.
RingPlan plan_rings(long long num_points_approx, double radius) {
  RingPlan p;
  const double pi = 3.141592653589793;
  long long n_rings = (long long)floor((-1.0 + sqrt(1.0 + 4.0 * (double)num_points_approx / pi)) / 2.0);
  if (n_rings < 1) n_rings = 1;
  std::vector<long long> circ(n_rings), per(n_rings);
  long long circ_sum = 0;
  for (long long k = 0; k < n_rings; ++k) {
    circ[k] = (long long)nearbyint(2.0 * pi * (double)(k + 1));
    circ_sum += circ[k];
  }
  const double per_unit = (double)num_points_approx / (double)circ_sum;
  p.ring_start.assign(1, 0);
  for (long long k = 0; k < n_rings; ++k) {
    per[k] = (long long)nearbyint((double)circ[k] * per_unit);
    if (per[k] < 0) per[k] = 0;
    p.ring_start.push_back(p.ring_start.back() + per[k]);
  }
  p.n = p.ring_start.back();
  // np.linspace(0, radius, n_rings + 1, endpoint=True)[1:]: arange * step + start, last element = stop
  const double step = (radius - 0.0) / (double)n_rings;
  p.radii.resize(n_rings);
  p.div_angle.resize(n_rings);
  for (long long k = 0; k < n_rings; ++k) {
    p.radii[k] = (k == n_rings - 1) ? radius : (double)(k + 1) * step + 0.0;
    p.div_angle[k] = 2.0 * pi / (double)per[k];
  }
  // multi_cumsum_inplace (utils.py:46-80) restarts its running sum when the element counter of the current
  // partition equals the partition length BEFORE being advanced: segment s covers per[s] + 1 elements, so the
  // restarts drift one element per ring behind the ring boundaries.  Reproduced as is.
  long long start = 0, k = 0;
  p.seg_start.assign(1, 0);
  while (start < p.n) {
    long long stop = start + per[k] + 1;
    if (stop > p.n) stop = p.n;
    p.seg_start.push_back(stop);
    start = stop;
    if (k + 1 < n_rings) ++k;
  }
  return p;
}

__device__ __forceinline__ int ring_of(const long long *__restrict__ ring_start, int n_rings, long long i) {
  int lo = 0, hi = n_rings;   // ring_start[lo] <= i < ring_start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ring_start[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// one thread per segment: sequential running sum of the per-ring angle steps (the order of the additions is the
// definition of the result)
__global__ void __launch_bounds__(128)
    ring_angle_kernel(int n_seg, const long long *__restrict__ seg_start, const long long *__restrict__ ring_start,
                      int n_rings, const double *__restrict__ div_angle, double *__restrict__ ang) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const long long a = seg_start[s], b = seg_start[s + 1];
  if (b <= a) return;
  double acc = 0.0;
  ang[a] = acc;
  int k = ring_of(ring_start, n_rings, a);
  for (long long i = a + 1; i < b; ++i) {
    while (k + 1 < n_rings && ring_start[k + 1] <= i) ++k;
    acc = div_angle[k] + acc;      // values[i] += values[i - 1]
    ang[i] = acc;
  }
}
__global__ void __launch_bounds__(256)
    ring_point_kernel(long long n, const long long *__restrict__ ring_start, int n_rings,
                      const double *__restrict__ radii, const double *__restrict__ ang, double *__restrict__ y,
                      double *__restrict__ x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = radii[ring_of(ring_start, n_rings, i)];
  double sn, cs;
  sincos(ang[i], &sn, &cs);
  y[i] = r * sn;
  x[i] = r * cs;
}

// ---- decompose_Q_inv (gaussian.py:35-89) ------------------------------------------------------------
struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, double b) { return {a.re * b, a.im * b}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }

// eigenvectors of the symmetric 2 x 2 [[a, b], [b, c]], eigenvalues ascending, columns (v0, v1) orthonormal
__device__ __forceinline__ void eigh2(double a, double b, double c, double v[2][2]) {
  if (b == 0.0) {
    if (a <= c) { v[0][0] = 1.0; v[1][0] = 0.0; v[0][1] = 0.0; v[1][1] = 1.0; }
    else        { v[0][0] = 0.0; v[1][0] = 1.0; v[0][1] = 1.0; v[1][1] = 0.0; }
    return;
  }
  // Jacobi rotation angle: tan(2 t) = 2 b / (a - c); take the stable form via tau
  const double tau = (c - a) / (2.0 * b);
  const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
  // rotation [[cs, sn], [-sn, cs]] diagonalises: eigenvalues l0 = a - t b (vector (cs, -sn)), l1 = c + t b ((sn, cs))
  const double l0 = a - t * b, l1 = c + t * b;
  if (l0 <= l1) { v[0][0] = cs; v[1][0] = -sn; v[0][1] = sn; v[1][1] = cs; }
  else          { v[0][0] = sn; v[1][0] = cs; v[0][1] = cs; v[1][1] = -sn; }
}
__device__ __forceinline__ void right_handed(double v[2][2]) {
  const double det = v[0][0] * v[1][1] - v[0][1] * v[1][0];
  if (det < 0.0) { v[0][1] = -v[0][1]; v[1][1] = -v[1][1]; }
}
__device__ __forceinline__ double waist_of(double im, double wavelength, double eps) {
  return fabs(im) > eps ? sqrt(fabs(wavelength / (3.141592653589793 * im))) : INFINITY;
}
__global__ void __launch_bounds__(128)
    decompose_qinv_kernel(long long n, const double *__restrict__ Q /* (n,2,2) complex interleaved */,
                          const double *__restrict__ wavelength, int wl_stride, double eps,
                          double *__restrict__ w1, double *__restrict__ w2, double *__restrict__ r1,
                          double *__restrict__ r2, double *__restrict__ theta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *q = Q + i * 8;
  const cplx q00{q[0], q[1]}, q01{q[2], q[3]}, q10{q[4], q[5]}, q11{q[6], q[7]};
  const double wl = wavelength[i * wl_stride];
  // Sm = (Im Q + Im Q^T) / 2
  double v[2][2];
  eigh2(0.5 * (q00.im + q00.im), 0.5 * (q01.im + q10.im), 0.5 * (q11.im + q11.im), v);
  right_handed(v);
  // Qd = V^T Q V, diagonal entries
  auto diag = [&](int j) {
    const cplx t0 = cadd(cmul(q00, v[0][j]), cmul(q01, v[1][j]));   // (Q v_j)_0
    const cplx t1 = cadd(cmul(q10, v[0][j]), cmul(q11, v[1][j]));   // (Q v_j)_1
    return cadd(cmul(t0, v[0][j]), cmul(t1, v[1][j]));
  };
  cplx d0 = diag(0), d1 = diag(1);
  if (waist_of(d0.im, wl, eps) < waist_of(d1.im, wl, eps)) {   // larger waist first
    const cplx t = d0; d0 = d1; d1 = t;
    const double a0 = v[0][0], a1 = v[1][0];
    v[0][0] = v[0][1]; v[1][0] = v[1][1]; v[0][1] = a0; v[1][1] = a1;
    right_handed(v);
  }
  w1[i] = waist_of(d0.im, wl, eps);
  w2[i] = waist_of(d1.im, wl, eps);
  r1[i] = fabs(d0.re) > eps ? 1.0 / d0.re : INFINITY;
  r2[i] = fabs(d1.re) > eps ? 1.0 / d1.re : INFINITY;
  theta[i] = atan2(v[1][0], v[0][0]);
}

}  // namespace

extern "C" int64_t tg_concentric_rings_count(int64_t num_points_approx, double radius) {
  if (num_points_approx < 0) return TG_EINVAL;
  return (int64_t)plan_rings(num_points_approx, radius).n;
}

extern "C" int tg_concentric_rings_f64(int64_t num_points_approx, double radius, int64_t capacity, double *y,
                                       double *x, void *stream) {
  TG_REQUIRE(num_points_approx >= 0, "negative point count");
  const RingPlan p = plan_rings(num_points_approx, radius);
  if (p.n == 0) return TG_OK;
  TG_REQUIRE(y && x, "null pointer");
  TG_REQUIRE(capacity >= p.n, "output arrays are smaller than tg_concentric_rings_count()");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_rings = (int)p.radii.size(), n_seg = (int)p.seg_start.size() - 1;
  // device tables: radii | div_angle | ring_start | seg_start | angles
  const size_t b_r = (size_t)n_rings * 8, b_rs = (size_t)(n_rings + 1) * 8, b_ss = (size_t)(n_seg + 1) * 8;
  TgAsyncBuf buf(st);
  TG_CUDA(buf.alloc(2 * b_r + b_rs + b_ss + (size_t)p.n * 8));
  unsigned char *d = buf.as<unsigned char>();
  double *d_radii = reinterpret_cast<double *>(d), *d_div = reinterpret_cast<double *>(d + b_r);
  long long *d_rs = reinterpret_cast<long long *>(d + 2 * b_r), *d_ss = reinterpret_cast<long long *>(d + 2 * b_r + b_rs);
  double *d_ang = reinterpret_cast<double *>(d + 2 * b_r + b_rs + b_ss);
  // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may go out of scope
  TG_CUDA(cudaMemcpyAsync(d_radii, p.radii.data(), b_r, cudaMemcpyHostToDevice, st));
  TG_CUDA(cudaMemcpyAsync(d_div, p.div_angle.data(), b_r, cudaMemcpyHostToDevice, st));
  TG_CUDA(cudaMemcpyAsync(d_rs, p.ring_start.data(), b_rs, cudaMemcpyHostToDevice, st));
  TG_CUDA(cudaMemcpyAsync(d_ss, p.seg_start.data(), b_ss, cudaMemcpyHostToDevice, st));
  ring_angle_kernel<<<(unsigned)((n_seg + 127) / 128), 128, 0, st>>>(n_seg, d_ss, d_rs, n_rings, d_div, d_ang);
  int rc = tg_launch_check("ring_angle_kernel");
  if (rc != TG_OK) return rc;
  ring_point_kernel<<<(unsigned)((p.n + 255) / 256), 256, 0, st>>>(p.n, d_rs, n_rings, d_radii, d_ang, y, x);
  return tg_launch_check("ring_point_kernel");
}

extern "C" int tg_decompose_qinv_f64(int64_t n, const double *Q_inv, const double *wavelength, int wavelength_is_scalar,
                                     double eps, double *waist1, double *waist2, double *radius1, double *radius2,
                                     double *theta, void *stream) {
  TG_REQUIRE(n >= 0, "negative n");
  if (n == 0) return TG_OK;
  TG_REQUIRE(Q_inv && wavelength && waist1 && waist2 && radius1 && radius2 && theta, "null pointer");
  decompose_qinv_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n, Q_inv, wavelength, wavelength_is_scalar ? 0 : 1, eps, waist1, waist2, radius1, radius2, theta);
  return tg_launch_check("decompose_qinv_kernel");
}
