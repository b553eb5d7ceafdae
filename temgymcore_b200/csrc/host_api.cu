// Error plumbing and the HOST-buffer entry points of the C ABI (include/temgym_b200.h).
// These are what a non-CUDA host program (the reference's Python with numpy buffers, or a
// jax.ffi CPU-side shim) binds: pointers are host memory, the H2D copy, kernels and the D2H
// copy run inside the call, chunked over several streams so PCIe up/down and the kernels
// overlap.  No CPU compute fallback exists.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <stdlib.h>
#include "tg_common.cuh"

static thread_local char g_err[512] = "";

void tg_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *tg_last_error(void) { return g_err; }
extern "C" int tg_abi_version(void) { return TG_ABI_VERSION; }
extern "C" int tg_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    tg_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return TG_ECUDA;
  }
  return n;
}

void tg_tune_mempool(int dev);

namespace {

constexpr int kSlots = 3;
constexpr int64_t kChunkRays = 1 << 18;

}  // namespace

// Keep freed stream-ordered allocations cached in the device's default pool instead of
// returning them to the driver at every synchronisation (the default threshold is 0).
void tg_tune_mempool(int dev) {
  static bool done[64] = {false};
  if (dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = cudaSetDevice(dev) == cudaSuccess;
    if (ok) tg_tune_mempool(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Streams {
  cudaStream_t s[kSlots] = {nullptr, nullptr, nullptr};
  int n = 0;
  int init(int count) {
    for (int i = 0; i < count; ++i) {
      TG_CUDA(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
      n = i + 1;
    }
    return TG_OK;
  }
  ~Streams() {
    for (int i = 0; i < n; ++i) cudaStreamDestroy(s[i]);
  }
};

}  // namespace

extern "C" int tg_trace_f64_host(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                                 double *const out[7], double *jac, int jac_layout, int device) {
  TG_REQUIRE(model_host && in, "null model or input");
  TG_REQUIRE(n >= 0, "negative n");
  if (n == 0) return TG_OK;
  DeviceGuard guard(device);
  if (!guard.ok) {
    tg_set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
    return TG_ECUDA;
  }
  const int jw = jac_layout == TG_JAC_ABCD5 ? 25 : (jac_layout == TG_JAC_FULL7 ? 49 : 0);
  TG_REQUIRE(jw == 0 || jac, "jac requested but pointer is null");
  const int64_t chunk = std::min<int64_t>(n, kChunkRays);
  const int nslots = (int)std::min<int64_t>(kSlots, (n + chunk - 1) / chunk);
  Streams st;
  int rc = st.init(nslots);
  if (rc != TG_OK) return rc;

  int n_in = 0, n_out = 0;
  for (int f = 0; f < 7; ++f) {
    n_in += in->ptr[f] ? 1 : 0;
    n_out += (out && out[f]) ? 1 : 0;
  }
  const size_t per_slot = (size_t)chunk * 8 * (size_t)(n_in + n_out + jw);
  unsigned char *dbuf[kSlots] = {nullptr, nullptr, nullptr};
  for (int s = 0; s < nslots; ++s) {
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&dbuf[s]), per_slot ? per_slot : 8, st.s[s]);
    if (e != cudaSuccess) {
      tg_set_error("cudaMallocAsync(%zu): %s", per_slot, cudaGetErrorString(e));
      for (int k = 0; k < s; ++k) cudaFreeAsync(dbuf[k], st.s[k]);
      return TG_ECUDA;
    }
  }
  rc = TG_OK;
  int slot = 0;
  for (int64_t b = 0; b < n && rc == TG_OK; b += chunk, slot = (slot + 1) % nslots) {
    const int64_t cnt = std::min<int64_t>(chunk, n - b);
    cudaStream_t s = st.s[slot];
    double *base = reinterpret_cast<double *>(dbuf[slot]);
    tg_ray_in din = *in;
    double *dout[7];
    double *p = base;
    cudaError_t e = cudaSuccess;
    for (int f = 0; f < 7 && e == cudaSuccess; ++f) {
      if (in->ptr[f]) {
        e = cudaMemcpyAsync(p, in->ptr[f] + b, (size_t)cnt * 8, cudaMemcpyHostToDevice, s);
        din.ptr[f] = p;
        p += chunk;
      }
    }
    for (int f = 0; f < 7; ++f) {
      dout[f] = nullptr;
      if (out && out[f]) {
        dout[f] = p;
        p += chunk;
      }
    }
    double *djac = jw ? p : nullptr;
    if (e != cudaSuccess) {
      tg_set_error("H2D copy: %s", cudaGetErrorString(e));
      rc = TG_ECUDA;
      break;
    }
    rc = tg_trace_f64(model_host, cnt, &din, dout, djac, jac_layout, s);
    if (rc != TG_OK) break;
    for (int f = 0; f < 7 && e == cudaSuccess; ++f)
      if (dout[f]) e = cudaMemcpyAsync(out[f] + b, dout[f], (size_t)cnt * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && jw)
      e = cudaMemcpyAsync(jac + b * jw, djac, (size_t)cnt * jw * 8, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) {
      tg_set_error("D2H copy: %s", cudaGetErrorString(e));
      rc = TG_ECUDA;
    }
  }
  for (int s = 0; s < nslots; ++s) {
    cudaFreeAsync(dbuf[s], st.s[s]);
    cudaError_t e = cudaStreamSynchronize(st.s[s]);
    if (e != cudaSuccess && rc == TG_OK) {
      tg_set_error("stream sync: %s", cudaGetErrorString(e));
      rc = TG_ECUDA;
    }
  }
  return rc;
}

extern "C" int tg_metres_to_pixels_host(int64_t n, const double *x, const double *y,
                                        const double m2px[9], void *py, void *px, int as_float,
                                        int device) {
  TG_REQUIRE(n >= 0 && m2px, "bad arguments");
  if (n == 0) return TG_OK;
  TG_REQUIRE(x && y && py && px, "null pointer");
  DeviceGuard guard(device);
  if (!guard.ok) {
    tg_set_error("cudaSetDevice(%d) failed", device);
    return TG_ECUDA;
  }
  Streams st;
  int rc = st.init(1);
  if (rc != TG_OK) return rc;
  cudaStream_t s = st.s[0];
  const size_t osz = as_float ? 8 : 4;
  unsigned char *d = nullptr;
  TG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d), (size_t)n * (16 + 2 * osz), s));
  double *dx = reinterpret_cast<double *>(d), *dy = dx + n;
  unsigned char *dpy = d + (size_t)n * 16, *dpx = dpy + (size_t)n * osz;
  cudaError_t e = cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, (size_t)n * 8, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    rc = tg_metres_to_pixels(n, dx, dy, m2px, dpy, dpx, as_float, s);
    if (rc == TG_OK) {
      e = cudaMemcpyAsync(py, dpy, (size_t)n * osz, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(px, dpx, (size_t)n * osz, cudaMemcpyDeviceToHost, s);
    }
  }
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  if (rc == TG_OK && (e != cudaSuccess || e2 != cudaSuccess)) {
    tg_set_error("metres_to_pixels_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    rc = TG_ECUDA;
  }
  return rc;
}

// Device-resident make_gaussian_image (gaussian.py:225-273) as ONE call: ray kernel with ABCD, one
// kernel for Q_inv + wave numbers + coefficients, field sum (method dispatch), all enqueued on `stream`.
extern "C" int tg_make_gaussian_image_f64(const tg_model *model_host, int64_t nb,
                                          const double *const rays[7], const double *amplitude,
                                          const double *waist_xy, const double *radii_xy,
                                          const double *wavelength, const double *theta,
                                          const double px2m[6], int H, int W, int row0, int nrows,
                                          void *out, int out_is_c128, int cull_bits, int method,
                                          void *stream) {
  TG_REQUIRE(model_host && rays && px2m && out, "null pointer");
  TG_REQUIRE(nb >= 0 && H > 0 && W > 0, "bad sizes");
  TG_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  if (nrows == 0) return TG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (nb == 0) return tg_field_sum(0, nullptr, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method, s);
  TG_REQUIRE(amplitude && waist_xy && radii_xy && wavelength && theta, "null pointer");
  int dev = 0;
  TG_CUDA(cudaGetDevice(&dev));
  tg_tune_mempool(dev);
  double *scratch = nullptr;  // abcd 25n | poly 12n
  TG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)nb * 37 * 8, s));
  double *dabcd = scratch, *dpoly = dabcd + 25 * nb;
  tg_ray_in in;
  for (int f = 0; f < 7; ++f) {
    in.ptr[f] = rays[f];
    in.value[f] = 0.0;
    if (!rays[f]) {
      cudaFreeAsync(scratch, s);
      tg_set_error("tg_make_gaussian_image_f64: ray field %d is null", f);
      return TG_EINVAL;
    }
  }
  double *no_out[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rc = tg_trace_f64(model_host, nb, &in, no_out, dabcd, TG_JAC_ABCD5, s);
  if (rc == TG_OK)   // Q_inv, k, p0 and the coefficients in one kernel
    rc = tg_coeffs_from_beam(nb, amplitude, rays[5], waist_xy, radii_xy, wavelength, theta, dabcd, rays[0], rays[1],
                             rays[2], rays[3], dpoly, s);
  if (rc == TG_OK)
    rc = tg_field_sum(nb, dpoly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method, s);
  cudaFreeAsync(scratch, s);
  return rc;
}

extern "C" int tg_make_gaussian_image_host(const tg_model *model_host, int64_t nb,
                                           const double *const rays[7], const double *amplitude,
                                           const double *waist_xy, const double *radii_xy,
                                           const double *wavelength, const double *theta,
                                           const double px2m[6], int H, int W, int row0, int nrows,
                                           void *out, int out_is_c128, int cull_bits, int method,
                                           int device) {
  TG_REQUIRE(model_host && rays && amplitude && waist_xy && radii_xy && wavelength && theta && px2m && out,
             "null pointer");
  TG_REQUIRE(nb >= 0 && H > 0 && W > 0, "bad sizes");
  TG_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  if (nrows == 0) return TG_OK;
  DeviceGuard guard(device);
  if (!guard.ok) {
    tg_set_error("cudaSetDevice(%d) failed", device);
    return TG_ECUDA;
  }
  Streams st;
  int rc = st.init(1);
  if (rc != TG_OK) return rc;
  cudaStream_t s = st.s[0];
  const size_t npix = (size_t)nrows * W, elt = out_is_c128 ? 16 : 8;
  // device layout (doubles): rays 7n | amp n | waist 2n | radii 2n | wl n | theta n | field
  const size_t nd = (size_t)nb * 14;
  unsigned char *d = nullptr;
  TG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d), ((nd * 8 + 255) / 256) * 256 + npix * elt, s));
  double *p = reinterpret_cast<double *>(d);
  double *dr[7];
  for (int f = 0; f < 7; ++f) { dr[f] = p; p += nb; }
  double *damp = p; p += nb;
  double *dw = p; p += 2 * nb;
  double *drad = p; p += 2 * nb;
  double *dwl = p; p += nb;
  double *dth = p; p += nb;
  void *dout = d + ((nd * 8 + 255) / 256) * 256;
  cudaError_t e = cudaSuccess;
  auto up = [&](double *dst, const double *src, size_t cnt) {
    if (e == cudaSuccess && cnt) e = cudaMemcpyAsync(dst, src, cnt * 8, cudaMemcpyHostToDevice, s);
  };
  static const bool timing = getenv("TG_HOST_TIMING") != nullptr;   // debug: per-phase times on stderr
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (timing) {
    for (auto &x : ev) cudaEventCreate(&x);
    cudaEventRecord(ev[0], s);
  }
  for (int f = 0; f < 7; ++f) up(dr[f], rays[f], nb);
  up(damp, amplitude, nb);
  up(dw, waist_xy, 2 * nb);
  up(drad, radii_xy, 2 * nb);
  up(dwl, wavelength, nb);
  up(dth, theta, nb);
  if (e != cudaSuccess) {
    tg_set_error("H2D copy: %s", cudaGetErrorString(e));
    rc = TG_ECUDA;
  }
  if (timing) cudaEventRecord(ev[1], s);
  if (rc == TG_OK)
    rc = tg_make_gaussian_image_f64(model_host, nb, dr, damp, dw, drad, dwl, dth, px2m, H, W, row0, nrows, dout,
                                    out_is_c128, cull_bits, method, s);
  if (timing) cudaEventRecord(ev[2], s);
  if (rc == TG_OK) {
    e = cudaMemcpyAsync(out, dout, npix * elt, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) {
      tg_set_error("D2H copy: %s", cudaGetErrorString(e));
      rc = TG_ECUDA;
    }
  }
  if (timing) cudaEventRecord(ev[3], s);
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  if (rc == TG_OK && e2 != cudaSuccess) {
    tg_set_error("stream sync: %s", cudaGetErrorString(e2));
    rc = TG_ECUDA;
  }
  if (timing) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    cudaEventElapsedTime(&c, ev[2], ev[3]);
    fprintf(stderr, "tg_make_gaussian_image_host: H2D %.3f ms, kernels %.3f ms, D2H %.3f ms\n", a, b, c);
    for (auto &x : ev) cudaEventDestroy(x);
  }
  return rc;
}
