// Error plumbing and the HOST-buffer entry points of the C ABI (include/temgym_b200.h).
// These are what a non-CUDA host program (the reference's Python with numpy buffers, or a
// jax.ffi CPU-side shim) binds: pointers are host memory, the H2D copy, kernels and the D2H
// copy run inside the call, chunked over several streams so PCIe up/down and the kernels
// overlap.  No CPU compute fallback exists.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <atomic>
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
  static std::atomic<bool> done[64];     // zero-initialised; setting the attribute twice is harmless
  if (dev < 0 || dev >= 64 || done[dev].load(std::memory_order_acquire)) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev].store(true, std::memory_order_release);
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
int tg_make_gaussian_image_impl(const tg_model *model_host, int64_t nb, const double *const rays[7],
                                const double *amplitude, const double *waist_xy, const double *radii_xy,
                                const double *wavelength, const double *theta, const double px2m[6], int H,
                                int W, int row0, int nrows, void *out, int out_is_c128, int cull_bits,
                                int method, cudaStream_t s, const TgEmit *emit, const TgPeers *peers) {
  TG_REQUIRE(model_host && rays && px2m && out, "null pointer");
  TG_REQUIRE(nb >= 0 && H > 0 && W > 0, "bad sizes");
  TG_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  if (nrows == 0) return TG_OK;
  if (nb == 0)
    return tg_field_sum_impl(0, nullptr, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method, s, peers,
                             emit);
  TG_REQUIRE(amplitude && waist_xy && radii_xy && wavelength && theta, "null pointer");
  int dev = 0;
  TG_CUDA(cudaGetDevice(&dev));
  tg_tune_mempool(dev);
  tg_ray_in in;
  for (int f = 0; f < 7; ++f) {
    in.ptr[f] = rays[f];
    in.value[f] = 0.0;
    if (!rays[f]) {
      tg_set_error("tg_make_gaussian_image_f64: ray field %d is null", f);
      return TG_EINVAL;
    }
  }
  TgAsyncBuf scratch(s);  // abcd 25n | poly 12n
  TG_CUDA(scratch.alloc((size_t)nb * 37 * 8));
  double *dabcd = scratch.as<double>(), *dpoly = dabcd + 25 * nb;
  double *no_out[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rc = tg_trace_f64(model_host, nb, &in, no_out, dabcd, TG_JAC_ABCD5, s);
  if (rc == TG_OK)   // Q_inv, k, p0 and the coefficients in one kernel
    rc = tg_coeffs_from_beam(nb, amplitude, rays[5], waist_xy, radii_xy, wavelength, theta, dabcd, rays[0], rays[1],
                             rays[2], rays[3], dpoly, s);
  if (rc == TG_OK)
    rc = tg_field_sum_impl(nb, dpoly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method, s, peers, emit);
  return rc;
}

extern "C" int tg_make_gaussian_image_f64(const tg_model *model_host, int64_t nb,
                                          const double *const rays[7], const double *amplitude,
                                          const double *waist_xy, const double *radii_xy,
                                          const double *wavelength, const double *theta,
                                          const double px2m[6], int H, int W, int row0, int nrows,
                                          void *out, int out_is_c128, int cull_bits, int method,
                                          void *stream) {
  return tg_make_gaussian_image_impl(model_host, nb, rays, amplitude, waist_xy, radii_xy, wavelength, theta, px2m, H,
                                     W, row0, nrows, out, out_is_c128, cull_bits, method,
                                     static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

namespace {
// Streams and events of the host-buffer pipeline, created once per host thread and device (creating and
// destroying two streams per call cost more than the H2D copies they carried).
constexpr int kMaxEmitBlocks = 64;
struct HostPipe {
  cudaStream_t compute = nullptr, copy = nullptr;
  cudaEvent_t ev[kMaxEmitBlocks] = {};
  unsigned int *flags_host = nullptr, *flags_dev = nullptr;   // pinned + mapped: block-complete words of a streamed launch
  bool ok = false;
};
HostPipe *host_pipe(int dev) {
  static thread_local HostPipe pipes[16];
  if (dev < 0 || dev >= 16) return nullptr;
  HostPipe &p = pipes[dev];
  if (!p.ok) {
    if (cudaStreamCreateWithFlags(&p.compute, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto &e : p.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    void *fh = nullptr, *fd = nullptr;
    if (cudaHostAlloc(&fh, kMaxEmitBlocks * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(&fd, fh, 0) == cudaSuccess) {
      memset(fh, 0, kMaxEmitBlocks * sizeof(unsigned int));
      p.flags_host = static_cast<unsigned int *>(fh);
      p.flags_dev = static_cast<unsigned int *>(fd);
    } else {
      cudaGetLastError();              // no mapped memory: the pipeline launches one GEMM per block instead
    }
    p.ok = true;
  }
  return &p;
}
int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}
}  // namespace

// Host-buffer make_gaussian_image.  Pipeline: (1) H2D of the beamlet parameters -- ONE copy when the twelve
// arrays lie back to back in host memory (pack them that way: temgymcore_b200.gaussian.pack_beamlets_pinned),
// otherwise one per contiguous run; (2) ray kernel, coefficient kernel, verdict; (3) the field sum in blocks of
// detector rows, the D2H of every finished block on a second stream while the next block is being computed.
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
  HostPipe *pipe = host_pipe(device);
  if (!pipe) {
    tg_set_error("could not create the streams of the host pipeline: %s", cudaGetErrorString(cudaGetLastError()));
    return TG_ECUDA;
  }
  cudaStream_t s = pipe->compute;
  const size_t npix = (size_t)nrows * W, elt = out_is_c128 ? 16 : 8;
  // device layout (doubles): rays 7n | amp n | waist 2n | radii 2n | wl n | theta n | field
  const size_t nd = (size_t)nb * 14;
  int rc = TG_OK;
  {
    TgAsyncBuf dbuf(s);
    TG_CUDA(dbuf.alloc(((nd * 8 + 255) / 256) * 256 + npix * elt));
    unsigned char *d = dbuf.as<unsigned char>();
    double *p = reinterpret_cast<double *>(d);
    // the twelve segments in device order
    const double *src[12] = {rays[0], rays[1], rays[2], rays[3], rays[4], rays[5], rays[6],
                             amplitude, waist_xy, radii_xy, wavelength, theta};
    const size_t len[12] = {(size_t)nb, (size_t)nb, (size_t)nb, (size_t)nb, (size_t)nb, (size_t)nb, (size_t)nb,
                            (size_t)nb, 2 * (size_t)nb, 2 * (size_t)nb, (size_t)nb, (size_t)nb};
    double *dst[12];
    for (int i = 0; i < 12; ++i) {
      TG_REQUIRE(src[i] || nb == 0, "null ray field");
      dst[i] = p;
      p += len[i];
    }
    void *dout = d + ((nd * 8 + 255) / 256) * 256;
    static const bool timing = getenv("TG_HOST_TIMING") != nullptr;   // debug: per-phase times on stderr
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (timing) {
      for (auto &x : ev) cudaEventCreate(&x);
      cudaEventRecord(ev[0], s);
    }
    int ncopies = 0;
    for (int i = 0; i < 12 && nb > 0;) {   // merge runs that are contiguous on the host
      int j = i;
      size_t cnt = len[i];
      while (j + 1 < 12 && src[j + 1] == src[j] + len[j]) cnt += len[++j];
      TG_CUDA(cudaMemcpyAsync(dst[i], src[i], cnt * 8, cudaMemcpyHostToDevice, s));
      ++ncopies;
      i = j + 1;
    }
    if (timing) cudaEventRecord(ev[1], s);
    // rows per emitted block: a few blocks per image so that the D2H of block i hides behind block i+1
    static const int blk_env = env_int("TG_E2E_BLOCK_ROWS", 0);
    int block_rows = blk_env > 0 ? ((blk_env + 127) / 128) * 128 : 256;
    if (blk_env < 0) block_rows = nrows;                     // TG_E2E_BLOCK_ROWS=-1: one block (A/B runs)
    while ((nrows + block_rows - 1) / block_rows > kMaxEmitBlocks) block_rows *= 2;
    block_rows = ((block_rows + 127) / 128) * 128;
    TgEmit emit;
    emit.host_out = static_cast<unsigned char *>(out);
    emit.block_rows = block_rows;
    emit.copy = pipe->copy;
    emit.ev = pipe->ev;
    emit.flags_host = pipe->flags_host;
    emit.flags_dev = pipe->flags_dev;
    emit.n_flags = pipe->flags_host ? kMaxEmitBlocks : 0;
    const double *dr[7] = {dst[0], dst[1], dst[2], dst[3], dst[4], dst[5], dst[6]};
    rc = tg_make_gaussian_image_impl(model_host, nb, dr, dst[7], dst[8], dst[9], dst[10], dst[11], px2m, H, W, row0,
                                     nrows, dout, out_is_c128, cull_bits, method, s, &emit, nullptr);
    if (timing) cudaEventRecord(ev[2], s);
    cudaError_t e1 = cudaStreamSynchronize(s);
    cudaError_t e2 = cudaStreamSynchronize(pipe->copy);
    if (rc == TG_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
      tg_set_error("stream sync: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      rc = TG_ECUDA;
    }
    if (timing) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, ev[0], ev[1]);
      cudaEventElapsedTime(&b, ev[1], ev[2]);
      fprintf(stderr, "tg_make_gaussian_image_host: H2D %.3f ms (%d copies), kernels %.3f ms, blocks of %d rows\n",
              a, ncopies, b, block_rows);
      for (auto &x : ev) cudaEventDestroy(x);
    }
  }   // device buffers released (stream-ordered) after both streams have drained
  return rc;
}
