// Shared helpers for the temgym_b200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "temgym_b200.h"

void tg_set_error(const char *fmt, ...);
void tg_tune_mempool(int dev);
// Optional by-products of the prep kernel for the tensor-core path (one launch instead of three):
// sep_key  <- atomicMax of the bits of max_n(cross term / tolerance)   (separability verdict)
// peak_key <- atomicMax of the ordered-uint key of max_n peak log2|U_n(row)| over rows [row0, row0+nrows)
struct TgPrepExtra {
  unsigned long long *sep_key;
  unsigned long long *peak_key;
  int row0, nrows;
};
int tg_launch_prep(int64_t nb, const double *poly, const double px2m[6], int H, int W, double *table,
                   unsigned long long *gref_key, cudaStream_t st, const TgPrepExtra *extra = nullptr);
// k = 2 pi / wavelength, p0 = k * pathlength (reference gaussian.py:253-255); internal helper
int tg_wave_numbers(int64_t nb, const double *wavelength, const double *pathlength, double *k,
                    double *p0, cudaStream_t st);

// Extra destinations of a field sum: the same row block inside the images of peer GPUs (device pointers
// mapped over NVLink, already offset to the block).  The kernels that produce the final values store them
// to `out` and to every peer, so the row-block "all-gather" of the row-sharded multi-GPU field sum happens
// inside the compute kernels, tile by tile, instead of as a separate collective.
struct TgPeers {
  int n;
  void *ptr[TG_MAX_PEERS];
  // set by tg_field_sum_peers / tg_make_gaussian_image_peers (0 elsewhere): the byte offset of this rank's row block
  // inside the images (ptr[] and `out` are already advanced by it), and this rank's place among the cyc_world ranks
  // that share the image -- the SFU path may then take cyclic tile rows of the whole image instead of the block
  size_t shard_off_bytes = 0;
  int cyc_world = 0, cyc_rank = 0;
};
// Row-block emission of a field sum into HOST memory (the host-buffer entry points): the image is computed in
// blocks of `block_rows` detector rows on the compute stream, and each finished block is copied to the host on
// `copy` (event + cudaMemcpyAsync) while the next blocks are still being computed.
struct TgEmit {
  unsigned char *host_out;  // rows [row0, row0 + nrows) of the call: (nrows, W) complex128 / complex64
  int block_rows;           // multiple of 128 (GEMM tile) -- hence of 32 (SFU tile)
  cudaStream_t copy;
  cudaEvent_t *ev;          // >= ceil(nrows / block_rows) events (cudaEventDisableTiming)
  // optional: n_flags words of pinned, device-mapped host memory (and their device address).  With them a one-batch
  // tensor-path call runs ONE GEMM launch that raises flag i when block i is complete, and the (synchronous)
  // caller polls the words instead of waiting on events of per-block launches
  unsigned int *flags_host = nullptr, *flags_dev = nullptr;
  int n_flags = 0;
};
#define TG_SEP_VERDICT_ONLY 1 /* tg_separable_run: table + separability / cost verdict into *key_async, nothing else */
#define TG_SEP_TRUSTED 2      /* tg_separable_run: the caller has read the verdict; skip the host-side check */
#define TG_SEP_VERDICT_SPLIT 4 /* verdict-only: key_async[0] = separability, key_async[1] = 1 when the beamlets are sparse */
#define TG_BIN_PROBE 8         /* tg_separable_binned_run (eager calls): also form AUTO's verdicts (separable? sparse?)
                                 and go on only when both hold -- one host read-back for the verdicts AND the operand
                                 count.  Otherwise returns TG_NOT_BINNED with *probe_verdict = 1 (dense GEMM) or 0 (SFU) */
#define TG_NOT_BINNED 1       /* internal return code of the probe, not an error */

// stream-ordered scratch allocation, released (stream-ordered) when the scope ends -- also on error returns
struct TgAsyncBuf {
  void *p = nullptr;
  cudaStream_t st;
  explicit TgAsyncBuf(cudaStream_t s) : st(s) {}
  TgAsyncBuf(const TgAsyncBuf &) = delete;
  TgAsyncBuf &operator=(const TgAsyncBuf &) = delete;
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, st); }
  template <class T> T *as() const { return static_cast<T *>(p); }
  ~TgAsyncBuf() {
    if (p) cudaFreeAsync(p, st);
  }
};
// copy rows of a finished block to the host behind an event (see TgEmit)
static inline int tg_emit_block(const TgEmit *emit, int block, cudaStream_t compute, const void *dev_rows,
                                size_t byte_offset, size_t bytes) {
  cudaError_t e = cudaEventRecord(emit->ev[block], compute);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(emit->copy, emit->ev[block], 0);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(emit->host_out + byte_offset, dev_rows, bytes, cudaMemcpyDeviceToHost, emit->copy);
  if (e != cudaSuccess) {
    tg_set_error("row-block D2H: %s", cudaGetErrorString(e));
    return TG_ECUDA;
  }
  return TG_OK;
}
static inline bool tg_stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

// Q_inv + wave numbers + coefficients from the traced ABCD in one kernel (coeffs.cu)
int tg_coeffs_from_beam(int64_t nb, const double *amp, const double *pathlength, const double *waist_xy,
                        const double *radii_xy, const double *wavelength, const double *theta,
                        const double *abcd, const double *r1x, const double *r1y, const double *thx,
                        const double *thy, double *poly, cudaStream_t st);
int tg_field_grid_run(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                      int nrows, void *out, int out_is_c128, int cull_bits, long long *n_evals_out,
                      const unsigned long long *sep_guard, cudaStream_t stream, const TgPeers *peers = nullptr,
                      const TgEmit *emit = nullptr);
// key_async != NULL: no host sync; the verdict stays on the device in *key_async and the kernels of
// this path return at once when it says "not separable".
// cost_cull_bits > 0 (async mode only): also estimate the culled SFU work and leave the call to the SFU
// path when that is clearly cheaper than the dense GEMM.
// f16 != 0: fp16 x 3 operands (kind::f16, device-side pre-scaling), else tf32 x 3.
int tg_separable_run(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                     int nrows, void *out, int out_is_c128, unsigned long long *key_async,
                     cudaStream_t stream, int cost_cull_bits, int f16, const TgPeers *peers = nullptr,
                     const TgEmit *emit = nullptr, int flags = 0);
// Tile-binned (block-sparse K) tensor-core sum for separable beamlets that each reach a small part of the detector
// (cull_bits > 0).  TG_EUNSUPPORTED when the operands would not fit (use the SFU kernel), TG_ENOTSEPARABLE like
// tg_separable_run.  Under stream capture the operand capacity comes from the last eager call of this host thread with
// the same shape (plans warm up eagerly), and beamlets that need more poison the output with NaN.
int tg_separable_binned_run(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0, int nrows,
                            void *out, int out_is_c128, int cull_bits, cudaStream_t stream,
                            const TgPeers *peers = nullptr, const TgEmit *emit = nullptr, int flags = 0,
                            int *probe_verdict = nullptr);
int tg_field_sum_impl(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0, int nrows,
                      void *out, int out_is_c128, int cull_bits, int method, cudaStream_t st, const TgPeers *peers,
                      const TgEmit *emit = nullptr);
// the whole of make_gaussian_image for device-resident inputs (host_api.cu); emit / peers optional
int tg_make_gaussian_image_impl(const tg_model *model_host, int64_t nb, const double *const rays[7],
                                const double *amplitude, const double *waist_xy, const double *radii_xy,
                                const double *wavelength, const double *theta, const double px2m[6], int H,
                                int W, int row0, int nrows, void *out, int out_is_c128, int cull_bits,
                                int method, cudaStream_t s, const TgEmit *emit, const TgPeers *peers);
// separability key = bits of max_n(cross term / tolerance) as a double (0 when there is none)
__host__ __device__ inline bool tg_key_is_separable(unsigned long long key) {
  union { unsigned long long u; double d; } c;
  c.u = key;
  return !(c.d > 1.0);
}

#ifdef __CUDACC__
// max over c in [0, W-1] of E1 c + E3 c^2 (the column part of the envelope exponent, bits)
__device__ __forceinline__ double tg_col_env_max(double E1, double E3, double Wm1) {
  double best = fmax(0.0, Wm1 * (E1 + E3 * Wm1));
  if (E3 < 0.0) {
    const double cs = fmin(fmax(-E1 / (2.0 * E3), 0.0), Wm1);
    best = fmax(best, cs * (E1 + E3 * cs));
  }
  return isfinite(best) ? best : 0.0;
}
// monotone map double -> uint64 (for atomicMax / atomicMin on doubles) and back
__device__ __forceinline__ unsigned long long tg_enc_ordered(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double tg_dec_ordered(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}
// pre-scaling exponent of the tensor-core factors: G = ceil(brightest row-factor peak, bits), clamped;
// key == 0: no finite beamlet set the peak
__device__ __forceinline__ double tg_prescale_G(unsigned long long peak_key) {
  double G = 0.0;
  if (peak_key != 0ULL) G = ceil(tg_dec_ordered(peak_key));
  return fmin(fmax(G, -960.0), 960.0);
}
#endif

#define TG_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t err__ = (call);                                                         \
    if (err__ != cudaSuccess) {                                                         \
      tg_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      return TG_ECUDA;                                                                  \
    }                                                                                   \
  } while (0)

#define TG_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      tg_set_error("%s:%d: %s", __FILE__, __LINE__, msg);       \
      return TG_EINVAL;                                         \
    }                                                           \
  } while (0)

static inline int tg_launch_check(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    tg_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return TG_ECUDA;
  }
  return TG_OK;
}

// ---- small PTX wrappers (Blackwell: bulk async copies through the TMA unit) ----------
__device__ __forceinline__ uint32_t tg_smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void tg_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tg_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tg_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tg_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool tg_mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(tg_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tg_mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!tg_mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (1D, TMA unit), completion on an mbarrier. 16 B aligned, size % 16 == 0
__device__ __forceinline__ void tg_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(tg_smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(tg_smem_u32(bar))
      : "memory");
}
// shared -> global bulk copy (1D, TMA unit)
__device__ __forceinline__ void tg_bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(tg_smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tg_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tg_bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tg_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
