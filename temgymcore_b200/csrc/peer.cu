// Multi-GPU plumbing of the row-sharded field sum over NVLink peer memory (one process per GPU):
// IPC-shareable image buffers, the fused compute + gather entry point, and a device-side barrier.
// The stores into the peers' images are issued by the kernels that produce the final values
// (GEMM epilogue in separable.cu, split-reduce in field.cu); see TgPeers in tg_common.cuh.
#include <stdlib.h>
#include <string.h>
#include "tg_common.cuh"

namespace {

struct FlagPtrs {
  unsigned long long *ptr[TG_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Thread p: publish `epoch` in slot `self` of peer p's flag array, then wait for slot p of the local
// array.  Runs after the compute kernels on the same stream, so their peer stores are ordered before
// the release store (fence.sys is cumulative over what happened-before the kernel launch).
// state (may be NULL): {epoch counter, status} in LOCAL memory.  With a state block the epoch is taken from
// the device-side counter (incremented here), so the same launch can be replayed from a CUDA graph, and a
// timeout is reported in state[1] (= 1 + the rank that did not arrive) instead of trapping: the context
// survives and the host can read the word (PeerImage.status()).  Without it: explicit epoch, trap on timeout.
__global__ void peer_barrier_kernel(const FlagPtrs f, int npeers, int self, unsigned long long epoch,
                                    long long timeout_cycles, unsigned long long *state) {
  __shared__ unsigned long long ep;
  if (threadIdx.x == 0) {
    if (state) {
      epoch = state[0] + 1;
      state[0] = epoch;
    }
    ep = epoch;
  }
  __syncthreads();
  epoch = ep;
  const int p = threadIdx.x;
  if (p >= npeers) return;
  __threadfence_system();
  st_release_sys(f.ptr[p] + self, epoch);
  const unsigned long long *mine = f.ptr[self] + p;
  const long long t0 = clock64();
  while (ld_acquire_sys(mine) < epoch) {
    if (clock64() - t0 > timeout_cycles) {
      printf("tg_peer_barrier: rank %d timed out waiting for rank %d (epoch %llu)\n", self, p, epoch);
      if (state) {
        atomicMax(state + 1, (unsigned long long)(p + 1));
        return;
      }
      __trap();
    }
    __nanosleep(200);
  }
}

// budget of the barrier's spin loop in SM cycles: TG_PEER_BARRIER_TIMEOUT_S seconds (default 10) at ~2 GHz
// (querying cudaDevAttrClockRate costs milliseconds per call: measured)
long long barrier_timeout_cycles(double timeout_s) {
  if (!(timeout_s > 0.0)) {
    static const double env_s = [] {
      const char *e = getenv("TG_PEER_BARRIER_TIMEOUT_S");
      const double v = e ? atof(e) : 10.0;
      return v > 0.0 ? v : 10.0;
    }();
    timeout_s = env_s;
  }
  const double c = timeout_s * 2.0e9;
  return c > 9.0e18 ? (long long)9.0e18 : (long long)c;
}

}  // namespace

extern "C" int tg_peer_alloc(uint64_t bytes, void **dptr, tg_ipc_handle *handle) {
  TG_REQUIRE(dptr && handle && bytes > 0, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(tg_ipc_handle), "IPC handle size");
  void *p = nullptr;
  TG_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    tg_set_error("tg_peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return TG_ECUDA;
  }
  memcpy(handle->bytes, &h, sizeof(h));
  *dptr = p;
  return TG_OK;
}

extern "C" int tg_peer_open(const tg_ipc_handle *handle, void **dptr) {
  TG_REQUIRE(handle && dptr, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof(h));
  TG_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TG_OK;
}

extern "C" int tg_peer_close(void *dptr) {
  if (dptr) TG_CUDA(cudaIpcCloseMemHandle(dptr));
  return TG_OK;
}

extern "C" int tg_peer_free(void *dptr) {
  if (dptr) TG_CUDA(cudaFree(dptr));
  return TG_OK;
}

static int make_peers(void *const images[], int npeers, int self, size_t off, TgPeers *pe, void **out) {
  TG_REQUIRE(images && npeers >= 1 && npeers <= TG_MAX_PEERS && self >= 0 && self < npeers, "bad peer set");
  pe->n = 0;
  for (int p = 0; p < npeers; ++p) {
    TG_REQUIRE(images[p], "null image pointer");
    if (p != self) pe->ptr[pe->n++] = static_cast<unsigned char *>(images[p]) + off;
  }
  *out = static_cast<unsigned char *>(images[self]) + off;
  pe->shard_off_bytes = off;
  static const bool cyclic = [] {          // TG_PEER_CYCLIC_ROWS=0: contiguous row blocks on the SFU path too
    const char *e = getenv("TG_PEER_CYCLIC_ROWS");
    return !(e && atoi(e) == 0);
  }();
  pe->cyc_world = cyclic ? npeers : 0;
  pe->cyc_rank = self;
  return TG_OK;
}

extern "C" int tg_field_sum_peers(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                                  int nrows, void *const images[], int npeers, int self, int out_is_c128,
                                  int cull_bits, int method, void *stream) {
  TG_REQUIRE(H > 0 && W > 0 && row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  TgPeers pe;
  void *out = nullptr;
  int rc = make_peers(images, npeers, self, (size_t)row0 * W * (out_is_c128 ? 16 : 8), &pe, &out);
  if (rc != TG_OK) return rc;
  return tg_field_sum_impl(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method,
                           static_cast<cudaStream_t>(stream), &pe);
}

extern "C" int tg_make_gaussian_image_peers(const tg_model *model_host, int64_t nb, const double *const rays[7],
                                            const double *amplitude, const double *waist_xy,
                                            const double *radii_xy, const double *wavelength, const double *theta,
                                            const double px2m[6], int H, int W, int row0, int nrows,
                                            void *const images[], int npeers, int self, int out_is_c128,
                                            int cull_bits, int method, void *stream) {
  TG_REQUIRE(H > 0 && W > 0 && row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  TgPeers pe;
  void *out = nullptr;
  int rc = make_peers(images, npeers, self, (size_t)row0 * W * (out_is_c128 ? 16 : 8), &pe, &out);
  if (rc != TG_OK) return rc;
  return tg_make_gaussian_image_impl(model_host, nb, rays, amplitude, waist_xy, radii_xy, wavelength, theta, px2m,
                                     H, W, row0, nrows, out, out_is_c128, cull_bits, method,
                                     static_cast<cudaStream_t>(stream), nullptr, &pe);
}

static int launch_barrier(void *const flags[], int npeers, int self, uint64_t epoch, void *state, double timeout_s,
                          void *stream) {
  TG_REQUIRE(flags && npeers >= 1 && npeers <= TG_MAX_PEERS && self >= 0 && self < npeers, "bad peer set");
  FlagPtrs f;
  for (int p = 0; p < npeers; ++p) {
    TG_REQUIRE(flags[p], "null flag pointer");
    f.ptr[p] = static_cast<unsigned long long *>(flags[p]);
  }
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      f, npeers, self, (unsigned long long)epoch, barrier_timeout_cycles(timeout_s),
      static_cast<unsigned long long *>(state));
  return tg_launch_check("peer_barrier_kernel");
}

extern "C" int tg_peer_barrier(void *const flags[], int npeers, int self, uint64_t epoch, void *stream) {
  return launch_barrier(flags, npeers, self, epoch, nullptr, 0.0, stream);
}

extern "C" int tg_peer_barrier_auto(void *const flags[], int npeers, int self, void *state, double timeout_s,
                                    void *stream) {
  TG_REQUIRE(state, "null state block");
  return launch_barrier(flags, npeers, self, 0, state, timeout_s, stream);
}
