// Multi-GPU plumbing of the row-sharded field sum over NVLink peer memory (one process per GPU):
// IPC-shareable image buffers, the fused compute + gather entry point, and a device-side barrier.
// The stores into the peers' images are issued by the kernels that produce the final values
// (GEMM epilogue in separable.cu, split-reduce in field.cu); see TgPeers in tg_common.cuh.
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
__global__ void peer_barrier_kernel(const FlagPtrs f, int npeers, int self, unsigned long long epoch,
                                    long long timeout_cycles) {
  const int p = threadIdx.x;
  if (p >= npeers) return;
  __threadfence_system();
  st_release_sys(f.ptr[p] + self, epoch);
  const unsigned long long *mine = f.ptr[self] + p;
  const long long t0 = clock64();
  while (ld_acquire_sys(mine) < epoch) {
    if (clock64() - t0 > timeout_cycles) {
      printf("tg_peer_barrier: rank %d timed out waiting for rank %d (epoch %llu)\n", self, p, epoch);
      __trap();
    }
    __nanosleep(200);
  }
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

extern "C" int tg_field_sum_peers(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                                  int nrows, void *const images[], int npeers, int self, int out_is_c128,
                                  int cull_bits, int method, void *stream) {
  TG_REQUIRE(images && npeers >= 1 && npeers <= TG_MAX_PEERS && self >= 0 && self < npeers, "bad peer set");
  TG_REQUIRE(H > 0 && W > 0 && row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  const size_t off = (size_t)row0 * W * (out_is_c128 ? 16 : 8);
  TgPeers pe;
  pe.n = 0;
  for (int p = 0; p < npeers; ++p) {
    TG_REQUIRE(images[p], "null image pointer");
    if (p != self) pe.ptr[pe.n++] = static_cast<unsigned char *>(images[p]) + off;
  }
  void *out = static_cast<unsigned char *>(images[self]) + off;
  return tg_field_sum_impl(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method,
                           static_cast<cudaStream_t>(stream), &pe);
}

extern "C" int tg_peer_barrier(void *const flags[], int npeers, int self, uint64_t epoch, void *stream) {
  TG_REQUIRE(flags && npeers >= 1 && npeers <= TG_MAX_PEERS && self >= 0 && self < npeers, "bad peer set");
  FlagPtrs f;
  for (int p = 0; p < npeers; ++p) {
    TG_REQUIRE(flags[p], "null flag pointer");
    f.ptr[p] = static_cast<unsigned long long *>(flags[p]);
  }
  // ~10 s of SM cycles at 2 GHz (querying cudaDevAttrClockRate costs milliseconds per call: measured)
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, npeers, self, (unsigned long long)epoch,
                                                                      20000000000LL);
  return tg_launch_check("peer_barrier_kernel");
}
