// K4: tensor-core field sum for SEPARABLE beamlets (tcgen05 / TMEM / TMA, sm_100a).
//
// When the pixel-space polynomial of every beamlet has no col*row term,
//     exp(i P_n(row, col)) = U_n(row) * V_n(col)        (complex),
// and the field sum of the reference (gaussian.py:319-369) is the complex GEMM
//     F[row, col] = sum_n U[n,row] V[n,col],           K = number of beamlets.
// BASELINE configs C2 and C3 are of this kind (isotropic beamlets through Lens / free space /
// Biprism onto an axis-aligned detector).
//
// Real formulation (one real GEMM, output already interleaved re/im):
//     A'[row][2n]   = Re U, A'[row][2n+1]   = Im U                    (M  x K', K' = 2 nb)
//     B'[2col][2n]  = Re V, B'[2col][2n+1]  = -Im V   -> column 2col   = Re F
//     B'[2col+1][2n]= Im V, B'[2col+1][2n+1]= Re V    -> column 2col+1 = Im F
//     D = A' * B'^T                                                   (M x 2W)
// Precision: operands are fp32 values split as x = hi + lo, and D accumulates
// hi*hi + hi*lo + lo*hi on the tensor cores with fp32 accumulation in TMEM -> ~2^-22 relative per
// product.  Two operand formats (template parameter F16):
//   * fp16 x 3 (default): hi = fp16(x), lo = fp16(x - hi); kind::f16 runs at twice the TF32 rate and
//     the operands are half the bytes.  fp16 has 5 exponent bits, so the factors are pre-scaled on
//     the device: U by 2^(14 - G) with G = ceil(log2 of the brightest row-factor peak of the call)
//     and V (|V| <= 1 by construction) by 2^14; the epilogue multiplies by 2^(G - 28) in fp64
//     (exact).  A factor keeps 22 bits down to 2^-17 of the brightest peak and degrades
//     gracefully to an absolute floor of 2^-38 of it (fp16 subnormals).
//   * tf32 x 3 (TG_METHOD_TENSOR_TF32): hi = tf32(x) (round to nearest), lo = x - hi; fp32 range.
// The TMEM accumulator is drained every 128 k-elements into fp32 registers (round-to-nearest
// adds), which bounds the tensor-core accumulation length.
//
// Kernel structure (one CTA per 128 x 128 output tile, 384 threads):
//   warp 0   : TMA producer  - 4 tiles (A_hi, A_lo, B_hi, B_lo; 128 rows x 128 B = 64 fp16 or
//              32 tf32, SWIZZLE_128B) per k-block into a 3-stage shared-memory ring, mbarrier expect_tx
//   warp 1   : MMA issuer    - one thread issues 8 tcgen05.mma per k-block (4 K-steps of 32 B x {N = 256:
//              A_hi [B_hi; B_lo]^T, N = 128: A_lo B_hi^T}) into one of two 256-column TMEM accumulators;
//              tcgen05.commit frees the smem stage / publishes the accumulator
//   warp 2   : TMEM allocator (512 columns)
//   warps 4-11: epilogue     - tcgen05.ld the finished accumulator (32x32b.x32), add into
//              registers, release the accumulator; finally write doubles to the output
#include <atomic>
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "tg_common.cuh"

namespace {

constexpr int BM = 128, BN = 128;
// One k-block = one swizzle row (BK_BYTES) per tile row.  The ring holds 192 KB either way; what matters is
// the share of it that can be in flight while one stage is being multiplied, (STAGES - 1) / STAGES: the
// operand fetch (L2 hit ~1500 clk, HBM more) has to fit into the time the other stages take to be consumed.
// 3 x 128-byte stages leave a window of 2 x 768 = 1536 clk, 6 x 64-byte stages 5 x 384 = 1920 clk -- but
// measured on B200 the 64-byte variant (SWIZZLE_64B boxes) is SLOWER: C2 0.256 vs 0.226 ms, C3 10.3 vs 8.7 ms
// (64-byte rows halve the TMA request size and the MMA's operand fetch efficiency), so 128 stays.
#ifndef TG_GEMM_BK_BYTES
#define TG_GEMM_BK_BYTES 128
#endif
constexpr int BK_BYTES = TG_GEMM_BK_BYTES;      // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
static_assert(BK_BYTES == 128 || BK_BYTES == 64, "k-block = one 128-byte or 64-byte swizzle row");
constexpr int STAGES = 384 / BK_BYTES;          // 3 or 6 stages of 4 tiles: 192 KiB
constexpr int TILE_BYTES = BM * BK_BYTES;       // 16 or 8 KiB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;     // A_hi, A_lo, B_hi, B_lo
// k-elements accumulated in TMEM before the epilogue drains them into registers.  256 (round 2): the epilogue's work per
// MMA halves -- decisive for the 3-product form, whose 16 epilogue warps pace the kernel at 128 (C2: 0.188 -> 0.178 ms;
// 4-multiplication form 0.209 -> 0.205 ms; row shards equal or better) -- at a GEMM error of 4.0e-7 instead of 3.2e-7
// relative to fp64 (the tensor core's accumulator truncates; the field's error stays at the ~1e-6 of its factors).
#ifndef TG_CHUNK_K
#define TG_CHUNK_K 256
#endif
constexpr int CHUNK_K = TG_CHUNK_K;             // k-elements per TMEM accumulation chunk (multiple of 64)
template <bool F16> struct GemmCfg {
  static constexpr int ELEM = F16 ? 2 : 4;               // operand bytes
  static constexpr int BK = BK_BYTES / ELEM;             // k-elements per k-block (64 fp16 / 32 tf32)
  static constexpr int CHUNK_KB = CHUNK_K / BK;          // k-blocks per accumulation chunk
};
constexpr int GEMM_THREADS = 384;                // real formulation: 4 service warps + 8 epilogue warps
constexpr int GEMM_THREADS_GAUSS = 640;          // complex 3-product formulation: 4 + 16 epilogue warps
constexpr int KCH = CHUNK_K;                     // beamlets per K'' group of the 3-product layout: one chunk per block
constexpr int ACC_COLS = 2 * BN;                // one accumulator = [A_hi B_hi | A_hi B_lo + A_lo B_hi], fp32
constexpr int TMEM_COLS = 2 * ACC_COLS;         // double-buffered: all 512 TMEM columns

// ---------------------------------------------------------------- PTX wrappers (tcgen05, TMA)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(tg_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(tg_smem_u32(bar)), "r"(c0),
      "r"(c1)
      : "memory");
}
// L2 prefetch of a tensor tile (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   tg_smem_u32(bar))
               : "memory");
}
template <bool F16>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (F16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//  [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4
//  (8 swizzle rows between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B) / 4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_smem_desc(const void *smem_tile) {
  const uint32_t addr = tg_smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * BK_BYTES) >> 4) << 32;            // SBO: 8 rows of one swizzle row each
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(BK_BYTES == 128 ? 2 : 4) << 61;        // UMMA::LayoutType SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a/b format at bits 7/10
// (F16 = 0, TF32 = 2), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
template <bool F16, int N> struct Idesc {
  static constexpr uint32_t value = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) |
                                    ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

struct GemmSmemCtl {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// ---- work decomposition: data-parallel tiles + a stream-K part that fills the machine ----------------
// The GEMM is a persistent kernel of G CTAs (one per SM).  T output tiles of 128 x 128, each nch accumulation
// chunks (CHUNK_K k-elements) deep.  Tiles [R, T) are "data-parallel": CTA c owns tiles R + c, R + c + G, ...
// over the whole K.  Tiles [0, R) (R < G: the partial last wave, or ALL tiles when the problem has fewer tiles
// than SMs, e.g. one rank's row shard) are split along K so that every CTA gets the same number of chunks:
//   head  : CTA r < R takes chunks [0, q) of tile r -- the R heads walk K in lockstep, so the operand tiles
//           they share are fetched from HBM once and served from L2 (a plain stream-K split loses exactly
//           that: measured in round 1);
//   tails : the remaining Tl = nch - q chunks of the R tiles are laid end to end and cut into runs of qh
//           chunks for the G - R helper CTAs.
// Every (tile, CTA) piece is a UNIT.  A unit that does not cover the whole K writes its fp32 partial tile to a
// scratch slot; the last unit of a tile to arrive (atomic counter per tile and epilogue warp) adds the slots
// in slot order and writes the output -- deterministic, no atomics on the data path.
struct SkSched {
  int tiles_n, T, nkb, nch, G, R, q, Tl, qh, maxparts;
  int prefetch;   // k-blocks the producer prefetches ahead into L2 (0 = none)
  int uniform;    // > 0: plain split-K -- CTA c owns chunks [split q, split q + q) of tile c % T, split = c / T
  int S;          // head / tail arrangement: full pieces of q chunks per stream-K tile (1 = heads only)
  // plain split-K in ROUNDS (the host pipeline's row-block streaming): the tiles are walked Tr at a time -- a block
  // of tile rows -- each round split along K over the whole machine, so that the image is finished block of rows by
  // block of rows inside ONE launch; rounds = 1, Tr = T is the plain split-K above
  int rounds, Tr;
  // RAGGED schedule of the tile-binned (block-sparse K) sum: tile t owns the GLOBAL accumulation chunks [P[t], P[t+1])
  // of one concatenated k axis, P and the chunks per CTA live in device memory (bins[]: built on the device from the
  // beamlets' bounding boxes, never read by the host); CTA c takes chunks [c q, (c + 1) q)
  const int *bins;
  // TILED operands of a static schedule (the 3-product path of tg_separable_run): k-block kb of A / B is stored as
  // tiled_a / tiled_b rows x 64 k-elements of contiguous memory (0 = row-major operands, rows a pitch apart)
  int tiled_a, tiled_b;
};
enum { BIN_CTOT = 0, BIN_Q = 1, BIN_OVERFLOW = 2, BIN_NEED = 3, BIN_HDR = 4 };   // bins[]: header, then P[0..T]
struct SkUnit {
  int tile, ch0, ch1, slot, nparts;
  int first, p0, rq;   // ragged: first CTA of the tile, P[tile], chunks per CTA
};
__host__ __device__ inline int sk_nparts(const SkSched &s, int tile) {
  if (s.uniform) return s.maxparts;
  if (s.Tl <= 0) return s.S;
  const int first = (int)(((long long)tile * s.Tl) / s.qh);
  const int last = (int)((((long long)tile + 1) * s.Tl - 1) / s.qh);
  return s.S + 1 + last - first;
}
struct SkIter {
  int phase;  // 0: head of a stream-K tile pending, 1: helper walking its run of tail chunks, 2: data-parallel tiles
  int f0, f1, t;
  template <bool RG = false>
  __host__ __device__ __forceinline__ void init(const SkSched &s, int cta) {
    phase = 2;
    f0 = f1 = 0;
    t = s.R + cta;
    if (RG) {                // ragged: f0 = position, f1 = end (global chunks), t = tile holding f0
      phase = 5;
      const int ctot = s.bins[BIN_CTOT], q = s.bins[BIN_Q];
      const long long a = (long long)cta * q;
      f0 = (int)(a < ctot ? a : ctot);
      f1 = (int)(a + q < ctot ? a + q : ctot);
      const int *P = s.bins + BIN_HDR;
      int lo = 0, hi = s.T;                       // P[0] = 0 <= f0 < P[T] = ctot (when the CTA has work)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (P[mid] <= f0) lo = mid;
        else hi = mid;
      }
      t = lo;
      return;
    }
    if (s.uniform) {
      phase = 3;
    } else if (s.R > 0) {
      if (cta < s.S * s.R) {
        phase = 0;
      } else {
        phase = 1;
        const long long tot = (long long)s.R * s.Tl;
        long long a = (long long)(cta - s.S * s.R) * s.qh, b = a + s.qh;
        if (a > tot) a = tot;
        if (b > tot) b = tot;
        f0 = (int)a;
        f1 = (int)b;
      }
    }
  }
  template <bool RG = false>
  __host__ __device__ __forceinline__ bool next(const SkSched &s, int cta, SkUnit &u) {
    if (RG) {                // ragged: the rest of this CTA's chunk range that lies inside the next non-empty tile
      if (f0 >= f1) return false;
      const int *P = s.bins + BIN_HDR;
      while (P[t + 1] <= f0) ++t;
      const int p0 = P[t], p1 = P[t + 1], q = s.bins[BIN_Q];
      u.tile = t;
      u.ch0 = f0;                                  // GLOBAL chunk indices: the k coordinate of the operands
      u.ch1 = f1 < p1 ? f1 : p1;
      u.first = p0 / q;
      u.nparts = (p1 - 1) / q - u.first + 1;
      u.slot = cta - u.first;
      u.p0 = p0;
      u.rq = q;
      f0 = u.ch1;
      return true;
    }
    if (phase == 3) {        // plain split-K: one unit per CTA and round
      const int split = cta / s.Tr, col = cta - split * s.Tr;
      while (f0 < s.rounds) {
        const int tile = f0 * s.Tr + col;
        ++f0;
        if (tile >= s.T) break;
        u.tile = tile;
        u.ch0 = split * s.q;
        u.ch1 = u.ch0 + s.q < s.nch ? u.ch0 + s.q : s.nch;
        u.slot = split;
        u.nparts = s.maxparts;
        if (u.ch0 < u.ch1) return true;
      }
      phase = 4;
      return false;
    }
    if (phase == 4) return false;
    if (phase == 0) {        // full piece `cta / R` of stream-K tile `cta % R`
      phase = 2;
      const int piece = cta / s.R;
      u.tile = cta - piece * s.R;
      u.ch0 = piece * s.q;
      u.ch1 = u.ch0 + s.q;
      u.slot = piece;
      u.nparts = sk_nparts(s, u.tile);
      return true;
    }
    if (phase == 1) {
      if (f0 < f1) {
        const int tile = f0 / s.Tl, c0 = f0 - tile * s.Tl;
        int c1 = c0 + (f1 - f0);
        if (c1 > s.Tl) c1 = s.Tl;
        u.tile = tile;
        u.ch0 = s.S * s.q + c0;
        u.ch1 = s.S * s.q + c1;
        u.slot = s.S + (cta - s.S * s.R) - (int)(((long long)tile * s.Tl) / s.qh);
        u.nparts = sk_nparts(s, tile);
        f0 += c1 - c0;
        return true;
      }
      phase = 2;
    }
    if (t < s.T) {
      u.tile = t;
      u.ch0 = 0;
      u.ch1 = s.nch;
      u.slot = 0;
      u.nparts = 1;
      t += s.G;
      return true;
    }
    return false;
  }
};

// D[M x Np] (+)= A[M x K] * B[Np x K]^T with A = A_hi + A_lo, B = B_hi + B_lo (3 products).
// out: doubles, row pitch ldo; rows >= M / columns >= Np are not written.  peak_key (device, may be NULL):
// the accumulators are multiplied in fp64 by 2^(G - 2 headroom), G = tg_prescale_G(*peak_key) -- the exact
// power of two that undoes the pre-scaling of the factors.
// scratch / counters: stream-K partial tiles (sched.R * sched.maxparts slots of 128 x 128 floats) and one
// arrival counter per (stream-K tile, epilogue warp), zero on entry and zero again on exit.
//
// GAUSS = true: COMPLEX product with three real multiplications per term.  The operands hold, per group of 128
// beamlets, three blocks of 128 k-elements each (row factors A: Ur + Ui | Ur | Ui, column factors B: Vr | Vi - Vr |
// Vr + Vi), so accumulation chunk ch of a tile is the real product k_j, j = ch mod 3, of one beamlet group:
//     k1 = (Ur + Ui) Vr,  k2 = Ur (Vi - Vr),  k3 = Ui (Vr + Vi);   Re(U V) = k1 - k3,  Im(U V) = k1 + k2.
// The tensor cores run the same real GEMM over K'' = 3 nb (instead of 4 nb real multiply-adds per complex
// term: 25 % less tensor work, same operand bytes); the epilogue adds each drained chunk into the (re, im)
// registers with the signs above.  A tile is 128 rows x 128 COMPLEX columns (Np = complex columns, the output row
// holds 2 Np doubles), 16 epilogue warps of 32 complex columns each.
// Experiment builds (TG_BUILD_DEFS=-DTG_GEMM_TRACE): %globaltimer stamps of every CTA's milestones, read back with
// tg_debug_gemm_trace -- where a launch spends the ~20 us it costs beyond its tensor work (tools/exp_gemm_trace.py)
#ifdef TG_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[148 * 8];
__device__ __forceinline__ void gemm_trace(int cta, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  if (cta < 148) g_gemm_trace[cta * 8 + slot] = t;
}
#define GEMM_TRACE(cta, slot) gemm_trace(cta, slot)
#else
#define GEMM_TRACE(cta, slot) ((void)0)
#endif

template <bool F16, bool GAUSS, bool RAGGED = false>
__global__ void __launch_bounds__(GAUSS ? GEMM_THREADS_GAUSS : GEMM_THREADS, 1)
    gemm_x3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   int M, int Np, int K, double *__restrict__ out, long long ldo, int accumulate_out,
                   const unsigned long long *__restrict__ peak_key, double headroom,
                   const unsigned long long *__restrict__ sep_guard, const __grid_constant__ TgPeers peers,
                   const __grid_constant__ SkSched sched, float *__restrict__ scratch,
                   unsigned int *__restrict__ counters, unsigned int *__restrict__ round_cnt,
                   unsigned int *round_flag, unsigned int round_epoch) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;  // the SFU path owns this call
  // Programmatic dependent launch: a following GEMM of the host pipeline (the next block of rows, launched with
  // programmatic stream serialisation) may start on SMs this grid leaves idle or frees -- the launches share
  // operands that were complete before the first of them started, and write disjoint rows, scratch and counters.
  // Without such a dependent the instruction does nothing.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int BK = GemmCfg<F16>::BK, CHUNK_KB = GemmCfg<F16>::CHUNK_KB;
  constexpr uint32_t kIdesc = Idesc<F16, BN>::value, kIdesc2 = Idesc<F16, 2 * BN>::value;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem) + 1023) &
                                                           ~uintptr_t(1023));
  GemmSmemCtl *ctl = reinterpret_cast<GemmSmemCtl *>(tiles + STAGES * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int nkb = sched.nkb;
  if (threadIdx.x == 0) GEMM_TRACE(cta, 0);           // kernel entry

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tg_mbar_init(&ctl->full[s], 1);
      tg_mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tg_mbar_init(&ctl->tmem_full[a], 1);
      tg_mbar_init(&ctl->tmem_empty[a], GAUSS ? 16 : 8);  // one arrive per epilogue warp
    }
    tg_fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tg_smem_u32(&ctl->tmem_base)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (threadIdx.x == 0) GEMM_TRACE(cta, 1);           // prologue done (barriers, TMEM)

  if (warp == 0) {
    // ===== TMA producer: the shared-memory ring runs on across unit boundaries =====
    if (lane == 0) {
      SkIter itr;
      itr.template init<RAGGED>(sched, cta);
      SkUnit u;
      uint32_t it = 0;
      while (itr.template next<RAGGED>(sched, cta, u)) {
        // (ragged schedule: every tile's operands are 128-row blocks of their own at the tile's global chunks)
        const int m0 = RAGGED ? 0 : (u.tile / sched.tiles_n) * BM, n0 = RAGGED ? 0 : (u.tile % sched.tiles_n) * BN;
        const int kb0 = u.ch0 * CHUNK_KB, kb1 = min(u.ch1 * CHUNK_KB, nkb);
        // Experiment knob (TG_GEMM_PREFETCH, default 0): stream-K pieces take ~1.8x longer per k-block than whole
        // tiles whose CTAs share operand tiles in lockstep; asking the TMA unit to prefetch the tiles `pf` k-blocks
        // ahead into L2 (cp.async.bulk.prefetch.tensor) was measured and does NOT help (it competes with the loads).
        const int pf = sched.prefetch;
        // (ragged schedule: a CTA walks ONE contiguous k-range through all its tiles and every operand byte comes from
        // HBM exactly once -- there the prefetch runs ahead across unit boundaries, up to the end of the CTA's range)
        const int pf_end = RAGGED ? itr.f1 * CHUNK_KB : kb1;
        if (pf > 0 && (!RAGGED || it == 0))
          for (int kb = kb0; kb < min(kb0 + pf, pf_end); ++kb) {
            tma_prefetch_2d(&tmA_hi, RAGGED ? 0 : kb * BK, RAGGED ? kb * BM : m0);
            tma_prefetch_2d(&tmA_lo, RAGGED ? 0 : kb * BK, RAGGED ? kb * BM : m0);
            tma_prefetch_2d(&tmB_hi, RAGGED ? 0 : kb * BK, RAGGED ? kb * BN : n0);
            tma_prefetch_2d(&tmB_lo, RAGGED ? 0 : kb * BK, RAGGED ? kb * BN : n0);
          }
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          const uint32_t ph = (it / STAGES) & 1u;
          if (pf > 0 && kb + pf < pf_end) {
            tma_prefetch_2d(&tmA_hi, RAGGED ? 0 : (kb + pf) * BK, RAGGED ? (kb + pf) * BM : m0);
            tma_prefetch_2d(&tmA_lo, RAGGED ? 0 : (kb + pf) * BK, RAGGED ? (kb + pf) * BM : m0);
            tma_prefetch_2d(&tmB_hi, RAGGED ? 0 : (kb + pf) * BK, RAGGED ? (kb + pf) * BN : n0);
            tma_prefetch_2d(&tmB_lo, RAGGED ? 0 : (kb + pf) * BK, RAGGED ? (kb + pf) * BN : n0);
          }
          tg_mbar_wait(&ctl->empty[s], ph ^ 1u);
          unsigned char *st = tiles + s * STAGE_BYTES;
          tg_mbar_expect_tx(&ctl->full[s], STAGE_BYTES);
          // (ragged schedule: operands are stored k-block by k-block, each 128-row x 128-byte tile contiguous -- a
          // tensor of 128-byte rows, tile kb at rows [128 kb, 128 kb + 128) -- so a load is ONE 16 KiB run of DRAM
          // instead of 128 lines a whole operand row apart)
          const bool tl = !RAGGED && sched.tiled_a > 0;          // tiled operands of a static schedule, same idea
          const int ck = (RAGGED || tl) ? 0 : kb * BK;
          const int cra = RAGGED ? kb * BM : tl ? kb * sched.tiled_a + m0 : m0;
          const int crb = RAGGED ? kb * BN : tl ? kb * sched.tiled_b + n0 : n0;
          tma_load_2d(st + 0 * TILE_BYTES, &tmA_hi, &ctl->full[s], ck, cra);
          tma_load_2d(st + 1 * TILE_BYTES, &tmA_lo, &ctl->full[s], ck, cra);
          tma_load_2d(st + 2 * TILE_BYTES, &tmB_hi, &ctl->full[s], ck, crb);
          tma_load_2d(st + 3 * TILE_BYTES, &tmB_lo, &ctl->full[s], ck, crb);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      SkIter itr;
      itr.template init<RAGGED>(sched, cta);
      SkUnit u;
      uint32_t it = 0, chn = 0;
      while (itr.template next<RAGGED>(sched, cta, u)) {
        const int kb0 = u.ch0 * CHUNK_KB, kb1 = min(u.ch1 * CHUNK_KB, nkb);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          const uint32_t ph = (it / STAGES) & 1u;
          const bool chunk_start = ((kb - kb0) % CHUNK_KB) == 0;
          const int acc = (int)(chn & 1u);
          if (chunk_start) {
            tg_mbar_wait(&ctl->tmem_empty[acc], ((chn >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
            tc_fence_after();
          }
          tg_mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          if (it == 0) GEMM_TRACE(cta, 2);             // first operand stage has landed
          unsigned char *st = tiles + s * STAGE_BYTES;
          const uint64_t dAh = make_smem_desc(st + 0 * TILE_BYTES), dAl = make_smem_desc(st + 1 * TILE_BYTES);
          const uint64_t dBh = make_smem_desc(st + 2 * TILE_BYTES);
          const uint32_t d = tmem_base + (uint32_t)(acc * ACC_COLS);
          // Two MMAs per K-step instead of three: the B_hi and B_lo tiles are adjacent in the stage, so one
          // N = 256 MMA forms A_hi [B_hi; B_lo]^T into columns [0,128) | [128,256) reading A_hi once, and
          // A_lo B_hi^T is added to the second half.  Same tensor work, 17 % less shared-memory operand
          // traffic (20 KB instead of 24 KB per K-step; the N = 128 form runs at the 128 B/clk smem limit),
          // and the small cross terms accumulate apart from the large hi*hi term.
#pragma unroll
          for (int k4 = 0; k4 < BK_BYTES / 32; ++k4) {
            const uint64_t ko = (uint64_t)(k4 * 32 >> 4);  // one MMA = 32 bytes (8 tf32 / 16 fp16) of the swizzled row
            tc_mma<F16>(d, dAh + ko, dBh + ko, kIdesc2, (chunk_start && k4 == 0) ? 0u : 1u);
            tc_mma<F16>(d + (uint32_t)BN, dAl + ko, dBh + ko, kIdesc, 1u);
          }
          tc_commit(&ctl->empty[s]);  // smem stage reusable once these MMAs have read it
          if (((kb - kb0) % CHUNK_KB) == CHUNK_KB - 1 || kb == kb1 - 1) {
            tc_commit(&ctl->tmem_full[acc]);
            ++chn;
          }
        }
        GEMM_TRACE(cta, 3);                            // all MMAs of the (last) unit issued
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue.  REAL: 8 warps; warp%4 selects the TMEM lane quarter, (warp-4)/4 the column half (64 real
    // columns).  GAUSS: 16 warps; (warp-4)/4 selects 32 of the 128 complex columns; accum = re[32] | im[32].
    constexpr int NE = GAUSS ? 16 : 8;
    constexpr int PART_FLOATS = NE * 2048;               // one stream-K partial tile: [warp][64][lane]
    const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
    const double sc = peak_key ? scalbn(1.0, (int)(tg_prescale_G(*peak_key) - 2.0 * headroom)) : 1.0;
    float accum[64];
    SkIter itr;
    itr.template init<RAGGED>(sched, cta);
    SkUnit u;
    uint32_t chn = 0;
    while (itr.template next<RAGGED>(sched, cta, u)) {
      const int m0 = (u.tile / sched.tiles_n) * BM, n0 = (u.tile % sched.tiles_n) * BN;
      const int kb0 = u.ch0 * CHUNK_KB, kb1 = min(u.ch1 * CHUNK_KB, nkb);
      const int nchunks = (kb1 - kb0 + CHUNK_KB - 1) / CHUNK_KB;
#pragma unroll
      for (int i = 0; i < 64; ++i) accum[i] = 0.f;
      int j3 = u.ch0 % 3;                                 // GAUSS: which of k1, k2, k3 the next chunk is
      for (int ch = 0; ch < nchunks; ++ch, ++chn) {
        const int acc = (int)(chn & 1u);
        tg_mbar_wait(&ctl->tmem_full[acc], (chn >> 1) & 1u);
        tc_fence_after();
        if constexpr (!GAUSS) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + h * 64);
          float v[32];
#pragma unroll
          for (int part = 0; part < 2; ++part) {   // hi*hi columns, then the cross-term columns
            tc_ld32(taddr + part * BN, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) accum[i] += v[i];
            tc_ld32(taddr + part * BN + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) accum[32 + i] += v[i];
          }
        } else {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + h * 32);
          // 16 columns at a time, hi*hi then the cross terms: 64 accumulators + 16 loaded values stay inside the
          // 102-register budget of a 640-thread CTA (32-column loads of both parts spilled 1 KB per thread)
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            float v[16];
            const int c16 = (pc & 1) * 16;
            tc_ld16(taddr + (pc >> 1) * BN + c16, v);
            if (j3 == 0) {                         // k1: + re, + im
#pragma unroll
              for (int i = 0; i < 16; ++i) { accum[c16 + i] += v[i]; accum[32 + c16 + i] += v[i]; }
            } else if (j3 == 1) {                  // k2: + im
#pragma unroll
              for (int i = 0; i < 16; ++i) accum[32 + c16 + i] += v[i];
            } else {                               // k3: - re
#pragma unroll
              for (int i = 0; i < 16; ++i) accum[c16 + i] -= v[i];
            }
          }
          j3 = j3 == 2 ? 0 : j3 + 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->tmem_empty[acc]);
      }
      if (ew == 0 && lane == 0) GEMM_TRACE(cta, 4);     // last chunk drained from TMEM
      if (u.nparts > 1) {
        // stream-K piece: park the fp32 partial ([slot][epilogue warp][column][lane]: coalesced), count the
        // arrival; the last piece of this tile sums the slots in slot order (fp32, like the chunk sums)
        // static schedules: slot s of tile t at [t maxparts + s]; ragged: two slots per CTA -- a CTA's first unit
        // (which may start inside a tile) parks in slot 2c, a later partial unit (it starts at its tile's first chunk
        // and is the CTA's last) in slot 2c + 1
        constexpr bool rg = RAGGED;
        float *slot0 = scratch + (rg ? (size_t)0 : (size_t)u.tile * sched.maxparts) * (size_t)PART_FLOATS + (size_t)ew * 2048 + lane;
        float *mine = slot0 + (rg ? (size_t)(2 * cta + (u.ch0 == cta * u.rq ? 0 : 1)) : (size_t)u.slot) * (size_t)PART_FLOATS;
#pragma unroll
        for (int i = 0; i < 64; ++i) __stcg(mine + i * 32, accum[i]);
        __threadfence();
        __syncwarp();
        unsigned int prev = 0;
        if (lane == 0) {
          __threadfence();
          prev = atomicAdd(counters + u.tile * NE + ew, 1u);
        }
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (ew == 0 && lane == 0) GEMM_TRACE(cta, 5);   // partial parked, arrival counted
        if (prev != (unsigned)(u.nparts - 1)) continue;   // not the last piece
        __threadfence();
        if (lane == 0) counters[u.tile * NE + ew] = 0u;   // leave the counters clean for the next launch
#pragma unroll
        for (int i = 0; i < 64; ++i) accum[i] = 0.f;
        for (int s = 0; s < u.nparts; ++s) {
          size_t si = (size_t)s;
          if (rg) {
            const int c = u.first + s;
            si = (size_t)(2 * c + ((long long)c * u.rq >= (long long)u.p0 ? 0 : 1));
          }
          const float *src = slot0 + si * (size_t)PART_FLOATS;
#pragma unroll
          for (int i = 0; i < 64; ++i) accum[i] += __ldcg(src + i * 32);
        }
      }
      const int row = m0 + q * 32 + lane;
      if (row < M) {
        if constexpr (!GAUSS) {
          double *o = out + (long long)row * ldo + n0 + h * 64;
          const int ncol = min(64, Np - (n0 + h * 64));
          if (ncol == 64 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              double2 w = make_double2((double)accum[i] * sc, (double)accum[i + 1] * sc);
              if (accumulate_out) {
                const double2 p = *reinterpret_cast<double2 *>(o + i);
                w.x += p.x;
                w.y += p.y;
              }
              *reinterpret_cast<double2 *>(o + i) = w;
              // row-sharded multi-GPU sum: the same values go straight into the peers' images (NVLink P2P stores)
              for (int p = 0; p < peers.n; ++p)
                *reinterpret_cast<double2 *>(static_cast<double *>(peers.ptr[p]) + (o - out) + i) = w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i < ncol) {
                const double wv = (accumulate_out ? o[i] : 0.0) + (double)accum[i] * sc;
                o[i] = wv;
                for (int p = 0; p < peers.n; ++p) (static_cast<double *>(peers.ptr[p]) + (o - out))[i] = wv;
              }
          }
        } else {
          // complex columns n0 + 32 h + i: (re, im) pairs, 16-byte aligned when the output base and pitch are
          const int c0 = n0 + h * 32;
          double *o = out + (long long)row * ldo + 2 * (long long)c0;
          const int ncol = min(32, Np - c0);
          const bool al = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncol) {
              double2 w = make_double2((double)accum[i] * sc, (double)accum[32 + i] * sc);
              if (accumulate_out) {
                w.x += o[2 * i];
                w.y += o[2 * i + 1];
              }
              if (al) {
                *reinterpret_cast<double2 *>(o + 2 * i) = w;
                for (int p = 0; p < peers.n; ++p)
                  *reinterpret_cast<double2 *>(static_cast<double *>(peers.ptr[p]) + (o - out) + 2 * i) = w;
              } else {
                o[2 * i] = w.x;
                o[2 * i + 1] = w.y;
                for (int p = 0; p < peers.n; ++p) {
                  double *po = static_cast<double *>(peers.ptr[p]) + (o - out) + 2 * i;
                  po[0] = w.x;
                  po[1] = w.y;
                }
              }
            }
        }
      }
      if (ew == 0 && lane == 0) GEMM_TRACE(cta, 6);     // output rows stored (final writer)
      if (round_flag) {
        // row-block streaming: this warp's part of the tile is in memory; the last warp of the round's last tile
        // publishes the round (system scope: a copy-engine read behind cuStreamWaitValue32 follows)
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          const int rnd = u.tile / sched.Tr;
          const int left = sched.T - rnd * sched.Tr;
          const unsigned need = (unsigned)((left < sched.Tr ? left : sched.Tr) * NE);
          if (atomicAdd(round_cnt + rnd, 1u) == need - 1u) {
            round_cnt[rnd] = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int *>(round_flag + rnd) = round_epoch;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
  if (threadIdx.x == 0) GEMM_TRACE(cta, 7);           // CTA done
}

// ---------------------------------------------------------------- CTA-pair GEMM (tcgen05 cta_group::2)
// The one-CTA kernel above is bound by what ONE SM can take in: a k-block of four 128-row operand tiles is
// 64 KB per 768 tensor-pipe cycles (83 B/clk against ~64 B/clk of L2 -> SM bandwidth; ncu: tensor pipe 73 %
// active, no other unit saturated).  Here two CTAs on the two SMs of a TPC (a cluster of 2) compute a
// 256 x 128 tile together: each CTA loads its own 128 rows of A_hi / A_lo but only HALF of the B_hi / B_lo tiles
// (64 rows each), and the leader CTA issues tcgen05.mma.cta_group::2 with M = 256, which reads A from both
// CTAs' shared memory and the two B halves from both -- 48 KB per k-block and SM instead of 64 KB, so four
// stages fit where three did.  Each CTA's TMEM receives the accumulators of its own 128 rows.
//   barriers : full[s]       lives in the LEADER; both CTAs' TMA loads complete_tx on it (peer bit masked off)
//              empty[s]      one per CTA; the leader's tcgen05.commit multicasts the arrival to both
//              tmem_full[a]  one per CTA, multicast commit; each CTA's epilogue drains its own TMEM
//              tmem_empty[a] in the LEADER, 16 arrivals: the 8 epilogue warps of both CTAs (remote arrive)
// Whole tiles only (persistent over pair tiles); shapes that would leave the machine idle go to the one-CTA
// stream-K kernel.
constexpr int P_STAGES = 4;
constexpr int P_TILE_A = BM * BK_BYTES;                      // 16 KiB: 128 rows
constexpr int P_TILE_B = (BN / 2) * BK_BYTES;                // 8 KiB: this CTA's 64 rows of the B tile
constexpr int P_STAGE_BYTES = 2 * P_TILE_A + 2 * P_TILE_B;   // 48 KiB
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;               // shared::cluster address of the even CTA of a pair

struct PairSmemCtl {
  uint64_t full[P_STAGES];
  uint64_t empty[P_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (executed by both CTAs)
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(tg_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(tg_smem_u32(bar) & kPeerBitMask),
      "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the barrier at this offset in both CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          tg_smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tg_smem_u32(bar) & kPeerBitMask) : "memory");
}
template <bool F16>
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if constexpr (F16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// instruction descriptor with M = 256 (pair) and N
template <bool F16, int N> struct IdescPair {
  static constexpr uint32_t value = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) |
                                    ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

template <bool F16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_x3_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                        const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                        int M, int Np, int K, double *__restrict__ out, long long ldo, int accumulate_out,
                        const unsigned long long *__restrict__ peak_key, double headroom,
                        const unsigned long long *__restrict__ sep_guard, const __grid_constant__ TgPeers peers,
                        int tiles_n, int n_pair_tiles) {
  // (both CTAs of a cluster read the same key: they leave together)
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  constexpr int BK = GemmCfg<F16>::BK, CHUNK_KB = GemmCfg<F16>::CHUNK_KB;
  constexpr uint32_t kIdesc = IdescPair<F16, BN>::value;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem) + 1023) &
                                                           ~uintptr_t(1023));
  PairSmemCtl *ctl = reinterpret_cast<PairSmemCtl *>(tiles + P_STAGES * P_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nkb = (K + BK - 1) / BK;
  const int nchunks_tile = (nkb + CHUNK_KB - 1) / CHUNK_KB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      tg_mbar_init(&ctl->full[s], 1);
      tg_mbar_init(&ctl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tg_mbar_init(&ctl->tmem_full[a], 1);
      tg_mbar_init(&ctl->tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs (used in the leader only)
    }
    tg_fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tg_smem_u32(&ctl->tmem_base)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows, own half of the B rows
    if (lane == 0) {
      uint32_t it = 0;
      for (int pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
        const int m0 = ((pt / tiles_n) * 2 + (int)rank) * BM;
        const int n0 = (pt % tiles_n) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = (int)(it % P_STAGES);
          const uint32_t ph = (it / P_STAGES) & 1u;
          tg_mbar_wait(&ctl->empty[s], ph ^ 1u);
          unsigned char *st = tiles + s * P_STAGE_BYTES;
          if (rank == 0) tg_mbar_expect_tx(&ctl->full[s], 2 * P_STAGE_BYTES);   // both CTAs' bytes land here
          tma_load_2d_pair(st, &tmA_hi, &ctl->full[s], kb * BK, m0);
          tma_load_2d_pair(st + P_TILE_A, &tmA_lo, &ctl->full[s], kb * BK, m0);
          tma_load_2d_pair(st + 2 * P_TILE_A, &tmB_hi, &ctl->full[s], kb * BK, n0);
          tma_load_2d_pair(st + 2 * P_TILE_A + P_TILE_B, &tmB_lo, &ctl->full[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA only
    if (lane == 0 && rank == 0) {
      uint32_t it = 0, chn = 0;
      for (int pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = (int)(it % P_STAGES);
          const uint32_t ph = (it / P_STAGES) & 1u;
          const bool chunk_start = (kb % CHUNK_KB) == 0;
          const int acc = (int)(chn & 1u);
          if (chunk_start) {
            tg_mbar_wait(&ctl->tmem_empty[acc], ((chn >> 1) & 1u) ^ 1u);   // both CTAs have drained this accumulator
            tc_fence_after();
          }
          tg_mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          unsigned char *st = tiles + s * P_STAGE_BYTES;
          const uint64_t dAh = make_smem_desc(st), dAl = make_smem_desc(st + P_TILE_A);
          const uint64_t dBh = make_smem_desc(st + 2 * P_TILE_A), dBl = make_smem_desc(st + 2 * P_TILE_A + P_TILE_B);
          const uint32_t d = tmem_base + (uint32_t)(acc * ACC_COLS);
#pragma unroll
          for (int k4 = 0; k4 < BK_BYTES / 32; ++k4) {
            const uint64_t ko = (uint64_t)(k4 * 32 >> 4);
            const uint32_t first = (chunk_start && k4 == 0) ? 0u : 1u;
            tc_mma_pair<F16>(d, dAh + ko, dBh + ko, kIdesc, first);                    // hi * hi
            tc_mma_pair<F16>(d + (uint32_t)BN, dAh + ko, dBl + ko, kIdesc, first);     // hi * lo
            tc_mma_pair<F16>(d + (uint32_t)BN, dAl + ko, dBh + ko, kIdesc, 1u);        // lo * hi
          }
          tc_commit_pair(&ctl->empty[s]);          // both CTAs' stage s may be refilled once these MMAs have read it
          if ((kb % CHUNK_KB) == CHUNK_KB - 1 || kb == nkb - 1) {
            tc_commit_pair(&ctl->tmem_full[acc]);
            ++chn;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): drain the own 128 rows
    const int q = warp & 3, h = (warp - 4) >> 2;
    const double sc = peak_key ? scalbn(1.0, (int)(tg_prescale_G(*peak_key) - 2.0 * headroom)) : 1.0;
    float accum[64];
    uint32_t chn = 0;
    for (int pt = cluster_id; pt < n_pair_tiles; pt += n_clusters) {
      const int m0 = ((pt / tiles_n) * 2 + (int)rank) * BM;
      const int n0 = (pt % tiles_n) * BN;
#pragma unroll
      for (int i = 0; i < 64; ++i) accum[i] = 0.f;
      for (int ch = 0; ch < nchunks_tile; ++ch, ++chn) {
        const int acc = (int)(chn & 1u);
        tg_mbar_wait(&ctl->tmem_full[acc], (chn >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + h * 64);
        float v[32];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          tc_ld32(taddr + part * BN, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) accum[i] += v[i];
          tc_ld32(taddr + part * BN + 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) accum[32 + i] += v[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&ctl->tmem_empty[acc]);
      }
      const int row = m0 + q * 32 + lane;
      if (row < M) {
        double *o = out + (long long)row * ldo + n0 + h * 64;
        const int ncol = min(64, Np - (n0 + h * 64));
        if (ncol == 64 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            double2 w = make_double2((double)accum[i] * sc, (double)accum[i + 1] * sc);
            if (accumulate_out) {
              const double2 p = *reinterpret_cast<double2 *>(o + i);
              w.x += p.x;
              w.y += p.y;
            }
            *reinterpret_cast<double2 *>(o + i) = w;
            for (int p = 0; p < peers.n; ++p)
              *reinterpret_cast<double2 *>(static_cast<double *>(peers.ptr[p]) + (o - out) + i) = w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < ncol) {
              const double wv = (accumulate_out ? o[i] : 0.0) + (double)accum[i] * sc;
              o[i] = wv;
              for (int p = 0; p < peers.n; ++p) (static_cast<double *>(peers.ptr[p]) + (o - out))[i] = wv;
            }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // neither CTA leaves (or frees TMEM) while the pair still reads its shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- factor builders
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// table: pixel-space {T0..T5, E0..E5} per beamlet (prep_kernel of field.cu).
//
// One thread per beamlet walks a strip of FS consecutive rows (or columns): the table entry and
// the fp64 strip setup are paid once per strip; along the strip the phase advances in 32-bit
// fixed-point turns by exact integer differences and the envelope is the pivot-form fp32
// quadratic of the SFU kernel (field.cu).  Stores are coalesced: for a fixed row, consecutive
// threads (beamlets) write consecutive k' pairs.
// AUTO hands a separable problem to the culled SFU kernel when its estimated executed evaluations are
// below this fraction of nb*H*W.  Measured on B200 (tools/exp_auto.py, threshold sweeps through the
// environment variable TG_SFU_WINS_BELOW) on the C3 geometry (~11 px envelopes) with the fp16 x 3 GEMM
// (2.1e-14 s per nominal evaluation) and the gather-mode culled SFU kernel: 1e5 beamlets on 2048^2 (estimate
// between 2 and 3 %): SFU 6.1 ms vs GEMM 8.7 ms; 2e4 beamlets: 1.34 vs 1.79 ms; 4e3 beamlets on 1024^2
// (estimate 3-4.5 %): about equal.  Break-even estimate ~3.6 %.
constexpr double kSfuWinsBelow = 0.03;
constexpr int FS = 32;
constexpr long long kBatch = 16384;  // beamlets per GEMM pass
constexpr int kMaxRounds = 64;       // row blocks of one streamed launch (host pipeline)
constexpr double kMagicF = 1572864.0;  // 1.5 * 2^20

struct Strip1D {
  uint32_t t0, d0, dd;
  float jp, ep, e1p, e2;
};
// quadratic phase (turns) p0 + p1 s + p2 s^2 and envelope (bits) q0 + q1 s + q2 s^2 on s = s0 + j
__device__ __forceinline__ Strip1D strip_setup(double p0, double p1, double p2, double q0, double q1,
                                               double q2, double s0) {
  Strip1D r;
  const double ph = p0 + s0 * (p1 + p2 * s0);
  const double dl = p1 + p2 * (2.0 * s0 + 1.0);
  const double d2 = 2.0 * p2;
  r.t0 = (uint32_t)__double2loint((ph - rint(ph)) + kMagicF) + 0x100u;
  r.d0 = (uint32_t)__double2loint((dl - rint(dl)) + kMagicF);
  r.dd = (uint32_t)__double2loint((d2 - rint(d2)) + kMagicF);
  // pivot = strip pixel nearest the envelope vertex
  double js = 0.0;
  if (q2 < -1e-12) js = -0.5 * q1 / q2 - s0;
  const float jpf = fminf(fmaxf(rintf((float)js), 0.f), (float)(FS - 1));
  const double sp = s0 + (double)jpf;
  r.jp = jpf;
  r.ep = (float)(q0 + sp * (q1 + q2 * sp));
  r.e1p = (float)(q1 + 2.0 * q2 * sp);
  r.e2 = (float)q2;
  return r;
}
__device__ __forceinline__ void strip_eval(const Strip1D &r, int j, float &re, float &im) {
  const uint32_t tj = r.t0 + (uint32_t)j * r.d0 + (uint32_t)(j * (j - 1) / 2) * r.dd;
  const float ft = __uint_as_float((tj >> 9) | 0x3f800000u);
  const float ang = fmaf(ft, 6.28318530717958648f, -9.42477796076937972f);  // 2 pi frac - pi
  const float dj = (float)j - r.jp;
  const float ej = fmaf(dj, fmaf(dj, r.e2, r.e1p), r.ep);
  float amp;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(amp) : "f"(ej));
  re = -amp * __cosf(ang);  // exp(i 2 pi frac) = -(cos ang + i sin ang)
  im = -amp * __sinf(ang);
}

// operand stores: one (re, im) pair of hi and of lo per call
template <bool F16> struct Operand;
template <> struct Operand<false> {
  typedef float T;
  static __device__ __forceinline__ void split(float x, float &hi, float &lo) {
    hi = tf32_rna(x);
    lo = x - hi;
  }
  static __device__ __forceinline__ void store2(void *base, long long o, float a, float b) {
    *reinterpret_cast<float2 *>(static_cast<float *>(base) + o) = make_float2(a, b);
  }
};
template <> struct Operand<true> {
  typedef __half T;
  static __device__ __forceinline__ void split(float x, float &hi, float &lo) {
    hi = __half2float(__float2half_rn(x));
    lo = __half2float(__float2half_rn(x - hi));
  }
  static __device__ __forceinline__ void store2(void *base, long long o, float a, float b) {
    *reinterpret_cast<__half2 *>(static_cast<__half *>(base) + o) = __floats2half2_rn(a, b);  // exact: already fp16 values
  }
};
// factors are scaled so that their peaks are <= 2^headroom: 14 for fp16 (max 65504); the tf32 factors get
// the same treatment with headroom 0, which keeps weak beamlets out of the fp32 flush-to-zero range
template <bool F16> struct Headroom { static constexpr double value = F16 ? 14.0 : 0.0; };

// Pre-scaling: G = ceil(max over beamlets of the peak of log2|U_n(row)| over the call's rows) comes from
// the prep kernel's peak key (TgPrepExtra); the row factors are scaled by 2^(headroom - G), the column
// factors by 2^headroom, the GEMM epilogue by 2^(G - 2 headroom).
// A[(row - row0)][2n..2n+1] = U_n(row) for rows [row0, row0+M), beamlets [b0, b0+nbatch)
template <bool F16>
__global__ void __launch_bounds__(128)
    factor_rows_kernel(const double *__restrict__ table, long long b0, int nbatch, int row0, int M, int W,
                       long long ldk, void *__restrict__ Ahi, void *__restrict__ Alo,
                       const unsigned long long *__restrict__ peak_key,
                       const unsigned long long *__restrict__ sep_guard) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  // bounded 1-D grid walking the (beamlet block, strip) pairs: a launch that the device-side verdict cancels costs
  // microseconds instead of ~3 ns for each of ~10^4 CTAs (measured: 245 us of dead launches per C3 image)
  const int nbx = (nbatch + (int)blockDim.x - 1) / (int)blockDim.x, nby = (M + FS - 1) / FS;
  for (int vb = blockIdx.x; vb < nbx * nby; vb += gridDim.x) {
    const int n = (vb % nbx) * blockDim.x + threadIdx.x;
    const int m0 = (vb / nbx) * FS;
    if (n >= nbatch) continue;
    const double *t = table + (b0 + n) * 12;
    double mu = tg_col_env_max(t[6 + 1], t[6 + 3], (double)(W - 1));
    mu += Headroom<F16>::value - tg_prescale_G(*peak_key);
    const Strip1D st = strip_setup(t[0], t[2], t[5], t[6 + 0] + mu, t[6 + 2], t[6 + 5], (double)(row0 + m0));
#pragma unroll 4
    for (int j = 0; j < FS; ++j) {
      if (m0 + j < M) {
        float re, im, rh, rl, ih, il;
        strip_eval(st, j, re, im);
        Operand<F16>::split(re, rh, rl);
        Operand<F16>::split(im, ih, il);
        const long long o = (long long)(m0 + j) * ldk + 2 * n;
        Operand<F16>::store2(Ahi, o, rh, ih);
        Operand<F16>::store2(Alo, o, rl, il);
      }
    }
  }
}
// B[2c][2n..] = (Re V, -Im V), B[2c+1][2n..] = (Im V, Re V), V_n(col) without the constant term
template <bool F16>
__global__ void __launch_bounds__(128)
    factor_cols_kernel(const double *__restrict__ table, long long b0, int nbatch, int W, long long ldk,
                       void *__restrict__ Bhi, void *__restrict__ Blo,
                       const unsigned long long *__restrict__ sep_guard) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  const int nbx = (nbatch + (int)blockDim.x - 1) / (int)blockDim.x, nby = (W + FS - 1) / FS;
  for (int vb = blockIdx.x; vb < nbx * nby; vb += gridDim.x) {      // bounded grid, see factor_rows_kernel
    const int n = (vb % nbx) * blockDim.x + threadIdx.x;
    const int c0 = (vb / nbx) * FS;
    if (n >= nbatch) continue;
    const double *t = table + (b0 + n) * 12;
    const double mu = tg_col_env_max(t[6 + 1], t[6 + 3], (double)(W - 1)) - Headroom<F16>::value;
    const Strip1D st = strip_setup(0.0, t[1], t[3], -mu, t[6 + 1], t[6 + 3], (double)c0);
#pragma unroll 4
    for (int j = 0; j < FS; ++j) {
      if (c0 + j < W) {
        float re, im, rh, rl, ih, il;
        strip_eval(st, j, re, im);
        Operand<F16>::split(re, rh, rl);
        Operand<F16>::split(im, ih, il);
        const long long o0 = (long long)(2 * (c0 + j)) * ldk + 2 * n, o1 = o0 + ldk;
        Operand<F16>::store2(Bhi, o0, rh, -ih);
        Operand<F16>::store2(Blo, o0, rl, -il);
        Operand<F16>::store2(Bhi, o1, ih, rh);
        Operand<F16>::store2(Blo, o1, il, rl);
      }
    }
  }
}

// ---- 3-product (GAUSS) operand layout: per group of 128 beamlets three blocks of 128 k-elements --------------
//   rows    A''[r][384 g + 128 j + i]:  j = 0: Ur + Ui,  1: Ur,       2: Ui
//   columns B''[c][384 g + 128 j + i]:  j = 0: Vr,       1: Vi - Vr,  2: Vr + Vi          (beamlet n = 128 g + i)
// The sums / differences are formed in fp32 (one rounding, 2^-24) BEFORE the hi / lo split.  One thread owns
// the beamlet pair (2p, 2p + 1) -> 4-byte stores, consecutive threads consecutive addresses.  Beamlets
// n >= nbatch of the last group are written as zeros (the tensor map covers whole groups).
template <bool ROWS>
__global__ void __launch_bounds__(128)
    factor_gauss_kernel(const double *__restrict__ table, long long b0, int nbatch, int npad, int s_first, int S,
                        int W, long long ldk, __half *__restrict__ hi, __half *__restrict__ lo,
                        const unsigned long long *__restrict__ peak_key,
                        const unsigned long long *__restrict__ sep_guard, int rows_pad) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  const int nbx = (npad / 2 + (int)blockDim.x - 1) / (int)blockDim.x, nby = (S + FS - 1) / FS;
  for (int vb = blockIdx.x; vb < nbx * nby; vb += gridDim.x) {      // bounded grid, see factor_rows_kernel
    const int n = 2 * ((vb % nbx) * blockDim.x + threadIdx.x);
    const int s0 = (vb / nbx) * FS;
    if (n >= npad) continue;
    Strip1D st[2];
    bool live[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      live[e] = n + e < nbatch;
      const double *t = table + (b0 + (live[e] ? n + e : 0)) * 12;
      if (ROWS) {
        double mu = tg_col_env_max(t[6 + 1], t[6 + 3], (double)(W - 1));
        mu += Headroom<true>::value - tg_prescale_G(*peak_key);
        st[e] = strip_setup(t[0], t[2], t[5], t[6 + 0] + mu, t[6 + 2], t[6 + 5], (double)(s_first + s0));
      } else {
        const double mu = tg_col_env_max(t[6 + 1], t[6 + 3], (double)(W - 1)) - Headroom<true>::value;
        st[e] = strip_setup(0.0, t[1], t[3], -mu, t[6 + 1], t[6 + 3], (double)s0);
      }
    }
    const long long kk = (long long)(n / KCH) * (3 * KCH) + (n % KCH);
#pragma unroll 2
    for (int j = 0; j < FS; ++j) {
      if (s0 + j >= S) break;
      float v[3][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float re = 0.f, im = 0.f;
        if (live[e]) strip_eval(st[e], j, re, im);
        if (ROWS) {
          v[0][e] = re + im;
          v[1][e] = re;
          v[2][e] = im;
        } else {
          v[0][e] = re;
          v[1][e] = im - re;
          v[2][e] = re + im;
        }
      }
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float h0, l0, h1, l1;
        Operand<true>::split(v[b][0], h0, l0);
        Operand<true>::split(v[b][1], h1, l1);
        // rows_pad > 0: tiled operands -- k-block k / 64 holds rows_pad rows x 64 k-elements contiguously, so that the
        // rows of a strip are 128 bytes apart instead of a whole operand row (DRAM pages, see factor_binned_kernel)
        const long long k = kk + b * KCH;
        const long long o = rows_pad > 0 ? ((k >> 6) * rows_pad + (s0 + j)) * 64 + (k & 63) : (long long)(s0 + j) * ldk + k;
        *reinterpret_cast<__half2 *>(hi + o) = __floats2half2_rn(h0, h1);
        *reinterpret_cast<__half2 *>(lo + o) = __floats2half2_rn(l0, l1);
      }
    }
  }
}

// ---- cost-aware dispatch (AUTO only) -----------------------------------------------------
// The GEMM is dense: it spends 24 TF32 flops on every beamlet*pixel whether the beamlet reaches the
// pixel or not, while the SFU kernel skips (tile, beamlet) pairs below the culling threshold.  For
// narrow beamlets on a large detector (BASELINE C3: ~11 px envelopes on 2048^2) the culled SFU sum
// executes a few per cent of the nominal evaluations and wins.  This kernel estimates the SFU work
// from the bounding box of each beamlet's region {envelope >= brightest on-detector peak - cull_bits}
// in whole 32 x 128 tiles; verdict_kernel then hands the call to the SFU path (by raising the
// separability key above 1) when the estimate is below `ratio` of the nominal nb*H*W evaluations.
__global__ void __launch_bounds__(256)
    sfu_cost_kernel(const double *__restrict__ table, long long nb, int H, int W,
                    const unsigned long long *__restrict__ gref_key, int cull_bits, double *est_tiles) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double tiles = 0.0;
  const double all_tiles = (double)((W + 127) / 128) * (double)((H + 31) / 32);
  if (i < nb) {
    const unsigned long long k = *gref_key;
    const double *e = table + i * 12 + 6;               // E(c, r) = e0 + e1 c + e2 r + e3 c^2 + e4 c r + e5 r^2 [bits]
    tiles = all_tiles;                                   // default: the beamlet reaches everything
    if (k != ~0ULL) {
      // ordered-uint key of the smallest -E on the detector (field.cu enc_ordered): decode
      const unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
      const double e_thr = -(__longlong_as_double((long long)b) + (double)cull_bits);
      const double det = e[3] * e[5] - 0.25 * e[4] * e[4];
      if (e[3] < 0.0 && e[5] < 0.0 && det > 0.0) {
        const double cs = (0.5 * e[4] * e[2] - e[5] * e[1]) / (2.0 * det);   // vertex column
        const double rs = (0.5 * e[4] * e[1] - e[3] * e[2]) / (2.0 * det);   // vertex row
        const double emax = e[0] + 0.5 * (e[1] * cs + e[2] * rs);
        const double d = emax - e_thr;
        if (d < 0.0) {
          tiles = 0.0;
        } else if (isfinite(d) && isfinite(cs) && isfinite(rs)) {
          const double hc = sqrt(d * (-e[5]) / det), hr = sqrt(d * (-e[3]) / det);
          const double c_lo = fmax(cs - hc, 0.0), c_hi = fmin(cs + hc, (double)(W - 1));
          const double r_lo = fmax(rs - hr, 0.0), r_hi = fmin(rs + hr, (double)(H - 1));
          if (c_hi < c_lo || r_hi < r_lo) tiles = 0.0;
          else tiles = (floor(c_hi / 128.0) - floor(c_lo / 128.0) + 1.0) * (floor(r_hi / 32.0) - floor(r_lo / 32.0) + 1.0);
        }
      }
    }
    if (!(tiles == tiles)) tiles = all_tiles;
  }
  for (int o = 16; o > 0; o >>= 1) tiles += __shfl_xor_sync(0xffffffffu, tiles, o);
  if ((threadIdx.x & 31) == 0 && tiles > 0.0) atomicAdd(est_tiles, tiles);
}
// split != NULL: the cost verdict goes to *split (1 = the beamlets are sparse on this grid) and the key keeps the pure
// separability verdict -- the host then picks the tile-binned tensor sum for "separable and sparse"
__global__ void verdict_kernel(unsigned long long *key, const double *est_tiles, double nominal_evals, double ratio,
                               unsigned long long *split) {
  if (*est_tiles * 4096.0 < ratio * nominal_evals) {
    if (split) *split = 1ULL;
    else atomicMax(key, (unsigned long long)__double_as_longlong(2.0));   // > 1: "not for the tensor path"
  }
}

__global__ void __launch_bounds__(256)
    f64_to_c64_kernel(const double *__restrict__ in, float *__restrict__ out, size_t n,
                      const unsigned long long *__restrict__ sep_guard, const TgPeers peers) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = (float)in[i];
    out[i] = v;
    for (int p = 0; p < peers.n; ++p) static_cast<float *>(peers.ptr[p])[i] = v;
  }
}

// ---------------------------------------------------------------- tile-binned (block-sparse K) separable sum
// BASELINE C3 is separable but SPARSE: 1e5 beamlets whose footprints {envelope >= brightest on-detector peak - cull_bits}
// are ellipses of ~260 px on a 2048^2 detector, so the dense GEMM spends 98 % of its tensor work on (beamlet, pixel)
// pairs below the culling threshold, and the culled SFU kernel evaluates the remaining 1.4 % pixel by pixel.  Here the
// sum stays on the tensor cores but every 128-row x 64-column output tile only multiplies the beamlets whose footprint
// meets it (measured on B200: C3 0.93 ms as a graph replay against 5.6 ms culled SFU / 6.8 ms dense GEMM):
//   * bin_mark_kernel: one thread per beamlet sets its bit in the word of every tile its elliptical footprint
//     {envelope >= threshold} meets (hits[tile][beamlet / 32], atomicOr: ~1e6 of them at C3);
//   * bin_tiles_kernel<false>: per tile (one CTA, each of its 8 warps owns a contiguous eighth of the tile's words) the
//     number of set bits; bin_prefix_kernel: chunks of 128 beamlets per tile -> P[0..T] (exclusive scan), the chunks per
//     CTA of the ragged GEMM schedule, the overflow verdict against the operand capacity;
//   * bin_tiles_kernel<true>: the set bits expanded IN BEAMLET ORDER (deterministic sums) into the tile's slots
//     sel[P[t] * 128 ...], padded with -1 to whole chunks; chunk -> tile map c2t;
//   * factor_binned_kernel: row / column factors of slot s for the 128 rows / 64 columns of ITS tile, on one
//     concatenated k axis (chunk g at k = 256 g), stored k-block by k-block: each 128-row x 64-k tile of an operand is
//     16 KiB of contiguous memory, so the factor kernels write and the GEMM's TMA loads read whole DRAM pages;
//   * gemm_x3_kernel with the ragged schedule (SkSched::bins): 148 CTAs split the chunk axis evenly, tiles cut by a CTA
//     boundary are fixed up by the last arriver like every stream-K piece.
// Per (beamlet, tile) pair: 192 factor evaluations and 2 KB of operands instead of up to 8192 SFU pixel evaluations;
// the operands are read once each (no reuse across tiles), so the kernel streams them at the HBM rate.
constexpr int BIN_SLOTS = CHUNK_K / 2;   // beamlets per accumulation chunk (4-multiplication real form: 2 k per beamlet)
constexpr int BIN_TN = BN / 2;           // complex columns per tile
constexpr int BIN_WARPS = 8;

// One thread per beamlet: the axis-aligned ellipse {envelope >= threshold} in detector pixels (+1 px slack on both
// half-axes; separable beamlets have no cross term to speak of; same threshold as field.cu's bbox_kernel), then every
// tile of its bounding box whose pixel rectangle the ellipse meets gets bit (i mod 32) of word hits[t][i / 32] set.
// Non-concave or non-finite envelopes mark every tile, so NaN beamlets poison the image as in the dense sum.
// probe mode (TG_BIN_PROBE): ctl[0] = separability key, ctl[BIN_CTL_SPLIT] = 1 when the cost model found the beamlets
// sparse; the binning kernels of a call that AUTO will hand to another path return at once
constexpr int BIN_CTL_SPLIT = 9;
__device__ __forceinline__ bool bin_probe_skip(const unsigned long long *ctl) {
  return ctl && (!tg_key_is_separable(ctl[0]) || ctl[BIN_CTL_SPLIT] == 0ULL);
}
__global__ void __launch_bounds__(256)
    bin_mark_kernel(long long nb, const double *__restrict__ table, int H, int W, int row0, int nrows, int tiles_n,
                    const unsigned long long *__restrict__ gref_key, int cull_bits, unsigned *__restrict__ hits,
                    const unsigned long long *__restrict__ probe) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb || bin_probe_skip(probe)) return;
  const long long wpt = (nb + 31) / 32;
  int tn_lo = 0, tn_hi = (W - 1) / BIN_TN, tm_lo = 0, tm_hi = (nrows - 1) / BM;
  float cx = 0.f, cy = 0.f, hx = 1e30f, hy = 1e30f;
  const unsigned long long k = *gref_key;
  const double *e = table + i * 12 + 6;   // E(c, r) = e0 + e1 c + e2 r + e3 c^2 + e4 c r + e5 r^2 [bits]
  const double det = e[3] * e[5] - 0.25 * e[4] * e[4];
  if (k != ~0ULL && e[3] < 0.0 && e[5] < 0.0 && det > 0.0) {
    const double e_thr = -(tg_dec_ordered(k) + (double)cull_bits);
    const double cs = (0.5 * e[4] * e[2] - e[5] * e[1]) / (2.0 * det);
    const double rs = (0.5 * e[4] * e[1] - e[3] * e[2]) / (2.0 * det);
    const double d = (e[0] + 0.5 * (e[1] * cs + e[2] * rs)) - e_thr;
    if (d < 0.0) return;                                      // below the threshold everywhere
    if (isfinite(d) && isfinite(cs) && isfinite(rs)) {
      // (fp32 tests: the centre rounds by < 2^-24 of its magnitude -- widen the half-axes by that much)
      const double slack = (fabs(cs) + fabs(rs)) * 1.2e-7;
      const double hc = sqrt(d * (-e[5]) / det) + 1.0 + slack, hr = sqrt(d * (-e[3]) / det) + 1.0 + slack;
      const double c_lo = fmax(floor(cs - hc), 0.0), c_hi = fmin(ceil(cs + hc), (double)(W - 1));
      const double r_lo = fmax(floor(rs - hr), (double)row0), r_hi = fmin(ceil(rs + hr), (double)(row0 + nrows - 1));
      if (c_hi < c_lo || r_hi < r_lo) return;                 // off the detector / outside these rows
      tn_lo = (int)c_lo / BIN_TN;
      tn_hi = (int)c_hi / BIN_TN;
      tm_lo = ((int)r_lo - row0) / BM;
      tm_hi = ((int)r_hi - row0) / BM;
      cx = (float)cs;
      cy = (float)rs;
      hx = (float)fmin(hc, 1e30);
      hy = (float)fmin(hr, 1e30);
    }
  }
  const unsigned bit = 1u << (unsigned)(i & 31);
  const float hh = hx * hy;
  for (int tm = tm_lo; tm <= tm_hi; ++tm) {
    const float r0 = (float)(row0 + tm * BM), r1 = (float)(row0 + min(tm * BM + BM - 1, nrows - 1));
    const float dy = fmaxf(fmaxf(r0 - cy, cy - r1), 0.f) * hx;
    for (int tn = tn_lo; tn <= tn_hi; ++tn) {
      // closest point of the tile's pixel rectangle to the centre, inside the ellipse?
      const float c0 = (float)(tn * BIN_TN), c1 = (float)min(tn * BIN_TN + BIN_TN - 1, W - 1);
      const float dx = fmaxf(fmaxf(c0 - cx, cx - c1), 0.f) * hy;
      if (dx * dx + dy * dy <= hh * hh * 1.000001f) atomicOr(hits + (long long)(tm * tiles_n + tn) * wpt + (i >> 5), bit);
    }
  }
}

// One CTA per tile; warp w owns the beamlets [w seg, (w + 1) seg) (seg a multiple of 32) = a run of the tile's words.
// FILL = false: wc[t][w] = set bits of warp w's words.  FILL = true: the set bits ARE the tile's beamlets in order --
// append them behind the beamlets of the warps before, pad the tile's last chunk with -1, fill c2t.
template <bool FILL>
__global__ void __launch_bounds__(32 * BIN_WARPS)
    bin_tiles_kernel(long long nb, int T, int *__restrict__ wc, const unsigned *__restrict__ hits,
                     const int *__restrict__ bins, int *__restrict__ sel, int *__restrict__ c2t,
                     const unsigned long long *__restrict__ probe) {
  if (FILL && bins[BIN_OVERFLOW]) return;
  if (bin_probe_skip(probe)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seg = (((nb + BIN_WARPS - 1) / BIN_WARPS + 31) / 32) * 32;
  const long long wpt = (nb + 31) / 32;                      // words per tile
  const long long g_lo = (long long)warp * (seg / 32), g_hi = (g_lo + seg / 32 < wpt) ? g_lo + seg / 32 : wpt;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const unsigned *hw = hits + (long long)t * wpt;
    long long base = 0;
    int total = 0;
    if (FILL) {
      const int *P = bins + BIN_HDR;
      base = (long long)P[t] * BIN_SLOTS;
      for (int w = 0; w < BIN_WARPS; ++w) {
        const int c = wc[t * BIN_WARPS + w];
        base += (w < warp) ? c : 0;
        total += c;
      }
    }
    int pos = 0;
    for (long long g0 = g_lo; g0 < g_hi; g0 += 32) {
      const long long g = g0 + lane;
      unsigned m = g < g_hi ? hw[g] : 0u;
      const int p = __popc(m);
      if (!FILL) {
        pos += p;
      } else {
        int incl = p;                                        // the lanes' words in order: exclusive scan of the bit counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        long long dst = base + pos + incl - p;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          sel[dst++] = (int)(g * 32 + b);
        }
        pos += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if (!FILL) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
      if (lane == 0) wc[t * BIN_WARPS + warp] = pos;
    } else {
      const int *P = bins + BIN_HDR;
      const int p0 = P[t], p1 = P[t + 1];
      if (warp == 0)
        for (long long s = (long long)p0 * BIN_SLOTS + total + lane; s < (long long)p1 * BIN_SLOTS; s += 32) sel[s] = -1;
      if (warp == 1)
        for (int g = p0 + lane; g < p1; g += 32) c2t[g] = t;
    }
  }
}

// P[t] = chunks before tile t; bins header: total chunks (0 on overflow), chunks per GEMM CTA, overflow flag, chunks needed
__global__ void __launch_bounds__(1024)
    bin_prefix_kernel(int T, const int *__restrict__ wc, int *__restrict__ bins, int G, int cap_chunks,
                      const unsigned long long *__restrict__ probe) {
  if (bin_probe_skip(probe)) return;
  __shared__ long long wsum[32];
  __shared__ long long carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  int *P = bins + BIN_HDR;
  for (int base = 0; base < T; base += 1024) {
    const int t = base + threadIdx.x;
    long long cnt = 0;
    if (t < T)
      for (int w = 0; w < BIN_WARPS; ++w) cnt += wc[t * BIN_WARPS + w];
    const long long ch = (cnt + BIN_SLOTS - 1) / BIN_SLOTS;
    long long inc = ch;
    for (int o = 1; o < 32; o <<= 1) {
      const long long v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    long long before = carry_s;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    const long long excl = before + inc - ch;
    if (t < T) P[t] = (int)(excl < 0x7fffffffLL ? excl : 0x7fffffffLL);
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long long tot = carry_s;
    const bool over = tot > (long long)cap_chunks;
    bins[BIN_OVERFLOW] = over ? 1 : 0;
    bins[BIN_NEED] = (int)(tot < 0x7fffffffLL ? tot : 0x7fffffffLL);   // chunks the beamlets need (also when over capacity)
    bins[BIN_CTOT] = over ? 0 : (int)tot;
    const long long q = (tot + G - 1) / G;
    bins[BIN_Q] = (int)(q > 0 ? q : 1);
    P[T] = over ? 0 : (int)tot;
  }
  __syncthreads();
  if (bins[BIN_OVERFLOW])
    for (int t = threadIdx.x; t < T; t += 1024) P[t] = 0;
}

// One CTA per operand chunk (128 slots), one thread per slot s, n = sel[s]: the beamlet's table entry is read once, then
//   A[k-block s / 32][row][2 (s % 32) ..]       = U_n(row)                for the 128 rows of the chunk's tile,
//   B[k-block s / 32][2 c][..] = (Re V, -Im V),  B[..][2 c + 1][..] = (Im V, Re V)   for its 64 columns c
// in strips of 32 (fp64 strip setup, fp32 recurrences: the dense factor kernels' arithmetic), each value split into
// fp16 hi + lo as one packed pair.  Rows and columns in ONE kernel: the row strips are issue-bound, the column strips
// (four stores per value) write-bound, and CTAs in different phases overlap the two (separate kernels: 257 + 218 us).
// Tiled layout: k-block kb of an operand holds rows [0, 128) x 64 k-elements contiguously (16 KiB).
// (64 registers, 8 CTAs per SM.  Measured without effect on C3: 10 CTAs per SM at 48 registers, and an L1 prefetch of the
// next chunk's table rows -- 0.921-0.935 ms per image in all four combinations: the kernel is bound by its writes.)
template <bool F16>
__global__ void __launch_bounds__(BIN_SLOTS)
    factor_binned_kernel(const double *__restrict__ table, const int *__restrict__ bins, const int *__restrict__ sel,
                         const int *__restrict__ c2t, int tiles_n, int row0, int M, int W, void *__restrict__ Ahi,
                         void *__restrict__ Alo, void *__restrict__ Bhi, void *__restrict__ Blo,
                         const unsigned long long *__restrict__ peak_key,
                         const unsigned long long *__restrict__ sep_guard) {
  static_assert(F16, "the tile-binned sum uses fp16 x 3 operands");
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  const int ctot = bins[BIN_CTOT];
  const double G = tg_prescale_G(*peak_key);
  for (int g = blockIdx.x; g < ctot; g += gridDim.x) {
    const long long s = (long long)g * BIN_SLOTS + threadIdx.x;
    const int n = sel[s];
    const int tile = c2t[g], tm = tile / tiles_n, tn = tile - tm * tiles_n;
    double t[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = n >= 0 ? __ldg(table + (long long)n * 12 + i) : 0.0;
    const double cm = tg_col_env_max(t[6 + 1], t[6 + 3], (double)(W - 1));
    const long long kb_rows = (s >> 5) * BM;              // first row of the slot's k-block in the tiled operands
    const int kk = 2 * (int)(s & 31);
#pragma unroll 1
    for (int sidx = 0; sidx < BM / FS; ++sidx) {
      const int m0 = tm * BM + sidx * FS;                 // first row of the strip, relative to row0
      const Strip1D st = strip_setup(t[0], t[2], t[5], t[6 + 0] + cm + Headroom<F16>::value - G, t[6 + 2], t[6 + 5],
                                     (double)(row0 + m0));
#pragma unroll 4
      for (int j = 0; j < FS; ++j) {
        float re = 0.f, im = 0.f;
        if (n >= 0 && m0 + j < M) strip_eval(st, j, re, im);
        // (re, im) split as one packed pair: hi = fp16(x), lo = fp16(x - hi) -- the values Operand<true>::split gives
        const __half2 hi = __floats2half2_rn(re, im);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(re - hf.x, im - hf.y);
        const long long o = (kb_rows + (sidx * FS + j)) * 64 + kk;
        *reinterpret_cast<__half2 *>(static_cast<__half *>(Ahi) + o) = hi;
        *reinterpret_cast<__half2 *>(static_cast<__half *>(Alo) + o) = lo;
      }
    }
#pragma unroll 1
    for (int sidx = 0; sidx < BIN_TN / FS; ++sidx) {
      const int c0 = tn * BIN_TN + sidx * FS;
      const Strip1D st = strip_setup(0.0, t[1], t[3], Headroom<F16>::value - cm, t[6 + 1], t[6 + 3], (double)c0);
#pragma unroll 4
      for (int j = 0; j < FS; ++j) {
        float re = 0.f, im = 0.f;
        if (n >= 0 && c0 + j < W) strip_eval(st, j, re, im);
        const __half2 hi = __floats2half2_rn(re, im);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(re - hf.x, im - hf.y);
        // row 2 c: (Re, -Im) = the pair with the sign bit of its upper half flipped; row 2 c + 1: (Im, Re) = halves swapped
        const unsigned hb = *reinterpret_cast<const unsigned *>(&hi), lb = *reinterpret_cast<const unsigned *>(&lo);
        const long long o0 = (kb_rows + 2 * (sidx * FS + j)) * 64 + kk, o1 = o0 + 64;
        *reinterpret_cast<unsigned *>(static_cast<__half *>(Bhi) + o0) = hb ^ 0x80000000u;
        *reinterpret_cast<unsigned *>(static_cast<__half *>(Blo) + o0) = lb ^ 0x80000000u;
        *reinterpret_cast<unsigned *>(static_cast<__half *>(Bhi) + o1) = __byte_perm(hb, 0, 0x1032);
        *reinterpret_cast<unsigned *>(static_cast<__half *>(Blo) + o1) = __byte_perm(lb, 0, 0x1032);
      }
    }
  }
}
// tiles no beamlet reaches get no GEMM unit: their pixels are zeroed here (in `out` and in every peer image), every
// other tile is written whole by its final GEMM unit
__global__ void __launch_bounds__(256)
    bin_zero_empty_kernel(const int *__restrict__ bins, int T, int tiles_n, int M, int Np, double *__restrict__ out,
                          long long ldo, const TgPeers peers, const unsigned long long *__restrict__ sep_guard) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  const int *P = bins + BIN_HDR;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    if (P[t + 1] != P[t]) continue;
    const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
    for (int idx = threadIdx.x; idx < BM * (BN / 2); idx += blockDim.x) {
      const int row = m0 + idx / (BN / 2), col = n0 + 2 * (idx % (BN / 2));
      if (row < M && col < Np) {                   // Np = 2 W is even: (col, col + 1) is one complex pixel
        const long long o = (long long)row * ldo + col;
        *reinterpret_cast<double2 *>(out + o) = make_double2(0.0, 0.0);
        for (int p = 0; p < peers.n; ++p)
          *reinterpret_cast<double2 *>(static_cast<double *>(peers.ptr[p]) + o) = make_double2(0.0, 0.0);
      }
    }
  }
}
// Row-sharded multi-GPU sum: this rank's finished rows go into every peer's image as whole 512-byte warp stores over
// NVLink.  (The dense GEMM's epilogue stores its tiles to the peers itself, 16 bytes per lane a detector row apart: fine
// for the 2 MB of a C2 shard, but the 59 MB a C3 shard sends to seven peers took longer than the sum -- measured 0.64 ms
// per image at N = 8 against 0.67 ms at N = 2.)
__global__ void __launch_bounds__(256)
    peer_push_kernel(const double2 *__restrict__ src, size_t n, const TgPeers peers,
                     const unsigned long long *__restrict__ sep_guard) {
  if (sep_guard && !tg_key_is_separable(*sep_guard)) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = src[i];
    for (int p = 0; p < peers.n; ++p) static_cast<double2 *>(peers.ptr[p])[i] = v;
  }
}
// capacity overflow (a captured graph replayed on beamlets that need more operand chunks than it was built for) or a
// non-separable input under capture: poison the output instead of returning a partial sum
__global__ void __launch_bounds__(256)
    nan_fill_binned_kernel(double *__restrict__ out, size_t n, const int *__restrict__ bins,
                           const unsigned long long *__restrict__ key) {
  if (!bins[BIN_OVERFLOW] && (!key || tg_key_is_separable(*key))) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __longlong_as_double(0x7ff8000000000000LL);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// TG_E2E_STREAM=1: ONE GEMM launch for the whole image, walked in rounds of row blocks (SkSched::rounds).  Measured
// slower than a launch per block -- a round of 256 rows is split 9 ways along K and the last piece of every tile
// sums nine partial tiles (C2: 0.51 ms of kernels against 0.41 ms) -- so it is opt-in.  Read on every call.
bool stream_rounds_enabled() {
  const char *e = getenv("TG_E2E_STREAM");
  return e && atoi(e) != 0;
}
// TG_E2E_FLAGGED=1: the per-block launches are chained with programmatic dependent launch and raise flag words that
// the call polls, instead of being separated by events.  Measured on C2 (4 blocks of 256 rows): the kernels take the
// same 0.42 ms (a block's launch costs ~20 us beyond its tensor work whether or not its ramp overlaps the previous
// tail -- tools/exp_gemm3.py), and the host-polled copies start a little later than event-driven ones (0.585 vs
// 0.568 ms per image), so this too is opt-in.
bool flagged_blocks_enabled() {
  const char *e = getenv("TG_E2E_FLAGGED");
  return e && atoi(e) != 0;
}

// 2-D row-major operand (rows x K elements of 4 (tf32) or 2 (fp16) bytes, pitch ldk elements),
// box = 128 rows x 128 bytes, 128B swizzle
template <bool F16>
int make_map(CUtensorMap *m, const void *base, long long rows, long long K, long long ldk, int box_rows = BM) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    tg_set_error("cuTensorMapEncodeTiled entry point not available");
    return TG_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldk * GemmCfg<F16>::ELEM};
  cuuint32_t box[2] = {(cuuint32_t)GemmCfg<F16>::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  BK_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tg_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TG_ECUDA;
  }
  return TG_OK;
}

// Host side of the work decomposition (see SkSched).  Stream-K is used when it pays: the tiles do not fill
// whole waves of the `sms` CTAs and K is deep enough (>= 8 chunks per tile) for the pieces to amortise
// their partial-tile round trip.  TG_GEMM_STREAMK=0 forces the plain one-tile-at-a-time schedule (A/B runs).
// TG_GEMM_PAIR=1 enables the CTA-pair (cta_group::2) kernel for shapes whose pair tiles fill the machine
bool pair_mode() {
  static const bool on = [] {
    const char *e = getenv("TG_GEMM_PAIR");
    return e && atoi(e) != 0;
  }();
  return on;
}
int streamk_mode_default() {
  static const int mode = [] {
    const char *e = getenv("TG_GEMM_STREAMK");
    const int v = e ? atoi(e) : 1;
    return (v >= 0 && v <= 4) ? v : 1;
  }();
  return mode;
}
SkSched make_sched(int M, int Np, int K, int BK, int chunk_kb, int sms, int mode = -1, int round_tiles_m = 0) {
  if (mode < 0 || mode > 4) mode = streamk_mode_default();   // 3 = like 1 but head / tail only, 4 = like 1 with split-K tails
  SkSched s;
  const int tiles_m = (M + BM - 1) / BM;
  s.tiles_n = (Np + BN - 1) / BN;
  s.T = tiles_m * s.tiles_n;
  s.nkb = (K + BK - 1) / BK;
  s.nch = (s.nkb + chunk_kb - 1) / chunk_kb;
  s.G = s.T < sms ? s.T : sms;
  s.R = 0;
  s.q = s.nch;
  s.Tl = 0;
  s.qh = 1;
  s.maxparts = 1;
  s.S = 1;
  s.prefetch = 0;
  s.bins = nullptr;
  s.tiled_a = s.tiled_b = 0;
  s.uniform = 0;
  s.rounds = 1;
  s.Tr = s.T;
  if (round_tiles_m > 0 && tiles_m > round_tiles_m && s.tiles_n * round_tiles_m <= sms) {
    // row-block streaming (host pipeline): rounds of round_tiles_m tile rows, each split over the machine
    s.Tr = s.tiles_n * round_tiles_m;
    s.rounds = (tiles_m + round_tiles_m - 1) / round_tiles_m;
    int S = sms / s.Tr;
    if (S > s.nch / 4) S = s.nch / 4 > 0 ? s.nch / 4 : 1;
    s.q = (s.nch + S - 1) / S;
    s.maxparts = (s.nch + s.q - 1) / s.q;
    s.G = s.Tr * s.maxparts;
    s.R = s.T;
    s.uniform = 1;
    return s;
  }
  // TG_GEMM_STREAMK: 0 = never split, 2 = split whenever the tiles leave a partial wave (experiments); default
  // 1 = split only when whole tiles would leave more than 30 % of the machine idle.  Measured on B200 (fp16 x 3,
  // N = 2048, K = 20000): 1024 rows (128 tiles, 86 % of a wave) 0.212 ms whole tiles vs 0.27 ms split -- the
  // 20 extra SMs do not pay for the partial-tile traffic and the helpers' out-of-step operand reads; 512 rows
  // 0.197 -> 0.126 ms, 256 rows 0.203 -> 0.085 ms, 128 rows (the row shard of one of 8 ranks) 0.205 -> 0.065 ms.
  const double waves = (double)s.T / sms;
  const double dp_eff = waves / ceil(waves);
  const bool want = mode == 2 || ((mode == 1 || mode == 3 || mode == 4) && dp_eff < 0.7);
  if (want && s.nch >= 8 && sms > 1 && 2 * s.T <= sms && mode != 3) {      // (mode 4: split-K pieces + tails on the SMs left over)
    // Few tiles (a row shard, a row block of the host pipeline, the 64 complex tiles of C2 in the 3-product form):
    // PLAIN split-K, every tile cut into S = floor(sms / T) equal k-ranges, CTA c -> (tile c mod T, range c div T).
    // All tiles' range s is walked at the same time by the T CTAs of group s, so operand tiles are fetched from HBM
    // once and shared through L2 within a group -- the head / tail arrangement below uses every SM but its helper
    // CTAs each walk a k-range of their own, and with half the work in helpers (T = 64 on 148 SMs) their unshared
    // operand reads made the kernel HBM-bound (3-product C2: 0.220 ms head / tail, 0.191 ms plain split-K).
    int S = sms / s.T;
    if (S > s.nch / 4) S = s.nch / 4 > 0 ? s.nch / 4 : 1;      // at least ~4 chunks per piece
    // ... and when S pieces per tile leave SMs over (64 tiles x 2 = 128 of 148), the pieces shrink to q = ceil(T nch /
    // sms) chunks and the remaining nch - S q chunks of every tile -- the SAME k-range for all tiles -- are laid end
    // to end for the sms - S T helper CTAs: every SM gets the same work, and only a small part of it (13 % for C2)
    // is done outside the lockstep groups.  MEASURED (B200, C2 3-product GEMM): 0.186 ms against 0.178 ms for the plain
    // split on 128 SMs -- the GEMM runs at the board's power cap, where 20 more SMs reading k-ranges nobody shares buy
    // nothing -- so this arrangement is opt-in (mode 4) and the default stays the plain split
    const int helpers = sms - S * s.T;
    const int q_all = (int)(((long long)s.T * s.nch + sms - 1) / sms);
    if (S > 1 && mode == 4 && helpers >= sms / 16 && S * q_all < s.nch && s.nch - S * q_all >= 4) {
      s.S = S;
      s.q = q_all;
      s.Tl = s.nch - S * q_all;
      s.qh = (int)(((long long)s.T * s.Tl + helpers - 1) / helpers);
      s.R = s.T;
      s.G = sms;
      for (int t = 0; t < s.R; ++t) {
        const int n = sk_nparts(s, t);
        if (n > s.maxparts) s.maxparts = n;
      }
      return s;
    }
    if (S > 1) {
      s.q = (s.nch + S - 1) / S;
      s.maxparts = (s.nch + s.q - 1) / s.q;                    // non-empty ranges
      s.G = s.T * s.maxparts;
      s.R = s.T;
      s.uniform = 1;
      s.Tl = 0;
      return s;
    }
  }
  if (want && s.nch >= 8 && sms > 1) {
    int G = sms, R = s.T % sms;
    if (s.T < sms) {
      const long long cap = (long long)s.T * s.nch / 4;   // at least ~4 chunks per CTA
      G = (int)(cap < sms ? (cap > s.T ? cap : s.T) : sms);
      R = G > s.T ? s.T : 0;
    }
    if (R > 0 && G > R) {
      static const int pf = [] {            // TG_GEMM_PREFETCH: k-blocks of L2 prefetch distance in stream-K schedules
        const char *e = getenv("TG_GEMM_PREFETCH");
        const int v = e ? atoi(e) : 0;      // measured on B200: any distance makes the pieces SLOWER (512 rows:
        return (v >= 0 && v <= 64) ? v : 0;  // 0.125 ms -> 0.177 ms at 4..12 k-blocks), so the default is off
      }();
      s.prefetch = pf;
      s.G = G;
      s.R = R;
      s.q = (int)(((long long)R * s.nch + G - 1) / G);
      s.Tl = s.nch - s.q;
      if (s.Tl <= 0) {
        s.q = s.nch;
        s.Tl = 0;
      } else {
        s.qh = (int)(((long long)R * s.Tl + (G - R) - 1) / (G - R));
        for (int t = 0; t < R; ++t) {
          const int n = sk_nparts(s, t);
          if (n > s.maxparts) s.maxparts = n;
        }
      }
    } else if (s.T >= sms) {
      s.G = sms;
    }
  } else if (s.T >= sms) {
    s.G = sms;
  }
  return s;
}

// Stream-K scratch supplied by the caller (tg_separable_run keeps it in its workspace: a cudaMallocAsync +
// memset + free per GEMM launch cost 12-15 us, measured): [0, cnt_cap) arrival counters, zero (the kernel
// leaves them zero), then the partial tiles.
struct SkWs {
  unsigned char *p = nullptr;
  size_t cnt_cap = 0, bytes = 0;
  unsigned char *parts = nullptr;     // partial tiles when they do not follow the counters directly
};
// grow-only scratch of the direct GEMM entry points (see launch_gemm); counters occupy the first cnt_cap bytes
struct SkCache {
  unsigned char *p = nullptr;
  size_t bytes = 0;
  static constexpr size_t cnt_cap = 16384;      // >= 147 tiles x 16 warps x 4 bytes, rounded up
  cudaStream_t last = nullptr;
  cudaEvent_t ev = nullptr;
  bool used = false;
};
inline SkCache *sk_cache_get(size_t part_bytes, cudaStream_t st) {
  static thread_local SkCache caches[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SkCache &c = caches[dev];
  const size_t need = SkCache::cnt_cap + part_bytes;
  if (!c.ev && cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (c.bytes < need) {
    if (c.p) {
      cudaEventSynchronize(c.ev);
      cudaFree(c.p);
      c.p = nullptr;
      c.bytes = 0;
    }
    const size_t want = need + need / 4;
    if (cudaMalloc(reinterpret_cast<void **>(&c.p), want) != cudaSuccess) {
      cudaGetLastError();
      c.p = nullptr;
      return nullptr;
    }
    if (cudaMemset(c.p, 0, SkCache::cnt_cap) != cudaSuccess) return nullptr;
    c.bytes = want;
    c.used = false;
  }
  if (c.used && c.last != st) cudaStreamWaitEvent(st, c.ev, 0);
  return &c;
}
inline void sk_cache_release(SkCache *c, cudaStream_t st) {
  cudaEventRecord(c->ev, st);
  c->last = st;
  c->used = true;
}
inline int device_sms(int *sms_out) {
  int dev = 0, sms = 148;
  TG_CUDA(cudaGetDevice(&dev));
  static int sms_cache[64] = {0};
  if (dev >= 0 && dev < 64 && sms_cache[dev] > 0) {
    sms = sms_cache[dev];
  } else {
    TG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) sms_cache[dev] = sms;
  }
  *sms_out = sms;
  return TG_OK;
}
template <bool F16, bool GAUSS>
void sk_scratch_need(int M, int Np, int K, int sms, size_t *cnt_bytes, size_t *part_bytes, int round_tiles_m = 0) {
  constexpr int NE = GAUSS ? 16 : 8;
  const SkSched sched = make_sched(M, Np, K, GemmCfg<F16>::BK, GemmCfg<F16>::CHUNK_KB, sms, -1, round_tiles_m);
  const bool split = sched.R > 0 && sched.maxparts > 1;
  *cnt_bytes = split ? (((size_t)sched.R * NE * sizeof(unsigned int)) + 255) / 256 * 256 : 0;
  *part_bytes = split ? (size_t)sched.R * sched.maxparts * (size_t)(NE * 2048) * sizeof(float) : 0;
}

// Row-block streaming of one launch (see SkSched::rounds): cnt[rounds] arrival counters (zero on entry, left zero),
// flag[round] = epoch when every tile of the round is in memory.
struct SkStream {
  int round_tiles_m = 0;
  unsigned int *cnt = nullptr, *flag = nullptr;
  unsigned int epoch = 1;
  bool pdl = false;       // launch with programmatic stream serialisation (may overlap the GEMM launched before it)
};

template <bool F16, bool GAUSS = false>
int launch_gemm(const void *Ahi, const void *Alo, const void *Bhi, const void *Blo, int M, int Np, int K,
                long long ldk, double *out, long long ldo, int accumulate, const unsigned long long *peak_key,
                const unsigned long long *sep_guard, cudaStream_t st, const TgPeers &peers,
                const SkWs *skws = nullptr, const SkStream *strm = nullptr, int tiled_a = 0, int tiled_b = 0) {
  CUtensorMap ta, tb, tc, td;
  int rc;
  if (tiled_a > 0) {
    // tiled operands (K a whole number of k-blocks): tensors of 128-byte rows, k-block kb at rows [kb rows_pad, ...)
    constexpr int BKe = GemmCfg<F16>::BK;
    const long long nkb = (K + BKe - 1) / BKe;
    if ((rc = make_map<F16>(&ta, Ahi, nkb * tiled_a, BKe, BKe)) != TG_OK) return rc;
    if ((rc = make_map<F16>(&tb, Alo, nkb * tiled_a, BKe, BKe)) != TG_OK) return rc;
    if ((rc = make_map<F16>(&tc, Bhi, nkb * tiled_b, BKe, BKe)) != TG_OK) return rc;
    if ((rc = make_map<F16>(&td, Blo, nkb * tiled_b, BKe, BKe)) != TG_OK) return rc;
  } else {
    if ((rc = make_map<F16>(&ta, Ahi, M, K, ldk)) != TG_OK) return rc;
    if ((rc = make_map<F16>(&tb, Alo, M, K, ldk)) != TG_OK) return rc;
    if ((rc = make_map<F16>(&tc, Bhi, Np, K, ldk)) != TG_OK) return rc;
    if ((rc = make_map<F16>(&td, Blo, Np, K, ldk)) != TG_OK) return rc;
  }
  int sms = 148;
  if ((rc = device_sms(&sms)) != TG_OK) return rc;
  if constexpr (!GAUSS) {
    // CTA-pair kernel (cta_group::2) when whole 256 x 128 pair tiles keep >= 70 % of the SM pairs busy
    const int tiles_n = (Np + BN - 1) / BN, pair_rows = ((M + BM - 1) / BM + 1) / 2;
    const int n_pair_tiles = tiles_n * pair_rows, slots = sms / 2;
    const double waves = (double)n_pair_tiles / slots;
    if (pair_mode() && !strm && M > BM && slots > 0 && waves / ceil(waves) >= 0.7) {
      CUtensorMap tcp, tdp;
      if ((rc = make_map<F16>(&tcp, Bhi, Np, K, ldk, BN / 2)) != TG_OK) return rc;
      if ((rc = make_map<F16>(&tdp, Blo, Np, K, ldk, BN / 2)) != TG_OK) return rc;
      const size_t psmem = (size_t)P_STAGES * P_STAGE_BYTES + sizeof(PairSmemCtl) + 1024;
      TG_CUDA(cudaFuncSetAttribute(gemm_x3_pair_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      const int n_clusters = n_pair_tiles < slots ? n_pair_tiles : slots;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(2 * n_clusters));
      cfg.blockDim = dim3(GEMM_THREADS);
      cfg.dynamicSmemBytes = psmem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const double hr = Headroom<F16>::value;
      TG_CUDA(cudaLaunchKernelEx(&cfg, gemm_x3_pair_kernel<F16>, ta, tb, tcp, tdp, M, Np, K, out, ldo, accumulate,
                                 peak_key, hr, sep_guard, peers, tiles_n, n_pair_tiles));
      return tg_launch_check(F16 ? "gemm_x3_pair_kernel<f16>" : "gemm_x3_pair_kernel<tf32>");
    }
  }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + sizeof(GemmSmemCtl) + 1024;
  TG_CUDA(cudaFuncSetAttribute(gemm_x3_kernel<F16, GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rtm = strm ? strm->round_tiles_m : 0;
  SkSched sched = make_sched(M, Np, K, GemmCfg<F16>::BK, GemmCfg<F16>::CHUNK_KB, sms, -1, rtm);
  sched.tiled_a = tiled_a;
  sched.tiled_b = tiled_b;
  size_t cnt_bytes = 0, part_bytes = 0;
  sk_scratch_need<F16, GAUSS>(M, Np, K, sms, &cnt_bytes, &part_bytes, rtm);
  unsigned char *sk = nullptr, *parts = nullptr;
  bool own = false;
  SkCache *cache = nullptr;
  if (cnt_bytes) {
    if (skws && skws->p && cnt_bytes <= skws->cnt_cap && skws->cnt_cap + part_bytes <= skws->bytes) {
      sk = skws->p;                          // counters are zero on entry by contract
      parts = skws->parts ? skws->parts : skws->p + skws->cnt_cap;
    } else if (!tg_stream_is_capturing(st) && cnt_bytes <= SkCache::cnt_cap &&
               (cache = sk_cache_get(part_bytes, st)) != nullptr) {
      // direct ABI calls (tg_gemm_*): a per-thread, per-device scratch buffer that only grows; its counters are
      // zeroed when it is (re)allocated and left zero by every kernel; calls on another stream wait for the last user
      sk = cache->p;
      parts = cache->p + SkCache::cnt_cap;
    } else {
      int dev = 0;
      TG_CUDA(cudaGetDevice(&dev));
      tg_tune_mempool(dev);
      TG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&sk), cnt_bytes + part_bytes, st));
      own = true;
      parts = sk + cnt_bytes;
      cudaError_t e = cudaMemsetAsync(sk, 0, cnt_bytes, st);
      if (e != cudaSuccess) {
        cudaFreeAsync(sk, st);
        tg_set_error("stream-K counters: %s", cudaGetErrorString(e));
        return TG_ECUDA;
      }
    }
  }
  if (strm && strm->pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sched.G);
    cfg.blockDim = dim3(GAUSS ? GEMM_THREADS_GAUSS : GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const double hr = Headroom<F16>::value;
    float *pf = reinterpret_cast<float *>(parts);
    unsigned int *pc = reinterpret_cast<unsigned int *>(sk);
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_x3_kernel<F16, GAUSS>, ta, tb, tc, td, M, Np, K, out, ldo, accumulate,
                                        peak_key, hr, sep_guard, peers, sched, pf, pc, strm->cnt, strm->flag, strm->epoch);
    if (le != cudaSuccess) {
      tg_set_error("gemm_x3_kernel (dependent launch): %s", cudaGetErrorString(le));
      if (own) cudaFreeAsync(sk, st);
      if (cache) sk_cache_release(cache, st);
      return TG_ECUDA;
    }
  } else {
    gemm_x3_kernel<F16, GAUSS><<<(unsigned)sched.G, GAUSS ? GEMM_THREADS_GAUSS : GEMM_THREADS, smem, st>>>(
        ta, tb, tc, td, M, Np, K, out, ldo, accumulate, peak_key, Headroom<F16>::value, sep_guard, peers, sched,
        reinterpret_cast<float *>(parts), reinterpret_cast<unsigned int *>(sk), strm ? strm->cnt : nullptr,
        strm ? strm->flag : nullptr, strm ? strm->epoch : 0u);
  }
  rc = tg_launch_check(GAUSS ? "gemm_x3_kernel<f16, 3-product>" : F16 ? "gemm_x3_kernel<f16>" : "gemm_x3_kernel<tf32>");
  if (own) cudaFreeAsync(sk, st);
  if (cache) sk_cache_release(cache, st);
  return rc;
}

// grid of a kernel that walks `blocks` virtual blocks with a stride loop: enough CTAs to fill the machine, few enough
// that a launch cancelled by the device-side verdict costs microseconds
inline unsigned bounded_grid(long long blocks) {
  const long long cap = 148LL * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// explicit tensor method under stream capture: the verdict cannot be read on the host, so a non-separable
// input poisons the output instead of being silently ignored
__global__ void __launch_bounds__(256)
    nan_fill_kernel(double *__restrict__ out, size_t n, const unsigned long long *__restrict__ key) {
  if (tg_key_is_separable(*key)) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __longlong_as_double(0x7ff8000000000000LL);
}

// TG_GEMM_TILED=1 (experiment knob, default off): tiled operands on the dense 3-product path as well.  MEASURED on
// B200 (C2 graph replay, L2 flushed, two A/B pairs): 0.2717 ms tiled against 0.2696 ms row-major, same image -- unlike
// the tile-binned sum, whose operands are streamed once from HBM, the dense GEMM finds its operand tiles in L2 and the
// dense factor kernels were not limited by DRAM pages.  Validated (43 tensor-path tests) and left off.  Read once.
bool tiled_enabled() {
  static const bool on = [] {
    const char *e = getenv("TG_GEMM_TILED");
    return e && atoi(e) != 0;
  }();
  return on;
}
// Batches of kBatch beamlets (outer loop: the column factors of a batch are built once) x row blocks (inner
// loop: row factors + GEMM of one block; blocks exist for the host-buffer pipeline, whose D2H of a finished
// block overlaps the GEMMs of the following ones).  acc: fp64 (nrows x 2W) accumulator = the output itself
// for complex128.  A_hi / A_lo hold one row block.
template <bool F16, bool GAUSS>
int run_batches(int64_t nb, const double *table, int row0, int nrows, int W, long long ldk, void *Ahi, void *Alo,
                void *Bhi, void *Blo, double *acc, const unsigned long long *peak, const unsigned long long *guard,
                cudaStream_t st, const TgPeers &gemm_peers, int block_rows, const TgEmit *emit, void *out,
                int out_is_c128, const SkWs *skws, const SkStream *strm = nullptr, const SkStream *blk_sig = nullptr,
                size_t blk_cnt_stride = 0, size_t blk_part_stride = 0) {
  static_assert(!GAUSS || F16, "the 3-product formulation is implemented for fp16 x 3 operands");
  // blk_sig: one beamlet batch, the row factors of ALL rows built at once (A_hi / A_lo hold nrows rows), then one
  // GEMM launch per row block, back to back with programmatic dependent launch, each raising flag[block] when its
  // rows are in memory (the caller polls the flags and queues the D2H copies; no events between the launches)
  const int ldo = 2 * W;                     // doubles per output row (re, im interleaved)
  const int Np = GAUSS ? W : 2 * W;          // B rows: complex columns (3-product) or real columns
  // 3-product path, one launch per row block: optionally TILED operands (k-block by k-block, see SkSched::tiled_a and
  // tiled_enabled: measured without gain here)
  const bool tiled = GAUSS && !blk_sig && !strm && tiled_enabled();
  const int b_pad = tiled ? ((Np + BN - 1) / BN) * BN : 0;
  int rc = TG_OK;
  TgPeers none;
  none.n = 0;
  for (long long b0 = 0; b0 < nb && rc == TG_OK; b0 += kBatch) {
    const int nbatch = (int)((nb - b0) < kBatch ? (nb - b0) : kBatch);
    const int npad = ((nbatch + KCH - 1) / KCH) * KCH;
    const int K = GAUSS ? 3 * npad : 2 * nbatch;
    const bool last = b0 + kBatch >= nb;
    // (the tensor maps are encoded with the true K: the TMA unit zero-fills the K padding)
    if constexpr (GAUSS) {
      const unsigned gb = bounded_grid((long long)((npad / 2 + 127) / 128) * ((W + FS - 1) / FS));
      factor_gauss_kernel<false><<<gb, 128, 0, st>>>(table, b0, nbatch, npad, 0, W, W, ldk, static_cast<__half *>(Bhi),
                                                     static_cast<__half *>(Blo), peak, guard, b_pad);
    } else {
      const unsigned gb = bounded_grid((long long)((nbatch + 127) / 128) * ((W + FS - 1) / FS));
      factor_cols_kernel<F16><<<gb, 128, 0, st>>>(table, b0, nbatch, W, ldk, Bhi, Blo, guard);
    }
    rc = tg_launch_check("factor_cols_kernel");
    int blk = 0;
    if (blk_sig) {
      if constexpr (GAUSS) {
        const unsigned ga = bounded_grid((long long)((npad / 2 + 127) / 128) * ((nrows + FS - 1) / FS));
        factor_gauss_kernel<true><<<ga, 128, 0, st>>>(table, b0, nbatch, npad, row0, nrows, W, ldk,
                                                      static_cast<__half *>(Ahi), static_cast<__half *>(Alo), peak,
                                                      guard, 0);
      } else {
        const unsigned ga = bounded_grid((long long)((nbatch + 127) / 128) * ((nrows + FS - 1) / FS));
        factor_rows_kernel<F16><<<ga, 128, 0, st>>>(table, b0, nbatch, row0, nrows, W, ldk, Ahi, Alo, peak, guard);
      }
      rc = tg_launch_check("factor_rows_kernel");
      const size_t elem = GemmCfg<F16>::ELEM;
      for (int r = 0; r < nrows && rc == TG_OK; r += block_rows, ++blk) {
        const int nr = (nrows - r) < block_rows ? (nrows - r) : block_rows;
        SkStream sg = *blk_sig;
        sg.cnt += blk;
        sg.flag += blk;
        sg.pdl = blk > 0;
        SkWs w = *skws;
        if (w.p) {                         // per-block counters and partial tiles (the launches overlap)
          w.p += (size_t)blk * blk_cnt_stride;
          w.parts += (size_t)blk * blk_part_stride;
        }
        const unsigned char *a_hi = static_cast<const unsigned char *>(Ahi) + (size_t)r * ldk * elem;
        const unsigned char *a_lo = static_cast<const unsigned char *>(Alo) + (size_t)r * ldk * elem;
        rc = launch_gemm<F16, GAUSS>(a_hi, a_lo, Bhi, Blo, nr, Np, K, ldk, acc + (size_t)r * ldo, (long long)ldo, 0, peak,
                                     guard, st, none, &w, &sg);
      }
      return rc;
    }
    for (int r = 0; r < nrows && rc == TG_OK; r += block_rows, ++blk) {
      const int nr = (nrows - r) < block_rows ? (nrows - r) : block_rows;
      if constexpr (GAUSS) {
        const unsigned ga = bounded_grid((long long)((npad / 2 + 127) / 128) * ((nr + FS - 1) / FS));
        factor_gauss_kernel<true><<<ga, 128, 0, st>>>(table, b0, nbatch, npad, row0 + r, nr, W, ldk,
                                                      static_cast<__half *>(Ahi), static_cast<__half *>(Alo), peak,
                                                      guard, tiled ? ((nr + BM - 1) / BM) * BM : 0);
      } else {
        const unsigned ga = bounded_grid((long long)((nbatch + 127) / 128) * ((nr + FS - 1) / FS));
        factor_rows_kernel<F16><<<ga, 128, 0, st>>>(table, b0, nbatch, row0 + r, nr, W, ldk, Ahi, Alo, peak, guard);
      }
      rc = tg_launch_check("factor_rows_kernel");
      if (rc != TG_OK) break;
      TgPeers pe = none;
      if (last) {                                         // peers: final batch only
        pe = gemm_peers;
        for (int p = 0; p < pe.n; ++p) pe.ptr[p] = static_cast<double *>(pe.ptr[p]) + (size_t)r * ldo;
      }
      double *acc_r = acc + (size_t)r * ldo;
      rc = launch_gemm<F16, GAUSS>(Ahi, Alo, Bhi, Blo, nr, Np, K, ldk, acc_r, (long long)ldo, b0 > 0 ? 1 : 0, peak,
                                   guard, st, pe, skws, strm, tiled ? ((nr + BM - 1) / BM) * BM : 0, b_pad);
      if (rc == TG_OK && last && !out_is_c128) {
        const size_t n = (size_t)nr * ldo;
        TgPeers pc = none;   // complex64 peers are written by the conversion
        float *o = static_cast<float *>(out) + (size_t)r * ldo;
        f64_to_c64_kernel<<<bounded_grid((long long)((n + 255) / 256)), 256, 0, st>>>(acc_r, o, n, guard, pc);
        rc = tg_launch_check("f64_to_c64_kernel");
      }
      if (rc == TG_OK && last && emit) {
        const size_t row_bytes = (size_t)ldo * (out_is_c128 ? 8 : 4);
        rc = tg_emit_block(emit, blk, st, static_cast<unsigned char *>(out) + (size_t)r * row_bytes,
                           (size_t)r * row_bytes, (size_t)nr * row_bytes);
      }
    }
  }
  return rc;
}

// stream-K scratch of a whole call: the largest need over {full, last} row block x {full, last} beamlet batch
template <bool F16, bool GAUSS>
void sk_need_for_call(int64_t nb, int nrows, int block_rows, int W, int sms, size_t *cnt_cap, size_t *part_cap) {
  const int Np = GAUSS ? W : 2 * W;
  const long long last_b = nb % kBatch ? nb % kBatch : (nb < kBatch ? nb : kBatch);
  const long long full_b = nb < kBatch ? nb : kBatch;
  const int last_r = nrows % block_rows ? nrows % block_rows : block_rows;
  *cnt_cap = *part_cap = 0;
  for (long long bsz : {full_b, last_b})
    for (int nr : {block_rows, last_r}) {
      const int K = GAUSS ? 3 * (int)(((bsz + KCH - 1) / KCH) * KCH) : 2 * (int)bsz;
      size_t c = 0, p = 0;
      sk_scratch_need<F16, GAUSS>(nr, Np, K, sms, &c, &p);
      if (c > *cnt_cap) *cnt_cap = c;
      if (p > *part_cap) *part_cap = p;
    }
}

// Which formulation TG_METHOD_TENSOR / AUTO run.  The 3-product GEMM does 25 % less tensor work per image on tiles
// twice as large (128 rows x 128 complex columns), so it needs enough of them to keep the machine busy -- through whole
// tiles or the plain split-K schedule.  Measured on B200 (256-k chunks): C2 (64 such tiles, split in 2) 0.178 ms against
// 0.205 ms for the 4-multiplication form on 128 whole tiles; C3 on the tensor path 7.4 against 8.7 ms; 512 rows (32
// tiles, split in 4) 0.111 against 0.119 ms; on smaller row blocks and shards the 4-multiplication form wins (256 rows
// 0.080 vs 0.088 ms, 128 rows 0.059 vs 0.090 ms).  Hence: 3-product from 32 complex tiles on.  TG_TENSOR_GAUSS=0 / 1
// forces one form for every shape; TG_METHOD_TENSOR_3M / _4M select explicitly.
double sfu_wins_ratio() {                        // tuning knob: TG_SFU_WINS_BELOW overrides the constant
  static const double v = [] {
    const char *e = getenv("TG_SFU_WINS_BELOW");
    const double x = e ? atof(e) : kSfuWinsBelow;
    return (x >= 0.0 && x <= 1.0) ? x : kSfuWinsBelow;
  }();
  return v;
}
// TG_TENSOR_BINNED=0: AUTO never takes the tile-binned sum (sparse separable beamlets go to the culled SFU kernel as
// before) -- for A/B measurements.  Read on every call.
bool binned_enabled() {
  const char *e = getenv("TG_TENSOR_BINNED");
  return !(e && atoi(e) == 0);
}
bool use_gauss(int rows, int W) {
  static const int forced = [] {
    const char *e = getenv("TG_TENSOR_GAUSS");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  if (forced >= 0) return forced == 1;
  return (long long)((rows + BM - 1) / BM) * ((W + BN - 1) / BN) >= 32;
}

}  // namespace

// The work decomposition of the persistent GEMM for a given shape, as the kernel's three roles enumerate it
// (host-side mirror, no device needed): units[6 i..] = {cta, tile, chunk_begin, chunk_end, slot, nparts};
// sched_out[10] = SkSched fields.  Returns the number of units (written up to max_units), < 0 on error.
extern "C" int tg_gemm_schedule(int M, int N, int K, int f16, int sms, int mode, int32_t *units, int max_units,
                                int32_t *sched_out) {
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && sms > 0, "bad GEMM shape");
  const int rtm = mode >= 100 ? mode - 100 : 0;      // mode 100 + r: row-block streaming, rounds of r tile rows
  if (rtm) mode = -1;
  const SkSched s = f16 ? make_sched(M, N, K, GemmCfg<true>::BK, GemmCfg<true>::CHUNK_KB, sms, mode, rtm)
                        : make_sched(M, N, K, GemmCfg<false>::BK, GemmCfg<false>::CHUNK_KB, sms, mode, rtm);
  if (sched_out) {
    const int v[10] = {s.tiles_n, s.T, s.nkb, s.nch, s.G, s.R, s.q, s.Tl, s.qh, s.maxparts};
    for (int i = 0; i < 10; ++i) sched_out[i] = v[i];
  }
  int n = 0;
  for (int cta = 0; cta < s.G; ++cta) {
    SkIter it;
    it.init(s, cta);
    SkUnit u;
    while (it.next(s, cta, u)) {
      if (units && n < max_units) {
        int32_t *o = units + 6 * (size_t)n;
        o[0] = cta; o[1] = u.tile; o[2] = u.ch0; o[3] = u.ch1; o[4] = u.slot; o[5] = u.nparts;
      }
      ++n;
    }
  }
  return n;
}

// The RAGGED decomposition of the tile-binned sum (SkSched::bins), host-side mirror: chunks_per_tile[T] -> every unit
// as the kernel's roles enumerate it, units[8 i..] = {cta, tile, chunk_begin, chunk_end (global chunk indices), slot,
// nparts, scratch slot the unit parks its partial tile in, first CTA of the tile}; readers[] (may be NULL, sized by the
// caller as sum of nparts over the returned units) lists for every unit the scratch slots a last arriver would sum,
// part by part.  Returns the number of units, < 0 on error.
extern "C" int tg_gemm_schedule_ragged(int T, const int32_t *chunks_per_tile, int sms, int32_t *units, int max_units,
                                       int32_t *readers, int max_readers) {
  TG_REQUIRE(T > 0 && chunks_per_tile && sms > 0, "bad arguments");
  int *bins = static_cast<int *>(malloc(sizeof(int) * (size_t)(BIN_HDR + T + 1)));
  TG_REQUIRE(bins, "out of memory");
  long long tot = 0;
  for (int t = 0; t < T; ++t) {
    bins[BIN_HDR + t] = (int)tot;
    tot += chunks_per_tile[t] > 0 ? chunks_per_tile[t] : 0;
  }
  bins[BIN_HDR + T] = (int)tot;
  bins[BIN_CTOT] = (int)tot;
  const long long q = (tot + sms - 1) / sms;
  bins[BIN_Q] = (int)(q > 0 ? q : 1);
  bins[BIN_OVERFLOW] = 0;
  bins[BIN_NEED] = (int)tot;
  SkSched s = make_sched(BM, BN, CHUNK_K, GemmCfg<true>::BK, GemmCfg<true>::CHUNK_KB, sms, 0);
  s.T = T;
  s.G = sms;
  s.bins = bins;
  int n = 0, nr = 0;
  for (int cta = 0; cta < sms; ++cta) {
    SkIter it;
    it.init<true>(s, cta);
    SkUnit u;
    while (it.next<true>(s, cta, u)) {
      if (units && n < max_units) {
        int32_t *o = units + 8 * (size_t)n;
        o[0] = cta; o[1] = u.tile; o[2] = u.ch0; o[3] = u.ch1; o[4] = u.slot; o[5] = u.nparts;
        o[6] = 2 * cta + (u.ch0 == cta * u.rq ? 0 : 1);
        o[7] = u.first;
      }
      for (int part = 0; part < u.nparts; ++part, ++nr) {
        const int c = u.first + part;
        if (readers && nr < max_readers) readers[nr] = 2 * c + ((long long)c * u.rq >= (long long)u.p0 ? 0 : 1);
      }
      ++n;
    }
  }
  free(bins);
  return n;
}

#ifdef TG_GEMM_TRACE
extern "C" int tg_debug_gemm_trace(unsigned long long *out, int clear) {
  TG_REQUIRE(out, "null pointer");
  TG_CUDA(cudaDeviceSynchronize());
  TG_CUDA(cudaMemcpyFromSymbol(out, g_gemm_trace, sizeof(unsigned long long) * 148 * 8));
  if (clear) {
    static unsigned long long zero[148 * 8] = {};
    TG_CUDA(cudaMemcpyToSymbol(g_gemm_trace, zero, sizeof(zero)));
  }
  return TG_OK;
}
#endif

// D[M x N] = (A_hi + A_lo)[M x K] * (B_hi + B_lo)[N x K]^T on the tensor cores (3 x TF32), fp64 out.
extern "C" int tg_gemm_tf32x3(int M, int N, int K, const float *A_hi, const float *A_lo, const float *B_hi,
                              const float *B_lo, long long ldk, double *D, long long ldd, int accumulate,
                              void *stream) {
  TG_REQUIRE(M > 0 && N > 0 && K > 0, "bad GEMM shape");
  TG_REQUIRE(A_hi && A_lo && B_hi && B_lo && D, "null pointer");
  TG_REQUIRE(ldk >= K && (ldk % 4) == 0, "ldk must be >= K and a multiple of 4 (16-byte TMA pitch)");
  TG_REQUIRE(((uintptr_t)A_hi % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)B_hi % 16) == 0 &&
                 ((uintptr_t)B_lo % 16) == 0,
             "operands must be 16-byte aligned");
  TgPeers none;
  none.n = 0;
  return launch_gemm<false>(A_hi, A_lo, B_hi, B_lo, M, N, K, ldk, D, ldd, accumulate, nullptr, nullptr,
                            static_cast<cudaStream_t>(stream), none);
}
// the same with fp16 operands (IEEE binary16, 2 bytes each): 3 x kind::f16 at twice the TF32 rate
extern "C" int tg_gemm_f16x3(int M, int N, int K, const void *A_hi, const void *A_lo, const void *B_hi,
                             const void *B_lo, long long ldk, double *D, long long ldd, int accumulate,
                             void *stream) {
  TG_REQUIRE(M > 0 && N > 0 && K > 0, "bad GEMM shape");
  TG_REQUIRE(A_hi && A_lo && B_hi && B_lo && D, "null pointer");
  TG_REQUIRE(ldk >= K && (ldk % 8) == 0, "ldk must be >= K and a multiple of 8 (16-byte TMA pitch)");
  TG_REQUIRE(((uintptr_t)A_hi % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)B_hi % 16) == 0 &&
                 ((uintptr_t)B_lo % 16) == 0,
             "operands must be 16-byte aligned");
  TgPeers none;
  none.n = 0;
  return launch_gemm<true>(A_hi, A_lo, B_hi, B_lo, M, N, K, ldk, D, ldd, accumulate, nullptr, nullptr,
                           static_cast<cudaStream_t>(stream), none);
}

// Complex D[M x N] (+)= sum_n U[m, n] V[c, n] with three real products per term (see gemm_x3_kernel, GAUSS): the
// operands are in the 3-product layout (A'' / B'' of factor_gauss_kernel: per group of 128 terms the blocks
// Ur + Ui | Ur | Ui and Vr | Vi - Vr | Vr + Vi, fp16 hi and lo parts), K3 = 3 * 128 * ceil(nterms / 128) <= ldk.
// D: (M, N) complex128 as interleaved doubles, row pitch ldd doubles.  Exposed for testing.
extern "C" int tg_gemm_chunk_k(void) { return CHUNK_K; }

extern "C" int tg_cgemm3_f16x3(int M, int N, int K3, const void *A_hi, const void *A_lo, const void *B_hi,
                               const void *B_lo, long long ldk, double *D, long long ldd, int accumulate,
                               void *stream) {
  TG_REQUIRE(M > 0 && N > 0 && K3 > 0 && K3 % (3 * KCH) == 0, "bad shape (K3 must be a multiple of 3 * tg_gemm_chunk_k())");
  TG_REQUIRE(A_hi && A_lo && B_hi && B_lo && D, "null pointer");
  TG_REQUIRE(ldk >= K3 && (ldk % 8) == 0 && ldd >= 2LL * N, "bad pitch");
  TG_REQUIRE(((uintptr_t)A_hi % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)B_hi % 16) == 0 &&
                 ((uintptr_t)B_lo % 16) == 0,
             "operands must be 16-byte aligned");
  TgPeers none;
  none.n = 0;
  return launch_gemm<true, true>(A_hi, A_lo, B_hi, B_lo, M, N, K3, ldk, D, ldd, accumulate, nullptr, nullptr,
                                 static_cast<cudaStream_t>(stream), none);
}

extern "C" int tg_field_sum_separable(int64_t nb, const double *poly, const double px2m[6], int H, int W,
                                      int row0, int nrows, void *out, int out_is_c128, void *stream) {
  return tg_separable_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, nullptr,
                          static_cast<cudaStream_t>(stream), 0, 1);
}

int tg_separable_run(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                     int nrows, void *out, int out_is_c128, unsigned long long *key_async,
                     cudaStream_t stream, int cost_cull_bits, int f16, const TgPeers *peers, const TgEmit *emit,
                     int flags) {
  TG_REQUIRE(nb >= 0 && H > 0 && W > 0, "bad shape");
  TG_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  TG_REQUIRE(px2m && out, "null pointer");
  const bool verdict_only = (flags & TG_SEP_VERDICT_ONLY) != 0;
  TG_REQUIRE(!verdict_only || key_async, "verdict-only mode needs a device key");
  cudaStream_t st = stream;
  if (nrows == 0) return TG_OK;
  const size_t npix = (size_t)nrows * W;
  TgPeers none;
  none.n = 0;
  const TgPeers &pe = peers ? *peers : none;
  TG_REQUIRE(!(emit && pe.n > 0), "row-block emission and peer images are separate modes");
  if (nb == 0) {
    TG_REQUIRE(!verdict_only, "verdict of an empty beamlet set");
    TG_CUDA(cudaMemsetAsync(out, 0, npix * (out_is_c128 ? 16 : 8), st));
    for (int p = 0; p < pe.n; ++p) TG_CUDA(cudaMemsetAsync(pe.ptr[p], 0, npix * (out_is_c128 ? 16 : 8), st));
    if (emit) return tg_emit_block(emit, 0, st, out, 0, npix * (out_is_c128 ? 16 : 8));
    return TG_OK;
  }
  TG_REQUIRE(poly, "null poly");
  int dev = 0;
  TG_CUDA(cudaGetDevice(&dev));
  tg_tune_mempool(dev);

  // An explicit tensor method normally reads the separability verdict back (8 bytes + a stream synchronise)
  // to report TG_ENOTSEPARABLE; that is illegal while the stream is being captured into a CUDA graph.  Then
  // the verdict stays on the device like in AUTO mode: the kernels of this path return at once when the
  // beamlets are not separable and the output is filled with NaN instead.
  const bool capturing = !key_async && !(flags & TG_SEP_TRUSTED) && tg_stream_is_capturing(st);
  const bool host_check = !key_async && !(flags & TG_SEP_TRUSTED) && !capturing;

  int block_rows = nrows;
  // Variant (TG_E2E_STREAM=1; one beamlet batch, complex128): ONE GEMM launch that finishes the image block of rows by block of
  // rows (rounds of the split-K schedule) and raises a flag word in pinned host memory per block; this (synchronous)
  // call polls the flags and queues the D2H copy of each block as soon as it is complete, while the following rounds
  // run.  A GEMM launch per block (the fallback below) re-reads all column factors and pays ramp, tail and fix-up per
  // block: 4 x 0.088 ms for the four 256-row blocks of C2 against 0.18 ms for the whole image.  (cuStreamWaitValue32
  // on the copy stream instead of host polling was measured first: ~0.15 ms of latency per wait on this driver.)
  bool streaming = false;
  if (emit) {
    TG_REQUIRE(emit->block_rows > 0 && emit->block_rows % BM == 0 && emit->host_out && emit->ev, "bad emission block");
    block_rows = emit->block_rows < nrows ? emit->block_rows : nrows;
    streaming = out_is_c128 && nb <= kBatch && block_rows < nrows && !verdict_only && !key_async && stream_rounds_enabled() &&
                emit->flags_host && emit->flags_dev && (nrows + block_rows - 1) / block_rows <= emit->n_flags &&
                (nrows + block_rows - 1) / block_rows <= kMaxRounds && !tg_stream_is_capturing(st);
  }
  const int round_rows = block_rows;
  if (streaming) block_rows = nrows;
  // Variant (TG_E2E_FLAGGED=1; one beamlet batch, complex128): row factors of all rows at once, then one GEMM launch
  // per block of rows, chained with programmatic dependent launch so that a block's ramp (and its first CTAs, on the
  // SMs the previous block's split-K grid leaves idle) overlap the previous block's tail; each launch raises a flag
  // word in pinned host memory when its rows are in memory, this call polls the words and queues the D2H copies.
  // No event sits between the launches (an event record would serialise them).
  const int nblocks = (nrows + block_rows - 1) / block_rows;
  const bool flagged = emit && !streaming && out_is_c128 && nb <= kBatch && nblocks >= 2 && !verdict_only &&
                       !key_async && flagged_blocks_enabled() && emit->flags_host && emit->flags_dev &&
                       nblocks <= emit->n_flags && nblocks <= kMaxRounds && !tg_stream_is_capturing(st);
  const long long nbatch_max = nb < kBatch ? nb : kBatch;
  // f16: 0 = tf32 x 3, 1 = fp16 x 3 in the default formulation, 2 = fp16 x 3 4-multiplication, 3 = fp16 x 3 3-product
  const bool gauss = f16 == 3 || (f16 == 1 && use_gauss(block_rows, W));
  // operand row pitch: whole 128-byte k-blocks in either format; 3-product layout: 3 x 128 k per 128 beamlets
  const long long ldk = gauss ? 3 * KCH * ((nbatch_max + KCH - 1) / KCH) : ((2 * nbatch_max + 63) / 64) * 64;
  const size_t elem = f16 ? 2 : 4;
  const int Np = gauss ? W : 2 * W;          // B rows
  const size_t table_bytes = (((size_t)nb * 96 + 255) / 256) * 256;
  // (rows padded to whole 128-row tiles: the tiled layout of the 3-product path addresses k-blocks of whole tiles)
  auto pad128 = [](size_t r) { return (r + 127) / 128 * 128; };
  const size_t a_bytes = verdict_only ? 0 : ((pad128((size_t)(flagged ? nrows : block_rows)) * ldk * elem + 255) / 256) * 256;
  const size_t b_bytes = verdict_only ? 0 : ((pad128((size_t)Np) * ldk * elem + 255) / 256) * 256;
  const size_t acc_bytes = (out_is_c128 || verdict_only) ? 0 : npix * 16;
  size_t sk_cnt = 0, sk_part = 0;
  if (!verdict_only) {
    int sms = 148;
    int rcs = device_sms(&sms);
    if (rcs != TG_OK) return rcs;
    if (streaming) {
      const int K = gauss ? 3 * (int)(((nb + KCH - 1) / KCH) * KCH) : 2 * (int)nb, rtm = round_rows / BM;
      if (gauss) sk_scratch_need<true, true>(nrows, Np, K, sms, &sk_cnt, &sk_part, rtm);
      else if (f16) sk_scratch_need<true, false>(nrows, Np, K, sms, &sk_cnt, &sk_part, rtm);
      else sk_scratch_need<false, false>(nrows, Np, K, sms, &sk_cnt, &sk_part, rtm);
    } else if (gauss) sk_need_for_call<true, true>(nb, nrows, block_rows, W, sms, &sk_cnt, &sk_part);
    else if (f16) sk_need_for_call<true, false>(nb, nrows, block_rows, W, sms, &sk_cnt, &sk_part);
    else sk_need_for_call<false, false>(nb, nrows, block_rows, W, sms, &sk_cnt, &sk_part);
  }
  const size_t sig_bytes = (streaming || flagged) ? kMaxRounds * sizeof(unsigned int) : 0;     // arrival counters
  const size_t sk_cnt1 = sk_cnt, sk_part1 = sk_part;       // per launch
  if (flagged) {                                           // overlapping launches: scratch per block
    sk_cnt *= (size_t)nblocks;
    sk_part *= (size_t)nblocks;
  }
  TgAsyncBuf wsb(st);
  TG_CUDA(wsb.alloc(table_bytes + 256 + 2 * a_bytes + 2 * b_bytes + acc_bytes + sk_cnt + sk_part + sig_bytes));
  unsigned char *ws = wsb.as<unsigned char>();
  double *table = reinterpret_cast<double *>(ws);
  // control block after the table: key (8) | peak key (8) || gref (8) at +64 || est (8) at +128
  unsigned long long *key = key_async ? key_async : reinterpret_cast<unsigned long long *>(ws + table_bytes);
  unsigned long long *peak = reinterpret_cast<unsigned long long *>(ws + table_bytes + 8);
  const unsigned long long *guard = key_async ? key_async : (capturing ? key : nullptr);  // kernels decide on the device
  unsigned char *Ahi = ws + table_bytes + 256, *Alo = Ahi + a_bytes, *Bhi = Alo + a_bytes, *Blo = Bhi + b_bytes;
  double *acc = out_is_c128 ? static_cast<double *>(out) : reinterpret_cast<double *>(Blo + b_bytes);
  SkWs skws;
  if (sk_cnt) {
    skws.p = Blo + b_bytes + acc_bytes;
    skws.cnt_cap = sk_cnt1;
    skws.bytes = sk_cnt1 + sk_part1;
    skws.parts = skws.p + sk_cnt;
    TG_CUDA(cudaMemsetAsync(skws.p, 0, sk_cnt, st));   // arrival counters: zero once, the kernels leave them zero
  }
  SkStream strm;
  if (streaming || flagged) {
    strm.round_tiles_m = streaming ? round_rows / BM : 0;
    strm.cnt = reinterpret_cast<unsigned int *>(Blo + b_bytes + acc_bytes + sk_cnt + sk_part);
    strm.flag = emit->flags_dev;
    static std::atomic<unsigned int> epoch_counter{0};   // a value no earlier call has written into the flag words
    strm.epoch = ++epoch_counter;
    if (strm.epoch == 0) strm.epoch = ++epoch_counter;
    TG_CUDA(cudaMemsetAsync(strm.cnt, 0, sig_bytes, st));
  }
  TG_CUDA(cudaMemsetAsync(ws + table_bytes, 0, 16, st));  // own key slot, peak key
  if (key_async) TG_CUDA(cudaMemsetAsync(key, 0, (flags & TG_SEP_VERDICT_SPLIT) ? 16 : 8, st));
  // cost model (AUTO with culling enabled): brightest-peak key and tile estimate live after the key slot
  const bool cost = key_async != nullptr && cost_cull_bits > 0;
  unsigned long long *gref = reinterpret_cast<unsigned long long *>(ws + table_bytes + 64);
  double *est = reinterpret_cast<double *>(ws + table_bytes + 128);
  if (cost) {
    TG_CUDA(cudaMemsetAsync(gref, 0xFF, 8, st));
    TG_CUDA(cudaMemsetAsync(est, 0, 8, st));
  }
  // one O(nb) kernel: pixel-space table + separability verdict + pre-scaling peak (+ brightest-peak key)
  TgPrepExtra ex;
  ex.sep_key = key;
  ex.peak_key = peak;
  ex.row0 = row0;
  ex.nrows = nrows;
  int rc = tg_launch_prep(nb, poly, px2m, H, W, table, cost ? gref : nullptr, st, &ex);
  if (rc != TG_OK) return rc;
  if (cost) {
    sfu_cost_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(table, nb, H, W, gref, cost_cull_bits, est);
    const double sfu_wins_below = sfu_wins_ratio();
    verdict_kernel<<<1, 1, 0, st>>>(key, est, (double)nb * (double)H * (double)W, sfu_wins_below,
                                    (flags & TG_SEP_VERDICT_SPLIT) ? key + 1 : nullptr);
    rc = tg_launch_check("cost kernels");
    if (rc != TG_OK) return rc;
  }
  if (verdict_only) return TG_OK;
  if (host_check) {
    unsigned long long hkey = 0;
    TG_CUDA(cudaMemcpyAsync(&hkey, key, 8, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    double worst;
    memcpy(&worst, &hkey, 8);
    if (worst > 1.0) {
      tg_set_error("beamlets are not separable on this grid (cross term %.3g x tolerance)", worst);
      return TG_ENOTSEPARABLE;
    }
  }
  const TgEmit *emit_b = (streaming || flagged) ? nullptr : emit;          // the copies are queued below
  const SkStream *sp = streaming ? &strm : nullptr, *bs = flagged ? &strm : nullptr;
  rc = gauss ? run_batches<true, true>(nb, table, row0, nrows, W, ldk, Ahi, Alo, Bhi, Blo, acc, peak, guard, st,
                                       out_is_c128 ? pe : none, block_rows, emit_b, out, out_is_c128, &skws, sp, bs, sk_cnt1, sk_part1)
       : f16 ? run_batches<true, false>(nb, table, row0, nrows, W, ldk, Ahi, Alo, Bhi, Blo, acc, peak, guard, st,
                                        out_is_c128 ? pe : none, block_rows, emit_b, out, out_is_c128, &skws, sp, bs, sk_cnt1, sk_part1)
             : run_batches<false, false>(nb, table, row0, nrows, W, ldk, Ahi, Alo, Bhi, Blo, acc, peak, guard, st,
                                         out_is_c128 ? pe : none, block_rows, emit_b, out, out_is_c128, &skws, sp, bs, sk_cnt1, sk_part1);
  if (rc == TG_OK && (streaming || flagged)) {
    // per block: spin on the flag word (the kernel writes it after a system-scope fence), then queue the copy
    const size_t row_bytes = (size_t)2 * W * sizeof(double);
    volatile unsigned int *flags = emit->flags_host;
    int blk = 0;
    for (int r = 0; r < nrows && rc == TG_OK; r += round_rows, ++blk) {
      const int nr = (nrows - r) < round_rows ? (nrows - r) : round_rows;
      unsigned spins = 0;
      while (flags[blk] != strm.epoch) {
        if ((++spins & 0x3ffu) == 0) {                 // every 1024 polls: did the stream stop without the flag?
          const cudaError_t q = cudaStreamQuery(st);
          if (q == cudaErrorNotReady) continue;
          if (q != cudaSuccess) {
            tg_set_error("row-block streaming: %s", cudaGetErrorString(q));
            rc = TG_ECUDA;
          } else if (flags[blk] != strm.epoch) {
            tg_set_error("row-block streaming: the launch finished without completing block %d", blk);
            rc = TG_ECUDA;
          }
          break;
        }
      }
      if (rc != TG_OK) break;
      const cudaError_t e = cudaMemcpyAsync(emit->host_out + (size_t)r * row_bytes,
                                            static_cast<unsigned char *>(out) + (size_t)r * row_bytes,
                                            (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, emit->copy);
      if (e != cudaSuccess) {
        tg_set_error("row-block D2H: %s", cudaGetErrorString(e));
        rc = TG_ECUDA;
      }
    }
  }
  if (rc == TG_OK && !out_is_c128 && pe.n > 0) {
    // complex64 peer images: one more pass over the converted rows (the GEMM's peer stores are fp64-only)
    const size_t n = npix * 2;
    f64_to_c64_kernel<<<bounded_grid((long long)((n + 255) / 256)), 256, 0, st>>>(acc, static_cast<float *>(out), n, guard,
                                                                           pe);
    rc = tg_launch_check("f64_to_c64_kernel");
  }
  if (rc == TG_OK && capturing) {
    const size_t n = npix * (out_is_c128 ? 2 : 1);   // complex64: one NaN double covers (re, im)
    nan_fill_kernel<<<bounded_grid((long long)((n + 255) / 256)), 256, 0, st>>>(static_cast<double *>(out), n, key);
    rc = tg_launch_check("nan_fill_kernel");
  }
  return rc;
}

// ---- tile-binned tensor-core sum (see the kernels above) --------------------------------------------------------
namespace {
constexpr int kBinMaxChunks = 24576;        // operand capacity limit: 256 KiB per chunk -> 6 GiB
thread_local int g_bin_last_chunks = -1;    // operand chunks of this thread's last eager tile-binned call
struct BinCapKey {
  long long nb;
  int H, W, row0, nrows, cull, dev, need;
};
// chunks the last eager call of this thread needed, per shape: the capacity of a captured call
BinCapKey *bin_cap_slot(long long nb, int H, int W, int row0, int nrows, int cull, int dev, bool create) {
  static thread_local BinCapKey tab[16];
  static thread_local int used = 0, next = 0;
  for (int i = 0; i < used; ++i) {
    BinCapKey &k = tab[i];
    if (k.nb == nb && k.H == H && k.W == W && k.row0 == row0 && k.nrows == nrows && k.cull == cull && k.dev == dev) return &k;
  }
  if (!create) return nullptr;
  BinCapKey &k = tab[next];
  next = (next + 1) % 16;
  if (used < 16) ++used;
  k = BinCapKey{nb, H, W, row0, nrows, cull, dev, 0};
  return &k;
}
}  // namespace

int tg_separable_binned_run(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0, int nrows,
                            void *out, int out_is_c128, int cull_bits, cudaStream_t st, const TgPeers *peers,
                            const TgEmit *emit, int flags, int *probe_verdict) {
  if (probe_verdict) *probe_verdict = -1;
  TG_REQUIRE(nb >= 0 && H > 0 && W > 0, "bad shape");
  TG_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  TG_REQUIRE(px2m && out, "null pointer");
  TG_REQUIRE(cull_bits > 0, "the tile-binned sum needs a culling threshold (cull_bits > 0)");
  TG_REQUIRE(nb <= 0x7fffffffLL, "the tile-binned sum indexes beamlets with 32 bits");
  if (nrows == 0) return TG_OK;
  const size_t npix = (size_t)nrows * W;
  const size_t out_bytes = npix * (out_is_c128 ? 16 : 8);
  TgPeers none;
  none.n = 0;
  const TgPeers &pe = peers ? *peers : none;
  if (nb == 0) {
    TG_CUDA(cudaMemsetAsync(out, 0, out_bytes, st));
    for (int p = 0; p < pe.n; ++p) TG_CUDA(cudaMemsetAsync(pe.ptr[p], 0, out_bytes, st));
    if (emit) return tg_emit_block(emit, 0, st, out, 0, out_bytes);
    return TG_OK;
  }
  TG_REQUIRE(poly, "null poly");
  const int tiles_m = (nrows + BM - 1) / BM, tiles_n = (W + BIN_TN - 1) / BIN_TN;
  const long long Tl = (long long)tiles_m * tiles_n;
  if (tiles_m > 32767 || tiles_n > 32767 || Tl > (1 << 20)) {
    tg_set_error("tile-binned sum: %lld tiles are more than this path handles", Tl);
    return TG_EUNSUPPORTED;
  }
  const int T = (int)Tl;
  int dev = 0, sms = 148;
  TG_CUDA(cudaGetDevice(&dev));
  tg_tune_mempool(dev);
  int rc = device_sms(&sms);
  if (rc != TG_OK) return rc;
  const bool capturing = tg_stream_is_capturing(st);
  const bool probing = (flags & TG_BIN_PROBE) != 0;
  TG_REQUIRE(!probing || (!capturing && probe_verdict), "the probe reads its verdicts on the host: eager calls only");

  // ---- phase 1: table, verdicts, tile ranges, per-tile counts, prefix
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t table_bytes = al((size_t)nb * 96), wc_bytes = al((size_t)T * BIN_WARPS * sizeof(int)),
               bins_bytes = al((size_t)(BIN_HDR + T + 1) * sizeof(int));
  // one bit per (tile, beamlet): set by the beamlets (bin_mark_kernel), counted and expanded per tile
  const size_t hits_need = (size_t)T * (size_t)((nb + 31) / 32) * sizeof(unsigned);
  if (hits_need > ((size_t)1 << 30)) {
    tg_set_error("tile-binned sum: %d tiles x %lld beamlets are more than this path handles", T, (long long)nb);
    return TG_EUNSUPPORTED;
  }
  const size_t hits_bytes = al(hits_need);
  TgAsyncBuf ws1(st);
  TG_CUDA(ws1.alloc(table_bytes + 256 + wc_bytes + bins_bytes + hits_bytes));
  unsigned char *w1 = ws1.as<unsigned char>();
  double *table = reinterpret_cast<double *>(w1);
  unsigned long long *key = reinterpret_cast<unsigned long long *>(w1 + table_bytes);
  unsigned long long *peak = key + 1, *gref = reinterpret_cast<unsigned long long *>(w1 + table_bytes + 64);
  int *wc = reinterpret_cast<int *>(w1 + table_bytes + 256);
  int *bins = reinterpret_cast<int *>(w1 + table_bytes + 256 + wc_bytes);
  unsigned *hits = reinterpret_cast<unsigned *>(w1 + table_bytes + 256 + wc_bytes + bins_bytes);
  TG_CUDA(cudaMemsetAsync(hits, 0, hits_bytes, st));
  TG_CUDA(cudaMemsetAsync(key, 0, 16, st));
  TG_CUDA(cudaMemsetAsync(gref, 0xFF, 8, st));
  TgPrepExtra ex;
  ex.sep_key = key;
  ex.peak_key = peak;
  ex.row0 = row0;
  ex.nrows = nrows;
  rc = tg_launch_prep(nb, poly, px2m, H, W, table, gref, st, &ex);
  if (rc != TG_OK) return rc;
  const unsigned long long *probe = nullptr;
  if (probing) {
    // AUTO's cost verdict (see sfu_cost_kernel), kept apart from the separability key: ctl[BIN_CTL_SPLIT] = 1 if sparse
    double *est = reinterpret_cast<double *>(w1 + table_bytes + 128);
    TG_CUDA(cudaMemsetAsync(w1 + table_bytes + 8 * BIN_CTL_SPLIT, 0, 8, st));
    TG_CUDA(cudaMemsetAsync(est, 0, 8, st));
    TG_CUDA(cudaMemsetAsync(bins, 0, BIN_HDR * sizeof(int), st));
    sfu_cost_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(table, nb, H, W, gref, cull_bits, est);
    verdict_kernel<<<1, 1, 0, st>>>(key, est, (double)nb * (double)H * (double)W, sfu_wins_ratio(), key + BIN_CTL_SPLIT);
    probe = key;
  }
  bin_mark_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(nb, table, H, W, row0, nrows, tiles_n, gref, cull_bits, hits,
                                                                probe);
  bin_tiles_kernel<false><<<bounded_grid(T), 32 * BIN_WARPS, 0, st>>>(nb, T, wc, hits, nullptr, nullptr, nullptr, probe);
  rc = tg_launch_check("bin_tiles_kernel");
  if (rc != TG_OK) return rc;
  int cap = 0;
  if (!capturing) {
    bin_prefix_kernel<<<1, 1024, 0, st>>>(T, wc, bins, sms, 0x7fffffff, probe);
    rc = tg_launch_check("bin_prefix_kernel");
    if (rc != TG_OK) return rc;
    int hdr[BIN_HDR];
    unsigned long long hctl[BIN_CTL_SPLIT + 1];
    TG_CUDA(cudaMemcpyAsync(hdr, bins, sizeof(hdr), cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaMemcpyAsync(hctl, key, sizeof(hctl), cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    const unsigned long long hkey = hctl[0];
    if (probing) {
      if (!tg_key_is_separable(hkey)) {
        *probe_verdict = 0;
        return TG_NOT_BINNED;
      }
      if (hctl[BIN_CTL_SPLIT] == 0ULL) {
        *probe_verdict = 1;
        return TG_NOT_BINNED;
      }
      *probe_verdict = 2;
    }
    if (!(flags & TG_SEP_TRUSTED) && !tg_key_is_separable(hkey)) {
      double worst;
      memcpy(&worst, &hkey, 8);
      tg_set_error("beamlets are not separable on this grid (cross term %.3g x tolerance)", worst);
      return TG_ENOTSEPARABLE;
    }
    cap = hdr[BIN_NEED];
    g_bin_last_chunks = cap;
    if (BinCapKey *k = bin_cap_slot(nb, H, W, row0, nrows, cull_bits, dev, true)) k->need = cap;
  } else {
    const BinCapKey *k = bin_cap_slot(nb, H, W, row0, nrows, cull_bits, dev, false);
    if (!k) {
      tg_set_error("tile-binned sum under stream capture: run the same call once outside the capture first (the operand "
                   "capacity of the graph is taken from it)");
      return TG_EUNSUPPORTED;
    }
    const long long c = (long long)k->need + k->need / 4 + 64;
    cap = (int)(c < kBinMaxChunks ? c : kBinMaxChunks);
    bin_prefix_kernel<<<1, 1024, 0, st>>>(T, wc, bins, sms, cap, nullptr);
    rc = tg_launch_check("bin_prefix_kernel");
    if (rc != TG_OK) return rc;
  }
  if (cap > kBinMaxChunks) {
    tg_set_error("tile-binned sum: %d operand chunks exceed the capacity limit of %d (use the SFU kernel)", cap, kBinMaxChunks);
    return TG_EUNSUPPORTED;
  }
  if (cap == 0) {      // no beamlet reaches these rows (eager calls only: a captured call always has room for 64 chunks)
    TG_CUDA(cudaMemsetAsync(out, 0, out_bytes, st));
    for (int p = 0; p < pe.n; ++p) TG_CUDA(cudaMemsetAsync(pe.ptr[p], 0, out_bytes, st));
    if (emit) return tg_emit_block(emit, 0, st, out, 0, out_bytes);
    return TG_OK;
  }

  // ---- phase 2: slots, operands, ragged GEMM
  constexpr int NE = 8;
  const long long capK = (long long)cap * CHUNK_K;                       // k-elements of the concatenated axis
  const size_t c2t_bytes = al((size_t)cap * sizeof(int)), sel_bytes = al((size_t)cap * BIN_SLOTS * sizeof(int)),
               op_bytes = al((size_t)BM * capK * 2), cnt_bytes = al((size_t)T * NE * sizeof(unsigned int)),
               part_bytes = (size_t)2 * sms * (NE * 2048) * sizeof(float), acc_bytes = out_is_c128 ? 0 : npix * 16;
  TgAsyncBuf ws2(st);
  TG_CUDA(ws2.alloc(c2t_bytes + sel_bytes + 4 * op_bytes + cnt_bytes + part_bytes + acc_bytes));
  unsigned char *w2 = ws2.as<unsigned char>();
  int *c2t = reinterpret_cast<int *>(w2);
  int *sel = reinterpret_cast<int *>(w2 + c2t_bytes);
  unsigned char *Ahi = w2 + c2t_bytes + sel_bytes, *Alo = Ahi + op_bytes, *Bhi = Alo + op_bytes, *Blo = Bhi + op_bytes;
  unsigned int *counters = reinterpret_cast<unsigned int *>(Blo + op_bytes);
  float *parts = reinterpret_cast<float *>(Blo + op_bytes + cnt_bytes);
  double *acc = out_is_c128 ? static_cast<double *>(out) : reinterpret_cast<double *>(Blo + op_bytes + cnt_bytes + part_bytes);
  TG_CUDA(cudaMemsetAsync(counters, 0, cnt_bytes, st));
  const unsigned long long *guard = capturing ? key : nullptr;          // under capture the verdict stays on the device
  bin_tiles_kernel<true><<<bounded_grid(T), 32 * BIN_WARPS, 0, st>>>(nb, T, wc, hits, bins, sel, c2t, nullptr);
  factor_binned_kernel<true><<<bounded_grid(cap), BIN_SLOTS, 0, st>>>(table, bins, sel, c2t, tiles_n, row0, nrows, W, Ahi, Alo,
                                                                      Bhi, Blo, peak, guard);
  rc = tg_launch_check("binned factor kernels");
  if (rc != TG_OK) return rc;
  CUtensorMap ta, tb, tc, td;
  // tiled operands: tensors of 128-byte rows (64 fp16), k-block kb at rows [128 kb, 128 kb + 128)
  const long long op_rows = capK / GemmCfg<true>::BK * BM;
  if ((rc = make_map<true>(&ta, Ahi, op_rows, GemmCfg<true>::BK, GemmCfg<true>::BK)) != TG_OK) return rc;
  if ((rc = make_map<true>(&tb, Alo, op_rows, GemmCfg<true>::BK, GemmCfg<true>::BK)) != TG_OK) return rc;
  if ((rc = make_map<true>(&tc, Bhi, op_rows, GemmCfg<true>::BK, GemmCfg<true>::BK)) != TG_OK) return rc;
  if ((rc = make_map<true>(&td, Blo, op_rows, GemmCfg<true>::BK, GemmCfg<true>::BK)) != TG_OK) return rc;
  SkSched sc = make_sched(nrows, 2 * W, CHUNK_K, GemmCfg<true>::BK, GemmCfg<true>::CHUNK_KB, sms, 0);
  sc.T = T;
  sc.tiles_n = tiles_n;
  sc.nkb = (int)((long long)cap * GemmCfg<true>::CHUNK_KB);
  sc.G = sms;
  sc.bins = bins;
  // L2 prefetch distance of the operand stream in k-blocks of 64 KiB per CTA (TG_BIN_PREFETCH, experiment knob, default
  // off).  MEASURED on B200 (C3): cp.async.bulk.prefetch.tensor at distances 3 / 6 / 10 / 16 takes the GEMM from 0.58 to
  // 1.11 ms whatever the distance -- the prefetches queue in the same TMA unit as the loads they are meant to help
  static const int bin_pf = [] {
    const char *e = getenv("TG_BIN_PREFETCH");
    const int v = e ? atoi(e) : 0;
    return (v >= 0 && v <= 64) ? v : 0;
  }();
  sc.prefetch = bin_pf;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + sizeof(GemmSmemCtl) + 1024;
  TG_CUDA(cudaFuncSetAttribute(gemm_x3_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // peers: complex64 images are written by the conversion kernel, complex128 ones by peer_push_kernel below
  const TgPeers &gp = none;
  bin_zero_empty_kernel<<<bounded_grid(T), 256, 0, st>>>(bins, T, tiles_n, nrows, 2 * W, acc, (long long)(2 * W), gp, guard);
  gemm_x3_kernel<true, false, true><<<(unsigned)sms, GEMM_THREADS, smem, st>>>(
      ta, tb, tc, td, nrows, 2 * W, (int)(capK < 0x7fffffffLL ? capK : 0x7fffffffLL), acc, (long long)(2 * W), 0, peak,
      Headroom<true>::value, guard, gp, sc, parts, counters, nullptr, nullptr, 0u);
  rc = tg_launch_check("gemm_x3_kernel<f16> (tile-binned)");
  if (rc != TG_OK) return rc;
  if (out_is_c128 && pe.n > 0) {
    peer_push_kernel<<<bounded_grid((long long)((npix + 255) / 256)), 256, 0, st>>>(reinterpret_cast<const double2 *>(acc), npix,
                                                                                 pe, guard);
    rc = tg_launch_check("peer_push_kernel");
    if (rc != TG_OK) return rc;
  }
  if (!out_is_c128) {
    const size_t n = npix * 2;
    f64_to_c64_kernel<<<bounded_grid((long long)((n + 255) / 256)), 256, 0, st>>>(acc, static_cast<float *>(out), n, guard, pe);
    rc = tg_launch_check("f64_to_c64_kernel");
    if (rc != TG_OK) return rc;
  }
  if (capturing) {
    const size_t n = npix * (out_is_c128 ? 2 : 1);
    nan_fill_binned_kernel<<<bounded_grid((long long)((n + 255) / 256)), 256, 0, st>>>(static_cast<double *>(out), n, bins, key);
    rc = tg_launch_check("nan_fill_binned_kernel");
    if (rc != TG_OK) return rc;
  }
  if (emit) return tg_emit_block(emit, 0, st, out, 0, out_bytes);
  return TG_OK;
}

// Operand chunks (128 beamlet slots = 256 KiB of fp16 hi / lo row and column factors each) of the calling thread's last
// eager tile-binned sum, -1 if there was none: what bench.py turns into the path's HBM traffic.
extern "C" int tg_binned_last_chunks(void) { return g_bin_last_chunks; }

// The verdict TG_METHOD_AUTO reaches on the device, read back to the host (synchronises `stream`): 1 = the tensor-core
// path applies (separable, and not clearly more expensive than the culled SFU sum), 0 = SFU kernel.  Plans use it to
// freeze the dispatch at build time, so that the captured graph holds the launches of ONE path only.
extern "C" int tg_field_sum_verdict(int64_t nb, const double *poly, const double px2m[6], int H, int W, int cull_bits,
                                    int *use_tensor, void *stream) {
  TG_REQUIRE(use_tensor && px2m && H > 0 && W > 0 && nb >= 0, "bad arguments");
  *use_tensor = 0;
  if (nb == 0) return TG_OK;
  TG_REQUIRE(poly, "null poly");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TG_REQUIRE(!tg_stream_is_capturing(st), "the verdict is read back on the host: not inside a stream capture");
  TgAsyncBuf keyb(st);
  TG_CUDA(keyb.alloc(24));
  unsigned long long *key = keyb.as<unsigned long long>();
  // (the output pointer is only range-checked in verdict-only mode)
  int rc = tg_separable_run(nb, poly, px2m, H, W, 0, H, key + 2, 1, key, st, cull_bits, 1, nullptr, nullptr,
                            TG_SEP_VERDICT_ONLY | TG_SEP_VERDICT_SPLIT);
  if (rc != TG_OK) return rc;
  unsigned long long hkey[2] = {0, 0};
  TG_CUDA(cudaMemcpyAsync(hkey, key, 16, cudaMemcpyDeviceToHost, st));
  TG_CUDA(cudaStreamSynchronize(st));
  // 1 = dense GEMM, 2 = separable and sparse: tile-binned GEMM, 0 = SFU kernel
  *use_tensor = !tg_key_is_separable(hkey[0]) ? 0 : hkey[1] == 0 ? 1 : (cull_bits > 0 && binned_enabled()) ? 2 : 0;
  return TG_OK;
}

// method dispatch: AUTO enqueues BOTH paths with a device-side separability verdict (no host sync)
extern "C" int tg_field_sum(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                            int nrows, void *out, int out_is_c128, int cull_bits, int method, void *stream) {
  return tg_field_sum_impl(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, method,
                           static_cast<cudaStream_t>(stream), nullptr);
}
int tg_field_sum_impl(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0, int nrows,
                      void *out, int out_is_c128, int cull_bits, int method, cudaStream_t st, const TgPeers *peers,
                      const TgEmit *emit) {
  if (method == TG_METHOD_SFU)
    return tg_field_grid_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, nullptr, nullptr, st,
                             peers, emit);
  if (method == TG_METHOD_TENSOR || method == TG_METHOD_TENSOR_TF32 || method == TG_METHOD_TENSOR_4M ||
      method == TG_METHOD_TENSOR_3M)
    return tg_separable_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, nullptr, st, 0,
                            method == TG_METHOD_TENSOR ? 1 : method == TG_METHOD_TENSOR_4M ? 2
                            : method == TG_METHOD_TENSOR_3M ? 3 : 0, peers, emit);
  if (method == TG_METHOD_TENSOR_BINNED)
    return tg_separable_binned_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, st, peers, emit);
  TG_REQUIRE(method == TG_METHOD_AUTO, "unknown method");
  if (nb == 0 || nrows == 0)
    return tg_field_grid_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, nullptr, nullptr, st,
                             peers, emit);
  TgAsyncBuf keyb(st);
  TG_CUDA(keyb.alloc(16));
  unsigned long long *key = keyb.as<unsigned long long>();
  // Large problems outside a graph capture: every kernel of the path that does not apply still has to be launched
  // and exit (7 batches x {two factor grids of ~10^4 CTAs, GEMM} at C3: ~0.3 ms of dead launches, measured as the
  // gap between `auto` and the explicit method), which costs far more than reading 8 bytes back once.
  const bool host_verdict = emit != nullptr || (nb > kBatch && !tg_stream_is_capturing(st));
  if (host_verdict && cull_bits > 0 && binned_enabled()) {
    // one read-back for both verdicts and, when they say "separable and sparse", the operand count of the tile-binned
    // sum, which then simply goes on (TG_BIN_PROBE); the SFU kernel remains the fallback when its operands would not fit
    int verdict = -1;
    int rc = tg_separable_binned_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, st, peers, emit,
                                     TG_BIN_PROBE, &verdict);
    if (rc != TG_NOT_BINNED && !(rc == TG_EUNSUPPORTED && verdict != 1)) return rc;     // done, or a real error
    if (rc == TG_NOT_BINNED && verdict == 1)
      return tg_separable_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, nullptr, st, 0, 1, peers, emit,
                              TG_SEP_TRUSTED);
    if (rc == TG_NOT_BINNED || verdict == 2)             // not separable, or sparse but over the capacity limit
      return tg_field_grid_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, nullptr, nullptr, st,
                               peers, emit);
    // (TG_EUNSUPPORTED before any verdict: a shape the binned path does not take -- the two-step flow below decides)
  }
  if (host_verdict) {
    // (host-buffer pipeline: the call is synchronous anyway; only the path that applies is enqueued -- block by
    // block, each block's D2H behind its own event)
    int rc = tg_separable_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, key, st, cull_bits, 1, nullptr,
                              nullptr, TG_SEP_VERDICT_ONLY | TG_SEP_VERDICT_SPLIT);
    if (rc != TG_OK) return rc;
    unsigned long long hkey[2] = {0, 0};
    TG_CUDA(cudaMemcpyAsync(hkey, key, 16, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    if (tg_key_is_separable(hkey[0]) && hkey[1] == 0)
      return tg_separable_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, nullptr, st, 0, 1, peers, emit,
                              TG_SEP_TRUSTED);
    if (tg_key_is_separable(hkey[0]) && cull_bits > 0 && binned_enabled()) {
      // separable AND sparse (every beamlet reaches a small part of the detector): the tile-binned tensor-core sum;
      // the SFU kernel remains the fallback when the operands of the (beamlet, tile) pairs would not fit
      rc = tg_separable_binned_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, st, peers, emit,
                                   TG_SEP_TRUSTED);
      if (rc != TG_EUNSUPPORTED) return rc;
    }
    return tg_field_grid_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, nullptr, nullptr, st,
                             peers, emit);
  }
  int rc = tg_separable_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, key, st, cull_bits, 1, peers);
  if (rc == TG_OK)
    rc = tg_field_grid_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, nullptr, key, st, peers);
  return rc;
}
