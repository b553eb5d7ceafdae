// K3: the Gaussian-beamlet field sum  F[p] = sum_n exp(i P_n(p))  on sm_100a.
//
// Replaces map_reduce(_beam_field_outer, jnp.add, ...) of the reference
// (src/temgym_core/gaussian.py:319-369) -- every complex Gaussian on every pixel.
//
// Design (SFU / issue bound; 2 MUFU per beamlet*pixel: ex2 + sin/cos seeds of a phasor recurrence,
// instead of the naive 3):
//  * prep kernel: metre-space complex quadratic (K2 output) -> pixel-space table, phase in
//    TURNS (Re P / 2pi) and envelope in BITS (-Im P * log2 e), fp64, 96 B per beamlet.
//  * main kernel: a CTA owns a TR x TC pixel tile and streams the beamlet table through
//    shared memory in chunks, double-buffered with 1-D bulk async copies issued to the TMA
//    unit (cp.async.bulk ... mbarrier::complete_tx).  One thread per beamlet re-centres the
//    chunk's polynomials on the tile origin in fp64, reduces the phase coefficients mod 1
//    (they only ever multiply integers), optionally culls beamlets whose envelope over the
//    whole tile is negligible, and writes compact tile-local records.
//  * every thread owns a strip of L consecutive pixels of one row.  Per (thread, beamlet) it
//    evaluates the strip-start phase and first difference in fp64 and converts them to
//    32-bit FIXED-POINT turns; along the strip the phase advances by exact integer second
//    differences (wrap-around mod 1 turn is free).  Every 4 pixels the top 23 bits become an fp32
//    angle in [-pi, pi) for MUFU.SIN/COS -- exact seeds of the unit phasor z and of its step w;
//    in between z_{j+1} = z_j w_j, w_{j+1} = w_j c (c = exp(i dd), per beamlet, fp64 sincospi).
//    The envelope is a 2-FFMA Horner in fp32 for MUFU.EX2 on every pixel.
//  * fp32 partial sums over one chunk (<= 128 terms) are flushed into fp64 accumulators held
//    in shared memory, so accumulation error does not grow with the number of beamlets.
//  * grid = tiles x beamlet-splits, split count chosen so the CTA count fills whole waves
//    of 148 SMs x 2 resident CTAs; split partials are reduced by a second tiny kernel in a
//    fixed order (deterministic, no atomics on the data path).
#include <math.h>
#include "tg_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 128;          // beamlets per staged chunk
constexpr int kScanPer = 8;          // gather mode: bounding boxes tested per thread and scan round
constexpr int kCandRing = 4096;      // gather mode: capacity of the candidate ring (>= kChunk + kScanPer * kThreads)
constexpr int kRecDoubles = 18;      // tile-local record: 6 phase + 6 envelope + vertex(2) + {dd,e2} + {cr,ci}
constexpr double kMagic = 1572864.0; // 1.5 * 2^20: ulp = 2^-32 -> low mantissa word = frac * 2^32
constexpr double kInv2Pi = 0.15915494309189533577;
constexpr double kLog2e = 1.4426950408889634074;

struct FieldGeom {
  int H, W;          // full detector
  int row0, nrows;   // rows computed by this call
  int tiles_x, tiles_y;
  int nsplit;
  long long nb;
  int cull_bits;
  int via_partial;   // tiles go to the split-partials buffer (nsplit > 1, or peer destinations)
  // Cyclic tile rows (row-sharded multi-GPU sum over peer images): tile row ty of this call is tile row
  // ty_phase + ty * ty_stride of the detector (row0 = 0, nrows = H then); ty_stride = 0: contiguous rows.
  int ty_stride, ty_phase;
  int lrows;         // rows of this call's split-partials buffer (= nrows, or tiles_y * tile height when cyclic)
};

// ---- ordered-uint encoding of doubles for atomicMin ---------------------------------
__device__ __forceinline__ unsigned long long enc_ordered(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dec_ordered(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}

// minimum of a convex quadratic g(u,v) = G0 + G1 v + G2 u + G3 v^2 + G4 u v + G5 u^2 over the
// rectangle v in [0,V], u in [0,U].  Returns -inf when g is not convex (never cull then).
__device__ __forceinline__ double min1d(double a, double b, double c, double S) {
  // min over s in [0,S] of a + b s + c s^2
  double s = 0.0;
  if (c > 0.0) {
    s = fmin(fmax(-b / (2.0 * c), 0.0), S);
  } else {
    s = (b * S + c * S * S < 0.0) ? S : 0.0;
  }
  return a + s * (b + c * s);
}
__device__ __forceinline__ double quad_min_rect(const double *G, double U, double V) {
  const double G0 = G[0], G1 = G[1], G2 = G[2], G3 = G[3], G4 = G[4], G5 = G[5];
  const double det = 4.0 * G3 * G5 - G4 * G4;
  if (!(G3 >= 0.0 && G5 >= 0.0 && det >= 0.0)) return -INFINITY;
  if (det > 0.0) {
    const double vs = (-2.0 * G5 * G1 + G4 * G2) / det;
    const double us = (-2.0 * G3 * G2 + G4 * G1) / det;
    if (vs >= 0.0 && vs <= V && us >= 0.0 && us <= U) return G0 + 0.5 * (G1 * vs + G2 * us);
  }
  double m = min1d(G0, G2, G5, U);                                   // v = 0
  m = fmin(m, min1d(G0 + V * (G1 + G3 * V), G2 + G4 * V, G5, U));    // v = V
  m = fmin(m, min1d(G0, G1, G3, V));                                 // u = 0
  m = fmin(m, min1d(G0 + U * (G2 + G5 * U), G1 + G4 * U, G3, V));    // u = U
  return m;
}

// ---- prep: metre-space poly -> pixel-space table ------------------------------------
// poly: c0..c5 complex (radians) in (x,y) metres.  x = X0 + Xc col + Xr row, y likewise.
// table[n] = {T0..T5 (turns), E0..E5 (bits, = -Im * log2e)} over (col,row):
//   a0 + a1 col + a2 row + a3 col^2 + a4 col row + a5 row^2
__global__ void __launch_bounds__(128)
    prep_kernel(long long nb, const double *__restrict__ poly, double X0, double Xc, double Xr,
                double Y0, double Yc, double Yr, int H, int W, double *__restrict__ table,
                unsigned long long *__restrict__ gref_key, const TgPrepExtra ex) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double cross = 0.0, peak = -INFINITY;
  unsigned long long gmin_key = ~0ULL;          // ordered key of this beamlet's smallest on-detector exponent
  if (i < nb) {
    const double *c = poly + i * 12;
    double a[2][6];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const double c0 = c[0 + p], c1 = c[2 + p], c2 = c[4 + p], c3 = c[6 + p], c4 = c[8 + p],
                   c5 = c[10 + p];
      a[p][3] = c3 * Xc * Xc + c4 * Xc * Yc + c5 * Yc * Yc;
      a[p][5] = c3 * Xr * Xr + c4 * Xr * Yr + c5 * Yr * Yr;
      a[p][4] = 2.0 * c3 * Xc * Xr + c4 * (Xc * Yr + Xr * Yc) + 2.0 * c5 * Yc * Yr;
      a[p][1] = c1 * Xc + c2 * Yc + 2.0 * c3 * X0 * Xc + c4 * (X0 * Yc + Y0 * Xc) + 2.0 * c5 * Y0 * Yc;
      a[p][2] = c1 * Xr + c2 * Yr + 2.0 * c3 * X0 * Xr + c4 * (X0 * Yr + Y0 * Xr) + 2.0 * c5 * Y0 * Yr;
      a[p][0] = c0 + c1 * X0 + c2 * Y0 + c3 * X0 * X0 + c4 * X0 * Y0 + c5 * Y0 * Y0;
    }
    double *t = table + i * 12;
    double G[6], T4 = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double tj = a[0][j] * kInv2Pi;
      t[j] = tj;
      if (j == 4) T4 = tj;
      G[j] = a[1][j] * kLog2e;   // envelope exponent in bits: |field| = 2^-g
      t[6 + j] = -G[j];
    }
    if (gref_key) {
      const double gm = quad_min_rect(G, (double)(H - 1), (double)(W - 1));
      if (isfinite(gm)) gmin_key = enc_ordered(gm);
    }
    if (ex.sep_key) {
      // cross-term contribution across the detector, in units of the separability tolerances
      // (2^-24 turn of phase, 2^-20 bit of envelope); NaN beamlets are carried by the factors themselves
      const double hw = (double)H * (double)W;
      cross = fmax(fabs(T4) * hw * 16777216.0, fabs(G[4]) * hw * 1048576.0);
      if (!(cross == cross)) cross = 0.0;
    }
    if (ex.peak_key) {
      // peak over the call's rows of log2 |U_n(row)| = (E0 + max_col part) + E2 r + E5 r^2, E = -G
      const double q0 = -G[0] + tg_col_env_max(-G[1], -G[3], (double)(W - 1)), q1 = -G[2], q2 = -G[5];
      const double s_lo = (double)ex.row0, s_hi = (double)(ex.row0 + ex.nrows - 1);
      peak = fmax(q0 + s_lo * (q1 + q2 * s_lo), q0 + s_hi * (q1 + q2 * s_hi));
      if (q2 < 0.0) {
        const double sv = fmin(fmax(-0.5 * q1 / q2, s_lo), s_hi);
        peak = fmax(peak, q0 + sv * (q1 + q2 * sv));
      }
      if (!isfinite(peak)) peak = -INFINITY;   // NaN / inf beamlets do not set the scale
    }
  }
  // One atomic per warp and key, and only when it can still change the key (a plain read first: the keys move
  // monotonically, so a stale value only costs a redundant atomic).  The per-THREAD atomicMin this replaces
  // serialised 1e5 atomics on one address: 70 us of the 1e5-beamlet prep (ncu launch list, round 2).
  if (gref_key) {                              // block-uniform: every lane of every warp gets here
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, gmin_key, o);
      gmin_key = other < gmin_key ? other : gmin_key;
    }
    if ((threadIdx.x & 31) == 0 && gmin_key != ~0ULL && gmin_key < *reinterpret_cast<volatile unsigned long long *>(gref_key))
      atomicMin(gref_key, gmin_key);
  }
  if (ex.sep_key || ex.peak_key) {
    for (int o = 16; o > 0; o >>= 1) {
      cross = fmax(cross, __shfl_xor_sync(0xffffffffu, cross, o));
      peak = fmax(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (ex.sep_key && cross > 0.0) {
        const unsigned long long k = (unsigned long long)__double_as_longlong(cross);
        if (k > *reinterpret_cast<volatile unsigned long long *>(ex.sep_key)) atomicMax(ex.sep_key, k);
      }
      if (ex.peak_key && peak > -INFINITY) {
        const unsigned long long k = tg_enc_ordered(peak);
        if (k > *reinterpret_cast<volatile unsigned long long *>(ex.peak_key)) atomicMax(ex.peak_key, k);
      }
    }
  }
}

// Detector-space bounding box (whole pixels, clipped to the detector, +-1 px slack) of the region where a
// beamlet's envelope exceeds the culling threshold  E >= (brightest on-detector peak) - cull_bits.
// Non-concave envelopes get the whole detector; beamlets below the threshold everywhere an empty box.
__global__ void __launch_bounds__(256)
    bbox_kernel(long long nb, const double *__restrict__ table, int H, int W,
                const unsigned long long *__restrict__ gref_key, int cull_bits, short4 *__restrict__ bbox) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  short4 bb = make_short4(0, (short)min(W - 1, 32767), 0, (short)min(H - 1, 32767));
  const unsigned long long k = *gref_key;
  const double *e = table + i * 12 + 6;   // E(c, r) = e0 + e1 c + e2 r + e3 c^2 + e4 c r + e5 r^2 [bits]
  const double det = e[3] * e[5] - 0.25 * e[4] * e[4];
  if (k != ~0ULL && e[3] < 0.0 && e[5] < 0.0 && det > 0.0) {
    const double e_thr = -(dec_ordered(k) + (double)cull_bits);
    const double cs = (0.5 * e[4] * e[2] - e[5] * e[1]) / (2.0 * det);
    const double rs = (0.5 * e[4] * e[1] - e[3] * e[2]) / (2.0 * det);
    const double d = (e[0] + 0.5 * (e[1] * cs + e[2] * rs)) - e_thr;
    if (d < 0.0) {
      bb = make_short4(1, 0, 1, 0);        // empty
    } else if (isfinite(d) && isfinite(cs) && isfinite(rs)) {
      const double hc = sqrt(d * (-e[5]) / det) + 1.0, hr = sqrt(d * (-e[3]) / det) + 1.0;
      const double c_lo = floor(cs - hc), c_hi = ceil(cs + hc), r_lo = floor(rs - hr), r_hi = ceil(rs + hr);
      if (c_hi < 0.0 || r_hi < 0.0 || c_lo > (double)(W - 1) || r_lo > (double)(H - 1)) {
        bb = make_short4(1, 0, 1, 0);
      } else {
        bb.x = (short)fmin(fmax(c_lo, 0.0), 32767.0);
        bb.y = (short)fmin(fmin(c_hi, (double)(W - 1)), 32767.0);
        bb.z = (short)fmin(fmax(r_lo, 0.0), 32767.0);
        bb.w = (short)fmin(fmin(r_hi, (double)(H - 1)), 32767.0);
      }
    }
  }
  bbox[i] = bb;
}

// ---- main tiled kernel ----------------------------------------------------------------
struct __align__(16) Rec {
  double th[6];   // tile-local phase coefficients, turns, reduced to [-0.5, 0.5]
  double en[6];   // tile-local envelope coefficients (bits, amplitude = 2^en(u,v))
  float vs0, vs1; // column of the envelope's vertex along row u: vs0 + vs1 * u (0,0 if none); only picks the pivot pixel
  uint32_t dd;    // second difference of the phase along a row, fixed point 2^-32 turn
  float e2;       // en[3] as fp32
  float cr, ci;   // exp(i 2 pi dd): the phasor of the second difference (fp64 sincospi, rounded once)
  uint32_t cspan; // tile columns that can hold a contribution above the culling threshold: lo | hi << 16
  float rho;      // 2^(2 e2) when the envelope is SMOOTH on this tile (|d envelope / d column| <= 16 bits per pixel
                  // everywhere): the amplitude then follows a ratio recurrence like the phasor; 0 = use ex2 per pixel
  float c4r, c4i; // (rho c)^4: the ratio R of the smooth path advances by it from one 4-pixel group to the next
  float pad2[2];
};
static_assert(sizeof(Rec) == kRecDoubles * 8, "record size");

template <int L, int SPR>
struct FieldSmem {
  static constexpr int TC = L * SPR;
  static constexpr int TR = kThreads / SPR;
  alignas(128) double raw[2][kChunk * 12];
  alignas(16) Rec rec[kChunk];
  alignas(16) double acc[2 * L * kThreads];  // [component][j][thread]
  alignas(8) uint64_t full[2];
  int warp_cnt[kChunk / 32];
  int n_active;
  // gather mode: the candidate list (beamlets, relative to the split, whose bounding box meets the tile) is a RING
  // of kCandRing ints living in `raw` (the TMA staging buffers are not used in gather mode); gcnt holds the
  // per-(sub-round, warp) hit counts of one scan round
  int gcnt[kScanPer][kThreads / 32];
};

// WCULL: culling is enabled -> lane = row / warp = strip mapping with the per-warp column skip; the
// dense instantiation keeps lane = strip (8 strips x 4 rows per warp), which measures ~7 % faster.
template <int L, int SPR, bool WCULL>
__global__ void __launch_bounds__(kThreads, 2)
    field_grid_kernel(const double *__restrict__ table, const FieldGeom g,
                      const unsigned long long *__restrict__ gref_key, const short4 *__restrict__ bbox,
                      void *__restrict__ out, int out_is_c128, double2 *__restrict__ partial,
                      unsigned long long *__restrict__ evals,
                      const unsigned long long *__restrict__ sep_guard) {
  if (sep_guard && tg_key_is_separable(*sep_guard)) return;  // the tensor-core path owns this call
  using S = FieldSmem<L, SPR>;
  constexpr int TC = S::TC, TR = S::TR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  S &sm = *reinterpret_cast<S *>(smem_raw);

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int split = blockIdx.y;
  const int ty = tile / g.tiles_x, tx = tile % g.tiles_x;
  // tile origin in full-detector pixel coordinates
  const int r0 = g.ty_stride ? (g.ty_phase + ty * g.ty_stride) * TR : g.row0 + ty * TR;
  const int c0 = tx * TC;
  // lane = row inside the tile, warp = strip of L columns: a warp owns a TR x L pixel block, so a
  // beamlet can be skipped per warp (no divergence) when it cannot reach those columns -- culling at
  // 32 x 16 instead of 32 x 128 pixels for narrow beamlets (BASELINE C3)
  static_assert(S::TR == 32 && kThreads / 32 == SPR, "lane <-> row mapping");
  const int u = WCULL ? (tid & 31) : tid / SPR;                 // row inside tile
  const int v0 = (WCULL ? (tid >> 5) : (tid % SPR)) * L;        // first column of this thread's strip inside the tile

  // beamlet range of this split
  const long long per = (g.nb + g.nsplit - 1) / g.nsplit;
  const long long b_begin = (long long)split * per;
  const long long b_end = b_begin + per < g.nb ? b_begin + per : g.nb;
  const int nchunks = b_end > b_begin ? (int)((b_end - b_begin + kChunk - 1) / kChunk) : 0;

  if (tid == 0) {
    tg_mbar_init(&sm.full[0], 1);
    tg_mbar_init(&sm.full[1], 1);
    tg_fence_mbar_init();
  }
#pragma unroll
  for (int j = 0; j < 2 * L; ++j) sm.acc[j * kThreads + tid] = 0.0;
  __syncthreads();

  auto issue = [&](int c) {
    const long long b = b_begin + (long long)c * kChunk;
    const int cnt = (int)((b_end - b) < kChunk ? (b_end - b) : kChunk);
    const uint32_t bytes = (uint32_t)cnt * 96u;
    tg_mbar_expect_tx(&sm.full[c & 1], bytes);
    tg_bulk_g2s(sm.raw[c & 1], table + b * 12, bytes, &sm.full[c & 1]);
  };
  double thr_bits = INFINITY;  // cull when min envelope exponent g over tile > thr_bits
  if (g.cull_bits > 0 && gref_key) {
    const unsigned long long key = *gref_key;
    if (key != ~0ULL) thr_bits = dec_ordered(key) + (double)g.cull_bits;
  }
  // GATHER mode (culling with bounding boxes): instead of streaming the whole table through shared memory
  // and evaluating the few survivors of every 128-beamlet chunk (one barrier-bound evaluation phase per
  // chunk: ncu showed 2.4 barrier stalls per issue and the XU pipe at 41 % on BASELINE C3), all 256 threads
  // scan the 8-byte bounding boxes, survivors are appended IN ORDER to a candidate list, and an evaluation
  // phase starts only when 128 candidates are waiting (or the input is exhausted); their table rows are read
  // straight from L2.  Same beamlet order per pixel, so the sums keep their summation order.
  const bool gather = WCULL && bbox != nullptr && thr_bits < INFINITY;
  if (tid == 0 && nchunks > 0 && !gather) issue(0);

  // per-thread strip constants (exact small integers in fp64)
  const double ud = (double)u, vd = (double)v0;
  const float uf = (float)u, vf = (float)v0;


  float pr[L], pi[L];
#pragma unroll
  for (int j = 0; j < L; ++j) { pr[j] = 0.f; pi[j] = 0.f; }
  unsigned long long my_active = 0;
  int pending = 0;                   // terms in the fp32 partials since the last flush (<= 1.5 kChunk)

  int ncand = 0, cand_head = 0;      // gather mode: candidates waiting in the ring / ring position of the first
  long long bpos = b_begin;          // gather mode: next beamlet to scan
  int *cand = reinterpret_cast<int *>(sm.raw);
  static_assert(sizeof(sm.raw) >= kCandRing * sizeof(int) && kCandRing >= kChunk + kScanPer * kThreads &&
                    (kCandRing & (kCandRing - 1)) == 0,
                "candidate ring");
  for (int c = 0;; ++c) {
    int cnt;
    bool near;
    const double *a = nullptr;
    if (!gather) {
      if (c >= nchunks) break;
      if (tid == 0 && c + 1 < nchunks) issue(c + 1);
      tg_mbar_wait(&sm.full[c & 1], (uint32_t)((c >> 1) & 1));
      const long long b = b_begin + (long long)c * kChunk;
      cnt = (int)((b_end - b) < kChunk ? (b_end - b) : kChunk);
      near = tid < cnt;
      if (near && bbox && thr_bits < INFINITY) {
        // cheap reject: the beamlet's detector-space bounding box of {envelope >= threshold} (bbox_kernel)
        const short4 bb = __ldg(bbox + b + tid);        // col_lo, col_hi, row_lo, row_hi (detector pixels)
        near = !(bb.y < c0 || bb.x > c0 + TC - 1 || bb.w < r0 || bb.z > r0 + TR - 1);
      }
      a = sm.raw[c & 1] + tid * 12;
    } else {
      // One scan round tests kScanPer * 256 bounding boxes (kScanPer independent 8-byte loads per thread in
      // flight, two barriers per round): with one box per thread and round the scan was latency-bound -- at the
      // ~1 % hit rate of BASELINE C3 it took ~58 rounds of an L2 round trip each to collect the 128 candidates
      // of one evaluation phase, about as long as the phase itself.  Hits are appended in beamlet order:
      // sub-round j covers boxes bpos + j * 256 + tid.
      while (ncand < kChunk && bpos < b_end) {           // block-uniform conditions
        const int warp = tid >> 5, lane = tid & 31;
        unsigned ballots[kScanPer];
#pragma unroll
        for (int j = 0; j < kScanPer; ++j) {
          const long long i = bpos + (long long)j * kThreads + tid;
          bool hit = false;
          if (i < b_end) {
            const short4 bb = __ldg(bbox + i);
            hit = !(bb.y < c0 || bb.x > c0 + TC - 1 || bb.w < r0 || bb.z > r0 + TR - 1);
          }
          ballots[j] = __ballot_sync(0xffffffffu, hit);
        }
        if (lane < kScanPer) {
          unsigned b = 0;
#pragma unroll
          for (int j = 0; j < kScanPer; ++j) b = (lane == j) ? ballots[j] : b;
          sm.gcnt[lane][warp] = __popc(b);
        }
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int j = 0; j < kScanPer; ++j) {
          int base = ncand + tot;
#pragma unroll
          for (int w = 0; w < kThreads / 32; ++w) {
            const int n_w = sm.gcnt[j][w];
            base += (w < warp) ? n_w : 0;
            tot += n_w;
          }
          if (ballots[j] & (1u << lane))
            cand[(cand_head + base + __popc(ballots[j] & ((1u << lane) - 1u))) & (kCandRing - 1)] =
                (int)(bpos + (long long)j * kThreads + tid - b_begin);
        }
        __syncthreads();                                 // candidates visible, gcnt reusable
        ncand += tot;
        bpos += (long long)kScanPer * kThreads;
      }
      if (ncand == 0) break;
      cnt = ncand < kChunk ? ncand : kChunk;
      near = tid < cnt;
      if (near) a = table + (b_begin + (long long)cand[(cand_head + tid) & (kCandRing - 1)]) * 12;
    }

    // ---- stage: one thread per beamlet re-centres on the tile origin, culls, compacts
    bool keep = false;
    Rec rec;
    if (near) {
      const double cc = (double)c0, rr = (double)r0;
      // phase, turns
      double t0 = a[0] + cc * (a[1] + a[3] * cc + a[4] * rr) + rr * (a[2] + a[5] * rr);
      double t1 = a[1] + 2.0 * a[3] * cc + a[4] * rr;
      double t2 = a[2] + 2.0 * a[5] * rr + a[4] * cc;
      double t3 = a[3], t4 = a[4], t5 = a[5];
      rec.th[0] = t0 - rint(t0);
      rec.th[1] = t1 - rint(t1);
      rec.th[2] = t2 - rint(t2);
      rec.th[3] = t3 - rint(t3);
      rec.th[4] = t4 - rint(t4);
      rec.th[5] = t5 - rint(t5);
      const double dd = 2.0 * rec.th[3];
      rec.dd = (uint32_t)__double2loint(dd + kMagic);
      // envelope, bits (amplitude = 2^e)
      const double *e = a + 6;
      rec.en[0] = e[0] + cc * (e[1] + e[3] * cc + e[4] * rr) + rr * (e[2] + e[5] * rr);
      rec.en[1] = e[1] + 2.0 * e[3] * cc + e[4] * rr;
      rec.en[2] = e[2] + 2.0 * e[5] * rr + e[4] * cc;
      rec.en[3] = e[3];
      rec.en[4] = e[4];
      rec.en[5] = e[5];
      rec.e2 = (float)e[3];
      rec.rho = 0.f;
      {
        // smooth-envelope test: the per-pixel step of the exponent along a row, E1 + E4 u + E3 (2 v + 1), at
        // the four corners of the tile (it is affine in u, v); NaN / inf coefficients fail the test
        const double s00 = rec.en[1] + rec.en[3];
        const double s01 = s00 + rec.en[3] * (2.0 * (TC - 1));
        const double su = rec.en[4] * (double)(TR - 1);
        const double worst = fmax(fmax(fabs(s00), fabs(s01)), fmax(fabs(s00 + su), fabs(s01 + su)));
        rec.c4r = rec.c4i = 0.f;
        rec.pad2[0] = rec.pad2[1] = 0.f;
        if (worst <= 16.0) {
          rec.rho = (float)exp2(2.0 * rec.en[3]);
          const double d4 = 4.0 * dd, r4 = exp2(8.0 * rec.en[3]);
          double s4, c4;
          sincospi(2.0 * (d4 - rint(d4)), &s4, &c4);
          rec.c4r = (float)(r4 * c4);
          rec.c4i = (float)(r4 * s4);
        }
      }
      rec.cspan = (uint32_t)(TC - 1) << 16;      // all columns unless the culling test below narrows it
      {
        double sd, cd;
        sincospi(2.0 * (dd - rint(dd)), &sd, &cd);
        rec.cr = (float)cd;
        rec.ci = (float)sd;
      }
      // vertex of the (concave) envelope exponent along a row: d/dv = 0
      rec.vs0 = 0.f;
      rec.vs1 = 0.f;
      if (rec.en[3] < -1e-12) {
        const double inv = -0.5 / rec.en[3];
        const double a0 = rec.en[1] * inv, a1 = rec.en[4] * inv;
        // clamped far outside the tile: the pivot is clamped to the strip anyway
        if (isfinite(a0) && isfinite(a1)) {
          rec.vs0 = (float)fmin(fmax(a0, -1e6), 1e6);
          rec.vs1 = (float)fmin(fmax(a1, -1e4), 1e4);
        }
      }
      keep = true;
      if (thr_bits < INFINITY) {
        double G[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) G[j] = -rec.en[j];
        const double gmin = quad_min_rect(G, (double)(TR - 1), (double)(TC - 1));
        keep = !(gmin > thr_bits);  // NaN keeps
        // column extent of {G <= thr} (bounding box of the ellipse), for the per-warp skip
        const double det4 = 4.0 * G[3] * G[5] - G[4] * G[4];
        if (keep && G[3] > 0.0 && G[5] > 0.0 && det4 > 0.0) {
          const double vs = (-2.0 * G[5] * G[1] + G[4] * G[2]) / det4;
          const double us = (-2.0 * G[3] * G[2] + G[4] * G[1]) / det4;
          const double dlt = thr_bits - (G[0] + 0.5 * (G[1] * vs + G[2] * us));
          const double hv = sqrt(fmax(dlt, 0.0) * 4.0 * G[5] / det4);
          const double lo = floor(vs - hv) - 1.0, hi = ceil(vs + hv) + 1.0;   // one pixel of slack
          if (isfinite(lo) && isfinite(hi)) {
            const int ilo = (int)fmin(fmax(lo, 0.0), (double)(TC - 1));
            const int ihi = (int)fmin(fmax(hi, 0.0), (double)(TC - 1));
            if (hi < 0.0 || lo > (double)(TC - 1)) keep = false;             // wholly beside the tile
            rec.cspan = (uint32_t)ilo | ((uint32_t)ihi << 16);
          }
        }
      }
    }
    int n_active;
    if (thr_bits < INFINITY) {
      // order-preserving compaction (determinism: beamlets stay in natural order)
      const unsigned ballot = __ballot_sync(0xffffffffu, keep);
      const int warp = tid >> 5, lane = tid & 31;
      if (warp < kChunk / 32 && lane == 0) sm.warp_cnt[warp] = __popc(ballot);
      __syncthreads();
      if (warp < kChunk / 32) {
        int base = 0;
#pragma unroll
        for (int w = 0; w < kChunk / 32; ++w) base += (w < warp) ? sm.warp_cnt[w] : 0;
        if (keep) sm.rec[base + __popc(ballot & ((1u << lane) - 1u))] = rec;
      }
      if (tid == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kChunk / 32; ++w) tot += sm.warp_cnt[w];
        sm.n_active = tot;
      }
      __syncthreads();
      n_active = sm.n_active;
    } else {
      if (keep) sm.rec[tid] = rec;
      __syncthreads();
      n_active = cnt;
    }

    // ---- evaluate: every thread, its strip of L pixels, all active beamlets of the chunk
    for (int n = 0; n < n_active; ++n) {
      const Rec &q = sm.rec[n];
      if constexpr (WCULL) {
        const uint32_t cs = q.cspan;                                  // warp-uniform: same q, same v0
        if (v0 + L - 1 < (int)(cs & 0xffffu) || v0 > (int)(cs >> 16)) continue;
      }
      ++my_active;
      ++pending;
      // strip-start phase and first difference (turns), fp64 -> 2^-32 fixed point
      // (Horner in v then u: only the two per-thread constants ud, vd stay live in registers)
      const double rt = fma(q.th[4], ud, q.th[1]);                  // T1 + T4 u
      const double wt = fma(q.th[3], vd, rt);                       // T1 + T4 u + T3 v0
      const double th0 = fma(vd, wt, fma(ud, fma(q.th[5], ud, q.th[2]), q.th[0]));
      const double dl0 = fma(q.th[3], vd, wt + q.th[3]);            // T1 + T4 u + T3 (2 v0 + 1)
      const uint32_t t0 = (uint32_t)__double2loint(th0 + kMagic) + 0x100u;  // +0.5 ulp of the 23-bit angle
      const uint32_t d0 = (uint32_t)__double2loint(dl0 + kMagic);
      const uint32_t dd = q.dd;
      // envelope exponent (bits) expanded about an integer PIVOT pixel jp of the strip, the one
      // nearest the envelope's vertex: e(j) = ep + (j-jp) (e1p + (j-jp) e2).  Near a narrow
      // (even sub-pixel) peak the two terms are small, away from it they have equal signs, so
      // fp32 never cancels where the amplitude matters; broad envelopes are insensitive to jp.
      // pivot = clamp(round(vertex column - strip start), 0, L-1), all without the XU pipe:
      // round via the 1.5*2^52 mantissa trick, clamp as integer, rebuild float/double by bits
      // (fp32: the vertex only selects which pixel of the strip the expansion is centred on)
      const float jsf = fmaf(q.vs1, uf, q.vs0 - vf) + 12582912.0f;       // 1.5 * 2^23: low mantissa bits = round(js)
      const int ji = min(max((int)(__float_as_uint(jsf) & 0x7fffffu) - 0x400000, 0), L - 1);
      const float jp = __uint_as_float(0x4b000000u | (uint32_t)ji) - 8388608.0f;
      const double vp = __hiloint2double(0x43300000, ji) + (vd - 4503599627370496.0);
      const double r1 = q.en[1] + q.en[4] * ud;
      const float ep = (float)(q.en[0] + ud * (q.en[2] + q.en[5] * ud) + vp * (r1 + q.en[3] * vp));
      const float e1p = (float)(r1 + 2.0 * q.en[3] * vp);
      const float e2 = q.e2;
      // Phasor recurrence: along the strip the phase is t0 + j d0 + j(j-1)/2 dd, so
      //   z_{j+1} = z_j w_j,  w_{j+1} = w_j c,   z_j = exp(i phase_j), w_j = exp(i (d0 + j dd)), c = exp(i dd).
      // z and w are re-seeded EXACTLY from the fixed-point phase every 4 pixels (MUFU sin / cos), c is
      // a per-beamlet constant from the staging step: 2.0 MUFU (+ 1 ex2... see below) per pixel become
      // 4 sin/cos per 4 pixels -- 2 MUFU per pixel instead of 3 -- and the error of a seed is carried
      // for at most 3 multiplications (<= ~1e-6 absolute on a unit phasor).
      const float cr = q.cr, ci = q.ci;
      const float rho = q.rho;
      if (rho > 0.f) {
        // SMOOTH envelope (warp-uniform: every thread works on the same beamlet): the whole complex term
        // V_j = amp_j z_j follows the same kind of recurrence, V_{j+1} = V_j R_j, R_{j+1} = R_j C' with
        // R_j = 2^(e_{j+1} - e_j) w_j and C' = 2^(2 e2) c.  Seeds every 4 pixels: amp_j, the amplitude ratio
        // (2 ex2) and z_j, w_j (4 sin / cos): 1.5 MUFU and ~12 issue slots per pixel instead of 2 and ~14.5.
        // |exponent step| <= 16 bits per pixel keeps every ratio far inside the fp32 range; steep (sub-pixel)
        // envelopes take the per-pixel path below.
        const float Cr = rho * cr, Ci = rho * ci;
        const float C4r = q.c4r, C4i = q.c4i;
        // the ratio R is seeded once per strip (ex2 + sin / cos) and advanced from group to group by C'^4 (fp64-
        // rounded in the staging step); V is re-seeded exactly for every group: 3 MUFU per 4 pixels after the first
        float gr, gi;
        {
          const uint32_t dj = d0 + 0x100u;
          const float fw = __uint_as_float((dj >> 9) | 0x3f800000u);
          const float aw = fmaf(fw, 6.28318530717958648f, -9.42477796076937972f);
          const float dj_ = -jp;
          const float gj = fmaf(e2, fmaf(2.0f, dj_, 1.0f), e1p);          // exponent step from pixel 0 to pixel 1
          float rat;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(rat) : "f"(gj));
          gr = -rat * __cosf(aw);
          gi = -rat * __sinf(aw);
        }
#pragma unroll
        for (int s4 = 0; s4 < L; s4 += 4) {
          const uint32_t tj = t0 + (uint32_t)s4 * d0 + (uint32_t)(s4 * (s4 - 1) / 2) * dd;
          const float fz = __uint_as_float((tj >> 9) | 0x3f800000u);       // 1 + frac(turns)
          const float az = fmaf(fz, 6.28318530717958648f, -9.42477796076937972f);
          const float dj_ = (float)s4 - jp;
          const float ej = fmaf(dj_, fmaf(dj_, e2, e1p), ep);            // exponent at the group's first pixel
          float amp;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(amp) : "f"(ej));
          float vr = -amp * __cosf(az), vi = -amp * __sinf(az);          // V = amp exp(i 2 pi frac)
          float rr = gr, ri = gi;                                        // R of the group's first pixel
          if (s4 + 4 < L) {                                              // R of the next group
            const float ngr = fmaf(gr, C4r, -(gi * C4i));
            const float ngi = fmaf(gr, C4i, gi * C4r);
            gr = ngr;
            gi = ngi;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = s4 + k;
            pr[j] += vr;
            pi[j] += vi;
            if (k < 3) {
              const float nvr = fmaf(vr, rr, -(vi * ri));
              const float nvi = fmaf(vr, ri, vi * rr);
              vr = nvr;
              vi = nvi;
            }
            if (k < 2) {
              const float nrr = fmaf(rr, Cr, -(ri * Ci));
              const float nri = fmaf(rr, Ci, ri * Cr);
              rr = nrr;
              ri = nri;
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int s4 = 0; s4 < L; s4 += 4) {
        const uint32_t tj = t0 + (uint32_t)s4 * d0 + (uint32_t)(s4 * (s4 - 1) / 2) * dd;
        const uint32_t dj = d0 + (uint32_t)s4 * dd + 0x100u;
        float zr, zi, wr, wi;
        {
          // exp(i 2pi frac) = -(cos ang + i sin ang), ang = 2 pi frac - pi in [-pi, pi)
          const float fz = __uint_as_float((tj >> 9) | 0x3f800000u);       // 1 + frac(turns)
          const float az = fmaf(fz, 6.28318530717958648f, -9.42477796076937972f);
          const float fw = __uint_as_float((dj >> 9) | 0x3f800000u);
          const float aw = fmaf(fw, 6.28318530717958648f, -9.42477796076937972f);
          zr = -__cosf(az);
          zi = -__sinf(az);
          wr = -__cosf(aw);
          wi = -__sinf(aw);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = s4 + k;
          const float dj_ = (float)j - jp;
          const float ej = fmaf(dj_, fmaf(dj_, e2, e1p), ep);
          float amp;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(amp) : "f"(ej));
          pr[j] = fmaf(amp, zr, pr[j]);
          pi[j] = fmaf(amp, zi, pi[j]);
          if (k < 3) {
            const float nzr = fmaf(zr, wr, -(zi * wi));
            const float nzi = fmaf(zr, wi, zi * wr);
            zr = nzr;
            zi = nzi;
          }
          if (k < 2) {
            const float nwr = fmaf(wr, cr, -(wi * ci));
            const float nwi = fmaf(wr, ci, wi * cr);
            wr = nwr;
            wi = nwi;
          }
        }
      }
    }

    // ---- flush the fp32 partials into the fp64 accumulators once they hold ~kChunk terms (bounds the
    // fp32 accumulation length; when most beamlets are culled for this strip the flush is rare)
    if (pending >= kChunk / 2) {
#pragma unroll
      for (int j = 0; j < L; ++j) {
        sm.acc[j * kThreads + tid] += (double)pr[j];
        sm.acc[(L + j) * kThreads + tid] += (double)pi[j];
        pr[j] = 0.f;
        pi[j] = 0.f;
      }
      pending = 0;
    }
    __syncthreads();  // records and raw[c&1] are free for the next stage / TMA
    if (gather) {       // drop the processed candidates: the ring's head moves on (no data movement, no barrier)
      cand_head = (cand_head + cnt) & (kCandRing - 1);
      ncand -= cnt;
    }
  }
  if (pending) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      sm.acc[j * kThreads + tid] += (double)pr[j];
      sm.acc[(L + j) * kThreads + tid] += (double)pi[j];
    }
  }

  // ---- write the tile
  const int row = r0 + u;
  const bool row_ok = row < g.row0 + g.nrows && row < g.H;
  if (row_ok) {
    const size_t base = (size_t)(row - g.row0) * g.W + c0 + v0;              // in the output (row0-relative)
    const size_t pbase = (size_t)(ty * TR + u) * g.W + c0 + v0;               // in the call's partial buffer
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (c0 + v0 + j < g.W) {
        const double re = sm.acc[j * kThreads + tid], im = sm.acc[(L + j) * kThreads + tid];
        if (g.via_partial) {
          partial[(size_t)split * ((size_t)g.lrows * g.W) + pbase + j] = make_double2(re, im);
        } else if (out_is_c128) {
          static_cast<double2 *>(out)[base + j] = make_double2(re, im);
        } else {
          static_cast<float2 *>(out)[base + j] = make_float2((float)re, (float)im);
        }
      }
    }
  }
  if (evals) {   // executed evaluations: beamlets this thread evaluated x valid pixels of its strip
    const int rows_valid = min(TR, g.row0 + g.nrows - r0);
    const int cols_valid = min(L, g.W - c0 - v0);
    if (u < rows_valid && cols_valid > 0 && my_active) atomicAdd(evals, my_active * (unsigned long long)cols_valid);
  }
}

__global__ void __launch_bounds__(256)
    split_reduce_kernel(const double2 *__restrict__ partial, int nsplit, size_t npix,
                        void *__restrict__ out, int out_is_c128,
                        const unsigned long long *__restrict__ sep_guard, const TgPeers peers, int W, int H,
                        int tile_rows, int ty_stride, int ty_phase) {
  if (sep_guard && tg_key_is_separable(*sep_guard)) return;
  // bounded grid with a stride loop: when the device-side verdict cancels the launch it costs ~2 us, not 3 ns for
  // each of the thousands of CTAs a pixel-per-thread grid needs (13.7 us per C2 step, measured)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    size_t o = i;                      // position in the output
    if (ty_stride) {                   // cyclic tile rows: local row -> detector row
      const size_t lr = i / (size_t)W, col = i - lr * (size_t)W;
      const size_t grow = ((size_t)ty_phase + (lr / tile_rows) * (size_t)ty_stride) * tile_rows + lr % tile_rows;
      if (grow >= (size_t)H) continue;
      o = grow * (size_t)W + col;
    }
    double re = 0.0, im = 0.0;
    for (int s = 0; s < nsplit; ++s) {
      const double2 v = partial[(size_t)s * npix + i];
      re += v.x;
      im += v.y;
    }
    if (out_is_c128) {
      static_cast<double2 *>(out)[o] = make_double2(re, im);
      for (int p = 0; p < peers.n; ++p) static_cast<double2 *>(peers.ptr[p])[o] = make_double2(re, im);  // NVLink P2P
    } else {
      static_cast<float2 *>(out)[o] = make_float2((float)re, (float)im);
      for (int p = 0; p < peers.n; ++p) static_cast<float2 *>(peers.ptr[p])[o] = make_float2((float)re, (float)im);
    }
  }
}

// ---- arbitrary observation points ------------------------------------------------------
// One thread per point; beamlet polynomials (metre space, radians) streamed through smem.
constexpr int kPtChunk = 128;
__global__ void __launch_bounds__(128)
    field_points_kernel(long long nb, const double *__restrict__ poly, long long npts,
                        const double *__restrict__ r_xy, void *__restrict__ out, int out_is_c128) {
  __shared__ double s_poly[kPtChunk * 12];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < npts;
  const double x = ok ? r_xy[i * 2] : 0.0, y = ok ? r_xy[i * 2 + 1] : 0.0;
  double are = 0.0, aim = 0.0;
  for (long long b = 0; b < nb; b += kPtChunk) {
    const int cnt = (int)((nb - b) < kPtChunk ? (nb - b) : kPtChunk);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt * 12; k += blockDim.x) s_poly[k] = poly[b * 12 + k];
    __syncthreads();
    float pr = 0.f, pi = 0.f;
    for (int n = 0; n < cnt; ++n) {
      const double *c = s_poly + n * 12;
      const double re = c[0] + x * (c[2] + c[6] * x + c[8] * y) + y * (c[4] + c[10] * y);
      const double im = c[1] + x * (c[3] + c[7] * x + c[9] * y) + y * (c[5] + c[11] * y);
      const double t = re * kInv2Pi;
      const float fr = (float)(t - rint(t));              // turns in [-0.5, 0.5]
      const float ang = fr * 6.28318530717958648f;
      float amp;
      const float ex = (float)(-im * kLog2e);
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(amp) : "f"(ex));
      pr = fmaf(amp, __cosf(ang), pr);
      pi = fmaf(amp, __sinf(ang), pi);
    }
    are += (double)pr;
    aim += (double)pi;
  }
  if (ok) {
    if (out_is_c128) static_cast<double2 *>(out)[i] = make_double2(are, aim);
    else static_cast<float2 *>(out)[i] = make_float2((float)are, (float)aim);
  }
}

inline unsigned reduce_grid(size_t npix) {
  const size_t blocks = (npix + 255) / 256, cap = 148 * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// choose the beamlet-split count so tiles*S fills whole waves of resident CTAs
int choose_split(long long tiles, long long nb, int slots, bool culling, size_t npix) {
  if (tiles <= 0) return 1;
  const long long max_by_chunks = nb / kChunk > 0 ? nb / kChunk : 1;
  long long max_s = 32 < max_by_chunks ? 32 : max_by_chunks;
  const size_t ws_cap = (size_t)1 << 30;  // <= 1 GiB of split partials
  while (max_s > 1 && (size_t)max_s * npix * 16 > ws_cap) --max_s;
  (void)culling;  // culled work is irregular; finer granularity only helps balance
  double best_eff = -1.0;
  int best = 1;
  for (long long s = 1; s <= max_s; ++s) {
    const double waves = (double)(tiles * s) / slots;
    const double eff = waves / ceil(waves);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = (int)s;
    }
  }
  return best;
}

}  // namespace

// shared with separable.cu: metre-space polynomials -> pixel-space {turns, bits} table
int tg_launch_prep(int64_t nb, const double *poly, const double px2m[6], int H, int W, double *table,
                   unsigned long long *gref_key, cudaStream_t st, const TgPrepExtra *extra) {
  TgPrepExtra ex;
  ex.sep_key = ex.peak_key = nullptr;
  ex.row0 = 0;
  ex.nrows = H;
  if (extra) ex = *extra;
  prep_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(nb, poly, px2m[0], px2m[1], px2m[2], px2m[3],
                                                            px2m[4], px2m[5], H, W, table, gref_key, ex);
  return tg_launch_check("prep_kernel");
}

extern "C" int tg_field_sum_grid(int64_t nb, const double *poly, const double px2m[6], int H, int W,
                                 int row0, int nrows, void *out, int out_is_c128, int cull_bits,
                                 long long *n_evals_out, void *stream) {
  return tg_field_grid_run(nb, poly, px2m, H, W, row0, nrows, out, out_is_c128, cull_bits, n_evals_out,
                           nullptr, static_cast<cudaStream_t>(stream));
}

// sep_guard != NULL: every kernel of this path returns at once when the device-side key says the
// beamlets are separable (the tensor-core path, enqueued by the same call, does the work instead).
int tg_field_grid_run(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                      int nrows, void *out, int out_is_c128, int cull_bits, long long *n_evals_out,
                      const unsigned long long *sep_guard, cudaStream_t stream, const TgPeers *peers,
                      const TgEmit *emit) {
  TG_REQUIRE(nb >= 0 && H > 0 && W > 0, "bad shape");
  TG_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "bad row range");
  TG_REQUIRE(px2m && out, "null pointer");
  TG_REQUIRE(cull_bits >= 0 && cull_bits < 1000, "bad cull_bits");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_evals_out) *n_evals_out = 0;
  if (nrows == 0) return TG_OK;
  if (n_evals_out && tg_stream_is_capturing(st)) {
    tg_set_error("counting executed evaluations reads a counter back and cannot be captured into a CUDA graph");
    return TG_EUNSUPPORTED;
  }
  const size_t npix = (size_t)nrows * W;
  const size_t elt = out_is_c128 ? 16 : 8;
  TgPeers no_peers;
  no_peers.n = 0;
  const TgPeers &pe = peers ? *peers : no_peers;
  TG_REQUIRE(!(emit && pe.n > 0), "row-block emission and peer images are separate modes");
  if (nb == 0) {
    TG_CUDA(cudaMemsetAsync(out, 0, npix * elt, st));
    for (int p = 0; p < pe.n; ++p) TG_CUDA(cudaMemsetAsync(pe.ptr[p], 0, npix * elt, st));
    if (emit) return tg_emit_block(emit, 0, st, out, 0, npix * elt);
    return TG_OK;
  }
  TG_REQUIRE(poly, "null poly");

  constexpr int L = 16, SPR = 8;
  using S = FieldSmem<L, SPR>;
  // Row-sharded sum over peer images: instead of its contiguous row block this rank takes every cyc_world-th tile
  // row (32 rows) of the WHOLE detector -- narrow beamlets illuminate a disc, so contiguous blocks leave the ranks
  // that own the middle rows with 1.27x the average work (measured: C3 at 8 ranks 1.18 ms against 0.71 ms of ideal
  // compute).  All images are addressed from row 0 then.
  if (pe.cyc_world > 1 && !n_evals_out) {
    TgPeers base = pe;
    for (int p = 0; p < base.n; ++p) base.ptr[p] = static_cast<unsigned char *>(base.ptr[p]) - pe.shard_off_bytes;
    void *out0 = static_cast<unsigned char *>(out) - pe.shard_off_bytes;
    FieldGeom g;
    g.H = H; g.W = W; g.row0 = 0; g.nrows = H;
    g.tiles_x = (W + S::TC - 1) / S::TC;
    g.nb = nb;
    g.cull_bits = cull_bits;
    g.ty_stride = pe.cyc_world;
    g.ty_phase = pe.cyc_rank;
    const int tile_rows_total = (H + S::TR - 1) / S::TR;
    g.tiles_y = tile_rows_total > pe.cyc_rank ? (tile_rows_total - pe.cyc_rank + pe.cyc_world - 1) / pe.cyc_world : 0;
    if (g.tiles_y == 0) return TG_OK;
    g.lrows = g.tiles_y * S::TR;
    int dev = 0, sms = 148;
    TG_CUDA(cudaGetDevice(&dev));
    tg_tune_mempool(dev);
    TG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (long long)g.tiles_x * g.tiles_y;
    const size_t lpix = (size_t)g.lrows * W;
    g.nsplit = choose_split(tiles, nb, 2 * sms, cull_bits > 0, lpix);
    g.via_partial = 1;
    const size_t table_bytes = (size_t)nb * 96;
    const size_t part_bytes = (size_t)g.nsplit * lpix * 16;
    const bool use_bbox = cull_bits > 0 && H <= 32768 && W <= 32768;
    const size_t bbox_bytes = use_bbox ? (size_t)nb * sizeof(short4) : 0;
    TgAsyncBuf wsb(st);
    TG_CUDA(wsb.alloc(table_bytes + 256 + part_bytes + bbox_bytes));
    unsigned char *ws = wsb.as<unsigned char>();
    short4 *bbox = use_bbox ? reinterpret_cast<short4 *>(ws + table_bytes + 256 + part_bytes) : nullptr;
    double *table = reinterpret_cast<double *>(ws);
    unsigned long long *gref = reinterpret_cast<unsigned long long *>(ws + table_bytes);
    double2 *partial = reinterpret_cast<double2 *>(ws + table_bytes + 256);
    TG_CUDA(cudaMemsetAsync(gref, 0xFF, 8, st));
    int rc = tg_launch_prep(nb, poly, px2m, H, W, table, cull_bits > 0 ? gref : nullptr, st);
    if (rc != TG_OK) return rc;
    if (use_bbox) {
      bbox_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(nb, table, H, W, gref, cull_bits, bbox);
      if ((rc = tg_launch_check("bbox_kernel")) != TG_OK) return rc;
    }
    const size_t smem = sizeof(S);
    auto kern = cull_bits > 0 ? field_grid_kernel<L, SPR, true> : field_grid_kernel<L, SPR, false>;
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)tiles, (unsigned)g.nsplit);
    kern<<<grid, kThreads, smem, st>>>(table, g, cull_bits > 0 ? gref : nullptr, bbox, out0, out_is_c128, partial,
                                       nullptr, sep_guard);
    if ((rc = tg_launch_check("field_grid_kernel")) != TG_OK) return rc;
    split_reduce_kernel<<<reduce_grid(lpix), 256, 0, st>>>(partial, g.nsplit, lpix, out0, out_is_c128,
                                                                       sep_guard, base, W, H, S::TR, g.ty_stride,
                                                                       g.ty_phase);
    return tg_launch_check("split_reduce_kernel");
  }
  int block_rows = nrows;
  if (emit) {
    TG_REQUIRE(emit->block_rows > 0 && emit->block_rows % S::TR == 0 && emit->host_out && emit->ev,
               "bad emission block");
    block_rows = emit->block_rows < nrows ? emit->block_rows : nrows;
  }
  FieldGeom g;
  g.H = H; g.W = W;
  g.tiles_x = (W + S::TC - 1) / S::TC;
  g.nb = nb;
  g.cull_bits = cull_bits;
  g.ty_stride = 0;
  g.ty_phase = 0;
  int dev = 0, sms = 148;
  TG_CUDA(cudaGetDevice(&dev));
  tg_tune_mempool(dev);
  TG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // the split count of a full block (the last, shorter block uses its own)
  const size_t blk_pix = (size_t)block_rows * W;
  const int nsplit_max = choose_split((long long)g.tiles_x * ((block_rows + S::TR - 1) / S::TR), nb, 2 * sms,
                                      cull_bits > 0, blk_pix);
  const bool via_partial_max = nsplit_max > 1 || pe.n > 0;

  // workspace: table (nb*96 B) | gref key (8) | evals (8) | split partials of one block | per-beamlet bounding boxes
  const size_t table_bytes = (size_t)nb * 96;
  const size_t part_bytes = via_partial_max ? (size_t)nsplit_max * blk_pix * 16 : 0;
  const bool use_bbox = cull_bits > 0 && H <= 32768 && W <= 32768;
  const size_t bbox_bytes = use_bbox ? (size_t)nb * sizeof(short4) : 0;
  TgAsyncBuf wsb(st);
  TG_CUDA(wsb.alloc(table_bytes + 256 + part_bytes + bbox_bytes));
  unsigned char *ws = wsb.as<unsigned char>();
  short4 *bbox = use_bbox ? reinterpret_cast<short4 *>(ws + table_bytes + 256 + part_bytes) : nullptr;
  double *table = reinterpret_cast<double *>(ws);
  unsigned long long *gref = reinterpret_cast<unsigned long long *>(ws + table_bytes);
  unsigned long long *evals = gref + 1;
  double2 *partial = part_bytes ? reinterpret_cast<double2 *>(ws + table_bytes + 256) : nullptr;
  TG_CUDA(cudaMemsetAsync(gref, 0xFF, 8, st));
  TG_CUDA(cudaMemsetAsync(evals, 0, 8, st));

  int rc = tg_launch_prep(nb, poly, px2m, H, W, table, cull_bits > 0 ? gref : nullptr, st);
  if (rc != TG_OK) return rc;
  if (use_bbox) {
    bbox_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(nb, table, H, W, gref, cull_bits, bbox);
    rc = tg_launch_check("bbox_kernel");
    if (rc != TG_OK) return rc;
  }
  const size_t smem = sizeof(S);
  auto kern = cull_bits > 0 ? field_grid_kernel<L, SPR, true> : field_grid_kernel<L, SPR, false>;
  TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blk = 0;
  for (int r = 0; r < nrows; r += block_rows, ++blk) {
    const int nr = (nrows - r) < block_rows ? (nrows - r) : block_rows;
    g.row0 = row0 + r;
    g.nrows = nr;
    g.lrows = nr;
    g.tiles_y = (nr + S::TR - 1) / S::TR;
    const long long tiles = (long long)g.tiles_x * g.tiles_y;
    const size_t bpix = (size_t)nr * W;
    g.nsplit = nr == block_rows ? nsplit_max : choose_split(tiles, nb, 2 * sms, cull_bits > 0, bpix);
    if (g.nsplit > nsplit_max) g.nsplit = nsplit_max;          // the partial buffer is sized for nsplit_max
    g.via_partial = (g.nsplit > 1 || pe.n > 0) ? 1 : 0;        // peer stores are issued by the reduce kernel (coalesced)
    unsigned char *out_r = static_cast<unsigned char *>(out) + (size_t)r * W * elt;
    dim3 grid((unsigned)tiles, (unsigned)g.nsplit);
    kern<<<grid, kThreads, smem, st>>>(table, g, cull_bits > 0 ? gref : nullptr, bbox, out_r, out_is_c128, partial,
                                       n_evals_out ? evals : nullptr, sep_guard);
    rc = tg_launch_check("field_grid_kernel");
    if (rc != TG_OK) return rc;
    if (g.via_partial) {
      TgPeers pr = pe;
      for (int p = 0; p < pr.n; ++p) pr.ptr[p] = static_cast<unsigned char *>(pr.ptr[p]) + (size_t)r * W * elt;
      split_reduce_kernel<<<reduce_grid(bpix), 256, 0, st>>>(partial, g.nsplit, bpix, out_r,
                                                                         out_is_c128, sep_guard, pr, W, H, S::TR, 0, 0);
      rc = tg_launch_check("split_reduce_kernel");
      if (rc != TG_OK) return rc;
    }
    if (emit) {
      rc = tg_emit_block(emit, blk, st, out_r, (size_t)r * W * elt, bpix * elt);
      if (rc != TG_OK) return rc;
    }
  }
  if (n_evals_out) {
    unsigned long long h = 0;
    TG_CUDA(cudaMemcpyAsync(&h, evals, 8, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    *n_evals_out = (long long)h;
  }
  return TG_OK;
}

extern "C" int tg_field_sum_points(int64_t nb, const double *poly, int64_t npts, const double *r_xy,
                                   void *out, int out_is_c128, void *stream) {
  TG_REQUIRE(nb >= 0 && npts >= 0, "bad sizes");
  if (npts == 0) return TG_OK;
  TG_REQUIRE(out && r_xy && (poly || nb == 0), "null pointer");
  field_points_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      nb, poly, npts, r_xy, out, out_is_c128);
  return tg_launch_check("field_points_kernel");
}
