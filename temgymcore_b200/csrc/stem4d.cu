// Fused 4D-STEM shadow-image backprojection (BASELINE config C5).
//
// Every (scan position s, detector pixel p) pair is one ray.  The reference ships the building
// blocks only (Scanner / Descanner / DescanError components.py:27-115, 252-372; ScanGrid /
// Detector pixel<->metre maps grid.py:120-182; transfer_rays_pt_src transfer.py:57-123;
// inplace_sum utils.py:83-114) -- the composite workflow lives in a private downstream repo
// (SURVEY.md F8).  This kernel fuses them:
//   (spx, spy) = ScanGrid.pixels_to_metres(sy, sx)                      scan position [m]
//   (xd, yd)   = Detector.pixels_to_metres(dy, dx)                      detector pixel centre [m]
//   theta      = Bdet^-1 ((xd, yd) - Adet r0 - edet(s))                 slopes at the point source
//   (xs, ys)   = Asamp r0 + Bsamp theta + esamp(s)                      transfer_rays_pt_src to the sample plane
//   (py, px)   = OutGrid.metres_to_pixels(xs, ys), round half to even   sample-plane pixel
//   out[py, px] += data[sy, sx, dy, dx]                                 inplace_sum (bounds-checked)
// where the 5th ABCD column (descan error!) is affine in the scan position:
//   e(s) = e0 + spx e1 + spy e2   (Scanner / Descanner add offsets linear in scan_pos times _one).
// All coordinate arithmetic is fp64 with a fixed operation order and no FMA contraction
// (-fmad=false), so pixel indices are bit-reproducible against the numpy oracle.
//
// Mapping: one CTA per scan position streams that position's detector frame (coalesced fp32 /
// uint loads, 4 B per ray: the kernel is HBM-read-bound on the 4D dataset) and accumulates into a
// 64 x 64 shared-memory tile centred on the frame's footprint on the sample plane (a frame lands
// on a few hundred sample pixels: privatisation turns 65 536 contended global atomics per frame
// into a few hundred); rays falling outside the tile go straight to global atomics.
#include <math.h>
#include <stdlib.h>
#include "tg_common.cuh"

namespace {

constexpr int kTile = 64;
constexpr int kThreads4d = 256;

struct Stem4dGeom {
  int Sy, Sx, Dy, Dx, Oy, Ox;
  double Ts[6];      // scan px->m : y = Ts0*sy + Ts1*sx + Ts2 ; x = Ts3*sy + Ts4*sx + Ts5
  double Td[6];      // detector px->m, same layout
  double To[6];      // output grid m->px : py = To0*y + To1*x + To2 ; px = To3*y + To4*x + To5
  double cdet[2];    // Adet r0          (x, y)
  double edet[6];    // e0x e0y e1x e1y e2x e2y
  double Binv[4];    // Bdet^-1 row-major
  double csamp[2];   // Asamp r0
  double Bs[4];      // Bsamp row-major
  double esamp[6];
};

__device__ __forceinline__ void ray_to_pixel(const Stem4dGeom &g, double spx, double spy, double edx, double edy,
                                             double esx, double esy, int dy, int dx, int &py, int &px) {
  const double fy = (double)dy, fx = (double)dx;
  const double yd = (g.Td[0] * fy + g.Td[1] * fx) + g.Td[2];
  const double xd = (g.Td[3] * fy + g.Td[4] * fx) + g.Td[5];
  const double rx = (xd - g.cdet[0]) - edx;
  const double ry = (yd - g.cdet[1]) - edy;
  const double tx = g.Binv[0] * rx + g.Binv[1] * ry;
  const double ty = g.Binv[2] * rx + g.Binv[3] * ry;
  const double xs = (g.csamp[0] + (g.Bs[0] * tx + g.Bs[1] * ty)) + esx;
  const double ys = (g.csamp[1] + (g.Bs[2] * tx + g.Bs[3] * ty)) + esy;
  const double fpy = (g.To[0] * ys + g.To[1] * xs) + g.To[2];
  const double fpx = (g.To[3] * ys + g.To[4] * xs) + g.To[5];
  py = isnan(fpy) ? 0 : __double2int_rn(fpy);
  px = isnan(fpx) ? 0 : __double2int_rn(fpx);
}

template <typename T>
__global__ void __launch_bounds__(kThreads4d)
    stem4d_backproject_kernel(const __grid_constant__ Stem4dGeom g, const T *__restrict__ data,
                              float *__restrict__ out, int s_begin) {
  __shared__ float tile[kTile * kTile];
  const int s = s_begin + blockIdx.x;
  const int sy = s / g.Sx, sx = s % g.Sx;
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) tile[k] = 0.f;
  // per-scan-position quantities (every thread computes the same values)
  const double fsy = (double)sy, fsx = (double)sx;
  const double spy = (g.Ts[0] * fsy + g.Ts[1] * fsx) + g.Ts[2];
  const double spx = (g.Ts[3] * fsy + g.Ts[4] * fsx) + g.Ts[5];
  const double edx = (g.edet[0] + spx * g.edet[2]) + spy * g.edet[4];
  const double edy = (g.edet[1] + spx * g.edet[3]) + spy * g.edet[5];
  const double esx = (g.esamp[0] + spx * g.esamp[2]) + spy * g.esamp[4];
  const double esy = (g.esamp[1] + spx * g.esamp[3]) + spy * g.esamp[5];
  int cy, cx;  // footprint centre = image of the central detector pixel
  ray_to_pixel(g, spx, spy, edx, edy, esx, esy, g.Dy / 2, g.Dx / 2, cy, cx);
  const int ty0 = cy - kTile / 2, tx0 = cx - kTile / 2;
  __syncthreads();

  const long long npix = (long long)g.Dy * g.Dx;
  const T *frame = data + (long long)s * npix;
  for (long long p = threadIdx.x; p < npix; p += kThreads4d) {
    const float v = (float)frame[p];
    const int dy = (int)(p / g.Dx), dx = (int)(p % g.Dx);
    int py, px;
    ray_to_pixel(g, spx, spy, edx, edy, esx, esy, dy, dx, py, px);
    if (py < 0 || py >= g.Oy || px < 0 || px >= g.Ox) continue;  // inplace_sum bounds check
    const int ly = py - ty0, lx = px - tx0;
    if (ly >= 0 && ly < kTile && lx >= 0 && lx < kTile)
      atomicAdd(&tile[ly * kTile + lx], v);
    else
      atomicAdd(&out[(long long)py * g.Ox + px], v);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) {
    const float v = tile[k];
    if (v != 0.f) {
      const int py = ty0 + k / kTile, px = tx0 + k % kTile;
      if (py >= 0 && py < g.Oy && px >= 0 && px < g.Ox) atomicAdd(&out[(long long)py * g.Ox + px], v);
    }
  }
}

// ---- fast path -------------------------------------------------------------------------
// The map (dy, dx) -> (fpy, fpx) is affine for a fixed scan position.  The fast kernel evaluates
// it with two FMAs per coordinate from per-CTA coefficients; whenever the result lies within a
// small tolerance of a half-integer (a rounding tie, or the edge of the grid) the ray is
// re-evaluated with the step-wise arithmetic above, so the pixel index is ALWAYS the step-wise
// one (bit-exact parity) while ~all rays cost 4 FMAs instead of ~37 fp64 operations.
// Each thread owns 8 consecutive detector pixels of a row (two 16-byte loads) and merges runs of
// equal target pixels in registers before touching the shared-memory tile.
struct AffineYX {
  double by, bx;      // value at (dy, dx) = (0, 0)
  double ry, rx;      // per detector row
  double cy, cx;      // per detector column
};

// round-half-even of f via the 1.5*2^52 mantissa trick (3 DADD): the integer is the low word of
// f + magic.  `flag` is raised when f is within 1e-6 of a half-integer (composite-vs-step-wise
// discrepancy is < 1e-7 for |f| < 2^20) or is NaN; |f| >= 2^20 is reported as far out of bounds
// (both evaluations agree there: grids are limited to 2^20 pixels per side).
__device__ __forceinline__ int round_guarded(double f, bool &flag, bool &far) {
  const double kMagic = 6755399441055744.0;
  const double t = f + kMagic;
  const double frac = f - (t - kMagic);
  flag |= !(fabs(frac) <= 0.5 - 1e-6);
  far |= !(fabs(f) < 1048576.0);
  return __double2loint(t);
}

template <typename T>
__device__ __forceinline__ void load8(const T *p, float v[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float v[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
  const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<unsigned short>(const unsigned short *p, float v[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
  v[0] = (float)(a.x & 0xffffu); v[1] = (float)(a.x >> 16); v[2] = (float)(a.y & 0xffffu); v[3] = (float)(a.y >> 16);
  v[4] = (float)(a.z & 0xffffu); v[5] = (float)(a.z >> 16); v[6] = (float)(a.w & 0xffffu); v[7] = (float)(a.w >> 16);
}

template <typename T>
__global__ void __launch_bounds__(kThreads4d)
    stem4d_backproject_fast_kernel(const __grid_constant__ Stem4dGeom g, const T *__restrict__ data,
                                   float *__restrict__ out, int s_begin) {
  __shared__ float tile[kTile * kTile];
  const int s = s_begin + blockIdx.x;
  const int sy = s / g.Sx, sx = s % g.Sx;
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) tile[k] = 0.f;
  const double fsy = (double)sy, fsx = (double)sx;
  const double spy = (g.Ts[0] * fsy + g.Ts[1] * fsx) + g.Ts[2];
  const double spx = (g.Ts[3] * fsy + g.Ts[4] * fsx) + g.Ts[5];
  const double edx = (g.edet[0] + spx * g.edet[2]) + spy * g.edet[4];
  const double edy = (g.edet[1] + spx * g.edet[3]) + spy * g.edet[5];
  const double esx = (g.esamp[0] + spx * g.esamp[2]) + spy * g.esamp[4];
  const double esy = (g.esamp[1] + spx * g.esamp[3]) + spy * g.esamp[5];
  // composite affine coefficients (same for every thread of the CTA)
  AffineYX a;
  {
    // constant term: the step-wise chain at (dy, dx) = (0, 0), kept in floating point
    const double yd = g.Td[2], xd = g.Td[5];
    const double rx = (xd - g.cdet[0]) - edx, ry = (yd - g.cdet[1]) - edy;
    const double tx = g.Binv[0] * rx + g.Binv[1] * ry, ty = g.Binv[2] * rx + g.Binv[3] * ry;
    const double xs = (g.csamp[0] + (g.Bs[0] * tx + g.Bs[1] * ty)) + esx;
    const double ys = (g.csamp[1] + (g.Bs[2] * tx + g.Bs[3] * ty)) + esy;
    a.by = (g.To[0] * ys + g.To[1] * xs) + g.To[2];
    a.bx = (g.To[3] * ys + g.To[4] * xs) + g.To[5];
    // slopes: d(out px)/d(det row) and /d(det col) through Td -> Binv -> Bs -> To
    auto slope = [&](double dyd, double dxd, double &oy, double &ox) {
      const double ttx = g.Binv[0] * dxd + g.Binv[1] * dyd, tty = g.Binv[2] * dxd + g.Binv[3] * dyd;
      const double dxs = g.Bs[0] * ttx + g.Bs[1] * tty, dys = g.Bs[2] * ttx + g.Bs[3] * tty;
      oy = g.To[0] * dys + g.To[1] * dxs;
      ox = g.To[3] * dys + g.To[4] * dxs;
    };
    slope(g.Td[0], g.Td[3], a.ry, a.rx);
    slope(g.Td[1], g.Td[4], a.cy, a.cx);
  }
  int cy, cx;
  ray_to_pixel(g, spx, spy, edx, edy, esx, esy, g.Dy / 2, g.Dx / 2, cy, cx);
  const int ty0 = cy - kTile / 2, tx0 = cx - kTile / 2;
  __syncthreads();

  const int groups_per_row = g.Dx >> 3;
  const int ngroups = g.Dy * groups_per_row;
  const T *frame = data + (long long)s * g.Dy * g.Dx;
  for (int grp = threadIdx.x; grp < ngroups; grp += kThreads4d) {
    const int dy = grp / groups_per_row, dx0 = (grp - dy * groups_per_row) << 3;
    float v[8];
    load8<T>(frame + (long long)dy * g.Dx + dx0, v);
    const double fx0 = (double)dx0;
    const double y0v = fma(fx0, a.cy, fma((double)dy, a.ry, a.by));
    const double x0v = fma(fx0, a.cx, fma((double)dy, a.rx, a.bx));
    int cur = -2;        // shared-tile slot with a pending partial sum (-2: none)
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bool tie = false, far = false;
      int py = round_guarded(fma((double)j, a.cy, y0v), tie, far);
      int px = round_guarded(fma((double)j, a.cx, x0v), tie, far);
      if (far) {
        if (!(fma((double)j, a.cy, y0v) == fma((double)j, a.cy, y0v)) ||
            !(fma((double)j, a.cx, x0v) == fma((double)j, a.cx, x0v)))
          tie = true;          // NaN geometry: let the step-wise path decide (it maps NaN to pixel 0)
        else
          continue;            // far outside every admissible grid
      }
      if (tie) ray_to_pixel(g, spx, spy, edx, edy, esx, esy, dy, dx0 + j, py, px);  // exact step-wise path
      if ((unsigned)py >= (unsigned)g.Oy || (unsigned)px >= (unsigned)g.Ox) continue;  // inplace_sum bounds
      const int ly = py - ty0, lx = px - tx0;
      if ((unsigned)ly < (unsigned)kTile && (unsigned)lx < (unsigned)kTile) {
        const int key = ly * kTile + lx;
        if (key != cur) {
          if (cur >= 0) atomicAdd(&tile[cur], sum);
          cur = key;
          sum = 0.f;
        }
        sum += v[j];
      } else {
        atomicAdd(&out[(long long)py * g.Ox + px], v[j]);
      }
    }
    if (cur >= 0) atomicAdd(&tile[cur], sum);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) {
    const float t = tile[k];
    if (t != 0.f) {
      const int py = ty0 + k / kTile, px = tx0 + k % kTile;
      if (py >= 0 && py < g.Oy && px >= 0 && px < g.Ox) atomicAdd(&out[(long long)py * g.Ox + px], t);
    }
  }
}

// ---- integer DDA path ---------------------------------------------------------------------
// When a frame's footprint fits inside the 64 x 64 tile (checked on the host from the slopes, which
// do not depend on the scan position) the per-ray work is integer only.  Tile-local coordinates
// u = f - tile_origin + 0.5 + g are carried in fixed point: a 64-bit accumulator with 40 fractional
// bits walks the thread's strips (exact integer stepping down the frame), the 8 rays of a strip use
// 32-bit values with 24 fractional bits (F' = F0 + j*C).  floor(u) is the rounded pixel unless u is
// within the guard g = 2^-19 px of a rounding tie; total fixed-point + composite-vs-step-wise error is
// < 10 * 2^-24 px << g, so a strip with any ray inside the guard window -- or any coordinate outside
// [0, 64) -- is redone with the step-wise fp64 chain (ray_to_pixel) and indices ALWAYS equal the
// step-wise definition.  Per ray: 2 IADD, 4 ops of guard (LOP, LOP, MIN, SETP.OR), 3 ops of key,
// compare + FADD for the run merge -- no fp64 in the steady state.  Rays outside the output grid land
// in tile cells that the flush discards (the reference's inplace_sum bounds check).
constexpr int kFrac = 24;                     // fractional bits of the per-strip 32-bit coordinates
constexpr int kQFrac = 40;                    // fractional bits of the 64-bit strip accumulators
constexpr int kGuardLog = 19;                 // guard g = 2^-19 px; window [0, 2g) on F' = F + g
constexpr unsigned kFracMask = (1u << kFrac) - 1u;
constexpr unsigned kWindow = 1u << (kFrac - kGuardLog + 1);

template <typename T>
__global__ void __launch_bounds__(kThreads4d, 4)
    stem4d_backproject_dda_kernel(const __grid_constant__ Stem4dGeom g, const T *__restrict__ data,
                                  float *__restrict__ out, int s_begin) {
  __shared__ float tile[kTile * kTile];
  const int s = s_begin + blockIdx.x;
  const int sy = s / g.Sx, sx = s % g.Sx;
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) tile[k] = 0.f;
  const double fsy = (double)sy, fsx = (double)sx;
  const double spy = (g.Ts[0] * fsy + g.Ts[1] * fsx) + g.Ts[2];
  const double spx = (g.Ts[3] * fsy + g.Ts[4] * fsx) + g.Ts[5];
  const double edx = (g.edet[0] + spx * g.edet[2]) + spy * g.edet[4];
  const double edy = (g.edet[1] + spx * g.edet[3]) + spy * g.edet[5];
  const double esx = (g.esamp[0] + spx * g.esamp[2]) + spy * g.esamp[4];
  const double esy = (g.esamp[1] + spx * g.esamp[3]) + spy * g.esamp[5];
  int cy, cx;
  ray_to_pixel(g, spx, spy, edx, edy, esx, esy, g.Dy / 2, g.Dx / 2, cy, cx);
  const int ty0 = cy - kTile / 2, tx0 = cx - kTile / 2;

  const int gpr = g.Dx >> 3;                   // strips (groups of 8 pixels) per detector row
  const int ngroups = g.Dy * gpr;
  // 64-bit fixed-point strip origin for this thread's first strip, and its strides
  long long qy, qx, stepy, stepx, wrapy, wrapx;
  int cy32, cx32;
  // each warp streams its own contiguous chunk of the frame (see the single-crossing kernel below)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = (((ngroups + (kThreads4d / 32) - 1) / (kThreads4d / 32)) + 31) & ~31;
  const int grp_end = min(ngroups, (warp + 1) * chunk);
  int grp = warp * chunk + lane;
  int dy = grp / gpr, cg = grp - dy * gpr;
  const int drow = 32 / gpr, dcol = 32 - drow * gpr;
  {
    const double yd = g.Td[2], xd = g.Td[5];
    const double rx = (xd - g.cdet[0]) - edx, ry = (yd - g.cdet[1]) - edy;
    const double tx = g.Binv[0] * rx + g.Binv[1] * ry, ty = g.Binv[2] * rx + g.Binv[3] * ry;
    const double xs = (g.csamp[0] + (g.Bs[0] * tx + g.Bs[1] * ty)) + esx;
    const double ys = (g.csamp[1] + (g.Bs[2] * tx + g.Bs[3] * ty)) + esy;
    const double by = (g.To[0] * ys + g.To[1] * xs) + g.To[2];
    const double bx = (g.To[3] * ys + g.To[4] * xs) + g.To[5];
    double ry_, rx_, cy_, cx_;
    auto slope = [&](double dyd, double dxd, double &oy, double &ox) {
      const double ttx = g.Binv[0] * dxd + g.Binv[1] * dyd, tty = g.Binv[2] * dxd + g.Binv[3] * dyd;
      const double dxs = g.Bs[0] * ttx + g.Bs[1] * tty, dys = g.Bs[2] * ttx + g.Bs[3] * tty;
      oy = g.To[0] * dys + g.To[1] * dxs;
      ox = g.To[3] * dys + g.To[4] * dxs;
    };
    slope(g.Td[0], g.Td[3], ry_, rx_);
    slope(g.Td[1], g.Td[4], cy_, cx_);
    const double kS = 1099511627776.0;         // 2^40
    const double off = 0.5 + 1.0 / (double)(1 << kGuardLog);
    const long long By = __double2ll_rn(((by - (double)ty0) + off) * kS);
    const long long Bx = __double2ll_rn(((bx - (double)tx0) + off) * kS);
    const long long Ry = __double2ll_rn(ry_ * kS), Rx = __double2ll_rn(rx_ * kS);
    const long long Cy = __double2ll_rn(cy_ * kS), Cx = __double2ll_rn(cx_ * kS);
    qy = By + (long long)dy * Ry + (long long)(cg * 8) * Cy;
    qx = Bx + (long long)dy * Rx + (long long)(cg * 8) * Cx;
    stepy = (long long)drow * Ry + (long long)(dcol * 8) * Cy;
    stepx = (long long)drow * Rx + (long long)(dcol * 8) * Cx;
    wrapy = Ry - (long long)(gpr * 8) * Cy;
    wrapx = Rx - (long long)(gpr * 8) * Cx;
    cy32 = (int)((Cy + (1ll << (kQFrac - kFrac - 1))) >> (kQFrac - kFrac));
    cx32 = (int)((Cx + (1ll << (kQFrac - kFrac - 1))) >> (kQFrac - kFrac));
  }
  __syncthreads();

  // one strip of 8 rays: fixed-point keys + guard, run merge into the tile; then advance to this
  // thread's next strip (exact integer stepping)
  auto strip = [&](const float (&v)[8]) {
    const int fy0 = (int)(qy >> (kQFrac - kFrac)), fx0 = (int)(qx >> (kQFrac - kFrac));
    int key[8];
    unsigned near = 0xffffffffu, range = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int fy = fy0 + j * cy32, fx = fx0 + j * cx32;
      near = min(near, min((unsigned)fy & kFracMask, (unsigned)fx & kFracMask));
      if (j == 0 || j == 7) range |= (unsigned)fy | (unsigned)fx;
      key[j] = ((fy >> kFrac) << 6) + (fx >> kFrac);
    }
    if (near >= kWindow && (range >> (kFrac + 6)) == 0u) {
      int cur = key[0];
      float sum = v[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        if (key[j] != cur) {
          if (sum != 0.f) atomicAdd(&tile[cur], sum);
          cur = key[j];
          sum = v[j];
        } else {
          sum += v[j];
        }
      }
      if (sum != 0.f) atomicAdd(&tile[cur], sum);
    } else {                                   // a tie / out-of-tile strip: exact step-wise chain
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        float vj = v[0];                        // register select (no dynamic indexing -> no stack)
#pragma unroll
        for (int t = 1; t < 8; ++t) vj = (j == t) ? v[t] : vj;
        int py, px;
        ray_to_pixel(g, spx, spy, edx, edy, esx, esy, dy, cg * 8 + j, py, px);
        if ((unsigned)py >= (unsigned)g.Oy || (unsigned)px >= (unsigned)g.Ox) continue;
        const int ly = py - ty0, lx = px - tx0;
        if ((unsigned)ly < (unsigned)kTile && (unsigned)lx < (unsigned)kTile)
          atomicAdd(&tile[ly * kTile + lx], vj);
        else
          atomicAdd(&out[(long long)py * g.Ox + px], vj);
      }
    }
    dy += drow; cg += dcol; qy += stepy; qx += stepx;
    if (cg >= gpr) { cg -= gpr; ++dy; qy += wrapy; qx += wrapx; }
  };

  const T *frame = data + (long long)s * g.Dy * g.Dx;
  float va[8], vb[8];                          // ping-pong: the next strip's loads are always in flight
  if (grp < grp_end) load8<T>(frame + (long long)grp * 8, va);
  while (grp < grp_end) {
    if (grp + 32 < grp_end) load8<T>(frame + (long long)(grp + 32) * 8, vb);
    strip(va);
    grp += 32;
    if (grp >= grp_end) break;
    if (grp + 32 < grp_end) load8<T>(frame + (long long)(grp + 32) * 8, va);
    strip(vb);
    grp += 32;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) {
    const float t = tile[k];
    if (t != 0.f) {
      const int py = ty0 + k / kTile, px = tx0 + k % kTile;
      if (py >= 0 && py < g.Oy && px >= 0 && px < g.Ox) atomicAdd(&out[(long long)py * g.Ox + px], t);
    }
  }
}

// ---- integer DDA, single-crossing strips ---------------------------------------------------
// When additionally 7 |slope| < 1 px per detector column for both coordinates (checked on the host),
// a strip of 8 rays crosses at most ONE pixel boundary per coordinate, so its rays fall into at most
// three cells: (y, x) not crossed / one crossed / both crossed.  Each coordinate is carried MIRRORED
// if it decreases along the strip (U = 63 - u), so both always increase; t = frac(U_0) - 1 + j |c| is
// negative before the crossing and >= 0 after it (the class predicate is a sign test), and the
// smallest non-negative t is the only candidate for a near-tie apart from frac(U_0) itself.  The strip
// is summed branch-free into three class sums with predicated FADDs and flushed with at most three
// shared-memory atomics: no divergent run-merge loop (which cost 2/3 of the general DDA kernel's
// issue slots).  The tile is indexed in mirrored cells; the final flush un-mirrors.
__device__ __forceinline__ void smem_red_add(unsigned addr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// 3 CTAs/SM (80 registers): measured 3.87 ms vs 4.16 ms at 4 CTAs/SM (64 registers force the compiler
// to rematerialise per-CTA constants inside the strip loop); the kernel is ALU-pipe-bound (77 %).
template <typename T>
__global__ void __launch_bounds__(kThreads4d, 3)
    stem4d_backproject_dda1x_kernel(const __grid_constant__ Stem4dGeom g, const T *__restrict__ data,
                                    float *__restrict__ out, int s_begin) {
  __shared__ float tile[kTile * kTile];
  const int s = s_begin + blockIdx.x;
  const int sy = s / g.Sx, sx = s % g.Sx;
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) tile[k] = 0.f;
  const double fsy = (double)sy, fsx = (double)sx;
  const double spy = (g.Ts[0] * fsy + g.Ts[1] * fsx) + g.Ts[2];
  const double spx = (g.Ts[3] * fsy + g.Ts[4] * fsx) + g.Ts[5];
  const double edx = (g.edet[0] + spx * g.edet[2]) + spy * g.edet[4];
  const double edy = (g.edet[1] + spx * g.edet[3]) + spy * g.edet[5];
  const double esx = (g.esamp[0] + spx * g.esamp[2]) + spy * g.esamp[4];
  const double esy = (g.esamp[1] + spx * g.esamp[3]) + spy * g.esamp[5];
  int cy, cx;
  ray_to_pixel(g, spx, spy, edx, edy, esx, esy, g.Dy / 2, g.Dx / 2, cy, cx);
  const int ty0 = cy - kTile / 2, tx0 = cx - kTile / 2;
  const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);

  const int gpr = g.Dx >> 3;
  const int ngroups = g.Dy * gpr;
  long long qy, qx, stepy, stepx, wrapy, wrapx;
  int cy32, cx32;            // |slope| per detector column, 24 fractional bits (>= 0)
  bool my, mx;               // coordinate is carried mirrored
  // Each warp streams its own contiguous chunk of the frame (1/8 of the strips, a multiple of 32):
  // the 8 warps of the CTA then work ~Dy/8 detector rows apart and update different tile cells, so
  // the shared-memory CAS loops of concurrent warps do not collide (interleaved rows made every
  // warp hit the same ~30 cells: 2.9 CAS iterations per flush, measured).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = (((ngroups + (kThreads4d / 32) - 1) / (kThreads4d / 32)) + 31) & ~31;
  const int grp_end = min(ngroups, (warp + 1) * chunk);
  int grp = warp * chunk + lane;
  int dy = grp / gpr, cg = grp - dy * gpr;
  const int drow = 32 / gpr, dcol = 32 - drow * gpr;
  {
    const double yd = g.Td[2], xd = g.Td[5];
    const double rx = (xd - g.cdet[0]) - edx, ry = (yd - g.cdet[1]) - edy;
    const double tx = g.Binv[0] * rx + g.Binv[1] * ry, ty = g.Binv[2] * rx + g.Binv[3] * ry;
    const double xs = (g.csamp[0] + (g.Bs[0] * tx + g.Bs[1] * ty)) + esx;
    const double ys = (g.csamp[1] + (g.Bs[2] * tx + g.Bs[3] * ty)) + esy;
    const double by = (g.To[0] * ys + g.To[1] * xs) + g.To[2];
    const double bx = (g.To[3] * ys + g.To[4] * xs) + g.To[5];
    double ry_, rx_, cy_, cx_;
    auto slope = [&](double dyd, double dxd, double &oy, double &ox) {
      const double ttx = g.Binv[0] * dxd + g.Binv[1] * dyd, tty = g.Binv[2] * dxd + g.Binv[3] * dyd;
      const double dxs = g.Bs[0] * ttx + g.Bs[1] * tty, dys = g.Bs[2] * ttx + g.Bs[3] * tty;
      oy = g.To[0] * dys + g.To[1] * dxs;
      ox = g.To[3] * dys + g.To[4] * dxs;
    };
    slope(g.Td[0], g.Td[3], ry_, rx_);
    slope(g.Td[1], g.Td[4], cy_, cx_);
    my = cy_ < 0.0;
    mx = cx_ < 0.0;
    const double kS = 1099511627776.0;         // 2^40
    const double off = 0.5 + 1.0 / (double)(1 << kGuardLog);
    const double uy = by - (double)ty0, ux = bx - (double)tx0;
    const long long By = __double2ll_rn(((my ? (double)(kTile - 1) - uy : uy) + off) * kS);
    const long long Bx = __double2ll_rn(((mx ? (double)(kTile - 1) - ux : ux) + off) * kS);
    const long long Ry = __double2ll_rn((my ? -ry_ : ry_) * kS), Rx = __double2ll_rn((mx ? -rx_ : rx_) * kS);
    const long long Cy = __double2ll_rn(fabs(cy_) * kS), Cx = __double2ll_rn(fabs(cx_) * kS);
    qy = By + (long long)dy * Ry + (long long)(cg * 8) * Cy;
    qx = Bx + (long long)dy * Rx + (long long)(cg * 8) * Cx;
    stepy = (long long)drow * Ry + (long long)(dcol * 8) * Cy;
    stepx = (long long)drow * Rx + (long long)(dcol * 8) * Cx;
    wrapy = Ry - (long long)(gpr * 8) * Cy;
    wrapx = Rx - (long long)(gpr * 8) * Cx;
    cy32 = (int)((Cy + (1ll << (kQFrac - kFrac - 1))) >> (kQFrac - kFrac));
    cx32 = (int)((Cx + (1ll << (kQFrac - kFrac - 1))) >> (kQFrac - kFrac));
  }
  __syncthreads();

  auto strip = [&](const float (&v)[8]) {
    const int wy0 = (int)(qy >> (kQFrac - kFrac)), wx0 = (int)(qx >> (kQFrac - kFrac));
    const int ky = wy0 >> kFrac, kx = wx0 >> kFrac;                   // mirrored cell of ray 0
    const unsigned fy = (unsigned)wy0 & kFracMask, fx = (unsigned)wx0 & kFracMask;
    int ty = (int)fy - (1 << kFrac), tx = (int)fx - (1 << kFrac);     // < 0 until the crossing
    int yonly = 0;                             // some ray has crossed in y but not (yet) in x
    unsigned near = min(fy, fx);
    float s00 = 0.f, s11 = 0.f, smid = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // class selection by sign masks on the float's bits (branch-free: 2 SHF, 3 LOP3, 3 FADD)
      const int ny = ty >> 31, nx = tx >> 31;            // all ones until the coordinate has crossed
      const int bits = __float_as_int(v[j]);
      near = min(near, min((unsigned)ty, (unsigned)tx));
      s00 += __int_as_float(bits & ny & nx);
      s11 += __int_as_float(bits & ~ny & ~nx);
      smid += __int_as_float(bits & (ny ^ nx));
      yonly |= ~ny & nx;
      ty += cy32;
      tx += cx32;
    }
    if (near >= kWindow && (unsigned)(ky - 1) < (unsigned)(kTile - 2) && (unsigned)(kx - 1) < (unsigned)(kTile - 2)) {
      const unsigned a00 = tile_s + 4u * (unsigned)(ky * kTile + kx);
      const bool yfirst = yonly != 0;          // the middle class is (y+1, x) rather than (y, x+1)
      if (s00 != 0.f) smem_red_add(a00, s00);
      if (smid != 0.f) smem_red_add(a00 + (yfirst ? 4u * kTile : 4u), smid);
      if (s11 != 0.f) smem_red_add(a00 + 4u * (kTile + 1), s11);
    } else {                                   // a tie / out-of-tile strip: exact step-wise chain
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        float vj = v[0];
#pragma unroll
        for (int t = 1; t < 8; ++t) vj = (j == t) ? v[t] : vj;
        int py, px;
        ray_to_pixel(g, spx, spy, edx, edy, esx, esy, dy, cg * 8 + j, py, px);
        if ((unsigned)py >= (unsigned)g.Oy || (unsigned)px >= (unsigned)g.Ox) continue;
        const int ly = py - ty0, lx = px - tx0;
        if ((unsigned)ly < (unsigned)kTile && (unsigned)lx < (unsigned)kTile)
          atomicAdd(&tile[(my ? kTile - 1 - ly : ly) * kTile + (mx ? kTile - 1 - lx : lx)], vj);
        else
          atomicAdd(&out[(long long)py * g.Ox + px], vj);
      }
    }
    dy += drow; cg += dcol; qy += stepy; qx += stepx;
    if (cg >= gpr) { cg -= gpr; ++dy; qy += wrapy; qx += wrapx; }
  };

  const T *frame = data + (long long)s * g.Dy * g.Dx;
  float va[8], vb[8];
  if (grp < grp_end) load8<T>(frame + (long long)grp * 8, va);
  while (grp < grp_end) {
    if (grp + 32 < grp_end) load8<T>(frame + (long long)(grp + 32) * 8, vb);
    strip(va);
    grp += 32;
    if (grp >= grp_end) break;
    if (grp + 32 < grp_end) load8<T>(frame + (long long)(grp + 32) * 8, va);
    strip(vb);
    grp += 32;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kTile * kTile; k += kThreads4d) {
    const float t = tile[k];
    if (t != 0.f) {
      const int ly = k / kTile, lx = k % kTile;
      const int py = ty0 + (my ? kTile - 1 - ly : ly), px = tx0 + (mx ? kTile - 1 - lx : lx);
      if (py >= 0 && py < g.Oy && px >= 0 && px < g.Ox) atomicAdd(&out[(long long)py * g.Ox + px], t);
    }
  }
}

// host-side dispatch test: does every frame's footprint stay inside the tile?  (The linear part of
// the detector-pixel -> sample-pixel map does not depend on the scan position.)
int footprint_fits_tile(const Stem4dGeom &g) {
  auto slope = [&](double dyd, double dxd, double &oy, double &ox) {
    const double ttx = g.Binv[0] * dxd + g.Binv[1] * dyd, tty = g.Binv[2] * dxd + g.Binv[3] * dyd;
    const double dxs = g.Bs[0] * ttx + g.Bs[1] * tty, dys = g.Bs[2] * ttx + g.Bs[3] * tty;
    oy = g.To[0] * dys + g.To[1] * dxs;
    ox = g.To[3] * dys + g.To[4] * dxs;
  };
  double ry, rx, cy, cx;
  slope(g.Td[0], g.Td[3], ry, rx);
  slope(g.Td[1], g.Td[4], cy, cx);
  const double hy = 0.5 * g.Dy + 1.0, hx = 0.5 * g.Dx + 1.0;
  const double ey = fabs(ry) * hy + fabs(cy) * hx, ex = fabs(rx) * hy + fabs(cx) * hx;
  if (!(ey <= kTile / 2 - 3 && ex <= kTile / 2 - 3)) return 0;   // (NaN -> 0)
  // single-crossing strips: 7 |slope per column| < 1 with a margin for the fixed-point rounding
  return (7.0 * fabs(cy) < 0.999 && 7.0 * fabs(cx) < 0.999) ? 2 : 1;
}

// pixel indices only (parity / debugging): idx[(s*npix + p)*2 + {0,1}] = (py, px)
__global__ void __launch_bounds__(kThreads4d)
    stem4d_indices_kernel(const __grid_constant__ Stem4dGeom g, int32_t *__restrict__ idx, int s_begin) {
  const int s = s_begin + blockIdx.x;
  const int sy = s / g.Sx, sx = s % g.Sx;
  const double fsy = (double)sy, fsx = (double)sx;
  const double spy = (g.Ts[0] * fsy + g.Ts[1] * fsx) + g.Ts[2];
  const double spx = (g.Ts[3] * fsy + g.Ts[4] * fsx) + g.Ts[5];
  const double edx = (g.edet[0] + spx * g.edet[2]) + spy * g.edet[4];
  const double edy = (g.edet[1] + spx * g.edet[3]) + spy * g.edet[5];
  const double esx = (g.esamp[0] + spx * g.esamp[2]) + spy * g.esamp[4];
  const double esy = (g.esamp[1] + spx * g.esamp[3]) + spy * g.esamp[5];
  const long long npix = (long long)g.Dy * g.Dx;
  for (long long p = threadIdx.x; p < npix; p += kThreads4d) {
    int py, px;
    ray_to_pixel(g, spx, spy, edx, edy, esx, esy, (int)(p / g.Dx), (int)(p % g.Dx), py, px);
    idx[((long long)s * npix + p) * 2 + 0] = py;
    idx[((long long)s * npix + p) * 2 + 1] = px;
  }
}

int fill_geom(Stem4dGeom &g, const int shapes[6], const double *geom44) {
  g.Sy = shapes[0]; g.Sx = shapes[1]; g.Dy = shapes[2]; g.Dx = shapes[3]; g.Oy = shapes[4]; g.Ox = shapes[5];
  TG_REQUIRE(g.Sy > 0 && g.Sx > 0 && g.Dy > 0 && g.Dx > 0 && g.Oy > 0 && g.Ox > 0, "bad shapes");
  TG_REQUIRE(g.Oy <= (1 << 20) && g.Ox <= (1 << 20), "output grid larger than 2^20 pixels per side");
  const double *p = geom44;
  for (int i = 0; i < 6; ++i) g.Ts[i] = *p++;
  for (int i = 0; i < 6; ++i) g.Td[i] = *p++;
  for (int i = 0; i < 6; ++i) g.To[i] = *p++;
  for (int i = 0; i < 2; ++i) g.cdet[i] = *p++;
  for (int i = 0; i < 6; ++i) g.edet[i] = *p++;
  for (int i = 0; i < 4; ++i) g.Binv[i] = *p++;
  for (int i = 0; i < 2; ++i) g.csamp[i] = *p++;
  for (int i = 0; i < 4; ++i) g.Bs[i] = *p++;
  for (int i = 0; i < 6; ++i) g.esamp[i] = *p++;
  return TG_OK;
}

}  // namespace

extern "C" int tg_stem4d_backproject(const int shapes[6], const double geom[42], const void *data4d,
                                     int data_is_f32, int s_begin, int s_count, float *out, void *stream) {
  TG_REQUIRE(shapes && geom && data4d && out, "null pointer");
  Stem4dGeom g;
  int rc = fill_geom(g, shapes, geom);
  if (rc != TG_OK) return rc;
  TG_REQUIRE(s_begin >= 0 && s_count >= 0 && (long long)s_begin + s_count <= (long long)g.Sy * g.Sx, "bad scan range");
  if (s_count == 0) return TG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool force_stepwise = (data_is_f32 & 2) != 0;
  const bool no_dda = (data_is_f32 & 4) != 0;      // A/B switch: guarded fp64 affine kernel instead of the DDA
  const bool general_dda = (data_is_f32 & 8) != 0; // A/B switch: run-merging DDA even for single-crossing strips
  data_is_f32 &= 1;
  const size_t esz = data_is_f32 ? 4 : 2;
  const bool fast = !force_stepwise && (g.Dx % 8) == 0 && ((reinterpret_cast<uintptr_t>(data4d) +
                                         (size_t)s_begin * g.Dy * g.Dx * esz) % 16) == 0 &&
                    (((size_t)g.Dy * g.Dx * esz) % 16) == 0;
  const int fits = (fast && !no_dda) ? footprint_fits_tile(g) : 0;
  if (fits == 2 && !general_dda) {
    if (data_is_f32)
      stem4d_backproject_dda1x_kernel<float><<<(unsigned)s_count, kThreads4d, 0, st>>>(
          g, static_cast<const float *>(data4d), out, s_begin);
    else
      stem4d_backproject_dda1x_kernel<unsigned short><<<(unsigned)s_count, kThreads4d, 0, st>>>(
          g, static_cast<const unsigned short *>(data4d), out, s_begin);
    return tg_launch_check("stem4d_backproject_dda1x_kernel");
  }
  if (fits) {
    if (data_is_f32)
      stem4d_backproject_dda_kernel<float><<<(unsigned)s_count, kThreads4d, 0, st>>>(
          g, static_cast<const float *>(data4d), out, s_begin);
    else
      stem4d_backproject_dda_kernel<unsigned short><<<(unsigned)s_count, kThreads4d, 0, st>>>(
          g, static_cast<const unsigned short *>(data4d), out, s_begin);
    return tg_launch_check("stem4d_backproject_dda_kernel");
  }
  if (fast) {
    if (data_is_f32)
      stem4d_backproject_fast_kernel<float><<<(unsigned)s_count, kThreads4d, 0, st>>>(
          g, static_cast<const float *>(data4d), out, s_begin);
    else
      stem4d_backproject_fast_kernel<unsigned short><<<(unsigned)s_count, kThreads4d, 0, st>>>(
          g, static_cast<const unsigned short *>(data4d), out, s_begin);
    return tg_launch_check("stem4d_backproject_fast_kernel");
  }
  if (data_is_f32)
    stem4d_backproject_kernel<float><<<(unsigned)s_count, kThreads4d, 0, st>>>(g, static_cast<const float *>(data4d), out, s_begin);
  else
    stem4d_backproject_kernel<unsigned short><<<(unsigned)s_count, kThreads4d, 0, st>>>(
        g, static_cast<const unsigned short *>(data4d), out, s_begin);
  return tg_launch_check("stem4d_backproject_kernel");
}

extern "C" int tg_stem4d_indices(const int shapes[6], const double geom[42], int s_begin, int s_count,
                                 int32_t *idx, void *stream) {
  TG_REQUIRE(shapes && geom && idx, "null pointer");
  Stem4dGeom g;
  int rc = fill_geom(g, shapes, geom);
  if (rc != TG_OK) return rc;
  TG_REQUIRE(s_begin >= 0 && s_count >= 0 && (long long)s_begin + s_count <= (long long)g.Sy * g.Sx, "bad scan range");
  if (s_count == 0) return TG_OK;
  stem4d_indices_kernel<<<(unsigned)s_count, kThreads4d, 0, static_cast<cudaStream_t>(stream)>>>(g, idx, s_begin);
  return tg_launch_check("stem4d_indices_kernel");
}
