/*
 * temgym_b200.h -- C ABI of the B200-native TemGymCore hot path.
 *
 * The reference (TemGym/TemGymCore, pure Python on JAX) has no plugin / FFI layer;
 * its hot path is reached through Python calls.  Each entry point below states the
 * reference call it replaces (file:line relative to the reference repository).
 * The Python package `temgymcore_b200` binds these symbols with ctypes; a JAX user
 * would bind the same symbols through jax.ffi.
 *
 * Conventions
 *  - plain C types only; no torch / XLA types.
 *  - `*_f64` / `tg_field_*` entry points take DEVICE pointers owned by the caller and
 *    enqueue work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream).  They are asynchronous and retain no pointer after return.
 *  - `*_host` entry points take HOST pointers (pinned or pageable), do the H2D copy,
 *    the kernels and the D2H copy themselves, and return after the result is in the
 *    host buffers.
 *  - return value 0 = ok, negative = error (TG_E*); tg_last_error() gives a
 *    thread-local message.  No CPU fallback exists: without a CUDA device every
 *    compute entry point returns TG_ECUDA.
 *  - state vector order everywhere: [x, y, dx, dy, z, pathlength, _one]
 *    (reference src/temgym_core/ray.py:39-45).
 */
#ifndef TEMGYM_B200_H
#define TEMGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_ABI_VERSION 1

/* error codes */
#define TG_OK 0
#define TG_EINVAL (-1)        /* bad argument */
#define TG_ECUDA (-2)         /* CUDA runtime error (message in tg_last_error) */
#define TG_ENOTSEPARABLE (-3) /* tg_field_sum_separable: a beamlet has an xy cross term */
#define TG_EUNSUPPORTED (-4)

/* ---- model descriptor --------------------------------------------------- */

#define TG_MAX_COMPS 24
#define TG_NPARAM 48

/* opcodes: one per reference component class on the path */
enum tg_op {
  TG_OP_PLANE = 0,     /* Plane / Detector / ScanGrid / Source: identity
                          (components.py:133-134, 248-249, 405-406; source.py:21-34) */
  TG_OP_LENS = 1,      /* Lens                (components.py:161-174)  p[0]=focal_length */
  TG_OP_DEFLECTOR = 2, /* Deflector           (components.py:476-482)  p[0]=def_x p[1]=def_y */
  TG_OP_BIPRISM = 3,   /* Biprism             (components.py:553-559)  p[0]=def_x */
  TG_OP_KRIVANEK = 4,  /* AberratedLensKrivanek (components.py:192-215; aberrations.py:42-108)
                          p[0]=focal_length, p[1..25]=KrivanekCoeffs in field order,
                          p[26+2t], p[27+2t] = cos(m phi0), sin(m phi0) of harmonic term t in the
                          order C12 C21 C23 C32 C34 C41 C43 C45 C52 C54 C56 */
  TG_OP_OFFSET = 5,    /* Scanner / Descanner (components.py:279-285, 343-372)
                          p[0..3] = offsets added to x, y, dx, dy times _one */
  TG_OP_THICKLENS = 6, /* ThickLens           (components.py:431-452)  p[0]=focal_length,
                          p[1]=z_po - z_pi ; z = z_po */
  TG_OP_ROTATOR = 7    /* Rotator             (components.py:503-523)  p[0]=cos p[1]=sin */
};

/* flags */
#define TG_F_NOPROP 1 /* do not insert the free-space step before this component
                         (used to expose run_iter's individual steps, run.py:75-82) */
#define TG_F_DIST 2   /* `z` holds a fixed propagation DISTANCE instead of a plane position:
                         Propagator(distance, FreeSpaceParaxial) (propagator.py:22-40) */

typedef struct {
  int32_t op;
  int32_t flags;
  double z; /* component.z : free space of (z - ray.z) is inserted first, also when 0
               (run.py:76-80; propagator.py:52-72) */
  double p[TG_NPARAM];
} tg_comp;

typedef struct {
  int32_t n_comp;
  int32_t reserved;
  tg_comp comp[TG_MAX_COMPS];
} tg_model;

/* Ray input: SoA.  ptr[i] == NULL means "field i is the scalar value[i] for every ray"
 * (a reference Ray may mix Python floats and arrays, ray.py:58-77; source.py:73-79). */
typedef struct {
  const double *ptr[7];
  double value[7];
} tg_ray_in;

/* Per-ray component parameters: Scanner / Descanner (TG_OP_OFFSET) components whose parameters are ARRAYS over
 * the ray batch -- what the reference gets from jax.vmap over scan positions (components.py:252-372,
 * run.py:85-116).  ptr[k][j] (device, n doubles; NULL = keep the descriptor's scalar p[j]) replaces offset j
 * (added to x, y, dx, dy times _one) of model component comp[k]. */
#define TG_MAX_PERRAY 4
typedef struct {
  int32_t n;
  int32_t comp[TG_MAX_PERRAY];
  const double *ptr[TG_MAX_PERRAY][4];
} tg_perray;

/* Jacobian layouts for tg_trace_f64 */
#define TG_JAC_NONE 0   /* rays only                    == run_to_end (run.py:85-116) */
#define TG_JAC_ABCD5 1  /* jac = n*25 doubles, (n,5,5) row-major over [x,y,dx,dy,_one]
                           == vmap(jacobian(run_to_end)) + custom_jacobian_matrix
                           (gaussian.py:234-239; utils.py:7-43) */
#define TG_JAC_FULL7 2  /* jac = n*49 doubles, (n,7,7): d out_i / d in_j over all seven
                           Ray leaves == the full Ray-of-Ray pytree of jax.jacobian
                           (README.md:227-234) */

const char *tg_last_error(void);
int tg_abi_version(void);
/* number of visible CUDA devices, or TG_ECUDA */
int tg_device_count(void);

/* ---- K1: batched ray propagation + Jacobian ------------------------------ */
/* replaces run_to_end / run_iter steps / jax.jacobian(run_to_end) / solve_model steps
 * (run.py:47-116, 150-179).  out[i] may be NULL to skip field i. */
int tg_trace_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                 double *const out[7], double *jac, int jac_layout, void *stream);

/* tg_trace_f64 with per-ray Scanner / Descanner offsets (perray may be NULL) */
int tg_trace_perray_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in, const tg_perray *perray,
                        double *const out[7], double *jac, int jac_layout, void *stream);
/* host-buffer variant: pointers in `in`, `out`, `jac` are HOST memory. */
int tg_trace_f64_host(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                      double *const out[7], double *jac, int jac_layout, int device);

/* ---- parameter tangents (run_with_grads, run.py:182-267) -------------------- */
/* Jacobian of the output ray w.r.t. up to TG_GRAD_LANES directions per launch.  A direction
 * ("lane") is seeded either on an input ray field (ray_lane[f] = lane, or -1) or on component
 * parameters: seed {comp, slot, lane, weight} adds `weight` to lane `lane` of the tangent of
 * parameter `slot` of component `comp`, slot 0 = z, slot k+1 = p[k] (so derived parameters are
 * seeded as linear combinations by the caller).  jac = n * 7 * TG_GRAD_LANES doubles,
 * jac[i][r][lane] = d out_r / d direction_lane, rows r in state-vector order. */
#define TG_GRAD_LANES 8
#define TG_MAX_SEEDS 32
typedef struct {
  int32_t comp, slot, lane, reserved;
  double weight;
} tg_seed;
int tg_trace_grad_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in,
                      const int32_t ray_lane[7], const tg_seed *seeds, int n_seeds,
                      double *const out[7], double *jac, void *stream);

/* ---- higher-order derivatives w.r.t. the input ray (calculate_derivatives, run.py:119-147) ---- */
/* What `order` nested jax.jacfwd(run_to_end, argnums=0) calls produce, as dense tensors in Ray
 * field order: d1 (n,7,7) = d out_f / d in_a, d2 (n,7,7,7) = d2 out_f / d in_a d in_b,
 * d3 (n,7,7,7,7); order in 1..3 (TG_EUNSUPPORTED beyond), tensors above `order` may be NULL.  One
 * kernel launch evaluates the model in hyper-dual arithmetic, one thread per sorted index tuple;
 * out[f] (may be NULL) receives the output ray like tg_trace_f64.  Device pointers, async on stream. */
int tg_trace_jets_f64(const tg_model *model_host, int64_t n, const tg_ray_in *in, int order,
                      double *const out[7], double *d1, double *d2, double *d3, void *stream);

/* W_krivanek / grad_W_krivanek (aberrations.py:51-108) on n slope pairs (device fp64): coeffs = the 25
 * KrivanekCoeffs fields in declaration order followed by 11 (cos, sin)(m * phi0) pairs in the order
 * (2,phi12) (1,phi21) (3,phi23) (2,phi32) (4,phi34) (1,phi41) (3,phi43) (5,phi45) (2,phi52) (4,phi54)
 * (6,phi56) -- the parameter block of an AberratedLensKrivanek component minus its focal length.
 * Any of W, dWx, dWy may be NULL. */
int tg_krivanek_f64(int64_t n, const double *alpha_x, const double *alpha_y, const double coeffs[47],
                    double *W, double *dWx, double *dWy, void *stream);

/* concentric_rings(num_points_approx, radius) (utils.py:117-175; the deterministic disc sampler behind
 * ParallelBeam / PointSource.make_rays, source.py:58-188) generated on the device: (y, x) of every point, ring
 * by ring, the running angle sums restarted exactly where the reference's multi_cumsum_inplace restarts them
 * (utils.py:46-80).  tg_concentric_rings_count gives the number of points N (host arithmetic only); y and x must
 * hold `capacity` >= N doubles. */
int64_t tg_concentric_rings_count(int64_t num_points_approx, double radius);
int tg_concentric_rings_f64(int64_t num_points_approx, double radius, int64_t capacity, double *y, double *x,
                            void *stream);
/* decompose_Q_inv(Q_inv, wavelength, eps) (gaussian.py:35-89): principal waists (larger first), radii of
 * curvature and orientation of n complex 2x2 Q_inv (device, (n,2,2) complex128 interleaved); wavelength is n
 * doubles, or one when wavelength_is_scalar. */
int tg_decompose_qinv_f64(int64_t n, const double *Q_inv, const double *wavelength, int wavelength_is_scalar,
                          double eps, double *waist1, double *waist2, double *radius1, double *radius2,
                          double *theta, void *stream);
/* fibonacci_spiral(nb_samples, radius, alpha) (utils.py:297-325) written to device arrays x, y: the
 * beamlet-centre sampler of the aperture / biprism examples, generated on the GPU. */
int tg_fibonacci_spiral_f64(int64_t n, double radius, double alpha, double *x, double *y, void *stream);

/* transfer_rays (transfer.py:6-54): out[n][k][i] = sum_j T[k][i][j] rays[n][j] for m <= 32
 * (already cumulative) 5x5 matrices given in HOST memory; rays (n,5), out (n,m,5) device fp64. */
#define TG_MAX_TRANSFER 32
int tg_transfer_rays_f64(int64_t n, const double *rays, int m, const double *matrices_host,
                         double *out, void *stream);

/* ---- fused 4D-STEM shadow-image backprojection (BASELINE config C5) -------- */
/* One ray per (scan position, detector pixel): detector pixel -> metres (Detector.pixels_to_metres,
 * grid.py:155-182) -> slopes at the point source -> sample plane (transfer_rays_pt_src,
 * transfer.py:57-123, with the descan-error 5th column affine in the scan position,
 * components.py:343-372) -> sample-grid pixel (metres_to_pixels, grid.py:120-153) -> bounds-checked
 * accumulation (inplace_sum, utils.py:83-114).  shapes = {Sy,Sx,Dy,Dx,Oy,Ox}; geom = 42 doubles:
 * Ts[6] scan px->m | Td[6] detector px->m | To[6] out grid m->px (rows: y-form then x-form,
 * value = (a*row + b*col) + c) | cdet[2]=Adet r0 | edet[6]=e0,e1,e2 (x,y each) | Binv[4] |
 * csamp[2] | Bsamp[4] | esamp[6].  data4d: device (Sy,Sx,Dy,Dx) float32 (data_is_f32 bit 0 set) or
 * uint16; bit 1 of data_is_f32 forces the plain step-wise kernel (the default fast kernel evaluates
 * the per-frame affine map with FMAs and re-evaluates step-wise near rounding ties, giving
 * identical pixel indices); scan positions
 * [s_begin, s_begin+s_count) are processed (multi-GPU shards); out: device (Oy,Ox) float32,
 * accumulated into (zero it first). */
int tg_stem4d_backproject(const int shapes[6], const double geom[42], const void *data4d,
                          int data_is_f32, int s_begin, int s_count, float *out, void *stream);
/* the (py, px) int32 pixel pairs of the same rays, idx[(s*Dy*Dx + p)*2 + {0,1}] (parity checks) */
int tg_stem4d_indices(const int shapes[6], const double geom[42], int s_begin, int s_count,
                      int32_t *idx, void *stream);

/* ---- K5: metres -> pixels ------------------------------------------------ */
/* replaces Grid.metres_to_pixels(cast=True) (grid.py:120-153) given the inverse 3x3
 * m2px (row-major, acting on [y_m, x_m, 1], grid.py:50-63).  Evaluation order
 * (m[0]*y + m[1]*x) + m[2] without FMA contraction; round half to even; saturating;
 * NaN -> 0.  With as_float != 0 writes doubles (cast=False) into py/px instead. */
int tg_metres_to_pixels(int64_t n, const double *x, const double *y, const double m2px[9],
                        void *py, void *px, int as_float, void *stream);
int tg_metres_to_pixels_host(int64_t n, const double *x, const double *y,
                             const double m2px[9], void *py, void *px, int as_float,
                             int device);

/* Grid.into_image (grid.py:231-280; utils.py:83-114): bounds-checked scatter-add of
 * 1 per ray into an int64 image of shape (H, W). */
int tg_into_image_i64(int64_t n, const int32_t *py, const int32_t *px, int H, int W,
                      long long *image, void *stream);

/* ---- K2: per-beamlet coefficient builder --------------------------------- */
/* Collapses _beam_field (gaussian.py:276-316) + Qinv_ABCD (gaussian.py:92-96) to a
 * complex quadratic in the observation point (x, y) [metres]:
 *   field_n(x,y) = exp(i * P_n),  P_n = c0 + c1 x + c2 y + c3 x^2 + c4 x y + c5 y^2
 * poly[n*12 + 2*j + {0,1}] = Re/Im c_j.  All arrays are DEVICE fp64; complex inputs
 * are interleaved (re,im).  Shapes as propagate_misaligned_gaussian_jax_scan
 * (gaussian.py:319-337): amp,phase_offset,k (nb,); Q1_inv (nb,2,2) complex;
 * A,B,C,D (nb,2,2); e,f,r1m,theta1m (nb,2). */
int tg_beamlet_coeffs_f64(int64_t nb, const double *amp, const double *phase_offset,
                          const double *Q1_inv, const double *A, const double *B,
                          const double *C, const double *D, const double *e,
                          const double *f, const double *r1m, const double *theta1m,
                          const double *k, double *poly, void *stream);

/* Same, but reads A,B,C,D,e,f from the (nb,5,5) ABCD array K1 wrote (gaussian.py:244-249). */
int tg_beamlet_coeffs_abcd_f64(int64_t nb, const double *amp, const double *phase_offset,
                               const double *Q1_inv, const double *abcd, const double *r1m_x,
                               const double *r1m_y, const double *th_x, const double *th_y,
                               const double *k, double *poly, void *stream);

/* _input_beam_field (gaussian.py:402-407) in the same 6-coefficient format. */
int tg_input_coeffs_f64(int64_t nb, const double *amp, const double *phase_offset,
                        const double *Q1_inv, const double *r1m, const double *theta1m,
                        const double *k, double *poly, void *stream);

/* GaussianRay.Q_inv (gaussian.py:138-177): (nb,2,2) complex from waists, radii, theta. */
int tg_gaussian_qinv_f64(int64_t nb, const double *waist_xy, const double *radii_xy,
                         const double *wavelength, const double *theta, double *Q_inv,
                         void *stream);

/* k = 2 pi / wavelength and phase_offset = k * pathlength (gaussian.py:253-255). */
int tg_wave_numbers_f64(int64_t nb, const double *wavelength, const double *pathlength,
                        double *k, double *phase_offset, void *stream);

/* ---- K3: field sum -------------------------------------------------------- */
/* replaces map_reduce(_beam_field_outer, add) (gaussian.py:325-332, 340-369).
 * Grid form: observation points are the pixel centres of an (H, W) grid with
 *   x_m = px2m[0] + px2m[1]*col + px2m[2]*row,  y_m = px2m[3] + px2m[4]*col + px2m[5]*row
 * (Grid.coords, grid.py:82-100).  Computes rows [row0, row0+nrows) and writes them to
 * out[(row-row0)*W + col] as complex64 (out_is_c128 == 0) or complex128.
 * cull_bits > 0 enables tile-level culling of beamlets whose envelope over the whole
 * tile is below 2^-cull_bits of the brightest beamlet peak; 0 = dense (every beamlet
 * evaluated on every pixel).  n_evals_out (host pointer, may be NULL) receives the
 * number of executed beamlet*pixel evaluations (synchronises the stream if non-NULL). */
int tg_field_sum_grid(int64_t nb, const double *poly, const double px2m[6], int H, int W,
                      int row0, int nrows, void *out, int out_is_c128, int cull_bits,
                      long long *n_evals_out, void *stream);

/* Arbitrary observation points r (npts,2) (x,y) [metres], device fp64: the r2 argument
 * of propagate_misaligned_gaussian_jax_scan (gaussian.py:319-337). */
int tg_field_sum_points(int64_t nb, const double *poly, int64_t npts, const double *r_xy,
                        void *out, int out_is_c128, void *stream);

/* Tensor-core path (tcgen05, 3xTF32) for separable beamlets: when no beamlet has a col*row
 * term on this pixel grid, exp(iP) = U(row) V(col) and the sum is the complex GEMM
 * F = U^T V with K = nb.  Returns TG_ENOTSEPARABLE otherwise (then use tg_field_sum_grid).
 * Synchronises `stream` once (separability verdict). */
int tg_field_sum_separable(int64_t nb, const double *poly, const double px2m[6], int H,
                           int W, int row0, int nrows, void *out, int out_is_c128,
                           void *stream);

/* Method dispatch for the grid field sum.  TG_METHOD_AUTO enqueues both paths with a DEVICE-side
 * separability verdict (no host synchronisation): the kernels of the path that does not apply
 * return immediately.  Exception: for more than 16 384 beamlets outside a stream capture the verdict
 * (8 bytes) is read back once and only the path that applies is enqueued -- the dead launches of the
 * other path cost more than the read-back there; under capture the device-side form is always used. */
#define TG_METHOD_AUTO 0   /* tensor-core path when the beamlets are separable, else SFU kernel */
#define TG_METHOD_SFU 1    /* tg_field_sum_grid */
#define TG_METHOD_TENSOR 2 /* tg_field_sum_separable (TG_ENOTSEPARABLE if it does not apply): fp16 x 3 */
#define TG_METHOD_TENSOR_TF32 3 /* the same GEMM with tf32 x 3 operands (fp32 exponent range, half the rate) */
#define TG_METHOD_TENSOR_4M 4   /* fp16 x 3 with four real multiplications per complex term (the real GEMM on
                                   interleaved re/im operands) */
#define TG_METHOD_TENSOR_3M 5   /* fp16 x 3 with three real multiplications per complex term (Gauss: 25 % less tensor
                                   work, tiles twice as large).  TG_METHOD_TENSOR / AUTO run the 4-multiplication
                                   form unless the environment sets TG_TENSOR_GAUSS=1 */
#define TG_METHOD_TENSOR_BINNED 6 /* tile-binned (block-sparse K) tensor-core sum for separable beamlets that each reach a
                                   small part of the detector (BASELINE C3): every 128-row x 64-column output tile only
                                   multiplies the beamlets whose bounding box {envelope >= brightest peak - cull_bits}
                                   meets it.  Needs cull_bits > 0.  TG_METHOD_AUTO picks it (outside stream capture, for
                                   more than 16384 beamlets or host-buffer calls) when the beamlets are separable and
                                   sparse; TG_EUNSUPPORTED when the operands would exceed the capacity limit, and under
                                   stream capture unless the same call ran once eagerly on this host thread before */
int tg_field_sum(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                 int nrows, void *out, int out_is_c128, int cull_bits, int method, void *stream);

/* The verdict TG_METHOD_AUTO reaches on the device for these beamlets on this grid, read back to the host
 * (synchronises `stream`; not inside a stream capture): *use_tensor = 1 -> dense tensor-core path (separable, and not
 * clearly more expensive than the culled SFU sum), 2 -> tile-binned tensor-core path (separable and sparse,
 * TG_METHOD_TENSOR_BINNED), 0 -> SFU kernel.  Plans (GaussianImagePlan, PeerImagePlan) call it once at
 * build time and capture the launches of that one path. */
int tg_field_sum_verdict(int64_t nb, const double *poly, const double px2m[6], int H, int W, int cull_bits,
                         int *use_tensor, void *stream);

/* D[M x N] = (A_hi + A_lo)[M x K] * (B_hi + B_lo)[N x K]^T, products hi*hi + hi*lo + lo*hi on the
 * tensor cores (kind::tf32, fp32 TMEM accumulation drained every 128 k), fp64 output with row pitch
 * ldd.  Operands: device fp32, row pitch ldk elements (multiple of 4), 16-byte aligned; hi parts
 * must be TF32-representable.  The GEMM engine of tg_field_sum_separable, exposed for testing. */
int tg_gemm_tf32x3(int M, int N, int K, const float *A_hi, const float *A_lo, const float *B_hi,
                   const float *B_lo, long long ldk, double *D, long long ldd, int accumulate,
                   void *stream);
/* The same with fp16 operands (device IEEE binary16, row pitch ldk elements, multiple of 8; hi and lo
 * parts fp16 values): kind::f16 at twice the TF32 rate.  This is the engine tg_field_sum_separable and
 * TG_METHOD_AUTO/TENSOR use; they pre-scale the factors on the device so that fp16's exponent range
 * suffices and undo the scaling in the fp64 epilogue. */
int tg_gemm_f16x3(int M, int N, int K, const void *A_hi, const void *A_lo, const void *B_hi,
                  const void *B_lo, long long ldk, double *D, long long ldd, int accumulate,
                  void *stream);

/* Complex product with THREE real multiplications per term (Gauss): D[m, c] (+)= sum_n U[m, n] V[c, n], D (M, N)
 * complex128 as interleaved doubles (row pitch ldd doubles).  Operands in the 3-product layout the field sum
 * builds for its tensor path: per group g of C = tg_gemm_chunk_k() terms, k-elements [3 C g, 3 C g + 3 C) of a row hold
 * Ur + Ui | Ur | Ui (A) resp. Vr | Vi - Vr | Vr + Vi (B), each as fp16 hi and lo parts; K3 = 3 C * groups.
 * Re = k1 - k3, Im = k1 + k2 with k1 = (Ur + Ui) Vr, k2 = Ur (Vi - Vr), k3 = Ui (Vr + Vi): 25 % less tensor work
 * than the 4-multiplication real GEMM (tg_gemm_f16x3 on the interleaved layout).  Exposed for testing. */
int tg_gemm_chunk_k(void); /* terms per group of the 3-product layout (= k-elements per TMEM drain): 256 */
int tg_cgemm3_f16x3(int M, int N, int K3, const void *A_hi, const void *A_lo, const void *B_hi, const void *B_lo,
                    long long ldk, double *D, long long ldd, int accumulate, void *stream);
/* The GEMM is one persistent kernel (one CTA per SM): whole 128 x 128 tiles per CTA while they fill waves of
 * `sms` CTAs, and a stream-K split of the remaining tiles (or of ALL tiles when the problem has fewer tiles
 * than SMs, e.g. the row shard of one rank of a multi-GPU field sum) so that every SM gets the same amount of
 * K.  This host-only call returns that decomposition for a shape (no device needed; used by the CPU tests):
 * units[6 i ..] = {cta, tile, chunk_begin, chunk_end, slot, nparts} (chunks of tg_gemm_chunk_k() k-elements),
 * sched_out[10] = {tiles_n, tiles, k_blocks, chunks, ctas, streamk_tiles, head_chunks, tail_chunks,
 * helper_chunks, max_parts}.  mode: 0 = whole tiles only, 1 = split when whole tiles would leave more than 30 % of
 * the machine idle (what the library does: plain split-K into floor(sms / tiles) equal k-ranges per tile when
 * 2 tiles <= sms -- then streamk_tiles == tiles and tail_chunks == 0 -- else the head / tail arrangement),
 * 2 = split whenever the tiles leave a partial wave, 3 = like 1 with the head / tail arrangement only,
 * -1 = the library's setting (environment TG_GEMM_STREAMK, default 1).  Returns the number of units (only
 * max_units are written). */
int tg_gemm_schedule(int M, int N, int K, int f16, int sms, int mode, int32_t *units, int max_units,
                     int32_t *sched_out);
/* The RAGGED decomposition the tile-binned sum (TG_METHOD_TENSOR_BINNED) builds on the device, mirrored on the host (no
 * device needed): tile t owns chunks_per_tile[t] accumulation chunks of one concatenated k axis, `sms` CTAs split that
 * axis evenly.  units[8 i..] = {cta, tile, chunk_begin, chunk_end, slot, nparts, scratch slot written, first CTA of the
 * tile}; readers[] = for every unit in turn the nparts scratch slots the tile's last arriver sums.  Returns the number
 * of units. */
int tg_gemm_schedule_ragged(int T, const int32_t *chunks_per_tile, int sms, int32_t *units, int max_units,
                            int32_t *readers, int max_readers);
/* Operand chunks (128 beamlet slots, 256 KiB of fp16 hi / lo row and column factors each) that the calling thread's
 * last eager TG_METHOD_TENSOR_BINNED sum built; -1 if there was none.  (Measurement aid: bench.py's HBM roofline.) */
int tg_binned_last_chunks(void);

/* ---- host-buffer field sum (make_gaussian_image end to end, gaussian.py:225-273) -- */
/* All pointers HOST.  rays[7] are the central rays (length nb each), waist_xy /
 * radii_xy (nb,2), wavelength / theta / amplitude (nb,).  Traces the rays through
 * `model`, builds Q_inv, the coefficients, sums the field on the (H,W) grid given by
 * px2m and writes rows [row0, row0+nrows) as (nrows,W) complex to out (the row range is what a
 * rank of a row-sharded multi-GPU job computes; row0=0,nrows=H gives the whole image). */
int tg_make_gaussian_image_host(const tg_model *model_host, int64_t nb,
                                const double *const rays[7], const double *amplitude,
                                const double *waist_xy, const double *radii_xy,
                                const double *wavelength, const double *theta,
                                const double px2m[6], int H, int W, int row0, int nrows,
                                void *out, int out_is_c128, int cull_bits, int method,
                                int device);

/* Same pipeline with DEVICE pointers, enqueued on `stream` (asynchronous): the whole of
 * make_gaussian_image (gaussian.py:225-273) in one call for device-resident inputs. */
int tg_make_gaussian_image_f64(const tg_model *model_host, int64_t nb, const double *const rays[7],
                               const double *amplitude, const double *waist_xy,
                               const double *radii_xy, const double *wavelength,
                               const double *theta, const double px2m[6], int H, int W, int row0,
                               int nrows, void *out, int out_is_c128, int cull_bits, int method,
                               void *stream);

/* ---- multi-GPU: row-sharded field sum over peer memory -----------------------------
 * One process per GPU.  Each rank owns a full (H, W) image buffer allocated with tg_peer_alloc and maps
 * the buffers of the other ranks of the node with tg_peer_open (CUDA IPC; the handles travel through
 * torch.distributed / any host channel).  tg_field_sum_peers computes this rank's detector rows
 * [row0, row0 + nrows) (the reference has no multi-device path; this shards gaussian.py:319-369 by
 * detector rows) and the kernel that produces the final values stores them into ALL `npeers` images at
 * that row offset -- its own and, through NVLink P2P stores, its peers' -- so the gather of the row
 * blocks is fused into the compute kernels instead of being a separate collective.  tg_peer_barrier is
 * the matching device-side barrier over peer memory (flag arrays of npeers uint64, one per rank, inside
 * peer allocations): after it completes on every rank's stream, every rank's image is complete. */
#define TG_MAX_PEERS 8
typedef struct tg_ipc_handle { unsigned char bytes[64]; } tg_ipc_handle;
/* cudaMalloc (zero-filled) + IPC handle of the allocation */
int tg_peer_alloc(uint64_t bytes, void **dptr, tg_ipc_handle *handle);
/* map another process's allocation into this process (peer access enabled lazily) */
int tg_peer_open(const tg_ipc_handle *handle, void **dptr);
int tg_peer_close(void *dptr);
int tg_peer_free(void *dptr);
/* images[p]: base of rank p's (H, W) image as seen from this process (images[self] is local). */
int tg_field_sum_peers(int64_t nb, const double *poly, const double px2m[6], int H, int W, int row0,
                       int nrows, void *const images[], int npeers, int self, int out_is_c128,
                       int cull_bits, int method, void *stream);
/* The whole of make_gaussian_image (gaussian.py:225-273: ray kernel + ABCD, coefficients, field sum) for this
 * rank's rows, with the peer stores of tg_field_sum_peers: the per-step call of a row-sharded multi-GPU job,
 * capturable into a CUDA graph (every rank builds the same table from the same inputs: no broadcast). */
int tg_make_gaussian_image_peers(const tg_model *model_host, int64_t nb, const double *const rays[7],
                                 const double *amplitude, const double *waist_xy, const double *radii_xy,
                                 const double *wavelength, const double *theta, const double px2m[6], int H,
                                 int W, int row0, int nrows, void *const images[], int npeers, int self,
                                 int out_is_c128, int cull_bits, int method, void *stream);
/* flags[p]: base of rank p's flag array (npeers uint64, zero-initialised) as seen from this process.
 * Signals `epoch` (strictly increasing per use) to every peer and waits until every peer has signalled it;
 * traps after TG_PEER_BARRIER_TIMEOUT_S seconds (environment, default 10) instead of hanging. */
int tg_peer_barrier(void *const flags[], int npeers, int self, uint64_t epoch, void *stream);
/* The same barrier with its epoch kept on the device: state = 2 uint64 in LOCAL device memory, zero-initialised,
 * {epoch counter, status}.  Every launch increments the counter and uses the new value, so the launch can be
 * captured into a CUDA graph and replayed (all ranks must run the same number of barriers).  A timeout
 * (timeout_s seconds; <= 0: TG_PEER_BARRIER_TIMEOUT_S, default 10) does not trap: status becomes 1 + the rank
 * that did not arrive and the kernel returns, leaving the context usable. */
int tg_peer_barrier_auto(void *const flags[], int npeers, int self, void *state, double timeout_s, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TEMGYM_B200_H */
