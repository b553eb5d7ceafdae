// XLA FFI custom-call handlers over the C ABI of include/temgym_b200.h -- the `jax.ffi` layer
// BASELINE.json's north_star asks for, so that jax.jit / jax.vmap / jax.jacobian keep working on the
// reference's tracing surface while the arithmetic runs in the sm_100a kernels.
//
// NOT BUILT IN THIS IMAGE: jax / jaxlib and the XLA FFI headers (xla/ffi/api/ffi.h) are not installed
// and cannot be (no network).  build_xla_shim.sh compiles this file when
// `python -c "import jax.ffi; print(jax.ffi.include_dir())"` works; temgymcore_b200/jax_ffi.py registers
// the targets.  Until then this file is documentation that compiles elsewhere, and the tested drop-in
// boundary is the C ABI + the Python mirror of the reference's call surface.
//
// Reference calls replaced (file:line in the reference repository):
//   tg_trace           run_to_end                        src/temgym_core/run.py:85-116
//   tg_trace_abcd      vmap(jacobian(run_to_end)) + custom_jacobian_matrix   gaussian.py:234-239, utils.py:7-43
//   tg_trace_full7     jax.jacobian(run_to_end)          README.md:227-234 (the JVP rule's Jacobian)
//   tg_field_sum_grid  propagate_misaligned_gaussian_jax_scan / map_reduce   gaussian.py:319-369
#if __has_include("xla/ffi/api/ffi.h")
#include <cstring>

#include "temgym_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status(int rc) {
  if (rc == TG_OK) return ffi::Error::Success();
  const char *msg = tg_last_error();
  return ffi::Error(rc == TG_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    msg ? msg : "temgym_b200 call failed");
}

// The model descriptor (tg_model, 9.6 KB of plain data) travels as a byte-string attribute: it is a
// compile-time constant of the jitted computation, exactly like the Python model tuple the reference
// closes over (run.py:87).
ffi::Error load_model(std::string_view bytes, tg_model *m) {
  if (bytes.size() != sizeof(tg_model))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "model attribute is not a tg_model");
  std::memcpy(m, bytes.data(), sizeof(tg_model));
  return ffi::Error::Success();
}

// rays: (7, N) fp64 SoA in Ray field order x, y, dx, dy, z, pathlength, _one
ffi::Error TraceImpl(cudaStream_t stream, std::string_view model, int64_t jac_layout,
                     ffi::Buffer<ffi::F64> rays, ffi::ResultBuffer<ffi::F64> out,
                     ffi::ResultBuffer<ffi::F64> jac) {
  tg_model m;
  if (auto e = load_model(model, &m); e.failure()) return e;
  const auto dims = rays.dimensions();
  if (dims.size() != 2 || dims[0] != 7) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "rays must be (7, N)");
  const int64_t n = dims[1];
  tg_ray_in in{};
  double *o[7];
  for (int f = 0; f < 7; ++f) {
    in.ptr[f] = rays.typed_data() + f * n;
    o[f] = out->typed_data() + f * n;
  }
  return status(tg_trace_f64(&m, n, &in, o, jac_layout == TG_JAC_NONE ? nullptr : jac->typed_data(),
                             (int)jac_layout, stream));
}

// poly: (nb, 12) fp64 coefficient table from tg_beamlet_coeffs_abcd_f64 / tg_input_coeffs_f64;
// out: (nrows, W) complex128
ffi::Error FieldSumImpl(cudaStream_t stream, ffi::Span<const double> px2m, int64_t H, int64_t W, int64_t row0,
                        int64_t nrows, int64_t cull_bits, int64_t method, ffi::Buffer<ffi::F64> poly,
                        ffi::ResultBuffer<ffi::C128> out) {
  if (px2m.size() != 6) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "px2m must hold 6 doubles");
  const int64_t nb = poly.dimensions()[0];
  return status(tg_field_sum(nb, poly.typed_data(), px2m.begin(), (int)H, (int)W, (int)row0, (int)nrows,
                             out->typed_data(), 1, (int)cull_bits, (int)method, stream));
}

// the six complex coefficients per beamlet from the (nb,5,5) ABCD array (gaussian.py:241-310);
// rays: the INPUT central rays as (7, nb) SoA (r1m = rows 0,1; theta1m = rows 2,3)
ffi::Error CoeffsImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> amp, ffi::Buffer<ffi::F64> p0,
                      ffi::Buffer<ffi::C128> q1inv, ffi::Buffer<ffi::F64> abcd, ffi::Buffer<ffi::F64> rays,
                      ffi::Buffer<ffi::F64> k, ffi::ResultBuffer<ffi::F64> poly) {
  const int64_t nb = amp.dimensions()[0];
  const double *r = rays.typed_data();
  return status(tg_beamlet_coeffs_abcd_f64(nb, amp.typed_data(), p0.typed_data(),
                                           reinterpret_cast<const double *>(q1inv.typed_data()), abcd.typed_data(),
                                           r + 0 * nb, r + 1 * nb, r + 2 * nb, r + 3 * nb, k.typed_data(),
                                           poly->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(TgTrace, TraceImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<std::string_view>("model")
                                  .Attr<int64_t>("jac_layout")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(TgFieldSum, FieldSumImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const double>>("px2m")
                                  .Attr<int64_t>("H")
                                  .Attr<int64_t>("W")
                                  .Attr<int64_t>("row0")
                                  .Attr<int64_t>("nrows")
                                  .Attr<int64_t>("cull_bits")
                                  .Attr<int64_t>("method")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::C128>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(TgBeamletCoeffs, CoeffsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::C128>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
#else
// XLA FFI headers not available: nothing to build (see the header comment).
#endif
