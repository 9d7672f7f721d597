// XLA FFI custom-call handlers over the C-ABI of libvajax_b200.so (BASELINE.json north_star: "calls a thin
// C-ABI library through a JAX FFI custom call").  Built only where jaxlib's headers exist
// (vajax_b200/ffi.py::build_ffi compiles this file with -I$(python -c "import jax; print(jax.ffi.include_dir())")
// and links it against libvajax_b200.so); the build image has no jaxlib, so this file is NOT part of the
// default build and the ctypes binding (vajax_b200/lib.py) is what the tests execute.
//
// Modelled on the reference's UMFPACK handler, vajax/sparse/umfpack_ffi.cpp:189-338 (operands
// (S32 indptr, S32 indices, F64 data, F64 b) -> F64 x; registration :531-551).  Differences: the pattern,
// the symbolic analysis and the pivot order live in a vjx handle, which travels as the int64 attribute
// "handle" (vjx_handle_id / vjx_handle_from_id), and everything runs on the CUDA stream XLA provides.
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/vajax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

struct Sizes {
  int n_total, n_external, n_vs, n_is, n_unknowns, n_aug, nse, n_lim;
};

ffi::Error Resolve(int64_t id, vjx_handle** h, Sizes* sz) {
  *h = vjx_handle_from_id((long long)id);
  if (!*h) return ffi::Error::InvalidArgument("vajax_b200: unknown or destroyed handle id");
  int out[8];
  if (vjx_sizes(*h, out) != VJX_OK) return ffi::Error::Internal(vjx_last_error());
  *sz = Sizes{out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]};
  return ffi::Error::Success();
}

ffi::Error Status(int rc) { return rc == VJX_OK ? ffi::Error::Success() : ffi::Error::Internal(vjx_last_error()); }

// Optional operands arrive as zero-length buffers (JAX has no None operand): NULL for the C-ABI.
const double* OrNull(const ffi::Buffer<ffi::F64>& b) { return b.element_count() == 0 ? nullptr : b.typed_data(); }

// ---- S1: linear_solve_fn(csr_data, b) -> x            (solver_factories.py:985-1004) --------------------
ffi::Error SolveImpl(cudaStream_t stream, int64_t handle, ffi::Buffer<ffi::F64> data, ffi::Buffer<ffi::F64> b,
                     ffi::ResultBuffer<ffi::F64> x) {
  vjx_handle* h;
  Sizes sz;
  if (auto e = Resolve(handle, &h, &sz); e.failure()) return e;
  if ((int)data.element_count() != sz.nse) return ffi::Error::InvalidArgument("csr_data length != nse");
  if ((int)b.element_count() != sz.n_aug || (int)x->element_count() != sz.n_aug)
    return ffi::Error::InvalidArgument("rhs / solution length != n_aug");
  return Status(vjx_lu_solve(h, data.typed_data(), b.typed_data(), x->typed_data(), stream));
}

// every operand / result is checked against vjx_sizes before the C-ABI writes through it (a zero-length optional
// operand is None)
bool LenIs(const ffi::Buffer<ffi::F64>& b, int n) { return (int)b.element_count() == n; }
bool OptLenIs(const ffi::Buffer<ffi::F64>& b, int n) { return b.element_count() == 0 || (int)b.element_count() == n; }
// a result the library writes n elements into (nothing is written when n == 0, any length will do then)
bool ResLenIs(ffi::ResultBuffer<ffi::F64>& b, int n) { return n <= 0 || (int)b->element_count() == n; }

// ---- S2: build_system(...) -> (csr_data, f, Q, I_vsource, limit_state_out, max_res_contrib)
//          (mna_builder.py:528-550, result order :1010) ----------------------------------------------------
ffi::Error BuildSystemImpl(cudaStream_t stream, int64_t handle, int32_t nr_iteration, double integ_c0, double gmin,
                           double gshunt, double integ_c1, double integ_d1, double integ_c2, ffi::Buffer<ffi::F64> X,
                           ffi::Buffer<ffi::F64> vsrc, ffi::Buffer<ffi::F64> isrc, ffi::Buffer<ffi::F64> Q_prev,
                           ffi::Buffer<ffi::F64> dQdt_prev, ffi::Buffer<ffi::F64> Q_prev2, ffi::Buffer<ffi::F64> limit_in,
                           ffi::ResultBuffer<ffi::F64> csr, ffi::ResultBuffer<ffi::F64> f, ffi::ResultBuffer<ffi::F64> Q,
                           ffi::ResultBuffer<ffi::F64> Ivs, ffi::ResultBuffer<ffi::F64> limit_out,
                           ffi::ResultBuffer<ffi::F64> max_res) {
  vjx_handle* h;
  Sizes sz;
  if (auto e = Resolve(handle, &h, &sz); e.failure()) return e;
  if (!LenIs(X, sz.n_total + sz.n_vs)) return ffi::Error::InvalidArgument("X length != n_total + n_vsources");
  if (!OptLenIs(vsrc, sz.n_vs) || !OptLenIs(isrc, sz.n_is)) return ffi::Error::InvalidArgument("vsource_vals / isource_vals length");
  if (!LenIs(Q_prev, sz.n_unknowns) || !OptLenIs(dQdt_prev, sz.n_unknowns) || !OptLenIs(Q_prev2, sz.n_unknowns))
    return ffi::Error::InvalidArgument("Q_prev / dQdt_prev / Q_prev2 length != n_unknowns");
  if (!OptLenIs(limit_in, sz.n_lim)) return ffi::Error::InvalidArgument("limit_state_in length != total limit states");
  if (!ResLenIs(csr, sz.nse) || !ResLenIs(f, sz.n_aug)) return ffi::Error::InvalidArgument("result shapes: csr_data[nse], f[n_aug]");
  if (!ResLenIs(Q, sz.n_unknowns) || !ResLenIs(max_res, sz.n_unknowns) || !ResLenIs(Ivs, sz.n_vs) || !ResLenIs(limit_out, sz.n_lim))
    return ffi::Error::InvalidArgument("result shapes: Q[n_unknowns], I_vsource[n_vs], limit_state_out[n_lim], max_res_contrib[n_unknowns]");
  return Status(vjx_build_system(h, X.typed_data(), OrNull(vsrc), OrNull(isrc), Q_prev.typed_data(), integ_c0, gmin, gshunt,
                                 integ_c1, integ_d1, OrNull(dQdt_prev), integ_c2, OrNull(Q_prev2), OrNull(limit_in),
                                 (int)nr_iteration, csr->typed_data(), f->typed_data(), Q->typed_data(), Ivs->typed_data(),
                                 limit_out->typed_data(), max_res->typed_data(), stream));
}

// ---- S3: nr_solve(...) -> (X, iterations, converged, max_f, Q, dQdt, I_vsource, limit_state, max_res_contrib)
//          (solver_factories.py:604-725).  Stream-ordered: vjx_nr_solve_async runs the NR loop as a CUDA-graph WHILE
//          node on XLA's stream and the DEVICE writes (iterations, converged, max_f, error_flags) into the F64[4]
//          result - no host read, no synchronisation inside the handler (the very first call on a handle analyses
//          the first Jacobian and builds the graph, which synchronises once). ------------------------------------
ffi::Error NrSolveImpl(cudaStream_t stream, int64_t handle, int32_t max_iterations, double integ_c0, double gmin,
                       double gshunt, double integ_c1, double integ_d1, double integ_c2, ffi::Buffer<ffi::F64> X_init,
                       ffi::Buffer<ffi::F64> vsrc, ffi::Buffer<ffi::F64> isrc, ffi::Buffer<ffi::F64> Q_prev,
                       ffi::Buffer<ffi::F64> dQdt_prev, ffi::Buffer<ffi::F64> Q_prev2, ffi::Buffer<ffi::F64> limit_in,
                       ffi::Buffer<ffi::F64> res_tol_floor, ffi::ResultBuffer<ffi::F64> X, ffi::ResultBuffer<ffi::F64> scalars,
                       ffi::ResultBuffer<ffi::F64> Q, ffi::ResultBuffer<ffi::F64> dQdt, ffi::ResultBuffer<ffi::F64> Ivs,
                       ffi::ResultBuffer<ffi::F64> limit_out, ffi::ResultBuffer<ffi::F64> max_res) {
  vjx_handle* h;
  Sizes sz;
  if (auto e = Resolve(handle, &h, &sz); e.failure()) return e;
  if (!LenIs(X_init, sz.n_total + sz.n_vs) || !ResLenIs(X, sz.n_total + sz.n_vs))
    return ffi::Error::InvalidArgument("X length != n_total + n_vsources");
  if (!OptLenIs(vsrc, sz.n_vs) || !OptLenIs(isrc, sz.n_is)) return ffi::Error::InvalidArgument("vsource_vals / isource_vals length");
  if (!LenIs(Q_prev, sz.n_unknowns) || !OptLenIs(dQdt_prev, sz.n_unknowns) || !OptLenIs(Q_prev2, sz.n_unknowns) ||
      !OptLenIs(res_tol_floor, sz.n_unknowns))
    return ffi::Error::InvalidArgument("Q_prev / dQdt_prev / Q_prev2 / res_tol_floor length != n_unknowns");
  if (!OptLenIs(limit_in, sz.n_lim)) return ffi::Error::InvalidArgument("limit_state_in length != total limit states");
  if ((int)scalars->element_count() != 4) return ffi::Error::InvalidArgument("scalars result must be F64[4]");
  if (!ResLenIs(Q, sz.n_unknowns) || !ResLenIs(dQdt, sz.n_unknowns) || !ResLenIs(max_res, sz.n_unknowns) ||
      !ResLenIs(Ivs, sz.n_vs) || !ResLenIs(limit_out, sz.n_lim))
    return ffi::Error::InvalidArgument("result shapes: Q / dQdt / max_res_contrib[n_unknowns], I_vsource[n_vs], limit_state[n_lim]");
  return Status(vjx_nr_solve_async(h, X_init.typed_data(), OrNull(vsrc), OrNull(isrc), Q_prev.typed_data(), integ_c0, gmin,
                                   gshunt, integ_c1, integ_d1, OrNull(dQdt_prev), integ_c2, OrNull(Q_prev2), OrNull(limit_in),
                                   OrNull(res_tol_floor), (int)max_iterations, X->typed_data(), scalars->typed_data(),
                                   Q->typed_data(), dQdt->typed_data(), Ivs->typed_data(), limit_out->typed_data(),
                                   max_res->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(vjx_ffi_solve_f64, SolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<ffi::Buffer<ffi::F64>>()   // csr_data[nse]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // b[n_aug]
                                  .Ret<ffi::Buffer<ffi::F64>>()); // x[n_aug]

XLA_FFI_DEFINE_HANDLER_SYMBOL(vjx_ffi_build_system_f64, BuildSystemImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int32_t>("nr_iteration")
                                  .Attr<double>("integ_c0")
                                  .Attr<double>("gmin")
                                  .Attr<double>("gshunt")
                                  .Attr<double>("integ_c1")
                                  .Attr<double>("integ_d1")
                                  .Attr<double>("integ_c2")
                                  .Arg<ffi::Buffer<ffi::F64>>()   // X
                                  .Arg<ffi::Buffer<ffi::F64>>()   // vsource_vals
                                  .Arg<ffi::Buffer<ffi::F64>>()   // isource_vals
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Q_prev
                                  .Arg<ffi::Buffer<ffi::F64>>()   // dQdt_prev   (zero length = None)
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Q_prev2     (zero length = None)
                                  .Arg<ffi::Buffer<ffi::F64>>()   // limit_state_in (zero length = None)
                                  .Ret<ffi::Buffer<ffi::F64>>()   // csr_data
                                  .Ret<ffi::Buffer<ffi::F64>>()   // f
                                  .Ret<ffi::Buffer<ffi::F64>>()   // Q
                                  .Ret<ffi::Buffer<ffi::F64>>()   // I_vsource
                                  .Ret<ffi::Buffer<ffi::F64>>()   // limit_state_out
                                  .Ret<ffi::Buffer<ffi::F64>>()); // max_res_contrib

XLA_FFI_DEFINE_HANDLER_SYMBOL(vjx_ffi_nr_solve_f64, NrSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int32_t>("max_iterations")
                                  .Attr<double>("integ_c0")
                                  .Attr<double>("gmin")
                                  .Attr<double>("gshunt")
                                  .Attr<double>("integ_c1")
                                  .Attr<double>("integ_d1")
                                  .Attr<double>("integ_c2")
                                  .Arg<ffi::Buffer<ffi::F64>>()   // X_init
                                  .Arg<ffi::Buffer<ffi::F64>>()   // vsource_vals
                                  .Arg<ffi::Buffer<ffi::F64>>()   // isource_vals
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Q_prev
                                  .Arg<ffi::Buffer<ffi::F64>>()   // dQdt_prev
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Q_prev2
                                  .Arg<ffi::Buffer<ffi::F64>>()   // limit_state_in
                                  .Arg<ffi::Buffer<ffi::F64>>()   // res_tol_floor
                                  .Ret<ffi::Buffer<ffi::F64>>()   // X
                                  .Ret<ffi::Buffer<ffi::F64>>()   // (iterations, converged, max_f, error_flags)
                                  .Ret<ffi::Buffer<ffi::F64>>()   // Q
                                  .Ret<ffi::Buffer<ffi::F64>>()   // dQdt
                                  .Ret<ffi::Buffer<ffi::F64>>()   // I_vsource
                                  .Ret<ffi::Buffer<ffi::F64>>()   // limit_state_out
                                  .Ret<ffi::Buffer<ffi::F64>>()); // max_res_contrib
