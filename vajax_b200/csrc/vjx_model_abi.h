// ABI between libvajax_b200.so and the per-model CUDA plugins emitted by
// vajax_b200/codegen/cuda_emitter.py.  One plugin = one Verilog-A model lowered from OpenVAF
// MIR to a single fp64 eval kernel (one thread per instance) plus one init kernel.
//
// The reference's counterpart of this interface is the generated Python
//   eval_fn(shared_params, device_params, shared_cache, device_cache, simparams,
//           limit_state_in, limit_funcs) -> 9-tuple
// (reference: openvaf_jax/codegen/function_builder.py:971-996) wrapped in jax.vmap
// (reference: vajax/analysis/openvaf_models.py:997).
#pragma once
#include <cuda_runtime.h>

// Device-resident control block written by the step/NR control kernels and read by the model
// kernels, so that a captured CUDA graph sees new coefficients without re-capture.
enum VjxCtl {
  VJX_CTL_C0 = 0,      // integ_c0
  VJX_CTL_C1 = 1,      // integ_c1
  VJX_CTL_D1 = 2,      // integ_d1
  VJX_CTL_C2 = 3,      // integ_c2
  VJX_CTL_GMIN = 4,    // $simparam("gmin") seen by the models (gmin_arg)
  VJX_CTL_GSHUNT = 5,  // gshunt
  VJX_CTL_ITER = 6,    // NR iteration counter (as double)
  VJX_CTL_DIAGREG = 7, // min_diag_reg (options.gmin captured at factory time)
  VJX_CTL_COUNT = 16
};

struct VjxEvalArgs {
  int n_dev;
  const double* V;           // node voltages, ground at index 0            [n_total]
  const int* vnode1;         // SoA (n_volt, n_dev): circuit node of V(hi,..)
  const int* vnode2;         // SoA (n_volt, n_dev): circuit node of V(..,lo) (0 = ground)
  const double* dev_params;  // SoA (n_dev_params, n_dev): varying non-voltage params
  const double* dev_cache;   // SoA (n_dev_cache, n_dev): varying init-cache columns
  const double* ctl;         // VjxCtl block (device)
  const double* limit_in;    // (n_dev, n_lim) row-major slice of limit_state_in, or null
  double* limit_out;         // (n_dev, n_lim) row-major slice of limit_state_out, or null
  // Scatter destinations: each contribution is written once, directly at its position in
  // the slot-sorted contribution arrays (no per-model intermediate tensors).
  const int* jdest;          // SoA (J, n_dev): index into jsorted, -1 = masked (ground)
  double* jsorted;           // slot-sorted Jacobian contributions: jr + c0*jc; NULL = the caller needs no Jacobian
                             //   (residual-side outputs only: the build at the converged X)
  const int* vdest;          // SoA (R, n_dev): record index into vsorted, -1 = masked
  double* vsorted;           // row-sorted records of 4 doubles: res_resist, res_react(Q),
                             //   lim_rhs_resist, lim_rhs_react
};

struct VjxInitArgs {
  int n_dev;
  const double* vals;        // packed init inputs
  const int* col_off;        // per init param: offset into vals
  const int* col_stride;     // per init param: 0 = shared scalar, 1 = per-instance column
  double* cache;             // SoA (cache_size, n_dev)
  float* collapse;           // SoA (n_collapse_outputs, n_dev)
};

typedef int (*vjx_model_eval_fn)(const VjxEvalArgs*, cudaStream_t);
typedef int (*vjx_model_init_fn)(const VjxInitArgs*, cudaStream_t);
typedef int (*vjx_model_set_const_fn)(const double* shared_params, int n_shared,
                                      const double* shared_cache, int n_shared_cache,
                                      const double* simparam_defaults, int n_simparams,
                                      cudaStream_t);
