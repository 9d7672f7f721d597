/* libvajax_b200.so - C-ABI of the B200-native transient Newton-Raphson hot path of VAJAX.
 *
 * Three nested seams of the reference are exported; each entry point names
 * the reference interface it replaces.  All pointers are plain; `*_dev` arguments are CUDA
 * device pointers owned by the caller (XLA / torch), everything else is host memory.  No call
 * allocates caller-visible outputs; scratch lives in the handle.  One in-flight call per
 * handle.  Every call returns 0 on success or a negative vjx_status; vjx_last_error() gives
 * the message (the reference's FFI handlers return ffi::Error::InvalidArgument / Internal,
 * vajax/sparse/umfpack_ffi.cpp:201-204,297-300).
 */
#ifndef VAJAX_B200_H
#define VAJAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vjx_handle vjx_handle;
typedef void* vjx_stream; /* cudaStream_t */

enum vjx_status {
  VJX_OK = 0,
  VJX_ERR_INVALID = -1,  /* shape / index validation failed                      */
  VJX_ERR_CUDA = -2,     /* CUDA runtime error                                   */
  VJX_ERR_PLUGIN = -3,   /* model plugin missing or ABI mismatch                 */
  VJX_ERR_SINGULAR = -4, /* structurally or numerically singular at analysis     */
  VJX_ERR_HANG = -5,     /* dependency wait watchdog fired inside a kernel       */
  VJX_ERR_NODEVICE = -6  /* no CUDA device: there is no CPU fallback             */
};

/* Source waveform descriptor (reference: vajax/analysis/sources.py:26-187). */
enum vjx_source_kind { VJX_SRC_DC = 0, VJX_SRC_PULSE = 1, VJX_SRC_SINE = 2, VJX_SRC_PWL = 3,
                       VJX_SRC_EXP = 4, VJX_SRC_AM = 5, VJX_SRC_FM = 6 };
typedef struct {
  int kind;
  double p[8];         /* dc | val0,val1,rise,fall,width,period,delay | sinedc,ampl,freq,phase |
                          (pwl) period | val0,val1,delay,td2,tau1,tau2 |
                          sinedc,ampl,freq,phase,delay,modfreq,modphase,modindex            */
  int pwl_offset;      /* first index into pwl_times / pwl_values                           */
  int pwl_count;
} vjx_source_desc;

/* One Verilog-A model type = one emitted CUDA plugin + its instance tables
 * (reference: compiled_models[model_type] + static_inputs_cache[model_type],
 *  vajax/analysis/openvaf_models.py:1032-1056, vajax/analysis/mna_builder.py:44-118). */
typedef struct {
  const char* plugin_path;        /* shared object emitted by vajax_b200.codegen.cuda_emitter   */
  int n_dev, n_volt, n_dev_params, n_dev_cache, n_res, n_jac, n_lim;
  const int* vnode1;              /* (n_dev, n_volt) row-major circuit node of V(hi,..)         */
  const int* vnode2;              /* (n_dev, n_volt) row-major circuit node of V(..,lo)         */
  const double* dev_params;       /* (n_dev, n_dev_params) row-major varying non-voltage params */
  const double* dev_cache;        /* (n_dev, n_dev_cache) row-major varying cache columns       */
  const double* shared_params;    int n_shared;
  const double* shared_cache;     int n_shared_cache;
  const double* simparam_defaults; int n_simparams;
  const int* res_idx;             /* (n_dev, n_res) residual row or -1   (build_stamp_index_mapping) */
  const int* jac_row;             /* (n_dev, n_jac) Jacobian row or -1                           */
  const int* jac_col;             /* (n_dev, n_jac) Jacobian col or -1                           */
} vjx_model_desc;

/* Options (reference: SimulationOptions vajax/analysis/options.py:34-190 and AdaptiveConfig
 * vajax/analysis/transient/adaptive.py:37-91, hot-path subset). */
typedef struct {
  double nr_damping, reltol, abstol, vntol, gmin, max_step;
  int tran_itl;
  /* adaptive step control */
  double lte_ratio, redo_factor, lte_reltol, lte_abstol, min_dt, grow_factor, tran_ft;
  double gshunt_init, gshunt_target;
  int gshunt_steps, warmup_steps, max_order, max_consecutive_rejects;
  int integration_method;   /* 0 = backward Euler, 1 = trapezoidal, 2 = gear2/BDF2 */
  double pivot_tol;         /* threshold for diagonal preference at analysis (KLU-like), e.g. 1e-3 */
  int use_device_graph;     /* 1: NR + step loops run as CUDA-graph WHILE nodes on the device   */
} vjx_options;

typedef struct {
  int n_total;              /* nodes incl. ground (external + surviving internal)               */
  int n_external;           /* recorded columns of V_out (incl. ground column 0)                */
  int n_vsources, n_isources;
  int nse;                  /* CSR stored entries                                               */
  const int* csr_indptr;    /* [n_aug + 1], n_aug = n_total - 1 + n_vsources                    */
  const int* csr_indices;   /* [nse] sorted, strictly increasing per row (full_mna.py:461-513)  */
  int n_models;
  const vjx_model_desc* models;   /* in the reference's model_types iteration order             */
  const int* vs_node_p;     /* [n_vsources] circuit node indices                                */
  const int* vs_node_n;
  const int* is_node_p;     /* [n_isources]                                                     */
  const int* is_node_n;
  const vjx_source_desc* vsrc;    /* [n_vsources]                                               */
  const vjx_source_desc* isrc;    /* [n_isources]                                               */
  const double* pwl_times; const double* pwl_values; int n_pwl;
  const int* noi_nodes;     int n_noi;       /* node indices (in V) clamped to 0 V              */
  const int* internal_nodes; int n_internal; /* all internal device node indices (full_mna.py:360-374) */
  vjx_options opts;
} vjx_circuit_desc;

const char* vjx_last_error(void);
int vjx_version(void);
int vjx_device_count(void);

/* Handle lifetime.  Replaces the process-global caches keyed by (n, nnz) in
 * vajax/sparse/umfpack_ffi.cpp:140-179 with an explicit handle. */
int vjx_create(const vjx_circuit_desc* desc, int device, vjx_handle** out);
void vjx_destroy(vjx_handle* h);
void vjx_clear_cache(void); /* API parity with umfpack_jax.clear_cache (umfpack_jax.py:71-77) */

/* Replace per-instance data (corner / Monte-Carlo sweeps re-use topology, symbolic analysis
 * and plugins; reference engine.py:2235-2320 re-runs prepare() per corner). */
int vjx_update_model_data(vjx_handle* h, int model, const double* dev_params, const double* dev_cache,
                          const double* shared_params, const double* shared_cache);
int vjx_update_sources(vjx_handle* h, const vjx_source_desc* vsrc, const vjx_source_desc* isrc);

/* ---- Seam S2: build_system (vajax/analysis/mna_builder.py:528-550, 759-1076) ------------
 * (X, vsource_vals, isource_vals, Q_prev, c0, gmin, gshunt, c1, d1, dQdt_prev, c2, Q_prev2,
 *  limit_state_in, nr_iteration) -> (csr_data, f, Q, I_vsource, limit_state_out, max_res_contrib)
 * Null output pointers are skipped; null optional inputs read as zeros. */
int vjx_build_system(vjx_handle* h, const double* X_dev, const double* vsrc_dev, const double* isrc_dev,
                     const double* Q_prev_dev, double c0, double gmin, double gshunt, double c1, double d1,
                     const double* dQdt_prev_dev, double c2, const double* Q_prev2_dev,
                     const double* limit_in_dev, int nr_iteration,
                     double* csr_out_dev, double* f_out_dev, double* Q_out_dev, double* Ivs_out_dev,
                     double* limit_out_dev, double* maxres_out_dev, vjx_stream stream);

/* ---- Seam S1: linear solve (vajax/sparse/umfpack_ffi.cpp:189-195,329-338:
 * (S32[n+1] indptr, S32[nnz] indices, F64[nnz] data, F64[n] b) -> F64[n] x; pattern is the
 * handle's).  Numeric refactorisation on the fixed symbolic pattern + pivot order, then
 * level-scheduled triangular solves.  The first call performs the host analysis. */
int vjx_lu_solve(vjx_handle* h, const double* csr_data_dev, const double* b_dev, double* x_dev, vjx_stream stream);
/* Force a new analysis (ordering + pivot sequence) from the given values. */
int vjx_lu_analyze(vjx_handle* h, const double* csr_data_dev, vjx_stream stream);
/* Symbolic statistics, 10 doubles: [0]=nnz(L+U) incl. diag, [1]=factor flops, [2]=levels(L),
 * [3]=levels(U), [4]=max row length, [5]=n, [6]=peeled singleton pivots, [7]=core size,
 * [8]=longest L part of a row, [9]=entries of the scatter program */
int vjx_lu_stats(vjx_handle* h, double* out10);
/* Sweep tuning (no reference counterpart: the reference runs corners one after the other, engine.py:2235-2320):
 * the persistent factorisation / back-solve kernels of this handle take 1/div of the GPU's CTA slots so that several
 * handles in flight on one GPU share the SMs.  Results do not depend on it.  Call before the first solve. */
int vjx_lu_set_grid_div(vjx_handle* h, int div);
/* A12, the reference's DENSE linear-solve path (make_dense_full_mna_solver.linear_solve,
 * vajax/analysis/solver_factories.py:784-788: solve(J + 1e-14 * I, -f); the reference's default below ~1000 nodes):
 * reg is added to every diagonal entry - node rows and branch rows - inside each NR solve, after the NOI rows were
 * rewritten, exactly where the reference adds it; what vjx_build_system returns stays the unregularised J.  The
 * factorisation itself remains this library's sparse LU.  Needs the branch rows' diagonal entries in the CSR pattern
 * given to vjx_create (VJX_ERR_INVALID otherwise); 0 = the sparse paths' semantics (default).  Call before the first solve. */
int vjx_set_dense_regularization(vjx_handle* h, double reg);

/* ---- Seam S3: nr_solve (vajax/analysis/solver_factories.py:604-725) ---------------------- */
int vjx_nr_solve(vjx_handle* h, const double* X_init_dev, const double* vsrc_dev, const double* isrc_dev,
                 const double* Q_prev_dev, double c0, double gmin, double gshunt, double c1, double d1,
                 const double* dQdt_prev_dev, double c2, const double* Q_prev2_dev,
                 const double* limit_in_dev, const double* res_tol_floor_dev, int max_iterations,
                 double* X_out_dev, int* iterations, int* converged, double* max_f,
                 double* Q_out_dev, double* dQdt_out_dev, double* Ivs_out_dev, double* limit_out_dev,
                 double* maxres_out_dev, vjx_stream stream);

/* Stream-ordered form of vjx_nr_solve for callers that must not block the host (an XLA FFI handler runs inside
 * XLA's stream; the reference's handler contract is vajax/sparse/umfpack_ffi.cpp:189-338): same operands, the NR
 * loop runs as a CUDA graph WHILE node on `stream`, and (iterations, converged, max_f, error_flags) are written
 * to scalars4_out_dev by the device.  Nothing is read back and nothing synchronises, except on the very first
 * call on a handle (symbolic analysis of the first Jacobian + graph construction).  Inputs are copied into
 * handle-owned staging buffers, so the caller's buffers may be reused as soon as the call returns in stream
 * order.  error_flags != 0 (LU watchdog / zero pivot) is the caller's to check. */
int vjx_nr_solve_async(vjx_handle* h, const double* X_init_dev, const double* vsrc_dev, const double* isrc_dev,
                       const double* Q_prev_dev, double c0, double gmin, double gshunt, double c1, double d1,
                       const double* dQdt_prev_dev, double c2, const double* Q_prev2_dev,
                       const double* limit_in_dev, const double* res_tol_floor_dev, int max_iterations,
                       double* X_out_dev, double* scalars4_out_dev,
                       double* Q_out_dev, double* dQdt_out_dev, double* Ivs_out_dev, double* limit_out_dev,
                       double* maxres_out_dev, vjx_stream stream);

/* ---- Seam S3': device-resident adaptive transient loop (full_mna.py:1840-2194 body_fn,
 * chunked like _run_with_checkpoints :1443-1699).  State lives in the handle. */
int vjx_tran_begin(vjx_handle* h, const double* X0_dev, const double* Q_init_dev, const double* I0_dev,
                   const double* hist_maxres_dev, double t_stop, double dt, double max_dt, int max_steps,
                   vjx_stream stream);
/* N4: device-side save list (netlist `save v(p0) ...`, vajax/benchmarks/data/mul64/runme.sim:296-308; the
 * reference records all external nodes, full_mna.py:1017-1023,2122).  node_idx_host[k] = index into X of V_out
 * column k; n_save = 0 = every external node.  Call before vjx_tran_begin. */
int vjx_tran_set_save_list(vjx_handle* h, const int* node_idx_host, int n_save);
/* Run until t >= t_stop, step_idx >= max_steps, or `max_attempts` step attempts were made.
 * Returns the number of attempts made in *attempts. */
int vjx_tran_run(vjx_handle* h, int max_attempts, int* attempts, vjx_stream stream);
/* Change the accepted-step bound (1 <= step_limit <= max_steps of vjx_tran_begin); the next
 * vjx_tran_run stops when step_idx reaches it.  Chunked runs a la _run_with_checkpoints
 * (full_mna.py:1443-1699) and bench.py's warm-up / timed split use it. */
int vjx_tran_set_step_limit(vjx_handle* h, int step_limit, vjx_stream stream);
/* scalars: [0]=t [1]=dt [2]=step_idx [3]=total_nr_iters [4]=rejected [5]=min_dt_used
 *          [6]=max_dt_used [7]=history_count [8]=warmup_count [9]=consecutive_rejects
 *          [10]=attempts [11]=error_flags */
int vjx_tran_scalars(vjx_handle* h, double* out12, vjx_stream stream);
/* Copy recorded outputs [row0, row0+rows) to host buffers: times[rows], V_out[rows*n_external],
 * I_out[rows*max(n_vsources,1)]. */
int vjx_tran_read(vjx_handle* h, int row0, int rows, double* times, double* V_out, double* I_out, vjx_stream stream);
/* Per-attempt debug records (step, t_next, dt, nr_iters, accept, lte_reject, nr_failed, lte_norm)
 * = the fields of the reference's debug_steps line (full_mna.py:2019-2035); 8 doubles per attempt. */
int vjx_tran_records(vjx_handle* h, int first, int count, double* out, vjx_stream stream);
/* Current solution / charge state (device -> host): X[n_total+n_vs], Q_prev[n_unknowns]. */
int vjx_tran_state(vjx_handle* h, double* X, double* Q_prev, double* dQdt_prev, vjx_stream stream);

/* Sizes of a handle: out8 = {n_total, n_external, n_vsources, n_isources, n_unknowns, n_aug, nse,
 * total_limit_states}.  What an FFI handler needs to validate operand shapes (the reference's handler
 * reads them from the buffer dimensions, vajax/sparse/umfpack_ffi.cpp:196-215). */
int vjx_sizes(vjx_handle* h, int* out8);
/* XLA FFI attributes cannot carry pointers, so a handle travels through a jitted function as a 64-bit id
 * (the reference needs none: its caches are process-global, umfpack_ffi.cpp:140-179).  vjx_handle_id
 * registers the handle (idempotent), vjx_handle_from_id resolves it (NULL if unknown or destroyed). */
long long vjx_handle_id(vjx_handle* h);
vjx_handle* vjx_handle_from_id(long long id);

/* Kernel timing probe used by bench.py: average device time (ms) of `reps` launches of one phase
 * on the handle's current state: 0=device eval, 1=jacobian stamp, 2=residual, 3=LU refactorisation
 * with the forward solve folded in, 4=(nothing: reports 0), 5=back solve, 6=NR update. Also returns
 * the number of kernels launched per phase. */
int vjx_time_phase(vjx_handle* h, int phase, int reps, double* ms_avg, int* launches, vjx_stream stream);
/* Device time in ms (CUDA events recorded on `stream` tightly around the loop launch) of the last
 * vjx_tran_run and the total since vjx_tran_begin. */
int vjx_tran_run_ms(vjx_handle* h, double* last_ms, double* total_ms);
/* Debug probe (not on the product path): per permuted row, completion time in ns since kernel start
 * and dependency level of one LU phase (3 = refactorisation + forward solve, 5 = back solve). */
int vjx_lu_timeline(vjx_handle* h, int phase, double* ns_out, int* level_out, vjx_stream stream);
/* Number of kernels launched by this handle since creation (for bench.py's gpu_launches). */
long long vjx_launch_count(vjx_handle* h);

#ifdef __cplusplus
}
#endif
#endif
