// libvajax_b200: B200 (sm_100a) implementation of VAJAX's transient Newton-Raphson hot path.
// See include/vajax_b200.h for the boundary and DESIGN.md for the data layout.
//
// Kernel inventory (all fp64, all deterministic - no floating-point atomics):
//   model eval (emitted plugin)  K1+K2  one thread per instance, scatters into slot-sorted arrays
//   k_isrc                       K4     current-source contributions
//   k_stamp_jacobian             K3,K5,K8  segmented sum over slot-sorted contributions -> CSR values
//   k_residual                   K4,K6,K7  row-sorted vector sums, residual assembly, convergence
//   k_lu_load / k_lu_rows /      K9     dependency-driven (sync-free) row-wise refactorisation with the forward
//   k_lu_dense_tail                     solve folded in; dense trailing block in the registers of one CTA
//   k_lu_usolve                  K9     dependency-driven back solve
//   k_delta_reduce / k_nr_update K7,K10 delta convergence, step limiting, X update, NOI clamp
//   k_pre_step .. k_step_commit  K11,K12 integration coefficients, sources, predictor, LTE,
//                                       accept/reject, history, outputs - all on the device
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/vajax_b200.h"
#include "vjx_model_abi.h"
#include "vjx_symbolic.h"

static thread_local std::string g_err;
static void* plugin_open(const char* path, std::string& key);
static void plugin_close(void* dl, const std::string& key);
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(VJX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

typedef unsigned long long ull;

// ------------------------------------------------------------------------------------------
// Device-side state
// ------------------------------------------------------------------------------------------
struct NRStat {
  int iter;          // NR iteration counter
  int converged;
  int bad_res;       // any |f_check| !< res_tol
  int bad_delta;     // any |conv_delta| !< tol
  int max_iter;
  int cont;          // loop-continue flag (host-polled mode mirror of the WHILE handle)
  int error;         // watchdog / pivot flags
  int pad;
  ull max_f_bits;    // max |f_check|   (bit pattern of a non-negative double: order-preserving)
  ull max_dv_bits;   // max |delta V|
  ull lim_delta_bits;
  ull lim_ref_bits;
  unsigned epoch;    // count of NR iterations started (diagnostics)
  unsigned ticket_f, ticket_l, ticket_u;   // row tickets: k_lu_rows launch 1 / launch 2, k_lu_usolve
  unsigned ticket_hf, ticket_hl, pad1;      // heavy-row tickets of the two k_lu_rows launches
};

// State every NR iteration starts from: per-iteration reductions cleared, LU row tickets reset.  Inside the NR loop
// this is done by the kernels that are launched anyway (the loop's entry - k_nr_begin / k_pre_step - for the first
// iteration, k_nr_update / k_nr_reduce_update at the end of an iteration for the next one); k_iter_begin is the
// stand-alone form for the single builds and solves outside the loop.
__device__ __forceinline__ void nr_clear_delta(NRStat* st) {
  st->bad_delta = 0;
  st->max_dv_bits = 0ULL;
  st->lim_delta_bits = 0ULL;
  st->lim_ref_bits = 0ULL;
}
__device__ __forceinline__ void nr_clear_tickets(NRStat* st) {
  st->epoch += 1u;
  st->ticket_f = 0u;
  st->ticket_l = 0u;
  st->ticket_u = 0u;
  st->ticket_hf = 0u;
  st->ticket_hl = 0u;
}
__device__ __forceinline__ void nr_clear_iter(NRStat* st) {
  st->bad_res = 0;
  st->max_f_bits = 0ULL;
  nr_clear_delta(st);
  nr_clear_tickets(st);
}

struct TranScalars {
  double t, dt, t_stop, max_dt, min_dt_used, max_dt_used;
  double dt_hist[4];
  int hist_count, step_idx, warmup, total_nr, rejected, consec, attempts, hist_head, max_steps, cont;
  int attempts_limit, pad0;
  // per-attempt
  double t_next, dt_cur, eci, a[3], pred_err, gshunt, new_dt;
  int can_predict, accept, use_new, nr_fail_min, lte_reject, nr_failed, reject, pad;
  ull lte_bits;
};

struct LUDev {
  int n;
  long long nnz;
  const int* ptr;
  const int* col;
  const int* diag;
  const int* src;
  const int* init;             // per L+U entry: what k_lu_load writes (slot of A / 0.0 / sentinel / nothing)
  const int* kub;              // per L entry (i,k): start of U(k,:) in val (= diag[k] + 1)
  const int* order_u;
  const int* perm_r;
  const int* perm_c;
  const int* order_f;          // all rows by L-dependency level; the first n_nol have no L part
  int n_nol;                   // rows without an L part (finished by k_lu_load)
  int max_lrow;                // longest row with an L part
  double* scratch;             // per-warp global working vectors, only when max_lrow > rowbuf
  // Dense trailing block: rows [dense_start, n) stop their row updates at column dense_start and hand
  // the rest (+ their partial y in column n_dense) to k_lu_dense_tail as a dense n_dense x (n_dense+1)
  // row-major matrix.
  int dense_start, n_dense;
  const int* dense_lstop;      // per dense row: first position of the row whose column is >= dense_start
  double* dense;
  int tail_start;              // rows [tail_start, n): chain tail of the back solve, one thread block
  const int* long_u;           // rows with a very long U part: one thread block each (last blocks of the grid)
  int n_long_u, n_bulk_u;      // order_u holds n_bulk_u rows
  int n_flat_u;                // ... the first n_flat_u of them have no U part (x = y / pivot, no dependency)
  const long long* upd_ptr;    // [nnz+1] start of each L entry's scatter list
  const int* upd_pos;          // position inside row i of each column of U(k,:)
  long long n_upd_prefetch;    // entries of upd_pos k_lu_load prefetches into L2 (0: the description is larger than L2)
  // heavy rows (vjx_symbolic.h, Heavy): walked depth by depth by a whole thread block
  const int* hv_prog;
  const long long* hv_off;
  const int* hv_rows;
  const unsigned char* hv_flag; // per position of order_f: the row is a heavy row (the warps skip it)
  unsigned long long* ts;      // optional per-row completion timestamps (globaltimer, debug probe)
  double* val;
  double* inv_diag;
  double* y;
  double* xp;
};

__device__ __forceinline__ double nan0(double v) { return (v != v) ? 0.0 : v; }
__device__ __forceinline__ ull global_ns() {
  ull t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void atomic_max_pos(ull* addr, double v) {
  atomicMax(addr, (ull)__double_as_longlong(v));
}

// ------------------------------------------------------------------------------------------
// Sources (reference: vajax/analysis/sources.py:26-187)
// ------------------------------------------------------------------------------------------
__device__ double py_mod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}
__device__ double eval_source(const vjx_source_desc& s, double t, const double* pt, const double* pv) {
  const double* p = s.p;
  const double PI = 3.141592653589793;
  switch (s.kind) {
    case VJX_SRC_DC: return p[0];
    case VJX_SRC_PULSE: {
      double v0 = p[0], v1 = p[1], r = p[2], f = p[3], w = p[4], per = p[5], d = p[6];
      double tp = py_mod(t - d, per);
      double rising = v0 + (v1 - v0) * tp / r;
      double falling = v1 - (v1 - v0) * (tp - r - w) / f;
      if (t < d) return v0;
      if (tp < r) return rising;
      if (tp < r + w) return v1;
      if (tp < r + w + f) return falling;
      return v0;
    }
    case VJX_SRC_SINE: return p[0] + p[1] * sin(2 * PI * p[2] * t + p[3]);
    case VJX_SRC_PWL: {
      double per = p[0];
      if (per > 0) t = py_mod(t, per);
      const double* xp = pt + s.pwl_offset;
      const double* fp = pv + s.pwl_offset;
      int n = s.pwl_count;
      if (n <= 0) return 0.0;
      if (t <= xp[0]) return fp[0];
      if (t >= xp[n - 1]) return fp[n - 1];
      int lo = 0, hi = n - 1;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= t) lo = mid; else hi = mid;
      }
      double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
      return slope * (t - xp[lo]) + fp[lo];
    }
    case VJX_SRC_EXP: {
      double v0 = p[0], v1 = p[1], d = p[2], td = p[3], t1 = p[4], t2 = p[5];
      double tf = d + td;
      double rising = v0 + (v1 - v0) * (1 - exp(-(t - d) / t1));
      double raf = v0 + (v1 - v0) * (1 - exp(-td / t1));
      double falling = raf + (v0 - raf) * (1 - exp(-(t - tf) / t2));
      return t < d ? v0 : (t < tf ? rising : falling);
    }
    case VJX_SRC_AM: {
      double dc = p[0], a = p[1], f = p[2], ph = p[3], d = p[4], mf = p[5], mph = p[6], mi = p[7];
      double initial = dc + a * sin(ph * PI / 180) * (1 + mi * sin(mph * PI / 180));
      double carrier = sin(2 * PI * f * (t - d) + ph * PI / 180);
      double mod = 1 + mi * sin(2 * PI * mf * (t - d) + mph * PI / 180);
      return t < d ? initial : dc + a * carrier * mod;
    }
    case VJX_SRC_FM: {
      double dc = p[0], a = p[1], f = p[2], ph = p[3], d = p[4], mf = p[5], mph = p[6], mi = p[7];
      double initial = dc + a * sin(ph * PI / 180 + mi * sin(mph * PI / 180));
      double cph = 2 * PI * f * (t - d) + ph * PI / 180;
      double mc = mi * sin(2 * PI * mf * (t - d) + mph * PI / 180);
      return t < d ? initial : dc + a * sin(cph + mc);
    }
  }
  return 0.0;
}

// ------------------------------------------------------------------------------------------
// Stamping kernels
// ------------------------------------------------------------------------------------------
__global__ void k_isrc(int n2, const int* __restrict__ dest, const double* __restrict__ isrc, double* vsorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  int rec = dest[i];
  if (rec < 0) return;
  double v = isrc[i >> 1] * ((i & 1) ? -1.0 : 1.0);
  double* q = vsorted + (size_t)rec * 4;
  q[0] = nan0(v);
  q[1] = 0.0;
  q[2] = 0.0;
  q[3] = 0.0;
}

enum { SLOT_DIAG = 1, SLOT_NOI_ZERO = 2, SLOT_NOI_ONE = 4, SLOT_BRANCH_DIAG = 8 };

// Segments longer than this are reduced by a whole CTA (supply-rail diagonals collect one
// contribution per attached transistor: 10^4 at c6288 size, 10^5 at mul64 size).  Shorter ones are summed
// by one thread, which sets the kernel's duration: at c6288 size 99 % of the node rows hold <= 18
// contributions, 34 input nets hold 64-96.  Measured (residual / stamp, us): 96 -> 22 / 12.4, 32 -> 14 / 11.6,
// 16 -> 16 / 18.5.
#define VJX_LONG_SEG 32

// Deterministic CTA sum: fixed strided partials, fixed shuffle butterfly, fixed warp order.
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) t += sh[w];
  return t;
}
__device__ __forceinline__ double block_max_256(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) t = fmax(t, sh[w]);
  return t;
}

// csr[s] = sum of the slot's contributions in the reference's scatter order
// (model order, instance, entry; vajax/analysis/mna_builder.py:953-958), then the diagonal
// regularisation (:969) and the vsource incidence constants (:972-989).  With enforce_noi the
// NOI rows/columns are rewritten as in solver_factories.py:1268-1276.
// dreg != 0 is the reference's DENSE linear-solve path (A12, make_dense_full_mna_solver.linear_solve,
// solver_factories.py:784-788): J + dreg * I on every diagonal, node rows and branch rows alike, added where the
// reference adds it - inside the NR solve, after the NOI rows were rewritten - and not in what build_system returns.
__device__ __forceinline__ double finish_slot(double acc, unsigned char f, float c, double reg, int enforce_noi, double dreg) {
  if (f & SLOT_DIAG) acc += reg;
  if (c != 0.0f) acc += (double)c;
  if (enforce_noi) {
    if (f & SLOT_NOI_ZERO) acc = 0.0;
    if (f & SLOT_NOI_ONE) acc = 1.0;
    if (dreg != 0.0 && (f & (SLOT_DIAG | SLOT_BRANCH_DIAG))) acc += dreg;
  }
  return acc;
}

// Long segments are cut into chunks of VJX_CHUNK entries; one CTA reduces one chunk (fixed strided
// partials, fixed butterfly), publishes the partial and takes a ticket on the segment's counter; the
// CTA that takes the last ticket adds the partials in chunk order (warp 0: lane-strided, butterfly) and
// finishes the slot.  The order of additions is a function of the tables only -> bit-identical runs.
#define VJX_CHUNK 4096
struct LongSeg {
  const int* chunk_seg;      // [n_chunks] index into seg_id / first / count
  const int* chunk_begin;    // [n_chunks] first entry of the chunk
  const int* chunk_end;      // [n_chunks]
  const int* seg_id;         // [n_long] CSR slot (or node row) of the long segment
  const int* first;          // [n_long + 1] first chunk of each long segment
  unsigned* count;           // [n_long] arrival counters (self-resetting)
  double* partial;           // [n_chunks * width]
  int n_chunks;
};

// Returns true in every thread of the CTA that took the last ticket of long segment k.
__device__ __forceinline__ bool long_seg_arrive(const LongSeg& L, int k, int* sh_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned need = (unsigned)(L.first[k + 1] - L.first[k]);
    const unsigned old = atomicAdd(L.count + k, 1u);
    *sh_flag = (old + 1u == need);
    if (old + 1u == need) { L.count[k] = 0u; __threadfence(); }
  }
  __syncthreads();
  return *sh_flag != 0;
}

// Grid layout: CTAs [0, n_chunks) each reduce one chunk of a long segment; the rest stride over all
// slots and skip the long ones.
__global__ void __launch_bounds__(256) k_stamp_jacobian(int nse, const int* __restrict__ seg,
                                 const double* __restrict__ jsorted, const unsigned char* __restrict__ flags,
                                 const float* __restrict__ cst, const double* __restrict__ ctl, int enforce_noi,
                                 LongSeg L, double* __restrict__ csr, double dreg) {
  __shared__ double sh[8];
  __shared__ int sh_flag;
  const double reg = ctl[VJX_CTL_DIAGREG] + ctl[VJX_CTL_GSHUNT];
  if ((int)blockIdx.x < L.n_chunks) {
    const int c = blockIdx.x;
    const int k = L.chunk_seg[c];
    const int b = L.chunk_begin[c], e = L.chunk_end[c];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int p = b + threadIdx.x;
    for (; p + 3 * 256 < e; p += 4 * 256) {
      a0 += jsorted[p]; a1 += jsorted[p + 256]; a2 += jsorted[p + 512]; a3 += jsorted[p + 768];
    }
    for (; p < e; p += 256) a0 += jsorted[p];
    double acc = block_sum_256((a0 + a1) + (a2 + a3), sh);
    if (threadIdx.x == 0) L.partial[c] = acc;
    if (!long_seg_arrive(L, k, &sh_flag)) return;
    if (threadIdx.x < 32) {
      double t = 0.0;
      for (int q = L.first[k] + threadIdx.x; q < L.first[k + 1]; q += 32) t += __ldcg(L.partial + q);
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      const int s = L.seg_id[k];
      if (threadIdx.x == 0) csr[s] = finish_slot(t, flags[s], cst[s], reg, enforce_noi, dreg);
    }
    return;
  }
  const int nb = gridDim.x - L.n_chunks;
  for (int s = (blockIdx.x - L.n_chunks) * blockDim.x + threadIdx.x; s < nse; s += nb * blockDim.x) {
    int b = seg[s], e = seg[s + 1];
    if (e - b > VJX_LONG_SEG) continue;
    double acc = 0.0;
    for (int p = b; p < e; ++p) acc += jsorted[p];
    csr[s] = finish_slot(acc, flags[s], cst[s], reg, enforce_noi, dreg);
  }
}

struct ResidualArgs {
  int n_unknowns, n_vs, n_total;
  const int* vseg;            // [n_unknowns+1] row segments of vsorted records
  const double* vsorted;
  const int* att_ptr;         // [n_unknowns+1] vsources attached to each node row
  const int* att_idx;         // vsource index
  const signed char* att_sign;  // +1 p terminal, -1 n terminal
  const int* vs_p;
  const int* vs_n;
  const double* X;
  const double* vsrc;
  const double* Q_prev;
  const double* dQdt_prev;
  const double* Q_prev2;
  const double* res_tol_floor;  // null -> abstol
  const unsigned char* conv_mask;
  const double* ctl;
  double reltol, abstol, vntol;
  double* f;
  double* Q;
  double* Ivs;
  double* maxres;
  NRStat* st;
  int check;                  // accumulate convergence statistics
  LongSeg L;                  // node rows with more than VJX_LONG_SEG contributions, in chunks
};

// One node row of the residual once its five segment reductions are known
// (reference: _build_residual vajax/analysis/mna_builder.py:1012-1076, combine_transient_residual
// vajax/analysis/mna.py:673-719, residual tolerance solver_factories.py:271-281).
__device__ __forceinline__ void finish_node_row(const ResidualArgs& a, int i, double fr, double q, double lr, double lq,
                                                double mx, double c0, double c1, double d1, double c2, double eff,
                                                double& f, double& tol) {
  fr = fr - lr;
  double fb = 0.0;
  for (int p = a.att_ptr[i]; p < a.att_ptr[i + 1]; ++p) {
    int v = a.att_idx[p];
    if (a.att_sign[p] > 0) {
      if (a.Ivs) a.Ivs[v] = -fr;
      fb += a.X[a.n_total + v];
    } else {
      fb += -a.X[a.n_total + v];
    }
  }
  if (a.n_vs > 0) fr = fr + fb;
  double qp = a.Q_prev ? a.Q_prev[i] : 0.0;
  double dq = a.dQdt_prev ? a.dQdt_prev[i] : 0.0;
  double q2 = a.Q_prev2 ? a.Q_prev2[i] : 0.0;
  f = fr - 0.0 + c0 * (q - lq) + c1 * qp + d1 * dq + c2 * q2;
  f = f + eff * a.X[i + 1];
  if (a.Q) a.Q[i] = q;
  if (a.maxres) a.maxres[i] = mx;
  double fl = a.res_tol_floor ? a.res_tol_floor[i] : a.abstol;
  tol = fmax(mx * a.reltol, fl);
  if (mx != mx || fl != fl) tol = mx * 0.0 + fl * 0.0 + tol;  // jnp.maximum propagates NaN
}

// Grid layout: CTAs [0, n_chunks) each reduce one chunk of a long node row (five partials per chunk);
// the rest stride over all rows.
__global__ void __launch_bounds__(256) k_residual(ResidualArgs a) {
  __shared__ double sh[8];
  __shared__ int sh_flag;
  const int n_aug = a.n_unknowns + a.n_vs;
  const double c0 = a.ctl[VJX_CTL_C0], c1 = a.ctl[VJX_CTL_C1], d1 = a.ctl[VJX_CTL_D1], c2 = a.ctl[VJX_CTL_C2];
  const double eff = a.ctl[VJX_CTL_DIAGREG] + a.ctl[VJX_CTL_GSHUNT];
  int bad = 0;
  double mf = 0.0;
  bool mfnan = false;
  if ((int)blockIdx.x < a.L.n_chunks) {
    const int c = blockIdx.x;
    const int k = a.L.chunk_seg[c];
    double fr = 0.0, q = 0.0, lr = 0.0, lq = 0.0, mx = 0.0;
    for (int p = a.L.chunk_begin[c] + threadIdx.x; p < a.L.chunk_end[c]; p += blockDim.x) {
      const double2* r = reinterpret_cast<const double2*>(a.vsorted + (size_t)p * 4);
      double2 r0 = r[0], r1 = r[1];
      fr += r0.x;
      mx = fmax(mx, fabs(r0.x));
      q += r0.y;
      lr += r1.x;
      lq += r1.y;
    }
    fr = block_sum_256(fr, sh);
    q = block_sum_256(q, sh);
    lr = block_sum_256(lr, sh);
    lq = block_sum_256(lq, sh);
    mx = block_max_256(mx, sh);
    if (threadIdx.x == 0) {
      double* o = a.L.partial + (size_t)c * 5;
      o[0] = fr; o[1] = q; o[2] = lr; o[3] = lq; o[4] = mx;
    }
    if (!long_seg_arrive(a.L, k, &sh_flag)) return;
    if (threadIdx.x < 32) {
      fr = 0.0; q = 0.0; lr = 0.0; lq = 0.0; mx = 0.0;
      for (int g = a.L.first[k] + threadIdx.x; g < a.L.first[k + 1]; g += 32) {
        const double* o = a.L.partial + (size_t)g * 5;
        fr += __ldcg(o); q += __ldcg(o + 1); lr += __ldcg(o + 2); lq += __ldcg(o + 3); mx = fmax(mx, __ldcg(o + 4));
      }
      for (int o = 16; o > 0; o >>= 1) {
        fr += __shfl_xor_sync(0xffffffffu, fr, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
        lr += __shfl_xor_sync(0xffffffffu, lr, o);
        lq += __shfl_xor_sync(0xffffffffu, lq, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
    }
    if (threadIdx.x == 0) {
      const int i = a.L.seg_id[k];
      double f, tol;
      finish_node_row(a, i, fr, q, lr, lq, mx, c0, c1, d1, c2, eff, f, tol);
      if (a.f) a.f[i] = f;
      if (a.check) {
        double fc = a.conv_mask[i] ? f : 0.0;
        double af = fabs(fc);
        if (af != af) atomicMax(&a.st->max_f_bits, 0x7ff8000000000000ULL);
        else if (af > 0.0) atomic_max_pos(&a.st->max_f_bits, af);
        if (!(af < tol)) atomicOr(&a.st->bad_res, 1);
      }
    }
    return;
  }
  const int nb = gridDim.x - a.L.n_chunks;
  for (int i = (blockIdx.x - a.L.n_chunks) * blockDim.x + threadIdx.x; i < n_aug; i += nb * blockDim.x) {
    double f, tol;
    if (i < a.n_unknowns) {
      if (a.vseg[i + 1] - a.vseg[i] > VJX_LONG_SEG) continue;
      double fr = 0.0, q = 0.0, lr = 0.0, lq = 0.0, mx = 0.0;
      for (int p = a.vseg[i]; p < a.vseg[i + 1]; ++p) {
        const double2* r = reinterpret_cast<const double2*>(a.vsorted + (size_t)p * 4);
        double2 r0 = r[0], r1 = r[1];
        fr += r0.x;
        mx = fmax(mx, fabs(r0.x));
        q += r0.y;
        lr += r1.x;
        lq += r1.y;
      }
      finish_node_row(a, i, fr, q, lr, lq, mx, c0, c1, d1, c2, eff, f, tol);
    } else {
      int v = i - a.n_unknowns;
      f = a.X[a.vs_p[v]] - a.X[a.vs_n[v]] - a.vsrc[v];
      tol = a.vntol;
      if (a.Ivs && a.vs_p[v] <= 0) a.Ivs[v] = 0.0;
    }
    if (a.f) a.f[i] = f;
    if (a.check) {
      double fc = a.conv_mask[i] ? f : 0.0;
      double af = fabs(fc);
      if (af != af) mfnan = true; else mf = fmax(mf, af);
      if (!(af < tol)) bad = 1;
    }
  }
  if (a.check) {
    // warp-level pre-reduction, then integer atomics (deterministic)
    for (int o = 16; o > 0; o >>= 1) {
      mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o));
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
      mfnan |= (bool)__shfl_xor_sync(0xffffffffu, (int)mfnan, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (mfnan) atomicMax(&a.st->max_f_bits, 0x7ff8000000000000ULL);
      else if (mf > 0.0) atomic_max_pos(&a.st->max_f_bits, mf);
      if (bad) atomicOr(&a.st->bad_res, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Sparse LU: numeric refactorisation on the fixed pattern / pivot order, dependency driven, with the
// forward solve folded in (y = L^-1 P b is one more column carried through the elimination), then a
// dependency-driven back solve.
//
// Readiness travels WITH the data.  Everything another row may wait for - the U part of a row, its
// 1/pivot, its y, and x in the back solve - is set to a sentinel bit pattern by k_lu_load at the start
// of the factorisation and overwritten exactly once by its producer.  A consumer polls the value
// itself (64-bit relaxed gpu-scope loads are single-copy atomic, every datum validates itself, no
// fence is needed) instead of polling a flag and then fetching the data: one L2 round trip per
// dependency link instead of two (profiles/r1_lu_notes.md).  The sentinel is a quiet-NaN payload that
// arithmetic never produces (a computed NaN is the canonical 0x7ff8000000000000).
//
// Rows are handed to warps through a ticket counter in dependency-level order: a warp holding a
// ticket only ever waits for rows with smaller tickets, which are held by running warps - no
// residency requirement, no deadlock.
// ------------------------------------------------------------------------------------------
#define VJX_SPIN_LIMIT (1u << 24)
#define VJX_SENTINEL 0x7ff8dead0badbeefULL

__device__ __forceinline__ ull ld_poll_u64(const double* p) {
  ull v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_pub_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double bits_f64(ull b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Flat pass, full grid: rows without an L part (peeled singletons, leaves of the elimination forest -
// 80 % of all rows) are plain copies of A's values and are finished here, pivot reciprocal and y
// included; the U part / pivot / y of every other row and all of x are set to the sentinel.
// init[p]: >= 0 slot of A, -1 structural fill (0.0), -2 sentinel, -3 leave (L part of an L row).
__global__ void __launch_bounds__(256) k_lu_load(LUDev lu, const double* __restrict__ csr, const double* __restrict__ b,
                                                 double scale, const unsigned char* __restrict__ zero_mask, NRStat* st) {
  const long long nt = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double sent = bits_f64(VJX_SENTINEL);
  // (the reductions of the Newton update that follows this solve: their last reader was the previous k_nr_update)
  if (t0 == 0) nr_clear_delta(st);
  // When the static description of the factorisation fits in L2, pull it there now, with full-grid coalesced
  // prefetches: the row kernels read it inside dependency chains, where a miss to HBM costs a k-step its
  // whole latency budget (first NR iteration after anything evicted it; free when it is resident already).
  if (lu.n_upd_prefetch > 0) {
    for (long long q = t0 * 32; q < lu.n_upd_prefetch; q += nt * 32) prefetch_l2(lu.upd_pos + q);   // 128 B lines
    for (long long q = t0 * 32; q < lu.nnz; q += nt * 32) {
      prefetch_l2(lu.col + q);
      prefetch_l2(lu.kub + q);
      prefetch_l2(lu.src + q);
    }
    for (long long q = t0 * 16; q < lu.nnz + 1; q += nt * 16) prefetch_l2(lu.upd_ptr + q);
  }
  for (long long p = t0; p < lu.nnz; p += nt) {
    const int c = lu.init[p];
    if (c >= 0) lu.val[p] = csr[c];
    else if (c == -1) lu.val[p] = 0.0;
    else if (c == -2) lu.val[p] = sent;
  }
  for (long long p = t0; p < (long long)lu.n_dense * (lu.n_dense + 1); p += nt) lu.dense[p] = 0.0;
  for (long long i = t0; i < lu.n; i += nt) {
    lu.xp[i] = sent;
    const int dg = lu.diag[i];
    if (dg == lu.ptr[i]) {
      const int s = lu.src[dg];
      const double d = s >= 0 ? csr[s] : 0.0;
      if (d == 0.0 || d != d) atomicOr(&st->error, 2);
      lu.inv_diag[i] = 1.0 / d;
      const int r = lu.perm_r[i];
      lu.y[i] = (zero_mask && zero_mask[r]) ? 0.0 : scale * b[r];
      if (lu.ts) lu.ts[i] = global_ns();
    } else {
      lu.inv_diag[i] = sent;
      lu.y[i] = sent;
    }
  }
}

// Rows with an L part, up-looking and k-sequential, one warp per row.  The row lives in a
// shared-memory working vector w (sized at analysis time for the longest such row; rows beyond the
// cap use a per-warp global scratch): for every L entry (i,k) in ascending k the warp forms
// l = w[k] / U(k,k), scatters w[pos] -= l * U(k,j) through the precomputed position list (lanes
// striding over U(k,:)) and lane 0 carries y_i -= l * y_k.  Rounding of the factor is that of the
// classic row update.
//
// The L entries are taken 32 at a time: lane b fetches the static description of step b (coalesced)
// and issues ONE non-blocking load of 1/pivot and y of its row k_b; the steps then run in order, and
// only a step whose inputs still hold the sentinel polls.  The first 32 U values / positions of the
// next VJX_PF steps are kept in flight in a register ring (with a look-ahead of one step every k-step
// waited a full L2 latency for its own prefetch: 0.35 us per step), so a row whose dependencies are
// long finished streams its k-steps at shared-memory speed, and a row at the head of a dependency
// chain pays one L2 round trip per link.  All polling is CONVERGENT (whole warp iterates until __all_sync): lanes
// parked in different spin loops starve the lane whose value has arrived (profiles/r1_lu_notes.md).
//
// Publication order: U part, pivot reciprocal and y first (what other rows wait for), the L part
// (needed by nobody before the next kernel) afterwards.
#define VJX_FACTOR_WARPS 4
#ifndef VJX_UG
#define VJX_UG 4   // chunks of 32 U entries requested together per k-step (beyond the first)
#endif
#ifndef VJX_PF
#define VJX_PF 4   // k-steps whose first chunks are in flight
#endif
extern __shared__ double vjx_lu_smem[];

// Heavy rows: the first n_hv_cta thread blocks of a k_lu_rows launch take the rows with many k-steps, one row per
// BLOCK, and walk them group by group (vjx_symbolic.h, Heavy): the k-steps of one dependency depth are independent,
// so every warp forms their multipliers at once (lane s: l_s = w[kpos_s] / U(k_s,k_s), redundantly in each warp - no
// hand-over between warps), then the warps share the group's records: a lane owns one position of the row and applies
// its <= 4 contributions -l_slot * U in ascending k, l_slot by shuffle.  One barrier per group instead of ~550
// cycles of dependent latency per k-step.  The row lives in the block's whole shared memory (all four warps'
// working vectors); L entries go straight to global memory (w[kpos] stays the undivided value the other warps still
// read).  The program text is read through L1 (read-only path) and touched a few KB ahead, the U values / pivots
// of the next record / group are requested before the current one is applied.  Readiness is the sentinel protocol
// of the other rows: polling is convergent per warp.
#define VJX_HV_PF_INTS 6144
#define VJX_HV_WARPS 3   // of the block's 4: the last warp keeps taking ordinary rows, so a launch makes progress even
                         // if only its heavy-row blocks are resident (several handles sharing the GPU)
__device__ __forceinline__ void hv_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * VJX_HV_WARPS) : "memory"); }
// Pulls a line of the program into L1 ahead of its use: a real read-only load whose value is folded into a checksum
// one call later (ptxas deletes a load nobody consumes; the deferred use never waits for it).
__device__ __forceinline__ void hv_touch(const int* p, unsigned& acc, unsigned& last) {
  acc ^= last;
  last = (unsigned)__ldg(p);
}

struct HvRec {
  int pos, R;
  int e[4];
  ull u[4];
};

__device__ __forceinline__ void hv_fetch(const LUDev& lu, const int* G, int r, int lane, HvRec& o) {
  const int info = __ldg(G + VJX_HV_GRP_INTS + r);
  const int* Rp = G + (info & 0x0fffffff);
  o.R = info >> 28;
  o.pos = __ldg(Rp + lane);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o.e[j] = -1;
    o.u[j] = 0ULL;
    if (j < o.R) {
      o.e[j] = __ldg(Rp + 32 + 32 * j + lane);
      if (o.e[j] != -1) o.u[j] = ld_poll_u64(lu.val + (((unsigned)o.e[j]) >> 5));
    }
  }
}

__device__ __forceinline__ void lu_heavy_role(const LUDev& lu, const double* __restrict__ csr, const double* __restrict__ b,
                                              double scale, const unsigned char* __restrict__ zero_mask, NRStat* st,
                                              unsigned* ticket_h, int hv_begin, int hv_end) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* const w = vjx_lu_smem;
  __shared__ unsigned s_tk;
  unsigned pf_acc = 0u, pf_last = 0u;
  while (true) {
    hv_bar();
    if (tid == 0) s_tk = atomicAdd(ticket_h, 1u);
    hv_bar();
    const unsigned tk = s_tk;
    if (tk >= (unsigned)(hv_end - hv_begin)) {
      if ((pf_acc ^ pf_last) == 0x9e3779b9u && tk == 0xffffffffu) atomicOr(&st->error, 16);   // (keeps the touches alive)
      break;
    }
    const int hidx = hv_begin + (int)tk;
    const int i = lu.hv_rows[hidx];
    const int rb = lu.ptr[i], dg = lu.diag[i], re = lu.ptr[i + 1];
    const int* P = lu.hv_prog + lu.hv_off[hidx];
    const int* const Pend = lu.hv_prog + lu.hv_off[hidx + 1];
    const int* pf = P;
    {
      const int* lim = P + VJX_HV_PF_INTS < Pend ? P + VJX_HV_PF_INTS : Pend;
      for (const int* q = pf + tid * 32; q < lim; q += 32 * 32 * VJX_HV_WARPS) hv_touch(q, pf_acc, pf_last);
      pf = lim;
    }
    for (int p = rb + tid; p < re; p += 32 * VJX_HV_WARPS) {
      const int sidx = lu.src[p];
      w[p - rb] = sidx >= 0 ? csr[sidx] : 0.0;
    }
    double yacc = 0.0, l = 0.0;
    int nrec = 0, keep = 0, gints = 0, kpos = -1, kk = 0;
    ull inv_b = 0ULL, y_b = 0ULL;
    HvRec la;
    la.pos = -1; la.R = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { la.e[j] = -1; la.u[j] = 0ULL; }
    if (P < Pend) {
      nrec = __ldg(P); keep = __ldg(P + 1); gints = __ldg(P + 3);
      kpos = __ldg(P + 4 + lane); kk = __ldg(P + 36 + lane);
      if (kpos >= 0) { inv_b = ld_poll_u64(lu.inv_diag + kk); y_b = ld_poll_u64(lu.y + kk); }
      if (warp < nrec) hv_fetch(lu, P, warp, lane, la);
    }
    hv_bar();
    while (P < Pend) {
      {
        const int* lim = P + VJX_HV_PF_INTS < Pend ? P + VJX_HV_PF_INTS : Pend;
        for (const int* q = pf + tid * 32; q < lim; q += 32 * 32 * VJX_HV_WARPS) hv_touch(q, pf_acc, pf_last);
        if (lim > pf) pf = lim;
      }
      if (!keep) {
        unsigned spins = 0;
        while (true) {
          bool ok = true;
          if (kpos >= 0) {
            if (inv_b == VJX_SENTINEL) { inv_b = ld_poll_u64(lu.inv_diag + kk); ok = false; }
            if (y_b == VJX_SENTINEL) { y_b = ld_poll_u64(lu.y + kk); ok = false; }
          }
          if (!__all_sync(0xffffffffu, ok)) {
            // still waiting for a row of this group: refresh, in the same round trip, the U values of the record that
            // is applied next (a row publishes its U part before its pivot: they are there when the pivot is)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (la.e[j] != -1 && la.u[j] == VJX_SENTINEL) la.u[j] = ld_poll_u64(lu.val + (((unsigned)la.e[j]) >> 5));
            if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
            continue;
          }
          break;
        }
        l = 0.0;
        if (kpos >= 0) {
          l = w[kpos] * bits_f64(inv_b);
          if (warp == 0) {
            lu.val[rb + kpos] = l;
            yacc = __fma_rn(-l, bits_f64(y_b), yacc);
          }
        }
      }
      // header of the next group: its pivots / y are requested now, one group ahead
      const int* Pn = P + gints;
      int nrec_n = 0, keep_n = 0, gints_n = 0, kpos_n = -1, kk_n = 0;
      ull inv_n = 0ULL, y_n = 0ULL;
      if (Pn < Pend) {
        nrec_n = __ldg(Pn); keep_n = __ldg(Pn + 1); gints_n = __ldg(Pn + 3);
        kpos_n = __ldg(Pn + 4 + lane); kk_n = __ldg(Pn + 36 + lane);
        if (kpos_n >= 0) { inv_n = ld_poll_u64(lu.inv_diag + kk_n); y_n = ld_poll_u64(lu.y + kk_n); }
      }
      for (int r = warp; r < nrec; r += VJX_HV_WARPS) {
        HvRec c = la;
        if (r + VJX_HV_WARPS < nrec) hv_fetch(lu, P, r + VJX_HV_WARPS, lane, la);
        else if (Pn < Pend && warp < nrec_n) hv_fetch(lu, Pn, warp, lane, la);
        unsigned spins = 0;
        while (true) {
          bool ok = true;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c.e[j] != -1 && c.u[j] == VJX_SENTINEL) { c.u[j] = ld_poll_u64(lu.val + (((unsigned)c.e[j]) >> 5)); ok = false; }
          if (__all_sync(0xffffffffu, ok)) break;
          if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
        }
        double acc = c.pos >= 0 ? w[c.pos] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < c.R) {   // warp-uniform
            const double lv = __shfl_sync(0xffffffffu, l, c.e[j] & 31);
            if (c.e[j] != -1) acc = __fma_rn(-lv, bits_f64(c.u[j]), acc);
          }
        }
        if (c.pos >= 0) w[c.pos] = acc;
      }
      if (warp >= nrec && Pn < Pend && warp < nrec_n) hv_fetch(lu, Pn, warp, lane, la);
      hv_bar();
      P = Pn; nrec = nrec_n; keep = keep_n; gints = gints_n; kpos = kpos_n; kk = kk_n; inv_b = inv_n; y_b = y_n;
    }
    // y of the row: the lanes' partial sums in lane order (fixed), on top of the right-hand side
    double ysum = 0.0;
    if (warp == 0) {
      if (lane == 0) {
        const int r = lu.perm_r[i];
        ysum = (zero_mask && zero_mask[r]) ? 0.0 : scale * b[r];
      }
      for (int sl = 0; sl < 32; ++sl) {
        const double ys = __shfl_sync(0xffffffffu, yacc, sl);
        if (lane == 0) ysum += ys;
      }
    }
    if (i >= lu.dense_start) {
      const int lend = lu.dense_lstop[i - lu.dense_start];
      double* srow = lu.dense + (size_t)(i - lu.dense_start) * (lu.n_dense + 1);
      for (int p = lend + tid; p < re; p += 32 * VJX_HV_WARPS) srow[lu.col[p] - lu.dense_start] = w[p - rb];
      if (tid == 0) {
        srow[lu.n_dense] = ysum;
        if (lu.ts) lu.ts[lu.n + 8 + (i - lu.dense_start)] = global_ns();
      }
    } else {
      const double d = w[dg - rb];
      for (int p = dg + 1 + tid; p < re; p += 32 * VJX_HV_WARPS) st_pub_f64(lu.val + p, w[p - rb]);
      if (tid == 0) {
        if (d == 0.0 || d != d) atomicOr(&st->error, 2);
        st_pub_f64(lu.y + i, ysum);
        st_pub_f64(lu.inv_diag + i, 1.0 / d);
        lu.val[dg] = d;
        if (lu.ts) lu.ts[i] = global_ns();
      }
    }
  }
}

// Block rows: the same k-sequential row update as the warps below - every position receives the same fused
// multiply-adds in the same order, so the values are bit-identical - but by three warps of a block: the entries of
// U(k,:) are spread over 96 threads (a 1,000-entry U row of mul64's trailing part is 11 per thread, requested together,
// instead of 32 chunks walked by one warp) with one named barrier per k-step.  The next two k-steps' pivot, y and
// first three U values per thread (a whole 288-entry U row) are in flight while the current step is applied; a step whose row is not finished polls
// pivot and U values in the same round trip, exactly like the warp path.  For rows whose work (entries of their
// scatter program) is at least VJX_BR_WORK; same block layout / fourth warp / tickets as the heavy-row path above.
#define VJX_BR_E 3   // entries of U(k,:) per thread that travel with the look-ahead (3 x 96 = 288: a whole U row of mul64's heavy part)
struct BrStep {
  int k, ub, ue;
  long long mb;
  ull inv, y;
  int pos[VJX_BR_E];
  ull u[VJX_BR_E];
};

__device__ __forceinline__ void br_meta(const LUDev& lu, int p, int lend, BrStep& o) {
  o.k = -1; o.ub = 0; o.ue = 0; o.mb = 0; o.inv = 0ULL; o.y = 0ULL;
#pragma unroll
  for (int e = 0; e < VJX_BR_E; ++e) { o.pos[e] = -1; o.u[e] = 0ULL; }
  if (p < lend) {
    o.k = __ldg(lu.col + p);
    o.ub = __ldg(lu.kub + p);
    o.mb = __ldg(lu.upd_ptr + p);
    o.ue = o.ub + (int)(__ldg(lu.upd_ptr + p + 1) - o.mb);
  }
}
__device__ __forceinline__ void br_request(const LUDev& lu, int tid, BrStep& o) {
  if (o.k >= 0) {
    o.inv = ld_poll_u64(lu.inv_diag + o.k);
    if (tid == 0) o.y = ld_poll_u64(lu.y + o.k);
#pragma unroll
    for (int e = 0; e < VJX_BR_E; ++e) {
      const int q = o.ub + tid + e * 32 * VJX_HV_WARPS;
      if (q < o.ue) {
        o.pos[e] = __ldg(lu.upd_pos + o.mb + (q - o.ub));
        o.u[e] = ld_poll_u64(lu.val + q);
      }
    }
  }
}

__device__ __forceinline__ void lu_block_row_role(const LUDev& lu, const double* __restrict__ csr, const double* __restrict__ b,
                                                  double scale, const unsigned char* __restrict__ zero_mask, NRStat* st,
                                                  unsigned* ticket_h, int hv_begin, int hv_end) {
  constexpr int NT = 32 * VJX_HV_WARPS;
  const int tid = threadIdx.x;
  double* const w = vjx_lu_smem;
  __shared__ unsigned s_tk;
  while (true) {
    hv_bar();
    if (tid == 0) s_tk = atomicAdd(ticket_h, 1u);
    hv_bar();
    const unsigned tk = s_tk;
    if (tk >= (unsigned)(hv_end - hv_begin)) break;
    const int i = lu.hv_rows[hv_begin + (int)tk];
    const int rb = lu.ptr[i], dg = lu.diag[i], re = lu.ptr[i + 1];
    const bool dense_row = i >= lu.dense_start;
    const int lend = dense_row ? lu.dense_lstop[i - lu.dense_start] : dg;
    BrStep c, n1, n2, n3;
    br_meta(lu, rb, lend, c);
    br_meta(lu, rb + 1, lend, n1);
    br_meta(lu, rb + 2, lend, n2);
    br_request(lu, tid, c);
    br_request(lu, tid, n1);
    for (int p = rb + tid; p < re; p += NT) {
      const int sidx = lu.src[p];
      w[p - rb] = sidx >= 0 ? csr[sidx] : 0.0;
    }
    double ysum = 0.0;
    if (tid == 0) {
      const int r = lu.perm_r[i];
      ysum = (zero_mask && zero_mask[r]) ? 0.0 : scale * b[r];
    }
    hv_bar();
    for (int p = rb; p < lend; ++p) {
      br_request(lu, tid, n2);
      br_meta(lu, p + 3, lend, n3);
      // pivot (and y, thread 0) of row k: convergent per warp; the own first U value rides along
      {
        unsigned spins = 0;
        while (true) {
          bool ok = true;
          if (c.inv == VJX_SENTINEL) { c.inv = ld_poll_u64(lu.inv_diag + c.k); ok = false; }
          if (tid == 0 && c.y == VJX_SENTINEL) { c.y = ld_poll_u64(lu.y + c.k); ok = false; }
#pragma unroll
          for (int e = 0; e < VJX_BR_E; ++e)
            if (c.pos[e] >= 0 && c.u[e] == VJX_SENTINEL) c.u[e] = ld_poll_u64(lu.val + c.ub + tid + e * NT);
          if (__all_sync(0xffffffffu, ok)) break;
          if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
        }
      }
      const double l = w[p - rb] * bits_f64(c.inv);
      // the entries that came with the look-ahead
      if (c.ub + (tid & ~31) < c.ue) {   // warp-uniform: this warp has entries in the first pass
        unsigned spins = 0;
        while (true) {
          bool ok = true;
#pragma unroll
          for (int e = 0; e < VJX_BR_E; ++e)
            if (c.pos[e] >= 0 && c.u[e] == VJX_SENTINEL) { c.u[e] = ld_poll_u64(lu.val + c.ub + tid + e * NT); ok = false; }
          if (__all_sync(0xffffffffu, ok)) break;
          if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
        }
#pragma unroll
        for (int e = 0; e < VJX_BR_E; ++e)
          if (c.pos[e] >= 0) w[c.pos[e]] = __fma_rn(-l, bits_f64(c.u[e]), w[c.pos[e]]);
      }
      // rest of a long U(k,:): 6 entries per thread requested together
      for (int q0 = c.ub + VJX_BR_E * NT; q0 < c.ue; q0 += 6 * NT) {
        if (q0 + (tid & ~31) >= c.ue) break;   // warp-uniform
        ull u6[6];
        int pos6[6];
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          u6[g] = 0ULL; pos6[g] = -1;
          const int q = q0 + g * NT + tid;
          if (q < c.ue) { pos6[g] = __ldg(lu.upd_pos + c.mb + (q - c.ub)); u6[g] = ld_poll_u64(lu.val + q); }
        }
        unsigned spins = 0;
        while (true) {
          bool ok = true;
#pragma unroll
          for (int g = 0; g < 6; ++g)
            if (pos6[g] >= 0 && u6[g] == VJX_SENTINEL) { u6[g] = ld_poll_u64(lu.val + q0 + g * NT + tid); ok = false; }
          if (__all_sync(0xffffffffu, ok)) break;
          if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
        }
#pragma unroll
        for (int g = 0; g < 6; ++g)
          if (pos6[g] >= 0) w[pos6[g]] = __fma_rn(-l, bits_f64(u6[g]), w[pos6[g]]);
      }
      if (tid == 0) {
        lu.val[p] = l;
        ysum = __fma_rn(-l, bits_f64(c.y), ysum);
      }
      hv_bar();
      c = n1; n1 = n2; n2 = n3;
    }
    if (dense_row) {
      double* srow = lu.dense + (size_t)(i - lu.dense_start) * (lu.n_dense + 1);
      for (int p = lend + tid; p < re; p += NT) srow[lu.col[p] - lu.dense_start] = w[p - rb];
      if (tid == 0) {
        srow[lu.n_dense] = ysum;
        if (lu.ts) lu.ts[lu.n + 8 + (i - lu.dense_start)] = global_ns();
      }
    } else {
      const double d = w[dg - rb];
      for (int p = dg + 1 + tid; p < re; p += NT) st_pub_f64(lu.val + p, w[p - rb]);
      if (tid == 0) {
        if (d == 0.0 || d != d) atomicOr(&st->error, 2);
        st_pub_f64(lu.y + i, ysum);
        st_pub_f64(lu.inv_diag + i, 1.0 / d);
        lu.val[dg] = d;
        if (lu.ts) lu.ts[i] = global_ns();
      }
    }
  }
}

template <int UG, bool PIPE, int RC, bool HV>
__global__ void __launch_bounds__(32 * VJX_FACTOR_WARPS) k_lu_rows(LUDev lu, const double* __restrict__ csr,
                                                                    const double* __restrict__ b, double scale,
                                                                    const unsigned char* __restrict__ zero_mask, NRStat* st,
                                                                    unsigned* ticket, int t_begin, int t_end, int rowbuf,
                                                                    int n_hv_cta, int hv_begin, int hv_end, unsigned* ticket_h) {
  if constexpr (HV) {   // (a launch without heavy rows keeps its register count: the wide launch lives on occupancy)
    if ((int)blockIdx.x < n_hv_cta && (threadIdx.x >> 5) < VJX_HV_WARPS) {
      if (lu.hv_prog) lu_heavy_role(lu, csr, b, scale, zero_mask, st, ticket_h, hv_begin, hv_end);
      else lu_block_row_role(lu, csr, b, scale, zero_mask, st, ticket_h, hv_begin, hv_end);
      return;
    }
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double* const wsm = vjx_lu_smem + (size_t)warp * rowbuf;
  while (true) {
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(ticket, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= (unsigned)(t_end - t_begin)) break;
    if (lu.hv_flag[t_begin + tk]) continue;   // taken by a heavy-row block
    const int i = lu.order_f[t_begin + tk];
    const int rb = lu.ptr[i], dg = lu.diag[i], re = lu.ptr[i + 1];
    double* w = (re - rb) <= rowbuf ? wsm : lu.scratch + ((size_t)blockIdx.x * VJX_FACTOR_WARPS + warp) * lu.max_lrow;
    for (int p = rb + lane; p < re; p += 32) {
      const int s = lu.src[p];
      w[p - rb] = s >= 0 ? csr[s] : 0.0;
    }
    double ysum = 0.0;
    if (lane == 0) {
      const int r = lu.perm_r[i];
      ysum = (zero_mask && zero_mask[r]) ? 0.0 : scale * b[r];
    }
    __syncwarp();
    const bool dense_row = i >= lu.dense_start;
    const int lend = dense_row ? lu.dense_lstop[i - lu.dense_start] : dg;
    for (int pb = rb; pb < lend; pb += 32) {
      const int nb = min(32, lend - pb);
      int k = 0, ub = 0, ue = 0;
      long long mb = 0;
      ull inv_l = 0ULL, yk_l = 0ULL;
      if (lane < nb) {
        k = lu.col[pb + lane];
        ub = lu.kub[pb + lane];
        mb = lu.upd_ptr[pb + lane];
        ue = ub + (int)(lu.upd_ptr[pb + lane + 1] - mb);
        inv_l = ld_poll_u64(lu.inv_diag + k);
        yk_l = ld_poll_u64(lu.y + k);
      }
      // ring of VJX_PF steps in flight: the first 32 U values / positions of steps bb .. bb+VJX_PF-1
      ull u_r[VJX_PF][RC];
      int pos_r[VJX_PF][RC];
      // (single-launch variant, registers to spare: the bounds of the prefetched steps stay in registers too,
      // which saves the step its own three shuffles)
      constexpr bool KEEP = RC > 1;
      int ub_k[KEEP ? VJX_PF : 1], ue_k[KEEP ? VJX_PF : 1];
      long long mb_k[KEEP ? VJX_PF : 1];
      ull inv_k[KEEP ? VJX_PF : 1], yk_k[KEEP ? VJX_PF : 1];
#pragma unroll
      for (int j = 0; j < VJX_PF; ++j) {
#pragma unroll
        for (int c = 0; c < RC; ++c) { u_r[j][c] = 0ULL; pos_r[j][c] = 0; }
        if (j < nb) {
          const int ub_j = __shfl_sync(0xffffffffu, ub, j), ue_j = __shfl_sync(0xffffffffu, ue, j);
          const long long mb_j = __shfl_sync(0xffffffffu, mb, j);
          if constexpr (KEEP) {
            ub_k[j] = ub_j; ue_k[j] = ue_j; mb_k[j] = mb_j;
            inv_k[j] = __shfl_sync(0xffffffffu, inv_l, j);
            yk_k[j] = __shfl_sync(0xffffffffu, yk_l, j);
          }
#pragma unroll
          for (int c = 0; c < RC; ++c) {
            const int q = ub_j + 32 * c + lane;
            if (q < ue_j) { pos_r[j][c] = lu.upd_pos[mb_j + 32 * c + lane]; u_r[j][c] = ld_poll_u64(lu.val + q); }
          }
        }
      }
      for (int bb0 = 0; bb0 < nb; bb0 += VJX_PF) {
#pragma unroll
        for (int j = 0; j < VJX_PF; ++j) {
          const int bb = bb0 + j;
          if (bb < nb) {   // warp-uniform
            int ub_c, ue_c;
            long long mb_c;
            if constexpr (KEEP) {
              ub_c = ub_k[j]; ue_c = ue_k[j]; mb_c = mb_k[j];
            } else {
              ub_c = __shfl_sync(0xffffffffu, ub, bb); ue_c = __shfl_sync(0xffffffffu, ue, bb);
              mb_c = __shfl_sync(0xffffffffu, mb, bb);
            }
            ull inv_s, yk_s;
            if constexpr (KEEP) {
              inv_s = inv_k[j]; yk_s = yk_k[j];
            } else {
              inv_s = __shfl_sync(0xffffffffu, inv_l, bb); yk_s = __shfl_sync(0xffffffffu, yk_l, bb);
            }
            if (inv_s == VJX_SENTINEL || yk_s == VJX_SENTINEL) {
              // row k is not finished yet: poll its pivot / y and, in the same round trip, its first 32 U
              // values, so that the step can start as soon as the row appears
              const int k_s = __shfl_sync(0xffffffffu, k, bb);
              const bool has0 = ub_c + lane < ue_c;
              unsigned spins = 0;
              while (true) {
                if (lane == 0) { inv_s = ld_poll_u64(lu.inv_diag + k_s); yk_s = ld_poll_u64(lu.y + k_s); }
                if (has0) u_r[j][0] = ld_poll_u64(lu.val + ub_c + lane);
                inv_s = __shfl_sync(0xffffffffu, inv_s, 0);
                yk_s = __shfl_sync(0xffffffffu, yk_s, 0);
                if (inv_s != VJX_SENTINEL && yk_s != VJX_SENTINEL) break;
                if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
              }
            }
            const int kpos = pb + bb - rb;
            const double l = w[kpos] * bits_f64(inv_s);
            // first RC chunks: from the ring (re-polled only while they still hold the sentinel)
            {
              unsigned spins = 0;
              while (true) {
                bool ok = true;
#pragma unroll
                for (int c = 0; c < RC; ++c) {
                  const int q = ub_c + 32 * c + lane;
                  if (q < ue_c && u_r[j][c] == VJX_SENTINEL) { u_r[j][c] = ld_poll_u64(lu.val + q); ok = false; }
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
              }
#pragma unroll
              for (int c = 0; c < RC; ++c)
                if (ub_c + 32 * c + lane < ue_c) w[pos_r[j][c]] = __fma_rn(-l, bits_f64(u_r[j][c]), w[pos_r[j][c]]);
            }
            // refill this ring slot with step bb + VJX_PF before the long tail of this step
#pragma unroll
            for (int c = 0; c < RC; ++c) u_r[j][c] = 0ULL;
            if (bb + VJX_PF < nb) {
              const int ub_n = __shfl_sync(0xffffffffu, ub, bb + VJX_PF), ue_n = __shfl_sync(0xffffffffu, ue, bb + VJX_PF);
              const long long mb_n = __shfl_sync(0xffffffffu, mb, bb + VJX_PF);
              if constexpr (KEEP) {
                ub_k[j] = ub_n; ue_k[j] = ue_n; mb_k[j] = mb_n;
                inv_k[j] = __shfl_sync(0xffffffffu, inv_l, bb + VJX_PF);
                yk_k[j] = __shfl_sync(0xffffffffu, yk_l, bb + VJX_PF);
              }
#pragma unroll
              for (int c = 0; c < RC; ++c) {
                const int q = ub_n + 32 * c + lane;
                if (q < ue_n) { pos_r[j][c] = lu.upd_pos[mb_n + 32 * c + lane]; u_r[j][c] = ld_poll_u64(lu.val + q); }
              }
            }
            if constexpr (PIPE) {
              // long rows (their launch is limited by shared memory, registers are free): the next request of
              // UG chunks is in flight while the current one is validated and applied, so a 300-entry U row
              // costs about one L2 round trip instead of three
              if (ub_c + 32 * RC < ue_c) {
                ull u_a[UG], u_b[UG];
                int pos_a[UG], pos_b[UG];
                int q0 = ub_c + 32 * RC;
#pragma unroll
                for (int g = 0; g < UG; ++g) {
                  u_a[g] = 0ULL; pos_a[g] = 0;
                  const int q = q0 + g * 32 + lane;
                  if (q < ue_c) { pos_a[g] = lu.upd_pos[mb_c + (q - ub_c)]; u_a[g] = ld_poll_u64(lu.val + q); }
                }
                while (true) {
                  const int q1 = q0 + 32 * UG;
                  const bool more = q1 < ue_c;
                  if (more) {
#pragma unroll
                    for (int g = 0; g < UG; ++g) {
                      u_b[g] = 0ULL; pos_b[g] = 0;
                      const int q = q1 + g * 32 + lane;
                      if (q < ue_c) { pos_b[g] = lu.upd_pos[mb_c + (q - ub_c)]; u_b[g] = ld_poll_u64(lu.val + q); }
                    }
                  }
                  unsigned spins = 0;
                  while (true) {
                    bool ok = true;
#pragma unroll
                    for (int g = 0; g < UG; ++g) {
                      const int q = q0 + g * 32 + lane;
                      if (q < ue_c && u_a[g] == VJX_SENTINEL) { u_a[g] = ld_poll_u64(lu.val + q); ok = false; }
                    }
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
                  }
#pragma unroll
                  for (int g = 0; g < UG; ++g)
                    if (q0 + g * 32 + lane < ue_c) w[pos_a[g]] = __fma_rn(-l, bits_f64(u_a[g]), w[pos_a[g]]);
                  if (!more) break;
#pragma unroll
                  for (int g = 0; g < UG; ++g) { u_a[g] = u_b[g]; pos_a[g] = pos_b[g]; }
                  q0 = q1;
                }
              }
            } else {
              // rest of U(k,:), UG chunks of 32 requested together (keeping the next request in flight
              // costs 6 registers: one block per SM less in the wide launch, which is bound by occupancy)
              for (int q0 = ub_c + 32 * RC; q0 < ue_c; q0 += 32 * UG) {
                ull u_c[UG];
                int pos_c[UG];
  #pragma unroll
                for (int g = 0; g < UG; ++g) {
                  u_c[g] = 0ULL; pos_c[g] = 0;
                  const int q = q0 + g * 32 + lane;
                  if (q < ue_c) { pos_c[g] = lu.upd_pos[mb_c + (q - ub_c)]; u_c[g] = ld_poll_u64(lu.val + q); }
                }
                unsigned spins = 0;
                while (true) {
                  bool ok = true;
  #pragma unroll
                  for (int g = 0; g < UG; ++g) {
                    const int q = q0 + g * 32 + lane;
                    if (q < ue_c && u_c[g] == VJX_SENTINEL) { u_c[g] = ld_poll_u64(lu.val + q); ok = false; }
                  }
                  if (__all_sync(0xffffffffu, ok)) break;
                  if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, 1); break; }
                }
  #pragma unroll
                for (int g = 0; g < UG; ++g)
                  if (q0 + g * 32 + lane < ue_c) w[pos_c[g]] = __fma_rn(-l, bits_f64(u_c[g]), w[pos_c[g]]);
              }
            }
            __syncwarp();
            if (lane == 0) {
              w[kpos] = l;
              ysum = __fma_rn(-l, bits_f64(yk_s), ysum);
            }
            __syncwarp();
          }
        }
      }
    }
    if (dense_row) {
      // the part of the row inside the dense trailing block goes to k_lu_dense_tail, nobody waits for it
      double* srow = lu.dense + (size_t)(i - lu.dense_start) * (lu.n_dense + 1);
      for (int p = lend + lane; p < re; p += 32) srow[lu.col[p] - lu.dense_start] = w[p - rb];
      if (lane == 0) srow[lu.n_dense] = ysum;
      if (lu.ts && lane == 0) lu.ts[lu.n + 8 + (i - lu.dense_start)] = global_ns();
      for (int p = rb + lane; p < lend; p += 32) lu.val[p] = w[p - rb];
      __syncwarp();
      continue;
    }
    const double d = w[dg - rb];
    for (int p = dg + 1 + lane; p < re; p += 32) st_pub_f64(lu.val + p, w[p - rb]);
    if (lane == 0) {
      if (d == 0.0 || d != d) atomicOr(&st->error, 2);
      st_pub_f64(lu.y + i, ysum);
      st_pub_f64(lu.inv_diag + i, 1.0 / d);
      lu.val[dg] = d;
      if (lu.ts) lu.ts[i] = global_ns();
    }
    for (int p = rb + lane; p < dg; p += 32) lu.val[p] = w[p - rb];
    __syncwarp();
  }
}

// Dense trailing block.  The last rows of the elimination order form a near-dense block whose rows
// depend on each other one after the other (118 single-row levels at c6288 size): walked row by row
// through global memory every link costs an L2 round trip (~1.2 us).  Here ONE thread block holds the
// whole block in REGISTERS (10 x 5 doubles per thread, 512 threads: 159 rows + the y column), warp w
// owning rows w, w+16, ... and lane l owning columns l, l+32, ..., and runs the right-looking
// elimination with one __syncthreads per pivot: the pivot row and the pivot column travel through a
// double-buffered shared-memory line.  Every entry receives its updates in ascending pivot order, so
// the values are those of the row algorithm (both use fused multiply-add).  y rides along as column n_dense.  The block is
// then spilled to shared memory once and the back substitution of these rows runs on it
// (column-oriented, one barrier per unknown), publishing x for the rest of the back solve.
#define VJX_DT_THREADS 512
#define VJX_DT_R 10
#define VJX_DT_C 5
#define VJX_DT_MAX 159
#define VJX_DT_LDS 161
#define VJX_DT_SMEM (160 * VJX_DT_LDS * 8)
// (rows beyond T hold zeros in r and publish zeros into the pivot column, so they need no guard: -0 * u
// leaves them zero; lanes left of the pivot inside its column block see u = 0 instead of a predicate)
#define VJX_DT_UPDATE_ROW(A2, BC)                                                                  \
  {                                                                                                \
    const double nl_ = -(lcw[16 * (A2)] * inv);                                                    \
    _Pragma("unroll") for (int b = (BC); b < VJX_DT_C; ++b) r[A2][b] = __fma_rn(nl_, u[b], r[A2][b]); \
  }
#define VJX_DT_PUBLISH_ROW(A2, BC, LP)                                                             \
  {                                                                                                \
    _Pragma("unroll") for (int b = (BC); b < VJX_DT_C; ++b) urn[lane + 32 * b] = r[A2][b];         \
    if (lane == (LP)) {                                                                            \
      const double piv = r[A2][BC];                                                                \
      if (piv == 0.0 || piv != piv) bad = true;                                                    \
      urn[160] = 1.0 / piv;                                                                        \
    }                                                                                              \
  }
template <int A>
__device__ __forceinline__ void vjx_dt_panel(double (&r)[VJX_DT_R][VJX_DT_C], double* urow, double* lcol, const int T,
                                             const int wi, const int lane, bool& bad) {
  {
    constexpr int bcol = A >> 1;
    for (int w = 0; w < 16; ++w) {
      const int k = 16 * A + w;
      if (k >= T) return;
      const double* const ur = urow + (k & 1) * 192;
      const double* const lc = lcol + (k & 1) * 160;
      double* const urn = urow + ((k + 1) & 1) * 192;
      double* const lcn = lcol + ((k + 1) & 1) * 160;
      const int lpiv = 16 * (A & 1) + w;
      __syncthreads();
      const double inv = ur[160];
      const double* const lcw = lc + wi;
      double u[VJX_DT_C];
#pragma unroll
      for (int b = bcol; b < VJX_DT_C; ++b) u[b] = ur[lane + 32 * b];
      if (lane <= lpiv) u[bcol] = 0.0;   // inside the pivot's column block only the columns right of it change
      const bool has_next = k + 1 < T;
      if (w < 15) {
        // next pivot: row 16 A + w + 1 = register row A of warp w + 1, same column block
        if (wi == w + 1) {
          VJX_DT_UPDATE_ROW(A, bcol)
          if (has_next) VJX_DT_PUBLISH_ROW(A, bcol, lpiv + 1)
        } else if (wi > w + 1) {
          VJX_DT_UPDATE_ROW(A, bcol)
        }
#pragma unroll
        for (int a2 = A + 1; a2 < VJX_DT_R; ++a2) VJX_DT_UPDATE_ROW(a2, bcol)
        if (has_next && lane == lpiv + 1) {
#pragma unroll
          for (int a2 = A; a2 < VJX_DT_R; ++a2) lcn[wi + 16 * a2] = r[a2][bcol];
        }
      } else if constexpr (A + 1 < VJX_DT_R) {
        // last pivot of the panel: the next one is register row A + 1 of warp 0
        constexpr int An = A + 1;
        constexpr int bn = An >> 1;
        if (wi == 0) {
          VJX_DT_UPDATE_ROW(An, bcol)
          if (has_next) VJX_DT_PUBLISH_ROW(An, bn, 16 * (An & 1))
#pragma unroll
          for (int a2 = An + 1; a2 < VJX_DT_R; ++a2) VJX_DT_UPDATE_ROW(a2, bcol)
        } else {
#pragma unroll
          for (int a2 = An; a2 < VJX_DT_R; ++a2) VJX_DT_UPDATE_ROW(a2, bcol)
        }
        if (has_next && lane == 16 * (An & 1)) {
#pragma unroll
          for (int a2 = An; a2 < VJX_DT_R; ++a2) lcn[wi + 16 * a2] = r[a2][bn];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(VJX_DT_THREADS, 1) k_lu_dense_tail(LUDev lu, double* __restrict__ x_out, NRStat* st) {
  extern __shared__ double vjx_dt_smem[];
  const int T = lu.n_dense, ldg = T + 1;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double r[VJX_DT_R][VJX_DT_C];
#pragma unroll
  for (int a = 0; a < VJX_DT_R; ++a)
#pragma unroll
    for (int b = 0; b < VJX_DT_C; ++b) {
      const int i = wi + 16 * a, j = lane + 32 * b;
      r[a][b] = (i < T && j <= T) ? lu.dense[(size_t)i * ldg + j] : 0.0;
    }
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 1] = global_ns();
  double* const urow = vjx_dt_smem;         // [2][192] pivot row, slot 160 = 1 / pivot
  double* const lcol = vjx_dt_smem + 384;   // [2][160] pivot column
  bool bad = false;
  __syncthreads();
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 2] = global_ns();
  // The register tile can only be indexed with compile-time constants, so the pivots are walked in
  // panels of 16 (one per register row A, fully unrolled): inside panel A pivot k = 16 A + w belongs to
  // warp w, register row A, column block A / 2, lane 16 (A & 1) + w - all but w known at compile time.
  // Finished register rows (< A) and column blocks (< A / 2) drop out of the unrolled code.
  //
  // Look-ahead: in step k the owner of row k+1 updates that row FIRST, takes the reciprocal of the next
  // pivot and publishes the row into the other buffer while the remaining warps are still busy with
  // their step-k updates, so the division is off the critical path; every warp publishes its entries
  // of column k+1 after its own updates.  One barrier per pivot.  Row updates are branch-free
  // (a structural zero of L multiplies through as -0) so that independent rows overlap in the pipes.
  {
    double* const urn = urow;
    if (wi == 0) VJX_DT_PUBLISH_ROW(0, 0, 0)
    if (lane == 0) {
#pragma unroll
      for (int a2 = 0; a2 < VJX_DT_R; ++a2) lcol[wi + 16 * a2] = r[a2][0];
    }
  }
  vjx_dt_panel<0>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<1>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<2>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<3>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<4>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<5>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<6>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<7>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<8>(r, urow, lcol, T, wi, lane, bad);
  vjx_dt_panel<9>(r, urow, lcol, T, wi, lane, bad);
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&st->error, 2);
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 3] = global_ns();
  double* const S = vjx_dt_smem;
#pragma unroll
  for (int a = 0; a < VJX_DT_R; ++a)
#pragma unroll
    for (int b = 0; b < VJX_DT_C; ++b) {
      const int i = wi + 16 * a, j = lane + 32 * b;
      if (i < T && j <= T) S[i * VJX_DT_LDS + j] = r[a][b];
    }
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < T) {
    const double inv = 1.0 / S[tid * VJX_DT_LDS + tid];
    S[tid * VJX_DT_LDS + tid] = inv;
    lu.inv_diag[lu.dense_start + tid] = inv;
    lu.y[lu.dense_start + tid] = S[tid * VJX_DT_LDS + T];
  }
  __syncthreads();
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 4] = global_ns();
  double xmine = 0.0;
  for (int k = T - 1; k >= 0; --k) {
    const double xk = S[k * VJX_DT_LDS + T] * S[k * VJX_DT_LDS + k];
    if (tid < k) S[tid * VJX_DT_LDS + T] = __fma_rn(-S[tid * VJX_DT_LDS + k], xk, S[tid * VJX_DT_LDS + T]);
    if (tid == k) xmine = xk;
    __syncthreads();
  }
  if (tid < T) {
    const int i = lu.dense_start + tid;
    st_pub_f64(lu.xp + i, xmine);
    x_out[lu.perm_c[i]] = xmine;
    if (lu.ts) lu.ts[i] = global_ns();
  }
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 5] = global_ns();
}
#undef VJX_DT_UPDATE_ROW
#undef VJX_DT_PUBLISH_ROW

// ------------------------------------------------------------------------------------------
// Blocked dense trailing block (round 2).  Same arithmetic as k_lu_dense_tail - every entry receives
// a(i,j) = fma(-l(i,k), u(k,j), a(i,j)) in ascending pivot order with l(i,k) = a(i,k) * (1 / u(k,k)), so the
// results are bit-identical to it and to the row algorithm - but organised in panels of P = 8 pivots:
//   1. gather   the panel's columns of this CTA's rows go to shared memory (LpT, transposed), its 8 pivot rows
//               go to the shared memory of EVERY CTA of the cluster (Up; distributed shared memory),
//   2. diag     8 lanes of warp 0 factorise the 8 x 8 diagonal block with shuffles (no CTA barrier per pivot),
//   3. trsm     L21 = A21 U11^-1 (one thread per row) and U12 = L11^-1 A12 (one thread per column, y included),
//   4. update   rank-8 update of the register tiles: 8 (NR + NC) shared-memory loads for 8 NR NC FMAs.
// Three barriers per PANEL instead of one per pivot.  The matrix lives in the registers of a cluster of C CTAs
// (512 threads each): warp gw = 16 * rank + w of the cluster owns rows gw, gw + 16 C, ..., lane l owns columns
// l, l + 32, ...  Every CTA repeats steps 2 and 3b on identical data (cheaper than another exchange), so one
// cluster barrier per panel is the only inter-SM synchronisation.  C = 1: up to 159 rows in one CTA; C = 2: the
// same rows with twice the fp64 pipes; C = 4: up to 319 rows (mul64's 312-row trailing block).
// The finished U rows (y in column T) go back to lu.dense; the back substitution runs in CTA 0, one thread per
// row, in blocks of 16 unknowns: the warp holding those rows solves the 16 x 16 diagonal block in registers, the other
// rows apply the block with 16 FMAs from values they prefetched one block ahead.  Per-row operation order = descending column,
// as in k_lu_dense_tail.
#define VJX_DB_THREADS 512
#define VJX_DB_P 8
#define VJX_DB_MAX 319

template <int C>
__device__ __forceinline__ void vjx_db_sync() {
  if constexpr (C > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}
__device__ __forceinline__ unsigned vjx_cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of `p` (a shared-memory address of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ unsigned vjx_map_shared(const void* p, unsigned rank) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p), o;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(rank));
  return o;
}
__device__ __forceinline__ void vjx_st_cluster_f64(unsigned addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

template <int C, int NR, int NC>
struct VjxDB {
  static constexpr int NW = 16 * C;            // warps of the cluster = row classes
  static constexpr int LDS = NC * 32 + 1;      // row stride of Up
  static constexpr int NRL = NR * 16;          // rows held by one CTA
  static constexpr int LDL = NRL + 1;          // row stride of LpT
  static constexpr int TMAX = (NR * NW < NC * 32 ? NR * NW : NC * 32) - 1;
  static constexpr int NPANEL = (TMAX + VJX_DB_P) / VJX_DB_P;
  // shared memory (doubles): Up[2][P][LDS] | LpT[2][P][LDL] | invp[2][P] | ys[TMAX + 1] | xs[32]
  static constexpr int OFF_LPT = 2 * VJX_DB_P * LDS;
  static constexpr int OFF_INV = OFF_LPT + 2 * VJX_DB_P * LDL;
  static constexpr int OFF_YS = OFF_INV + 2 * VJX_DB_P;
  static constexpr int OFF_XS = OFF_YS + TMAX + 1;
  static constexpr int SMEM_BYTES = (OFF_XS + 32) * 8;
};

// Steps 2 and 3 of a panel and the write-back of its U rows: nothing here touches the register tile, so it is ONE
// copy of code for all panels (k0 is a run-time value).  Only the gather and the rank-8 update are instantiated per
// panel (vjx_db_panel<A>), and the update loops over the 8 pivots instead of unrolling them: fully unrolled the kernel
// was ~0.8 MB of straight-line code executed once and instruction fetch set its speed (ncu: no_instruction 8 stalls
// per issue); with every register index made run-time instead (predicated selects) it issued 4.4 x the
// instructions.  profiles/r2_dense_notes.md has the three variants.
// 1 / x for the pivots of the blocked kernel.  The pivot chain is the critical path of the whole dense phase (one
// reciprocal -> multiply -> FMA link per pivot, 159 - 319 links), and IEEE division costs it ~250 cycles a link.
// MUFU.RCP64H (rcp.approx.ftz.f64, 2^-20) and two Newton steps give 1 / x to within an ulp in 4 dependent DFMAs; the
// result is not the correctly rounded quotient, which nothing downstream requires (the factorisation is held to the
// residual and to agreement with SuperLU, not to the bits of another kernel).  Zero / NaN / Inf pivots fall through
// to the plain division so that the error flag logic sees what it always saw.
__device__ __forceinline__ double vjx_fast_rcp(double x) {
  const int hi = __double2hiint(x) & 0x7ff00000;
  if (hi == 0 || hi >= 0x7fd00000 || hi <= 0x00300000) return 1.0 / x;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = __fma_rn(-x, y, 1.0);
  y = __fma_rn(y, e, y);
  e = __fma_rn(-x, y, 1.0);
  return __fma_rn(y, e, y);
}

// debug probe of the blocked kernel (only when the timeline buffer lu.ts is set): cycles per phase of a panel,
// accumulated by thread 0 (on the critical path: warp 0 does the diagonal blocks) and by the last thread
__device__ long long vjx_db_cycles[2][8];
#define VJX_DB_TICK(slot)                                                          \
  if (lu.ts && (threadIdx.x == 0 || threadIdx.x == VJX_DB_THREADS - 1)) {         \
    const long long now_ = clock64();                                            \
    vjx_db_cycles[threadIdx.x == 0 ? 0 : 1][slot] += now_ - tick_;               \
    tick_ = now_;                                                                \
  }

template <int C, int NR, int NC>
__device__ __noinline__ void vjx_db_middle(double* Up, double* LpT, double* invp, const LUDev& lu, const int T, const int k0,
                                           const int crank, bool* bad_out) {
  using D = VjxDB<C, NR, NC>;
  constexpr int P = VJX_DB_P, LDS = D::LDS, LDL = D::LDL, NW = D::NW;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int np = min(P, T - k0);
  long long tick_ = lu.ts ? clock64() : 0;
  // ---- 2. diagonal block: lane t of warp 0 holds row k0 + t of the 8 x 8 block.  Nothing inside the pivot loop
  //         diverges or touches memory: every lane computes every reciprocal (same value), the stores happen once
  //         after the loop.  (With "if (lane == m) { store 1/pivot }" inside, each pivot paid a divergent region whose
  //         store addresses were loaded from the argument block: ~570 cycles per pivot instead of the chain's ~150.)
  if (wi == 0) {
    double a[P], invs[P];
    const int t = lane & (P - 1);
    bool anybad = false;
#pragma unroll
    for (int j = 0; j < P; ++j) a[j] = Up[t * LDS + k0 + j];
#pragma unroll
    for (int m = 0; m < P; ++m) {
      invs[m] = 0.0;
      if (m < np) {
        const double piv = __shfl_sync(0xffffffffu, a[m], m);
        const double inv = vjx_fast_rcp(piv);
        anybad |= (piv == 0.0) | (piv != piv);
        invs[m] = inv;
        const double l = a[m] * inv;
#pragma unroll
        for (int j = m + 1; j < P; ++j) {
          const double u = __shfl_sync(0xffffffffu, a[j], m);
          if (t > m) a[j] = __fma_rn(-l, u, a[j]);
        }
        if (t > m) a[m] = l;
      }
    }
    if (lane < P) {
      double myinv = invs[0];
#pragma unroll
      for (int m = 1; m < P; ++m) myinv = (lane == m) ? invs[m] : myinv;
      invp[lane] = myinv;
#pragma unroll
      for (int j = 0; j < P; ++j) Up[lane * LDS + k0 + j] = a[j];
      if (anybad && lane == 0) *bad_out = true;
      // the block's part of the finished U rows goes back to lu.dense for the back substitution (the rest: step 3b)
      if (crank == 0 && lane < np) {
        lu.inv_diag[lu.dense_start + k0 + lane] = myinv;
        double* const drow = lu.dense + (size_t)(k0 + lane) * (T + 1) + k0;
#pragma unroll
        for (int j = 0; j < P; ++j)
          if (j >= lane && k0 + j <= T) drow[j] = a[j];
      }
    }
  }
  VJX_DB_TICK(1)
  __syncthreads();
  VJX_DB_TICK(2)
  // ---- 3a. L21: this CTA's rows below the diagonal block, one thread per row (right-looking over the 8 pivots)
  if (threadIdx.x < D::NRL) {
    const int lr = threadIdx.x;
    const int i = crank * 16 + (lr & 15) + NW * (lr >> 4);
    if (i >= k0 + P && i < T) {
      double a[P];
#pragma unroll
      for (int j = 0; j < P; ++j) a[j] = LpT[j * LDL + lr];
#pragma unroll
      for (int m = 0; m < P; ++m) {
        if (m < np) {
          const double l = a[m] * invp[m];
          a[m] = l;
#pragma unroll
          for (int j = m + 1; j < P; ++j) a[j] = __fma_rn(-l, Up[m * LDS + k0 + j], a[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < P; ++j) LpT[j * LDL + lr] = a[j];
    }
  } else {
    // ---- 3b. U12 (and y, column T): one thread per column right of the block; every CTA computes all of it
    for (int j = k0 + P + ((int)threadIdx.x - D::NRL); j <= T; j += VJX_DB_THREADS - D::NRL) {
      double c[P];
#pragma unroll
      for (int m = 0; m < P; ++m) c[m] = Up[m * LDS + j];
#pragma unroll
      for (int mp = 0; mp < P - 1; ++mp)
#pragma unroll
        for (int m = mp + 1; m < P; ++m) c[m] = __fma_rn(-Up[m * LDS + k0 + mp], c[mp], c[m]);
#pragma unroll
      for (int m = 1; m < P; ++m) Up[m * LDS + j] = c[m];
      if (crank == 0) {   // finished U entries of this column -> lu.dense (consecutive threads, consecutive j)
        double* const dcol = lu.dense + (size_t)k0 * (T + 1) + j;
#pragma unroll
        for (int m = 0; m < P; ++m)
          if (m < np) dcol[(size_t)m * (T + 1)] = c[m];
      }
    }
  }
  VJX_DB_TICK(3)
  __syncthreads();
  VJX_DB_TICK(4)
}

template <int C, int NR, int NC, int A>
__device__ __forceinline__ void vjx_db_panel(double (&r)[NR][NC], double* sm, const LUDev& lu, const int T, const int crank,
                                             const int wi, const int lane, bool& bad) {
  using D = VjxDB<C, NR, NC>;
  constexpr int P = VJX_DB_P, LDS = D::LDS, LDL = D::LDL, NW = D::NW;
  constexpr int k0 = P * A;
  if (k0 >= T) return;
  constexpr int aA = k0 / NW;                // register row of the pivot rows
  constexpr int g0 = k0 % NW;                // cluster warp of the first pivot row
  constexpr int cA = g0 / 16, w0 = g0 % 16;  // owner CTA / first owner warp
  constexpr int bcol = k0 / 32, l0 = k0 % 32;
  double* const Up = sm + (A & 1) * P * LDS;
  double* const LpT = sm + D::OFF_LPT + (A & 1) * P * LDL;
  double* const invp = sm + D::OFF_INV + (A & 1) * P;
  const int np = min(P, T - k0);
  const int gw = crank * 16 + wi;
  long long tick_ = lu.ts ? clock64() : 0;
  // ---- 1. gather
  if (lane >= l0 && lane < l0 + P) {
#pragma unroll
    for (int a = aA; a < NR; ++a) LpT[(lane - l0) * LDL + a * 16 + wi] = r[a][bcol];
  }
  if (crank == cA && wi >= w0 && wi < w0 + P) {
    const int t = wi - w0;
    if constexpr (C > 1) {
#pragma unroll
      for (int dst = 0; dst < C; ++dst) {
        const unsigned base = vjx_map_shared(Up + t * LDS, dst);
#pragma unroll
        for (int b = bcol; b < NC; ++b) vjx_st_cluster_f64(base + 8u * (unsigned)(lane + 32 * b), r[aA][b]);
      }
    } else {
#pragma unroll
      for (int b = bcol; b < NC; ++b) Up[t * LDS + lane + 32 * b] = r[aA][b];
    }
  }
  VJX_DB_TICK(6)
  vjx_db_sync<C>();
  VJX_DB_TICK(0)
  vjx_db_middle<C, NR, NC>(Up, LpT, invp, lu, T, k0, crank, &bad);
  if (lu.ts) tick_ = clock64();
  // ---- 4. rank-8 update of the register tiles (rows below the block, columns right of it; a finished column inside
  //         the first live column block is updated with whatever Up holds there - it is never read again)
  constexpr int b0 = (l0 + P == 32) ? bcol + 1 : bcol;
  if constexpr (b0 < NC) {
    const bool first_row = gw >= g0 + P;
    const double* lp = LpT + wi;
    const double* up = Up + lane;
#pragma unroll 1
    for (int m = 0; m < np; ++m, lp += LDL, up += LDS) {
      double u[NC - b0];
#pragma unroll
      for (int b = b0; b < NC; ++b) u[b - b0] = up[32 * b];
      if (first_row) {
        const double nl = -lp[aA * 16];
#pragma unroll
        for (int b = b0; b < NC; ++b) r[aA][b] = __fma_rn(nl, u[b - b0], r[aA][b]);
      }
#pragma unroll
      for (int a = aA + 1; a < NR; ++a) {
        const double nl = -lp[a * 16];
#pragma unroll
        for (int b = b0; b < NC; ++b) r[a][b] = __fma_rn(nl, u[b - b0], r[a][b]);
      }
    }
  }
  VJX_DB_TICK(7)
  if constexpr (A + 1 < D::NPANEL) vjx_db_panel<C, NR, NC, A + 1>(r, sm, lu, T, crank, wi, lane, bad);
}

template <int C, int NR, int NC>
__global__ void __launch_bounds__(VJX_DB_THREADS, 1) k_lu_dense_blocked(LUDev lu, double* __restrict__ x_out, NRStat* st) {
  using D = VjxDB<C, NR, NC>;
  extern __shared__ double vjx_db_smem[];
  double* const sm = vjx_db_smem;
  const int T = lu.n_dense, ldg = T + 1;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = C > 1 ? (int)vjx_cluster_rank() : 0;
  const int gw = crank * 16 + wi;
  double r[NR][NC];
  if (lu.ts && threadIdx.x == 0 && crank == 0) lu.ts[lu.n + 1] = global_ns();
#pragma unroll
  for (int a = 0; a < NR; ++a)
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      const int i = gw + D::NW * a, j = lane + 32 * b;
      r[a][b] = (i < T && j <= T) ? lu.dense[(size_t)i * ldg + j] : 0.0;
    }
  bool bad = false;
  vjx_db_sync<C>();   // every CTA has read the block before CTA 0 starts writing U rows back into it
  if (lu.ts && threadIdx.x == 0 && crank == 0) lu.ts[lu.n + 2] = global_ns();
  vjx_db_panel<C, NR, NC, 0>(r, sm, lu, T, crank, wi, lane, bad);
  if (__syncthreads_or(bad) && threadIdx.x == 0 && crank == 0) atomicOr(&st->error, 2);
  if constexpr (C > 1) {
    vjx_db_sync<C>();   // nobody leaves while its shared memory may still be written / read by the others
    if (crank != 0) return;
  }
  if (lu.ts && threadIdx.x == 0) {
    lu.ts[lu.n + 3] = global_ns();
    for (int w = 0; w < 2; ++w) {
      printf("dense blocked C=%d T=%d thread %d cycles: wait-gather-sync %lld  diag %lld  sync %lld  trsm %lld  sync %lld  writeback %lld  gather %lld  update %lld\n",
             C, T, w ? VJX_DB_THREADS - 1 : 0, vjx_db_cycles[w][0], vjx_db_cycles[w][1], vjx_db_cycles[w][2], vjx_db_cycles[w][3],
             vjx_db_cycles[w][4], vjx_db_cycles[w][5], vjx_db_cycles[w][6], vjx_db_cycles[w][7]);
      for (int q = 0; q < 8; ++q) vjx_db_cycles[w][q] = 0;
    }
  }
  // ---- back substitution, CTA 0: thread t owns row t; blocks of 16 unknowns, last block first
  double* const ys = sm + D::OFF_YS;
  double* const xs = sm + D::OFF_XS;
  const int t = threadIdx.x;
  const bool row = t < T;
  const double* const urow = lu.dense + (size_t)(row ? t : 0) * ldg;
  if (row) {
    const double yv = urow[T];
    ys[t] = yv;
    lu.y[lu.dense_start + t] = yv;
  }
  const double myinv = row ? lu.inv_diag[lu.dense_start + t] : 0.0;   // 1 / u(t,t), written by the elimination
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 4] = global_ns();
  constexpr int BS = 16;
  const int nblk = (T + BS - 1) / BS;
  double un[BS];   // this row's entries in the columns of the block being solved (prefetched one block ahead)
  {
    const int k0 = BS * (nblk - 1);
#pragma unroll
    for (int m = 0; m < BS; ++m) un[m] = (row && k0 + m < T && k0 + m >= t) ? urow[k0 + m] : 0.0;
  }
  __syncthreads();
  double xmine = 0.0;
  for (int B = nblk - 1; B >= 0; --B) {
    const int k0 = BS * B;
    double uc[BS];
#pragma unroll
    for (int m = 0; m < BS; ++m) uc[m] = un[m];
    if (B > 0) {
      const int kn = k0 - BS;
#pragma unroll
      for (int m = 0; m < BS; ++m) un[m] = (row && kn + m >= t) ? urow[kn + m] : 0.0;
    }
    if (wi == (B >> 1)) {
      // the warp that holds rows k0 .. k0 + 15 (lanes lb .. lb + 15) solves the diagonal block
      const int lb = (B & 1) * BS;
      const bool mine = lane >= lb && lane < lb + BS;
      double yv = (mine && t < T) ? ys[t] : 0.0;
#pragma unroll
      for (int m = BS - 1; m >= 0; --m) {
        if (k0 + m < T) {
          const int src = lb + m;
          const double xk = __shfl_sync(0xffffffffu, yv * myinv, src);
          if (mine && lane < src) yv = __fma_rn(-uc[m], xk, yv);
          if (lane == src) xmine = xk;
        }
      }
      if (mine && t < T) xs[lane - lb] = xmine;
    }
    __syncthreads();
    if (row && t < k0) {
      double yv = ys[t];
#pragma unroll
      for (int m = BS - 1; m >= 0; --m)
        if (k0 + m < T) yv = __fma_rn(-uc[m], xs[m], yv);
      ys[t] = yv;
    }
    __syncthreads();
  }
  if (row) {
    const int i = lu.dense_start + t;
    st_pub_f64(lu.xp + i, xmine);
    x_out[lu.perm_c[i]] = xmine;
    if (lu.ts) lu.ts[i] = global_ns();
  }
  if (lu.ts && threadIdx.x == 0) lu.ts[lu.n + 5] = global_ns();
}

// Deterministic warp sum: fixed butterfly order.
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The chain tail of the back solve passes values through shared memory (same sentinel).
#define VJX_TAIL_MAX 4096
// The back-solve kernel runs 1024-thread blocks: the chain tail is walked by ONE block whose warps
// take the tail rows round-robin, and a row costs a warp ~10 us of (mostly latency) time, so the tail
// advances at (row time) / (warps per block) per row - 32 warps instead of 8 is a 4x.
#define VJX_SOLVE_THREADS 1024

// Dot product of one row part with already solved unknowns.  `tid` is the thread's index among the
// STRIDE threads that share the row (lane for a warp, threadIdx.x for a whole block).  Four
// (column, value) pairs are loaded before the first dependency wait so that the matrix loads overlap
// the waits.  The waits are CONVERGENT: all lanes of a warp poll together until every lane's
// dependency of the current slot has arrived.  Letting each lane spin on its own (lanes parked in
// different unrolled spin loops) costs ~3,000 cycles per chain link on sm_100a - the scheduler keeps
// running one divergent group's loop and the lane whose value has arrived is not looked at
// (profiles/r1_lu_notes.md).  `peek(c, &x)` returns true once x_c is available.
template <int STRIDE, typename Peek>
__device__ __forceinline__ double row_dot(const LUDev& lu, int begin, int end, int tid, NRStat* st, int errbit, Peek peek) {
  double part = 0.0;
  // highest columns first: in the back solve they are the unknowns that become available first, so a long
  // row consumes its early dependencies while it waits and is left with one batch when the last one arrives
  const int n_batches = (end - begin + 4 * STRIDE - 1) / (4 * STRIDE);
  for (int bi = n_batches - 1; bi >= 0; --bi) {
    const int base = begin + bi * 4 * STRIDE;
    int c[4];
    double v[4], x[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = base + tid + u * STRIDE;
      c[u] = p < end ? lu.col[p] : -1;
      v[u] = p < end ? lu.val[p] : 0.0;
    }
    // request all four unknowns of the batch at once; only those still missing are polled again
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x[u] = 0.0;
      ok[u] = c[u] < 0 || peek(c[u], &x[u]);
    }
#pragma unroll
    for (int u = 3; u >= 0; --u) {
      if (base + u * STRIDE >= end) continue;   // uniform: no thread has an entry in this slot
      unsigned spins = 0;
      while (true) {
        if (__all_sync(0xffffffffu, ok[u])) break;
        if (!ok[u]) ok[u] = peek(c[u], &x[u]);
        if (++spins > VJX_SPIN_LIMIT) { atomicOr(&st->error, errbit); break; }
      }
      if (c[u] >= 0) part += v[u] * x[u];
    }
  }
  return part;
}

// Back solve x = U^-1 y, scattered through the column permutation.  The chain comes first (the
// trailing block of U): block 0 walks rows [tail_start, n) in reverse through shared memory and
// publishes every x for the wide remainder, which the other blocks solve by ticket in level order,
// polling x itself.  The last n_long_u blocks each take one very long row (a rail kept as
// column-singleton pivot: one entry per attached device).
__global__ void __launch_bounds__(VJX_SOLVE_THREADS) k_lu_usolve(LUDev lu, double* __restrict__ x_out, NRStat* st) {
  __shared__ double xtail[VJX_TAIL_MAX];
  const int lane = threadIdx.x & 31;
  const int n = lu.n;
  const int n_tail = n - lu.tail_start;
  auto peek_x = [&](int c, double* x) {
    const ull bits = ld_poll_u64(lu.xp + c);
    *x = bits_f64(bits);
    return bits != VJX_SENTINEL;
  };
  if (blockIdx.x == 0 && n_tail > 0) {
    // the dense trailing block is already solved (k_lu_dense_tail): its x is preloaded
    const int n_done = min(lu.n_dense, n_tail);
    for (int t = threadIdx.x; t < n_tail; t += blockDim.x) {
      if (t < n_done) xtail[t] = lu.xp[n - 1 - t];
      else reinterpret_cast<ull*>(xtail)[t] = VJX_SENTINEL;
    }
    __syncthreads();
    for (int t = n_done + (threadIdx.x >> 5); t < n_tail; t += (blockDim.x >> 5)) {
      const int i = n - 1 - t;
      const int ub = lu.diag[i] + 1, re = lu.ptr[i + 1];
      double yi = 0.0, di = 0.0;
      int pc = 0;
      if (lane == 0) {
        yi = lu.y[i];
        di = lu.inv_diag[i];
        pc = lu.perm_c[i];
      }
      // a tail row's U entries point into the tail
      double part = row_dot<32>(lu, ub, re, lane, st, 8, [&](int c, double* x) {
        const ull bits = *reinterpret_cast<const volatile ull*>(xtail + (n - 1 - c));
        *x = bits_f64(bits);
        return bits != VJX_SENTINEL;
      });
      if (re - ub > 1) part = warp_sum(part);
      else part = __shfl_sync(0xffffffffu, part, 0);
      if (lane == 0) {
        const double xi = (yi - part) * di;
        *reinterpret_cast<volatile double*>(xtail + t) = xi;
        st_pub_f64(lu.xp + i, xi);
        x_out[pc] = xi;
        if (lu.ts) lu.ts[i] = global_ns();
      }
    }
    return;
  }
  if ((int)blockIdx.x >= (int)gridDim.x - lu.n_long_u) {
    double* sh = xtail;
    const int i = lu.long_u[blockIdx.x - (gridDim.x - lu.n_long_u)];
    const int ub = lu.diag[i] + 1, re = lu.ptr[i + 1];
    double part = row_dot<VJX_SOLVE_THREADS>(lu, ub, re, threadIdx.x, st, 8, peek_x);
    part = block_sum_256(part, sh);
    if (threadIdx.x == 0) {
      const double xi = (lu.y[i] - part) * lu.inv_diag[i];
      st_pub_f64(lu.xp + i, xi);
      x_out[lu.perm_c[i]] = xi;
      if (lu.ts) lu.ts[i] = global_ns();
    }
    return;
  }
  // Rows without a U part (the peeled singletons: 20 k of c6288's 25 k rows, 533 k of mul64's 666 k) are x = y / pivot
  // with everything known when this kernel starts: one thread each, before the tickets.  Handed out by ticket, one
  // row per warp, they kept every warp busy for the first 18 us of c6288's back solve while the rows of the
  // dependency chain waited for a ticket (profiles/r2_ab timeline).  Same value: (y - 0) * (1 / pivot).
  {
    const int first = n_tail > 0 ? 1 : 0;
    const int nb = (int)gridDim.x - lu.n_long_u - first;
    for (int t = ((int)blockIdx.x - first) * (int)blockDim.x + (int)threadIdx.x; t < lu.n_flat_u; t += nb * (int)blockDim.x) {
      const int i = lu.order_u[t];
      const double xi = lu.y[i] * lu.inv_diag[i];
      st_pub_f64(lu.xp + i, xi);
      x_out[lu.perm_c[i]] = xi;
      if (lu.ts) lu.ts[i] = global_ns();
    }
  }
  while (true) {
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(&st->ticket_u, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= (unsigned)(lu.n_bulk_u - lu.n_flat_u)) break;
    const int i = lu.order_u[lu.n_flat_u + tk];
    const int ub = lu.diag[i] + 1, re = lu.ptr[i + 1];
    double yi = 0.0, di = 0.0;
    int pc = 0;
    if (lane == 0) {
      yi = lu.y[i];
      di = lu.inv_diag[i];
      pc = lu.perm_c[i];
    }
    double part = row_dot<32>(lu, ub, re, lane, st, 8, peek_x);
    if (re - ub > 1) part = warp_sum(part);
    else part = __shfl_sync(0xffffffffu, part, 0);
    if (lane == 0) {
      const double xi = (yi - part) * di;
      st_pub_f64(lu.xp + i, xi);
      x_out[pc] = xi;
      if (lu.ts) lu.ts[i] = global_ns();
    }
  }
}

// ------------------------------------------------------------------------------------------
// NR control (reference: _nr_body_common vajax/analysis/solver_factories.py:230-325)
// ------------------------------------------------------------------------------------------
__global__ void k_nr_begin(NRStat* st, double* ctl, int max_iter, int reset_iter) {
  if (reset_iter) {
    st->iter = 0;
    st->converged = 0;
    ctl[VJX_CTL_ITER] = 0.0;
  }
  st->max_iter = max_iter;
  st->cont = 1;
  st->error = 0;   // a watchdog bit of an earlier call must not fail this one
  nr_clear_iter(st);
}

__global__ void k_iter_begin(NRStat* st) { nr_clear_iter(st); }

struct UpdateArgs {
  int n_unknowns, n_vs, n_total, n_lim;
  const double* delta;
  double* X;
  const unsigned char* delta_mask;  // residual_mask (NOI excluded)
  const unsigned char* is_noi;      // per node (index in V)
  const double* limit_in;
  const double* limit_out;
  double damping, reltol, vntol, abstol, max_step;
  NRStat* st;
};

__global__ void k_delta_reduce(UpdateArgs a) {
  const int n_aug = a.n_unknowns + a.n_vs;
  int bad = 0;
  double mv = 0.0, ld = 0.0, lr = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_aug; i += gridDim.x * blockDim.x) {
    double d = a.delta[i];
    double cd = (i < a.n_unknowns) ? d * a.damping : d;
    double xr = a.X[i + 1];
    double tol = fmax(fabs(xr) * a.reltol, (i < a.n_unknowns) ? a.vntol : a.abstol);
    if (!a.delta_mask[i]) cd = 0.0;
    if (!(fabs(cd) < tol)) bad = 1;
    if (i < a.n_unknowns) {
      double ad = fabs(d);
      mv = (ad != ad) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(mv, ad);
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_lim; i += gridDim.x * blockDim.x) {
    ld = fmax(ld, fabs(a.limit_out[i] - a.limit_in[i]));
    lr = fmax(lr, fabs(a.limit_in[i]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    double t = __shfl_xor_sync(0xffffffffu, mv, o);
    mv = (mv != mv || t != t) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(mv, t);
    ld = fmax(ld, __shfl_xor_sync(0xffffffffu, ld, o));
    lr = fmax(lr, __shfl_xor_sync(0xffffffffu, lr, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicOr(&a.st->bad_delta, 1);
    if (mv != mv) atomicMax(&a.st->max_dv_bits, 0x7ff8000000000000ULL);
    else if (mv > 0.0) atomic_max_pos(&a.st->max_dv_bits, mv);
    if (ld > 0.0) atomic_max_pos(&a.st->lim_delta_bits, ld);
    if (lr > 0.0) atomic_max_pos(&a.st->lim_ref_bits, lr);
  }
}

__global__ void k_nr_update(UpdateArgs a, double* ctl, double* limit_state, cudaGraphConditionalHandle handle,
                            int use_handle) {
  const int n_aug = a.n_unknowns + a.n_vs;
  const double maxv = __longlong_as_double((long long)a.st->max_dv_bits);
  const double scale = fmin(1.0, a.max_step / fmax(maxv, 1e-30));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_aug; i += gridDim.x * blockDim.x) {
    double d = a.delta[i];
    double x = a.X[i + 1];
    if (i < a.n_unknowns) {
      x = x + d * scale * a.damping;
      if (a.is_noi[i + 1]) x = 0.0;
    } else {
      x = x + d;
    }
    a.X[i + 1] = x;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_lim; i += gridDim.x * blockDim.x)
    limit_state[i] = a.limit_out[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    NRStat* st = a.st;
    int it = st->iter;
    bool conv = (st->bad_res == 0) && (it == 0 || st->bad_delta == 0);
    if (a.n_lim > 0) {
      double ld = __longlong_as_double((long long)st->lim_delta_bits);
      double lref = fmax(__longlong_as_double((long long)st->lim_ref_bits) * a.reltol, a.vntol);
      conv = conv && (ld < lref) && (it >= 1);
    }
    st->converged = conv ? 1 : 0;
    st->iter = it + 1;
    ctl[VJX_CTL_ITER] = (double)(it + 1);
    int cont = (!conv && (it + 1) < st->max_iter) ? 1 : 0;
    st->cont = cont;
    if (use_handle) cudaGraphSetConditional(handle, cont);
    // state of the next iteration (the residual's reductions of the LAST iteration stay readable: max_f of the solve;
    // max_dv / the limit reductions are still being read by the other blocks of this launch: k_lu_load clears them)
    if (cont) { st->bad_res = 0; st->max_f_bits = 0ULL; }
    nr_clear_tickets(st);
  }
}

// k_delta_reduce + k_nr_update as ONE block (circuits up to VJX_NRU_MAX unknowns: the reductions stay in shared
// memory, no second launch; same values, same convergence logic).
#define VJX_NRU_THREADS 1024
#define VJX_NRU_MAX 65536
__global__ void __launch_bounds__(VJX_NRU_THREADS) k_nr_reduce_update(UpdateArgs a, double* ctl, double* limit_state,
                                                                      cudaGraphConditionalHandle handle, int use_handle) {
  __shared__ double s_mv[32], s_ld[32], s_lr[32];
  __shared__ int s_bad[32];
  const int n_aug = a.n_unknowns + a.n_vs;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  int bad = 0;
  double mv = 0.0, ld = 0.0, lr = 0.0;
  for (int i = tid; i < n_aug; i += VJX_NRU_THREADS) {
    const double d = a.delta[i];
    double cd = (i < a.n_unknowns) ? d * a.damping : d;
    const double xr = a.X[i + 1];
    const double tol = fmax(fabs(xr) * a.reltol, (i < a.n_unknowns) ? a.vntol : a.abstol);
    if (!a.delta_mask[i]) cd = 0.0;
    if (!(fabs(cd) < tol)) bad = 1;
    if (i < a.n_unknowns) {
      const double ad = fabs(d);
      mv = (ad != ad || mv != mv) ? qnan : fmax(mv, ad);
    }
  }
  for (int i = tid; i < a.n_lim; i += VJX_NRU_THREADS) {
    ld = fmax(ld, fabs(a.limit_out[i] - a.limit_in[i]));
    lr = fmax(lr, fabs(a.limit_in[i]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    const double t = __shfl_xor_sync(0xffffffffu, mv, o);
    mv = (mv != mv || t != t) ? qnan : fmax(mv, t);
    ld = fmax(ld, __shfl_xor_sync(0xffffffffu, ld, o));
    lr = fmax(lr, __shfl_xor_sync(0xffffffffu, lr, o));
  }
  if (lane == 0) { s_bad[wp] = bad; s_mv[wp] = mv; s_ld[wp] = ld; s_lr[wp] = lr; }
  __syncthreads();
  bad = 0; mv = 0.0; ld = 0.0; lr = 0.0;
  for (int q = 0; q < VJX_NRU_THREADS / 32; ++q) {
    bad |= s_bad[q];
    const double t = s_mv[q];
    mv = (mv != mv || t != t) ? qnan : fmax(mv, t);
    ld = fmax(ld, s_ld[q]);
    lr = fmax(lr, s_lr[q]);
  }
  // (max |dV| as the two-kernel path reads it back: a NaN anywhere is the canonical quiet NaN)
  const double scale = fmin(1.0, a.max_step / fmax(mv, 1e-30));
  for (int i = tid; i < n_aug; i += VJX_NRU_THREADS) {
    const double d = a.delta[i];
    double x = a.X[i + 1];
    if (i < a.n_unknowns) {
      x = x + d * scale * a.damping;
      if (a.is_noi[i + 1]) x = 0.0;
    } else {
      x = x + d;
    }
    a.X[i + 1] = x;
  }
  for (int i = tid; i < a.n_lim; i += VJX_NRU_THREADS) limit_state[i] = a.limit_out[i];
  if (tid == 0) {
    NRStat* st = a.st;
    st->bad_delta = bad;
    st->max_dv_bits = (mv != mv) ? 0x7ff8000000000000ULL : (mv > 0.0 ? (ull)__double_as_longlong(mv) : 0ULL);
    st->lim_delta_bits = ld > 0.0 ? (ull)__double_as_longlong(ld) : 0ULL;
    st->lim_ref_bits = lr > 0.0 ? (ull)__double_as_longlong(lr) : 0ULL;
    const int it = st->iter;
    bool conv = (st->bad_res == 0) && (it == 0 || bad == 0);
    if (a.n_lim > 0) {
      const double lref = fmax(lr * a.reltol, a.vntol);
      conv = conv && (ld < lref) && (it >= 1);
    }
    st->converged = conv ? 1 : 0;
    st->iter = it + 1;
    ctl[VJX_CTL_ITER] = (double)(it + 1);
    const int cont = (!conv && (it + 1) < st->max_iter) ? 1 : 0;
    st->cont = cont;
    if (use_handle) cudaGraphSetConditional(handle, cont);
    if (cont) { st->bad_res = 0; st->max_f_bits = 0ULL; }
    nr_clear_tickets(st);
  }
}

__global__ void k_dqdt(int n, const double* ctl, const double* Q, const double* Qp, const double* dq,
                       const double* q2, double* out) {
  const double c0 = ctl[VJX_CTL_C0], c1 = ctl[VJX_CTL_C1], d1 = ctl[VJX_CTL_D1], c2 = ctl[VJX_CTL_C2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double a = Qp ? Qp[i] : 0.0, b = dq ? dq[i] : 0.0, c = q2 ? q2[i] : 0.0;
    out[i] = c0 * Q[i] + c1 * a + d1 * b + c2 * c;
  }
}

__global__ void k_set_ctl(double* ctl, double c0, double c1, double d1, double c2, double gmin, double gshunt,
                          double iter, double diagreg) {
  ctl[VJX_CTL_C0] = c0;
  ctl[VJX_CTL_C1] = c1;
  ctl[VJX_CTL_D1] = d1;
  ctl[VJX_CTL_C2] = c2;
  ctl[VJX_CTL_GMIN] = gmin;
  ctl[VJX_CTL_GSHUNT] = gshunt;
  ctl[VJX_CTL_ITER] = iter;
  ctl[VJX_CTL_DIAGREG] = diagreg;
}

__global__ void k_copy(int n, const double* src, double* dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src ? src[i] : 0.0;
}
__global__ void k_fill(int n, double v, double* dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = v;
}

// ------------------------------------------------------------------------------------------
// Transient step control (reference: body_fn vajax/analysis/transient/full_mna.py:1840-2194,
// predict_voltage_jax / compute_lte_timestep_jax vajax/analysis/transient/adaptive.py:94-339)
// ------------------------------------------------------------------------------------------
struct StepArgs {
  TranScalars* ts;
  NRStat* st;
  double* ctl;
  vjx_options o;
  int n_total, n_unknowns, n_vs, n_is, n_ext, n_lim, H;
  int n_out;             // columns of V_out: n_ext, or the length of the save list
  const int* save_idx;   // node index (into X) per V_out column; null = the first n_ext nodes
  const vjx_source_desc* vsrc_desc;
  const vjx_source_desc* isrc_desc;
  const double* pwl_t;
  const double* pwl_v;
  double* vsrc;       // [n_vs] source values at t_next
  double* isrc;       // [n_is]
  double* X;          // accepted solution [n_total + n_vs]
  double* Xw;         // NR working vector
  double* Vpred;      // [n_total]
  double* Vhist;      // [H][n_total] ring
  double* Vmax;       // [n_total]
  double* Q_prev;     // [n_unknowns]
  double* dQdt_prev;
  double* Q_prev2;
  double* Qw;         // Q from the final build
  double* dQdtw;
  double* Ivs;        // [n_vs]
  double* maxres_w;   // [n_unknowns]
  double* hist_mrc;   // [n_unknowns]
  double* res_tol_floor;
  double* limit_state;  // accepted
  double* limit_w;      // working (NR in/out)
  double* times_out;
  double* V_out;
  double* I_out;
  double* records;    // 8 doubles per attempt
  int max_records;
};

__global__ void k_pre_step(StepArgs a, cudaGraphConditionalHandle nr_handle, int use_handle) {
  TranScalars* s = a.ts;
  __shared__ double sh_tnext;
  if (threadIdx.x == 0) {
    const double dt_cur = s->dt;
    const double inv_dt = 1.0 / dt_cur;
    double c0, c1, d1, c2, eci;
    if (a.o.integration_method == 0) {
      c0 = inv_dt; c1 = -inv_dt; d1 = 0.0; c2 = 0.0; eci = -0.5;
    } else if (a.o.integration_method == 1) {
      c0 = 2.0 * inv_dt; c1 = -2.0 * inv_dt; d1 = -1.0; c2 = 0.0; eci = 1.0 / 24.0;
    } else {
      double dt_prev = s->dt_hist[0];
      double omega = dt_prev > 0 ? dt_cur / dt_prev : 1.0;
      double op1 = 1.0 + omega;
      c0 = (1.0 + 2.0 * omega) / (op1 * dt_cur);
      c1 = -op1 / dt_cur;
      d1 = 0.0;
      c2 = (omega * omega) / (op1 * dt_cur);
      eci = -2.0 / 9.0;
    }
    const double t_next = s->t + dt_cur;
    s->t_next = t_next;
    s->dt_cur = dt_cur;
    s->eci = eci;
    sh_tnext = t_next;
    s->can_predict = (s->warmup >= a.o.warmup_steps) && (s->hist_count >= 2);
    // predictor coefficients (adaptive.py:114-168)
    int order = min(s->hist_count - 1, a.o.max_order);
    order = max(order, 0);
    const double new_dt = dt_cur;
    double cum0 = s->dt_hist[0];
    double cum1 = cum0 + s->dt_hist[1];
    double tau1 = -(new_dt + cum0) / new_dt;
    double tau2 = -(new_dt + cum1) / new_dt;
    double ratio1 = new_dt / fmax(s->dt_hist[0], 1e-30);
    double tau0 = -1.0;
    double den0 = (tau0 - tau1) * (tau0 - tau2);
    double den1 = (tau1 - tau0) * (tau1 - tau2);
    double den2 = (tau2 - tau0) * (tau2 - tau1);
    den0 = fabs(den0) < 1e-30 ? 1e-30 : den0;
    den1 = fabs(den1) < 1e-30 ? 1e-30 : den1;
    den2 = fabs(den2) < 1e-30 ? 1e-30 : den2;
    double l0 = (0 - tau1) * (0 - tau2) / den0;
    double l1 = (0 - tau0) * (0 - tau2) / den1;
    double l2 = (0 - tau0) * (0 - tau1) / den2;
    double aa[3], err;
    if (order == 0) { aa[0] = 1.0; aa[1] = 0.0; aa[2] = 0.0; err = 1.0; }
    else if (order == 1) { aa[0] = 1.0 + ratio1; aa[1] = -ratio1; aa[2] = 0.0; err = -ratio1 * (1 + ratio1) / 2; }
    else { aa[0] = l0; aa[1] = l1; aa[2] = l2; err = -tau0 * tau1 * tau2 / 6; }
    for (int k = 0; k < 3; ++k) s->a[k] = (k <= order) ? aa[k] : 0.0;
    s->pred_err = err;
    // gshunt ramp (full_mna.py:1914-1920)
    double ramp = 1.0;
    if (a.o.gshunt_steps > 0) ramp = fmin(fmax((double)s->step_idx / (double)a.o.gshunt_steps, 0.0), 1.0);
    double gsh = a.o.gshunt_init + ramp * (a.o.gshunt_target - a.o.gshunt_init);
    s->gshunt = gsh;
    a.ctl[VJX_CTL_C0] = c0;
    a.ctl[VJX_CTL_C1] = c1;
    a.ctl[VJX_CTL_D1] = d1;
    a.ctl[VJX_CTL_C2] = c2;
    a.ctl[VJX_CTL_GMIN] = 1e-12;  // body_fn passes the literal 1e-12 (full_mna.py:1948)
    a.ctl[VJX_CTL_GSHUNT] = gsh;
    a.ctl[VJX_CTL_ITER] = 0.0;
    a.ctl[VJX_CTL_DIAGREG] = a.o.gmin;
    s->lte_bits = 0ULL;
    a.st->iter = 0;
    a.st->converged = 0;
    a.st->max_iter = a.o.tran_itl;
    a.st->cont = 1;
    nr_clear_iter(a.st);
    s->attempts += 1;
    if (use_handle) cudaGraphSetConditional(nr_handle, 1);
  }
  __syncthreads();
  const double t_next = sh_tnext;
  for (int v = threadIdx.x; v < a.n_vs; v += blockDim.x) a.vsrc[v] = eval_source(a.vsrc_desc[v], t_next, a.pwl_t, a.pwl_v);
  for (int v = threadIdx.x; v < a.n_is; v += blockDim.x) a.isrc[v] = eval_source(a.isrc_desc[v], t_next, a.pwl_t, a.pwl_v);
}

__global__ void k_predict(StepArgs a) {
  const TranScalars* s = a.ts;
  const int H = a.H, head = s->hist_head;
  const double a0 = s->a[0], a1 = s->a[1], a2 = s->a[2];
  const double* h0 = a.Vhist + (size_t)((head + 0) % H) * a.n_total;
  const double* h1 = a.Vhist + (size_t)((head + 1) % H) * a.n_total;
  const double* h2 = a.Vhist + (size_t)((head + 2) % H) * a.n_total;
  const int can = s->can_predict;
  const int nx = a.n_total + a.n_vs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    if (i < a.n_total) {
      double vp = a0 * h0[i] + a1 * h1[i] + a2 * h2[i];
      a.Vpred[i] = vp;
      a.Xw[i] = can ? vp : a.X[i];
    } else {
      a.Xw[i] = a.X[i];
    }
    if (i < a.n_unknowns) a.res_tol_floor[i] = fmax(a.hist_mrc[i] * a.o.lte_reltol, a.o.lte_abstol);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_lim; i += gridDim.x * blockDim.x)
    a.limit_w[i] = a.limit_state[i];
}

__global__ void k_lte(StepArgs a) {
  const TranScalars* s = a.ts;
  const double scale = s->eci / (s->eci - s->pred_err);
  double m = 0.0;
  bool isnan_ = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_total; i += gridDim.x * blockDim.x) {
    double vn = a.Xw[i], vp = a.Vpred[i];
    double lte = (vn - vp) * scale;
    double vc = fmax(fabs(vn), fabs(vp));
    double vs = fmax(a.Vmax[i], vc);
    double tol = a.o.lte_reltol * vs + a.o.lte_abstol;
    double r = fabs(lte) / tol;
    if (r != r) isnan_ = true; else m = fmax(m, r);
  }
  for (int o = 16; o > 0; o >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    isnan_ |= (bool)__shfl_xor_sync(0xffffffffu, (int)isnan_, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (isnan_) atomicMax(&a.ts->lte_bits, 0x7ff8000000000000ULL);
    else if (m > 0.0) atomic_max_pos(&a.ts->lte_bits, m);
  }
}

__global__ void k_step_decide(StepArgs a, cudaGraphConditionalHandle handle, int use_handle, int max_attempts_total) {
  TranScalars* s = a.ts;
  const NRStat* st = a.st;
  const vjx_options& o = a.o;
  const double dt_cur = s->dt_cur, t_next = s->t_next;
  const bool converged = st->converged != 0;
  const int iterations = st->iter;
  const bool can_predict = s->can_predict != 0;
  const double lte_norm = __longlong_as_double((long long)s->lte_bits);
  // compute_lte_timestep_jax (adaptive.py:333-339)
  int order = min(s->hist_count - 1, o.max_order);
  double factor = pow(o.lte_ratio / fmax(lte_norm, 1e-10), 1.0 / (double)(order + 1));
  if (lte_norm != lte_norm) factor = lte_norm;
  factor = fmin(fmax(factor, 0.1), o.grow_factor);
  if (lte_norm != lte_norm) factor = lte_norm;
  const double dt_lte = dt_cur * factor;
  const bool nr_failed = !converged;
  const bool at_min_dt = dt_cur <= o.min_dt;
  const bool nr_fail_min = nr_failed && at_min_dt;
  const bool lte_reject = (dt_cur / dt_lte > o.redo_factor) && can_predict && converged;
  const bool nr_reject = nr_failed && !at_min_dt;
  const bool force_accept = s->consec >= o.max_consecutive_rejects;
  const bool reject = (lte_reject || nr_reject) && !force_accept;
  const bool accept = !reject;
  const int consec = accept ? 0 : (lte_reject ? s->consec + 1 : s->consec);
  const double dt_from_lte = can_predict ? dt_lte : dt_cur;
  const double dt_nr = fmax(dt_cur * o.tran_ft, o.min_dt);
  const double last_good = s->hist_count > 0 ? s->dt_hist[0] : o.min_dt * 16.0;
  const double dt_rec = fmax(last_good / 8.0, o.min_dt * 2.0);
  double new_dt = nr_failed ? (at_min_dt ? dt_rec : dt_nr) : dt_from_lte;
  new_dt = fmin(fmax(new_dt, o.min_dt), s->max_dt);
  if (dt_from_lte != dt_from_lte && !nr_failed) new_dt = dt_from_lte;
  new_dt = fmin(new_dt, s->t_stop - t_next);
  const bool use_new = accept && !nr_fail_min;
  s->accept = accept;
  s->use_new = use_new;
  s->nr_fail_min = nr_fail_min;
  s->lte_reject = lte_reject;
  s->nr_failed = nr_failed;
  s->reject = reject;
  s->new_dt = new_dt;
  if (s->attempts <= a.max_records) {
    double* r = a.records + (size_t)(s->attempts - 1) * 8;
    r[0] = (double)s->step_idx;
    r[1] = t_next;
    r[2] = dt_cur;
    r[3] = (double)iterations;
    r[4] = accept ? 1.0 : 0.0;
    r[5] = lte_reject ? 1.0 : 0.0;
    r[6] = nr_failed ? 1.0 : 0.0;
    r[7] = lte_norm;
  }
  // scalar state update (vector state is committed by k_step_commit using the flags above)
  if (accept) {
    a.times_out[s->step_idx] = t_next;
    s->t = t_next;
    s->min_dt_used = fmin(s->min_dt_used, dt_cur);
    s->max_dt_used = fmax(s->max_dt_used, dt_cur);
  }
  s->total_nr += iterations;
  s->rejected += reject ? 1 : 0;
  s->consec = consec;
  s->dt = new_dt;
  // hist/step counters are bumped in k_step_commit's thread 0 after the vector commit is issued
  (void)handle; (void)use_handle; (void)max_attempts_total;
}

__global__ void k_step_commit(StepArgs a, cudaGraphConditionalHandle handle, int use_handle, int attempts_limit) {
  TranScalars* s = a.ts;
  const int accept = s->accept, use_new = s->use_new, nr_fail_min = s->nr_fail_min;
  const int H = a.H;
  const int new_head = (s->hist_head + H - 1) % H;
  const int row = s->step_idx;
  const int nx = a.n_total + a.n_vs;
  const int nvs1 = a.n_vs > 0 ? a.n_vs : 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    double xn = a.Xw[i];
    if (use_new) {
      a.X[i] = xn;
      if (i < a.n_total) {
        a.Vhist[(size_t)new_head * a.n_total + i] = xn;
        a.Vmax[i] = fmax(a.Vmax[i], fabs(xn));
      }
      if (i < a.n_unknowns) {
        double qp = a.Q_prev[i];
        a.Q_prev2[i] = qp;
        a.Q_prev[i] = a.Qw[i];
        a.dQdt_prev[i] = a.dQdtw[i];
        a.hist_mrc[i] = fmax(a.hist_mrc[i], a.maxres_w[i]);
      }
    }
    if (accept) {
      if (i < a.n_out) {
        const int src = a.save_idx ? a.save_idx[i] : i;   // Xw is never written here, X only when use_new
        a.V_out[(size_t)row * a.n_out + i] = use_new ? a.Xw[src] : a.X[src];
      }
      if (i < nvs1) a.I_out[(size_t)row * nvs1 + i] = (a.n_vs > 0 && !nr_fail_min) ? a.Ivs[i] : 0.0;
    }
  }
  if (use_new)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_lim; i += gridDim.x * blockDim.x)
      a.limit_state[i] = a.limit_w[i];
}

// Runs after k_step_commit (separate launch => ordered after all vector writes).
__global__ void k_step_finish(StepArgs a, cudaGraphConditionalHandle handle, int use_handle, int attempts_limit) {
  TranScalars* s = a.ts;
  if (s->use_new) {
    s->hist_head = (s->hist_head + a.H - 1) % a.H;
    s->dt_hist[3] = s->dt_hist[2];
    s->dt_hist[2] = s->dt_hist[1];
    s->dt_hist[1] = s->dt_hist[0];
    s->dt_hist[0] = s->dt_cur;
    s->hist_count = min(s->hist_count + 1, a.H);
    s->warmup += 1;
  }
  if (s->accept) s->step_idx += 1;
  // Only a watchdog bit (1 | 4 | 8: a dependency wait that never ended) stops the loop.  Bit 2 (zero / NaN pivot)
  // does not: the NaNs it produces make the NR solve fail, which the step control treats like the reference does -
  // reject, cut dt by tran_ft, carry on (full_mna.py:1990-2010); the bit stays set as a diagnostic in
  // stats["error_flags"].
  int cont = (s->t < s->t_stop) && (s->step_idx < s->max_steps) && (s->attempts < s->attempts_limit) &&
             ((a.st->error & (1 | 4 | 8)) == 0);
  s->cont = cont;
  if (use_handle) cudaGraphSetConditional(handle, cont);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    n = count;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(p, 0, count * sizeof(T));
    return e;
  }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
};

// Host owner of a LongSeg table (chunked long segments of the stamping kernels).
struct LongSegBuf {
  DevBuf<int> chunk_seg, chunk_begin, chunk_end, seg_id, first;
  DevBuf<unsigned> count;
  DevBuf<double> partial;
  int n_chunks = 0, n_long = 0;
  // seg: [n+1] segment pointers; ids: the long segments; width: partials per chunk
  cudaError_t build(const std::vector<int>& seg, const std::vector<int>& ids, int width) {
    std::vector<int> cs, cb, ce, first;
    for (size_t k = 0; k < ids.size(); ++k) {
      first.push_back((int)cs.size());
      for (int b = seg[ids[k]]; b < seg[ids[k] + 1]; b += VJX_CHUNK) {
        cs.push_back((int)k);
        cb.push_back(b);
        ce.push_back(std::min(b + VJX_CHUNK, seg[ids[k] + 1]));
      }
    }
    first.push_back((int)cs.size());
    n_chunks = (int)cs.size();
    n_long = (int)ids.size();
    std::vector<int> idv(ids);
    if (cs.empty()) { cs.push_back(0); cb.push_back(0); ce.push_back(0); }
    if (idv.empty()) idv.push_back(0);
    cudaError_t e;
    if ((e = chunk_seg.upload(cs)) != cudaSuccess) return e;
    if ((e = chunk_begin.upload(cb)) != cudaSuccess) return e;
    if ((e = chunk_end.upload(ce)) != cudaSuccess) return e;
    if ((e = seg_id.upload(idv)) != cudaSuccess) return e;
    if ((e = first.size() ? this->first.upload(first) : cudaSuccess) != cudaSuccess) return e;
    if ((e = count.alloc(std::max<size_t>(idv.size(), 1))) != cudaSuccess) return e;
    return partial.alloc(std::max<size_t>((size_t)n_chunks * width, 1));
  }
  LongSeg dev() const {
    LongSeg L;
    L.chunk_seg = chunk_seg.p; L.chunk_begin = chunk_begin.p; L.chunk_end = chunk_end.p;
    L.seg_id = seg_id.p; L.first = first.p; L.count = count.p; L.partial = partial.p; L.n_chunks = n_chunks;
    return L;
  }
};

struct ModelState {
  void* dl = nullptr;
  vjx_model_eval_fn eval = nullptr;
  vjx_model_set_const_fn set_const = nullptr;
  std::string plugin_key;   // non-empty: this handle holds the in-tree instance of the plugin
  int n_dev = 0, n_volt = 0, n_dev_params = 0, n_dev_cache = 0, R = 0, J = 0, n_lim = 0;
  int n_shared = 0, n_shared_cache = 0, n_simparams = 0;
  int limit_offset = 0;
  std::vector<double> simparam_defaults;
  DevBuf<int> vnode1, vnode2, jdest, vdest;
  DevBuf<double> dev_params, dev_cache;
};

struct vjx_handle {
  int device = 0;
  int n_total = 0, n_unknowns = 0, n_ext = 0, n_vs = 0, n_is = 0, n_aug = 0, nse = 0, total_lim = 0;
  vjx_options o{};
  std::vector<int> h_indptr, h_indices;
  long long lu_nnz_a = 0;                    // entries of A the factorisation reads (CSR slots minus structural zeros)
  std::vector<unsigned char> h_slot_flags;   // host copy: SLOT_NOI_ZERO slots are structural zeros for the LU
  std::vector<ModelState*> models;
  // stamping tables
  DevBuf<int> jseg, vseg, isrc_dest, att_ptr, att_idx, vs_p, vs_n;
  LongSegBuf long_slots, long_rows;
  DevBuf<signed char> att_sign;
  DevBuf<unsigned char> slot_flags, conv_mask, delta_mask, is_noi, noi_row_mask;
  DevBuf<float> slot_const;
  DevBuf<double> jsorted, vsorted;
  size_t n_jcontrib = 0, n_vcontrib = 0;
  // work vectors
  DevBuf<double> ctl, csr, f, delta, Xw, Qw, dQdtw, Ivs, maxres_w, limit_w, limit_out, zeros, abstol_floor;
  DevBuf<NRStat> st;
  // sources
  DevBuf<vjx_source_desc> vsrc_desc, isrc_desc;
  DevBuf<double> pwl_t, pwl_v, vsrc_val, isrc_val;
  // LU
  bool lu_ready = false;
  int lu_grid_div = 0;   // > 1: persistent LU kernels take 1 / lu_grid_div of the SM slots (vjx_lu_set_grid_div)
  double dense_reg = 0.0;   // != 0: the reference's dense linear-solve path, J + dense_reg * I (vjx_set_dense_regularization)
  VjxSymbolic sym;
  DevBuf<int> lu_ptr, lu_col, lu_diag, lu_src, lu_init, lu_kub, lu_order_u, lu_perm_r, lu_perm_c;
  long long ffi_id = 0;           // id handed out by vjx_handle_id (0 = none)
  cudaStream_t last_stream = 0;   // most recent stream an entry point was called with (vjx_destroy waits for it)
  DevBuf<double> lu_scratch, lu_dense;
  DevBuf<int> lu_dense_lstop;
  int lu_grid_load = 0;
  // k_lu_rows runs as one or two launches over consecutive ranges of order_f (wide low levels with a
  // small working vector and many warps per SM, then the narrow high levels with long rows)
  int lu_nseg = 0, lu_seg_t0[2] = {0, 0}, lu_seg_t1[2] = {0, 0}, lu_seg_rowbuf[2] = {0, 0}, lu_seg_grid[2] = {0, 0};
  DevBuf<int> lu_order_f, lu_long_u;
  // heavy rows of each launch: blocks [0, lu_seg_hv_cta) walk rows lu_hv_rows[lu_seg_hv0 .. lu_seg_hv1)
  int lu_seg_hv_cta[2] = {0, 0}, lu_seg_hv0[2] = {0, 0}, lu_seg_hv1[2] = {0, 0};
  DevBuf<int> lu_hv_prog, lu_hv_rows;
  DevBuf<long long> lu_hv_off;
  DevBuf<unsigned char> lu_hv_flag;
  int lu_grid_u = 0;
  DevBuf<long long> lu_upd_ptr;
  DevBuf<int> lu_upd_pos;
  DevBuf<double> lu_val, lu_inv, lu_y, lu_xp;
  LUDev lud{};
  int lu_grid_f = 0;
  // transient
  DevBuf<TranScalars> ts;
  DevBuf<double> X, Vpred, Vhist, Vmax, Q_prev, dQdt_prev, Q_prev2, hist_mrc, res_tol_floor, limit_state;
  DevBuf<double> times_out, V_out, I_out, records;
  DevBuf<int> save_idx;                // N4: node per V_out column (empty = all external nodes)
  int n_out = 0;                       // columns of V_out
  int max_steps = 0, max_records = 0, H = 4;
  // graphs
  cudaGraphExec_t g_iter = nullptr;    // one NR iteration
  cudaGraphExec_t g_final = nullptr;   // final build + dQdt
  cudaGraphExec_t g_step_pre = nullptr, g_step_post = nullptr;
  bool full_final_eval = false;        // VJX_FULL_FINAL_EVAL=1: the build at the converged X runs the full model kernel
  int dense_cluster = 0;               // CTAs of the blocked dense kernel (0 = round-1 single-CTA kernel)
  cudaGraphExec_t g_nrs = nullptr;     // stand-alone NR loop (WHILE node) of vjx_nr_solve_async, on staging buffers
  DevBuf<double> nrs_X, nrs_in;        // staging: X | vsrc, isrc, Q_prev, dQdt_prev, Q_prev2, res_tol_floor
  cudaGraphExec_t g_tran = nullptr;    // whole device-resident loop (WHILE nodes)
  int g_tran_limit = 0;
  cudaStream_t own_stream = nullptr;
  // the Jacobian stamp and the residual of an NR iteration both depend on the model kernels only: the residual runs
  // on a side stream between two events (a parallel branch once captured into the iteration graph)
  cudaStream_t fork_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
  void* pinned = nullptr;
  long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // bracket the loop launch inside vjx_tran_run
  double last_run_ms = 0.0, total_run_ms = 0.0;
  int sm_count = 148;
  ~vjx_handle();
};

vjx_handle::~vjx_handle() {
  for (auto* m : models) {
    if (m) {
      if (m->dl) plugin_close(m->dl, m->plugin_key);
      delete m;
    }
  }
  if (g_iter) cudaGraphExecDestroy(g_iter);
  if (g_final) cudaGraphExecDestroy(g_final);
  if (g_step_pre) cudaGraphExecDestroy(g_step_pre);
  if (g_step_post) cudaGraphExecDestroy(g_step_post);
  if (g_nrs) cudaGraphExecDestroy(g_nrs);
  if (g_tran) cudaGraphExecDestroy(g_tran);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (fork_stream) cudaStreamDestroy(fork_stream);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (ev_fork2) cudaEventDestroy(ev_fork2);
  if (ev_join2) cudaEventDestroy(ev_join2);
  if (pinned) cudaFreeHost(pinned);
}

static inline cudaStream_t use_stream(vjx_handle* h, vjx_stream stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (h) h->last_stream = s;
  return s;
}

static inline int grid_for(const vjx_handle* h, long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)h->sm_count * 8;
  if (g < 1) g = 1;
  return (int)std::min(g, cap);
}

static int csr_find(const std::vector<int>& indptr, const std::vector<int>& indices, int r, int c) {
  auto b = indices.begin() + indptr[r], e = indices.begin() + indptr[r + 1];
  auto it = std::lower_bound(b, e, c);
  if (it == e || *it != c) return -1;
  return (int)(it - indices.begin());
}

// A model plugin keeps the parameters shared by all instances in ITS OWN __constant__ bank, so two live
// circuit handles (corners of a sweep: different temperature / process values) must not share one loaded
// instance of the plugin.  The first handle that asks for a path gets the in-tree file itself; while it
// is in use, every further handle loads a private copy (a temporary file, unlinked once mapped), which
// the CUDA runtime registers as a separate module with its own constants.
static std::mutex g_plugin_mu;
static std::set<std::string> g_plugin_in_use;

static void* plugin_open(const char* path, std::string& key) {
  key.clear();
  {
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    if (g_plugin_in_use.insert(path).second) key = path;
  }
  if (!key.empty()) {
    void* dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) {
      std::lock_guard<std::mutex> lk(g_plugin_mu);
      g_plugin_in_use.erase(key);
      key.clear();
    }
    return dl;
  }
  // private copy: $TMPDIR, /tmp, or next to the plugin itself
  const char* td = getenv("TMPDIR");
  std::string self(path);
  const size_t slash = self.rfind('/');
  const std::string dirs[3] = {td && *td ? td : "/tmp", "/tmp", slash == std::string::npos ? "." : self.substr(0, slash)};
  std::vector<char> name;
  int fd = -1;
  for (const std::string& dir : dirs) {
    const std::string tmpl = dir + "/vjx_plugin_XXXXXX.so";
    name.assign(tmpl.begin(), tmpl.end());
    name.push_back(0);
    fd = mkstemps(name.data(), 3);
    if (fd >= 0) break;
  }
  if (fd < 0) return nullptr;
  FILE* in = fopen(path, "rb");
  bool ok = in != nullptr;
  if (ok) {
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), in)) > 0)
      if (write(fd, buf, got) != (ssize_t)got) { ok = false; break; }
    fclose(in);
  }
  close(fd);
  void* dl = ok ? dlopen(name.data(), RTLD_NOW | RTLD_LOCAL) : nullptr;
  unlink(name.data());
  return dl;
}

static void plugin_close(void* dl, const std::string& key) {
  dlclose(dl);
  if (!key.empty()) {
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    g_plugin_in_use.erase(key);
  }
}

extern "C" const char* vjx_last_error(void) { return g_err.c_str(); }
extern "C" int vjx_version(void) { return 100; }
extern "C" int vjx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
extern "C" void vjx_clear_cache(void) {}

// ---- handle ids for FFI attributes (include/vajax_b200.h) ------------------------------------------
static std::mutex g_id_mu;
static std::map<long long, vjx_handle*> g_id_to_handle;
static long long g_next_id = 1;

extern "C" long long vjx_handle_id(vjx_handle* h) {
  if (!h) return 0;
  std::lock_guard<std::mutex> lk(g_id_mu);
  if (h->ffi_id == 0) {
    h->ffi_id = g_next_id++;
    g_id_to_handle[h->ffi_id] = h;
  }
  return h->ffi_id;
}

extern "C" vjx_handle* vjx_handle_from_id(long long id) {
  std::lock_guard<std::mutex> lk(g_id_mu);
  auto it = g_id_to_handle.find(id);
  return it == g_id_to_handle.end() ? nullptr : it->second;
}

extern "C" int vjx_sizes(vjx_handle* h, int* out8) {
  if (!h || !out8) return fail(VJX_ERR_INVALID, "null argument");
  out8[0] = h->n_total; out8[1] = h->n_ext; out8[2] = h->n_vs; out8[3] = h->n_is;
  out8[4] = h->n_unknowns; out8[5] = h->n_aug; out8[6] = h->nse; out8[7] = h->total_lim;
  return VJX_OK;
}

extern "C" void vjx_destroy(vjx_handle* h) {
  if (!h) return;
  if (h->ffi_id) {
    std::lock_guard<std::mutex> lk(g_id_mu);
    g_id_to_handle.erase(h->ffi_id);
  }
  cudaSetDevice(h->device);
  // not cudaDeviceSynchronize: another host thread may be capturing a graph for its own handle
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  cudaStreamSynchronize(h->last_stream);
  cudaStreamSynchronize(0);
  delete h;
}

extern "C" int vjx_create(const vjx_circuit_desc* d, int device, vjx_handle** out) {
  if (!d || !out) return fail(VJX_ERR_INVALID, "null descriptor");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(VJX_ERR_NODEVICE, "no CUDA device visible: libvajax_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(VJX_ERR_INVALID, "bad device index");
  CK(cudaSetDevice(device));
  std::unique_ptr<vjx_handle> h(new vjx_handle());
  h->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  h->n_total = d->n_total;
  h->n_unknowns = d->n_total - 1;
  h->n_ext = d->n_external;
  h->n_vs = d->n_vsources;
  h->n_is = d->n_isources;
  h->n_aug = h->n_unknowns + h->n_vs;
  h->nse = d->nse;
  h->o = d->opts;
  const int nu = h->n_unknowns, n_aug = h->n_aug, nse = h->nse;
  if (nu < 0 || nse <= 0 || !d->csr_indptr || !d->csr_indices) return fail(VJX_ERR_INVALID, "bad sizes");
  // ---- CSR validation, once (umfpack_ffi.cpp:60-93; full_mna.py:483-513)
  h->h_indptr.assign(d->csr_indptr, d->csr_indptr + n_aug + 1);
  h->h_indices.assign(d->csr_indices, d->csr_indices + nse);
  if (h->h_indptr[0] != 0 || h->h_indptr[n_aug] != nse) return fail(VJX_ERR_INVALID, "CSR indptr ends mismatch");
  for (int r = 0; r < n_aug; ++r) {
    if (h->h_indptr[r + 1] < h->h_indptr[r]) return fail(VJX_ERR_INVALID, "CSR indptr not monotonic");
    for (int p = h->h_indptr[r]; p < h->h_indptr[r + 1]; ++p) {
      int c = h->h_indices[p];
      if (c < 0 || c >= n_aug) return fail(VJX_ERR_INVALID, "CSR column index out of range");
      if (p > h->h_indptr[r] && c <= h->h_indices[p - 1]) return fail(VJX_ERR_INVALID, "CSR columns not strictly sorted");
    }
  }
  // ---- Jacobian contributions: count per slot in the reference's scatter order
  std::vector<int> jcount(nse, 0);
  std::vector<std::vector<int>> slots(d->n_models);
  h->total_lim = 0;
  for (int m = 0; m < d->n_models; ++m) {
    const vjx_model_desc& md = d->models[m];
    auto* ms = new ModelState();
    h->models.push_back(ms);
    ms->n_dev = md.n_dev; ms->n_volt = md.n_volt; ms->n_dev_params = md.n_dev_params;
    ms->n_dev_cache = md.n_dev_cache; ms->R = md.n_res; ms->J = md.n_jac; ms->n_lim = md.n_lim;
    ms->n_shared = md.n_shared; ms->n_shared_cache = md.n_shared_cache; ms->n_simparams = md.n_simparams;
    ms->limit_offset = h->total_lim;
    h->total_lim += md.n_lim * md.n_dev;
    auto& sl = slots[m];
    sl.assign((size_t)md.n_dev * md.n_jac, -1);
    for (int dv = 0; dv < md.n_dev; ++dv)
      for (int e = 0; e < md.n_jac; ++e) {
        int r = md.jac_row[(size_t)dv * md.n_jac + e], c = md.jac_col[(size_t)dv * md.n_jac + e];
        if (r < 0 || c < 0) continue;
        if (r >= n_aug || c >= n_aug) return fail(VJX_ERR_INVALID, "stamp index out of range");
        int s = csr_find(h->h_indptr, h->h_indices, r, c);
        if (s < 0) return fail(VJX_ERR_INVALID, "device Jacobian entry outside the CSR pattern");
        sl[(size_t)dv * md.n_jac + e] = s;
        jcount[s]++;
      }
  }
  std::vector<int> jseg(nse + 1, 0);
  for (int s = 0; s < nse; ++s) jseg[s + 1] = jseg[s] + jcount[s];
  h->n_jcontrib = (size_t)jseg[nse];
  {
    std::vector<int> nxt(jseg.begin(), jseg.end() - 1);
    for (int m = 0; m < d->n_models; ++m) {
      const vjx_model_desc& md = d->models[m];
      std::vector<int> dest((size_t)md.n_dev * md.n_jac, -1);  // SoA (J, n_dev)
      for (int dv = 0; dv < md.n_dev; ++dv)
        for (int e = 0; e < md.n_jac; ++e) {
          int s = slots[m][(size_t)dv * md.n_jac + e];
          if (s >= 0) dest[(size_t)e * md.n_dev + dv] = nxt[s]++;
        }
      CK(h->models[m]->jdest.upload(dest));
    }
  }
  CK(h->jseg.upload(jseg));
  CK(h->jsorted.alloc(std::max<size_t>(h->n_jcontrib, 1)));
  {
    std::vector<int> ls;
    for (int s = 0; s < nse; ++s)
      if (jcount[s] > VJX_LONG_SEG) ls.push_back(s);
    CK(h->long_slots.build(jseg, ls, 1));
  }
  // ---- slot flags: diagonal regularisation, vsource incidence, NOI rows/cols
  std::vector<unsigned char> flags(nse, 0);
  std::vector<float> cst(nse, 0.0f);
  for (int i = 0; i < nu; ++i) {
    int s = csr_find(h->h_indptr, h->h_indices, i, i);
    if (s < 0) return fail(VJX_ERR_INVALID, "diagonal entry missing from the CSR pattern");
    flags[s] |= SLOT_DIAG;
  }
  for (int i = nu; i < h->n_aug; ++i) {   // present only when the caller put them into the pattern (dense-path mode)
    int s = csr_find(h->h_indptr, h->h_indices, i, i);
    if (s >= 0) flags[s] |= SLOT_BRANCH_DIAG;
  }
  for (int v = 0; v < h->n_vs; ++v) {
    int p = d->vs_node_p[v], n = d->vs_node_n[v];
    if (p < 0 || p >= h->n_total || n < 0 || n >= h->n_total) return fail(VJX_ERR_INVALID, "vsource node out of range");
    int br = nu + v;
    struct { int r, c; float val; } ent[4] = {{p - 1, br, 1.f}, {n - 1, br, -1.f}, {br, p - 1, 1.f}, {br, n - 1, -1.f}};
    for (int k = 0; k < 4; ++k) {
      if (ent[k].r < 0 || ent[k].c < 0) continue;
      int s = csr_find(h->h_indptr, h->h_indices, ent[k].r, ent[k].c);
      if (s < 0) return fail(VJX_ERR_INVALID, "vsource incidence entry missing from the CSR pattern");
      cst[s] += ent[k].val;
    }
  }
  std::vector<unsigned char> is_noi(h->n_total, 0), noi_row(n_aug, 0);
  for (int k = 0; k < d->n_noi; ++k) {
    int nd = d->noi_nodes[k];
    if (nd <= 0 || nd >= h->n_total) return fail(VJX_ERR_INVALID, "NOI node index out of range");
    is_noi[nd] = 1;
    noi_row[nd - 1] = 1;
  }
  if (d->n_noi > 0)
    for (int r = 0; r < n_aug; ++r)
      for (int p = h->h_indptr[r]; p < h->h_indptr[r + 1]; ++p) {
        int c = h->h_indices[p];
        if (noi_row[r]) flags[p] |= SLOT_NOI_ZERO;
        if (r < nu && noi_row[c]) flags[p] |= SLOT_NOI_ZERO;
        if (noi_row[r] && c == r) flags[p] |= SLOT_NOI_ONE;
      }
  h->h_slot_flags = flags;
  CK(h->slot_flags.upload(flags));
  CK(h->slot_const.upload(cst));
  CK(h->is_noi.upload(is_noi));
  CK(h->noi_row_mask.upload(noi_row));
  // ---- convergence masks (solver_factories.py:86-101, 1018-1036); negative indices wrap on the
  // n_unknowns-long node mask exactly like jnp's .at[]
  std::vector<unsigned char> conv(n_aug, 1), dmask(n_aug, 1);
  for (int k = 0; k < d->n_noi; ++k) dmask[d->noi_nodes[k] - 1] = 0;
  if (d->n_internal > 0) {
    for (int k = 0; k < d->n_internal; ++k) {
      int idx = d->internal_nodes[k] - 1;
      if (idx < 0) idx += nu;
      if (idx >= 0 && idx < nu) conv[idx] = 0;
    }
  } else {
    conv = dmask;
  }
  CK(h->conv_mask.upload(conv));
  CK(h->delta_mask.upload(dmask));
  // ---- vector contributions sorted by row: isources first, then models in order
  std::vector<int> vcount(std::max(nu, 1), 0);
  auto is_idx = [&](int node) { return node != 0 ? node - 1 : -1; };
  for (int i = 0; i < h->n_is; ++i) {
    int a = is_idx(d->is_node_p[i]), b = is_idx(d->is_node_n[i]);
    if (a >= nu || b >= nu) return fail(VJX_ERR_INVALID, "isource node out of range");
    if (a >= 0) vcount[a]++;
    if (b >= 0) vcount[b]++;
  }
  for (int m = 0; m < d->n_models; ++m) {
    const vjx_model_desc& md = d->models[m];
    for (size_t k = 0; k < (size_t)md.n_dev * md.n_res; ++k) {
      int r = md.res_idx[k];
      if (r >= nu) return fail(VJX_ERR_INVALID, "residual index out of range");
      if (r >= 0) vcount[r]++;
    }
  }
  std::vector<int> vseg(nu + 1, 0);
  for (int i = 0; i < nu; ++i) vseg[i + 1] = vseg[i] + vcount[i];
  h->n_vcontrib = (size_t)vseg[nu];
  {
    std::vector<int> nxt(vseg.begin(), vseg.end() - 1);
    std::vector<int> idest((size_t)std::max(2 * h->n_is, 1), -1);
    for (int i = 0; i < h->n_is; ++i) {
      int a = is_idx(d->is_node_p[i]), b = is_idx(d->is_node_n[i]);
      if (a >= 0) idest[2 * i] = nxt[a]++;
      if (b >= 0) idest[2 * i + 1] = nxt[b]++;
    }
    CK(h->isrc_dest.upload(idest));
    for (int m = 0; m < d->n_models; ++m) {
      const vjx_model_desc& md = d->models[m];
      std::vector<int> dest((size_t)md.n_dev * md.n_res, -1);  // SoA (R, n_dev)
      for (int dv = 0; dv < md.n_dev; ++dv)
        for (int r = 0; r < md.n_res; ++r) {
          int row = md.res_idx[(size_t)dv * md.n_res + r];
          if (row >= 0) dest[(size_t)r * md.n_dev + dv] = nxt[row]++;
        }
      CK(h->models[m]->vdest.upload(dest));
    }
  }
  CK(h->vseg.upload(vseg));
  CK(h->vsorted.alloc(std::max<size_t>(h->n_vcontrib * 4, 4)));
  {
    std::vector<int> lr;
    for (int i = 0; i < nu; ++i)
      if (vcount[i] > VJX_LONG_SEG) lr.push_back(i);
    CK(h->long_rows.build(vseg, lr, 5));
  }
  // ---- vsources attached to each node row, in the order of the concatenated [p..., n...] scatter
  {
    std::vector<int> cnt(std::max(nu, 1), 0);
    for (int v = 0; v < h->n_vs; ++v) {
      if (d->vs_node_p[v] > 0) cnt[d->vs_node_p[v] - 1]++;
      if (d->vs_node_n[v] > 0) cnt[d->vs_node_n[v] - 1]++;
    }
    std::vector<int> ap(nu + 1, 0);
    for (int i = 0; i < nu; ++i) ap[i + 1] = ap[i] + cnt[i];
    std::vector<int> ai(std::max(ap[nu], 1), 0), nxt(ap.begin(), ap.end() - 1);
    std::vector<signed char> as(std::max(ap[nu], 1), 0);
    for (int v = 0; v < h->n_vs; ++v)
      if (d->vs_node_p[v] > 0) { int q = nxt[d->vs_node_p[v] - 1]++; ai[q] = v; as[q] = 1; }
    for (int v = 0; v < h->n_vs; ++v)
      if (d->vs_node_n[v] > 0) { int q = nxt[d->vs_node_n[v] - 1]++; ai[q] = v; as[q] = -1; }
    CK(h->att_ptr.upload(ap));
    CK(h->att_idx.upload(ai));
    CK(h->att_sign.upload(as));
    std::vector<int> vp(d->vs_node_p, d->vs_node_p + h->n_vs), vn(d->vs_node_n, d->vs_node_n + h->n_vs);
    if (vp.empty()) { vp.push_back(0); vn.push_back(0); }
    CK(h->vs_p.upload(vp));
    CK(h->vs_n.upload(vn));
  }
  // ---- models: plugin + instance tables (SoA on the device)
  for (int m = 0; m < d->n_models; ++m) {
    const vjx_model_desc& md = d->models[m];
    ModelState* ms = h->models[m];
    ms->dl = plugin_open(md.plugin_path, ms->plugin_key);
    if (!ms->dl) return fail(VJX_ERR_PLUGIN, std::string("cannot load model plugin: ") + (dlerror() ? dlerror() : md.plugin_path));
    ms->eval = (vjx_model_eval_fn)dlsym(ms->dl, "vjx_model_eval");
    ms->set_const = (vjx_model_set_const_fn)dlsym(ms->dl, "vjx_model_set_const");
    if (!ms->eval || !ms->set_const) return fail(VJX_ERR_PLUGIN, "model plugin lacks vjx_model_eval/vjx_model_set_const");
    auto to_soa_i = [&](const int* src, int rows, int cols) {
      std::vector<int> o((size_t)rows * cols);
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) o[(size_t)c * rows + r] = src[(size_t)r * cols + c];
      return o;
    };
    auto to_soa_d = [&](const double* src, int rows, int cols) {
      std::vector<double> o((size_t)rows * cols);
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) o[(size_t)c * rows + r] = src[(size_t)r * cols + c];
      return o;
    };
    for (size_t k = 0; k < (size_t)md.n_dev * md.n_volt; ++k)
      if (md.vnode1[k] < 0 || md.vnode1[k] >= h->n_total || md.vnode2[k] < 0 || md.vnode2[k] >= h->n_total)
        return fail(VJX_ERR_INVALID, "voltage node index out of range");
    CK(ms->vnode1.upload(to_soa_i(md.vnode1, md.n_dev, md.n_volt)));
    CK(ms->vnode2.upload(to_soa_i(md.vnode2, md.n_dev, md.n_volt)));
    CK(ms->dev_params.upload(to_soa_d(md.dev_params, md.n_dev, md.n_dev_params)));
    CK(ms->dev_cache.upload(to_soa_d(md.dev_cache, md.n_dev, md.n_dev_cache)));
    ms->simparam_defaults.assign(md.simparam_defaults, md.simparam_defaults + md.n_simparams);
    int rc = ms->set_const(md.shared_params, md.n_shared, md.shared_cache, md.n_shared_cache,
                           md.simparam_defaults, md.n_simparams, 0);
    if (rc != 0) return fail(VJX_ERR_PLUGIN, "vjx_model_set_const failed (layout mismatch?) rc=" + std::to_string(rc));
  }
  // ---- sources
  {
    std::vector<vjx_source_desc> vd(d->vsrc, d->vsrc + h->n_vs), id(d->isrc, d->isrc + h->n_is);
    if (vd.empty()) vd.push_back(vjx_source_desc{});
    if (id.empty()) id.push_back(vjx_source_desc{});
    CK(h->vsrc_desc.upload(vd));
    CK(h->isrc_desc.upload(id));
    std::vector<double> pt(d->pwl_times, d->pwl_times + d->n_pwl), pv(d->pwl_values, d->pwl_values + d->n_pwl);
    if (pt.empty()) { pt.push_back(0); pv.push_back(0); }
    CK(h->pwl_t.upload(pt));
    CK(h->pwl_v.upload(pv));
    CK(h->vsrc_val.alloc(std::max(h->n_vs, 1)));
    CK(h->isrc_val.alloc(std::max(h->n_is, 1)));
  }
  // ---- work vectors
  const int nl = std::max(h->total_lim, 1);
  CK(h->ctl.alloc(VJX_CTL_COUNT));
  CK(h->csr.alloc(nse));
  CK(h->f.alloc(n_aug));
  CK(h->delta.alloc(n_aug));
  CK(h->Xw.alloc(h->n_total + h->n_vs));
  CK(h->Qw.alloc(std::max(nu, 1)));
  CK(h->dQdtw.alloc(std::max(nu, 1)));
  CK(h->Ivs.alloc(std::max(h->n_vs, 1)));
  CK(h->maxres_w.alloc(std::max(nu, 1)));
  CK(h->limit_w.alloc(nl));
  CK(h->limit_out.alloc(nl));
  CK(h->zeros.alloc(std::max(std::max(n_aug, nl), h->n_total + h->n_vs)));
  CK(h->st.alloc(1));
  CK(h->ts.alloc(1));
  CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->full_final_eval = getenv("VJX_FULL_FINAL_EVAL") != nullptr;
  if (!getenv("VJX_NO_FORK")) {
    CK(cudaStreamCreateWithFlags(&h->fork_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    if (!getenv("VJX_NO_EVAL_FORK")) {
      CK(cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
    }
  }
  CK(cudaHostAlloc(&h->pinned, 4096, cudaHostAllocDefault));
  CK(cudaStreamSynchronize(0));   // allocations are zero-filled on the legacy stream (not a device-wide sync: see vjx_destroy)
  *out = h.release();
  return VJX_OK;
}

extern "C" int vjx_update_model_data(vjx_handle* h, int model, const double* dev_params, const double* dev_cache,
                                     const double* shared_params, const double* shared_cache) {
  if (!h || model < 0 || model >= (int)h->models.size()) return fail(VJX_ERR_INVALID, "bad model index");
  CK(cudaSetDevice(h->device));
  ModelState* ms = h->models[model];
  auto soa = [&](const double* src, int rows, int cols) {
    std::vector<double> o((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) o[(size_t)c * rows + r] = src[(size_t)r * cols + c];
    return o;
  };
  if (dev_params && ms->n_dev_params > 0) {
    auto v = soa(dev_params, ms->n_dev, ms->n_dev_params);
    CK(cudaMemcpy(ms->dev_params.p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (dev_cache && ms->n_dev_cache > 0) {
    auto v = soa(dev_cache, ms->n_dev, ms->n_dev_cache);
    CK(cudaMemcpy(ms->dev_cache.p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (shared_params || shared_cache) {
    int rc = ms->set_const(shared_params, shared_params ? ms->n_shared : 0, shared_cache,
                           shared_cache ? ms->n_shared_cache : 0, nullptr, 0, 0);
    if (rc != 0) return fail(VJX_ERR_PLUGIN, "vjx_model_set_const failed");
  }
  return VJX_OK;
}

extern "C" int vjx_update_sources(vjx_handle* h, const vjx_source_desc* vsrc, const vjx_source_desc* isrc) {
  if (!h) return fail(VJX_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  if (vsrc && h->n_vs > 0) CK(cudaMemcpy(h->vsrc_desc.p, vsrc, sizeof(vjx_source_desc) * h->n_vs, cudaMemcpyHostToDevice));
  if (isrc && h->n_is > 0) CK(cudaMemcpy(h->isrc_desc.p, isrc, sizeof(vjx_source_desc) * h->n_is, cudaMemcpyHostToDevice));
  return VJX_OK;
}

// ---- launch helpers -----------------------------------------------------------------------
struct BuildPtrs {
  const double* X; const double* vsrc; const double* isrc; const double* Q_prev; const double* dQdt_prev;
  const double* Q_prev2; const double* limit_in; double* limit_out; const double* res_tol_floor;
  double* csr; double* f; double* Q; double* Ivs; double* maxres;
};

static int launch_eval(vjx_handle* h, const BuildPtrs& b, cudaStream_t s) {
  if (h->n_is > 0) {
    k_isrc<<<(2 * h->n_is + 127) / 128, 128, 0, s>>>(2 * h->n_is, h->isrc_dest.p, b.isrc, h->vsorted.p);
    h->launches++;
  }
  // The model with the most work runs on the caller's stream, the others (a handful of resistors or diodes beside ten
  // thousand MOSFETs: launches of a few microseconds each) on the side stream beside it - a parallel branch of the
  // captured iteration graph, like the residual beside the stamp.  They write disjoint parts of jsorted / vsorted.
  size_t main_m = 0;
  for (size_t m = 1; m < h->models.size(); ++m)
    if ((double)h->models[m]->n_dev * (h->models[m]->J + h->models[m]->R + 1) >
        (double)h->models[main_m]->n_dev * (h->models[main_m]->J + h->models[main_m]->R + 1)) main_m = m;
  const bool side = h->fork_stream && h->ev_fork2 && h->models.size() > 1;
  if (side) {
    CK(cudaEventRecord(h->ev_fork2, s));
    CK(cudaStreamWaitEvent(h->fork_stream, h->ev_fork2, 0));
  }
  for (size_t m = 0; m < h->models.size(); ++m) {
    ModelState* ms = h->models[m];
    cudaStream_t ms_stream = (side && m != main_m) ? h->fork_stream : s;
    VjxEvalArgs a;
    a.n_dev = ms->n_dev;
    a.V = b.X;
    a.vnode1 = ms->vnode1.p;
    a.vnode2 = ms->vnode2.p;
    a.dev_params = ms->dev_params.p;
    a.dev_cache = ms->dev_cache.p;
    a.ctl = h->ctl.p;
    a.limit_in = (ms->n_lim > 0 && b.limit_in) ? b.limit_in + ms->limit_offset : nullptr;
    a.limit_out = (ms->n_lim > 0 && b.limit_out) ? b.limit_out + ms->limit_offset : nullptr;
    a.jdest = ms->jdest.p;
    a.jsorted = (b.csr || h->full_final_eval) ? h->jsorted.p : nullptr;   // no Jacobian wanted: residual-side kernel
    a.vdest = ms->vdest.p;
    a.vsorted = h->vsorted.p;
    int rc = ms->eval(&a, ms_stream);
    if (rc != 0) return fail(VJX_ERR_CUDA, "model eval launch failed: " + std::string(cudaGetErrorString((cudaError_t)rc)));
    h->launches++;
  }
  if (side) {
    CK(cudaEventRecord(h->ev_join2, h->fork_stream));
    CK(cudaStreamWaitEvent(s, h->ev_join2, 0));
  }
  return VJX_OK;
}

static void launch_stamp(vjx_handle* h, double* csr, int enforce_noi, cudaStream_t s) {
  k_stamp_jacobian<<<grid_for(h, h->nse) + h->long_slots.n_chunks, 256, 0, s>>>(
      h->nse, h->jseg.p, h->jsorted.p, h->slot_flags.p, h->slot_const.p, h->ctl.p, enforce_noi, h->long_slots.dev(), csr,
      h->dense_reg);
  h->launches++;
}

static void launch_residual(vjx_handle* h, const BuildPtrs& b, int check, cudaStream_t s) {
  ResidualArgs a;
  a.n_unknowns = h->n_unknowns; a.n_vs = h->n_vs; a.n_total = h->n_total;
  a.vseg = h->vseg.p; a.vsorted = h->vsorted.p;
  a.att_ptr = h->att_ptr.p; a.att_idx = h->att_idx.p; a.att_sign = h->att_sign.p;
  a.vs_p = h->vs_p.p; a.vs_n = h->vs_n.p;
  a.X = b.X; a.vsrc = b.vsrc; a.Q_prev = b.Q_prev; a.dQdt_prev = b.dQdt_prev; a.Q_prev2 = b.Q_prev2;
  a.res_tol_floor = b.res_tol_floor; a.conv_mask = h->conv_mask.p; a.ctl = h->ctl.p;
  a.reltol = h->o.reltol; a.abstol = h->o.abstol; a.vntol = h->o.vntol;
  a.f = b.f; a.Q = b.Q; a.Ivs = b.Ivs; a.maxres = b.maxres; a.st = h->st.p; a.check = check;
  a.L = h->long_rows.dev();
  k_residual<<<grid_for(h, h->n_aug) + h->long_rows.n_chunks, 256, 0, s>>>(a);
  h->launches++;
}


// Dense trailing block: the blocked kernel on a cluster of C CTAs (h->dense_cluster; 0 = the round-1 kernel).
template <int C, int NR, int NC>
static cudaError_t launch_dense_blocked(const LUDev& lud, double* x, NRStat* st, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, 1, 1);
  cfg.blockDim = dim3(VJX_DB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = VjxDB<C, NR, NC>::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_lu_dense_blocked<C, NR, NC>, lud, x, st);
}
template <int C, int NR, int NC>
static cudaError_t prepare_dense_blocked() {
  return cudaFuncSetAttribute(k_lu_dense_blocked<C, NR, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              VjxDB<C, NR, NC>::SMEM_BYTES);
}
static void launch_dense(vjx_handle* h, const LUDev& lud, double* x, cudaStream_t s) {
  switch (h->dense_cluster) {
    case 1: launch_dense_blocked<1, 10, 5>(lud, x, h->st.p, s); break;
    case 2: launch_dense_blocked<2, 5, 5>(lud, x, h->st.p, s); break;
    case 4: if (lud.n_dense <= 159) launch_dense_blocked<4, 3, 5>(lud, x, h->st.p, s);
            else launch_dense_blocked<4, 5, 10>(lud, x, h->st.p, s);
            break;
    default: k_lu_dense_tail<<<1, VJX_DT_THREADS, VJX_DT_SMEM, s>>>(lud, x, h->st.p);
  }
}

static int ensure_lu(vjx_handle* h, const double* csr_dev, cudaStream_t s);

typedef void (*lu_rows_fn)(LUDev, const double*, const double*, double, const unsigned char*, NRStat*, unsigned*, int, int, int,
                           int, int, int, unsigned*);
static lu_rows_fn lu_rows_kernel(int rowbuf, bool two_launches, bool heavy) {
  if (rowbuf > 1024) return heavy ? k_lu_rows<VJX_UG, true, 1, true> : k_lu_rows<VJX_UG, true, 1, false>;
  if (two_launches) return heavy ? k_lu_rows<VJX_UG, false, 1, true> : k_lu_rows<VJX_UG, false, 1, false>;
  return heavy ? k_lu_rows<VJX_UG, false, 2, true> : k_lu_rows<VJX_UG, false, 2, false>;
}

static void launch_lu_factor(vjx_handle* h, const LUDev& lud, const double* csr, const double* b, double scale,
                             const unsigned char* zero_mask, double* x, cudaStream_t s) {
  k_lu_load<<<h->lu_grid_load, 256, 0, s>>>(lud, csr, b, scale, zero_mask, h->st.p);
  h->launches += 1;
  for (int g = 0; g < h->lu_nseg; ++g) {
    auto kfn = lu_rows_kernel(h->lu_seg_rowbuf[g], h->lu_nseg > 1 || h->lu_seg_rowbuf[g] > 1024, h->lu_seg_hv_cta[g] > 0);
    kfn<<<h->lu_seg_grid[g], 32 * VJX_FACTOR_WARPS, h->lu_seg_rowbuf[g] * VJX_FACTOR_WARPS * sizeof(double), s>>>(
        lud, csr, b, scale, zero_mask, h->st.p, g == 0 ? &h->st.p->ticket_f : &h->st.p->ticket_l, h->lu_seg_t0[g],
        h->lu_seg_t1[g], h->lu_seg_rowbuf[g], h->lu_seg_hv_cta[g], h->lu_seg_hv0[g], h->lu_seg_hv1[g],
        g == 0 ? &h->st.p->ticket_hf : &h->st.p->ticket_hl);
    h->launches += 1;
  }
  if (lud.n_dense > 0) {
    launch_dense(h, lud, x, s);
    h->launches += 1;
  }
}

// Refactorise on the fixed pattern with the forward solve folded in, then back solve into x.
static void launch_lu(vjx_handle* h, const double* csr, const double* b, double scale, const unsigned char* zero_mask,
                      double* x, cudaStream_t s) {
  launch_lu_factor(h, h->lud, csr, b, scale, zero_mask, x, s);
  k_lu_usolve<<<h->lu_grid_u, VJX_SOLVE_THREADS, 0, s>>>(h->lud, x, h->st.p);
  h->launches += 1;
}

static UpdateArgs make_update_args(vjx_handle* h, double* X, const double* limit_in, const double* limit_out) {
  UpdateArgs a;
  a.n_unknowns = h->n_unknowns; a.n_vs = h->n_vs; a.n_total = h->n_total; a.n_lim = h->total_lim;
  a.delta = h->delta.p; a.X = X; a.delta_mask = h->delta_mask.p; a.is_noi = h->is_noi.p;
  a.limit_in = limit_in; a.limit_out = limit_out;
  a.damping = h->o.nr_damping; a.reltol = h->o.reltol; a.vntol = h->o.vntol; a.abstol = h->o.abstol;
  a.max_step = h->o.max_step; a.st = h->st.p;
  return a;
}

// One NR iteration on (X, limit_w): reference _nr_body_common.
static int launch_nr_iteration(vjx_handle* h, const BuildPtrs& b, double* X, cudaGraphConditionalHandle handle,
                               int use_handle, cudaStream_t s) {
  // (no k_iter_begin: the loop's entry kernel / the previous iteration's update left the state cleared)
  int rc = launch_eval(h, b, s);
  if (rc) return rc;
  if (h->fork_stream) {
    CK(cudaEventRecord(h->ev_fork, s));
    CK(cudaStreamWaitEvent(h->fork_stream, h->ev_fork, 0));
    launch_residual(h, b, 1, h->fork_stream);
    CK(cudaEventRecord(h->ev_join, h->fork_stream));
    launch_stamp(h, h->csr.p, 1, s);
    CK(cudaStreamWaitEvent(s, h->ev_join, 0));
  } else {
    launch_stamp(h, h->csr.p, 1, s);
    launch_residual(h, b, 1, s);
  }
  launch_lu(h, h->csr.p, h->f.p, -1.0, h->noi_row_mask.p, h->delta.p, s);
  UpdateArgs ua = make_update_args(h, X, b.limit_in, b.limit_out);
  // (opt-in: one launch less, 7.6 instead of 8.6 us timed alone, but the c6288 step got 5 % SLOWER with it -
  // 0.804 against 0.762 ms over 200 steps, gpurun_out/r2_z_variants.txt - so the two-kernel form stays the default)
  if (std::max(h->n_aug, h->total_lim) <= VJX_NRU_MAX && getenv("VJX_NRU_MERGE")) {
    k_nr_reduce_update<<<1, VJX_NRU_THREADS, 0, s>>>(ua, h->ctl.p, h->limit_w.p, handle, use_handle);
    h->launches += 1;
  } else {
    k_delta_reduce<<<grid_for(h, std::max(h->n_aug, h->total_lim)), 256, 0, s>>>(ua);
    k_nr_update<<<grid_for(h, std::max(h->n_aug, h->total_lim)), 256, 0, s>>>(ua, h->ctl.p, h->limit_w.p, handle, use_handle);
    h->launches += 2;
  }
  return VJX_OK;
}

static int upload_symbolic(vjx_handle* h) {
  VjxSymbolic& S = h->sym;
  const int n = S.n;
  CK(h->lu_ptr.upload(S.lu_ptr));
  CK(h->lu_col.upload(S.lu_col));
  CK(h->lu_diag.upload(S.lu_diag));
  CK(h->lu_src.upload(S.lu_src));
  CK(h->lu_perm_r.upload(S.perm_r));
  CK(h->lu_perm_c.upload(S.perm_c));
  std::vector<long long> pp(S.upd_ptr.begin(), S.upd_ptr.end());
  CK(h->lu_upd_ptr.upload(pp));
  {
    std::vector<int> up(S.upd_pos);
    if (up.empty()) up.push_back(0);
    CK(h->lu_upd_pos.upload(up));
  }
  // k_lu_load program and the per-L-entry start of U(k,:); rows without an L part come first in order_f
  int n_nol = 0, max_lrow = 0;
  {
    std::vector<int> init((size_t)std::max<int64_t>(S.nnz_lu, 1), -3), kub((size_t)std::max<int64_t>(S.nnz_lu, 1), 0);
    for (int i = 0; i < n; ++i) {
      const int rb = S.lu_ptr[i], dg = S.lu_diag[i], re = S.lu_ptr[i + 1];
      if (dg == rb) {
        ++n_nol;
        for (int p = rb; p < re; ++p) init[p] = S.lu_src[p] >= 0 ? S.lu_src[p] : -1;
      } else {
        max_lrow = std::max(max_lrow, re - rb);
        for (int p = rb; p < dg; ++p) kub[p] = S.lu_diag[S.lu_col[p]] + 1;
        for (int p = dg; p < re; ++p) init[p] = -2;
      }
    }
    for (int t = 0; t < n; ++t) {
      const int i = S.order_f[t];
      const bool nol = S.lu_diag[i] == S.lu_ptr[i];
      if (nol != (t < n_nol)) return fail(VJX_ERR_INVALID, "factorisation order does not start with the rows without an L part");
    }
    CK(h->lu_init.upload(init));
    CK(h->lu_kub.upload(kub));
  }
  CK(h->lu_val.alloc((size_t)S.nnz_lu));
  CK(h->lu_inv.alloc(n));
  CK(h->lu_y.alloc(n));
  CK(h->lu_xp.alloc(n));
  LUDev& L = h->lud;
  L.n = n;
  L.nnz = S.nnz_lu;
  L.ptr = h->lu_ptr.p; L.col = h->lu_col.p; L.diag = h->lu_diag.p; L.src = h->lu_src.p;
  L.init = h->lu_init.p; L.kub = h->lu_kub.p;
  L.perm_r = h->lu_perm_r.p; L.perm_c = h->lu_perm_c.p;
  L.ts = nullptr;
  L.upd_ptr = h->lu_upd_ptr.p; L.upd_pos = h->lu_upd_pos.p;
  {
    // static bytes the row kernels walk: scatter program + per-entry tables; prefetch when well inside L2
    const double static_bytes = 4.0 * (double)S.upd_pos.size() + (4.0 * 3 + 8.0) * (double)S.nnz_lu;
    L.n_upd_prefetch = static_bytes < 48e6 ? (long long)S.upd_pos.size() : 0;
  }
  CK(h->lu_order_f.upload(S.order_f));
  L.order_f = h->lu_order_f.p;
  L.n_nol = n_nol;
  L.max_lrow = max_lrow;
  // dense trailing block: chosen by the symbolic analysis (which also orders the L parts around it)
  const int n_dense = S.n_dense;
  const int ds = n - n_dense;
  {
    std::vector<int> lstop((size_t)std::max(n_dense, 1), 0);
    for (int t = 0; t < n_dense; ++t) {
      const int i = ds + t;
      int p = S.lu_ptr[i];
      while (p < S.lu_diag[i] && S.lu_col[p] < ds) ++p;
      lstop[t] = p;
      for (int q = p; q < S.lu_diag[i]; ++q)
        if (S.lu_col[q] < ds) return fail(VJX_ERR_INVALID, "L part of a dense-block row is not ordered bulk-first");
    }
    CK(h->lu_dense_lstop.upload(lstop));
    CK(h->lu_dense.alloc((size_t)std::max(n_dense, 1) * (n_dense + 1)));
  }
  L.dense_start = ds;
  L.n_dense = n_dense;
  L.dense_lstop = h->lu_dense_lstop.p;
  L.dense = h->lu_dense.p;
  if (n_dense > 0) {
    CK(cudaFuncSetAttribute(k_lu_dense_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, VJX_DT_SMEM));
    CK((prepare_dense_blocked<1, 10, 5>()));
    CK((prepare_dense_blocked<2, 5, 5>()));
    CK((prepare_dense_blocked<4, 3, 5>()));
    CK((prepare_dense_blocked<4, 5, 10>()));
    // which kernel: VJX_DT_CLUSTER = 1 / 2 / 4 CTAs of the blocked kernel (4 takes up to 319 rows), 0 = the
    // one-barrier-per-pivot kernel of round 1 (default until the blocked one beats it, profiles/r2_dense_notes.md)
    int c = getenv("VJX_DT_CLUSTER") ? atoi(getenv("VJX_DT_CLUSTER")) : 0;
    if (c != 0 && c != 1 && c != 2 && c != 4) c = 0;
    if (n_dense > VJX_DT_MAX) c = 4;
    h->dense_cluster = c;
  }
  // back solve: rows of the dense block are solved by k_lu_dense_tail; the chain tail walked by block 0
  // of k_lu_usolve is what is left of [tail_start, n) in front of it
  L.tail_start = std::min(S.tail_start, ds);
  std::vector<int> bulk_u, lg;
  // rows without a U part first (one thread each at the start of k_lu_usolve), then the others in level order
  L.n_flat_u = 0;
  if (!getenv("VJX_NO_FLAT_U")) {
    for (int i : S.order_u) if (i < L.tail_start && S.lu_ptr[i + 1] - S.lu_diag[i] - 1 == 0) bulk_u.push_back(i);
    L.n_flat_u = (int)bulk_u.size();
  }
  for (int i : S.order_u) if (i < L.tail_start && (L.n_flat_u == 0 || S.lu_ptr[i + 1] - S.lu_diag[i] - 1 > 0)) bulk_u.push_back(i);
  for (int i : S.long_u) if (i < L.tail_start) lg.push_back(i);
  const int n_long_u = (int)lg.size();
  {
    if (bulk_u.empty()) bulk_u.push_back(0);
    if (lg.empty()) lg.push_back(0);
    L.n_bulk_u = 0;
    for (int i : S.order_u) if (i < L.tail_start) L.n_bulk_u++;
    CK(h->lu_order_u.upload(bulk_u));
    CK(h->lu_long_u.upload(lg));
  }
  L.order_u = h->lu_order_u.p;
  L.long_u = h->lu_long_u.p;
  L.n_long_u = n_long_u;
  L.val = h->lu_val.p; L.inv_diag = h->lu_inv.p; L.y = h->lu_y.p; L.xp = h->lu_xp.p;
  // k_lu_rows: the shared-memory working vector of a warp holds the longest row of its range, up to a
  // cap; longer rows go through a global scratch.  When the long rows would cost occupancy (> 1024
  // entries), the wide low levels get their own launch with a small vector and many warps per SM.
  {
    const int kRowbufCap = 6144;   // doubles per warp: 4 warps x 48 KB = 192 KB per block at the cap
    auto rowlen = [&](int t) { const int i = S.order_f[t]; return S.lu_ptr[i + 1] - S.lu_ptr[i]; };
    int n_cut = n;
    if (max_lrow > 1024) {
      std::vector<int> cl(S.n_levels_l, 0);
      for (int i = 0; i < n; ++i) cl[S.level_l[i]]++;
      int lcut = S.n_levels_l;
      while (lcut > 1 && cl[lcut - 1] <= 128) --lcut;
      n_cut = n_nol;
      while (n_cut < n && S.level_l[S.order_f[n_cut]] < lcut) ++n_cut;
      if (n_cut - n_nol < 2048 || n_cut == n) n_cut = n;
    }
    h->lu_nseg = 0;
    int bounds[3] = {n_nol, n_cut, n};
    int max_grid = 1;
    // The row and back-solve kernels are persistent (tickets + polling) and by default sized to fill the GPU, which is
    // right for one transient; with several corners of a sweep in flight on one GPU each of them holding every SM slot
    // while it polls serialises the corners.  lu_grid_div > 1 shrinks their grids (tickets make any grid size safe).
    const int grid_div = std::max(1, h->lu_grid_div > 0 ? h->lu_grid_div : (getenv("VJX_LU_GRID_DIV") ? atoi(getenv("VJX_LU_GRID_DIV")) : 1));
    for (int sgi = 0; sgi < 2; ++sgi) {
      const int t0 = bounds[sgi], t1 = bounds[sgi + 1];
      if (t1 <= t0) continue;
      std::vector<int> lens;
      lens.reserve(t1 - t0);
      for (int t = t0; t < t1; ++t) lens.push_back(rowlen(t));
      std::sort(lens.begin(), lens.end());
      int want = lens.back();
      // a handful of outliers in a wide range should not set the vector for everyone: they use the scratch
      if (t1 - t0 >= 4096) want = std::max(lens[(size_t)((lens.size() - 1) * 0.999)], std::min(want, 512));
      const int rowbuf = std::max(64, std::min(kRowbufCap, (want + 31) / 32 * 32));
      const int smem = rowbuf * VJX_FACTOR_WARPS * (int)sizeof(double);
      int occ = 1;
      // (request size for the long rows, pipelined: 2 / 4 / 8 chunks -> mul64 1.78 / 1.54 / 1.98 ms)
      // ring slots of 2 chunks (64 U entries per prefetched step) cost 24 registers: worth it when one launch
      // holds everything and the heaviest rows set its duration (c6288: rows phase 113 -> 108 us), not when the
      // wide launch lives on occupancy (mul64: 1.54 -> 1.62 ms)
      auto kfn = lu_rows_kernel(rowbuf, n_cut < n, false);
      CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kRowbufCap * VJX_FACTOR_WARPS * (int)sizeof(double)));
      CK(cudaFuncSetAttribute(lu_rows_kernel(rowbuf, n_cut < n, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kRowbufCap * VJX_FACTOR_WARPS * (int)sizeof(double)));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, 32 * VJX_FACTOR_WARPS, smem));
      occ = std::max(1, occ);
      const int s_i = h->lu_nseg++;
      h->lu_seg_t0[s_i] = t0;
      h->lu_seg_t1[s_i] = t1;
      h->lu_seg_rowbuf[s_i] = rowbuf;
      h->lu_seg_grid[s_i] = (int)std::max<long long>(1, std::min<long long>((long long)h->sm_count * occ / grid_div,
                                                     ((long long)(t1 - t0) + VJX_FACTOR_WARPS - 1) / VJX_FACTOR_WARPS));
      max_grid = std::max(max_grid, h->lu_seg_grid[s_i]);
    }
    L.scratch = nullptr;
    if (max_lrow > 0) {
      CK(h->lu_scratch.alloc((size_t)max_grid * VJX_FACTOR_WARPS * (size_t)max_lrow));
      L.scratch = h->lu_scratch.p;
    }
    // Heavy rows (vjx_symbolic.h, Heavy): rows with at least VJX_HV_MIN k-steps of their own that fit the block's
    // shared memory go to the first blocks of their launch, one row per block.  The program costs ~4 bytes per
    // contribution / fill: the threshold doubles until it fits VJX_HV_MAX_MB (default 192).  OFF by default
    // (VJX_HV_MIN unset or 0): measured slower than the k-sequential warps on both benchmark circuits - c6288
    // 203 vs 144 us, mul64 2.0 vs 1.5 ms per refactorisation - because after the minimum-degree + depth ordering the
    // rows phase is bound by the chain of dependency levels (~1.9 us per level through L2), not by the k-steps of a
    // row, and a block with a barrier per group crosses a level later than a warp does (profiles/r2_heavy_rows.md).
    {
      int hv_min = getenv("VJX_HV_MIN") ? atoi(getenv("VJX_HV_MIN")) : 0;
      const double max_mb = getenv("VJX_HV_MAX_MB") ? atof(getenv("VJX_HV_MAX_MB")) : 192.0;
      std::vector<unsigned char> flag((size_t)std::max(n, 1), 0);
      std::vector<int> rows_all;
      int seg_n[2] = {0, 0};
      // Block rows (lu_block_row_role): the same blocks, k-sequential, no program; rows whose scatter program has at
      // least VJX_BR_WORK entries.  Takes precedence over the heavy-row path when both are asked for.
      const long long br_work = getenv("VJX_BR_WORK") ? atoll(getenv("VJX_BR_WORK")) : 0;
      while (true) {
        rows_all.clear();
        for (int g = 0; g < h->lu_nseg; ++g) {
          if (g + 1 < h->lu_nseg) { seg_n[g] = 0; continue; }   // (the wide launch of two keeps to warps)
          std::vector<int> r;
          if (br_work > 0) {
            for (int t = h->lu_seg_t0[g]; t < h->lu_seg_t1[g]; ++t) {
              const int i = S.order_f[t];
              if (S.lu_ptr[i + 1] - S.lu_ptr[i] > VJX_HV_WARPS * h->lu_seg_rowbuf[g]) continue;
              const int le = vjx_row_lend(S, i);
              if (le > S.lu_ptr[i] && (long long)(S.upd_ptr[(size_t)le] - S.upd_ptr[(size_t)S.lu_ptr[i]]) >= br_work) r.push_back(i);
            }
          } else {
            r = vjx_select_heavy(S, h->lu_seg_t0[g], h->lu_seg_t1[g], hv_min, VJX_HV_WARPS * h->lu_seg_rowbuf[g]);
          }
          seg_n[g] = (int)r.size();
          rows_all.insert(rows_all.end(), r.begin(), r.end());
        }
        if (br_work > 0) { S.heavy = VjxSymbolic::Heavy(); break; }
        vjx_build_heavy(S, rows_all);
        if (rows_all.empty() || 4e-6 * (double)S.heavy.prog.size() <= max_mb) break;
        hv_min *= 2;
      }
      std::vector<int> pos_of_row((size_t)n, -1);
      for (int t = 0; t < n; ++t) pos_of_row[S.order_f[t]] = t;
      for (int i : rows_all) flag[(size_t)pos_of_row[i]] = 1;
      int o = 0;
      for (int g = 0; g < 2; ++g) {
        h->lu_seg_hv0[g] = o;
        h->lu_seg_hv1[g] = o + (g < h->lu_nseg ? seg_n[g] : 0);
        o = h->lu_seg_hv1[g];
        h->lu_seg_hv_cta[g] = 0;
        if (g >= h->lu_nseg || seg_n[g] == 0) continue;
        // every heavy row a block of its own while that leaves two thirds of the resident blocks to the other rows
        // (the heavy variant of the kernel needs more registers: its own occupancy sizes the launch)
        int occ_hv = 1;
        auto kfn = lu_rows_kernel(h->lu_seg_rowbuf[g], h->lu_nseg > 1 || h->lu_seg_rowbuf[g] > 1024, true);
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_hv, kfn, 32 * VJX_FACTOR_WARPS,
                                                         h->lu_seg_rowbuf[g] * VJX_FACTOR_WARPS * (int)sizeof(double)));
        const int cap = std::max(1, h->sm_count * std::max(1, occ_hv) / grid_div);
        h->lu_seg_grid[g] = std::max(1, std::min(h->lu_seg_grid[g], cap));
        h->lu_seg_hv_cta[g] = std::max(1, std::min(seg_n[g], h->lu_seg_grid[g] / 3));
      }
      std::vector<int> prog(S.heavy.prog), rows(rows_all);
      std::vector<long long> off(S.heavy.off);
      const bool block_rows = br_work > 0;
      if (prog.empty()) prog.push_back(0);
      if (rows.empty()) rows.push_back(0);
      if (off.empty()) off.push_back(0);
      CK(h->lu_hv_prog.upload(prog));
      CK(h->lu_hv_rows.upload(rows));
      CK(h->lu_hv_off.upload(off));
      CK(h->lu_hv_flag.upload(flag));
      L.hv_prog = block_rows ? nullptr : h->lu_hv_prog.p;   // no program: the blocks run lu_block_row_role
      L.hv_off = h->lu_hv_off.p; L.hv_rows = h->lu_hv_rows.p; L.hv_flag = h->lu_hv_flag.p;
      if (getenv("VJX_HV_VERBOSE") && block_rows)
        fprintf(stderr, "[vjx] block rows %zu (work >= %lld); blocks %d + %d\n", rows_all.size(), br_work, h->lu_seg_hv_cta[0], h->lu_seg_hv_cta[1]);
      if (getenv("VJX_HV_VERBOSE"))
        fprintf(stderr, "[vjx] heavy rows %zu (min k-steps %d): k-steps %d groups %d records %d fill %.2f program %.1f MB; blocks %d + %d\n",
                rows_all.size(), hv_min, S.heavy.n_ksteps, S.heavy.n_groups, S.heavy.n_records,
                S.heavy.n_slots ? (double)S.heavy.n_contrib / (double)S.heavy.n_slots : 0.0, 4e-6 * (double)S.heavy.prog.size(),
                h->lu_seg_hv_cta[0], h->lu_seg_hv_cta[1]);
      S.heavy.prog.clear(); S.heavy.prog.shrink_to_fit();
    }
  }
  int occ_u = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, k_lu_usolve, VJX_SOLVE_THREADS, 0));
  occ_u = std::max(1, occ_u);
  h->lu_grid_load = (int)std::max<long long>(1, std::min<long long>((long long)h->sm_count * 8, (S.nnz_lu + 255) / 256));
  // back solve: one block per SM slot, minus room for the long-row blocks; block 0 is dedicated to the
  // chain tail, the other blocks take the tickets
  const int wps = VJX_SOLVE_THREADS / 32;
  const int grid_div_u = std::max(1, h->lu_grid_div > 0 ? h->lu_grid_div : (getenv("VJX_LU_GRID_DIV") ? atoi(getenv("VJX_LU_GRID_DIV")) : 1));
  int grid_s = (int)std::max<long long>(1, std::min<long long>((long long)h->sm_count * occ_u / grid_div_u - n_long_u, ((long long)n + wps - 1) / wps));
  if (L.tail_start < n && grid_s < 2) grid_s = 2;
  h->lu_grid_u = grid_s + n_long_u;
  h->lu_ready = true;
  return VJX_OK;
}

static int ensure_lu(vjx_handle* h, const double* csr_dev, cudaStream_t s) {
  if (h->lu_ready) return VJX_OK;
  std::vector<double> vals(h->nse);
  CK(cudaStreamSynchronize(s));
  CK(cudaMemcpy(vals.data(), csr_dev, sizeof(double) * h->nse, cudaMemcpyDeviceToHost));
  std::string err;
  double tol = h->o.pivot_tol > 0 ? h->o.pivot_tol : 1e-3;
  // rows of the dense trailing block: up to 319 with the 4-CTA cluster kernel (mul64's trailing block has 312),
  // 159 when VJX_DT_CLUSTER pins a smaller configuration
  int dense_cap = VJX_DT_MAX;
  if (getenv("VJX_DT_CLUSTER") && atoi(getenv("VJX_DT_CLUSTER")) == 4) dense_cap = VJX_DB_MAX;
  if (getenv("VJX_DT_ROWS")) dense_cap = -std::min(dense_cap, std::max(atoi(getenv("VJX_DT_ROWS")), 1));   // < 0: forced size
  if (getenv("VJX_NO_DENSE_TAIL")) dense_cap = 0;
  // The linear solve only ever sees the NOI-enforced matrix (solver_factories.py:1268-1276 precedes every
  // linear_solve): the slots that enforce_noi rewrites to zero are STRUCTURAL zeros.  Leaving them in the
  // pattern would keep every noise node of every MOSFET (a row reaching D G S B, a column reaching G S D) out
  // of the singleton peel and in the core of the factorisation.  The analysis runs on the filtered pattern;
  // lu_src is mapped back to the slots of the full CSR afterwards.
  bool filtered = false;
  for (unsigned char f : h->h_slot_flags) if ((f & SLOT_NOI_ZERO) && !(f & SLOT_NOI_ONE)) { filtered = true; break; }
  if (getenv("VJX_LU_KEEP_NOI_SLOTS")) filtered = false;
  int rc;
  if (filtered) {
    std::vector<int> f_indptr(h->n_aug + 1, 0), f_indices, f_slot;
    std::vector<double> f_vals;
    f_indices.reserve(h->nse); f_slot.reserve(h->nse); f_vals.reserve(h->nse);
    for (int r = 0; r < h->n_aug; ++r) {
      for (int p = h->h_indptr[r]; p < h->h_indptr[r + 1]; ++p) {
        const unsigned char f = h->h_slot_flags[p];
        if ((f & SLOT_NOI_ZERO) && !(f & SLOT_NOI_ONE)) continue;
        f_indices.push_back(h->h_indices[p]); f_slot.push_back(p); f_vals.push_back(vals[p]);
      }
      f_indptr[r + 1] = (int)f_indices.size();
    }
    rc = vjx_symbolic_analyze(h->n_aug, f_indptr.data(), f_indices.data(), f_vals.data(), tol, true, h->sym, err, dense_cap);
    if (rc == 0)
      for (int& src : h->sym.lu_src) if (src >= 0) src = f_slot[src];
    h->lu_nnz_a = (long long)f_indices.size();
  } else {
    rc = vjx_symbolic_analyze(h->n_aug, h->h_indptr.data(), h->h_indices.data(), vals.data(), tol, true, h->sym, err, dense_cap);
    h->lu_nnz_a = h->nse;
  }
  if (rc != 0) return fail(VJX_ERR_SINGULAR, "LU analysis failed: " + err);
  return upload_symbolic(h);
}

extern "C" int vjx_lu_analyze(vjx_handle* h, const double* csr_data_dev, vjx_stream stream) {
  if (!h) return fail(VJX_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  h->lu_ready = false;
  if (h->g_iter) { cudaGraphExecDestroy(h->g_iter); h->g_iter = nullptr; }
  if (h->g_nrs) { cudaGraphExecDestroy(h->g_nrs); h->g_nrs = nullptr; }
  if (h->g_tran) { cudaGraphExecDestroy(h->g_tran); h->g_tran = nullptr; }
  return ensure_lu(h, csr_data_dev, (cudaStream_t)stream);
}

extern "C" int vjx_lu_stats(vjx_handle* h, double* out6 /* [10] */) {
  if (!h || !h->lu_ready) return fail(VJX_ERR_INVALID, "LU not analysed yet");
  out6[0] = (double)h->sym.nnz_lu;
  out6[1] = h->sym.flops;
  out6[2] = h->sym.n_levels_l;
  out6[3] = h->sym.n_levels_u;
  out6[4] = h->sym.max_row_len;
  out6[5] = h->sym.n;
  out6[6] = (double)h->sym.n_singletons;
  out6[7] = (double)h->sym.n_core;
  out6[8] = (double)h->sym.max_l_len;
  out6[9] = (double)h->sym.upd_pos.size();
  return VJX_OK;
}

static int check_device_error(vjx_handle* h, cudaStream_t s) {
  NRStat* hs = (NRStat*)h->pinned;
  CK(cudaMemcpyAsync(hs, h->st.p, sizeof(NRStat), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (hs->error & (1 | 4 | 8)) return fail(VJX_ERR_HANG, "dependency wait watchdog fired in the LU kernels");
  return VJX_OK;
}

extern "C" int vjx_lu_solve(vjx_handle* h, const double* csr_data_dev, const double* b_dev, double* x_dev, vjx_stream stream) {
  if (!h || !csr_data_dev || !b_dev || !x_dev) return fail(VJX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  int rc = ensure_lu(h, csr_data_dev, s);
  if (rc) return rc;
  k_iter_begin<<<1, 1, 0, s>>>(h->st.p);
  h->launches++;
  launch_lu(h, csr_data_dev, b_dev, 1.0, nullptr, x_dev, s);
  CK(cudaGetLastError());
  return VJX_OK;
}

extern "C" int vjx_build_system(vjx_handle* h, const double* X, const double* vsrc, const double* isrc,
                                const double* Q_prev, double c0, double gmin, double gshunt, double c1, double d1,
                                const double* dQdt_prev, double c2, const double* Q_prev2, const double* limit_in,
                                int nr_iteration, double* csr_out, double* f_out, double* Q_out, double* Ivs_out,
                                double* limit_out, double* maxres_out, vjx_stream stream) {
  if (!h || !X) return fail(VJX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  k_set_ctl<<<1, 1, 0, s>>>(h->ctl.p, c0, c1, d1, c2, gmin, gshunt, (double)nr_iteration, h->o.gmin);
  h->launches++;
  BuildPtrs b{};
  b.X = X; b.vsrc = vsrc ? vsrc : h->zeros.p; b.isrc = isrc ? isrc : h->zeros.p;
  b.Q_prev = Q_prev; b.dQdt_prev = dQdt_prev; b.Q_prev2 = Q_prev2;
  b.limit_in = limit_in ? limit_in : h->zeros.p;
  b.limit_out = limit_out ? limit_out : h->limit_out.p;
  b.res_tol_floor = nullptr;
  b.csr = csr_out; b.f = f_out ? f_out : h->f.p; b.Q = Q_out; b.Ivs = Ivs_out; b.maxres = maxres_out;
  int rc = launch_eval(h, b, s);
  if (rc) return rc;
  if (csr_out) launch_stamp(h, csr_out, 0, s);
  launch_residual(h, b, 0, s);
  CK(cudaGetLastError());
  return VJX_OK;
}

static void fill_build_ptrs_nr(vjx_handle* h, BuildPtrs& b, const double* X, const double* vsrc, const double* isrc,
                               const double* Q_prev, const double* dQdt_prev, const double* Q_prev2,
                               const double* rtf) {
  b.X = X; b.vsrc = vsrc ? vsrc : h->zeros.p; b.isrc = isrc ? isrc : h->zeros.p;
  b.Q_prev = Q_prev; b.dQdt_prev = dQdt_prev; b.Q_prev2 = Q_prev2;
  b.limit_in = h->limit_w.p; b.limit_out = h->limit_out.p; b.res_tol_floor = rtf;
  b.csr = h->csr.p; b.f = h->f.p; b.Q = nullptr; b.Ivs = nullptr; b.maxres = h->maxres_w.p;
}

// Host-polled NR loop: one graph-free chain per iteration + a 64-byte status read.
static int run_nr_host(vjx_handle* h, BuildPtrs& b, double* X, int max_iter, bool reset, cudaStream_t s) {
  k_nr_begin<<<1, 1, 0, s>>>(h->st.p, h->ctl.p, max_iter, reset ? 1 : 0);
  h->launches++;
  NRStat* hs = (NRStat*)h->pinned;
  if (!h->lu_ready) {
    // first factorisation of this handle: build the system once to analyse its values
    k_iter_begin<<<1, 1, 0, s>>>(h->st.p);
    int rc = launch_eval(h, b, s);
    if (rc) return rc;
    launch_stamp(h, h->csr.p, 1, s);
    rc = ensure_lu(h, h->csr.p, s);
    if (rc) return rc;
  }
  if (max_iter <= 0) return VJX_OK;
  while (true) {
    int rc = launch_nr_iteration(h, b, X, 0, 0, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(hs, h->st.p, sizeof(NRStat), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (hs->error & (1 | 4 | 8)) return fail(VJX_ERR_HANG, "dependency wait watchdog fired in the LU kernels");
    if (!hs->cont) break;
  }
  return VJX_OK;
}

extern "C" int vjx_nr_solve(vjx_handle* h, const double* X_init, const double* vsrc, const double* isrc,
                            const double* Q_prev, double c0, double gmin, double gshunt, double c1, double d1,
                            const double* dQdt_prev, double c2, const double* Q_prev2, const double* limit_in,
                            const double* res_tol_floor, int max_iterations, double* X_out, int* iterations,
                            int* converged, double* max_f, double* Q_out, double* dQdt_out, double* Ivs_out,
                            double* limit_out, double* maxres_out, vjx_stream stream) {
  if (!h || !X_init || !X_out) return fail(VJX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  const int nx = h->n_total + h->n_vs;
  if (max_iterations < 0) max_iterations = h->o.tran_itl;
  k_set_ctl<<<1, 1, 0, s>>>(h->ctl.p, c0, c1, d1, c2, gmin, gshunt, 0.0, h->o.gmin);
  if (X_out != X_init) k_copy<<<grid_for(h, nx), 256, 0, s>>>(nx, X_init, X_out);
  if (h->total_lim > 0) k_copy<<<grid_for(h, h->total_lim), 256, 0, s>>>(h->total_lim, limit_in, h->limit_w.p);
  h->launches += 3;
  BuildPtrs b{};
  fill_build_ptrs_nr(h, b, X_out, vsrc, isrc, Q_prev, dQdt_prev, Q_prev2, res_tol_floor);
  int rc = run_nr_host(h, b, X_out, max_iterations, true, s);
  if (rc) return rc;
  // final build at X_final for Q, I_vsource, max_res_contrib (solver_factories.py:689-713)
  BuildPtrs bf = b;
  bf.csr = nullptr; bf.f = h->f.p; bf.Q = Q_out ? Q_out : h->Qw.p; bf.Ivs = Ivs_out ? Ivs_out : h->Ivs.p;
  bf.maxres = maxres_out ? maxres_out : h->maxres_w.p;
  bf.limit_out = h->limit_out.p;
  rc = launch_eval(h, bf, s);
  if (rc) return rc;
  launch_residual(h, bf, 0, s);
  if (dQdt_out) {
    k_dqdt<<<grid_for(h, h->n_unknowns), 256, 0, s>>>(h->n_unknowns, h->ctl.p, bf.Q, Q_prev, dQdt_prev, Q_prev2, dQdt_out);
    h->launches++;
  }
  if (limit_out && h->total_lim > 0) {
    k_copy<<<grid_for(h, h->total_lim), 256, 0, s>>>(h->total_lim, h->limit_w.p, limit_out);
    h->launches++;
  }
  NRStat* hs = (NRStat*)h->pinned;
  CK(cudaMemcpyAsync(hs, h->st.p, sizeof(NRStat), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (iterations) *iterations = hs->iter;
  if (converged) *converged = hs->converged;
  if (max_f) { double mfv = INFINITY; if (max_iterations > 0) memcpy(&mfv, &hs->max_f_bits, sizeof(double)); *max_f = mfv; }
  CK(cudaGetLastError());
  return VJX_OK;
}


// ---- stream-ordered nr_solve (the form an XLA FFI handler needs: no host read, no synchronisation) ----------
__global__ void k_nr_scalars(const NRStat* st, int max_iterations, double* out4) {
  out4[0] = (double)st->iter;
  out4[1] = (double)st->converged;
  out4[2] = max_iterations > 0 ? __longlong_as_double((long long)st->max_f_bits) : INFINITY;
  out4[3] = (double)st->error;
}

struct NrsLayout { double *vsrc, *isrc, *Q_prev, *dQdt_prev, *Q_prev2, *rtf; };
static NrsLayout nrs_layout(vjx_handle* h) {
  const size_t nu = std::max(h->n_unknowns, 1), nv = std::max(h->n_vs, 1), ni = std::max(h->n_is, 1);
  NrsLayout L;
  double* p = h->nrs_in.p;
  L.vsrc = p; p += nv; L.isrc = p; p += ni; L.Q_prev = p; p += nu; L.dQdt_prev = p; p += nu; L.Q_prev2 = p; p += nu; L.rtf = p;
  return L;
}

static int build_nrs_graph(vjx_handle* h) {
  cudaGraph_t g;
  CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h_nr;
  CK(cudaGraphConditionalHandleCreate(&h_nr, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = h_nr;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t nr_node;
  CK(cudaGraphAddNode(&nr_node, g, nullptr, 0, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  NrsLayout L = nrs_layout(h);
  BuildPtrs b{};
  fill_build_ptrs_nr(h, b, h->nrs_X.p, L.vsrc, L.isrc, L.Q_prev, L.dQdt_prev, L.Q_prev2, L.rtf);
  cudaStream_t cs = h->own_stream;
  long long launches_before = h->launches;
  CK(cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  int rc = launch_nr_iteration(h, b, h->nrs_X.p, h_nr, 1, cs);
  cudaGraph_t tmp;
  CK(cudaStreamEndCapture(cs, &tmp));
  h->launches = launches_before;
  if (rc) return rc;
  CK(cudaGraphInstantiate(&h->g_nrs, g, 0));
  CK(cudaGraphDestroy(g));
  return VJX_OK;
}

extern "C" int vjx_nr_solve_async(vjx_handle* h, const double* X_init, const double* vsrc, const double* isrc,
                                  const double* Q_prev, double c0, double gmin, double gshunt, double c1, double d1,
                                  const double* dQdt_prev, double c2, const double* Q_prev2, const double* limit_in,
                                  const double* res_tol_floor, int max_iterations, double* X_out, double* scalars4_out,
                                  double* Q_out, double* dQdt_out, double* Ivs_out, double* limit_out, double* maxres_out,
                                  vjx_stream stream) {
  if (!h || !X_init || !X_out || !Q_prev || !scalars4_out) return fail(VJX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  const int nx = h->n_total + h->n_vs, nu = h->n_unknowns;
  if (max_iterations < 0) max_iterations = h->o.tran_itl;
  if (!h->nrs_X.p) {
    CK(h->nrs_X.alloc(nx));
    CK(h->nrs_in.alloc((size_t)std::max(h->n_vs, 1) + std::max(h->n_is, 1) + 4 * (size_t)std::max(nu, 1)));
  }
  NrsLayout L = nrs_layout(h);
  k_set_ctl<<<1, 1, 0, s>>>(h->ctl.p, c0, c1, d1, c2, gmin, gshunt, 0.0, h->o.gmin);
  k_copy<<<grid_for(h, nx), 256, 0, s>>>(nx, X_init, h->nrs_X.p);
  auto stage = [&](const double* src, double* dst, int n, double dflt) {
    if (n <= 0) return;
    if (src) k_copy<<<grid_for(h, n), 256, 0, s>>>(n, src, dst);
    else k_fill<<<grid_for(h, n), 256, 0, s>>>(n, dflt, dst);
    h->launches++;
  };
  stage(vsrc, L.vsrc, h->n_vs, 0.0);
  stage(isrc, L.isrc, h->n_is, 0.0);
  stage(Q_prev, L.Q_prev, nu, 0.0);
  stage(dQdt_prev, L.dQdt_prev, nu, 0.0);
  stage(Q_prev2, L.Q_prev2, nu, 0.0);
  stage(res_tol_floor, L.rtf, nu, h->o.abstol);
  if (h->total_lim > 0) stage(limit_in, h->limit_w.p, h->total_lim, 0.0);
  h->launches += 2;
  BuildPtrs b{};
  fill_build_ptrs_nr(h, b, h->nrs_X.p, L.vsrc, L.isrc, L.Q_prev, L.dQdt_prev, L.Q_prev2, L.rtf);
  k_nr_begin<<<1, 1, 0, s>>>(h->st.p, h->ctl.p, max_iterations, 1);
  h->launches++;
  if (!h->lu_ready) {
    // one-time set-up of the handle (symbolic analysis on the first Jacobian): this call, and only this one,
    // synchronises the stream
    k_iter_begin<<<1, 1, 0, s>>>(h->st.p);
    int rc = launch_eval(h, b, s);
    if (rc) return rc;
    launch_stamp(h, h->csr.p, 1, s);
    rc = ensure_lu(h, h->csr.p, s);
    if (rc) return rc;
  }
  if (max_iterations > 0) {
    if (!h->g_nrs) {
      CK(cudaStreamSynchronize(s));   // graph construction captures on the handle's own stream (one-time)
      int rc = build_nrs_graph(h);
      if (rc) return rc;
    }
    CK(cudaGraphLaunch(h->g_nrs, s));
  }
  // final build at X_final for Q, I_vsource, max_res_contrib (solver_factories.py:689-713)
  BuildPtrs bf = b;
  bf.csr = nullptr; bf.f = h->f.p; bf.Q = Q_out ? Q_out : h->Qw.p; bf.Ivs = Ivs_out ? Ivs_out : h->Ivs.p;
  bf.maxres = maxres_out ? maxres_out : h->maxres_w.p;
  bf.limit_out = h->limit_out.p;
  int rc = launch_eval(h, bf, s);
  if (rc) return rc;
  launch_residual(h, bf, 0, s);
  if (dQdt_out) {
    k_dqdt<<<grid_for(h, nu), 256, 0, s>>>(nu, h->ctl.p, bf.Q, L.Q_prev, L.dQdt_prev, L.Q_prev2, dQdt_out);
    h->launches++;
  }
  if (limit_out && h->total_lim > 0) {
    k_copy<<<grid_for(h, h->total_lim), 256, 0, s>>>(h->total_lim, h->limit_w.p, limit_out);
    h->launches++;
  }
  k_copy<<<grid_for(h, nx), 256, 0, s>>>(nx, h->nrs_X.p, X_out);
  k_nr_scalars<<<1, 1, 0, s>>>(h->st.p, max_iterations, scalars4_out);
  h->launches += 2;
  CK(cudaGetLastError());
  return VJX_OK;
}

// ---- transient --------------------------------------------------------------------------
static StepArgs make_step_args(vjx_handle* h) {
  StepArgs a;
  a.ts = h->ts.p; a.st = h->st.p; a.ctl = h->ctl.p; a.o = h->o;
  a.n_total = h->n_total; a.n_unknowns = h->n_unknowns; a.n_vs = h->n_vs; a.n_is = h->n_is; a.n_ext = h->n_ext;
  a.n_out = h->n_out > 0 ? h->n_out : h->n_ext; a.save_idx = h->n_out > 0 ? h->save_idx.p : nullptr;
  a.n_lim = h->total_lim; a.H = h->H;
  a.vsrc_desc = h->vsrc_desc.p; a.isrc_desc = h->isrc_desc.p; a.pwl_t = h->pwl_t.p; a.pwl_v = h->pwl_v.p;
  a.vsrc = h->vsrc_val.p; a.isrc = h->isrc_val.p;
  a.X = h->X.p; a.Xw = h->Xw.p; a.Vpred = h->Vpred.p; a.Vhist = h->Vhist.p; a.Vmax = h->Vmax.p;
  a.Q_prev = h->Q_prev.p; a.dQdt_prev = h->dQdt_prev.p; a.Q_prev2 = h->Q_prev2.p; a.Qw = h->Qw.p; a.dQdtw = h->dQdtw.p;
  a.Ivs = h->Ivs.p; a.maxres_w = h->maxres_w.p; a.hist_mrc = h->hist_mrc.p; a.res_tol_floor = h->res_tol_floor.p;
  a.limit_state = h->limit_state.p; a.limit_w = h->limit_w.p;
  a.times_out = h->times_out.p; a.V_out = h->V_out.p; a.I_out = h->I_out.p; a.records = h->records.p;
  a.max_records = h->max_records;
  return a;
}

__global__ void k_gather(int n, const int* idx, const double* src, double* dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void k_abs_copy(int n, const double* src, double* dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = fabs(src[i]);
}
__global__ void k_dec_attempts(TranScalars* s) { s->attempts -= 1; }
__global__ void k_set_attempts_limit(TranScalars* s, int limit) { s->attempts_limit = limit; }

__global__ void k_tran_init(TranScalars* s, NRStat* st, double t_stop, double dt, double max_dt, int max_steps, int H) {
  memset(s, 0, sizeof(TranScalars));
  s->t = 0.0; s->dt = dt; s->t_stop = t_stop; s->max_dt = max_dt; s->min_dt_used = dt; s->max_dt_used = dt;
  for (int k = 0; k < 4; ++k) s->dt_hist[k] = dt;
  s->hist_count = 1; s->step_idx = 1; s->warmup = 0; s->max_steps = max_steps; s->hist_head = 0; s->cont = 1;
  st->error = 0;
  (void)H;
}

extern "C" int vjx_tran_begin(vjx_handle* h, const double* X0, const double* Q_init, const double* I0,
                              const double* hist_maxres, double t_stop, double dt, double max_dt, int max_steps,
                              vjx_stream stream) {
  if (!h || !X0) return fail(VJX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  const int nt = h->n_total, nu = std::max(h->n_unknowns, 1), nx = h->n_total + h->n_vs, nvs1 = std::max(h->n_vs, 1);
  h->H = h->o.max_order + 2;
  if (h->H > 4 || h->H < 3) return fail(VJX_ERR_INVALID, "max_order must be 1 or 2");
  const int n_out = h->n_out > 0 ? h->n_out : h->n_ext;
  if (max_steps != h->max_steps || !h->X.p || h->V_out.n != (size_t)max_steps * n_out) {
    h->max_steps = max_steps;
    h->max_records = 4 * max_steps + 64;
    CK(h->X.alloc(nx)); CK(h->Vpred.alloc(nt)); CK(h->Vhist.alloc((size_t)h->H * nt)); CK(h->Vmax.alloc(nt));
    CK(h->Q_prev.alloc(nu)); CK(h->dQdt_prev.alloc(nu)); CK(h->Q_prev2.alloc(nu)); CK(h->hist_mrc.alloc(nu));
    CK(h->res_tol_floor.alloc(nu)); CK(h->limit_state.alloc(std::max(h->total_lim, 1)));
    CK(h->times_out.alloc(max_steps)); CK(h->V_out.alloc((size_t)max_steps * n_out));
    CK(h->I_out.alloc((size_t)max_steps * nvs1)); CK(h->records.alloc((size_t)h->max_records * 8));
    if (h->g_tran) { cudaGraphExecDestroy(h->g_tran); h->g_tran = nullptr; }
  }
  CK(cudaMemsetAsync(h->Vhist.p, 0, sizeof(double) * h->H * nt, s));
  CK(cudaMemsetAsync(h->dQdt_prev.p, 0, sizeof(double) * nu, s));
  CK(cudaMemsetAsync(h->Q_prev2.p, 0, sizeof(double) * nu, s));
  CK(cudaMemsetAsync(h->limit_state.p, 0, sizeof(double) * std::max(h->total_lim, 1), s));
  CK(cudaMemsetAsync(h->times_out.p, 0, sizeof(double) * max_steps, s));
  CK(cudaMemsetAsync(h->V_out.p, 0, sizeof(double) * (size_t)max_steps * n_out, s));
  CK(cudaMemsetAsync(h->I_out.p, 0, sizeof(double) * (size_t)max_steps * nvs1, s));
  CK(cudaMemcpyAsync(h->X.p, X0, sizeof(double) * nx, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(h->Vhist.p, X0, sizeof(double) * nt, cudaMemcpyDeviceToDevice, s));
  if (h->n_out > 0) k_gather<<<grid_for(h, n_out), 256, 0, s>>>(n_out, h->save_idx.p, X0, h->V_out.p);
  else CK(cudaMemcpyAsync(h->V_out.p, X0, sizeof(double) * h->n_ext, cudaMemcpyDeviceToDevice, s));
  k_copy<<<grid_for(h, nu), 256, 0, s>>>(h->n_unknowns, Q_init, h->Q_prev.p);
  k_copy<<<grid_for(h, nu), 256, 0, s>>>(h->n_unknowns, hist_maxres, h->hist_mrc.p);
  if (h->n_vs > 0 && I0) CK(cudaMemcpyAsync(h->I_out.p, I0, sizeof(double) * h->n_vs, cudaMemcpyDeviceToDevice, s));
  h->last_run_ms = 0.0;
  h->total_run_ms = 0.0;
  k_tran_init<<<1, 1, 0, s>>>(h->ts.p, h->st.p, t_stop, dt, max_dt, max_steps, h->H);
  h->launches += 3;
  k_abs_copy<<<grid_for(h, nt), 256, 0, s>>>(nt, X0, h->Vmax.p);
  h->launches++;
  CK(cudaGetLastError());
  return VJX_OK;
}


static void tran_build_ptrs(vjx_handle* h, BuildPtrs& b) {
  fill_build_ptrs_nr(h, b, h->Xw.p, h->vsrc_val.p, h->isrc_val.p, h->Q_prev.p, h->dQdt_prev.p, h->Q_prev2.p,
                     h->res_tol_floor.p);
}

static int launch_step_pre(vjx_handle* h, cudaGraphConditionalHandle nr_handle, int use_handle, cudaStream_t s) {
  StepArgs a = make_step_args(h);
  k_pre_step<<<1, 256, 0, s>>>(a, nr_handle, use_handle);
  k_predict<<<grid_for(h, h->n_total + h->n_vs + h->total_lim), 256, 0, s>>>(a);
  h->launches += 2;
  return VJX_OK;
}

static int launch_step_post(vjx_handle* h, cudaGraphConditionalHandle handle, int use_handle, int attempts_limit,
                            cudaStream_t s) {
  BuildPtrs b{};
  tran_build_ptrs(h, b);
  BuildPtrs bf = b;
  bf.csr = nullptr; bf.f = h->f.p; bf.Q = h->Qw.p; bf.Ivs = h->Ivs.p; bf.maxres = h->maxres_w.p;
  int rc = launch_eval(h, bf, s);
  if (rc) return rc;
  launch_residual(h, bf, 0, s);
  k_dqdt<<<grid_for(h, h->n_unknowns), 256, 0, s>>>(h->n_unknowns, h->ctl.p, h->Qw.p, h->Q_prev.p, h->dQdt_prev.p,
                                                   h->Q_prev2.p, h->dQdtw.p);
  StepArgs a = make_step_args(h);
  k_lte<<<grid_for(h, h->n_total), 256, 0, s>>>(a);
  k_step_decide<<<1, 1, 0, s>>>(a, handle, 0, attempts_limit);
  k_step_commit<<<grid_for(h, h->n_total + h->n_vs + h->total_lim), 256, 0, s>>>(a, handle, 0, attempts_limit);
  k_step_finish<<<1, 1, 0, s>>>(a, handle, use_handle, attempts_limit);
  h->launches += 5;
  return VJX_OK;
}

// Build the device-resident loop: WHILE(step){ pre; WHILE(nr){ iteration }; post } as one graph.
static int build_tran_graph(vjx_handle* h, int attempts_limit, cudaStream_t s) {
  if (h->g_tran) { cudaGraphExecDestroy(h->g_tran); h->g_tran = nullptr; }
  cudaGraph_t g;
  CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h_step, h_nr;
  CK(cudaGraphConditionalHandleCreate(&h_step, g, 1, cudaGraphCondAssignDefault));
  CK(cudaGraphConditionalHandleCreate(&h_nr, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams cp = {};
  cp.type = cudaGraphNodeTypeConditional;
  cp.conditional.handle = h_step;
  cp.conditional.type = cudaGraphCondTypeWhile;
  cp.conditional.size = 1;
  cudaGraphNode_t step_node;
  CK(cudaGraphAddNode(&step_node, g, nullptr, 0, &cp));
  cudaGraph_t step_body = cp.conditional.phGraph_out[0];
  // capture pre-step into step_body
  cudaStream_t cs = h->own_stream;
  long long launches_before = h->launches;
  CK(cudaStreamBeginCaptureToGraph(cs, step_body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  launch_step_pre(h, h_nr, 1, cs);
  cudaGraph_t tmp;
  CK(cudaStreamEndCapture(cs, &tmp));
  // find leaf nodes of step_body to chain the NR while node after them
  size_t nn = 0;
  CK(cudaGraphGetNodes(step_body, nullptr, &nn));
  std::vector<cudaGraphNode_t> nodes(nn);
  CK(cudaGraphGetNodes(step_body, nodes.data(), &nn));
  // the capture is linear: the last node added is the leaf
  std::vector<cudaGraphNode_t> leaves;
  for (auto nd : nodes) {
    size_t nd_dep = 0;
    CK(cudaGraphNodeGetDependentNodes(nd, nullptr, &nd_dep));
    if (nd_dep == 0) leaves.push_back(nd);
  }
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = h_nr;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t nr_node;
  CK(cudaGraphAddNode(&nr_node, step_body, leaves.data(), leaves.size(), &np));
  cudaGraph_t nr_body = np.conditional.phGraph_out[0];
  BuildPtrs b{};
  tran_build_ptrs(h, b);
  CK(cudaStreamBeginCaptureToGraph(cs, nr_body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  int rc = launch_nr_iteration(h, b, h->Xw.p, h_nr, 1, cs);
  CK(cudaStreamEndCapture(cs, &tmp));
  if (rc) return rc;
  // post-step after the NR while node
  CK(cudaStreamBeginCaptureToGraph(cs, step_body, &nr_node, nullptr, 1, cudaStreamCaptureModeRelaxed));
  rc = launch_step_post(h, h_step, 1, attempts_limit, cs);
  CK(cudaStreamEndCapture(cs, &tmp));
  if (rc) return rc;
  h->launches = launches_before;
  CK(cudaGraphInstantiate(&h->g_tran, g, 0));
  CK(cudaGraphDestroy(g));
  h->g_tran_limit = attempts_limit;
  (void)s;
  return VJX_OK;
}


// N4: device-side save list.  The reference records every external node (full_mna.py:1017-1023,
// 2122) and filters afterwards; for mul64 that is 133 k columns per accepted step where the netlist's `save`
// lines ask for 128 (vajax/benchmarks/data/mul64/runme.sim:296-308).  With a save list V_out has one column per
// listed node (index into X: 0 = ground, 1..n_total-1), written by k_step_commit.  n_save = 0 restores "all
// external nodes".  Takes effect at the next vjx_tran_begin.
extern "C" int vjx_tran_set_save_list(vjx_handle* h, const int* node_idx_host, int n_save) {
  if (!h || n_save < 0 || (n_save > 0 && !node_idx_host)) return fail(VJX_ERR_INVALID, "bad save list");
  if (n_save > h->n_total) return fail(VJX_ERR_INVALID, "save list longer than the node count");
  CK(cudaSetDevice(h->device));
  for (int k = 0; k < n_save; ++k)
    if (node_idx_host[k] < 0 || node_idx_host[k] >= h->n_total) return fail(VJX_ERR_INVALID, "save list: node index out of range");
  if (n_save > 0) {
    std::vector<int> v(node_idx_host, node_idx_host + n_save);
    CK(h->save_idx.upload(v));
  }
  h->n_out = n_save;
  if (h->g_tran) { cudaGraphExecDestroy(h->g_tran); h->g_tran = nullptr; }   // kernel arguments change
  h->max_steps = -1;                                                         // force re-allocation of V_out
  return VJX_OK;
}

// Sweep tuning: the persistent LU kernels (tickets + polling) of a handle take 1 / div of the GPU's CTA slots, so that
// several handles in flight on one GPU share the SMs.  Any grid size is safe (tickets); results do not depend on it.
// Must be called before the first factorisation / analysis of the handle.
extern "C" int vjx_lu_set_grid_div(vjx_handle* h, int div) {
  if (!h || div < 1) return fail(VJX_ERR_INVALID, "bad grid divisor");
  if (h->lu_ready) return fail(VJX_ERR_INVALID, "vjx_lu_set_grid_div: the LU of this handle is already analysed");
  h->lu_grid_div = div;
  return VJX_OK;
}

// A12: the reference's dense linear-solve path adds reg * I (1e-14) to the Jacobian inside every NR solve
// (solver_factories.py:784-788).  Needs the branch rows' diagonal entries in the CSR pattern handed to vjx_create.
// Call before the first solve of the handle (captured graphs hold the value).
extern "C" int vjx_set_dense_regularization(vjx_handle* h, double reg) {
  if (!h || !(reg >= 0.0)) return fail(VJX_ERR_INVALID, "bad dense regularisation");
  if (reg != 0.0)
    for (int i = h->n_unknowns; i < h->n_aug; ++i)
      if (csr_find(h->h_indptr, h->h_indices, i, i) < 0)
        return fail(VJX_ERR_INVALID, "vjx_set_dense_regularization: a branch row has no diagonal entry in the CSR pattern");
  h->dense_reg = reg;
  return VJX_OK;
}

extern "C" int vjx_tran_run(vjx_handle* h, int max_attempts, int* attempts, vjx_stream stream) {
  if (!h || !h->X.p) return fail(VJX_ERR_INVALID, "vjx_tran_begin not called");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  TranScalars* hts = (TranScalars*)((char*)h->pinned + 512);
  CK(cudaMemcpyAsync(hts, h->ts.p, sizeof(TranScalars), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int start_attempts = hts->attempts;
  const int limit = start_attempts + (max_attempts > 0 ? max_attempts : 0x3fffffff - start_attempts);
  if (!(hts->t < hts->t_stop) || hts->step_idx >= hts->max_steps) {
    if (attempts) *attempts = 0;
    return VJX_OK;
  }
  BuildPtrs b{};
  tran_build_ptrs(h, b);
  if (!h->lu_ready) {
    // analyse on the first system of the run (coefficients of the first step)
    launch_step_pre(h, 0, 0, s);
    k_iter_begin<<<1, 1, 0, s>>>(h->st.p);
    int rc = launch_eval(h, b, s);
    if (rc) return rc;
    launch_stamp(h, h->csr.p, 1, s);
    rc = ensure_lu(h, h->csr.p, s);
    if (rc) return rc;
    // undo the attempt counter bump of the probe
    k_dec_attempts<<<1, 1, 0, s>>>(h->ts.p);
  }
  k_set_attempts_limit<<<1, 1, 0, s>>>(h->ts.p, limit);
  if (!h->ev0) {
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
  }
  CK(cudaEventRecord(h->ev0, s));
  if (h->o.use_device_graph) {
    if (!h->g_tran) {
      int rc = build_tran_graph(h, limit, s);
      if (rc) return rc;
    }
    CK(cudaGraphLaunch(h->g_tran, s));
    CK(cudaStreamSynchronize(s));
  } else {
    NRStat* hs = (NRStat*)h->pinned;
    while (true) {
      launch_step_pre(h, 0, 0, s);
      while (true) {
        int rc = launch_nr_iteration(h, b, h->Xw.p, 0, 0, s);
        if (rc) return rc;
        CK(cudaMemcpyAsync(hs, h->st.p, sizeof(NRStat), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (hs->error & (1 | 4 | 8)) return fail(VJX_ERR_HANG, "dependency wait watchdog fired in the LU kernels");
        if (!hs->cont) break;
      }
      int rc = launch_step_post(h, 0, 0, limit, s);
      if (rc) return rc;
      CK(cudaMemcpyAsync(hts, h->ts.p, sizeof(TranScalars), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (!hts->cont) break;
    }
  }
  CK(cudaEventRecord(h->ev1, s));
  CK(cudaMemcpyAsync(hts, h->ts.p, sizeof(TranScalars), cudaMemcpyDeviceToHost, s));
  int rc = check_device_error(h, s);
  if (rc) return rc;
  {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_run_ms = ms;
    h->total_run_ms += ms;
  }
  if (attempts) *attempts = hts->attempts - start_attempts;
  CK(cudaGetLastError());
  return VJX_OK;
}


__global__ void k_set_step_limit(TranScalars* s, int limit) {
  s->max_steps = limit;
  s->cont = (s->t < s->t_stop) && (s->step_idx < limit);
}

// Lower / raise the accepted-step bound of the running transient (<= the rows allocated by
// vjx_tran_begin): "run <= K steps, return state", the contract of the reference's
// _run_with_checkpoints (full_mna.py:1443-1699).
extern "C" int vjx_tran_set_step_limit(vjx_handle* h, int step_limit, vjx_stream stream) {
  if (!h || !h->X.p) return fail(VJX_ERR_INVALID, "vjx_tran_begin not called");
  if (step_limit < 1 || step_limit > h->max_steps) return fail(VJX_ERR_INVALID, "step limit outside the allocated rows");
  CK(cudaSetDevice(h->device));
  k_set_step_limit<<<1, 1, 0, (cudaStream_t)stream>>>(h->ts.p, step_limit);
  h->launches++;
  CK(cudaGetLastError());
  return VJX_OK;
}

extern "C" int vjx_tran_scalars(vjx_handle* h, double* out, vjx_stream stream) {
  if (!h || !h->ts.p) return fail(VJX_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  TranScalars* hts = (TranScalars*)((char*)h->pinned + 512);
  NRStat* hs = (NRStat*)h->pinned;
  CK(cudaMemcpyAsync(hts, h->ts.p, sizeof(TranScalars), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hs, h->st.p, sizeof(NRStat), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  out[0] = hts->t; out[1] = hts->dt; out[2] = hts->step_idx; out[3] = hts->total_nr; out[4] = hts->rejected;
  out[5] = hts->min_dt_used; out[6] = hts->max_dt_used; out[7] = hts->hist_count; out[8] = hts->warmup;
  out[9] = hts->consec; out[10] = hts->attempts; out[11] = hs->error;
  return VJX_OK;
}

extern "C" int vjx_tran_read(vjx_handle* h, int row0, int rows, double* times, double* V_out, double* I_out, vjx_stream stream) {
  if (!h || !h->X.p) return fail(VJX_ERR_INVALID, "no transient state");
  if (row0 < 0 || rows < 0 || row0 + rows > h->max_steps) return fail(VJX_ERR_INVALID, "row range");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  const int nvs1 = std::max(h->n_vs, 1);
  if (times) CK(cudaMemcpyAsync(times, h->times_out.p + row0, sizeof(double) * rows, cudaMemcpyDeviceToHost, s));
  const int n_out = h->n_out > 0 ? h->n_out : h->n_ext;
  if (V_out) CK(cudaMemcpyAsync(V_out, h->V_out.p + (size_t)row0 * n_out, sizeof(double) * (size_t)rows * n_out, cudaMemcpyDeviceToHost, s));
  if (I_out) CK(cudaMemcpyAsync(I_out, h->I_out.p + (size_t)row0 * nvs1, sizeof(double) * (size_t)rows * nvs1, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VJX_OK;
}

extern "C" int vjx_tran_records(vjx_handle* h, int first, int count, double* out, vjx_stream stream) {
  if (!h || !h->records.p) return fail(VJX_ERR_INVALID, "no transient state");
  if (first < 0 || count < 0 || first + count > h->max_records) return fail(VJX_ERR_INVALID, "record range");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(out, h->records.p + (size_t)first * 8, sizeof(double) * 8 * count, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return VJX_OK;
}

extern "C" int vjx_tran_state(vjx_handle* h, double* X, double* Q_prev, double* dQdt_prev, vjx_stream stream) {
  if (!h || !h->X.p) return fail(VJX_ERR_INVALID, "no transient state");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  if (X) CK(cudaMemcpyAsync(X, h->X.p, sizeof(double) * (h->n_total + h->n_vs), cudaMemcpyDeviceToHost, s));
  if (Q_prev) CK(cudaMemcpyAsync(Q_prev, h->Q_prev.p, sizeof(double) * h->n_unknowns, cudaMemcpyDeviceToHost, s));
  if (dQdt_prev) CK(cudaMemcpyAsync(dQdt_prev, h->dQdt_prev.p, sizeof(double) * h->n_unknowns, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VJX_OK;
}

// Device time (CUDA events on the launch stream, tight around the loop launch) of the last
// vjx_tran_run call and the running total since vjx_tran_begin.
extern "C" int vjx_tran_run_ms(vjx_handle* h, double* last_ms, double* total_ms) {
  if (!h) return fail(VJX_ERR_INVALID, "null handle");
  if (last_ms) *last_ms = h->last_run_ms;
  if (total_ms) *total_ms = h->total_run_ms;
  return VJX_OK;
}

__global__ void k_stamp_now(ull* out) { *out = global_ns(); }

// Debug probe: run one LU phase (3 = factor, 4 = forward solve, 5 = back solve) on the handle's
// current state and return, per permuted row, the completion time in ns since the kernel started,
// together with the row's dependency level.  Used by scripts/lu_timeline.py to see where a
// dependency chain spends its time; never called by the product path.
extern "C" int vjx_lu_timeline(vjx_handle* h, int phase, double* ns_out, int* level_out, vjx_stream stream) {
  if (!h || !h->X.p || !h->lu_ready) return fail(VJX_ERR_INVALID, "run a transient first");
  if (phase != 3 && phase != 5) return fail(VJX_ERR_INVALID, "phase must be 3 (factor + forward solve) or 5 (back solve)");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  const int n = h->sym.n;
  DevBuf<ull> ts;
  CK(ts.alloc((size_t)n + 1 + 4100));
  LUDev lud = h->lud;
  k_iter_begin<<<1, 1, 0, s>>>(h->st.p);
  if (phase == 5) launch_lu_factor(h, h->lud, h->csr.p, h->f.p, -1.0, h->noi_row_mask.p, h->delta.p, s);
  lud.ts = ts.p;
  CK(cudaMemsetAsync(ts.p, 0, sizeof(ull) * ((size_t)n + 8 + VJX_DB_MAX), s));
  k_stamp_now<<<1, 1, 0, s>>>(ts.p + n);
  if (phase == 3) launch_lu_factor(h, lud, h->csr.p, h->f.p, -1.0, h->noi_row_mask.p, h->delta.p, s);
  if (phase == 5) k_lu_usolve<<<h->lu_grid_u, VJX_SOLVE_THREADS, 0, s>>>(lud, h->delta.p, h->st.p);
  std::vector<ull> hts((size_t)n + 8 + VJX_DB_MAX);
  CK(cudaStreamSynchronize(s));
  CK(cudaMemcpy(hts.data(), ts.p, sizeof(ull) * ((size_t)n + 8 + VJX_DB_MAX), cudaMemcpyDeviceToHost));
  if (phase == 3 && h->lud.n_dense > 0 && getenv("VJX_DT_TRACE")) {
    auto us = [&](int q) { return ((long long)hts[n + q] - (long long)hts[n]) / 1e3; };
    fprintf(stderr, "dense trailing block (%d rows), us since factorisation start: begin %.1f, loaded %.1f, eliminated %.1f, "
            "spilled %.1f, solved %.1f\n", h->lud.n_dense, us(1), us(2), us(3), us(4), us(5));
    fprintf(stderr, "row updates of the dense-block rows done at (us): row(offset) Lbulk Ldense -> t\n");
    for (int t = 0; t < h->lud.n_dense; t += 6) {
      const int i = n - h->lud.n_dense + t;
      int lb = 0;
      for (int p = h->sym.lu_ptr[i]; p < h->sym.lu_diag[i]; ++p) lb += h->sym.lu_col[p] < n - h->lud.n_dense;
      fprintf(stderr, "  %3d %3d %3d -> %.1f\n", t, lb, h->sym.lu_diag[i] - h->sym.lu_ptr[i] - lb, us(8 + t));
    }
  }
  for (int i = 0; i < n; ++i) {
    ns_out[i] = (double)((long long)hts[i] - (long long)hts[n]);
    if (level_out) level_out[i] = phase == 5 ? h->sym.level_u[i] : h->sym.level_l[i];
  }
  return check_device_error(h, s);
}

extern "C" long long vjx_launch_count(vjx_handle* h) { return h ? h->launches : 0; }

// Phase timing probe on the transient working state (Xw etc.), CUDA events on the launch stream.
extern "C" int vjx_time_phase(vjx_handle* h, int phase, int reps, double* ms_avg, int* launches, vjx_stream stream) {
  if (!h || !h->X.p || !h->lu_ready) return fail(VJX_ERR_INVALID, "run a transient first");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, stream);
  BuildPtrs b{};
  tran_build_ptrs(h, b);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int nl = 0;
  UpdateArgs ua = make_update_args(h, h->Xw.p, b.limit_in, b.limit_out);
  float total = 0.f;
  for (int r = 0; r < reps + 1; ++r) {
    long long l0 = h->launches;
    if (phase == 3 || phase == 4 || phase == 5) {
      // each LU phase needs fresh tickets and the preceding phases' results
      k_iter_begin<<<1, 1, 0, s>>>(h->st.p);
      if (phase == 5) launch_lu_factor(h, h->lud, h->csr.p, h->f.p, -1.0, h->noi_row_mask.p, h->delta.p, s);
    }
    CK(cudaEventRecord(e0, s));
    l0 = h->launches;
    switch (phase) {
      case 0: { int rc = launch_eval(h, b, s); if (rc) return rc; break; }
      case 1: launch_stamp(h, h->csr.p, 1, s); break;
      case 2: launch_residual(h, b, 0, s); break;
      case 3: launch_lu_factor(h, h->lud, h->csr.p, h->f.p, -1.0, h->noi_row_mask.p, h->delta.p, s); break;
      case 4: break;   // the forward solve is part of the factorisation kernels (phase 3)
      case 5: k_lu_usolve<<<h->lu_grid_u, VJX_SOLVE_THREADS, 0, s>>>(h->lud, h->delta.p, h->st.p); h->launches++; break;
      case 6: k_delta_reduce<<<grid_for(h, std::max(h->n_aug, h->total_lim)), 256, 0, s>>>(ua); h->launches++; break;
      default: return fail(VJX_ERR_INVALID, "unknown phase");
    }
    nl = (int)(h->launches - l0);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_avg) *ms_avg = total / std::max(reps, 1);
  if (launches) *launches = nl;
  return check_device_error(h, s);
}
