// Host-side symbolic analysis for the GPU sparse LU.
//
// The reference delegates this to cuDSS (via spineax, vajax/analysis/solver_factories.py:882-930)
// or UMFPACK (vajax/sparse/umfpack_ffi.cpp:253-300: symbolic cached by (n, nnz), numeric + solve per
// call, pivot tolerance 0.1).  Here the analysis runs once per circuit on the first Jacobian
// ("pivot order fixed at the DC operating point", BASELINE.json north_star item 3):
//   0. singleton peeling (UMFPACK / KLU-BTF first phase): rail nodes pinned by voltage sources,
//      branch-current columns and identity rows leave the core without fill,
//   1. approximate-minimum-degree ordering of the core's pattern of A + A^T (dense rows deferred),
//   2. left-looking Gilbert-Peierls LU with threshold partial pivoting that prefers the diagonal
//      (the KLU strategy) to fix the row permutation and the structural pattern of L and U,
//   3. row-wise (CSR) storage of L+U in the permuted space, the map from CSR slots of A into it,
//   4. dependency levels for the factorisation / forward solve (L pattern) and back solve (U),
//   5. the scatter program (position of every U(k,j) inside row i, per L entry) of the numeric kernel.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct VjxSymbolic {
  int n = 0;
  std::vector<int> perm_r;     // permuted row i  <- original row perm_r[i]
  std::vector<int> perm_c;     // permuted col j  <- original col perm_c[j]
  std::vector<int> lu_ptr;     // [n+1] row pointers of L+U (permuted space)
  std::vector<int> lu_col;     // column indices, ascending per row
  std::vector<int> lu_diag;    // [n] position of the diagonal entry of row i
  std::vector<int> lu_src;     // per L+U entry: slot of A's CSR values, or -1 for fill
  std::vector<int> order_f;    // all rows sorted by L-dependency level (factorisation order)
  std::vector<int> order_l;    // bulk rows sorted by L-dependency level (forward solve)
  std::vector<int> order_u;    // bulk rows sorted by U-dependency level (back solve)
  std::vector<int> level_l, level_u;
  int n_levels_l = 0, n_levels_u = 0, max_row_len = 0;
  // Scatter program of the numeric refactorisation (up-looking, k-sequential): row i is held in a
  // working vector w; for every L entry (i,k) in ascending k,  l = w[k] / U(k,k)  and
  // w[j] -= l * U(k,j) for every j in U(k,:).  upd_pos lists, for each L entry, the position inside
  // row i (relative to the row start) of each column of U(k,:), in the order U(k,:) is stored, so the
  // kernel never searches.  upd_ptr is indexed by the absolute position of the L entry.
  std::vector<int64_t> upd_ptr;   // [nnz_lu+1]
  std::vector<int> upd_pos;       // [sum over L entries (i,k) of |U(k,:)|]
  int max_l_len = 0;              // longest L part of a row (sequential k-steps of one row)
  // Dependency-chain tail.  The last rows of the elimination order form the near-dense trailing
  // block of the factor: a chain of hundreds of sequentially dependent rows with a handful of rows
  // per dependency level.  Rows [tail_start, n) are solved by ONE thread block that hands values
  // over in shared memory (~0.2 us per link) instead of global flags (~3 us per link): in index
  // order for L (after the bulk), in reverse index order for U (before the bulk).  order_l / order_u
  // list the bulk rows [0, tail_start) only, sorted by dependency level.
  int tail_start = 0;
  // Rows whose U part is longer than kLongRow (a rail row kept as a column-singleton pivot has one
  // entry per attached device) are left out of order_u and solved by a whole thread block each.
  std::vector<int> long_u;
  // Dense trailing block: rows [n - n_dense, n) hand their part right of column n - n_dense to a dense
  // kernel.  The L part of every row is stored in EXECUTION order, not column order: sorted by an
  // estimate of when row k is finished (its dependency depth), columns of the dense block last.  Any
  // order that applies step k' before step k whenever U(k',k) != 0 is valid; taking the steps whose
  // rows finish early first keeps a row from idling at one late dependency with hundreds of ready
  // steps queued behind it.
  int n_dense = 0;
  std::vector<int> depth;         // the dependency depth used for that order (see vjx_symbolic.cpp)
  int n_singletons = 0;           // pivots peeled as row / column singletons (no fill, no updates)
  int n_core = 0;                 // unknowns left for AMD + Gilbert-Peierls
  double flops = 0.0;             // 2 * sum_k nnz(L(:,k)) * nnz(U(k,:)) + nnz(L)
  int64_t nnz_lu = 0;
  // Heavy rows (vjx_build_heavy): rows with many k-steps are not walked k-step by k-step by one warp but depth by
  // depth by a whole thread block.  The k-steps of a row that have the same dependency depth are independent of
  // each other (U(k1,k2) = 0), so their multipliers l = w[k] / U(k,k) are formed together (one lane each, at most
  // 32 per group) and their updates are regrouped BY TARGET POSITION: a lane owns a position of the row and applies
  // the contributions -l_s * U(k_s, j) to it in ascending k - no two lanes touch the same word, no atomics, the
  // order of every sum is fixed at analysis time.  Program of one row = a sequence of group blocks (all ints):
  //   [nrec, keep, nk, ints]  kpos[32]  k[32]  info[nrec]  then nrec records of  pos[32]  ent[R][32]
  // kpos / pos: position inside the row (-1: idle lane); ent = slot | (index into the LU values << 5), -1: none;
  // info = offset of the record from the start of its group | R << 28, R <= 4 the largest number of contributions
  // a lane of the record has (positions are sorted by that count, so a record's lanes have similar counts).  A
  // position with more than 4 contributions in its group continues in a `keep` block (same multipliers, no new
  // k-steps) behind a barrier.
  struct Heavy {
    std::vector<int> rows;           // heavy rows (permuted index) in factorisation order
    std::vector<long long> off;      // [rows.size() + 1] start of each row's program in prog
    std::vector<int> prog;
    long long n_contrib = 0, n_slots = 0;
    int n_groups = 0, n_records = 0, n_ksteps = 0;
  } heavy;
};

// Builds S.heavy for the given rows (permuted indices, in factorisation order).  Needs the scatter program.
void vjx_build_heavy(VjxSymbolic& S, const std::vector<int>& rows);
// End of the part of row i's L entries that k_lu_rows executes itself (the rest belongs to the dense block).
int vjx_row_lend(const VjxSymbolic& S, int i);
// Rows with at least min_ksteps such k-steps and at most max_len entries, in factorisation order.
std::vector<int> vjx_select_heavy(const VjxSymbolic& S, int t0, int t1, int min_ksteps, int max_len);
#define VJX_HV_GRP_INTS 68

// Returns 0 on success; <0 and fills err otherwise.
// dense_cap: upper bound on the rows of the dense trailing block (0 = none).  > 0: out.n_dense is the chain of
// single-row dependency levels at the end of the order (whole panels of 16, within the longest suffix of rows
// that all have an L part); < 0: that longest suffix itself, up to |dense_cap|.
int vjx_symbolic_analyze(int n, const int* indptr, const int* indices, const double* values,
                         double pivot_tol, bool build_update_stream, VjxSymbolic& out, std::string& err,
                         int dense_cap = 0);

// Approximate minimum degree ordering on the pattern of A + A^T. perm[k] = k-th pivot.
void vjx_amd_order(int n, const int* indptr, const int* indices, std::vector<int>& perm);
// Exact minimum degree, ties broken by the dependency level the row would get (lowest first): same fill, a
// markedly lower dependency DAG (the default ordering of the core; VJX_ORDER=amd selects the one above).
void vjx_mdh_order(int n, const int* indptr, const int* indices, std::vector<int>& perm);

// Host emulation of the GPU numeric algorithm (diagnostics for CPU-only tests of the symbolic
// structures; never used by the product path): refactor + solve, returns max-norm residual.
double vjx_symbolic_selfcheck(const VjxSymbolic& s, int n, const int* indptr, const int* indices,
                              const double* values, const double* b, double* x_out);
