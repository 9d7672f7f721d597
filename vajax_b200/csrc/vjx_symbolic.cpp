// See vjx_symbolic.h.  Plain C++17, no CUDA.
#include "vjx_symbolic.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <set>

// ---------------------------------------------------------------------------------------------
// Approximate minimum degree on the quotient graph (Amestoy/Davis/Duff style bound, no
// supervariables).  Variables with degree above max(16, 10*sqrt(n)) are ordered last, as AMD
// does for dense rows: supply rails touch a large fraction of all devices.
// ---------------------------------------------------------------------------------------------
void vjx_amd_order(int n, const int* indptr, const int* indices, std::vector<int>& perm) {
  perm.clear();
  perm.reserve(n);
  if (n == 0) return;
  // symmetric adjacency without diagonal
  std::vector<std::vector<int>> adjv(n);
  {
    std::vector<int> cnt(n, 0);
    for (int i = 0; i < n; ++i)
      for (int p = indptr[i]; p < indptr[i + 1]; ++p) {
        int j = indices[p];
        if (j != i && j >= 0 && j < n) { cnt[i]++; cnt[j]++; }
      }
    for (int i = 0; i < n; ++i) adjv[i].reserve(cnt[i]);
    for (int i = 0; i < n; ++i)
      for (int p = indptr[i]; p < indptr[i + 1]; ++p) {
        int j = indices[p];
        if (j != i && j >= 0 && j < n) { adjv[i].push_back(j); adjv[j].push_back(i); }
      }
    for (int i = 0; i < n; ++i) {
      auto& a = adjv[i];
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
  }
  const int dense_thr = std::max(16, (int)(10.0 * std::sqrt((double)n)));
  // status: 0 alive variable, 1 element, 2 absorbed element, 3 dense (deferred)
  std::vector<char> status(n, 0);
  std::vector<int> dense;
  for (int i = 0; i < n; ++i)
    if ((int)adjv[i].size() > dense_thr) { status[i] = 3; dense.push_back(i); }
  if (!dense.empty())
    for (int i = 0; i < n; ++i) {
      if (status[i] == 3) continue;
      auto& a = adjv[i];
      a.erase(std::remove_if(a.begin(), a.end(), [&](int v) { return status[v] == 3; }), a.end());
    }
  std::vector<std::vector<int>> adje(n), elem(n);
  std::vector<int> deg(n, 0), head(n + 1, -1), nxt(n, -1), prv(n, -1);
  auto bucket_insert = [&](int v) {
    int d = deg[v];
    nxt[v] = head[d];
    prv[v] = -1;
    if (head[d] >= 0) prv[head[d]] = v;
    head[d] = v;
  };
  auto bucket_remove = [&](int v) {
    int d = deg[v];
    if (prv[v] >= 0) nxt[prv[v]] = nxt[v]; else head[d] = nxt[v];
    if (nxt[v] >= 0) prv[nxt[v]] = prv[v];
  };
  int n_nd = 0;
  for (int i = 0; i < n; ++i)
    if (status[i] == 0) { deg[i] = (int)adjv[i].size(); bucket_insert(i); n_nd++; }
  std::vector<int> tag(n, -1), wtag(n, -1), w(n, 0), Lp;
  int mindeg = 0;
  for (int k = 0; k < n_nd; ++k) {
    while (mindeg < n && head[mindeg] < 0) mindeg++;
    int p = head[mindeg];
    bucket_remove(p);
    perm.push_back(p);
    // pattern of the new element
    Lp.clear();
    tag[p] = k;
    for (int v : adjv[p])
      if (status[v] == 0 && tag[v] != k) { tag[v] = k; Lp.push_back(v); }
    for (int e : adje[p]) {
      if (status[e] != 1) continue;
      for (int v : elem[e])
        if (status[v] == 0 && tag[v] != k) { tag[v] = k; Lp.push_back(v); }
      status[e] = 2;
      std::vector<int>().swap(elem[e]);
    }
    status[p] = 1;
    std::vector<int>().swap(adjv[p]);
    std::vector<int>().swap(adje[p]);
    const int lp = (int)Lp.size();
    // w[e] = |L_e \ L_p| for elements adjacent to L_p
    for (int i : Lp)
      for (int e : adje[i]) {
        if (status[e] != 1) continue;
        if (wtag[e] != k) { wtag[e] = k; w[e] = (int)elem[e].size(); }
        w[e]--;
      }
    const int remaining = n_nd - k - 1;
    for (int i : Lp) {
      bucket_remove(i);
      auto& ae = adje[i];
      int m = 0;
      long long d = 0;
      for (int e : ae) {
        if (status[e] != 1) continue;
        if (w[e] <= 0) { continue; }  // e is a subset of L_p: covered by the new element
        ae[m++] = e;
        d += w[e];
      }
      ae.resize(m);
      ae.push_back(p);
      auto& av = adjv[i];
      int mv = 0;
      for (int v : av)
        if (status[v] == 0 && tag[v] != k) av[mv++] = v;
      av.resize(mv);
      d += mv + (lp - 1);
      long long dold = (long long)deg[i] + (lp - 1);
      d = std::min(d, dold);
      d = std::min<long long>(d, remaining > 0 ? remaining - 1 : 0);
      deg[i] = (int)std::max<long long>(d, 0);
      bucket_insert(i);
      if (deg[i] < mindeg) mindeg = deg[i];
    }
    elem[p] = Lp;
    // elements fully covered by L_p are dead now
    for (int i : Lp)
      for (int e : adje[i]) (void)e;
  }
  std::sort(dense.begin(), dense.end(), [&](int a, int b) {
    if (adjv[a].size() != adjv[b].size()) return adjv[a].size() < adjv[b].size();
    return a < b;
  });
  for (int v : dense) perm.push_back(v);
}

// ---------------------------------------------------------------------------------------------
// Minimum degree with dependency-depth tie-breaking (exact external degrees on explicit elimination
// graphs).  Among the vertices of minimum degree the one whose row would sit lowest in the dependency
// DAG of the factor is eliminated first: pend[v] = 1 + max level of the eliminated vertices whose clique
// holds v, which is exactly the level the numeric kernels would give row v.  Minimum degree alone does
// not care in which order it takes equal-degree vertices, and a first-in / last-out bucket (the AMD
// above) grows long single-row chains out of that freedom: on the c6288-shaped core the factor has the
// same fill (+1 %) but 105 dependency levels instead of 158 and a trailing chain of 48 single-row
// levels instead of 118 (scripts/ordering_study.py, profiles/r2_ordering_notes.md).
// Vertices above the dense threshold are ordered last, as in vjx_amd_order.
// ---------------------------------------------------------------------------------------------
void vjx_mdh_order(int n, const int* indptr, const int* indices, std::vector<int>& perm) {
  perm.clear();
  perm.reserve(n);
  if (n == 0) return;
  std::vector<std::vector<int>> adj(n);
  {
    std::vector<int> cnt(n, 0);
    for (int i = 0; i < n; ++i)
      for (int p = indptr[i]; p < indptr[i + 1]; ++p) {
        int j = indices[p];
        if (j != i && j >= 0 && j < n) { cnt[i]++; cnt[j]++; }
      }
    for (int i = 0; i < n; ++i) adj[i].reserve(cnt[i]);
    for (int i = 0; i < n; ++i)
      for (int p = indptr[i]; p < indptr[i + 1]; ++p) {
        int j = indices[p];
        if (j != i && j >= 0 && j < n) { adj[i].push_back(j); adj[j].push_back(i); }
      }
    for (int i = 0; i < n; ++i) {
      auto& a = adj[i];
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
  }
  const int dense_thr = std::max(16, (int)(10.0 * std::sqrt((double)n)));
  std::vector<char> state(n, 0);   // 0 alive, 1 eliminated, 3 dense (deferred)
  std::vector<int> dense;
  for (int i = 0; i < n; ++i)
    if ((int)adj[i].size() > dense_thr) { state[i] = 3; dense.push_back(i); }
  if (!dense.empty())
    for (int i = 0; i < n; ++i) {
      if (state[i] == 3) continue;
      auto& a = adj[i];
      a.erase(std::remove_if(a.begin(), a.end(), [&](int v) { return state[v] == 3; }), a.end());
    }
  std::vector<int> pend(n, 0);
  struct Key {
    int deg, pend, v;
    bool operator<(const Key& o) const {
      if (deg != o.deg) return deg < o.deg;
      if (pend != o.pend) return pend < o.pend;
      return v < o.v;
    }
  };
  std::set<Key> heap;
  for (int i = 0; i < n; ++i)
    if (state[i] == 0) heap.insert(Key{(int)adj[i].size(), 0, i});
  std::vector<int> merged;
  while (!heap.empty()) {
    const Key k = *heap.begin();
    heap.erase(heap.begin());
    const int p = k.v;
    state[p] = 1;
    perm.push_back(p);
    std::vector<int> clique;
    clique.swap(adj[p]);   // sorted, alive vertices only (eliminated ones are removed from every list below)
    for (int u : clique) {
      heap.erase(Key{(int)adj[u].size(), pend[u], u});
      // adj[u] <- (adj[u] U clique) \ {u, p}
      merged.clear();
      merged.reserve(adj[u].size() + clique.size());
      size_t a = 0, b = 0;
      const auto& au = adj[u];
      while (a < au.size() || b < clique.size()) {
        int x;
        if (b >= clique.size() || (a < au.size() && au[a] < clique[b])) x = au[a++];
        else if (a >= au.size() || clique[b] < au[a]) x = clique[b++];
        else { x = au[a]; ++a; ++b; }
        if (x != u && x != p) merged.push_back(x);
      }
      adj[u].swap(merged);
      pend[u] = std::max(pend[u], k.pend + 1);
      heap.insert(Key{(int)adj[u].size(), pend[u], u});
    }
  }
  std::sort(dense.begin(), dense.end(), [&](int a, int b) {
    if (adj[a].size() != adj[b].size()) return adj[a].size() < adj[b].size();
    return a < b;
  });
  for (int v : dense) perm.push_back(v);
}

// ---------------------------------------------------------------------------------------------
// Gilbert-Peierls left-looking LU with threshold partial pivoting (diagonal preference).
// ---------------------------------------------------------------------------------------------
namespace {
struct CSC {
  int n;
  std::vector<int> p, i;
  std::vector<double> x;
  std::vector<int> slot;  // original CSR slot of each entry
};

CSC csr_to_csc(int n, const int* indptr, const int* indices, const double* values) {
  CSC c;
  c.n = n;
  int nnz = indptr[n];
  c.p.assign(n + 1, 0);
  c.i.resize(nnz);
  c.x.resize(nnz);
  c.slot.resize(nnz);
  for (int k = 0; k < nnz; ++k) c.p[indices[k] + 1]++;
  for (int j = 0; j < n; ++j) c.p[j + 1] += c.p[j];
  std::vector<int> nx(c.p.begin(), c.p.end() - 1);
  for (int r = 0; r < n; ++r)
    for (int k = indptr[r]; k < indptr[r + 1]; ++k) {
      int q = nx[indices[k]]++;
      c.i[q] = r;
      c.x[q] = values ? values[k] : 1.0;
      c.slot[q] = k;
    }
  return c;
}

// Singleton peeling (the first phase of UMFPACK / KLU-BTF).  A column with one active entry is a
// pivot whose row becomes a (possibly long) U row that nothing depends on; a row with one active
// entry is a pivot whose column becomes L entries without updates.  Either way no fill is created,
// and the supply rails of an MNA matrix (a rail node pinned by a voltage source: branch row = row
// singleton, branch-current column = column singleton) leave the core that AMD + Gilbert-Peierls
// have to handle, together with every NOI row that was rewritten to identity.
struct Peel {
  std::vector<int> rows, cols;          // pivot sequence of the peeled part
  std::vector<char> row_active, col_active;
};

Peel peel_singletons(int n, const int* indptr, const int* indices, const double* values, const CSC& A) {
  Peel P;
  P.row_active.assign(n, 1);
  P.col_active.assign(n, 1);
  std::vector<int> rc(n), cc(n);
  for (int i = 0; i < n; ++i) rc[i] = indptr[i + 1] - indptr[i];
  for (int j = 0; j < n; ++j) cc[j] = A.p[j + 1] - A.p[j];
  std::vector<int> q;  // encoded: idx * 2 + (1 = column singleton, 0 = row singleton)
  q.reserve(2 * (size_t)n);
  for (int j = 0; j < n; ++j) if (cc[j] == 1) q.push_back(2 * j + 1);
  for (int i = 0; i < n; ++i) if (rc[i] == 1) q.push_back(2 * i);
  for (size_t h = 0; h < q.size(); ++h) {
    const int idx = q[h] >> 1;
    if (q[h] & 1) {
      const int j = idx;
      if (!P.col_active[j] || cc[j] != 1) continue;
      int i = -1;
      double v = 0.0;
      for (int pp = A.p[j]; pp < A.p[j + 1]; ++pp)
        if (P.row_active[A.i[pp]]) { i = A.i[pp]; v = A.x[pp]; break; }
      if (i < 0 || !(std::fabs(v) > 0.0)) continue;
      P.rows.push_back(i);
      P.cols.push_back(j);
      P.row_active[i] = 0;
      P.col_active[j] = 0;
      for (int kk = indptr[i]; kk < indptr[i + 1]; ++kk) {
        int j2 = indices[kk];
        if (P.col_active[j2] && --cc[j2] == 1) q.push_back(2 * j2 + 1);
      }
    } else {
      const int i = idx;
      if (!P.row_active[i] || rc[i] != 1) continue;
      int j = -1;
      double v = 0.0;
      for (int kk = indptr[i]; kk < indptr[i + 1]; ++kk)
        if (P.col_active[indices[kk]]) { j = indices[kk]; v = values ? values[kk] : 1.0; break; }
      if (j < 0 || !(std::fabs(v) > 0.0)) continue;
      P.rows.push_back(i);
      P.cols.push_back(j);
      P.row_active[i] = 0;
      P.col_active[j] = 0;
      for (int pp = A.p[j]; pp < A.p[j + 1]; ++pp) {
        int i2 = A.i[pp];
        if (P.row_active[i2] && --rc[i2] == 1) q.push_back(2 * i2);
      }
      // the row's other entries sit in already peeled columns; removing the row does not
      // change any active column count
    }
  }
  return P;
}
}  // namespace

int vjx_symbolic_analyze(int n, const int* indptr, const int* indices, const double* values,
                         double pivot_tol, bool build_update_stream, VjxSymbolic& S, std::string& err, int dense_cap) {
  S = VjxSymbolic();
  S.n = n;
  if (n <= 0) { err = "empty matrix"; return -1; }
  CSC A = csr_to_csc(n, indptr, indices, values);
  // --- pivot column order: peeled singletons first, then AMD on the remaining core
  std::vector<int> q;
  std::vector<int> pref_row(n, -1);   // preferred pivot row of each column
  Peel P = peel_singletons(n, indptr, indices, values, A);
  const int n_single = (int)P.cols.size();
  S.n_singletons = n_single;
  q.reserve(n);
  for (int k = 0; k < n_single; ++k) {
    q.push_back(P.cols[k]);
    pref_row[P.cols[k]] = P.rows[k];
  }
  {
    // pair the core's rows and columns: (i, i) when both are active, the rest in sorted order
    std::vector<int> core_cols, loc(n, -1), pair_col_of_row(n, -1), free_rows, free_cols;
    for (int j = 0; j < n; ++j) {
      if (!P.col_active[j]) continue;
      if (P.row_active[j]) pair_col_of_row[j] = j; else free_cols.push_back(j);
    }
    for (int i = 0; i < n; ++i)
      if (P.row_active[i] && pair_col_of_row[i] < 0) free_rows.push_back(i);
    if (free_rows.size() != free_cols.size()) { err = "singleton peeling left a non-square core"; return -1; }
    for (size_t t = 0; t < free_rows.size(); ++t) pair_col_of_row[free_rows[t]] = free_cols[t];
    for (int j = 0; j < n; ++j)
      if (P.col_active[j]) { loc[j] = (int)core_cols.size(); core_cols.push_back(j); }
    const int nc = (int)core_cols.size();
    S.n_core = nc;
    std::vector<int> row_of_loc(nc, -1);
    for (int i = 0; i < n; ++i)
      if (P.row_active[i]) { row_of_loc[loc[pair_col_of_row[i]]] = i; pref_row[pair_col_of_row[i]] = i; }
    std::vector<int> lip(nc + 1, 0), lix;
    for (int a = 0; a < nc; ++a) {
      int i = row_of_loc[a];
      for (int kk = indptr[i]; kk < indptr[i + 1]; ++kk)
        if (P.col_active[indices[kk]]) lix.push_back(loc[indices[kk]]);
      std::sort(lix.begin() + lip[a], lix.end());
      lip[a + 1] = (int)lix.size();
    }
    std::vector<int> pl;
    // ordering of the core: minimum degree with dependency-depth tie-breaking (default), VJX_ORDER=amd for the
    // approximate-minimum-degree ordering of round 1
    const char* ord = getenv("VJX_ORDER");
    if (ord && std::strcmp(ord, "amd") == 0) vjx_amd_order(nc, lip.data(), lix.data(), pl);
    else vjx_mdh_order(nc, lip.data(), lix.data(), pl);
    if ((int)pl.size() != nc) { err = "ordering failed"; return -1; }
    for (int a = 0; a < nc; ++a) q.push_back(core_cols[pl[a]]);
  }
  if ((int)q.size() != n) { err = "ordering failed"; return -1; }

  // L: unit lower, column k holds original row ids (mapped through pinv at the end); U: upper.
  std::vector<int> Lp(n + 1, 0), Up(n + 1, 0), Li, Ui;
  std::vector<double> Lx, Ux;
  Li.reserve(4 * (size_t)A.p[n]);
  Lx.reserve(4 * (size_t)A.p[n]);
  Ui.reserve(4 * (size_t)A.p[n]);
  Ux.reserve(4 * (size_t)A.p[n]);
  std::vector<int> pinv(n, -1), xi(2 * (size_t)n), mark(n, -1);
  std::vector<double> x(n, 0.0);
  std::vector<int> stack_pos(n);
  for (int k = 0; k < n; ++k) {
    Lp[k] = (int)Li.size();
    Up[k] = (int)Ui.size();
    const int col = q[k];
    // --- reach: DFS from the rows of A(:,col) through the graph of L
    int top = n;
    for (int pa = A.p[col]; pa < A.p[col + 1]; ++pa) {
      int r0 = A.i[pa];
      if (mark[r0] == k) continue;
      // iterative DFS
      int head = 0;
      xi[n + head] = r0;
      while (head >= 0) {
        int j = xi[n + head];
        int J = pinv[j];
        if (mark[j] != k) {
          mark[j] = k;
          stack_pos[j] = (J < 0) ? 0 : Lp[J] + 1;  // skip the unit diagonal entry
        }
        bool done = true;
        if (J >= 0) {
          int p2 = Lp[J + 1];
          for (int pp = stack_pos[j]; pp < p2; ++pp) {
            int i2 = Li[pp];
            if (mark[i2] == k) continue;
            stack_pos[j] = pp + 1;
            xi[n + (++head)] = i2;
            done = false;
            break;
          }
        }
        if (done) {
          head--;
          xi[--top] = j;
        }
      }
    }
    // --- numeric sparse triangular solve x = L \ A(:,col)
    for (int p2 = top; p2 < n; ++p2) x[xi[p2]] = 0.0;
    for (int pa = A.p[col]; pa < A.p[col + 1]; ++pa) x[A.i[pa]] = A.x[pa];
    for (int px = top; px < n; ++px) {
      int j = xi[px];
      int J = pinv[j];
      if (J < 0) continue;
      double xj = x[j];
      for (int pp = Lp[J] + 1; pp < Lp[J + 1]; ++pp) x[Li[pp]] -= Lx[pp] * xj;
    }
    // --- pivot search among non-pivotal rows
    int ipiv = -1;
    double amax = -1.0;
    for (int p2 = top; p2 < n; ++p2) {
      int i2 = xi[p2];
      if (pinv[i2] < 0) {
        double t = std::fabs(x[i2]);
        if (t > amax) { amax = t; ipiv = i2; }
      } else {
        Ui.push_back(pinv[i2]);
        Ux.push_back(x[i2]);
      }
    }
    if (ipiv == -1 || !(amax > 0.0)) {
      // structurally present but numerically zero column: accept a zero-valued structural
      // candidate if any (pattern is what matters for refactorisation), else fail.
      if (ipiv == -1) { err = "structurally singular at column " + std::to_string(col); return -4; }
    }
    {
      // prefer the paired row (the diagonal for ordinary MNA unknowns; the peeled row for a
      // singleton pivot, which is taken unconditionally unless it is numerically zero)
      const int pr = pref_row[col];
      if (pr >= 0 && pinv[pr] < 0 && mark[pr] == k && std::fabs(x[pr]) > 0.0 &&
          (k < n_single || std::fabs(x[pr]) >= amax * pivot_tol))
        ipiv = pr;
    }
    double pivot = x[ipiv];
    Ui.push_back(k);
    Ux.push_back(pivot);
    pinv[ipiv] = k;
    Li.push_back(ipiv);
    Lx.push_back(1.0);
    for (int p2 = top; p2 < n; ++p2) {
      int i2 = xi[p2];
      if (pinv[i2] < 0) {
        Li.push_back(i2);
        Lx.push_back(pivot != 0.0 ? x[i2] / pivot : 0.0);
      }
      x[i2] = 0.0;
    }
  }
  Lp[n] = (int)Li.size();
  Up[n] = (int)Ui.size();
  for (auto& r : Li) r = pinv[r];

  // --- permutations
  S.perm_r.assign(n, 0);
  for (int r = 0; r < n; ++r) S.perm_r[pinv[r]] = r;
  S.perm_c = q;
  std::vector<int> qinv(n);
  for (int j = 0; j < n; ++j) qinv[q[j]] = j;

  // --- row-wise L+U pattern in the permuted space
  std::vector<int> rc(n, 0);
  for (int k = 0; k < n; ++k) {
    for (int pp = Lp[k] + 1; pp < Lp[k + 1]; ++pp) rc[Li[pp]]++;   // L(i,k), i > k
    for (int pp = Up[k]; pp < Up[k + 1]; ++pp) rc[Ui[pp]]++;        // U(i,k), i <= k
  }
  S.lu_ptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) S.lu_ptr[i + 1] = S.lu_ptr[i] + rc[i];
  S.nnz_lu = S.lu_ptr[n];
  S.lu_col.assign((size_t)S.nnz_lu, 0);
  std::vector<int> nx(S.lu_ptr.begin(), S.lu_ptr.end() - 1);
  for (int k = 0; k < n; ++k) {  // ascending k => ascending columns per row
    for (int pp = Up[k]; pp < Up[k + 1]; ++pp) S.lu_col[nx[Ui[pp]]++] = k;
    for (int pp = Lp[k] + 1; pp < Lp[k + 1]; ++pp) S.lu_col[nx[Li[pp]]++] = k;
  }
  // rows were filled column by column, so each row is already sorted ascending
  S.lu_diag.assign(n, -1);
  S.max_row_len = 0;
  for (int i = 0; i < n; ++i) {
    int b = S.lu_ptr[i], e = S.lu_ptr[i + 1];
    S.max_row_len = std::max(S.max_row_len, e - b);
    auto it = std::lower_bound(S.lu_col.begin() + b, S.lu_col.begin() + e, i);
    if (it == S.lu_col.begin() + e || *it != i) { err = "missing diagonal in LU pattern"; return -4; }
    S.lu_diag[i] = (int)(it - S.lu_col.begin());
  }
  // --- map A's CSR slots into L+U
  S.lu_src.assign((size_t)S.nnz_lu, -1);
  for (int r = 0; r < n; ++r) {
    int i = pinv[r];
    int b = S.lu_ptr[i], e = S.lu_ptr[i + 1];
    for (int kk = indptr[r]; kk < indptr[r + 1]; ++kk) {
      int j = qinv[indices[kk]];
      auto it = std::lower_bound(S.lu_col.begin() + b, S.lu_col.begin() + e, j);
      if (it == S.lu_col.begin() + e || *it != j) { err = "A entry outside LU pattern"; return -1; }
      S.lu_src[it - S.lu_col.begin()] = kk;
    }
  }
  // --- dependency levels
  S.level_l.assign(n, 0);
  S.level_u.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    int lv = 0;
    for (int pp = S.lu_ptr[i]; pp < S.lu_diag[i]; ++pp) lv = std::max(lv, S.level_l[S.lu_col[pp]] + 1);
    S.level_l[i] = lv;
    S.n_levels_l = std::max(S.n_levels_l, lv + 1);
  }
  for (int i = n - 1; i >= 0; --i) {
    int lv = 0;
    for (int pp = S.lu_diag[i] + 1; pp < S.lu_ptr[i + 1]; ++pp) lv = std::max(lv, S.level_u[S.lu_col[pp]] + 1);
    S.level_u[i] = lv;
    S.n_levels_u = std::max(S.n_levels_u, lv + 1);
  }
  S.order_l.resize(n);
  S.order_u.resize(n);
  std::iota(S.order_l.begin(), S.order_l.end(), 0);
  std::iota(S.order_u.begin(), S.order_u.end(), 0);
  std::stable_sort(S.order_l.begin(), S.order_l.end(), [&](int a, int b) { return S.level_l[a] < S.level_l[b]; });
  std::stable_sort(S.order_u.begin(), S.order_u.end(), [&](int a, int b) {
    if (S.level_u[a] != S.level_u[b]) return S.level_u[a] < S.level_u[b];
    return a > b;
  });
  // --- chain tail (see vjx_symbolic.h)
  {
    const int kMaxPerLevel = 16, kMaxTail = 4096;
    std::vector<int> cl(S.n_levels_l, 0);
    for (int i = 0; i < n; ++i) cl[S.level_l[i]]++;
    int lcut = S.n_levels_l;   // first level of the all-narrow suffix of levels
    while (lcut > 1 && cl[lcut - 1] <= kMaxPerLevel) --lcut;
    int i0 = n;
    while (i0 > 0 && n - (i0 - 1) <= kMaxTail && S.level_l[i0 - 1] >= lcut) --i0;
    if (n - i0 < 8) i0 = n;
    S.tail_start = i0;
    S.order_f = S.order_l;
    std::vector<int> bl, bu;
    for (int i : S.order_l) if (i < i0) bl.push_back(i);
    const int kLongRow = 1024;
    for (int i : S.order_u) {
      if (i >= i0) continue;
      if (S.lu_ptr[i + 1] - S.lu_diag[i] - 1 > kLongRow) S.long_u.push_back(i);
      else bu.push_back(i);
    }
    S.order_l.swap(bl);
    S.order_u.swap(bu);
  }
  // --- dense trailing block and execution order of the L parts (see vjx_symbolic.h)
  {
    // dense_cap < 0: exactly the longest suffix of rows with an L part, up to |dense_cap| (VJX_DT_ROWS).
    // dense_cap > 0: the block is the *chain* at the end of the elimination order - the suffix of rows that are alone
    // in their dependency level (each waits for the one before it: a global-memory link per row in k_lu_rows, a
    // barrier per row inside the block), rounded up to a whole panel of 16.  Rows in front of the chain have
    // company in their level and cost less as ordinary rows than as rows of the block, whose L parts all finish
    // late (c6288 / psp103w with the minimum-degree + depth ordering, chain of 48: 159 rows 163 us, 64 rows 147 us,
    // 48 rows 143 us, 32 rows 146 us per refactorisation; gpurun_out/r2_w_variants.txt).
    const bool forced = dense_cap < 0;
    if (forced) dense_cap = -dense_cap;
    int nd = 0;
    while (nd < dense_cap && nd < n && S.lu_diag[n - 1 - nd] > S.lu_ptr[n - 1 - nd]) ++nd;
    if (!forced && nd >= 16) {
      std::vector<int> cl(S.n_levels_l, 0);
      for (int i = 0; i < n; ++i) cl[S.level_l[i]]++;
      int nc = 0;
      while (nc < nd && cl[S.level_l[n - 1 - nc]] == 1) ++nc;
      if (nc >= 16) nd = std::min(nd, (nc + 15) / 16 * 16);
    }
    if (nd < 16) nd = 0;
    S.n_dense = nd;
    const int ds = n - nd;
    // depth[k]: dependency depth of row k counting both what it waits for (L(k,j) != 0: row j finished)
    // and what must be applied before it inside any row that holds both (U(j,k) != 0)
    std::vector<int>& depth = S.depth;
    depth.assign(n, 0);
    for (int k = 0; k < n; ++k) {
      int d = depth[k];   // pushed from rows j < k with U(j,k) != 0
      for (int pp = S.lu_ptr[k]; pp < S.lu_diag[k]; ++pp) d = std::max(d, depth[S.lu_col[pp]] + 1);
      depth[k] = d;
      for (int pp = S.lu_diag[k] + 1; pp < S.lu_ptr[k + 1]; ++pp) {
        int& t = depth[S.lu_col[pp]];
        t = std::max(t, d + 1);
      }
    }
    std::vector<int> idx, col2, src2;
    for (int i = 0; i < n; ++i) {
      const int rb = S.lu_ptr[i], dg = S.lu_diag[i];
      if (dg - rb < 2) continue;
      idx.resize(dg - rb);
      std::iota(idx.begin(), idx.end(), rb);
      std::sort(idx.begin(), idx.end(), [&](int a, int b) {
        const int ka = S.lu_col[a], kb = S.lu_col[b];
        const bool da = ka >= ds, db = kb >= ds;
        if (da != db) return db;
        if (depth[ka] != depth[kb]) return depth[ka] < depth[kb];
        return ka < kb;
      });
      col2.resize(dg - rb);
      src2.resize(dg - rb);
      for (int t = 0; t < dg - rb; ++t) { col2[t] = S.lu_col[idx[t]]; src2[t] = S.lu_src[idx[t]]; }
      std::copy(col2.begin(), col2.end(), S.lu_col.begin() + rb);
      std::copy(src2.begin(), src2.end(), S.lu_src.begin() + rb);
    }
  }
  // --- flops and the scatter program
  double fl = 0.0;
  int64_t nupd = 0;
  for (int i = 0; i < n; ++i) {
    int64_t c = 0;
    for (int pp = S.lu_ptr[i]; pp < S.lu_diag[i]; ++pp) {
      int k = S.lu_col[pp];
      c += S.lu_ptr[k + 1] - S.lu_diag[k] - 1;
    }
    nupd += c;
    fl += 2.0 * (double)c + (double)(S.lu_diag[i] - S.lu_ptr[i]);
    S.max_l_len = std::max(S.max_l_len, S.lu_diag[i] - S.lu_ptr[i]);
  }
  S.flops = fl;
  if (build_update_stream) {
    if (nupd > (int64_t)8000000000LL) { err = "scatter program too large"; return -1; }
    S.upd_ptr.assign((size_t)S.nnz_lu + 1, 0);
    S.upd_pos.resize((size_t)nupd);
    std::vector<int> pos_of(n, -1);
    int64_t o = 0;
    for (int i = 0; i < n; ++i) {
      const int b = S.lu_ptr[i], e = S.lu_ptr[i + 1];
      for (int pp = b; pp < e; ++pp) pos_of[S.lu_col[pp]] = pp - b;
      for (int pp = b; pp < e; ++pp) {
        S.upd_ptr[(size_t)pp] = o;
        if (pp >= S.lu_diag[i]) continue;
        const int k = S.lu_col[pp];
        for (int qq = S.lu_diag[k] + 1; qq < S.lu_ptr[k + 1]; ++qq) {
          const int t = pos_of[S.lu_col[qq]];
          if (t < 0) { err = "LU pattern not closed under elimination"; return -1; }
          S.upd_pos[(size_t)o++] = t;
        }
      }
      for (int pp = b; pp < e; ++pp) pos_of[S.lu_col[pp]] = -1;
    }
    S.upd_ptr[(size_t)S.nnz_lu] = o;
  }
  return 0;
}

int vjx_row_lend(const VjxSymbolic& S, int i) {
  const int ds = S.n - S.n_dense;
  const int dg = S.lu_diag[i];
  if (i < ds) return dg;
  int p = S.lu_ptr[i];
  while (p < dg && S.lu_col[p] < ds) ++p;   // columns of the dense block are stored last (execution order)
  return p;
}

std::vector<int> vjx_select_heavy(const VjxSymbolic& S, int t0, int t1, int min_ksteps, int max_len) {
  std::vector<int> rows;
  if (min_ksteps <= 0 || S.upd_ptr.empty() || S.nnz_lu >= (int64_t)(1 << 26)) return rows;
  for (int t = t0; t < t1; ++t) {
    const int i = S.order_f[t];
    if (S.lu_ptr[i + 1] - S.lu_ptr[i] > max_len) continue;
    if (vjx_row_lend(S, i) - S.lu_ptr[i] >= min_ksteps) rows.push_back(i);
  }
  return rows;
}

void vjx_build_heavy(VjxSymbolic& S, const std::vector<int>& rows) {
  VjxSymbolic::Heavy& H = S.heavy;
  H = VjxSymbolic::Heavy();
  H.rows = rows;
  H.off.assign(rows.size() + 1, 0);
  struct Contrib { int pos, slot, q; };
  std::vector<Contrib> cs;
  std::vector<int> tstart, torder;
  for (size_t r = 0; r < rows.size(); ++r) {
    const int i = rows[r];
    const int rb = S.lu_ptr[i], lend = vjx_row_lend(S, i);
    H.off[r] = (long long)H.prog.size();
    int p0 = rb;
    while (p0 < lend) {
      const int d0 = S.depth[S.lu_col[p0]];
      int p1 = p0 + 1;
      while (p1 < lend && p1 - p0 < 32 && S.depth[S.lu_col[p1]] == d0) ++p1;
      // contributions of the group's k-steps, regrouped by target position (ascending k per position)
      cs.clear();
      for (int p = p0; p < p1; ++p) {
        const int k = S.lu_col[p];
        const int ub = S.lu_diag[k] + 1, ue = S.lu_ptr[k + 1];
        const int64_t mb = S.upd_ptr[(size_t)p];
        for (int q = ub; q < ue; ++q) cs.push_back({S.upd_pos[(size_t)(mb + (q - ub))], p - p0, q});
      }
      std::stable_sort(cs.begin(), cs.end(), [](const Contrib& a, const Contrib& b) {
        return a.pos != b.pos ? a.pos < b.pos : a.slot < b.slot;
      });
      tstart.clear();
      for (size_t c = 0; c < cs.size(); ++c)
        if (c == 0 || cs[c].pos != cs[c - 1].pos) tstart.push_back((int)c);
      tstart.push_back((int)cs.size());
      const int nt = (int)tstart.size() - 1;
      H.n_contrib += (long long)cs.size();
      H.n_ksteps += p1 - p0;
      for (int round = 0;; ++round) {
        torder.clear();
        for (int t = 0; t < nt; ++t)
          if (tstart[t + 1] - tstart[t] > 4 * round) torder.push_back(t);
        if (torder.empty() && round > 0) break;
        // positions with many contributions left first: lanes of one record then have similar counts
        std::stable_sort(torder.begin(), torder.end(), [&](int a, int b) {
          const int ca = std::min(4, tstart[a + 1] - tstart[a] - 4 * round), cb = std::min(4, tstart[b + 1] - tstart[b] - 4 * round);
          return ca > cb;
        });
        const int nrec = ((int)torder.size() + 31) / 32;
        // record r: pos[32] (lane 0 also carries the record's depth R in bits 24+), then R rows of 32 entries;
        // R = contributions the record's first lane still has (the largest count of the record), at most 4
        const size_t g0 = H.prog.size();
        size_t gi = VJX_HV_GRP_INTS + (size_t)nrec;
        H.prog.resize(g0 + gi, -1);
        H.prog[g0 + 0] = nrec; H.prog[g0 + 1] = round > 0 ? 1 : 0; H.prog[g0 + 2] = round > 0 ? 0 : p1 - p0;
        if (round == 0)
          for (int p = p0; p < p1; ++p) { H.prog[g0 + 4 + (p - p0)] = p - rb; H.prog[g0 + 36 + (p - p0)] = S.lu_col[p]; }
        for (int rr = 0; rr < nrec; ++rr) {
          const int t0 = torder[(size_t)rr * 32];
          const int R = std::min(4, tstart[t0 + 1] - tstart[t0] - 4 * round);
          H.prog[g0 + VJX_HV_GRP_INTS + rr] = (int)gi | (R << 28);
          H.prog.resize(g0 + gi + 32 * (size_t)(1 + R), -1);
          int* Rp = H.prog.data() + g0 + gi;
          for (int lane = 0; lane < 32 && (size_t)rr * 32 + lane < torder.size(); ++lane) {
            const int t = torder[(size_t)rr * 32 + lane];
            Rp[lane] = cs[tstart[t]].pos;
            for (int j = 0; j < R; ++j) {
              const int c = tstart[t] + 4 * round + j;
              if (c < tstart[t + 1]) Rp[32 + 32 * j + lane] = cs[c].slot | (cs[c].q << 5);
            }
          }
          gi += 32 * (size_t)(1 + R);
          H.n_slots += 32 * R;
        }
        H.prog[g0 + 3] = (int)gi;
        H.n_groups++;
        H.n_records += nrec;
        if (torder.empty()) break;
      }
      p0 = p1;
    }
  }
  H.off[rows.size()] = (long long)H.prog.size();
}

// One heavy row exactly as the thread block of k_lu_rows walks it (host emulation for the CPU tests).
static void heavy_row_emulate(const VjxSymbolic& S, size_t r, const double* values, std::vector<double>& v,
                              const std::vector<double>& inv) {
  const int i = S.heavy.rows[r];
  const int rb = S.lu_ptr[i], dg = S.lu_diag[i], re = S.lu_ptr[i + 1], lend = vjx_row_lend(S, i);
  std::vector<double> w(re - rb);
  for (int pp = rb; pp < re; ++pp) w[pp - rb] = S.lu_src[pp] >= 0 ? values[S.lu_src[pp]] : 0.0;
  const int* P = S.heavy.prog.data() + S.heavy.off[r];
  const int* Pend = S.heavy.prog.data() + S.heavy.off[r + 1];
  double l[32] = {0};
  while (P < Pend) {
    const int nrec = P[0], keep = P[1];
    if (!keep)
      for (int s = 0; s < 32; ++s) {
        const int kpos = P[4 + s];
        l[s] = kpos >= 0 ? w[kpos] * inv[P[36 + s]] : 0.0;
        if (kpos >= 0) v[rb + kpos] = l[s];
      }
    for (int rr = 0; rr < nrec; ++rr) {
      const int info = P[VJX_HV_GRP_INTS + rr];
      const int* R = P + (info & 0x0fffffff);
      const int depth = info >> 28;
      for (int lane = 0; lane < 32; ++lane) {
        const int pos = R[lane];
        if (pos < 0) continue;
        double acc = w[pos];
        for (int j = 0; j < depth; ++j) {
          const int e = R[32 + 32 * j + lane];
          if (e != -1) acc -= l[e & 31] * v[(size_t)(((unsigned)e) >> 5)];
        }
        w[pos] = acc;
      }
    }
    P += P[3];
  }
  // the k-steps inside the dense block (dense rows only), in stored order
  for (int pp = lend; pp < dg; ++pp) {
    const int k = S.lu_col[pp];
    const double lk = w[pp - rb] * inv[k];
    const int ub = S.lu_diag[k] + 1, ue = S.lu_ptr[k + 1];
    const int64_t mb = S.upd_ptr[(size_t)pp];
    for (int q = ub; q < ue; ++q) w[S.upd_pos[(size_t)(mb + (q - ub))]] -= lk * v[q];
    v[pp] = lk;
  }
  for (int pp = dg; pp < re; ++pp) v[pp] = w[pp - rb];
}

double vjx_symbolic_selfcheck(const VjxSymbolic& S, int n, const int* indptr, const int* indices,
                              const double* values, const double* b, double* x_out) {
  std::vector<double> v((size_t)S.nnz_lu, 0.0), inv(n, 0.0), y(n, 0.0), xp(n, 0.0);
  std::vector<int> hv_of(n, -1);
  for (size_t r = 0; r < S.heavy.rows.size(); ++r) hv_of[S.heavy.rows[r]] = (int)r;
  for (int t = 0; t < n; ++t) {
    const int i = S.order_f[t];
    const int rb = S.lu_ptr[i], dg = S.lu_diag[i], re = S.lu_ptr[i + 1];
    if (hv_of[i] >= 0) {
      heavy_row_emulate(S, (size_t)hv_of[i], values, v, inv);
      inv[i] = 1.0 / v[dg];
      continue;
    }
    double* w = v.data() + rb;
    for (int pp = rb; pp < re; ++pp) w[pp - rb] = S.lu_src[pp] >= 0 ? values[S.lu_src[pp]] : 0.0;
    for (int pp = rb; pp < dg; ++pp) {
      const int k = S.lu_col[pp];
      const double l = w[pp - rb] * inv[k];
      const int ub = S.lu_diag[k] + 1, ue = S.lu_ptr[k + 1];
      const int64_t mb = S.upd_ptr[(size_t)pp];
      for (int q = ub; q < ue; ++q) w[S.upd_pos[(size_t)(mb + (q - ub))]] -= l * v[q];
      w[pp - rb] = l;
    }
    inv[i] = 1.0 / w[dg - rb];
  }
  // any topological order gives the same per-row arithmetic: ascending rows for L, descending for U
  for (int i = 0; i < n; ++i) {
    double acc = b[S.perm_r[i]];
    for (int pp = S.lu_ptr[i]; pp < S.lu_diag[i]; ++pp) acc -= v[pp] * y[S.lu_col[pp]];
    y[i] = acc;
  }
  for (int i = n - 1; i >= 0; --i) {
    double acc = y[i];
    for (int pp = S.lu_diag[i] + 1; pp < S.lu_ptr[i + 1]; ++pp) acc -= v[pp] * xp[S.lu_col[pp]];
    xp[i] = acc * inv[i];
  }
  {
    // the schedule tables must cover every row exactly once
    std::vector<char> seen_l(n, 0), seen_u(n, 0);
    for (int i : S.order_l) seen_l[i]++;
    for (int i : S.order_u) seen_u[i]++;
    for (int i : S.long_u) seen_u[i]++;
    for (int i = S.tail_start; i < n; ++i) { seen_l[i]++; seen_u[i]++; }
    for (int i = 0; i < n; ++i)
      if (seen_l[i] != 1 || seen_u[i] != 1) return 1e300;
  }
  std::vector<double> x(n);
  for (int j = 0; j < n; ++j) x[S.perm_c[j]] = xp[j];
  double rmax = 0.0;
  for (int r = 0; r < n; ++r) {
    double acc = -b[r];
    for (int kk = indptr[r]; kk < indptr[r + 1]; ++kk) acc += values[kk] * x[indices[kk]];
    rmax = std::max(rmax, std::fabs(acc));
  }
  if (x_out) std::memcpy(x_out, x.data(), sizeof(double) * n);
  return rmax;
}

// C hooks for the CPU-only tests of the host logic (ctypes).
extern "C" int vjx_symbolic_probe(int n, const int* indptr, const int* indices, const double* values,
                                  double pivot_tol, const double* b, double* x_out, double* stats8 /* [12] */) {
  VjxSymbolic S;
  std::string err;
  const int dcap = getenv("VJX_PROBE_DENSE_CAP") ? atoi(getenv("VJX_PROBE_DENSE_CAP")) : 0;
  int rc = vjx_symbolic_analyze(n, indptr, indices, values, pivot_tol, true, S, err, dcap);
  if (rc != 0) return rc;
  if (getenv("VJX_PROBE_HV_MIN"))   // CPU tests of the heavy-row program: build it and let the emulation walk it
    vjx_build_heavy(S, vjx_select_heavy(S, 0, n, atoi(getenv("VJX_PROBE_HV_MIN")), 1 << 30));
  double r = vjx_symbolic_selfcheck(S, n, indptr, indices, values, b, x_out);
  if (stats8) {
    stats8[0] = (double)S.nnz_lu;
    stats8[1] = S.flops;
    stats8[2] = S.n_levels_l;
    stats8[3] = S.n_levels_u;
    stats8[4] = S.max_row_len;
    stats8[5] = (double)n;
    stats8[6] = r;
    stats8[7] = (double)S.upd_pos.size();
    stats8[8] = S.n_singletons;
    stats8[9] = S.n_core;
    stats8[10] = S.max_l_len;
    stats8[11] = S.n - S.tail_start;
  }
  if (getenv("VJX_PROBE_HV_MIN") && getenv("VJX_PROBE_VERBOSE"))
    fprintf(stderr, "heavy rows %zu  k-steps %d  groups %d  records %d  contributions %lld  fill %.3f  program %.1f MB\n",
            S.heavy.rows.size(), S.heavy.n_ksteps, S.heavy.n_groups, S.heavy.n_records, S.heavy.n_contrib,
            S.heavy.n_slots ? (double)S.heavy.n_contrib / (double)S.heavy.n_slots : 0.0, 4e-6 * (double)S.heavy.prog.size());
  return 0;
}

// Per-row diagnostics of the analysis (CPU-only tests and offline tuning), indexed by permuted row: L / U
// dependency level, number of L and U entries, and for the k-steps the row kernel executes itself (the L entries
// left of the dense trailing block of `dense_cap` rows): their number (bulk_len) and the number of distinct
// dependency depths among them (waves) - k-steps of equal depth are independent of each other, so `waves` is the
// length of the row's sequential chain if a depth were processed at a time (DESIGN.md 8, round-2 plan).
extern "C" int vjx_symbolic_rows2(int n, const int* indptr, const int* indices, const double* values,
                                  double pivot_tol, int dense_cap, int* level_l, int* level_u, int* l_len, int* u_len,
                                  int* waves, int* bulk_len, int* n_dense_out) {
  VjxSymbolic S;
  std::string err;
  int rc = vjx_symbolic_analyze(n, indptr, indices, values, pivot_tol, true, S, err, dense_cap);
  if (rc != 0) return rc;
  const int ds = n - S.n_dense;
  if (n_dense_out) *n_dense_out = S.n_dense;
  for (int i = 0; i < n; ++i) {
    if (level_l) level_l[i] = S.level_l[i];
    if (level_u) level_u[i] = S.level_u[i];
    if (l_len) l_len[i] = S.lu_diag[i] - S.lu_ptr[i];
    if (u_len) u_len[i] = S.lu_ptr[i + 1] - S.lu_diag[i] - 1;
    int w = 0, last = -1, nb = 0;
    for (int pp = S.lu_ptr[i]; pp < S.lu_diag[i]; ++pp) {
      const int k = S.lu_col[pp];
      if (i >= ds && k >= ds) continue;   // handed to the dense kernel
      ++nb;
      if (S.depth[k] != last) { ++w; last = S.depth[k]; }
    }
    if (waves) waves[i] = w;
    if (bulk_len) bulk_len[i] = nb;
  }
  return 0;
}

extern "C" int vjx_symbolic_rows(int n, const int* indptr, const int* indices, const double* values,
                                 double pivot_tol, int* level_l, int* level_u, int* l_len, int* u_len, int* waves) {
  return vjx_symbolic_rows2(n, indptr, indices, values, pivot_tol, 0, level_l, level_u, l_len, u_len, waves, nullptr, nullptr);
}
