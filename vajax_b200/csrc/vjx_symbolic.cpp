// See vjx_symbolic.h.  Plain C++17, no CUDA.
#include "vjx_symbolic.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

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
}  // namespace

int vjx_symbolic_analyze(int n, const int* indptr, const int* indices, const double* values,
                         double pivot_tol, bool build_update_stream, VjxSymbolic& S, std::string& err) {
  S = VjxSymbolic();
  S.n = n;
  if (n <= 0) { err = "empty matrix"; return -1; }
  std::vector<int> q;
  vjx_amd_order(n, indptr, indices, q);
  if ((int)q.size() != n) { err = "ordering failed"; return -1; }
  CSC A = csr_to_csc(n, indptr, indices, values);

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
    if (pinv[col] < 0 && mark[col] == k && std::fabs(x[col]) >= amax * pivot_tol && std::fabs(x[col]) > 0.0)
      ipiv = col;
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
  // --- flops and the gather program
  double fl = 0.0;
  int64_t npairs = 0;
  for (int i = 0; i < n; ++i) {
    int64_t c = 0;
    for (int pp = S.lu_ptr[i]; pp < S.lu_diag[i]; ++pp) {
      int k = S.lu_col[pp];
      c += S.lu_ptr[k + 1] - S.lu_diag[k] - 1;
    }
    npairs += c;
    fl += 2.0 * (double)c + (double)(S.lu_diag[i] - S.lu_ptr[i]);
  }
  S.flops = fl;
  if (build_update_stream) {
    if (npairs > (int64_t)2000000000LL) { err = "gather program too large"; return -1; }
    S.pair_ptr.assign((size_t)S.nnz_lu + 1, 0);
    std::vector<int> pos_of(n, -1);
    // pass 1: count contributors per entry
    for (int i = 0; i < n; ++i) {
      int b = S.lu_ptr[i], e = S.lu_ptr[i + 1];
      for (int pp = b; pp < e; ++pp) pos_of[S.lu_col[pp]] = pp;
      for (int pp = b; pp < S.lu_diag[i]; ++pp) {
        int k = S.lu_col[pp];
        for (int qq = S.lu_diag[k] + 1; qq < S.lu_ptr[k + 1]; ++qq) {
          int t = pos_of[S.lu_col[qq]];
          if (t < 0) { err = "LU pattern not closed under elimination"; return -1; }
          S.pair_ptr[(size_t)t + 1]++;
        }
      }
      for (int pp = b; pp < e; ++pp) pos_of[S.lu_col[pp]] = -1;
    }
    for (int64_t p2 = 0; p2 < S.nnz_lu; ++p2) S.pair_ptr[(size_t)p2 + 1] += S.pair_ptr[(size_t)p2];
    S.pair_l.resize((size_t)npairs);
    S.pair_u.resize((size_t)npairs);
    std::vector<int64_t> fillp(S.pair_ptr.begin(), S.pair_ptr.end() - 1);
    std::vector<int> elev((size_t)S.nnz_lu, 0);
    S.ent_order.resize((size_t)S.nnz_lu);
    S.row_wave_ptr.assign((size_t)n + 1, 0);
    S.wave_end.clear();
    std::vector<int> cnt;
    for (int i = 0; i < n; ++i) {
      int b = S.lu_ptr[i], e = S.lu_ptr[i + 1];
      for (int pp = b; pp < e; ++pp) pos_of[S.lu_col[pp]] = pp;
      int maxlev = 0;
      for (int pp = b; pp < S.lu_diag[i]; ++pp) {   // ascending k: elev[pp] is final here
        int k = S.lu_col[pp];
        for (int qq = S.lu_diag[k] + 1; qq < S.lu_ptr[k + 1]; ++qq) {
          int t = pos_of[S.lu_col[qq]];
          int64_t o = fillp[t]++;
          S.pair_l[(size_t)o] = pp;
          S.pair_u[(size_t)o] = qq;
          if (elev[t] < elev[pp] + 1) elev[t] = elev[pp] + 1;
        }
      }
      for (int pp = b; pp < e; ++pp) maxlev = std::max(maxlev, elev[pp]);
      // counting sort of the row's entries by wave (stable: ascending column inside a wave)
      cnt.assign((size_t)maxlev + 2, 0);
      for (int pp = b; pp < e; ++pp) cnt[(size_t)elev[pp] + 1]++;
      for (int l = 0; l <= maxlev; ++l) cnt[(size_t)l + 1] += cnt[(size_t)l];
      S.row_wave_ptr[i] = (int)S.wave_end.size();
      for (int l = 0; l <= maxlev; ++l) S.wave_end.push_back(b + cnt[(size_t)l + 1]);
      std::vector<int> nx2(cnt.begin(), cnt.end() - 1);
      for (int pp = b; pp < e; ++pp) S.ent_order[(size_t)b + nx2[(size_t)elev[pp]]++] = pp;
      S.max_waves = std::max(S.max_waves, maxlev + 1);
      for (int pp = b; pp < e; ++pp) pos_of[S.lu_col[pp]] = -1;
    }
    S.row_wave_ptr[n] = (int)S.wave_end.size();
  }
  return 0;
}

double vjx_symbolic_selfcheck(const VjxSymbolic& S, int n, const int* indptr, const int* indices,
                              const double* values, const double* b, double* x_out) {
  std::vector<double> v((size_t)S.nnz_lu, 0.0), inv(n, 0.0), y(n, 0.0), xp(n, 0.0);
  for (int t = 0; t < n; ++t) {
    int i = S.order_l[t];
    int w0 = S.row_wave_ptr[i], w1 = S.row_wave_ptr[i + 1];
    int eb = S.lu_ptr[i];
    for (int w = w0; w < w1; ++w) {
      int ee = S.wave_end[w];
      for (int q = eb; q < ee; ++q) {
        int pos = S.ent_order[q];
        double acc = S.lu_src[pos] >= 0 ? values[S.lu_src[pos]] : 0.0;
        for (int64_t pr = S.pair_ptr[pos]; pr < S.pair_ptr[(size_t)pos + 1]; ++pr)
          acc -= v[S.pair_l[(size_t)pr]] * v[S.pair_u[(size_t)pr]];
        if (pos < S.lu_diag[i]) acc *= inv[S.lu_col[pos]];
        v[pos] = acc;
      }
      eb = ee;
    }
    inv[i] = 1.0 / v[S.lu_diag[i]];
  }
  for (int t = 0; t < n; ++t) {
    int i = S.order_l[t];
    double acc = b[S.perm_r[i]];
    for (int pp = S.lu_ptr[i]; pp < S.lu_diag[i]; ++pp) acc -= v[pp] * y[S.lu_col[pp]];
    y[i] = acc;
  }
  for (int t = 0; t < n; ++t) {
    int i = S.order_u[t];
    double acc = y[i];
    for (int pp = S.lu_diag[i] + 1; pp < S.lu_ptr[i + 1]; ++pp) acc -= v[pp] * xp[S.lu_col[pp]];
    xp[i] = acc * inv[i];
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
                                  double pivot_tol, const double* b, double* x_out, double* stats8) {
  VjxSymbolic S;
  std::string err;
  int rc = vjx_symbolic_analyze(n, indptr, indices, values, pivot_tol, true, S, err);
  if (rc != 0) return rc;
  double r = vjx_symbolic_selfcheck(S, n, indptr, indices, values, b, x_out);
  if (stats8) {
    stats8[0] = (double)S.nnz_lu;
    stats8[1] = S.flops;
    stats8[2] = S.n_levels_l;
    stats8[3] = S.n_levels_u;
    stats8[4] = S.max_row_len;
    stats8[5] = (double)n;
    stats8[6] = r;
    stats8[7] = (double)S.pair_l.size();
  }
  return 0;
}
