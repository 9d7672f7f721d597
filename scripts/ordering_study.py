#!/usr/bin/env python
"""CPU-only study of fill-reducing orderings for the LU core: is there an ordering whose
elimination tree is markedly lower than AMD's at comparable fill?  Test / tuning tool, not on the product path.

    python scripts/lu_structure.py 16      # writes /tmp/jac_psp103w_16.npz (the Jacobian at a mid-rail state)
    python scripts/ordering_study.py /tmp/jac_psp103w_16.npz

Works on the pattern of A + A^T of the core left after singleton peeling (the same first phase as
csrc/vjx_symbolic.cpp) and reports, per ordering, nnz(L) of the symbolic Cholesky factor (the LU pattern of a
diagonally pivoted structurally symmetric matrix is L + L^T), the height of the elimination tree (= the number of
dependency levels of the factorisation) and the size of the trailing chain (levels holding a single row)."""
import sys
from collections import deque

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
from scipy.sparse.csgraph import reverse_cuthill_mckee


def peel(A):
    """Row / column singletons, iteratively.  Returns the indices (rows, cols) of the core."""
    A = A.tocsr()
    n = A.shape[0]
    At = A.tocsc()
    row_alive, col_alive = np.ones(n, bool), np.ones(n, bool)
    rcnt = np.diff(A.indptr).astype(np.int64)
    ccnt = np.diff(At.indptr).astype(np.int64)
    q = deque([("r", i) for i in np.where(rcnt == 1)[0]] + [("c", j) for j in np.where(ccnt == 1)[0]])
    while q:
        kind, x = q.popleft()
        if kind == "r":
            if not row_alive[x] or rcnt[x] != 1:
                continue
            cols = [j for j in A.indices[A.indptr[x]:A.indptr[x + 1]] if col_alive[j]]
            r, c = x, cols[0]
        else:
            if not col_alive[x] or ccnt[x] != 1:
                continue
            rows = [i for i in At.indices[At.indptr[x]:At.indptr[x + 1]] if row_alive[i]]
            r, c = rows[0], x
        row_alive[r] = col_alive[c] = False
        for j in A.indices[A.indptr[r]:A.indptr[r + 1]]:
            if col_alive[j]:
                ccnt[j] -= 1
                if ccnt[j] == 1:
                    q.append(("c", j))
        for i in At.indices[At.indptr[c]:At.indptr[c + 1]]:
            if row_alive[i]:
                rcnt[i] -= 1
                if rcnt[i] == 1:
                    q.append(("r", i))
    return np.where(row_alive)[0], np.where(col_alive)[0]


def symbolic(G, perm):
    """nnz(L) incl. diagonal, etree height, per-level row counts for the symmetric pattern G under perm."""
    n = G.shape[0]
    P = G[perm][:, perm].tocsr()
    parent = np.full(n, -1, np.int64)
    anc = np.full(n, -1, np.int64)
    for i in range(n):
        for k in P.indices[P.indptr[i]:P.indptr[i + 1]]:
            while k != -1 and k < i:
                nxt = anc[k]
                anc[k] = i
                if nxt == -1:
                    parent[k] = i
                k = nxt
    # row counts by walking the row subtrees
    mark = np.full(n, -1, np.int64)
    nnz = n
    colcnt = np.ones(n, np.int64)
    for i in range(n):
        mark[i] = i
        for k in P.indices[P.indptr[i]:P.indptr[i + 1]]:
            while k < i and mark[k] != i:
                mark[k] = i
                nnz += 1
                colcnt[k] += 1
                k = parent[k]
    depth = np.zeros(n, np.int64)   # level of row i = 1 + max level of its etree children
    for k in range(n):
        if parent[k] >= 0:
            depth[parent[k]] = max(depth[parent[k]], depth[k] + 1)
    levels = int(depth.max()) + 1
    cnt = np.bincount(depth, minlength=levels)
    flops = float(np.sum(colcnt.astype(float) ** 2))
    return nnz, levels, int(np.sum(cnt == 1)), flops


def bfs_levels(adj_ptr, adj_idx, verts_mask, start):
    lvl = {start: 0}
    order = [start]
    dq = deque([start])
    while dq:
        u = dq.popleft()
        for v in adj_idx[adj_ptr[u]:adj_ptr[u + 1]]:
            if verts_mask[v] and v not in lvl:
                lvl[v] = lvl[u] + 1
                order.append(v)
                dq.append(v)
    return lvl, order


def nested_dissection(G, leaf=96):
    """Level-set nested dissection: separator = the middle BFS level from a pseudo-peripheral vertex, thinned to the
    vertices that really touch the far side; leaves in minimum-degree-ish (RCM-reversed) order."""
    G = G.tocsr()
    n = G.shape[0]
    ptr, idx = G.indptr, G.indices
    mask = np.zeros(n, bool)
    out = []

    def order_leaf(verts):
        if len(verts) <= 2:
            return list(verts)
        sub = G[verts][:, verts]
        p = amd_like(sub)
        return [verts[i] for i in p]

    def rec(verts):
        if len(verts) <= leaf:
            return order_leaf(verts)
        mask[:] = False
        mask[verts] = True
        # connected components first
        seen = set()
        comps = []
        for s in verts:
            if s in seen:
                continue
            lvl, order = bfs_levels(ptr, idx, mask, s)
            seen.update(order)
            comps.append(order)
        if len(comps) > 1:
            res = []
            for c in sorted(comps, key=len):
                res += rec(np.array(c))
            return res
        s = verts[0]
        for _ in range(3):   # pseudo-peripheral vertex
            lvl, order = bfs_levels(ptr, idx, mask, s)
            s = order[-1]
        lvl, order = bfs_levels(ptr, idx, mask, s)
        depth = max(lvl.values())
        if depth < 2:
            return order_leaf(verts)
        # choose the level that balances the halves best
        cnt = np.bincount(np.array(list(lvl.values())), minlength=depth + 1)
        cum = np.cumsum(cnt)
        best, bl = None, None
        for L in range(1, depth):
            a, b, sepn = cum[L - 1], cum[-1] - cum[L], cnt[L]
            score = sepn * (1.0 + abs(a - b) / max(a + b, 1)) ** 2
            if min(a, b) * 4 >= max(a, b) and (best is None or score < best):
                best, bl = score, L
        if bl is None:
            bl = int(np.searchsorted(cum, cum[-1] // 2))
            bl = min(max(bl, 1), depth - 1)
        sep = [v for v in order if lvl[v] == bl]
        far = set(v for v in order if lvl[v] > bl)
        # thin: a separator vertex without a neighbour on the far side joins the near side
        sep2 = [v for v in sep if any((u in far) for u in idx[ptr[v]:ptr[v + 1]])]
        sepset = set(sep2)
        near = np.array([v for v in order if lvl[v] <= bl and v not in sepset], dtype=np.int64)
        farv = np.array([v for v in order if lvl[v] > bl], dtype=np.int64)
        return rec(near) + rec(farv) + order_leaf(np.array(sep2, dtype=np.int64))

    return np.array(rec(np.arange(n)), dtype=np.int64)


def amd_like(G):
    """Plain minimum-degree with explicit elimination graphs (exact external degree; fine for <= a few thousand)."""
    G = G.tocsr()
    n = G.shape[0]
    adj = [set(G.indices[G.indptr[i]:G.indptr[i + 1]]) - {i} for i in range(n)]
    alive = np.ones(n, bool)
    import heapq
    h = [(len(adj[i]), i) for i in range(n)]
    heapq.heapify(h)
    order = []
    while h:
        d, v = heapq.heappop(h)
        if not alive[v] or d != len(adj[v]):
            continue
        alive[v] = False
        order.append(v)
        nb = adj[v]
        for u in nb:
            adj[u].discard(v)
        for u in nb:
            new = nb - adj[u] - {u}
            if new:
                adj[u] |= new
            heapq.heappush(h, (len(adj[u]), u))
        adj[v] = set()
    return order


def main():
    d = np.load(sys.argv[1])
    n = len(d["indptr"]) - 1
    A = sp.csr_matrix((np.where(d["data"] != 0, 1.0, 0.0), d["indices"], d["indptr"]), shape=(n, n))
    A.eliminate_zeros()   # the NOI-enforced slots are structural zeros for the factorisation (vjx_engine.cu ensure_lu)
    rows, cols = peel(A)
    print(f"n = {n}, core after singleton peeling: {len(rows)} rows x {len(cols)} cols")
    C = A[rows][:, cols]
    # bring the core to a zero-free diagonal (the library pairs rows and columns the same way: threshold pivoting
    # that prefers the paired row)
    from scipy.sparse.csgraph import maximum_bipartite_matching
    m = maximum_bipartite_matching(C.tocsr(), perm_type="column")
    C = C[:, m] if (m >= 0).all() else C
    G = ((C + C.T) != 0).astype(np.float64).tocsr()
    G.setdiag(0)
    G.eliminate_zeros()
    nc = G.shape[0]
    print(f"core graph: {nc} vertices, {G.nnz // 2} edges")
    res = {}
    res["natural"] = np.arange(nc)
    res["rcm"] = np.asarray(reverse_cuthill_mckee(G, symmetric_mode=True))
    res["min-degree (exact)"] = np.array(amd_like(G))
    spd = (G + sp.identity(nc) * (nc + 1.0)).tocsc()
    for spec in ("MMD_AT_PLUS_A", "COLAMD"):
        lu = spl.splu(spd, permc_spec=spec, diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        res[f"superlu {spec}"] = np.argsort(np.asarray(lu.perm_c))
    for leaf in (32, 96, 256):
        res[f"nested dissection (level sets, leaf {leaf})"] = nested_dissection(G, leaf)
    print(f"{'ordering':48s} {'nnz(L+U)':>10s} {'levels':>7s} {'1-row levels':>13s} {'flops':>12s}")
    for k, p in res.items():
        assert sorted(p) == list(range(nc)), k
        nnz, lv, single, fl = symbolic(G, p)
        print(f"{k:48s} {2 * nnz - nc:10d} {lv:7d} {single:13d} {2 * fl:12.3e}")


if __name__ == "__main__":
    main()
