"""CPU prototype (not product code): additive MULTILEVEL preconditioner that needs no coarse SpMV --
  z = D0^-1 r + P1 D1^-1 P1^T r + P1 P2 A2^-1 P2^T P1^T r
D_l = 6x6 diagonal blocks of A_l = P_l^T A_{l-1} P_l (rigid-body modes of vertex aggregates), A2 solved exactly
(small and dense on the GPU).  Compares PCG iteration counts with block-Jacobi alone.

  python tests/tools/proto_multilevel.py [lattice N | shell N] [aggregate edge] [levels]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from meshbeamfea_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from helpers import to_oracle  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "lattice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
agg = int(sys.argv[3]) if len(sys.argv) > 3 else 4
levels = int(sys.argv[4]) if len(sys.argv) > 4 else 3
job = syn.cubic_lattice(n) if kind == "lattice" else syn.gridshell(n)
O = to_oracle(job)
V = O.nodes.shape[0]
fm = orc.free_map(V, O.fixed)
rowptr, col = orc.bsr_pattern(O, fm)
vals = orc.bsr_assemble(O, fm, rowptr, col)
nb = rowptr.size - 1
K = sp.bsr_matrix((vals.reshape(-1, 6, 6), col, rowptr), shape=(6 * nb, 6 * nb)).tocsr()
f = orc.load_vector(V, O.forces).reshape(V, 6)[fm >= 0].ravel()
X = O.nodes[fm >= 0]
h = np.median(np.linalg.norm(O.nodes[O.elements[:, 0]] - O.nodes[O.elements[:, 1]], axis=1))


def block_diag_inv(A, m):
    A = A.tobsr(blocksize=(6, 6)); A.sort_indices()
    D = np.zeros((m, 6, 6))
    for r in range(m):
        k = A.indptr[r] + np.searchsorted(A.indices[A.indptr[r]:A.indptr[r + 1]], r)
        D[r] = A.data[k]
    Di = np.linalg.inv(D)
    return lambda r: np.einsum("rij,rj->ri", Di, r.reshape(m, 6)).ravel()


def rigid_prolongation(pts, edge):
    cell = np.floor((pts - pts.min(0)) / edge + 1e-9).astype(np.int64)
    _, aid = np.unique(cell, axis=0, return_inverse=True)
    aid = aid.ravel(); na = aid.max() + 1
    cent = np.zeros((na, 3)); np.add.at(cent, aid, pts); cent /= np.bincount(aid)[:, None]
    rel = pts - cent[aid]
    m = pts.shape[0]
    rows, cols, data = [], [], []
    for v in range(m):
        x, y, z = rel[v]
        B = np.eye(6); B[:3, 3:] = np.array([[0, z, -y], [-z, 0, x], [y, -x, 0]])
        for i in range(6):
            for j in range(6):
                if B[i, j] != 0.0:
                    rows.append(6 * v + i); cols.append(6 * aid[v] + j); data.append(B[i, j])
    return sp.csr_matrix((data, (rows, cols)), shape=(6 * m, 6 * na)), cent


A = [K]; P = []; pts = X; edge = h * agg
for l in range(1, levels):
    Pl, pts = rigid_prolongation(pts, edge)
    P.append(Pl); A.append((Pl.T @ A[-1] @ Pl).tocsr()); edge *= agg
sm = [block_diag_inv(A[l], A[l].shape[0] // 6) for l in range(levels - 1)]
lu = spla.splu(A[-1].tocsc())


def prec(r):
    rs = [r]
    for Pl in P:
        rs.append(Pl.T @ rs[-1])
    z = lu.solve(rs[-1])
    for l in range(levels - 2, -1, -1):
        z = sm[l](rs[l]) + P[l] @ z
    return z


def pcg(prec, tol=1e-8, maxit=200000):
    x = np.zeros_like(f); r = f.copy(); z = prec(r); p = z.copy(); rz = r @ z; nf = np.linalg.norm(f)
    for it in range(1, maxit + 1):
        Ap = K @ p; a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) <= tol * nf:
            return it, x
        z = prec(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit, x


print(f"{kind} {n}: {6 * nb} DOF; levels " + " -> ".join(str(a.shape[0]) for a in A) + f" DOF (aggregate edge {agg})")
t = time.time(); its0, _ = pcg(sm[0]); t0 = time.time() - t
t = time.time(); its1, x = pcg(prec); t1 = time.time() - t
print(f"  block-Jacobi: {its0} iterations ({t0:.1f} s);  additive {levels}-level: {its1} iterations ({t1:.1f} s), "
      f"true rel residual {np.linalg.norm(f - K @ x) / np.linalg.norm(f):.2e}")
