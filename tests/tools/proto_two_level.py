"""CPU prototype for SURVEY 8(f) rank 4 (not product code, nothing here runs on the GPU):
how many PCG iterations would a two-level preconditioner save on the BASELINE meshes?

  M1 = block-Jacobi (what libmbfea runs today)
  M2 = block-Jacobi + rigid-body coarse space over vertex aggregates, additive:
       z = D^-1 r + P (P^T K P)^-1 P^T r,  P = 6 rigid-body modes per aggregate
  M3 = the same, multiplicative and symmetric: smoother - coarse correction - smoother

  python tests/tools/proto_two_level.py [lattice N | shell N] [aggregate edge]
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

# block-Jacobi
D = np.zeros((nb, 6, 6))
for r in range(nb):
    k = rowptr[r] + np.searchsorted(col[rowptr[r]:rowptr[r + 1]], r)
    D[r] = vals.reshape(-1, 6, 6)[k]
Dinv = np.linalg.inv(D)
bj = lambda r: np.einsum("rij,rj->ri", Dinv, r.reshape(nb, 6)).ravel()

# aggregates: bricks of `agg` vertices per axis of the bounding box grid (mean spacing from the mesh)
h = np.median(np.linalg.norm(O.nodes[O.elements[:, 0]] - O.nodes[O.elements[:, 1]], axis=1))
if agg > 0:
    cell = np.floor((X - X.min(0)) / (h * agg) + 1e-9).astype(np.int64)
else:
    # agg = -N: a fixed budget of about N aggregates, bricks of the bounding box with near-cubic cells
    ext = np.maximum(X.max(0) - X.min(0), 1e-12)
    live = ext > 1e-3 * ext.max()
    side = (np.prod(ext[live]) / (-agg)) ** (1.0 / live.sum())
    counts = np.where(live, np.maximum(1, np.round(ext / side)), 1).astype(np.int64)
    cell = np.minimum((np.floor((X - X.min(0)) / ext * counts)).astype(np.int64), counts - 1)
    print("aggregate grid", counts)
_, aid = np.unique(cell, axis=0, return_inverse=True)
aid = aid.ravel(); na = aid.max() + 1
cent = np.zeros((na, 3)); np.add.at(cent, aid, X); cent /= np.bincount(aid)[:, None]
rel = X - cent[aid]
rows, cols, data = [], [], []
for v in range(nb):
    a = aid[v]; x, y, z = rel[v]
    # translations: u = e_k; rotations about the centroid: u = e_k x r, theta = e_k
    B = np.zeros((6, 6)); B[:3, :3] = np.eye(3); B[3:, 3:] = np.eye(3)
    B[:3, 3:] = np.array([[0, z, -y], [-z, 0, x], [y, -x, 0]])
    for i in range(6):
        for j in range(6):
            if B[i, j] != 0.0:
                rows.append(6 * v + i); cols.append(6 * a + j); data.append(B[i, j])
P = sp.csr_matrix((data, (rows, cols)), shape=(6 * nb, 6 * na))
Ac = (P.T @ K @ P).tocsc()
t = time.time(); lu = spla.splu(Ac); t_lu = time.time() - t
coarse = lambda r: P @ lu.solve(P.T @ r)


def pcg(prec, tol=1e-8, maxit=200000):
    x = np.zeros_like(f); r = f.copy(); z = prec(r); p = z.copy(); rz = r @ z; nf = np.linalg.norm(f)
    for it in range(1, maxit + 1):
        Ap = K @ p; a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) <= tol * nf:
            return it, x
        z = prec(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit, x


def mult(r):   # symmetric multiplicative: smoother, coarse correction, smoother
    z = bj(r); z = z + coarse(r - K @ z); return z + bj(r - K @ z)


print(f"{kind} {n}: {nb} free vertices ({6 * nb} DOF), {na} aggregates of edge {agg} -> coarse system {6 * na} DOF "
      f"(nnz {Ac.nnz}, LU {t_lu:.2f} s)")
for name, prec, spmv_per_it in (("block-Jacobi", bj, 1), ("additive two-level", lambda r: bj(r) + coarse(r), 1),
                                ("multiplicative two-level", mult, 3)):
    t = time.time(); its, x = pcg(prec); dt = time.time() - t
    print(f"  {name:26s}: {its:6d} iterations ({its * spmv_per_it} fine SpMVs), true rel residual "
          f"{np.linalg.norm(f - K @ x) / np.linalg.norm(f):.2e}, {dt:.1f} s")
