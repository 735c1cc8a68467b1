"""CPU prototype (not product code): multiplicative V-cycle over rigid-body aggregates as the PCG preconditioner.
Smoother: damped block-Jacobi (omega) or Chebyshev(k) on the block-Jacobi-scaled operator; coarsest level exact.
Reports PCG iterations and fine-level SpMVs per solve (the GPU cost driver).

  python tests/tools/proto_vcycle.py [lattice N | shell N] [aggregate edge] [levels] [smoother: j<omega> | c<degree>]
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
smoother = sys.argv[5] if len(sys.argv) > 5 else "j0.6"
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
spmv = [0]


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
Dinv = [block_diag_inv(A[l], A[l].shape[0] // 6) for l in range(levels - 1)]
lu = spla.splu(A[-1].tocsc())
# largest eigenvalue of D^-1 A per level (power iteration) for the Chebyshev interval / Jacobi damping
lam = []
for l in range(levels - 1):
    v = np.random.default_rng(0).standard_normal(A[l].shape[0])
    for _ in range(30):
        w = Dinv[l](A[l] @ v); lmax = np.linalg.norm(w) / np.linalg.norm(v); v = w / np.linalg.norm(w)
    lam.append(1.05 * lmax)


def mul(l, x):
    if l == 0:
        spmv[0] += 1
    return A[l] @ x


def smooth(l, x, b):
    """one application of the smoother to A_l x = b starting from x (x may be None = zero)"""
    if smoother[0] == "j":
        om = float(smoother[1:]) * 2.0 / lam[l] if smoother[1:] != "" else 1.0 / lam[l]
        r = b if x is None else b - mul(l, x)
        d = om * Dinv[l](r)
        return d if x is None else x + d
    k = int(smoother[1:])
    hi, lo = lam[l], lam[l] / 8.0           # damp the upper 7/8 of the spectrum
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma = theta / delta; rho = 1.0 / sigma
    r = b if x is None else b - mul(l, x)
    d = Dinv[l](r) / theta
    xx = d if x is None else x + d
    for _ in range(k - 1):
        r = r - mul(l, d)
        rho_new = 1.0 / (2.0 * sigma - rho)
        d = rho_new * rho * d + 2.0 * rho_new / delta * Dinv[l](r)
        xx = xx + d; rho = rho_new
    return xx


def vcycle(l, b):
    if l == levels - 1:
        return lu.solve(b)
    x = smooth(l, None, b)
    r = b - mul(l, x)
    x = x + P[l] @ vcycle(l + 1, P[l].T @ r)
    return smooth(l, x, b)


def pcg(prec, tol=1e-8, maxit=100000):
    x = np.zeros_like(f); r = f.copy(); z = prec(r); p = z.copy(); rz = r @ z; nf = np.linalg.norm(f)
    for it in range(1, maxit + 1):
        Ap = K @ p; spmv[0] += 1
        a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) <= tol * nf:
            return it, x
        z = prec(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit, x


print(f"{kind} {n}: levels " + " -> ".join(str(a.shape[0]) for a in A) + f" DOF, aggregate edge {agg}, smoother {smoother}, "
      f"lambda_max(D^-1 A) per level {[round(x, 2) for x in lam]}")
spmv[0] = 0; t = time.time(); its0, _ = pcg(Dinv[0]); s0 = spmv[0]
spmv[0] = 0; its1, x = pcg(lambda r: vcycle(0, r)); s1 = spmv[0]
print(f"  block-Jacobi PCG: {its0} iterations = {s0} fine SpMVs;  V-cycle PCG: {its1} iterations = {s1} fine SpMVs "
      f"({s0 / s1:.1f}x fewer), true rel residual {np.linalg.norm(f - K @ x) / np.linalg.norm(f):.2e}")
