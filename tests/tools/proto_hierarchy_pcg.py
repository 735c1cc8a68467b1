"""CPU prototype on the PRODUCT's aggregate hierarchy (mb.host_hierarchy): additive multilevel PCG
  z = sum over levels  P_0 .. P_{l-1} D_l^-1 P_{l-1}^T .. P_0^T r   (coarsest level solved exactly)
with the level matrices assembled from the hierarchy's Galerkin contributor lists -- i.e. exactly the data the
device kernels of the next round will consume.  Reports PCG iterations against block-Jacobi.

  python tests/tools/proto_hierarchy_pcg.py [lattice N | shell N] [cell0 in mean beam lengths] [ratio] [coarsest_rows]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb  # noqa: E402
from meshbeamfea_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from helpers import to_oracle  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "lattice"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cell_mult = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
ratio = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
coarsest = int(sys.argv[5]) if len(sys.argv) > 5 else 128
job = syn.cubic_lattice(n) if kind == "lattice" else syn.gridshell(n)
O = to_oracle(job)
V = O.nodes.shape[0]
fm = orc.free_map(V, O.fixed)
rp, col = orc.bsr_pattern(O, fm)
vals = orc.bsr_assemble(O, fm, rp, col).reshape(-1, 6, 6)
nb = rp.size - 1
f = orc.load_vector(V, O.forces).reshape(V, 6)[fm >= 0].ravel()
hmean = np.mean(np.linalg.norm(O.nodes[O.elements[:, 0]] - O.nodes[O.elements[:, 1]], axis=1))
H = mb.host_hierarchy(job.nodes, job.elements, job.fixedNodes, cell0=cell_mult * hmean, ratio=ratio, max_levels=12, coarsest_rows=coarsest)


def rigid(off):
    B = np.tile(np.eye(6), (off.shape[0], 1, 1))
    x, y, z = off[:, 0], off[:, 1], off[:, 2]
    B[:, 0, 4], B[:, 0, 5], B[:, 1, 3], B[:, 1, 5], B[:, 2, 3], B[:, 2, 4] = z, -y, -z, x, y, -x
    return B


levels = [dict(rp=rp, col=col, blocks=vals)]
P = []
for L in H:
    fine = levels[-1]
    B = rigid(L["off"])
    frow = np.repeat(np.arange(L["n_fine"]), np.diff(fine["rp"]))
    gs = L["gal_src"]; seg = np.repeat(np.arange(L["gal_ptr"].size - 1), np.diff(L["gal_ptr"]))
    contrib = np.einsum("kji,kjl,klm->kim", B[frow[gs]], fine["blocks"][gs], B[fine["col"][gs]])
    blocks = np.zeros((L["gal_ptr"].size - 1, 6, 6)); np.add.at(blocks, seg, contrib)
    levels.append(dict(rp=L["rowptr"], col=L["colidx"], blocks=blocks))
    P.append(sp.bsr_matrix((B, L["agg"], np.arange(L["n_fine"] + 1)), shape=(6 * L["n_fine"], 6 * L["n_coarse"])).tocsr())


def diag_inv(lv):
    m = lv["rp"].size - 1
    D = np.zeros((m, 6, 6))
    for r in range(m):
        k = lv["rp"][r] + np.searchsorted(lv["col"][lv["rp"][r]:lv["rp"][r + 1]], r)
        D[r] = lv["blocks"][k]
    Di = np.linalg.inv(D)
    return lambda r: np.einsum("rij,rj->ri", Di, r.reshape(m, 6)).ravel()


K = sp.bsr_matrix((vals, col, rp), shape=(6 * nb, 6 * nb)).tocsr()
sm = [diag_inv(lv) for lv in levels[:-1]]
last = levels[-1]; m = last["rp"].size - 1
Alast = sp.bsr_matrix((last["blocks"], last["col"], last["rp"]), shape=(6 * m, 6 * m)).toarray()
Ainv = np.linalg.inv(Alast)


def prec(r):
    rs = [r]
    for Pl in P:
        rs.append(Pl.T @ rs[-1])
    z = Ainv @ rs[-1]
    for l in range(len(P) - 1, -1, -1):
        z = sm[l](rs[l]) + P[l] @ z
    return z


def pcg(prec, tol=1e-8, maxit=100000):
    x = np.zeros_like(f); r = f.copy(); z = prec(r); p = z.copy(); rz = r @ z; nf = np.linalg.norm(f)
    for it in range(1, maxit + 1):
        Ap = K @ p; a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) <= tol * nf:
            return it, x
        z = prec(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit, x


its0, _ = pcg(sm[0] if sm else (lambda r: Ainv @ r))
its1, x = pcg(prec)
print(f"{kind} {n}: rows per level {[nb] + [L['n_coarse'] for L in H]} (cell0 = {cell_mult} beam lengths, ratio {ratio}); "
      f"block-Jacobi {its0} iterations, additive multilevel {its1} iterations ({its0 / its1:.2f}x), "
      f"true rel residual {np.linalg.norm(f - K @ x) / np.linalg.norm(f):.2e}")
