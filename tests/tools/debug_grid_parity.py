import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
from oracle import oracle as orc
from helpers import to_oracle, rel_l2
n = int(os.environ.get("MBFEA_N", "100"))
job = syn.planar_grid(n)
O = to_oracle(job)
V = O.nodes.shape[0]
K = orc.assemble_full(O)
f = orc.load_vector(V, O.forces)
fm = orc.free_map(V, O.fixed)
free = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
Kr = K[free][:, free].tocsr()
fr = f[free]
def resid(U):
    u = U.ravel()[free]
    r = fr.astype(np.longdouble) - (Kr.astype(np.float64) @ u).astype(np.longdouble)
    # long-double accumulate: do it blockwise to limit memory
    Kl = Kr.tocoo()
    acc = np.zeros(fr.size, np.longdouble)
    np.add.at(acc, Kl.row, Kl.data.astype(np.longdouble) * u[Kl.col].astype(np.longdouble))
    r = fr.astype(np.longdouble) - acc
    return float(np.sqrt((r * r).sum()) / np.linalg.norm(fr))
for refine in (0, 2, 6):
    Uo, Fo = orc.solve(O, "eliminate", refine=refine)
    print(f"oracle eliminate refine={refine}: true rel residual {resid(Uo):.3e}")
    if refine == 2: Uref = Uo
Uk, _ = orc.solve(O, "kkt", refine=2)
print(f"oracle kkt refine=2: residual {resid(Uk):.3e}  vs eliminate(2): {rel_l2(Uk, Uref):.3e}")
U6, _ = orc.solve(O, "eliminate", refine=6)
print("eliminate refine 6 vs 2:", rel_l2(U6, Uref))
for storage in (2, 1):
    for tol in (1e-8, 1e-12):
        c = mb.Context(device=0, tol=tol, storage=storage)
        c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
        c.execute()
        st = c.stats(); U = c.displacements()
        print(f"gpu storage={storage} tol={tol:g}: its {st['iterations']} restarts {st['restarts']} conv {st['converged']} rec {st['rel_residual']:.2e} "
              f"true(gpu) {st['true_rel_residual']:.2e} true(longdouble) {resid(U):.3e}  vs oracle(2) {rel_l2(U, Uref):.3e} vs oracle(6) {rel_l2(U, U6):.3e}")
        c.close()
