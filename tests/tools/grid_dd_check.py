"""Config 2 at full size (planar 100 x 100 cantilever, kappa ~ 1e10) against the double-double reference solve."""
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
Ud, Fd, hist = orc.solve_dd(O)
print("dd reference residual history:", ["%.1e" % h for h in hist])
U2, F2 = orc.solve(O, refine=2)
print(f"plain oracle (refine 2) vs dd: u {rel_l2(U2, Ud):.2e} F {rel_l2(F2, Fd):.2e}")
Uk, Fk = orc.solve(O, "kkt", refine=0)
print(f"raw KKT SuperLU (the reference's call) vs dd: u {rel_l2(Uk, Ud):.2e} F {rel_l2(Fk, Fd):.2e}")
for precond in (1, 2):
    for tol in (1e-8, 1e-10, 1e-12):
        with mb.Context(device=0, tol=tol, precond=precond, ml_smooth_from=1) as c:
            c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
            rc = c.execute(allow_unconverged=True)
            st = c.stats(); U = c.displacements(); F = c.edge_forces()
            print(f"gpu precond={precond} tol={tol:g}: rc {rc} its {st['iterations']} restarts {st['restarts']} conv {st['converged']} rec {st['rel_residual']:.2e} "
                  f"true {st['true_rel_residual']:.2e} | vs dd: u {rel_l2(U, Ud):.2e} F {rel_l2(F, Fd):.2e}", flush=True)
