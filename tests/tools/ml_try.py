"""GPU first contact of the multilevel preconditioner: iteration counts and agreement with the plain solve."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
from helpers import rel_l2


def run(job, label, cell=2.0, ratio=2.0, coarsest=128):
    try:
        with mb.Context(device=0, tol=1e-8, cg_variant=1) as c:
            c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
            c.execute(allow_unconverged=True)
            s0 = c.stats(); U0 = c.displacements()
            c.assemble()
            t = time.time()
            c.solve_multilevel(cell, ratio, coarsest)
            dt = time.time() - t
            s1 = c.stats()
            c.recover()
            U1 = c.displacements()
            print(f"{label}: plain its={s0['iterations']} pcg={s0['ms_pcg']:.1f}ms | ML(cell={cell},ratio={ratio}) its={s1['iterations']} "
                  f"pcg={s1['ms_pcg']:.1f}ms wall={dt*1e3:.0f}ms relres={s1['rel_residual']:.2e} | diff u={rel_l2(U1, U0):.2e}", flush=True)
    except Exception:
        print(f"[FAIL] {label}", flush=True); traceback.print_exc()


if __name__ == "__main__":
    which = sys.argv[1:] or ["small"]
    if "small" in which:
        run(syn.cubic_lattice(12), "lattice12")
        run(syn.cubic_lattice(20), "lattice20")
        run(syn.cubic_lattice(40), "lattice40")
        run(syn.gridshell(120), "gridshell120")
        run(syn.planar_grid(100), "grid100")
    if "big" in which:
        run(syn.cubic_lattice(100), "lattice100")
        run(syn.cubic_lattice(100), "lattice100", 2.0, 3.0)
        run(syn.gridshell(480), "gridshell480")
