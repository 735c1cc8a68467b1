"""GPU check of the multilevel preconditioner (options.precond = 2) against the block-Jacobi solve: iteration counts,
solve / setup times and agreement of the displacements, over hierarchy parameters.

  python tests/tools/ml_check.py small|big|shell [smooth_from ...]
"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
from helpers import rel_l2


def solve(job, **kw):
    with mb.Context(device=0, tol=1e-8, **kw) as c:
        c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
        c.execute(allow_unconverged=True)
        c.execute(allow_unconverged=True)   # second run: graph cached, allocations warm
        return c.stats(), c.displacements()


def run(job, label, variants):
    try:
        s0, U0 = solve(job, precond=1)
        print(f"{label}: block-Jacobi its={s0['iterations']} pcg={s0['ms_pcg']:.1f}ms precond={s0['ms_precond']:.2f}ms true={s0['true_rel_residual']:.2e}", flush=True)
        for kw in variants:
            try:
                s1, U1 = solve(job, precond=2, **kw)
                print(f"   ML {kw}: levels={s1['ml_levels']} its={s1['iterations']} pcg={s1['ms_pcg']:.1f}ms ({s1['ms_pcg']/max(1,s1['iterations']):.4f} ms/it) "
                      f"setup={s1['ms_ml_setup']:.2f}ms symbolic={s1['ms_ml_symbolic']:.1f}ms set_job prep={s1['ms_host_prep']:.0f}ms true={s1['true_rel_residual']:.2e} conv={s1['converged']} "
                      f"restarts={s1['restarts']} | diff u={rel_l2(U1, U0):.2e}", flush=True)
            except Exception:
                print(f"   [FAIL] ML {kw}", flush=True); traceback.print_exc()
    except Exception:
        print(f"[FAIL] {label}", flush=True); traceback.print_exc()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    V = [dict(ml_smooth_from=99), dict(ml_smooth_from=2), dict(ml_smooth_from=1)]
    if which == "small":
        run(syn.cubic_lattice(12), "lattice12", V)
        run(syn.cubic_lattice(40), "lattice40", V + [dict(ml_smooth_from=1, ml_cell=3.0)])
        run(syn.gridshell(240), "gridshell240", V)
        run(syn.planar_grid(100), "grid100", V)
    elif which == "big":
        run(syn.cubic_lattice(100), "lattice100", V + [dict(ml_smooth_from=2, ml_coarsest_rows=32), dict(ml_smooth_from=1, ml_ratio=3.0)])
        run(syn.cubic_lattice(119), "lattice119", V[1:])
    elif which == "shell":
        run(syn.gridshell(480), "gridshell480", V)
        run(syn.gridshell(1826), "gridshell1826", V[1:])
