"""Hierarchy-parameter sweep of the multilevel preconditioner on one workload:
   python tests/tools/ml_params.py gridshell1826 "cell=2,sf=1" "cell=3,sf=1" ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb
import bench
job = bench.make_job(sys.argv[1])
for spec in sys.argv[2:]:
    kw = dict(kv.split("=") for kv in spec.split(","))
    opts = dict(precond=2, ml_cell=float(kw.get("cell", 0)), ml_smooth_from=int(kw.get("sf", 0)), ml_ratio=float(kw.get("ratio", 0)),
                ml_coarsest_rows=int(kw.get("coarsest", 0)))
    try:
        with mb.Context(device=0, tol=1e-8, **opts) as c:
            c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
            c.execute(); c.execute()
            s = c.stats()
            print(f"{sys.argv[1]} {spec}: levels={s['ml_levels']} its={s['iterations']} pcg={s['ms_pcg']:.1f} ms ({s['ms_pcg']/s['iterations']:.4f} ms/it) "
                  f"setup={s['ms_ml_setup']:.1f} ms execute={s['ms_execute_total']:.1f} ms conv={s['converged']} true={s['true_rel_residual']:.2e}", flush=True)
    except Exception as e:
        print(f"{sys.argv[1]} {spec}: FAILED {e}", flush=True)
