"""One short multilevel-PCG run for a kernel launch list (ncu --metrics gpu__time_duration.sum): set_job + assemble +
a few iterations.   python tests/tools/ml_profile.py lattice100 [smooth_from] [max_iter] [precond]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "lattice100"
sf = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mi = int(sys.argv[3]) if len(sys.argv) > 3 else 16
pc = int(sys.argv[4]) if len(sys.argv) > 4 else 2
job = bench.make_job(wl)
with mb.Context(device=0, tol=1e-8, precond=pc, ml_smooth_from=sf, max_iter=mi, use_graph=False, verbose=True) as c:
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
    c.execute(allow_unconverged=True)
    s = c.stats()
    print({k: s[k] for k in ("iterations", "ms_pcg", "ms_precond", "ms_ml_setup", "ms_ml_symbolic", "ml_levels")})
