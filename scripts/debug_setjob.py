import os, sys, time
os.environ["MBFEA_PREP_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
job = syn.cubic_lattice(100)
c = mb.Context(device=0, tol=1e-8, max_iter=50)
for rep in range(3):
    t = time.time()
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
    print(f"=== set_job #{rep}: {time.time() - t:.3f} s", c.stats()["ms_host_prep"], c.stats()["ms_h2d"], flush=True)
    c.execute(allow_unconverged=True)
