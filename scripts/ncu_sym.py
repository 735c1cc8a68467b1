import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
job = syn.cubic_lattice(100)
c = mb.Context(device=0, tol=1e-8, storage=int(os.environ.get("MBFEA_STORAGE", "2")))
c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
c.assemble()
print("spmv ms", c.time_spmv(kernel=1, warmup=2, reps=5))
