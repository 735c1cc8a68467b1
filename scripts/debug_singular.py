import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
job = syn.two_line_segments(fixed=())
c = mb.Context(device=0, max_iter=500, verbose=True)
c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
try:
    rc = c.execute(allow_unconverged=True)
    print("rc", rc, c.stats())
    print(c.displacements())
except Exception as e:
    print("raised", type(e).__name__, e)
