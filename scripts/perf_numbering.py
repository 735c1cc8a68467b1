"""How much does the solve depend on the input vertex numbering?  Same lattice, vertices renumbered at random."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn

n = int(os.environ.get("MBFEA_N", "100"))
job = syn.cubic_lattice(n)
V = job.nodes.shape[0]
def shuffled(job, seed):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(V)            # new id of old vertex v
    inv = np.argsort(perm)
    fn, fd, fm = job.nodalForces
    return mb.SimulationJob(job.nodes[inv], perm[job.elements].astype(np.int32), job.elementalNormals,
                            perm[job.fixedNodes].astype(np.int32), (perm[fn], fd, fm), job.beamDimensions, job.beamMaterial), perm
cases = [("lexicographic", job, 0, 2), ("lexicographic", job, 0, 1), ("random numbering", shuffled(job, 1)[0], 0, 2),
         ("random numbering", shuffled(job, 1)[0], 0, 1)]
if os.environ.get("MBFEA_REORDER_AB"):
    cases = [("lexicographic", job, 1, 2)] + [("random numbering", shuffled(job, 1)[0], r, 2) for r in (1, 2, 3)]
    if os.environ.get("MBFEA_SHELL"):
        g = syn.gridshell(int(os.environ["MBFEA_SHELL"])); V = g.nodes.shape[0]
        cases = [("gridshell", g, 1, 2)] + [("gridshell random", shuffled(g, 1)[0], r, 2) for r in (2, 3)]
for label, jb, reorder, storage in cases:
    if True:
        c = mb.Context(device=0, tol=1e-8, storage=storage, max_iter=400, reorder=reorder)
        t = time.time()
        c.set_job(jb.nodes, jb.elements, jb.elementalNormals, jb.beamDimensions, jb.beamMaterial, jb.fixedNodes, jb.nodalForces)
        ts = time.time() - t
        c.assemble()
        st = c.stats()
        ms = c.time_spmv(kernel=1, warmup=2, reps=10)
        c.execute(allow_unconverged=True)
        s2 = c.stats()
        print(f"[numbering] lattice{n} {label:18s} reorder={reorder} storage={storage}: set_job {ts:.2f} s, spmv {ms*1e3:7.1f} us "
              f"({st['bytes_spmv']/ms/1e6:6.0f} GB/s), {s2['ms_pcg']/max(1,s2['iterations'])*1e3:7.1f} us/it", flush=True)
        c.close()
