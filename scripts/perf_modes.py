"""Single-GPU A/B of the solver matrix formats and SpMV kernels on one lattice."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meshbeamfea_b200 as mb  # noqa: E402
from meshbeamfea_b200 import synthetic as syn  # noqa: E402

n = int(os.environ.get("MBFEA_N", "100"))
job = syn.cubic_lattice(n)
modes = ((2, 1, "symmetric-half, row-thread"), (1, 1, "full rows, row-thread"), (1, 2, "full rows, TMA"))
if os.environ.get("MBFEA_MODES"):
    modes = tuple(modes[int(k)] for k in os.environ["MBFEA_MODES"].split(","))
for storage, kernel, label in modes:
    c = mb.Context(device=0, tol=1e-8, storage=storage, spmv_kernel=kernel, cg_variant=int(os.environ.get("MBFEA_CG", "0")))
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
    c.assemble()
    st = c.stats()
    ms = c.time_spmv(kernel=kernel, warmup=3, reps=30)
    t = time.time(); c.execute(); dt = time.time() - t
    s2 = c.stats()
    print(f"[perf] cg={os.environ.get('MBFEA_CG', '0')} {os.path.basename(os.environ.get('MBFEA_LIB', 'default')):12s} lattice{n} {label:28s}: spmv {ms * 1e3:7.1f} us ({st['bytes_spmv'] / 1e9:.3f} GB -> {st['bytes_spmv'] / ms / 1e6:6.0f} GB/s) | "
          f"its {s2['iterations']} relres {s2['rel_residual']:.2e} true {s2['true_rel_residual']:.2e} | {s2['ms_pcg'] / max(1, s2['iterations']) * 1e3:6.1f} us/it "
          f"pcg {s2['ms_pcg']:.0f} ms precond {s2['ms_precond']:.2f} ms asm {s2['ms_assemble']:.2f} ms wall {dt:.2f} s", flush=True)
    c.close()
