"""set_job phase timings (MBFEA_PREP_TRACE=1) on a resubmitted mesh: python tests/tools/setjob_trace.py lattice119"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MBFEA_PREP_TRACE", "1")
import meshbeamfea_b200 as mb
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "lattice119"
job = bench.make_job(wl)
A = bench.flat_arrays(job, pin=True)
V, E = job.nodes.shape[0], job.elements.shape[0]
with mb.Context(device=0, tol=1e-8) as c:
    for i in range(int(os.environ.get("MBFEA_SETJOBS", "3"))):
        print(f"===== set_job #{i}", file=sys.stderr, flush=True)
        t = time.perf_counter()
        c.set_job_raw(V, E, A["nodes_cm"], A["elems_cm"], A["normals_cm"], A["dims"], A["mat"], A["fixed"], A["fn"], A["fd"], A["fm"])
        print(f"===== set_job #{i}: {1e3 * (time.perf_counter() - t):.1f} ms  stats prep {c.stats()['ms_host_prep']:.1f} h2d {c.stats()['ms_h2d']:.1f} ml_symbolic {c.stats()['ms_ml_symbolic']:.1f}", file=sys.stderr, flush=True)
        c.execute()
