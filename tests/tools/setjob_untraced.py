import os,sys,time
os.environ.pop("MBFEA_PREP_TRACE",None)
sys.path.insert(0,"."); sys.path.insert(0,"tests")
import meshbeamfea_b200 as mb, bench
wl = sys.argv[1] if len(sys.argv) > 1 else "lattice119"
job=bench.make_job(wl); A=bench.flat_arrays(job,pin=True); V,E=job.nodes.shape[0],job.elements.shape[0]
with mb.Context(device=0,tol=1e-8) as c:
    for i in range(5):
        t=time.perf_counter(); c.set_job_raw(V,E,A["nodes_cm"],A["elems_cm"],A["normals_cm"],A["dims"],A["mat"],A["fixed"],A["fn"],A["fd"],A["fm"]); dt=time.perf_counter()-t
        s=c.stats(); print("untraced set_job %d: %.1f ms prep %.1f rest %.1f ml %.1f" % (i, 1e3*dt, s["ms_host_prep"], s["ms_h2d"], s["ms_ml_symbolic"]), flush=True)
        c.execute(); s=c.stats(); print("   execute: its %d pcg %.1f ms precond %.1f conv %d true %.2e" % (s["iterations"], s["ms_pcg"], s["ms_precond"], s["converged"], s["true_rel_residual"]), flush=True)
