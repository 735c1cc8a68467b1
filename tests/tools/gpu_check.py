"""First-contact GPU check: every stage against the oracle, verbose, no early exit."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
from oracle import oracle as orc
from helpers import to_oracle, rel_l2, block_rel

def stage(name):
    def deco(fn):
        t = time.time()
        try:
            fn(); print(f"[ OK ] {name} ({time.time()-t:.2f}s)", flush=True)
        except Exception:
            print(f"[FAIL] {name}", flush=True); traceback.print_exc()
    return deco

def check_job(job, label, tol=1e-12, big=False):
    O = to_oracle(job)
    with mb.Context(device=0, tol=tol, verbose=True) as c:
        c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
        props = orc.beam_props(O.dims, O.material)
        print(label, "props equal:", np.array_equal(c.props(), props))
        fm = orc.free_map(O.nodes.shape[0], O.fixed)
        print(label, "dofmap equal:", np.array_equal(c.dofmap(), fm))
        Ke = orc.kelem_all(O, props); Kg = c.kelem()
        print(label, "kelem bit-equal:", np.array_equal(Ke, Kg), "max block rel", block_rel(Kg.reshape(-1,12,12), Ke.reshape(-1,12,12)) if not np.array_equal(Ke,Kg) else 0.0)
        rp, col = orc.bsr_pattern(O, fm)
        c.assemble()
        rp2, col2, vals2, _ = c.bsr()
        print(label, "pattern equal:", np.array_equal(rp, rp2), np.array_equal(col, col2))
        vals = orc.bsr_assemble(O, fm, rp, col, props)
        print(label, "bsr vals bit-equal:", np.array_equal(vals, vals2), "block rel", block_rel(vals2, vals))
        f = orc.load_vector(O.nodes.shape[0], O.forces)
        free_dofs = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
        print(label, "load equal:", np.array_equal(c.load(), f[free_dofs]))
        x = np.random.default_rng(0).standard_normal(free_dofs.size)
        y0 = np.empty_like(x)
        import ctypes as C
        orc.lib().orc_bsr_spmv(C.c_int64(rp.size-1), rp.ctypes.data_as(orc._pi64), np.ascontiguousarray(col,np.int32).ctypes.data_as(orc._pi32), vals.ctypes.data_as(orc._pd), x.ctypes.data_as(orc._pd), y0.ctypes.data_as(orc._pd))
        for k in (1, 2):
            try:
                y = c.spmv(x, kernel=k); print(label, f"spmv kernel {k} rel err", rel_l2(y, y0))
            except Exception as e:
                print(label, f"spmv kernel {k} failed: {e}")
        for k in (1, 2):
            c2 = mb.Context(device=0, tol=tol, spmv_kernel=k, verbose=True)
            c2.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
            t = time.time(); c2.execute(); dt = time.time() - t
            st = c2.stats()
            U = c2.displacements(); F = c2.edge_forces()
            if not big:
                Uo, Fo = orc.solve(O)
                print(label, f"kernel {k}: its={st['iterations']} relres={st['rel_residual']:.2e} true={st['true_rel_residual']:.2e} err u={rel_l2(U,Uo):.2e} F={rel_l2(F,Fo):.2e} wall={dt:.3f}s pcg={st['ms_pcg']:.1f}ms")
                c2.recover_from(Uo); F2 = c2.edge_forces()
                print(label, "recover_from bit-equal:", np.array_equal(F2, Fo), rel_l2(F2, Fo))
            else:
                print(label, f"kernel {k}: its={st['iterations']} relres={st['rel_residual']:.2e} true={st['true_rel_residual']:.2e} wall={dt:.3f}s pcg={st['ms_pcg']:.1f}ms asm={st['ms_assemble']:.2f}ms rec={st['ms_recover']:.2f}ms")
            c2.close()

@stage("two_line_segments")
def _():
    check_job(syn.two_line_segments(), "2seg")

@stage("cantilever 5 beams rotated")
def _():
    cases = list(syn.tester_rotation_cases())
    ax, st, nb, xa, ya = cases[14]
    check_job(syn.cantilever(nb, xa, ya), "cant")

@stage("random mesh")
def _():
    check_job(syn.random_mesh(1234), "rand")

@stage("planar grid 20")
def _():
    check_job(syn.planar_grid(20), "grid20")

@stage("lattice 10")
def _():
    check_job(syn.cubic_lattice(10), "lat10")

@stage("lattice 100 timing")
def _():
    job = syn.cubic_lattice(int(os.environ.get("MBFEA_N", "100")))
    c = mb.Context(device=0, tol=1e-8, verbose=True, max_iter=int(os.environ.get("MBFEA_MAXIT", "2000")))
    t = time.time()
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
    print("set_job wall", time.time() - t, c.stats())
    c.assemble()
    st = c.stats()
    for k in (1, 2):
        ck = mb.Context(device=0, spmv_kernel=k)   # kernel 2 (TMA) implies full-row storage
        ck.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
        ck.assemble()
        sk = ck.stats()
        ms = ck.time_spmv(kernel=k, warmup=3, reps=20)
        print(f"spmv kernel {k}: {ms:.4f} ms  -> {sk['bytes_spmv']/ms/1e6:.1f} GB/s algorithmic ({sk['bytes_spmv']/ms/1e6/6560.3*100:.1f}% of 6560)")
        ck.close()
    ms = c.time_assemble(2, 5)
    print(f"assemble: {ms:.3f} ms -> {st['n_beams']/ms/1e6:.3f} G beams/s; {st['bytes_assemble']/ms/1e6:.1f} GB/s")
    for k in (1, 2):
        c2 = mb.Context(device=0, tol=1e-8, verbose=True, spmv_kernel=k, max_iter=int(os.environ.get("MBFEA_MAXIT", "2000")))
        c2.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
        t = time.time(); rc = c2.execute(allow_unconverged=True); dt = time.time() - t
        s2 = c2.stats()
        print(f"execute kernel {k}: rc={rc} its={s2['iterations']} relres={s2['rel_residual']:.3e} pcg={s2['ms_pcg']:.1f} ms ({s2['ms_pcg']/max(1,s2['iterations']):.4f} ms/it) asm={s2['ms_assemble']:.2f} rec={s2['ms_recover']:.2f} wall={dt:.2f}s launches={s2['kernel_launches']}")
        c2.close()
