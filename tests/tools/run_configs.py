"""Runs the BASELINE.json configs on one GPU and prints one JSON line per config
(parity where an oracle answer exists, throughput and solver statistics otherwise).

  python tests/tools/run_configs.py [1 2 3 3b 4 5]      (default: all)
    1   Models/TwoLineSegments.ply-equivalent fixture, three scenarios (golden vectors)
    2   planar 100x100 grid, cantilever load            (vs the oracle's direct solve)
    3   cubic lattice 100^3
    3b  cubic lattice 119^3 (10.1 M DOF: the time-to-solve metric)
    4   triangulated gridshell, ~10 M beams (n = 1826)
    5   4096 random ~500-beam meshes solved concurrently as one block-diagonal job
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb  # noqa: E402
from meshbeamfea_b200 import synthetic as syn  # noqa: E402

which = sys.argv[1:] or ["1", "2", "3", "3b", "4", "5"]


def run(job, label, tol=1e-8, oracle=False, extra=None):
    V, E = job.nodes.shape[0], job.elements.shape[0]
    c = mb.Context(device=0, tol=tol)
    t0 = time.perf_counter()
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
    t1 = time.perf_counter()
    c.execute()            # warm (graph build, first-touch)
    t2 = time.perf_counter()
    c.execute()
    t3 = time.perf_counter()
    st = c.stats()
    ms_spmv = c.time_spmv(warmup=2, reps=20) if E > 100000 else None
    ms_asm = c.time_assemble(warmup=1, reps=5)
    out = {"config": label, "vertices": V, "beams": E, "free_dof": 6 * st["n_block_rows_global"], "tol": tol,
           "iterations": st["iterations"], "rel_residual": st["rel_residual"], "true_rel_residual": st["true_rel_residual"],
           "set_job_s": t1 - t0, "execute_s": t3 - t2, "time_to_solve_s": st["ms_pcg"] / 1e3,
           "ms_per_iteration": st["ms_pcg"] / max(1, st["iterations"]), "beams_per_s": E / (t3 - t2),
           "assembled_beams_per_s": E / (ms_asm * 1e-3), "phase_ms": {k: st[k] for k in ("ms_assemble", "ms_precond", "ms_pcg", "ms_recover")},
           "preconditioner": "multilevel" if st["precond"] == 2 else "block-Jacobi", "ml_levels": st["ml_levels"],
           "ms_ml_setup": st["ms_ml_setup"], "ms_ml_symbolic": st["ms_ml_symbolic"], "converged": st["converged"], "restarts": st["restarts"]}
    if ms_spmv:
        out["spmv_ms"] = ms_spmv
        out["spmv_gbs"] = st["bytes_spmv"] / ms_spmv / 1e6
    if oracle:
        from oracle import oracle as orc
        from helpers import to_oracle, rel_l2
        c2 = mb.Context(device=0, tol=1e-12)
        c2.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
        c2.execute()
        U, F, hist = orc.solve_dd(to_oracle(job))   # SuperLU refined with double-double residuals (exact to ~1e-24)
        out["reference"] = "oracle.solve_dd, final residual %.1e" % hist[-1]
        out["parity_tol1e-12"] = {"iterations": c2.stats()["iterations"], "rel_l2_u": rel_l2(c2.displacements(), U),
                                  "rel_l2_forces": rel_l2(c2.edge_forces(), F)}
        out["parity_at_bench_tol"] = {"rel_l2_u": rel_l2(c.displacements(), U), "rel_l2_forces": rel_l2(c.edge_forces(), F)}
        c2.close()
    if extra:
        out.update(extra(c))
    print(json.dumps(out), flush=True)
    c.close()


if "1" in which:
    sim = mb.BeamSimulator(device=0, tol=1e-13)
    sim.setSimulation(syn.two_line_segments())
    r = sim.executeSimulation()
    a = {"Mz": list(r.edgeForces[5]), "u_y_tip": r.nodalDisplacements[2, 1]}
    sim.setSimulation(syn.two_line_segments(fixed=(1, 2), forces=((0, 0, 1000.0),)))
    r = sim.executeSimulation()
    b = {"u_x0": r.nodalDisplacements[0, 0], "N": list(r.edgeForces[0])}
    try:
        sim.setSimulation(syn.two_line_segments(fixed=(1, 2, 3), forces=((0, 0, 1000.0),)))
        cerr = "no error"
    except RuntimeError as e:
        cerr = str(e)
    print(json.dumps({"config": "1 TwoLineSegments", "1a_screenshot": a, "1b_axial": b, "1c_config_json_as_shipped": cerr}), flush=True)
if "2" in which:
    run(syn.planar_grid(100), "2 planar grid 100x100", oracle=True)
if "3" in which:
    run(syn.cubic_lattice(100), "3 cubic lattice 100^3")
if "3b" in which:
    run(syn.cubic_lattice(119), "3b cubic lattice 119^3 (10.1 M DOF)")
if "4" in which:
    n = int(os.environ.get("MBFEA_SHELL_N", "1826"))
    run(syn.gridshell(n), f"4 triangulated gridshell n={n}")
if "5" in which:
    m = int(os.environ.get("MBFEA_BATCH", "4096"))
    t = time.perf_counter()
    jobs = [syn.random_mesh(1234 + k) for k in range(m)]
    merged = syn.merge_jobs(jobs)
    gen = time.perf_counter() - t
    run(merged, f"5 batch of {m} random meshes (block-diagonal job)",
        extra=lambda c: {"meshes_per_s": m / (c.stats()["ms_execute_total"] / 1e3), "generation_s": gen})
