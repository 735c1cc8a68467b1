"""torchrun-able multi-GPU check: partitioned solve vs oracle on a small lattice,
then a timed strong-scaling run.  Usage (on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/tools/multi_check.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import meshbeamfea_b200 as mb  # noqa: E402
from meshbeamfea_b200 import synthetic as syn  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def run(n, tol, use_graph=True, max_iter=0, comm_mode=0, precond=1, cg_variant=0):
    job = syn.cubic_lattice(n)
    c = mb.Context(device=local, tol=tol, verbose=(rank == 0), use_graph=use_graph, max_iter=max_iter,
                   comm_mode=comm_mode, precond=precond, cg_variant=cg_variant)
    uid = [mb.Context.comm_unique_id() if rank == 0 else None]   # one fresh NCCL id per communicator
    dist.broadcast_object_list(uid, src=0)
    c.comm_init(uid[0], rank, world)
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial,
              job.fixedNodes, job.nodalForces)
    dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
    c.execute(allow_unconverged=True)
    dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    return job, c, dt


PN = int(os.environ.get("MBFEA_PARITY_N", "14"))
for mode in (0, 1):
    for precond, variant in ((1, 1), (1, 2), (2, 1)):
        job, c, dt = run(PN, 1e-12, comm_mode=mode, precond=precond, cg_variant=variant)
        U, F = c.displacements(), c.edge_forces()
        st = c.stats()
        bad, where = c.compare_plan()   # this rank's device-built plan against the host builders
        flags = [None] * world
        dist.all_gather_object(flags, (int(where), int(abs(bad).sum())))
        if rank == 0:
            print(f"[multi] PLAN world={world}: per rank (built on device: 0 no / 1 pattern / 2 pattern + solver format, mismatching entries "
                  f"vs host builders) = {flags} {'OK' if all(f[1] == 0 for f in flags) else 'FAIL'}", flush=True)
        if rank == 0:
            from oracle import oracle as orc
            from helpers import to_oracle, rel_l2
            Uo, Fo = orc.solve(to_oracle(job))
            print(f"[multi] PARITY world={world} comm_mode={mode} ({'NVLink peer memory' if mode == 0 else 'NCCL'}) precond={st['precond']} "
                  f"cg_variant={variant} lattice{PN}: its={st['iterations']} err vs oracle u={rel_l2(U, Uo):.2e} F={rel_l2(F, Fo):.2e} "
                  f"{'OK' if max(rel_l2(U, Uo), rel_l2(F, Fo)) < 1e-8 else 'FAIL'}", flush=True)
        c.close()
N = int(os.environ.get("MBFEA_N", "100"))
for mode, graph, precond in ((0, True, 1), (0, True, 2), (1, True, 2), (0, False, 2)):
    if N <= 0:
        break
    job, c, dt = run(N, 1e-8, use_graph=graph, comm_mode=mode, precond=precond)
    st = c.stats()
    if rank == 0:
        print(f"[multi] TIMED world={world} comm_mode={mode} graph={graph} precond={st['precond']} lattice{N}: its={st['iterations']} "
              f"relres={st['rel_residual']:.2e} true={st['true_rel_residual']:.2e} pcg={st['ms_pcg']:.1f} ms "
              f"({st['ms_pcg'] / max(1, st['iterations']):.4f} ms/it) asm={st['ms_assemble']:.2f} precond={st['ms_precond']:.2f} "
              f"ml_setup={st['ms_ml_setup']:.2f} symbolic={st['ms_ml_symbolic']:.0f} wall={dt:.3f}s ghosts={st['n_ghost_blocks']}", flush=True)
    c.close()
dist.barrier()
dist.destroy_process_group()
