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


def run(n, tol, use_graph=True, max_iter=0, comm_mode=0):
    job = syn.cubic_lattice(n)
    c = mb.Context(device=local, tol=tol, verbose=(rank == 0), use_graph=use_graph, max_iter=max_iter,
                   comm_mode=comm_mode)
    uid = [mb.Context.comm_unique_id() if rank == 0 else None]   # one fresh NCCL id per communicator
    dist.broadcast_object_list(uid, src=0)
    c.comm_init(uid[0], rank, world)
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial,
              job.fixedNodes, job.nodalForces)
    dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
    c.execute(allow_unconverged=True)
    dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    return job, c, dt


for mode in (0, 1):
    job, c, dt = run(12, 1e-12, comm_mode=mode)
    U, F = c.displacements(), c.edge_forces()
    if rank == 0:
        from oracle import oracle as orc
        from helpers import to_oracle, rel_l2
        Uo, Fo = orc.solve(to_oracle(job))
        print(f"[multi] world={world} comm_mode={mode} lattice12: its={c.stats()['iterations']} "
              f"err u={rel_l2(U, Uo):.2e} F={rel_l2(F, Fo):.2e}", flush=True)
    c.close()
for mode, graph in ((0, True), (0, False), (1, True)):
    job, c, dt = run(int(os.environ.get("MBFEA_N", "100")), 1e-8, use_graph=graph, comm_mode=mode)
    st = c.stats()
    if rank == 0:
        print(f"[multi] world={world} comm_mode={mode} graph={graph} lattice: its={st['iterations']} relres={st['rel_residual']:.2e} "
              f"true={st['true_rel_residual']:.2e} pcg={st['ms_pcg']:.1f} ms "
              f"({st['ms_pcg'] / max(1, st['iterations']):.4f} ms/it) asm={st['ms_assemble']:.2f} wall={dt:.3f}s "
              f"ghosts={st['n_ghost_blocks']}", flush=True)
    c.close()
dist.barrier()
dist.destroy_process_group()
