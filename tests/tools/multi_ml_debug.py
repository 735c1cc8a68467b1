"""torchrun-able: one multilevel-preconditioned solve on a small lattice with MBFEA_ML_DEBUG diagnostics."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("MBFEA_PARITY_N", "14"))
job = syn.cubic_lattice(n)
for mode in (1, 0):
    c = mb.Context(device=local, tol=1e-10, verbose=(rank == 0), comm_mode=mode, precond=2)
    uid = [mb.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c.comm_init(uid[0], rank, world)
    c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
    try:
        c.execute(allow_unconverged=True)
        st = c.stats()
        if rank == 0:
            print(f"[dbg] mode {mode}: its {st['iterations']} conv {st['converged']} true {st['true_rel_residual']:.2e}", flush=True)
    except Exception as e:
        print(f"[dbg] rank {rank} mode {mode}: {e}", flush=True)
    dist.barrier()
    c.close()
dist.destroy_process_group()
