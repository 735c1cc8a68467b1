"""torchrun: set_job phase timings on several ranks (MBFEA_PREP_TRACE=1 on rank 0)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if rank != 0:
    os.environ.pop("MBFEA_PREP_TRACE", None)
import torch, torch.distributed as dist
import meshbeamfea_b200 as mb
import bench
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[1] if len(sys.argv) > 1 else "lattice119"
job = bench.make_job(wl); A = bench.flat_arrays(job, pin=True); V, E = job.nodes.shape[0], job.elements.shape[0]
c = mb.Context(device=local, tol=1e-8)
uid = [mb.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
c.comm_init(uid[0], rank, world)
for i in range(4):
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0: print(f"===== set_job #{i}", file=sys.stderr, flush=True)
    t = time.perf_counter()
    c.set_job_raw(V, E, A["nodes_cm"], A["elems_cm"], A["normals_cm"], A["dims"], A["mat"], A["fixed"], A["fn"], A["fd"], A["fm"])
    dt = time.perf_counter() - t
    s = c.stats()
    print(f"===== rank {rank} set_job #{i}: {1e3*dt:.1f} ms prep {s['ms_host_prep']:.1f} rest {s['ms_h2d']:.1f}", file=sys.stderr, flush=True)
    c.execute()
c.close()
dist.barrier(); dist.destroy_process_group()
