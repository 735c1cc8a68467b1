"""Per-phase device-timer breakdown of the PCG iteration (MBFEA_TRACE=1).
Run directly (1 GPU) or under torchrun (N GPUs)."""
import os
import sys

import numpy as np

os.environ["MBFEA_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import meshbeamfea_b200 as mb  # noqa: E402
from meshbeamfea_b200 import synthetic as syn  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
job = syn.cubic_lattice(int(os.environ.get("MBFEA_N", "100")))
c = mb.Context(device=local, tol=1e-8, max_iter=300, cg_variant=int(os.environ.get("MBFEA_CG", "0")))
if world > 1:
    uid = [mb.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c.comm_init(uid[0], rank, world)
c.set_job(job.nodes, job.elements, job.elementalNormals, job.beamDimensions, job.beamMaterial, job.fixedNodes, job.nodalForces)
c.execute(allow_unconverged=True)
t = c.trace(256).astype(np.int64)
st = c.stats()
# batches end with the residual check: keep iterations whose successor is in the same batch
sel = t[20:250]
names = ["spmv_in", "spmv_go", "upd_in", "upd_go", "pupd_in", "pupd_go", "push_in", "push_out"]
nxt = t[21:251, 0]
variant = int(os.environ.get("MBFEA_CG", "0")) or (2 if world > 1 else 1)
if variant == 2:   # single-reduction CG: SpMV on the residual (stamps 0, 1), then one step kernel (stamps 2, 3)
    segs = {"spmv wait halo": sel[:, 1] - sel[:, 0], "spmv body+tail+launch": sel[:, 2] - sel[:, 1], "step wait dots": sel[:, 3] - sel[:, 2],
            "step body+push+tail+launch": nxt - sel[:, 3], "whole iteration": nxt - sel[:, 0]}
    if world > 1:   # end-of-kernel stamps of the last CTA (slots 4, 5)
        segs.update({"  spmv go -> all CTAs done (ticket won)": sel[:, 6] - sel[:, 1], "  ticket won -> dots published": sel[:, 4] - sel[:, 6],
                     "  spmv go -> last CTA published dots": sel[:, 4] - sel[:, 1], "  dots published -> step kernel starts": sel[:, 2] - sel[:, 4],
                     "  step go -> last CTA published halo tag": sel[:, 5] - sel[:, 3], "  halo tag -> next spmv starts": nxt - sel[:, 5]})
elif world > 1:
    segs = {"spmv wait halo": sel[:, 1] - sel[:, 0], "spmv body+tail": sel[:, 2] - sel[:, 1], "update wait pAp": sel[:, 3] - sel[:, 2],
            "update body+tail": sel[:, 4] - sel[:, 3], "pupdate wait rho": sel[:, 5] - sel[:, 4],
            "pupdate+push+tail": nxt - sel[:, 5], "whole iteration": nxt - sel[:, 0]}
else:
    segs = {"spmv": sel[:, 2] - sel[:, 0], "update": sel[:, 4] - sel[:, 2], "pupdate": nxt - sel[:, 4], "whole iteration": nxt - sel[:, 0]}
msg = [f"[trace] rank {rank}/{world}: ms/it (stats) {st['ms_pcg'] / max(1, st['iterations']):.4f}"]
ok = (nxt - sel[:, 0]) < 3 * np.median(nxt - sel[:, 0])
segs = {k: v[ok] for k, v in segs.items()}
for k, v in segs.items():
    msg.append(f"   {k:44s} median {np.median(v) / 1e3:8.2f} us   mean {np.mean(v) / 1e3:8.2f} us   max {np.max(v) / 1e3:8.2f} us")
print("\n".join(msg), flush=True)
c.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
