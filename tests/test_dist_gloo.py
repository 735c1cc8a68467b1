"""world_size-2/3 CPU test (gloo) of the N>1 protocol libmbfea uses on NVLink:
vertex-block partition from mbfea_plan_partition, halo exchange of boundary
x-blocks before each SpMV, and dot products combined by an all-gather of
per-rank partials summed in rank order.  The arithmetic here is numpy on the
oracle's BSR (checker code); what is under test is the HOST logic of the plan
(row split, local column ids, ghost/send maps, neighbour order)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import meshbeamfea_b200 as mb
    from meshbeamfea_b200 import synthetic as syn
    from oracle import oracle as orc
    from helpers import to_oracle
    job = syn.cubic_lattice(n)
    O = to_oracle(job)
    V = O.nodes.shape[0]
    fm = orc.free_map(V, O.fixed)
    rp, col = orc.bsr_pattern(O, fm)
    vals = orc.bsr_assemble(O, fm, rp, col)
    f = orc.load_vector(V, O.forces)
    free_dofs = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
    fr = f[free_dofs]
    p = mb.plan_partition(V, job.elements, job.fixedNodes, rank, world)
    r0, r1 = p["row_begin"], p["row_end"]
    nb, ng = r1 - r0, p["ghost_global"].size
    lvals = vals[rp[r0]:rp[r1]]                 # this rank's block rows (same element-order sums)
    lrp, lcol = p["rowptr"], p["colidx"]
    fl = fr[6 * r0:6 * r1]

    def halo(x):  # x: (nb+ng, 6); owned boundary blocks -> neighbours' ghost slots
        reqs, bufs = [], []
        for q in range(world):
            if q == rank:
                continue
            rows = p["send_rows"][p["send_dest"] == q]
            slots = np.nonzero(p["ghost_owner"] == q)[0]
            if rows.size:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[rows])), q))
            if slots.size:
                t = torch.empty((slots.size, 6), dtype=torch.float64)
                reqs.append(dist.irecv(t, q)); bufs.append((slots, t))
        for r in reqs:
            r.wait()
        for slots, t in bufs:
            x[nb + slots] = t.numpy()

    def spmv(x):
        y = np.zeros((nb, 6))
        for r in range(nb):
            for k in range(lrp[r], lrp[r + 1]):
                y[r] += lvals[k] @ x[lcol[k]]
        return y

    def gsum(*parts):  # all-gather of per-rank partials, summed in rank order
        t = torch.tensor(parts, dtype=torch.float64)
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        return [float(sum(gi[i].item() for gi in g)) for i in range(len(parts))]

    diag = np.array([lvals[lrp[r] + list(lcol[lrp[r]:lrp[r + 1]]).index(r)] for r in range(nb)])
    minv = np.linalg.inv(diag)
    x = np.zeros((nb, 6)); r_ = fl.reshape(nb, 6).copy()
    z = np.einsum("rij,rj->ri", minv, r_)
    pvec = np.zeros((nb + ng, 6)); pvec[:nb] = z
    rz, ff = gsum((r_ * z).sum(), (r_ * r_).sum())
    it = 0
    while it < 5000:
        halo(pvec)
        Ap = spmv(pvec)
        (pAp,) = gsum((pvec[:nb] * Ap).sum())
        alpha = rz / pAp
        x += alpha * pvec[:nb]; r_ -= alpha * Ap
        z = np.einsum("rij,rj->ri", minv, r_)
        rz_new, rr = gsum((r_ * z).sum(), (r_ * r_).sum())
        it += 1
        if rr <= 1e-24 * ff:
            break
        pvec[:nb] = z + (rz_new / rz) * pvec[:nb]
        rz = rz_new
    np.save(os.path.join(out_dir, f"x_{rank}.npy"), x)
    if rank == 0:
        xs, its, rel = orc.pcg_bj(rp, col, vals, fr, tol=1e-12)
        np.save(os.path.join(out_dir, "x_ref.npy"), xs)
        np.save(os.path.join(out_dir, "its.npy"), np.array([it, its]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_pcg_matches_single_rank(built, tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, 5, str(tmp_path)), nprocs=world, join=True)
    x = np.concatenate([np.load(os.path.join(str(tmp_path), f"x_{r}.npy")) for r in range(world)]).ravel()
    ref = np.load(os.path.join(str(tmp_path), "x_ref.npy"))
    it, its = np.load(os.path.join(str(tmp_path), "its.npy"))
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-9
    assert abs(int(it) - int(its)) <= 3  # same Krylov process, partial sums regrouped


def test_mailbox_protocol_simulation(tmp_path):
    """Protocol-level model of the NVLink peer-memory CG exchange (csrc/kernels.cu: Mailbox::ll with slots
    double-buffered by iteration parity, flag = epoch | iteration; neighbour halo tags): 8 ranks as host
    threads with random delays; every value a rank consumes must be the one its peer published for that
    iteration.  (With a single slot per rank the same model deadlocks -- the hang seen on 4 GPUs.)"""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = os.path.join(str(tmp_path), "mailbox_sim")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mailbox_protocol_sim.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", src, "-o", exe, "-lpthread"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert "0 inconsistencies" in out
