"""Synthetic edge meshes for the BASELINE.json configs and the reference's own
tester case.  Host-side numpy only; everything returns a ``SimulationJob``
(per-beam sections as (E,2) float32 arrays).  Generators are deterministic.

config 1  two_line_segments()        [REF Models/TwoLineSegments.ply]
tester    cantilever(...)            [REF src/beamsimulatortester.cpp:110-163, .hpp:7-18]
config 2  planar_grid(100)           20k beams, 60k DOF
config 3  cubic_lattice(100)         2.97M beams, 6M DOF  (119 -> 10.1M DOF)
config 4  gridshell(n)               triangulated shallow shell, ~3 n^2 beams
config 5  random_mesh(seed)          ~500-beam random connected edge mesh
"""
from __future__ import annotations

import numpy as np

from .simulator import SimulationJob, NodalForce


def _sections(E, b=0.01, h=0.01, nu=0.3, E_gpa=210.0):
    dims = np.empty((E, 2), np.float32); dims[:, 0] = b; dims[:, 1] = h
    mat = np.empty((E, 2), np.float32); mat[:, 0] = nu; mat[:, 1] = E_gpa
    return dims, mat


def two_line_segments(fixed=(0,), forces=((2, 1, 100.0),)) -> SimulationJob:
    """The shipped fixture: 3 collinear vertices, 2 beams, normal (0,1,0),
    b=h=0.01, nu=0.3, E=800 GPa.  Default scenario = the README screenshot."""
    nodes = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0]], np.float64)
    elements = np.array([[0, 1], [1, 2]], np.int32)
    normals = np.array([[0, 1, 0], [0, 1, 0]], np.float64)
    dims, mat = _sections(2, 0.01, 0.01, 0.3, 800.0)
    return SimulationJob(nodes, elements, normals, np.array(fixed, np.int32),
                         [NodalForce(*f) for f in forces], dims, mat)


def angle_axis(angle: float, axis) -> np.ndarray:
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    c, s = np.cos(angle), np.sin(angle)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return c * np.eye(3) + s * K + (1 - c) * np.outer(a, a)


def cantilever(n_beams=1, x_axis=(1, 0, 0), y_axis=(0, 1, 0), force=1000.0, b=0.1, h=0.2,
               E_gpa=210.0, nu=0.3) -> SimulationJob:
    """BeamSimulatorTester::CantileverBeam: unit length along XAxis split into
    n_beams, fixed at node 0, tip load force*YAxis split into 3 DOF forces."""
    x = np.asarray(x_axis, np.float64); y = np.asarray(y_axis, np.float64)
    step = x / np.linalg.norm(x) / n_beams
    nodes = np.arange(n_beams + 1)[:, None] * step[None, :]
    elements = np.stack([np.arange(n_beams), np.arange(1, n_beams + 1)], 1).astype(np.int32)
    normals = np.tile(y, (n_beams, 1))
    dims, mat = _sections(n_beams, b, h, nu, E_gpa)
    f32 = np.float32(force)
    forces = [NodalForce(n_beams, d, float(y[d] * f32)) for d in range(3)]
    return SimulationJob(nodes, elements, normals, np.array([0], np.int32), forces, dims, mat)


def tester_rotation_cases():
    """The 45 cases of BeamSimulatorTester::launchRotationTest
    [REF src/beamsimulatortester.cpp:8-41]: axes (0,0,1),(0,1,0),(1,0,0) x
    quarter turns 1..3 (angle is float(pi/2)) x 1..5 beams."""
    ang = float(np.float32(np.pi / 2))
    for axis in ((0, 0, 1), (0, 1, 0), (1, 0, 0)):
        for step in (1, 2, 3):
            R = angle_axis(step * ang, axis)
            xa, ya = R @ np.array([1.0, 0, 0]), R @ np.array([0, 1.0, 0])
            for nbeams in range(1, 6):
                yield axis, step, nbeams, xa, ya


def planar_grid(n=100, ell=0.1, b=0.01, h=0.01) -> SimulationJob:
    """Config 2: n x n grid in z=0, vertex id i*n+j, per vertex the +i edge then
    the +j edge; normals (0,0,1); row i=0 fixed; unit out-of-plane load (dof 2)
    on every vertex of row i=n-1."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    nodes = np.stack([i.ravel() * ell, j.ravel() * ell, np.zeros(n * n)], 1)
    vid = (i * n + j)
    cand = np.stack([np.stack([vid, vid + n], -1), np.stack([vid, vid + 1], -1)], 2)  # (n,n,2 kinds,2)
    ok = np.stack([i < n - 1, j < n - 1], -1)
    elements = cand[ok].astype(np.int32)
    E = elements.shape[0]
    normals = np.tile(np.array([0.0, 0.0, 1.0]), (E, 1))
    dims, mat = _sections(E, b, h)
    fixed = np.arange(n, dtype=np.int32)  # i = 0
    tip = ((n - 1) * n + np.arange(n)).astype(np.int64)
    forces = (tip, np.full(n, 2, np.int64), np.ones(n))
    return SimulationJob(nodes, elements, normals, fixed, forces, dims, mat)


def cubic_lattice(n=100, ell=0.1, b=0.01, h=0.01) -> SimulationJob:
    """Config 3: n^3 lattice, vertex id (i*n+j)*n+k; per vertex edges +i (normal
    (0,1,0)), +j (normal (0,0,1)), +k (normal (1,0,0)); layer k=0 fixed; unit
    load in x (dof 0) on every vertex of layer k=n-1."""
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    nodes = np.stack([i.ravel() * ell, j.ravel() * ell, k.ravel() * ell], 1).astype(np.float64)
    vid = (i * n + j) * n + k
    cand = np.stack([np.stack([vid, vid + n * n], -1), np.stack([vid, vid + n], -1), np.stack([vid, vid + 1], -1)], 3)
    ok = np.stack([i < n - 1, j < n - 1, k < n - 1], -1)   # (n,n,n,3)
    elements = cand[ok].astype(np.int32)
    kind = np.broadcast_to(np.arange(3), ok.shape)[ok]
    table = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    normals = table[kind]
    E = elements.shape[0]
    dims, mat = _sections(E, b, h)
    fixed = vid[:, :, 0].ravel().astype(np.int32)
    top = vid[:, :, n - 1].ravel().astype(np.int64)
    forces = (top, np.zeros(top.size, np.int64), np.ones(top.size))
    return SimulationJob(nodes, elements, normals, fixed, forces, dims, mat)


def gridshell(n=200, b=0.01, h=0.01, rise=0.1) -> SimulationJob:
    """Config 4: n x n grid with one diagonal per cell lifted to a shallow
    sine-sine shell; mean edge ~0.1; per-edge normal = mean of the adjacent face
    normals projected perpendicular to the edge and renormalised; boundary ring
    fixed; load dof 2 = -1 on every interior vertex."""
    ell = 0.1
    span = ell * (n - 1)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    x = i.ravel() * ell; y = j.ravel() * ell
    z = rise * span * np.sin(np.pi * x / span) * np.sin(np.pi * y / span)
    nodes = np.stack([x, y, z], 1)
    vid = i * n + j
    cand = np.stack([np.stack([vid, vid + n], -1), np.stack([vid, vid + 1], -1), np.stack([vid, vid + n + 1], -1)], 2)
    ok = np.stack([i < n - 1, j < n - 1, (i < n - 1) & (j < n - 1)], -1)
    elements = cand[ok].astype(np.int32)
    E = elements.shape[0]
    # smooth surface normal at the edge midpoint (analytic), made perpendicular to the edge
    mid = 0.5 * (nodes[elements[:, 0]] + nodes[elements[:, 1]])
    a = np.pi / span
    dzdx = rise * span * a * np.cos(a * mid[:, 0]) * np.sin(a * mid[:, 1])
    dzdy = rise * span * a * np.sin(a * mid[:, 0]) * np.cos(a * mid[:, 1])
    nrm = np.stack([-dzdx, -dzdy, np.ones(E)], 1)
    t = nodes[elements[:, 1]] - nodes[elements[:, 0]]
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    nrm -= (nrm * t).sum(1, keepdims=True) * t
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    dims, mat = _sections(E, b, h)
    border = (i == 0) | (i == n - 1) | (j == 0) | (j == n - 1)
    fixed = vid[border].astype(np.int32)
    inner = vid[~border].astype(np.int64)
    forces = (inner, np.full(inner.size, 2, np.int64), -np.ones(inner.size))
    return SimulationJob(nodes, elements, nrm, fixed, forces, dims, mat)


def random_mesh(seed=1234, V=200, k=5) -> SimulationJob:
    """Config 5 member: V points uniform in the unit cube, k-nearest-neighbour
    graph (undirected, de-duplicated) made connected by chaining components;
    random perpendicular unit normals; b,h ~ U[0.005,0.02], E ~ U[70,210] GPa,
    nu ~ U[0.2,0.35]; 4 lowest-z vertices fixed; 3 random nodal forces."""
    rng = np.random.default_rng(seed)
    P = rng.random((V, 3))
    d2 = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    nn = np.argsort(d2, axis=1)[:, :k]
    pairs = set()
    for a in range(V):
        for b_ in nn[a]:
            pairs.add((min(a, int(b_)), max(a, int(b_))))
    # connect components (union-find), linking each new component to the nearest seen vertex
    parent = list(range(V))

    def find(u):
        while parent[u] != u:
            parent[u] = parent[parent[u]]
            u = parent[u]
        return u
    for a, b_ in pairs:
        parent[find(a)] = find(b_)
    roots = sorted({find(v) for v in range(V)})
    for r in roots[1:]:
        comp = [v for v in range(V) if find(v) == r]
        rest = [v for v in range(V) if find(v) == find(roots[0])]
        sub = d2[np.ix_(comp, rest)]
        ia, ib = np.unravel_index(np.argmin(sub), sub.shape)
        a, b_ = comp[ia], rest[ib]
        pairs.add((min(a, b_), max(a, b_)))
        parent[find(a)] = find(b_)
    elements = np.array(sorted(pairs), np.int32)
    E = elements.shape[0]
    t = P[elements[:, 1]] - P[elements[:, 0]]
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    nrm = rng.standard_normal((E, 3))
    nrm -= (nrm * t).sum(1, keepdims=True) * t
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    dims = rng.uniform(0.005, 0.02, (E, 2)).astype(np.float32)
    mat = np.stack([rng.uniform(0.2, 0.35, E), rng.uniform(70, 210, E)], 1).astype(np.float32)
    fixed = np.argsort(P[:, 2])[:4].astype(np.int32)
    free = np.setdiff1d(np.arange(V), fixed)
    fn = rng.choice(free, 3, replace=False).astype(np.int64)
    fd = rng.integers(0, 3, 3).astype(np.int64)
    fm = rng.uniform(10, 100, 3)
    return SimulationJob(P, elements, nrm, fixed, (fn, fd, fm), dims, mat)


def merge_jobs(jobs) -> SimulationJob:
    """Disjoint union of edge meshes (config 5 batched as one block-diagonal job)."""
    nodes, elements, normals, fixed, dims, mat = [], [], [], [], [], []
    fn, fd, fm = [], [], []
    off = 0
    for jb in jobs:
        nodes.append(jb.nodes); elements.append(jb.elements + off); normals.append(jb.elementalNormals)
        fixed.append(np.asarray(jb.fixedNodes, np.int32) + off)
        dims.append(jb.beamDimensions); mat.append(jb.beamMaterial)
        a, b_, c = jb.nodalForces if isinstance(jb.nodalForces, tuple) else (
            np.array([f.index for f in jb.nodalForces], np.int64), np.array([f.dof for f in jb.nodalForces], np.int64),
            np.array([f.magnitude for f in jb.nodalForces], np.float64))
        fn.append(a + off); fd.append(b_); fm.append(c)
        off += jb.nodes.shape[0]
    return SimulationJob(np.concatenate(nodes), np.concatenate(elements).astype(np.int32), np.concatenate(normals),
                         np.concatenate(fixed).astype(np.int32), (np.concatenate(fn), np.concatenate(fd), np.concatenate(fm)),
                         np.concatenate(dims), np.concatenate(mat))
