"""CPU ORACLE (Python face) for the MeshBeamFEA simulation hot path.

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s cpu_baseline / ``--impl reference`` arm; never from the
product package ``meshbeamfea_b200``.

The element arithmetic lives in ``mbfea_oracle.c`` (see its header for the
reference citations and the "parity unpinned" statement).  This module adds the
linear solve the way the reference's dependency does it: a sparse *direct* LU.
threed-beam-fea calls ``Eigen::SparseLU`` (a port of sequential SuperLU) on the
Lagrange-multiplier KKT system; scipy's ``splu`` *is* SuperLU, so
``solve(..., form="kkt")`` is the closest executable stand-in that exists in
this image.  ``form="eliminate"`` solves the reduced SPD system the product's
PCG works on; tests check that both agree.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmbfea_oracle.so")
_lib = None

_i64 = C.c_int64
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "mbfea_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_kelem.restype = C.c_double
        _lib.orc_assemble_count.restype = _i64
        _lib.orc_free_map.restype = _i64
        _lib.orc_bsr_pattern.restype = _i64
        _lib.orc_pcg_bj.restype = _i64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


@dataclass
class Job:
    """Flat-array image of the reference's SimulationJob
    [REF src/beamsimulator.hpp:35-43]; matrices are column-major like Eigen's."""

    nodes: np.ndarray      # (V,3) float64
    elements: np.ndarray   # (E,2) int32
    normals: np.ndarray    # (E,3) float64
    fixed: np.ndarray      # (F,)  int32
    forces: list           # [(node, dof, magnitude)]
    dims: np.ndarray       # (E,2) float32  {b,h}
    material: np.ndarray   # (E,2) float32  {nu, E_GPa}

    def cm(self):
        """column-major flats (x[V],y[V],z[V]) etc."""
        return (np.ascontiguousarray(np.asarray(self.nodes, np.float64).T).ravel(),
                np.ascontiguousarray(np.asarray(self.elements, np.int32).T).ravel(),
                np.ascontiguousarray(np.asarray(self.normals, np.float64).T).ravel())


def beam_props(dims, material) -> np.ndarray:
    dims = np.ascontiguousarray(dims, np.float32).reshape(-1, 2)
    material = np.ascontiguousarray(material, np.float32).reshape(-1, 2)
    E = dims.shape[0]
    out = np.empty((E, 4), np.float64)
    lib().orc_beam_props(_i64(E), _p(dims, _pf), _p(material, _pf), _p(out, _pd))
    return out


def kelem(p0, p1, normal, props4):
    p0 = np.ascontiguousarray(p0, np.float64); p1 = np.ascontiguousarray(p1, np.float64)
    n = np.ascontiguousarray(normal, np.float64); pr = np.ascontiguousarray(props4, np.float64)
    Kl = np.empty((12, 12)); A = np.empty((12, 12)); Ke = np.empty((12, 12))
    L = lib().orc_kelem(_p(p0, _pd), _p(p1, _pd), _p(n, _pd), _p(pr, _pd), _p(Kl, _pd), _p(A, _pd), _p(Ke, _pd))
    return L, Kl, A, Ke


def kelem_all(job: Job, props=None) -> np.ndarray:
    nodes, elems, normals = job.cm()
    V, E = job.nodes.shape[0], job.elements.shape[0]
    props = beam_props(job.dims, job.material) if props is None else props
    out = np.empty((E, 12, 12), np.float64)
    lib().orc_kelem_all(_i64(V), _i64(E), _p(nodes, _pd), _p(elems, _pi32), _p(normals, _pd), _p(props, _pd), _p(out, _pd))
    return out


def assemble_full(job: Job, props=None):
    """Scalar CSR of order 6V, triplets summed in element order (exact zeros of
    K_e are never pushed, like Eigen's sparseView)."""
    import scipy.sparse as sp
    nodes, elems, normals = job.cm()
    V, E = job.nodes.shape[0], job.elements.shape[0]
    props = beam_props(job.dims, job.material) if props is None else props
    rowptr = np.zeros(6 * V + 1, np.int64)
    args = (_i64(V), _i64(E), _p(nodes, _pd), _p(elems, _pi32), _p(normals, _pd), _p(props, _pd))
    nnz = lib().orc_assemble_count(*args, _p(rowptr, _pi64))
    col = np.empty(max(nnz, 1), np.int32); val = np.empty(max(nnz, 1), np.float64)
    lib().orc_assemble_fill(*args, _p(rowptr, _pi64), _p(col, _pi32), _p(val, _pd))
    return sp.csr_matrix((val[:nnz], col[:nnz], rowptr), shape=(6 * V, 6 * V))


def free_map(V: int, fixed) -> np.ndarray:
    fixed = np.ascontiguousarray(fixed, np.int32)
    fm = np.empty(V, np.int32)
    nb = lib().orc_free_map(_i64(V), _i64(fixed.size), _p(fixed, _pi32), _p(fm, _pi32))
    if nb < 0:
        # [REF src/beamsimulator.cpp:134-137]
        raise RuntimeError("Vertex index is outside of valid range.")
    return fm


def load_vector(V: int, forces) -> np.ndarray:
    """f[6*node+dof] (+)= magnitude.  Duplicates accumulate (unpinned upstream
    behaviour; documented in DESIGN.md).  Out-of-range node raises like
    [REF src/beamsimulator.cpp:164-166]."""
    f = np.zeros(6 * V)
    for node, dof, mag in forces:
        if node > V - 1:
            raise IndexError("Node index is out of valid range.")
        f[6 * int(node) + int(dof)] += mag
    return f


def bsr_pattern(job: Job, fm: np.ndarray):
    _, elems, _ = job.cm()
    V, E = job.nodes.shape[0], job.elements.shape[0]
    nb = int((fm >= 0).sum())
    rowptr = np.zeros(nb + 1, np.int64)
    nnzb = lib().orc_bsr_pattern(_i64(V), _i64(E), _p(elems, _pi32), _p(fm, _pi32), _i64(nb), _p(rowptr, _pi64), None)
    col = np.empty(max(nnzb, 1), np.int32)
    lib().orc_bsr_pattern(_i64(V), _i64(E), _p(elems, _pi32), _p(fm, _pi32), _i64(nb), _p(rowptr, _pi64), _p(col, _pi32))
    return rowptr, col[:nnzb]


def bsr_assemble(job: Job, fm: np.ndarray, rowptr, col, props=None) -> np.ndarray:
    nodes, elems, normals = job.cm()
    V, E = job.nodes.shape[0], job.elements.shape[0]
    props = beam_props(job.dims, job.material) if props is None else props
    nb = rowptr.size - 1
    vals = np.empty((max(col.size, 1), 6, 6), np.float64)
    colc = np.ascontiguousarray(col, np.int32)
    lib().orc_bsr_assemble(_i64(V), _i64(E), _p(nodes, _pd), _p(elems, _pi32), _p(normals, _pd), _p(props, _pd),
                           _p(fm, _pi32), _i64(nb), _p(rowptr, _pi64), _p(colc, _pi32), _p(vals, _pd))
    return vals[:col.size]


def pcg_bj(rowptr, col, vals, f, tol=1e-8, max_iter=10**9):
    nb = rowptr.size - 1
    x = np.zeros(6 * nb); rel = C.c_double(0.0)
    colc = np.ascontiguousarray(col, np.int32); v = np.ascontiguousarray(vals, np.float64)
    f = np.ascontiguousarray(f, np.float64)
    it = lib().orc_pcg_bj(_i64(nb), _p(rowptr, _pi64), _p(colc, _pi32), _p(v, _pd), _p(f, _pd), _p(x, _pd),
                          C.c_double(tol), _i64(max_iter), C.byref(rel))
    return x, int(it), rel.value


def element_forces(job: Job, U_rowmajor: np.ndarray, props=None):
    nodes, elems, normals = job.cm()
    V, E = job.nodes.shape[0], job.elements.shape[0]
    props = beam_props(job.dims, job.material) if props is None else props
    U = np.ascontiguousarray(U_rowmajor, np.float64).reshape(V, 6)
    F = np.empty((6, 2 * E)); Fe = np.empty((E, 12))
    lib().orc_element_forces(_i64(V), _i64(E), _p(nodes, _pd), _p(elems, _pi32), _p(normals, _pd), _p(props, _pd),
                             _p(U, _pd), _p(F, _pd), _p(Fe, _pd))
    return F, Fe


def _lu_solve(A, rhs, refine: int):
    """SuperLU solve + `refine` steps of iterative refinement (residual in
    long double).  refine=0 is the raw factorisation the reference would use."""
    import scipy.sparse.linalg as spla
    lu = spla.splu(A.tocsc())
    x = lu.solve(rhs)
    Ac = A.tocsr()
    for _ in range(refine):
        r = (rhs.astype(np.longdouble) - (Ac @ x).astype(np.longdouble)).astype(np.float64)
        x = x + lu.solve(r)
    return x


def residual_dd(A_csr, x_hi, x_lo, f) -> np.ndarray:
    """r = f - A (x_hi + x_lo) with double-double products and sums (orc_csr_residual_dd), rounded to double."""
    A = A_csr.tocsr()
    A.sort_indices()
    n = A.shape[0]
    rp = np.ascontiguousarray(A.indptr, np.int64); ci = np.ascontiguousarray(A.indices, np.int32)
    va = np.ascontiguousarray(A.data, np.float64)
    xh = np.ascontiguousarray(x_hi, np.float64)
    xl = None if x_lo is None else np.ascontiguousarray(x_lo, np.float64)
    ff = np.ascontiguousarray(f, np.float64)
    r = np.empty(n)
    L = lib()
    L.orc_csr_residual_dd.restype = None
    L.orc_csr_residual_dd(C.c_int64(n), _p(rp, _pi64), _p(ci, _pi32), _p(va, _pd), _p(xh, _pd),
                          _p(xl, _pd) if xl is not None else None, _p(ff, _pd), _p(r, _pd))
    return r


def solve_dd(job: Job, steps: int = 6):
    """Higher-precision reference: SuperLU factors of the eliminated system + iterative refinement whose residual
    and solution update are carried in double-double.  Each step gains ~ -log10(kappa * 1e-16) digits, so a few
    steps reach the exact solution of the FP64-represented system to far below 1e-8 even at kappa ~ 1e10
    (planar 100 x 100 cantilever).  Returns (U (V,6) rounded to double, edge forces from it, history of
    ||f - K x|| / ||f|| per step)."""
    import scipy.sparse.linalg as spla
    V = job.nodes.shape[0]
    props = beam_props(job.dims, job.material)
    K = assemble_full(job, props)
    f = load_vector(V, job.forces)
    fm = free_map(V, job.fixed)
    free_dofs = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
    Kr = K[free_dofs][:, free_dofs].tocsr()
    fr = np.ascontiguousarray(f[free_dofs])
    lu = spla.splu(Kr.tocsc())
    xh = lu.solve(fr)
    xl = np.zeros_like(xh)
    nf = np.linalg.norm(fr)
    hist = []
    for _ in range(steps):
        r = residual_dd(Kr, xh, xl, fr)
        hist.append(float(np.linalg.norm(r) / nf))
        d = lu.solve(r)
        s = xh + d
        bb = s - xh
        e = (xh - (s - bb)) + (d - bb)
        xl = xl + e
        xh = s + xl
        xl = xl - (xh - s)
    hist.append(float(np.linalg.norm(residual_dd(Kr, xh, xl, fr)) / nf))
    u = np.zeros(6 * V)
    u[free_dofs] = xh
    U = u.reshape(V, 6)
    F, _ = element_forces(job, U, props)
    return U, F, hist


def solve_dd_pcg(job: Job, passes: int = 3, tol: float = 1e-9):
    """Higher-precision reference for meshes too large for the sparse direct solve (3-D lattices beyond ~20^3 take
    minutes in SuperLU): the oracle's own block-Jacobi PCG as the inner solver of an iterative refinement whose
    residual f - K x and solution update are carried in double-double.  Pass k solves K d = r_k to `tol`, so the
    error contracts by about `tol` per pass whatever the conditioning, and the double-double residual makes the
    limit the exact solution of the FP64-represented system.  Returns (U, edge forces, residual history)."""
    V = job.nodes.shape[0]
    props = beam_props(job.dims, job.material)
    K = assemble_full(job, props)
    f = load_vector(V, job.forces)
    fm = free_map(V, job.fixed)
    free_dofs = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
    Kr = K[free_dofs][:, free_dofs].tocsr()
    fr = np.ascontiguousarray(f[free_dofs])
    rp, col = bsr_pattern(job, fm)
    vals = bsr_assemble(job, fm, rp, col, props)
    nf = np.linalg.norm(fr)
    xh = np.zeros_like(fr); xl = np.zeros_like(fr)
    hist = []
    for _ in range(passes):
        r = residual_dd(Kr, xh, xl, fr)
        hist.append(float(np.linalg.norm(r) / nf))
        d, _, _ = pcg_bj(rp, col, vals, r, tol=tol)
        s = xh + d
        bb = s - xh
        e = (xh - (s - bb)) + (d - bb)
        xl = xl + e
        xh = s + xl
        xl = xl - (xh - s)
    hist.append(float(np.linalg.norm(residual_dd(Kr, xh, xl, fr)) / nf))
    u = np.zeros(6 * V)
    u[free_dofs] = xh
    U = u.reshape(V, 6)
    F, _ = element_forces(job, U, props)
    return U, F, hist


def solve(job: Job, form: str = "eliminate", refine: int = 2):
    """Full pipeline with a sparse direct solve.

    Returns (U (V,6), edge_forces (6, 2E)) in the layout of the reference's
    SimulationResults [REF src/beamsimulator.cpp:190-210].
    """
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    V = job.nodes.shape[0]
    props = beam_props(job.dims, job.material)
    K = assemble_full(job, props)
    f = load_vector(V, job.forces)
    fm = free_map(V, job.fixed)
    if form == "eliminate":
        free_dofs = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
        Kr = K[free_dofs][:, free_dofs].tocsc()
        u = np.zeros(6 * V)
        u[free_dofs] = _lu_solve(Kr, f[free_dofs], refine)
    elif form == "kkt":
        # one Lagrange row/col per BC, 6 per fixed vertex in the order the
        # reference emits them [REF src/beamsimulator.cpp:139-150]
        fixed = np.asarray(job.fixed, np.int64)
        bc = (6 * fixed[:, None] + np.arange(6)[None, :]).ravel()
        m = bc.size
        Cmat = sp.csr_matrix((np.ones(m), (np.arange(m), bc)), shape=(m, 6 * V))
        KKT = sp.bmat([[K, Cmat.T], [Cmat, None]], format="csc")
        rhs = np.concatenate([f, np.zeros(m)])
        u = _lu_solve(KKT, rhs, refine)[:6 * V]
    else:
        raise ValueError(form)
    U = u.reshape(V, 6)
    F, _ = element_forces(job, U, props)
    return U, F
