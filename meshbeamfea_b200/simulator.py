"""Host-side mirror of the reference's simulation API over the C ABI.

Names, argument meaning and error behaviour follow the reference's
``BeamSimulator`` boundary [REF src/beamsimulator.hpp:29-52,78-95;
src/beamsimulator.cpp:10-69,126-171,190-210] so that parity tests read like the
reference's own harness [REF src/beamsimulatortester.cpp:33-37,49-64].  All
arithmetic happens in libmbfea.so on the GPU; this module only marshals arrays.

``Context`` is the thin object view of ``mbfea_ctx`` (one per GPU / rank) and
also exposes the parity taps the tests compare against the oracle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import Options, Stats, PlanSizes, ptr

_pd, _pf, _pi32, _pi64 = L._pd, L._pf, L._pi32, L._pi64


class ContractViolation(Exception):
    """Where the reference hits a GSL ``Expects`` and terminates
    [REF src/beamsimulator.cpp:41,73-75,83; src/beam.hpp:14,25]."""


class RangeError(RuntimeError):
    """std::range_error("Node index is out of valid range.")
    [REF src/beamsimulator.cpp:164-166]."""


class SingularSystem(RuntimeError):
    pass


class NotConverged(RuntimeError):
    pass


class MbfeaError(RuntimeError):
    pass


def _raise(code: int, msg: str):
    if code == L.EPRECOND:
        raise ContractViolation(msg)
    if code == L.ERANGE_FIXED:
        raise RuntimeError(msg)  # std::runtime_error [REF src/beamsimulator.cpp:134-137]
    if code == L.ERANGE_FORCE:
        raise RangeError(msg)
    if code == L.ESINGULAR:
        raise SingularSystem(msg)
    if code == L.ENOCONV:
        raise NotConverged(msg)
    if code == L.EIO:
        raise OSError(msg)
    raise MbfeaError(f"mbfea status {code} ({L.lib().mbfea_status_string(code).decode()}): {msg}")


# ---------------------------------------------------------------------------
# reference value types  [REF src/beam.hpp:9-28; src/beamsimulator.hpp:29-52]
# ---------------------------------------------------------------------------
@dataclass
class BeamDimensions:
    b: float = 1.0
    h: float = 1.0

    def __post_init__(self):
        if not (self.b > 0 and self.h > 0):  # Expects(b > 0 && h > 0)
            raise ContractViolation("BeamDimensions: b and h must be positive")


@dataclass
class BeamMaterial:
    poissonsRatio: float = 0.3
    youngsModulusGPascal: float = 210.0

    def __post_init__(self):
        if not (-1 <= self.poissonsRatio <= 0.5):  # Expects(poissonsRatio <= 0.5 && >= -1)
            raise ContractViolation("BeamMaterial: Poisson's ratio outside [-1, 0.5]")


@dataclass
class NodalForce:
    index: int
    dof: int
    magnitude: float


@dataclass
class SimulationJob:
    """[REF src/beamsimulator.hpp:35-43].  ``beamDimensions`` / ``beamMaterial``
    accept either lists of the dataclasses above or (E,2) float32 arrays
    ({b,h} and {nu,E_GPa}) for large meshes."""
    nodes: np.ndarray                      # (V,3) float64
    elements: np.ndarray                   # (E,2) int32
    elementalNormals: np.ndarray           # (E,3) float64
    fixedNodes: np.ndarray                 # (F,)  int32
    nodalForces: Sequence = field(default_factory=list)
    beamDimensions: object = None
    beamMaterial: object = None


class SimulationResults:
    """[REF src/beamsimulator.hpp:45-52; src/beamsimulator.cpp:190-210]:
    ``edgeForces[c][2e]`` / ``[2e+1]`` = component c (N,Ty,Tz,Mx,My,Mz) at end 0 /
    end 1 of beam e (local frame); ``nodalDisplacements`` is V x 6."""

    def __init__(self, nodalDisplacements: Optional[np.ndarray] = None, edgeForces: Optional[np.ndarray] = None):
        self.nodalDisplacements = nodalDisplacements if nodalDisplacements is not None else np.zeros((0, 6))
        ef = edgeForces if edgeForces is not None else np.zeros((6, 0))
        self.edgeForces: List[np.ndarray] = [ef[c] for c in range(ef.shape[0])]


def _forces_arrays(forces):
    if isinstance(forces, tuple) and len(forces) == 3 and isinstance(forces[0], np.ndarray):
        n, d, m = forces
    else:
        forces = list(forces)
        n = np.array([f.index if isinstance(f, NodalForce) else f[0] for f in forces], np.int64)
        d = np.array([f.dof if isinstance(f, NodalForce) else f[1] for f in forces], np.int64)
        m = np.array([f.magnitude if isinstance(f, NodalForce) else f[2] for f in forces], np.float64)
    return (np.ascontiguousarray(n, np.int64), np.ascontiguousarray(d, np.int64), np.ascontiguousarray(m, np.float64))


def _pairs(x, E, names, default):
    """(E,2) float32 from a list of dataclasses, an array, or None (defaults)."""
    if x is None:
        return np.tile(np.asarray(default, np.float32), (E, 1))
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, np.float32).reshape(-1, 2)
    return np.array([[getattr(v, names[0]), getattr(v, names[1])] for v in x], np.float32).reshape(-1, 2)


# ---------------------------------------------------------------------------
class Context:
    """One ``mbfea_ctx``: device buffers, stream, PCG graph, optional NCCL comm."""

    def __init__(self, device: int = -1, tol: float = 1e-8, max_iter: int = 0, verbose: bool = False,
                 spmv_kernel: int = 0, check_every: int = 0, use_graph: bool = True, comm_mode: int = 0, storage: int = 0, cg_variant: int = 0, batch_mode: int = 0, reorder: int = 0,
                 precond: int = 0, ml_cell: float = 0.0, ml_ratio: float = 0.0, ml_coarsest_rows: int = 0, ml_smooth_from: int = 0):
        self._lib = L.lib()
        o = Options()
        self._lib.mbfea_options_default(C.byref(o))
        o.device, o.tol, o.max_iter, o.verbose = device, tol, max_iter, int(verbose)
        o.spmv_kernel, o.check_every, o.use_graph = spmv_kernel, check_every, int(use_graph)
        o.comm_mode = comm_mode
        o.storage = storage
        o.cg_variant = cg_variant
        o.batch_mode = batch_mode
        o.reorder = reorder
        o.precond, o.ml_cell, o.ml_ratio, o.ml_coarsest_rows, o.ml_smooth_from = precond, ml_cell, ml_ratio, ml_coarsest_rows, ml_smooth_from
        self._h = C.c_void_p()
        rc = self._lib.mbfea_create(C.byref(self._h), C.byref(o))
        if rc != L.OK:
            msg = self._lib.mbfea_last_error(None).decode()
            self._h = C.c_void_p()
            _raise(rc, msg)
        self.V = self.E = 0

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mbfea_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, allow=()):
        if rc != L.OK and rc not in allow:
            _raise(rc, self._lib.mbfea_last_error(self._h).decode())
        return rc

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = L.lib().mbfea_comm_unique_id(buf)
        if rc != L.OK:
            _raise(rc, L.lib().mbfea_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._lib.mbfea_comm_init(self._h, buf, rank, world))

    # -- job
    def set_job(self, nodes, elements, normals, dims_bh, mat_nuE, fixed, forces=()):
        nodes = np.asarray(nodes, np.float64).reshape(-1, 3)
        elements = np.asarray(elements, np.int32).reshape(-1, 2)
        normals = np.asarray(normals, np.float64).reshape(-1, 3)
        V, E = nodes.shape[0], elements.shape[0]
        dims = _pairs(dims_bh, E, ("b", "h"), (1.0, 1.0))
        mat = _pairs(mat_nuE, E, ("poissonsRatio", "youngsModulusGPascal"), (0.3, 210.0))
        # Expects(elements.rows() == normals.rows() == beamDimensions.size() == beamMaterial.size())
        # [REF src/beamsimulator.cpp:73-75]
        if not (normals.shape[0] == E and dims.shape[0] == E and mat.shape[0] == E):
            raise ContractViolation("precondition violated: per-beam arrays must all have #elements rows")
        # column-major flats == the memory image of the reference's Eigen members
        nodes_cm = np.ascontiguousarray(nodes.T).ravel()
        elems_cm = np.ascontiguousarray(elements.T).ravel()
        normals_cm = np.ascontiguousarray(normals.T).ravel()
        fixed = np.ascontiguousarray(np.asarray(fixed, np.int32).ravel())
        fn, fd, fm = _forces_arrays(forces)
        self._ck(self._lib.mbfea_set_job(self._h, V, E, ptr(nodes_cm, _pd), ptr(elems_cm, _pi32), ptr(normals_cm, _pd),
                                         ptr(dims, _pf), ptr(mat, _pf), fixed.size, ptr(fixed, _pi32),
                                         fn.size, ptr(fn, _pi64), ptr(fd, _pi64), ptr(fm, _pd)))
        self.V, self.E = V, E

    def set_job_raw(self, V, E, nodes_cm, elems_cm, normals_cm, dims, mat, fixed, fn, fd, fm):
        """Arrays already in the C-ABI layout (used by bench.py to keep marshalling out of the timed region)."""
        self._ck(self._lib.mbfea_set_job(self._h, V, E, ptr(nodes_cm, _pd), ptr(elems_cm, _pi32), ptr(normals_cm, _pd),
                                         ptr(dims, _pf), ptr(mat, _pf), fixed.size, ptr(fixed, _pi32),
                                         fn.size, ptr(fn, _pi64), ptr(fd, _pi64), ptr(fm, _pd)))
        self.V, self.E = V, E

    def load_scenario(self, config_json: str, ply_override: Optional[str] = None):
        self._ck(self._lib.mbfea_load_scenario(self._h, config_json.encode(), ply_override.encode() if ply_override else None))
        s = self.stats()
        self.V, self.E = s["n_vertices"], s["n_beams"]

    def set_forces(self, forces):
        fn, fd, fm = _forces_arrays(forces)
        self._ck(self._lib.mbfea_set_forces(self._h, fn.size, ptr(fn, _pi64), ptr(fd, _pi64), ptr(fm, _pd)))

    # -- the path
    def execute(self, allow_unconverged: bool = False) -> int:
        allow = (L.ENOCONV,) if allow_unconverged else ()
        return self._ck(self._lib.mbfea_execute(self._h), allow)

    def assemble(self):
        self._ck(self._lib.mbfea_assemble(self._h))

    def solve(self, allow_unconverged: bool = False) -> int:
        allow = (L.ENOCONV,) if allow_unconverged else ()
        return self._ck(self._lib.mbfea_solve(self._h), allow)

    def recover(self):
        self._ck(self._lib.mbfea_recover(self._h))

    def displacements(self) -> np.ndarray:
        """V x 6 (a Fortran-ordered view == Eigen's column-major MatrixXd)."""
        U = np.empty(6 * self.V, np.float64)
        self._ck(self._lib.mbfea_get_displacements(self._h, ptr(U, _pd)))
        return U.reshape(6, self.V).T

    def edge_forces(self) -> np.ndarray:
        """6 x 2E."""
        F = np.empty((6, 2 * self.E), np.float64)
        self._ck(self._lib.mbfea_get_edge_forces(self._h, ptr(F, _pd)))
        return F

    def force_ranges(self):
        """(min[6], max[6]) of the beam-end forces per component (colour-bar labels)."""
        lo, hi = np.empty(6), np.empty(6)
        self._ck(self._lib.mbfea_get_force_ranges(self._h, ptr(lo, _pd), ptr(hi, _pd)))
        return lo, hi

    def edge_forces_normalized(self) -> np.ndarray:
        """6 x 2E, each component mapped to [0,1] by its own min/max."""
        T = np.empty((6, 2 * self.E), np.float64)
        self._ck(self._lib.mbfea_get_edge_forces_normalized(self._h, ptr(T, _pd)))
        return T

    def displacements_into(self, U_flat: np.ndarray):
        self._ck(self._lib.mbfea_get_displacements(self._h, ptr(U_flat, _pd)))

    def edge_forces_into(self, F_flat: np.ndarray):
        self._ck(self._lib.mbfea_get_edge_forces(self._h, ptr(F_flat, _pd)))

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._lib.mbfea_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def write_displacements_csv(self, path: str):
        self._ck(self._lib.mbfea_write_displacements_csv(self._h, path.encode()))

    def write_element_forces_csv(self, path: str):
        self._ck(self._lib.mbfea_write_element_forces_csv(self._h, path.encode()))

    # -- parity taps
    def props(self) -> np.ndarray:
        out = np.empty((self.E, 4), np.float64)
        self._ck(self._lib.mbfea_get_props(self._h, ptr(out, _pd)))
        return out

    def dofmap(self) -> np.ndarray:
        out = np.empty(self.V, np.int32)
        self._ck(self._lib.mbfea_get_dofmap(self._h, ptr(out, _pi32)))
        return out

    def kelem(self) -> np.ndarray:
        out = np.empty((self.E, 12, 12), np.float64)
        self._ck(self._lib.mbfea_get_kelem(self._h, ptr(out, _pd)))
        return out

    def bsr_sizes(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self._lib.mbfea_get_bsr_sizes(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def bsr(self, values: bool = True):
        """(rowptr, colidx, vals, row_ids): the owned block rows in canonical reduced numbering
        (ascending row_ids; global canonical column ids).  On one rank row_ids == arange(nb)."""
        nb, nnzb, row0, ng = self.bsr_sizes()
        rowptr = np.empty(nb + 1, np.int64)
        col = np.empty(max(nnzb, 1), np.int32)
        vals = np.empty((max(nnzb, 1), 6, 6), np.float64) if values else None
        rows = np.empty(max(nb, 1), np.int64)
        self._ck(self._lib.mbfea_get_bsr(self._h, ptr(rowptr, _pi64), ptr(col, _pi32), ptr(vals, _pd), ptr(rows, _pi64)))
        return rowptr, col[:nnzb], (vals[:nnzb] if values else None), rows[:nb]

    def load(self) -> np.ndarray:
        nb = self.bsr_sizes()[0]
        out = np.empty(6 * nb, np.float64)
        self._ck(self._lib.mbfea_get_load(self._h, ptr(out, _pd)))
        return out

    def spmv(self, x: np.ndarray, kernel: int = 0) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty_like(x)
        self._ck(self._lib.mbfea_spmv_with(self._h, kernel, ptr(x, _pd), ptr(y, _pd)))
        return y

    def recover_from(self, U: np.ndarray):
        U_cm = np.ascontiguousarray(np.asarray(U, np.float64).reshape(self.V, 6).T).ravel()
        self._ck(self._lib.mbfea_recover_from(self._h, ptr(U_cm, _pd)))

    def trace(self, iters: int = 256) -> np.ndarray:
        """(iters, 8) device globaltimer stamps in ns (needs MBFEA_TRACE=1 during solve)."""
        out = np.zeros((iters, 8), np.uint64)
        n = C.c_int64()
        self._ck(self._lib.mbfea_debug_trace(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), iters, C.byref(n)))
        return out

    def compare_plan(self):
        """(mismatch counts of the 12 set-up arrays against the host builders, 0 / 1 / 2: what was built on the device)"""
        bad = np.zeros(12, np.int64)
        where = C.c_int32()
        self._ck(self._lib.mbfea_debug_compare_plan(self._h, bad.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(where)))
        return bad, int(where.value)

    def compare_hierarchy(self):
        """(mismatch counters [levels, 16] of the multilevel hierarchy against the host builder, built on the device?)"""
        bad = np.zeros(256, np.int64)
        levels, where = C.c_int32(), C.c_int32()
        self._ck(self._lib.mbfea_debug_compare_hierarchy(self._h, bad.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(levels), C.byref(where)))
        return bad.reshape(16, 16)[:levels.value], int(where.value)

    def time_spmv(self, kernel: int = 0, warmup: int = 3, reps: int = 20) -> float:
        ms = C.c_double()
        self._ck(self._lib.mbfea_time_spmv(self._h, kernel, warmup, reps, C.byref(ms)))
        return ms.value

    def time_assemble(self, warmup: int = 3, reps: int = 10) -> float:
        ms = C.c_double()
        self._ck(self._lib.mbfea_time_assemble(self._h, warmup, reps, C.byref(ms)))
        return ms.value


def plan_partition(V: int, elements: np.ndarray, fixed: np.ndarray, rank: int, world: int):
    """Host-only vertex-block partition plan (no GPU): see mbfea_plan_partition."""
    lib = L.lib()
    elements = np.asarray(elements, np.int32).reshape(-1, 2)
    E = elements.shape[0]
    elems_cm = np.ascontiguousarray(elements.T).ravel()
    fixed = np.ascontiguousarray(np.asarray(fixed, np.int32).ravel())
    sz = PlanSizes()
    args = (V, E, ptr(elems_cm, _pi32), fixed.size, ptr(fixed, _pi32), rank, world)
    rc = lib.mbfea_plan_partition(*args, C.byref(sz), None, None, None, None, None, None)
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    nb = sz.row_end - sz.row_begin
    rowptr = np.empty(nb + 1, np.int64)
    col = np.empty(max(sz.nnz_blocks, 1), np.int32)
    gg = np.empty(max(sz.n_ghost, 1), np.int64)
    go = np.empty(max(sz.n_ghost, 1), np.int32)
    sr = np.empty(max(sz.n_send, 1), np.int64)
    sd = np.empty(max(sz.n_send, 1), np.int32)
    rc = lib.mbfea_plan_partition(*args, C.byref(sz), ptr(rowptr, _pi64), ptr(col, _pi32), ptr(gg, _pi64),
                                  ptr(go, _pi32), ptr(sr, _pi64), ptr(sd, _pi32))
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    return dict(nb_global=sz.nb_global, row_begin=sz.row_begin, row_end=sz.row_end, rowptr=rowptr,
                colidx=col[:sz.nnz_blocks], ghost_global=gg[:sz.n_ghost], ghost_owner=go[:sz.n_ghost],
                send_rows=sr[:sz.n_send], send_dest=sd[:sz.n_send], n_neighbors=sz.n_neighbors,
                n_components=sz.n_components, max_component_rows=sz.max_component_rows)


def host_renumber(V: int, elements: np.ndarray, fixed: np.ndarray, mode: int = 0):
    """The internal breadth-first renumbering set_job would use (host only): (perm, reordered)."""
    lib = L.lib()
    elements = np.asarray(elements, np.int32).reshape(-1, 2)
    elems_cm = np.ascontiguousarray(elements.T).ravel()
    fixed = np.ascontiguousarray(np.asarray(fixed, np.int32).ravel())
    n_free = V - np.unique(fixed).size
    perm = np.empty(max(n_free, 1), np.int32)
    flag = C.c_int32()
    rc = lib.mbfea_host_renumber(V, elements.shape[0], ptr(elems_cm, _pi32), fixed.size, ptr(fixed, _pi32), mode,
                                 ptr(perm, _pi32), C.byref(flag))
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    return perm[:n_free], bool(flag.value)


def host_contributors(V: int, elements: np.ndarray, fixed: np.ndarray, rank: int = 0, world: int = 1):
    """Assembly plan of one rank (host only): (contrib_ptr, contrib) per block of the rank's pattern."""
    lib = L.lib()
    elements = np.asarray(elements, np.int32).reshape(-1, 2)
    elems_cm = np.ascontiguousarray(elements.T).ravel()
    fixed = np.ascontiguousarray(np.asarray(fixed, np.int32).ravel())
    sizes = np.zeros(2, np.int64)
    args = (V, elements.shape[0], ptr(elems_cm, _pi32), fixed.size, ptr(fixed, _pi32), rank, world, ptr(sizes, _pi64))
    rc = lib.mbfea_host_contributors(*args, None, None)
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    cptr = np.zeros(int(sizes[0]) + 1, np.int32)
    contrib = np.zeros(max(int(sizes[1]), 1), np.int32)
    rc = lib.mbfea_host_contributors(*args, ptr(cptr, _pi32), ptr(contrib, _pi32))
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    return cptr, contrib[:int(sizes[1])]


def host_spd_inverse(A: np.ndarray) -> np.ndarray:
    """Explicit inverse of a dense SPD matrix on the host (coarsest level of the multilevel preconditioner)."""
    lib = L.lib()
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    out = np.empty_like(A)
    rc = lib.mbfea_host_spd_inverse(n, ptr(A, _pd), ptr(out, _pd))
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    return out


def host_hierarchy(nodes: np.ndarray, elements: np.ndarray, fixed: np.ndarray, cell0: float = 0.0, ratio: float = 2.0,
                   max_levels: int = 8, coarsest_rows: int = 128) -> list:
    """Aggregate hierarchy of the multilevel preconditioner (host half of SURVEY 8(f) rank 4; not used by the
    solver yet): one dict per coarse level with agg, off, rowptr, colidx, gal_ptr, gal_src."""
    lib = L.lib()
    nodes = np.asarray(nodes, np.float64).reshape(-1, 3)
    elements = np.asarray(elements, np.int32).reshape(-1, 2)
    V = nodes.shape[0]
    nodes_cm = np.ascontiguousarray(nodes.T).ravel()
    elems_cm = np.ascontiguousarray(elements.T).ravel()
    fixed = np.ascontiguousarray(np.asarray(fixed, np.int32).ravel())
    h = C.c_void_p()
    rc = lib.mbfea_hierarchy_open(V, elements.shape[0], ptr(nodes_cm, _pd), ptr(elems_cm, _pi32), fixed.size,
                                  ptr(fixed, _pi32), cell0, ratio, max_levels, coarsest_rows, C.byref(h))
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    try:
        out = []
        for lvl in range(lib.mbfea_hierarchy_levels(h)):
            sz = np.zeros(6, np.int64)
            lib.mbfea_hierarchy_sizes(h, lvl, ptr(sz, _pi64))
            nf, nc, bc, bf = (int(v) for v in sz[:4])
            d = dict(agg=np.zeros(nf, np.int32), off=np.zeros(3 * nf), rowptr=np.zeros(nc + 1, np.int32),
                     colidx=np.zeros(bc, np.int32), gal_ptr=np.zeros(bc + 1, np.int32), gal_src=np.zeros(bf, np.int32))
            lib.mbfea_hierarchy_copy(h, lvl, ptr(d["agg"], _pi32), ptr(d["off"], _pd), ptr(d["rowptr"], _pi32),
                                     ptr(d["colidx"], _pi32), ptr(d["gal_ptr"], _pi32), ptr(d["gal_src"], _pi32))
            d["off"] = d["off"].reshape(nf, 3)
            d.update(n_fine=nf, n_coarse=nc)
            out.append(d)
        return out
    finally:
        lib.mbfea_hierarchy_close(h)


def host_solver_format(nodes: Optional[np.ndarray], elements: np.ndarray, fixed: np.ndarray, V: Optional[int] = None,
                       mode: int = 0, rank: int = 0, world: int = 1) -> dict:
    """The symmetric-half solver format set_job would build on one rank (host only, internal numbering,
    local block rows): which row stores each pair's block, the transposed references and the tile layout."""
    lib = L.lib()
    elements = np.asarray(elements, np.int32).reshape(-1, 2)
    elems_cm = np.ascontiguousarray(elements.T).ravel()
    fixed = np.ascontiguousarray(np.asarray(fixed, np.int32).ravel())
    nodes_cm = None
    if nodes is not None:
        nodes = np.asarray(nodes, np.float64).reshape(-1, 3)
        V = nodes.shape[0]
        nodes_cm = np.ascontiguousarray(nodes.T).ravel()
    sizes = np.zeros(8, np.int64)
    args = (V, elements.shape[0], ptr(nodes_cm, _pd) if nodes_cm is not None else None, ptr(elems_cm, _pi32), fixed.size,
            ptr(fixed, _pi32), mode, rank, world, ptr(sizes, _pi64))
    rc = lib.mbfea_host_solver_format(*args, None, None, None, None, None, None, None)
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    nb, _, stored, refs = (int(v) for v in sizes[:4])
    out = {k: np.zeros(max(n, 1), np.int32) for k, n in (("uptr", nb + 1), ("ucol", stored), ("lptr", nb + 1), ("lcol", refs),
                                                          ("lofs", refs), ("tile_base", (nb + 31) // 32 + 1),
                                                          ("perm", V - np.unique(fixed).size))}
    rc = lib.mbfea_host_solver_format(*args, *(ptr(out[k], _pi32) for k in ("uptr", "ucol", "lptr", "lcol", "lofs", "tile_base", "perm")))
    if rc != L.OK:
        _raise(rc, lib.mbfea_status_string(rc).decode())
    out = {"uptr": out["uptr"][:nb + 1], "ucol": out["ucol"][:stored], "lptr": out["lptr"][:nb + 1], "lcol": out["lcol"][:refs],
           "lofs": out["lofs"][:refs], "tile_base": out["tile_base"][:(nb + 31) // 32 + 1], "perm": out["perm"]}
    out.update(n_block_rows=nb, n_offdiag=int(sizes[1]), n_stored=stored, n_refs=refs, value_units=int(sizes[4]),
               balanced=bool(sizes[5]), longest_row=int(sizes[6]), renumbered=bool(sizes[7]))
    return out


def read_scenario(config_json: str, ply_override: Optional[str] = None):
    """Parse config.json + its PLY edge mesh on the host (no GPU) into a
    ``SimulationJob`` [REF src/utilities.hpp:50-95; src/edgemesh.cpp:115-135]."""
    lib = L.lib()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.mbfea_scenario_open(config_json.encode(), ply_override.encode() if ply_override else None,
                                 C.byref(h), err, 512)
    if rc != L.OK:
        _raise(rc, err.value.decode())
    try:
        V, E, F, nF = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        lib.mbfea_scenario_sizes(h, C.byref(V), C.byref(E), C.byref(F), C.byref(nF))
        V, E, F, nF = V.value, E.value, F.value, nF.value
        nodes = np.empty(3 * V); elems = np.empty(2 * E, np.int32); normals = np.empty(3 * E)
        dims = np.empty((E, 2), np.float32); mat = np.empty((E, 2), np.float32)
        fixed = np.empty(max(F, 1), np.int32)
        fn = np.empty(max(nF, 1), np.int64); fd = np.empty(max(nF, 1), np.int64); fm = np.empty(max(nF, 1))
        lib.mbfea_scenario_copy(h, ptr(nodes, _pd), ptr(elems, _pi32), ptr(normals, _pd), ptr(dims, _pf),
                                ptr(mat, _pf), ptr(fixed, _pi32), ptr(fn, _pi64), ptr(fd, _pi64), ptr(fm, _pd))
    finally:
        lib.mbfea_scenario_close(h)
    return SimulationJob(nodes=nodes.reshape(3, V).T.copy(), elements=elems.reshape(2, E).T.copy(),
                         elementalNormals=normals.reshape(3, E).T.copy(), fixedNodes=fixed[:F].copy(),
                         nodalForces=[NodalForce(int(a), int(b), float(c)) for a, b, c in zip(fn[:nF], fd[:nF], fm[:nF])],
                         beamDimensions=dims, beamMaterial=mat)


# ---------------------------------------------------------------------------
class BeamSimulator:
    """Drop-in for the reference class [REF src/beamsimulator.hpp:78-95].

    Differences, all opt-in knobs with reference-compatible defaults elsewhere:
    the O(V+E) stdout dump of ``printInfo`` [REF src/beamsimulator.cpp:28-38] is
    not reproduced, and the two result CSVs are written only after the
    corresponding ``setResults...CSVFilepath`` call (the reference writes them to
    the working directory by default, hpp:90-91), because at 10^6+ beams they
    would dominate the run.
    """

    def __init__(self, device: int = -1, tol: float = 1e-8, max_iter: int = 0, verbose: bool = False, **kw):
        self._ctx = Context(device=device, tol=tol, max_iter=max_iter, verbose=verbose, **kw)
        self._have_job = False
        self._nd_csv = ""
        self._ef_csv = ""
        self._results = SimulationResults()

    @property
    def context(self) -> Context:
        return self._ctx

    def setSimulation(self, simulationJob: SimulationJob):
        j = simulationJob
        self._have_job = False
        self._ctx.set_job(j.nodes, j.elements, j.elementalNormals, j.beamDimensions, j.beamMaterial,
                          j.fixedNodes, j.nodalForces)
        self._have_job = True

    def setResultsNodalDisplacementCSVFilepath(self, outputPath: str):
        self._nd_csv = outputPath

    def setResultsElementalForcesCSVFilepath(self, outputPath: str):
        self._ef_csv = outputPath

    def executeSimulation(self) -> SimulationResults:
        if not self._have_job:  # Expects(!simulationJob.isEmpty()) [REF src/beamsimulator.cpp:41]
            raise ContractViolation("precondition violated: empty simulation job")
        self._ctx.execute()
        if self._ef_csv:
            self._ctx.write_element_forces_csv(self._ef_csv)
        if self._nd_csv:
            self._ctx.write_displacements_csv(self._nd_csv)
        self._results = SimulationResults(self._ctx.displacements(), self._ctx.edge_forces())
        return self._results

    def getResults(self) -> SimulationResults:
        return self._results

    def getStats(self) -> dict:
        return self._ctx.stats()
