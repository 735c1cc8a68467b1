"""CPU-only tests of the host side of libmbfea.so: the library loads and exports
every symbol include/mbfea.h declares, validation/error conventions, the float
section-property arithmetic, the constraint map, the block pattern and the
vertex-block partition plan, and the scenario (config.json + PLY) reader.
No compute entry point is called here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import meshbeamfea_b200 as mb
from meshbeamfea_b200 import _lib as L
from meshbeamfea_b200 import synthetic as syn
from oracle import oracle as orc
from helpers import to_oracle, write_scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "mbfea.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbfea_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = header_functions()
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (mbfea_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(L.SIGNATURES), sorted(set(names) ^ set(L.SIGNATURES))
    lib = L.lib()   # also verifies sizeof(mbfea_options / mbfea_stats / mbfea_plan_sizes) against the ctypes mirrors
    sizes = (C.c_int64 * 3)()
    lib.mbfea_abi_sizes(sizes)
    assert tuple(sizes) == (C.sizeof(L.Options), C.sizeof(L.Stats), C.sizeof(L.PlanSizes))
    assert lib.mbfea_version() == 100
    assert lib.mbfea_status_string(2) == b"Vertex index is outside of valid range."
    assert lib.mbfea_status_string(3) == b"Node index is out of valid range."


def test_no_silent_cpu_fallback(built):
    """Without a GPU the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MbfeaError, match="no usable CUDA device"):
        mb.Context()
    with pytest.raises(mb.MbfeaError):
        mb.BeamSimulator()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under meshbeamfea_b200/ or
    include/ may import, link, load or call it."""
    banned = re.compile(r"import\s+oracle|from\s+oracle|libmbfea_oracle|\borc_[a-z]|oracle\.py|oracle/oracle")
    for top in ("meshbeamfea_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f)).read()
                    assert not banned.search(text), (dirpath, f)


def _validate(job, fixed=None, forces=None):
    lib = L.lib()
    nodes = np.asarray(job.nodes, np.float64); el = np.ascontiguousarray(np.asarray(job.elements, np.int32).T).ravel()
    dims = np.ascontiguousarray(job.beamDimensions, np.float32); mat = np.ascontiguousarray(job.beamMaterial, np.float32)
    fixed = np.ascontiguousarray(job.fixedNodes if fixed is None else fixed, np.int32)
    forces = [(f.index, f.dof, f.magnitude) for f in job.nodalForces] if forces is None else forces
    fn = np.array([f[0] for f in forces], np.int64); fd = np.array([f[1] for f in forces], np.int64)
    msg = C.create_string_buffer(256)
    rc = lib.mbfea_host_validate(nodes.shape[0], job.elements.shape[0], L.ptr(el, L._pi32), L.ptr(dims, L._pf),
                                 L.ptr(mat, L._pf), fixed.size, L.ptr(fixed, L._pi32), fn.size, L.ptr(fn, L._pi64),
                                 L.ptr(fd, L._pi64), msg, 256)
    return rc, msg.value.decode()


def test_validation_error_conventions(built):
    job = syn.two_line_segments()
    assert _validate(job) == (L.OK, "")
    # config.json as shipped fixes vertex 3 of a 3-vertex mesh [REF src/beamsimulator.cpp:134-137]
    assert _validate(job, fixed=[1, 2, 3]) == (L.ERANGE_FIXED, "Vertex index is outside of valid range.")
    assert _validate(job, fixed=[-1])[0] == L.ERANGE_FIXED
    # [REF src/beamsimulator.cpp:164-166]
    assert _validate(job, forces=[(3, 0, 1.0)]) == (L.ERANGE_FORCE, "Node index is out of valid range.")
    assert _validate(job, forces=[(0, 6, 1.0)])[0] == L.EPRECOND      # [REF src/utilities.hpp:77]
    bad = syn.two_line_segments(); bad.beamDimensions[1, 0] = 0.0     # [REF src/beam.hpp:14]
    assert _validate(bad)[0] == L.EPRECOND
    bad = syn.two_line_segments(); bad.beamMaterial[0, 0] = 0.51      # [REF src/beam.hpp:25]
    assert _validate(bad)[0] == L.EPRECOND
    bad = syn.two_line_segments(); bad.elements[0, 1] = 7
    assert _validate(bad)[0] == L.EPRECOND
    # the reference orders the checks elements -> fixed -> forces [REF src/beamsimulator.cpp:107-111]
    assert _validate(job, fixed=[9], forces=[(9, 0, 1.0)])[0] == L.ERANGE_FIXED
    with pytest.raises(mb.ContractViolation):
        mb.BeamDimensions(0.0, 1.0)
    with pytest.raises(mb.ContractViolation):
        mb.BeamMaterial(0.6, 210)
    assert (mb.BeamDimensions().b, mb.BeamDimensions().h) == (1.0, 1.0)           # [REF src/beam.hpp:16]
    assert (mb.BeamMaterial().poissonsRatio, mb.BeamMaterial().youngsModulusGPascal) == (0.3, 210.0)  # [:27]


def test_validation_of_a_large_job_reports_the_first_offending_beam(built):
    """The per-beam checks run in parallel chunks of >= 65536 beams; the message must still be the one a serial scan in
    the reference's order gives: the FIRST offending beam decides [REF src/beamsimulator.cpp:71-94, src/beam.hpp:14,25]."""
    job = syn.cubic_lattice(42)          # 215 k beams: several chunks
    E = job.elements.shape[0]
    F1 = [(1, 2, -5.0)]
    assert E > 3 * 65536 and _validate(job, forces=F1) == (L.OK, "")
    late, early = E - 5, 70000            # different chunks
    bad = syn.cubic_lattice(42)
    bad.beamMaterial = np.array(bad.beamMaterial, np.float32); bad.beamDimensions = np.array(bad.beamDimensions, np.float32)
    bad.beamMaterial[late, 0] = 0.9                                   # Poisson's ratio, late beam
    assert _validate(bad, forces=F1) == (L.EPRECOND, "precondition violated: Poisson's ratio outside [-1, 0.5]")
    bad.beamDimensions[early, 1] = -1.0                               # dimensions, earlier beam: wins
    assert _validate(bad, forces=F1) == (L.EPRECOND, "precondition violated: beam dimensions must be positive")
    bad.elements = np.array(bad.elements, np.int32); bad.elements[early - 1000, 0] = bad.elements[early - 1000, 1]
    assert _validate(bad, forces=F1) == (L.EPRECOND, "precondition violated: zero-length beam (both ends on one vertex)")
    bad.elements[5, 1] = job.nodes.shape[0]                           # out of range, first chunk: wins over everything
    assert _validate(bad, forces=F1) == (L.EPRECOND, "precondition violated: element node index out of range")


def test_host_props_bit_exact_vs_oracle(built):
    rng = np.random.default_rng(5)
    E = 20000
    dims = rng.uniform(0.001, 0.5, (E, 2)).astype(np.float32)
    mat = np.stack([rng.uniform(-1, 0.5, E), rng.uniform(1, 1000, E)], 1).astype(np.float32)
    out = np.empty((E, 4))
    assert L.lib().mbfea_host_props(E, L.ptr(dims, L._pf), L.ptr(mat, L._pf), L.ptr(out, L._pd)) == L.OK
    assert np.array_equal(out, orc.beam_props(dims, mat))
    one = np.empty((1, 4))
    L.lib().mbfea_host_props(1, L.ptr(np.array([[0.01, 0.01]], np.float32), L._pf),
                             L.ptr(np.array([[0.3, 800]], np.float32), L._pf), L.ptr(one, L._pd))
    assert list(one[0]) == [80000000.0, 666.6665649414062, 666.6665649414062, 512.8204345703125]


def test_host_free_map_bit_exact_vs_oracle(built):
    rng = np.random.default_rng(3)
    # (the last two sizes take the chunked, multi-threaded scan: several chunks of 65536 vertices; 300000 fixed entries
    #  also mark in parallel)
    for V, F in ((3, 1), (1000, 17), (50000, 7000), (400001, 90000), (700000, 300000)):
        fixed = rng.integers(0, V, F).astype(np.int32)  # duplicates on purpose
        fm = np.empty(V, np.int32); n = C.c_int64()
        assert L.lib().mbfea_host_free_map(V, F, L.ptr(fixed, L._pi32), L.ptr(fm, L._pi32), C.byref(n)) == L.OK
        ref = orc.free_map(V, fixed)
        assert np.array_equal(fm, ref) and n.value == (ref >= 0).sum()
    fm = np.empty(3, np.int32)
    assert L.lib().mbfea_host_free_map(3, 3, L.ptr(np.array([1, 2, 3], np.int32), L._pi32), L.ptr(fm, L._pi32), None) == L.ERANGE_FIXED


@pytest.mark.parametrize("job", [syn.two_line_segments(), syn.planar_grid(9), syn.cubic_lattice(5), syn.gridshell(7),
                                 syn.random_mesh(21, V=80)], ids=["2seg", "grid9", "lat5", "shell7", "rand80"])
def test_block_pattern_bit_exact_vs_oracle(built, job):
    O = to_oracle(job)
    V = O.nodes.shape[0]
    fm = orc.free_map(V, O.fixed)
    rp, col = orc.bsr_pattern(O, fm)
    p = mb.plan_partition(V, job.elements, job.fixedNodes, 0, 1)
    assert p["nb_global"] == (fm >= 0).sum() and p["row_begin"] == 0 and p["row_end"] == p["nb_global"]
    assert np.array_equal(p["rowptr"], rp) and np.array_equal(p["colidx"], col)
    assert p["ghost_global"].size == 0 and p["send_rows"].size == 0


def test_pattern_with_multi_edges_and_isolated_vertex(built):
    # two beams between the same pair of vertices share one block; a free vertex
    # without beams still owns an (empty) diagonal block
    elements = np.array([[0, 1], [1, 0], [1, 2]], np.int32)
    p = mb.plan_partition(4, elements, np.array([0], np.int32), 0, 1)
    assert p["nb_global"] == 3
    assert list(p["rowptr"]) == [0, 2, 4, 5] and list(p["colidx"]) == [0, 1, 0, 1, 2]


def test_batch_detection(built):
    """A job that is a batch of independent small meshes is recognised by its independent block-row
    ranges (a mesh may fall into several connected components once its fixed vertices go)."""
    jobs = [syn.random_mesh(300 + m, V=30 + (m % 4) * 10) for m in range(40)]
    merged = syn.merge_jobs(jobs)
    p = mb.plan_partition(merged.nodes.shape[0], merged.elements, merged.fixedNodes, 0, 1)
    assert p["n_components"] == 40
    assert p["max_component_rows"] == max(j.nodes.shape[0] - np.unique(j.fixedNodes).size for j in jobs)
    # one connected mesh, too few meshes, or a partitioned job: not a batch
    g = syn.planar_grid(12)
    assert mb.plan_partition(g.nodes.shape[0], g.elements, g.fixedNodes, 0, 1)["n_components"] == 0
    few = syn.merge_jobs(jobs[:8])
    assert mb.plan_partition(few.nodes.shape[0], few.elements, few.fixedNodes, 0, 1)["n_components"] == 0
    assert mb.plan_partition(merged.nodes.shape[0], merged.elements, merged.fixedNodes, 0, 2)["n_components"] == 0
    # interleaved vertex numbering of two meshes couples the ranges: falls back to the global solver
    a, b = syn.random_mesh(1, V=40), syn.random_mesh(2, V=40)
    two = syn.merge_jobs([a, b] * 20)
    perm = np.arange(two.nodes.shape[0]); perm[[3, 45]] = perm[[45, 3]]      # swap a vertex of mesh 0 with one of mesh 1
    inv = np.argsort(perm)
    el = inv[two.elements]
    fx = inv[two.fixedNodes]
    assert mb.plan_partition(two.nodes.shape[0], el, fx, 0, 1)["n_components"] == 39   # meshes 0 and 1 merge into one range


def _shuffle_vertices(job, seed):
    V = job.nodes.shape[0]
    perm = np.random.default_rng(seed).permutation(V)      # new id of old vertex v
    inv = np.argsort(perm)
    fn, fd, fm = job.nodalForces
    return mb.SimulationJob(job.nodes[inv], perm[job.elements].astype(np.int32), job.elementalNormals,
                            perm[job.fixedNodes].astype(np.int32), (perm[fn], fd, fm), job.beamDimensions,
                            job.beamMaterial)


def test_internal_renumbering(built):
    """A mesh whose vertex ids are in arbitrary order gets a breadth-first internal numbering
    (neighbouring rows close together); a mesh that already has locality keeps its order."""
    lat = syn.cubic_lattice(40)      # 62 400 block rows: more than the ~16 MB locality window of the heuristic
    V = lat.nodes.shape[0]
    perm, reordered = mb.host_renumber(V, lat.elements, lat.fixedNodes, 0)
    assert not reordered and np.array_equal(perm, np.arange(perm.size))
    shuf = _shuffle_vertices(lat, 3)
    perm, reordered = mb.host_renumber(V, shuf.elements, shuf.fixedNodes, 0)
    assert reordered and np.array_equal(np.sort(perm), np.arange(perm.size))        # a permutation
    fm = orc.free_map(V, shuf.fixedNodes)
    a, b = fm[shuf.elements[:, 0]], fm[shuf.elements[:, 1]]
    both = (a >= 0) & (b >= 0)
    span_before = np.abs(a[both] - b[both]).mean()
    span_after = np.abs(perm[a[both]] - perm[b[both]]).mean()
    assert span_after < 0.2 * span_before and span_after < 3 * 40 * 40            # ~ one BFS level
    small = _shuffle_vertices(syn.cubic_lattice(10), 4)                              # L2-resident anyway: left alone
    assert not mb.host_renumber(1000, small.elements, small.fixedNodes, 0)[1]
    assert not mb.host_renumber(V, shuf.elements, shuf.fixedNodes, 1)[1]           # never
    assert mb.host_renumber(V, lat.elements, lat.fixedNodes, 2)[1]                # always
    # a batch of independent meshes stays a sequence of contiguous row ranges after renumbering
    jobs = [syn.random_mesh(700 + m, V=30) for m in range(12)]
    merged = syn.merge_jobs(jobs)
    perm, reordered = mb.host_renumber(merged.nodes.shape[0], merged.elements, merged.fixedNodes, 2)
    assert reordered
    fm = orc.free_map(merged.nodes.shape[0], merged.fixedNodes)
    owner = np.repeat(np.arange(12), 30)[fm >= 0]          # mesh of each canonical free row
    by_internal = owner[np.argsort(perm)]
    assert np.all(np.diff(by_internal) >= 0)


@pytest.mark.parametrize("world,n", [(2, 7), (3, 7), (8, 7), (3, 40)], ids=["w2", "w3", "w8", "w3_threaded"])
def test_partition_plan_is_consistent(built, world, n):
    """(n = 40: large enough for the multi-threaded host passes and the owned-beam filter of world > 1.)"""
    job = syn.cubic_lattice(n)
    V = job.nodes.shape[0]
    full = mb.plan_partition(V, job.elements, job.fixedNodes, 0, 1)
    plans = [mb.plan_partition(V, job.elements, job.fixedNodes, r, world) for r in range(world)]
    nbg = full["nb_global"]
    assert plans[0]["row_begin"] == 0 and plans[-1]["row_end"] == nbg
    for r, p in enumerate(plans):
        assert p["row_begin"] == nbg * r // world and p["row_end"] == nbg * (r + 1) // world
        if r:
            assert p["row_begin"] == plans[r - 1]["row_end"]
        nb = p["row_end"] - p["row_begin"]
        # local pattern, mapped back to global columns, equals the rows of the global pattern
        lo, hi = full["rowptr"][p["row_begin"]], full["rowptr"][p["row_end"]]
        assert np.array_equal(p["rowptr"], full["rowptr"][p["row_begin"]:p["row_end"] + 1] - lo)
        gcol = np.where(p["colidx"] < nb, p["colidx"] + p["row_begin"],
                        p["ghost_global"][np.maximum(p["colidx"] - nb, 0)] if p["ghost_global"].size else 0)
        assert np.array_equal(gcol, full["colidx"][lo:hi])
        # ghosts ascending, owned by someone else, owner consistent with the row split
        gg = p["ghost_global"]
        assert np.all(np.diff(gg) > 0)
        assert np.all((gg < p["row_begin"]) | (gg >= p["row_end"]))
        for g, o in zip(gg, p["ghost_owner"]):
            assert plans[o]["row_begin"] <= g < plans[o]["row_end"]
    # what q receives from r (in order) is exactly what r sends to q (in order)
    for r, p in enumerate(plans):
        for q, pq in enumerate(plans):
            if q == r:
                continue
            sent = p["send_rows"][p["send_dest"] == q] + p["row_begin"]
            recv = pq["ghost_global"][pq["ghost_owner"] == r]
            assert np.array_equal(sent, recv)


@pytest.mark.parametrize("world,n", [(2, 9), (3, 40)], ids=["w2", "w3_threaded"])
def test_partitioned_assembly_plan_equals_the_single_rank_plan(built, world, n):
    """Every rank's contributor lists (which beams' sub-blocks are summed into a block, in which order)
    are the single-rank lists of the same rows: K rows are bit-identical for any number of GPUs.  The
    single-rank lists are ascending in the beam index = the order setFromTriplets sums duplicates in."""
    job = syn.cubic_lattice(n)
    V = job.nodes.shape[0]
    full = mb.plan_partition(V, job.elements, job.fixedNodes, 0, 1)
    fptr, fcon = mb.host_contributors(V, job.elements, job.fixedNodes)
    assert fptr[-1] == fcon.size and fptr.size == full["colidx"].size + 1
    seg = np.repeat(np.arange(fptr.size - 1), np.diff(fptr))
    same_block = seg[1:] == seg[:-1]
    assert np.all((fcon[1:] >> 2)[same_block] >= (fcon[:-1] >> 2)[same_block])          # ascending beams per block
    for r in range(world):
        p = mb.plan_partition(V, job.elements, job.fixedNodes, r, world)
        cptr, con = mb.host_contributors(V, job.elements, job.fixedNodes, r, world)
        k0, k1 = full["rowptr"][p["row_begin"]], full["rowptr"][p["row_end"]]
        assert np.array_equal(cptr - cptr[0], fptr[k0:k1 + 1] - fptr[k0])
        assert np.array_equal(con, fcon[fptr[k0]:fptr[k1]])


def test_scenario_reader_fixture_and_error_path(built, tmp_path):
    cfg, ply = write_scenario(tmp_path, [0], [[2, 1, 100]])
    job = mb.read_scenario(cfg)
    ref = syn.two_line_segments()
    assert np.array_equal(job.nodes, ref.nodes) and np.array_equal(job.elements, ref.elements)
    assert np.array_equal(job.elementalNormals, ref.elementalNormals)
    assert np.array_equal(job.beamDimensions, ref.beamDimensions) and np.array_equal(job.beamMaterial, ref.beamMaterial)
    assert list(job.fixedNodes) == [0]
    assert [(f.index, f.dof, f.magnitude) for f in job.nodalForces] == [(2, 1, 100.0)]
    # absolute plyFilename that does not exist (config.json as shipped) -> IO error, override works
    cfg2 = os.path.join(str(tmp_path), "shipped.json")
    with open(cfg2, "w") as f:
        f.write('{\n\t"plyFilename": "/home/iason/Coding/spiralEdgeMesh.ply",\n\t"fixedVertices": [1,2,3],\n\t"forces": [[0,0,1000]]\n}\n')
    with pytest.raises(OSError):
        mb.read_scenario(cfg2)
    job = mb.read_scenario(cfg2, ply)
    assert list(job.fixedNodes) == [1, 2, 3]  # vertex 3 is out of range: set_job must raise (GPU test)
    # Expects(dof in [0,6) && index >= 0 && magnitude >= 0) [REF src/utilities.hpp:77]
    cfg3, _ = write_scenario(tmp_path, [0], [[2, 1, -5]])
    with pytest.raises(mb.ContractViolation):
        mb.read_scenario(cfg3)
    cfg4, _ = write_scenario(tmp_path, [0], [[2, 6, 5]])
    with pytest.raises(mb.ContractViolation):
        mb.read_scenario(cfg4)


def test_scenario_reader_rejects_corrupt_files(built, tmp_path):
    """Hostile or truncated inputs end in an IO error, never in a huge allocation, a wrapped index or a crash
    (the readers were fuzzed under AddressSanitizer / UBSan; these are the cases that needed a guard)."""
    from helpers import PLY_TWO_SEGMENTS
    bad_plys = {
        "huge_count": PLY_TWO_SEGMENTS.replace("element vertex 3", "element vertex 99999999999999"),
        "negative_count": PLY_TWO_SEGMENTS.replace("element edge 2", "element edge -2"),
        "no_count": PLY_TWO_SEGMENTS.replace("element edge 2", "element edge"),
        "huge_list": PLY_TWO_SEGMENTS.replace("0 1 3 0.0 1.0 0.0", "0 1 4000000000 0.0 1.0 0.0"),
        "index_wraps_int32": PLY_TWO_SEGMENTS.replace("1 2 3 0.0", "4294967297 2 3 0.0"),
        "truncated": PLY_TWO_SEGMENTS[:-20],
        "bad_number": PLY_TWO_SEGMENTS.replace("2 0 0", "2 zero 0"),
    }
    for name, text in bad_plys.items():
        cfg, _ = write_scenario(tmp_path, [0], [[2, 1, 100]], ply_text=text, ply_name=name + ".ply")
        with pytest.raises(OSError):
            mb.read_scenario(cfg)
    _, ply = write_scenario(tmp_path, [0], [[2, 1, 100]])
    bad_json = {
        "deep_nesting": "[" * 100000,
        "fixed_not_integer": '{"plyFilename": "%s", "fixedVertices": [1e300], "forces": []}' % ply,
        "fixed_nan_like": '{"plyFilename": "%s", "fixedVertices": ["0"], "forces": []}' % ply,
        "force_not_numbers": '{"plyFilename": "%s", "fixedVertices": [0], "forces": [["a", 1, 2]]}' % ply,
        "force_node_huge": '{"plyFilename": "%s", "fixedVertices": [0], "forces": [[1e30, 1, 2]]}' % ply,
        "not_an_object": "[1, 2, 3]",
        "unterminated": '{"plyFilename": "%s", "fixedVertices": [0' % ply,
    }
    for name, text in bad_json.items():
        path = os.path.join(str(tmp_path), name + ".json")
        with open(path, "w") as f:
            f.write(text)
        with pytest.raises(OSError):
            mb.read_scenario(path)


def test_ply_reader_defaults_and_double_types(built, tmp_path):
    ply = """ply
format ascii 1.0
comment no custom attributes: defaults of VCGEdgeMesh::setDefaultAttributes
element vertex 2
property double x
property double y
property double z
element edge 1
property int vertex1
property int vertex2
end_header
0.1 0 0
0.30000000000000004 0 0
0 1
"""
    cfg, _ = write_scenario(tmp_path, [0], [[1, 2, 1.5]], ply_text=ply, ply_name="plain.ply")
    job = mb.read_scenario(cfg)
    assert job.nodes[0, 0] == 0.1 and job.nodes[1, 0] == 0.30000000000000004
    assert np.array_equal(job.elementalNormals, [[0, 0, 1]])          # [REF src/edgemesh.cpp:226-232]
    assert np.array_equal(job.beamDimensions, np.array([[0.01, 0.01]], np.float32))
    assert np.array_equal(job.beamMaterial, np.array([[0.3, 210]], np.float32))


def test_synthetic_configs_have_the_surveyed_sizes():
    g = syn.planar_grid(100)
    assert g.nodes.shape[0] == 10000 and g.elements.shape[0] == 19800 and g.fixedNodes.size == 100
    l = syn.cubic_lattice(10)
    assert l.elements.shape[0] == 3 * 10 * 10 * 9 and l.fixedNodes.size == 100
    # per vertex: +i (normal y), +j (normal z), +k (normal x)
    assert list(l.elements[0]) == [0, 100] and list(l.elementalNormals[0]) == [0, 1, 0]
    assert list(l.elements[1]) == [0, 10] and list(l.elementalNormals[1]) == [0, 0, 1]
    assert list(l.elements[2]) == [0, 1] and list(l.elementalNormals[2]) == [1, 0, 0]
    s = syn.gridshell(12)
    t = s.nodes[s.elements[:, 1]] - s.nodes[s.elements[:, 0]]
    assert np.abs((t * s.elementalNormals).sum(1)).max() < 1e-14      # normals exactly perpendicular
    assert np.allclose(np.linalg.norm(s.elementalNormals, axis=1), 1.0, atol=1e-15)
    r = syn.random_mesh(1234)
    assert r.nodes.shape[0] == 200 and 450 <= r.elements.shape[0] <= 700
    cases = list(syn.tester_rotation_cases())
    assert len(cases) == 45                                            # [REF src/beamsimulatortester.cpp:21-40]


def test_cpp_dropin_header_builds_and_fails_loudly_without_gpu(built):
    """The C++17 shim (same type/method names as the reference's beamsimulator.hpp)
    compiles and links against libmbfea.so; without a GPU it must throw, not compute."""
    exe = os.path.join(ROOT, "meshbeamfea_b200", "cpp", "tester_main")
    assert os.path.exists(exe)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no usable CUDA device" in (r.stderr + r.stdout)


def test_bench_reference_arm_runs_on_cpu(built):
    """bench.py --impl reference: the CPU oracle timed on this host, one JSON line with the contract keys."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "lattice8",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "beams/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"] == "lattice8"


def _shuffled(job, seed):
    V = job.nodes.shape[0]
    vperm = np.random.default_rng(seed).permutation(V)
    inv = np.argsort(vperm)
    return job.nodes[inv], vperm[job.elements].astype(np.int32), vperm[job.fixedNodes].astype(np.int32)


def _check_solver_format(fmt, elements, fixed, V):
    """Every off-diagonal pair of the reduced pattern is stored exactly once, its other direction is a
    transposed reference to exactly that stored block, and the tiles hold their longest row."""
    nb = fmt["n_block_rows"]
    isfixed = np.zeros(V, bool); isfixed[fixed] = True
    fc = np.full(V, -1, np.int64); fc[~isfixed] = np.arange(nb)
    a, b = fc[elements[:, 0]], fc[elements[:, 1]]
    m = (a >= 0) & (b >= 0) & (a != b)
    pa, pb = fmt["perm"][a[m]].astype(np.int64), fmt["perm"][b[m]].astype(np.int64)
    pairs = np.unique(np.minimum(pa, pb) * nb + np.maximum(pa, pb))
    uptr, ucol, lptr, lcol, lofs, tb = (fmt[k].astype(np.int64) for k in ("uptr", "ucol", "lptr", "lcol", "lofs", "tile_base"))
    urow = np.repeat(np.arange(nb), np.diff(uptr)); lrow = np.repeat(np.arange(nb), np.diff(lptr))
    assert fmt["n_stored"] == fmt["n_refs"] == pairs.size and fmt["n_offdiag"] == 2 * pairs.size
    assert np.array_equal(np.sort(np.minimum(urow, ucol) * nb + np.maximum(urow, ucol)), pairs)     # stored once
    assert np.array_equal(np.sort(lrow * nb + lcol), np.sort(ucol * nb + urow))                      # refs = transposes
    # a reference points at the first 16-byte chunk of scalar row 0 of the block stored in row lcol
    slot = np.arange(ucol.size) - uptr[urow]
    first_chunk = tb[urow // 32] + slot * 576 + (urow % 32) * 6
    where = {(int(r), int(c)): int(u) for r, c, u in zip(urow, ucol, first_chunk)}
    assert all(where[(int(c), int(r))] == int(o) for r, c, o in zip(lrow, lcol, lofs))
    rows_len = np.diff(uptr)
    padded = np.zeros((nb + 31) // 32 * 32, np.int64); padded[:nb] = rows_len
    assert np.array_equal(np.diff(tb), padded.reshape(-1, 32).max(1) * 576)
    assert fmt["value_units"] == tb[-1] and fmt["longest_row"] == rows_len.max()
    return tb[-1] / (576.0 / 32 * pairs.size)    # slots per stored block: 1.0 = no padding


def test_solver_format_upper_rule_on_regular_numbering():
    lat = syn.cubic_lattice(12)
    fmt = mb.host_solver_format(lat.nodes, lat.elements, lat.fixedNodes)
    assert not fmt["renumbered"] and not fmt["balanced"] and fmt["longest_row"] == 3
    assert np.all(fmt["ucol"] > np.repeat(np.arange(fmt["n_block_rows"]), np.diff(fmt["uptr"])))   # upper blocks only
    assert _check_solver_format(fmt, lat.elements, lat.fixedNodes, lat.nodes.shape[0]) < 1.10   # face rows store fewer than 3


@pytest.mark.parametrize("mode", [2, 3], ids=["breadth_first", "morton"])
def test_solver_format_balanced_on_irregular_numbering(mode):
    """Shuffled vertex ids: internal renumbering (either kind) + pairs oriented towards the emptier row
    keep the interleaved tiles almost free of padding."""
    lat = syn.cubic_lattice(14)
    nodes, elements, fixed = _shuffled(lat, 3)
    fmt = mb.host_solver_format(nodes, elements, fixed, mode=mode)
    assert fmt["renumbered"] and np.array_equal(np.sort(fmt["perm"]), np.arange(fmt["n_block_rows"]))
    pad = _check_solver_format(fmt, elements, fixed, nodes.shape[0])
    if mode == 2:
        assert fmt["balanced"] and fmt["longest_row"] <= 4 and pad < 1.12
    else:   # the Z-order is monotone along every axis of a lattice: three upper neighbours per row, nothing to balance
        assert fmt["longest_row"] == 3 and pad < 1.12
    shell = syn.gridshell(40)
    nodes, elements, fixed = _shuffled(shell, 4)
    fmt = mb.host_solver_format(nodes, elements, fixed, mode=mode)
    assert fmt["balanced"]
    assert _check_solver_format(fmt, elements, fixed, nodes.shape[0]) < 1.25   # upper rule: 1.36-1.45 on this mesh


@pytest.mark.parametrize("mode", [2, 3], ids=["breadth_first", "morton"])
def test_solver_format_on_every_rank_of_a_partition(built, mode):
    """Several ranks, irregular numbering: a rank's rows list ghost columns of lower ranks FIRST (global column
    order), which the pair orientation must skip.  Per rank: every local pair is stored exactly once and
    referenced from its other row, ghost columns are always stored, references point at the stored block."""
    nodes, elements, fixed = _shuffled(syn.gridshell(60), 11)   # balanced orientation on ranks 1 and 2 in both modes
    V = nodes.shape[0]
    world = 3
    saw_balanced = False
    for rank in range(world):
        fmt = mb.host_solver_format(nodes, elements, fixed, mode=mode, rank=rank, world=world)
        nb = fmt["n_block_rows"]
        uptr, ucol, lptr, lcol, lofs, tb = (fmt[k].astype(np.int64) for k in ("uptr", "ucol", "lptr", "lcol", "lofs", "tile_base"))
        urow = np.repeat(np.arange(nb), np.diff(uptr)); lrow = np.repeat(np.arange(nb), np.diff(lptr))
        # the rank's rows of the (internally renumbered) reduced pattern, from the beams
        isfixed = np.zeros(V, bool); isfixed[fixed] = True
        fc = np.full(V, -1, np.int64); fc[~isfixed] = np.arange((~isfixed).sum())
        nbg = fmt["perm"].size
        r0 = nbg * rank // world
        assert nb == nbg * (rank + 1) // world - r0
        a, b = fc[elements[:, 0]], fc[elements[:, 1]]
        m = (a >= 0) & (b >= 0) & (a != b)
        ga, gb = fmt["perm"][a[m]].astype(np.int64), fmt["perm"][b[m]].astype(np.int64)
        own_a, own_b = (ga >= r0) & (ga < r0 + nb), (gb >= r0) & (gb < r0 + nb)
        local = own_a & own_b
        pairs = np.unique(np.minimum(ga[local], gb[local]) * nbg + np.maximum(ga[local], gb[local]))
        ghost_entries = np.unique(np.concatenate([ga[own_a & ~own_b] * nbg + gb[own_a & ~own_b],
                                                  gb[own_b & ~own_a] * nbg + ga[own_b & ~own_a]]))
        is_ghost = ucol >= nb
        assert is_ghost.sum() == ghost_entries.size                       # every (owned row, foreign column) stored here
        lu, lc = urow[~is_ghost] + r0, ucol[~is_ghost] + r0
        assert np.array_equal(np.sort(np.minimum(lu, lc) * nbg + np.maximum(lu, lc)), pairs)   # each local pair once
        assert np.array_equal(np.sort((lrow + r0) * nbg + lcol + r0), np.sort(lc * nbg + lu))    # refs = their transposes
        slot = np.arange(ucol.size) - uptr[urow]
        first_chunk = tb[urow // 32] + slot * 576 + (urow % 32) * 6
        where = {(int(r), int(c)): int(u) for r, c, u in zip(urow, ucol, first_chunk)}
        assert all(where[(int(c), int(r))] == int(o) for r, c, o in zip(lrow, lcol, lofs))
        saw_balanced |= bool(fmt["balanced"]) and rank > 0
    assert saw_balanced


@pytest.mark.parametrize("mesh", ["lattice", "gridshell"])
def test_aggregate_hierarchy_matches_galerkin_products(built, mesh):
    """Host half of the multilevel preconditioner: for every level the coarse pattern
    is the pattern of P^T A P and the Galerkin contributor lists reproduce its values, with
    P = rigid-body modes of the brick aggregates (built independently with scipy from agg / off)."""
    import scipy.sparse as sp
    job = syn.cubic_lattice(12) if mesh == "lattice" else syn.gridshell(30)
    O = to_oracle(job)
    V = O.nodes.shape[0]
    fm = orc.free_map(V, O.fixed)
    rp, col = orc.bsr_pattern(O, fm)
    vals = orc.bsr_assemble(O, fm, rp, col)
    nb = rp.size - 1
    H = mb.host_hierarchy(job.nodes, job.elements, job.fixedNodes, ratio=2.0, max_levels=6, coarsest_rows=8)
    assert len(H) >= 2 and H[0]["n_fine"] == nb
    # the fine pattern the contributor lists of level 0 index into is the canonical reduced pattern
    full = mb.plan_partition(V, job.elements, job.fixedNodes, 0, 1)
    assert np.array_equal(full["rowptr"], rp) and np.array_equal(full["colidx"], col)
    A = sp.bsr_matrix((vals.reshape(-1, 6, 6), col, rp), shape=(6 * nb, 6 * nb))
    pos = O.nodes[fm >= 0]
    blocks, bptr, bcol = vals.reshape(-1, 6, 6), rp, col
    for lvl, L in enumerate(H):
        nf, nc = L["n_fine"], L["n_coarse"]
        assert nc < nf and (lvl == 0 or nf == H[lvl - 1]["n_coarse"])
        agg, off = L["agg"], L["off"]
        # centroids and offsets
        cent = np.zeros((nc, 3)); np.add.at(cent, agg, pos); cent /= np.bincount(agg, minlength=nc)[:, None]
        assert np.allclose(off, pos - cent[agg], rtol=0, atol=1e-12)
        # B(d) = [[I, -[d]x], [0, I]]
        B = np.tile(np.eye(6), (nf, 1, 1))
        x, y, z = off[:, 0], off[:, 1], off[:, 2]
        B[:, 0, 4], B[:, 0, 5], B[:, 1, 3], B[:, 1, 5], B[:, 2, 3], B[:, 2, 4] = z, -y, -z, x, y, -x
        P = sp.bsr_matrix((B, agg, np.arange(nf + 1)), shape=(6 * nf, 6 * nc))
        Ac = (P.T @ A @ P).tobsr(blocksize=(6, 6)); Ac.sort_indices()
        # pattern: structural (a coarse block may be numerically zero), so compare with the aggregate graph
        frow = np.repeat(np.arange(nf), np.diff(bptr))
        want = np.unique(np.concatenate([agg[frow].astype(np.int64) * nc + agg[bcol], np.arange(nc) * (nc + 1)]))
        crow = np.repeat(np.arange(nc), np.diff(L["rowptr"]))
        assert np.array_equal(crow.astype(np.int64) * nc + L["colidx"], want)
        # contributor lists: every fine block exactly once, ascending per coarse block, values = Galerkin product
        gp, gs = L["gal_ptr"], L["gal_src"]
        assert np.array_equal(np.sort(gs), np.arange(bcol.size))
        seg = np.repeat(np.arange(gp.size - 1), np.diff(gp))
        assert np.all(np.diff(gs)[seg[1:] == seg[:-1]] > 0)
        assert np.array_equal(agg[frow[gs]], crow[seg]) and np.array_equal(agg[bcol[gs]], L["colidx"][seg])
        contrib = np.einsum("kji,kjl,klm->kim", B[frow[gs]], blocks[gs], B[bcol[gs]])
        mine = np.zeros((gp.size - 1, 6, 6)); np.add.at(mine, seg, contrib)
        ref = np.zeros_like(mine)
        Ad = Ac.tocsr()
        for k in range(gp.size - 1):
            I, J = crow[k], L["colidx"][k]
            ref[k] = Ad[6 * I:6 * I + 6, 6 * J:6 * J + 6].toarray()
        assert np.abs(mine - ref).max() <= 1e-9 * np.abs(ref).max()
        # next level works on the coarse matrix
        A, pos, blocks, bptr, bcol = sp.bsr_matrix((mine, L["colidx"], L["rowptr"]), shape=(6 * nc, 6 * nc)), cent, mine, L["rowptr"], L["colidx"]


def test_dense_spd_inverse(built):
    """Coarsest-level solve of the multilevel preconditioner: blocked Cholesky + explicit inverse on the host."""
    rng = np.random.default_rng(2)
    for n in (1, 6, 47, 48, 49, 300, 768):
        M = rng.standard_normal((n, n))
        A = M @ M.T + n * np.eye(n)
        Ai = mb.host_spd_inverse(A)
        assert np.abs(Ai @ A - np.eye(n)).max() < 1e-10
        assert np.abs(Ai - Ai.T).max() < 1e-12 * np.abs(Ai).max() + 1e-15
        assert np.abs(Ai - np.linalg.inv(A)).max() < 1e-10 * np.abs(Ai).max()
    with pytest.raises(mb.SingularSystem):
        mb.host_spd_inverse(np.array([[1.0, 2.0], [2.0, 1.0]]))          # indefinite
    with pytest.raises(mb.SingularSystem):
        mb.host_spd_inverse(np.zeros((3, 3)))


def test_multilevel_transfer_math(tmp_path):
    """The rigid-body transfer arithmetic the multilevel kernels share with the host:
    B(d) c, B(d)^T r, entries of B(d) and the Galerkin block entry, against explicit 6x6 matrices."""
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(str(tmp_path), "mlmath")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(here, "..", "meshbeamfea_b200", "csrc"),
                    os.path.join(here, "multilevel_math_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr


def _ml_image(job, world, smooth_from=0):
    """(header dict, arrays per level) of the hierarchy image rank 0 broadcasts on `world` ranks."""
    import ctypes as C
    from meshbeamfea_b200 import _lib as L
    lib = L.lib()
    nodes = np.ascontiguousarray(np.asarray(job.nodes, np.float64).T.ravel())
    el = np.ascontiguousarray(np.asarray(job.elements, np.int32).T.ravel())
    fx = np.ascontiguousarray(job.fixedNodes, np.int32)
    n = C.c_int64(0)
    args = (job.nodes.shape[0], job.elements.shape[0], L.ptr(nodes, L._pd), L.ptr(el, L._pi32), fx.size, L.ptr(fx, L._pi32), world, smooth_from)
    assert lib.mbfea_host_ml_image(*args, C.byref(n), None) == 0
    buf = np.zeros(n.value, np.uint8)
    assert lib.mbfea_host_ml_image(*args, C.byref(n), buf.ctypes.data_as(C.c_void_p)) == 0
    w = buf[:8 * 8].view(np.int64)
    hd = dict(magic=int(w[0]), levels=int(w[1]), world=int(w[2]), group0=int(w[3]), nbg=int(w[4]), fine_nnz=int(w[5]), total=int(w[6]))
    per_rank = buf[64:64 + 5 * 17 * 8].view(np.int64).reshape(5, 17)
    for k, name in enumerate(("own_first", "agg_begin", "blk_begin", "aggptr_at", "galptr_at")):
        hd[name] = per_rank[k, :world + 1].copy()
    base = 64 + 5 * 17 * 8
    levels = []
    for l in range(hd["levels"]):
        rec = buf[base + l * (4 + 24) * 8: base + (l + 1) * (4 + 24) * 8].view(np.int64)
        off, cnt = rec[4:16], rec[16:28]
        is_f64 = {1}   # array 1 of every level is the offsets (double), the rest are int32
        arrs = []
        for k in range(12):
            dt = np.float64 if k in is_f64 else np.int32
            arrs.append(buf[off[k]: off[k] + cnt[k] * np.dtype(dt).itemsize].view(dt) if cnt[k] else np.zeros(0, dt))
        levels.append(dict(n_fine=int(rec[0]), n_coarse=int(rec[1]), smoothed=int(rec[2]), arrays=arrs))
    return hd, levels


@pytest.mark.parametrize("world", [2, 3, 8])
def test_multilevel_hierarchy_image_is_sliceable_by_rank(world):
    """Several ranks: rank 0 builds the multilevel hierarchy and broadcasts one flat image; every rank slices its share of level
    0 out of it on the device.  That only works if fine-level aggregates never straddle a rank and are numbered rank-major,
    and if the member / contributor lists of a rank's aggregates and level-1 block rows reference its own rows and blocks
    only -- checked here on the image itself, against the partition of mbfea_plan_partition."""
    job = syn.cubic_lattice(9)
    hd, lv = _ml_image(job, world)
    assert hd["magic"] == 0x6d6c696d67303031 and hd["world"] == world and hd["levels"] >= 1
    nbg = hd["nbg"]
    agg, off, agg_ptr, agg_rows, rp1, ci1, diag1, gal_ptr, gal_src = lv[0]["arrays"][:9]
    assert agg.size == nbg and off.size == 3 * nbg and lv[0]["n_fine"] == nbg
    bounds = [nbg * q // world for q in range(world + 1)]
    for q in range(world):
        plan = mb.plan_partition(job.nodes.shape[0], job.elements, job.fixedNodes, q, world)
        assert (plan["row_begin"], plan["row_end"]) == (bounds[q], bounds[q + 1])
        # the image's block offsets are those of the rank's own local pattern (same block order, see build_plan)
        assert hd["own_first"][q + 1] - hd["own_first"][q] == plan["rowptr"][-1]
        a0, a1 = hd["agg_begin"][q], hd["agg_begin"][q + 1]
        mine = agg[bounds[q]:bounds[q + 1]]
        assert mine.min() >= a0 and mine.max() < a1                       # rank-major, no straddling
        members = agg_rows[agg_ptr[a0]:agg_ptr[a1]]
        assert members.min() >= bounds[q] and members.max() < bounds[q + 1]
        assert hd["aggptr_at"][q] == agg_ptr[a0] and hd["blk_begin"][q] == rp1[a0]
        b0, b1 = hd["blk_begin"][q], hd["blk_begin"][q + 1]
        src = gal_src[gal_ptr[b0]:gal_ptr[b1]]
        assert hd["galptr_at"][q] == gal_ptr[b0]
        assert src.min() >= hd["own_first"][q] and src.max() < hd["own_first"][q + 1]   # fine blocks of the rank's own rows only
    assert hd["agg_begin"][world] == lv[0]["n_coarse"] and gal_src.size == hd["fine_nnz"]
    assert np.array_equal(np.sort(gal_src), np.arange(hd["fine_nnz"]))   # every fine block contributes exactly once
    # diagonal positions of the level-1 pattern
    assert all(ci1[diag1[i]] == i for i in range(lv[0]["n_coarse"]))
    # the one-rank image of the same mesh has fewer (unsplit) aggregates
    hd1, lv1 = _ml_image(job, 1)
    assert lv1[0]["n_coarse"] <= lv[0]["n_coarse"]
