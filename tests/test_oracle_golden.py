"""Pins the CPU oracle against every golden / known-answer vector the reference
carries for the hot path: the README screenshot, the closed forms
behind BeamSimulatorTester's 45 cases, the float section-property arithmetic,
and the committed fixtures under tests/golden/."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc
from meshbeamfea_b200 import synthetic as syn
from helpers import to_oracle, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))


def test_screenshot_golden_vector():
    # images/MeshBeamFEA_postSimulation.png: fixed {0}, force (2, dof 1, 100),
    # colour bar of Mz: min -200, max 100
    U, F = orc.solve(to_oracle(syn.two_line_segments()))
    assert np.allclose(F[5], [-200.0, 100.0, -100.0, 0.0], rtol=0, atol=1e-9)
    assert F[5].min() == pytest.approx(-200.0, abs=1e-9) and F[5].max() == pytest.approx(100.0, abs=1e-9)
    # float quirk: EIz = 666.6665649414062, not 666.666...
    assert U[2, 1] == pytest.approx(100 * 8 / (3 * 666.6665649414062), rel=1e-13)
    assert U[2, 1] == pytest.approx(0.40000006103516, abs=1e-13)
    assert U[2, 5] == pytest.approx(0.3000000457763742, rel=1e-12)
    assert np.all(U[0] == 0)


def test_axial_known_answer():
    # config.json-like scenario on the fixture: fixed {1,2}, force (0,0,1000)
    U, F = orc.solve(to_oracle(syn.two_line_segments(fixed=(1, 2), forces=((0, 0, 1000.0),))))
    assert U[0, 0] == pytest.approx(1.25e-05, rel=1e-13)
    assert np.allclose(F[0], [1000.0, -1000.0, 0.0, 0.0], atol=1e-9)


def test_float_section_properties():
    # [REF src/beamsimulator.cpp:173-188] evaluated in IEEE float
    p = orc.beam_props(np.array([[0.01, 0.01]], np.float32), np.array([[0.3, 800]], np.float32))[0]
    assert p[0] == 80000000.0
    assert p[1] == 666.6665649414062 and p[2] == 666.6665649414062
    assert p[1].hex() == "0x1.4d55520000000p+9"
    assert p[3] == 512.8204345703125
    q = orc.beam_props(np.array([[0.1, 0.2]], np.float32), np.array([[0.3, 210]], np.float32))[0]
    assert list(q) == [4200000256.0, 14000000.0, 3500000.0, 6730769.5]


@pytest.mark.parametrize("case", list(syn.tester_rotation_cases()), ids=lambda c: f"{c[0]}-{c[1]}-{c[2]}")
def test_tester_rotation_cases_closed_form(case):
    # [REF src/beamsimulatortester.cpp:8-41]: results must not depend on the rigid
    # rotation nor on the number of sub-beams; closed form P L^3 / 3 EIz etc.
    axis, step, nbeams, xa, ya = case
    job = syn.cantilever(nbeams, xa, ya)
    U, F = orc.solve(to_oracle(job))
    EIz = 14000000.0
    tip = U[-1]
    za = np.cross(xa, ya)
    assert tip[:3] @ ya == pytest.approx(1000.0 / (3 * EIz), rel=1e-9)
    assert tip[:3] @ ya == pytest.approx(2.380952380952381e-05, rel=1e-9)
    assert tip[3:] @ za == pytest.approx(3.571428571428572e-05, rel=1e-9)
    assert abs(tip[:3] @ xa) < 1e-15 and abs(tip[:3] @ za) < 1e-15
    # clamped-end reactions: shear 1000, moment P*L (sign convention K*u)
    assert F[1][0] == pytest.approx(-1000.0, rel=1e-9)
    assert F[5][0] == pytest.approx(-1000.0, rel=1e-9)
    assert F[5][-1] == pytest.approx(0.0, abs=1e-6)


def test_kkt_and_elimination_agree():
    # upstream imposes BCs with Lagrange multipliers (KKT + SparseLU); the product
    # eliminates.  For zero prescribed values both give the same u.
    for job in (syn.random_mesh(7, V=60), syn.planar_grid(7), syn.cubic_lattice(4)):
        O = to_oracle(job)
        U1, F1 = orc.solve(O, "eliminate")
        U2, F2 = orc.solve(O, "kkt")
        assert rel_l2(U1, U2) < 1e-10
        assert rel_l2(F1, F2) < 1e-10
        # the raw (unrefined) KKT factorisation -- what the reference's SparseLU
        # call amounts to -- carries ~1e-8 of its own round-off on these meshes
        U3, _ = orc.solve(O, "kkt", refine=0)
        assert rel_l2(U3, U1) < 1e-6


def test_element_matrix_properties():
    O = to_oracle(syn.random_mesh(11, V=40))
    Ke = orc.kelem_all(O)
    assert np.abs(Ke - Ke.transpose(0, 2, 1)).max() <= 1e-9 * np.abs(Ke).max()
    # rigid translation is in the null space of every element matrix
    t = np.zeros(12); t[[0, 6]] = 1.0
    assert np.abs(Ke @ t).max() <= 1e-9 * np.abs(Ke).max()


def test_full_assembly_matches_block_assembly():
    job = syn.cubic_lattice(4)
    O = to_oracle(job)
    K = orc.assemble_full(O).toarray()
    fm = orc.free_map(O.nodes.shape[0], O.fixed)
    rp, col = orc.bsr_pattern(O, fm)
    vals = orc.bsr_assemble(O, fm, rp, col)
    free = np.nonzero(fm >= 0)[0]
    for r in range(rp.size - 1):
        for k in range(rp[r], rp[r + 1]):
            gi, gj = free[r], free[col[k]]
            blk = K[6 * gi:6 * gi + 6, 6 * gj:6 * gj + 6]
            assert np.array_equal(blk, vals[k])  # same element-order sums, bit for bit
    # block pattern == {(v,v)} U {(i,j),(j,i) per edge} on free vertices
    pairs = {(r, int(col[k])) for r in range(rp.size - 1) for k in range(rp[r], rp[r + 1])}
    want = {(int(fm[v]), int(fm[v])) for v in free}
    for a, b in O.elements:
        if fm[a] >= 0 and fm[b] >= 0:
            want |= {(int(fm[a]), int(fm[b])), (int(fm[b]), int(fm[a]))}
    assert pairs == want


def test_pcg_oracle_matches_direct():
    job = syn.planar_grid(8)
    O = to_oracle(job)
    U, _ = orc.solve(O)
    fm = orc.free_map(O.nodes.shape[0], O.fixed)
    rp, col = orc.bsr_pattern(O, fm)
    vals = orc.bsr_assemble(O, fm, rp, col)
    f = orc.load_vector(O.nodes.shape[0], O.forces)
    free_dofs = (6 * np.nonzero(fm >= 0)[0][:, None] + np.arange(6)[None, :]).ravel()
    x, it, rel = orc.pcg_bj(rp, col, vals, f[free_dofs], tol=1e-12)
    assert rel <= 1e-12 and it > 0
    assert rel_l2(x, U.ravel()[free_dofs]) < 1e-8


def test_error_conventions():
    with pytest.raises(RuntimeError, match="Vertex index is outside of valid range."):
        orc.free_map(3, [1, 2, 3])          # [REF src/beamsimulator.cpp:134-137], config.json as shipped
    with pytest.raises(IndexError, match="Node index is out of valid range."):
        orc.load_vector(3, [(3, 0, 1.0)])   # [REF src/beamsimulator.cpp:164-166]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))), ids=os.path.basename)
def test_committed_golden_fixtures(path):
    g = np.load(path)
    O = orc.Job(g["nodes"], g["elements"], g["normals"], g["fixed"],
                [(int(a), int(b), float(c)) for a, b, c in g["forces"]], g["dims"], g["material"])
    U, F = orc.solve(O)
    assert rel_l2(U, g["U"]) < 1e-10
    assert rel_l2(F, g["F"]) < 1e-10
    assert np.array_equal(orc.beam_props(O.dims, O.material), g["props"])
    assert np.array_equal(orc.free_map(O.nodes.shape[0], O.fixed), g["free_index"])
