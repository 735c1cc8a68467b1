"""Writes the committed golden fixtures tests/golden/*.npz.

The reference cannot be built or imported in this environment (its numerical
core is the un-vendored threed-beam-fea library + Eigen; DESIGN.md "Oracle"),
so these vectors come from the CPU oracle after it has been pinned against the
reference's own golden data (tests/test_oracle_golden.py).  Displacements are
solved with scipy's SuperLU in the upstream Lagrange-multiplier (KKT) form, the
closest executable stand-in for Eigen::SparseLU, followed by two steps of
iterative refinement (the raw KKT factorisation alone carries ~1e-8 relative
round-off, see test_kkt_and_elimination_agree).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc          # noqa: E402
from meshbeamfea_b200 import synthetic as syn  # noqa: E402
from helpers import to_oracle, forces_list  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
cases = {
    "two_line_segments_screenshot": syn.two_line_segments(),
    "two_line_segments_axial": syn.two_line_segments(fixed=(1, 2), forces=((0, 0, 1000.0),)),
    "cantilever_3beams_rotZ90": syn.cantilever(3, *[c for c in syn.tester_rotation_cases() if c[2] == 3][0][3:5]),
    "planar_grid_6": syn.planar_grid(6),
    "cubic_lattice_4": syn.cubic_lattice(4),
    "gridshell_6": syn.gridshell(6),
    "random_mesh_seed1234_V60": syn.random_mesh(1234, V=60),
}
for name, job in cases.items():
    O = to_oracle(job)
    U, F = orc.solve(O, "kkt")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), nodes=O.nodes, elements=O.elements, normals=O.normals,
                        fixed=O.fixed, forces=np.array(forces_list(job), np.float64).reshape(-1, 3), dims=O.dims,
                        material=O.material, U=U, F=F, props=orc.beam_props(O.dims, O.material),
                        free_index=orc.free_map(O.nodes.shape[0], O.fixed))
    print(name, U.shape, F.shape)
