"""Shared test helpers: SimulationJob -> oracle Job, norms, the fixture writer."""
import numpy as np

from oracle import oracle as orc


def forces_list(job):
    f = job.nodalForces
    if isinstance(f, tuple):
        return [(int(a), int(b), float(c)) for a, b, c in zip(*f)]
    return [(f_.index, f_.dof, f_.magnitude) if hasattr(f_, "index") else tuple(f_) for f_ in f]


def to_oracle(job) -> orc.Job:
    return orc.Job(np.asarray(job.nodes, np.float64), np.asarray(job.elements, np.int32),
                   np.asarray(job.elementalNormals, np.float64), np.asarray(job.fixedNodes, np.int32),
                   forces_list(job), np.asarray(job.beamDimensions, np.float32), np.asarray(job.beamMaterial, np.float32))


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def block_rel(a, b):
    """max over 6x6 blocks of |a-b| / max|b_block|  (hard part 3)."""
    a = np.asarray(a).reshape(-1, 36); b = np.asarray(b).reshape(-1, 36)
    scale = np.abs(b).max(axis=1)
    scale[scale == 0] = 1.0
    return (np.abs(a - b).max(axis=1) / scale).max() if a.size else 0.0


PLY_TWO_SEGMENTS = """ply
format ascii 1.0
element vertex 3
property float x
property float y
property float z
element edge 2
property int vertex1
property int vertex2
property list uchar float edge_normal
property list uchar float beam_dimensions
property list uchar float beam_material {poissons ratio , young's modulus in GPascal}
end_header

0 0 0
1 0 0
2 0 0
0 1 3 0.0 1.0 0.0 2 0.01 0.01 2 0.3 800
1 2 3 0.0 1.0 0.0 2 0.01 0.01 2 0.3 800"""


def write_scenario(tmpdir, fixed, forces, ply_text=PLY_TWO_SEGMENTS, ply_name="TwoLineSegments.ply"):
    """A config.json in the reference's schema [REF src/utilities.hpp:50-95] next to its PLY."""
    import json
    import os
    ply = os.path.join(str(tmpdir), ply_name)
    with open(ply, "w") as f:
        f.write(ply_text)
    cfg = os.path.join(str(tmpdir), "config.json")
    with open(cfg, "w") as f:
        json.dump({"plyFilename": ply_name, "fixedVertices": list(fixed), "forces": [list(x) for x in forces]}, f)
    return cfg, ply
