"""meshbeamfea_b200 -- B200-native (sm_100a) replacement for the simulation hot
path of IasonManolas/MeshBeamFEA: element stiffness + rotation, deterministic
BSR assembly, constraint elimination, FP64 block-Jacobi PCG and beam end-force
recovery, behind the reference's own ``BeamSimulator`` API.

The package is a thin ctypes face over ``libmbfea.so`` (C ABI in
``include/mbfea.h``, CUDA in ``meshbeamfea_b200/csrc``).  No CPU fallback.
"""
from .simulator import (BeamDimensions, BeamMaterial, BeamSimulator, Context, ContractViolation, MbfeaError,
                        NodalForce, NotConverged, RangeError, SimulationJob, SimulationResults, SingularSystem,
                        host_contributors, host_hierarchy, host_renumber, host_spd_inverse, host_solver_format, plan_partition, read_scenario)
from . import synthetic

__all__ = ["BeamDimensions", "BeamMaterial", "BeamSimulator", "Context", "ContractViolation", "MbfeaError",
           "NodalForce", "NotConverged", "RangeError", "SimulationJob", "SimulationResults", "SingularSystem",
           "host_contributors", "host_hierarchy", "host_renumber", "host_spd_inverse", "host_solver_format", "plan_partition", "read_scenario", "synthetic"]
