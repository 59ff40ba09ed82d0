"""Public names of the package (what ``tdgl/__init__.py`` exports for this path)."""
from .device import Device, Layer, Polygon, box, circle
from .engine import DeviceEngine, StepFailed
from .mesh import EdgeMesh, Mesh, make_film_mesh
from .options import SolverOptions, SolverOptionsError, SparseSolver
from .solution import DynamicsData, Solution, TDGLData
from .sharded import DistributedEngine, LocalShardGroup
from .solver import SolverResult, TDGLSolver, solve
from .sources import ConstantField, LinearRamp
from .synthetic import TerminalInfo

__all__ = [
    "Device", "Layer", "Polygon", "box", "circle", "DeviceEngine", "StepFailed", "EdgeMesh",
    "Mesh", "make_film_mesh", "SolverOptions", "SolverOptionsError", "SparseSolver",
    "DynamicsData", "Solution", "TDGLData", "SolverResult", "TDGLSolver", "solve",
    "TerminalInfo", "DistributedEngine", "LocalShardGroup", "ConstantField", "LinearRamp",
]
