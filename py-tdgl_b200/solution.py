"""``Solution`` / ``TDGLData`` / ``DynamicsData``: the output format of the path.

Field names, shapes and the per-save grouping follow the reference's HDF5 layout
(``DataHandler.save_time_step`` tdgl/solver/runner.py:155-183; ``TDGLData``
tdgl/solution/data.py:68-93; ``DynamicsData`` data.py:146-168, 385-425).  h5py/libhdf5 are
not in this image, so saves are kept in memory and ``Solution.to_npz`` writes the same
tree (keys ``data/<k>/psi`` ...) to one ``.npz``; post-processing (fluxoids, Biot-Savart,
plotting) is outside the hot path and not re-implemented.
"""

from __future__ import annotations

import dataclasses
from datetime import datetime
from typing import Any, Dict, List, Optional, Union

import numpy as np


@dataclasses.dataclass(eq=False)
class TDGLData:
    """Raw solver data at one saved step (reference data.py:68-93)."""

    step: int
    epsilon: np.ndarray
    psi: np.ndarray
    mu: np.ndarray
    applied_vector_potential: np.ndarray
    induced_vector_potential: np.ndarray
    supercurrent: np.ndarray
    normal_current: np.ndarray
    state: Dict[str, Any]


@dataclasses.dataclass(eq=False)
class DynamicsData:
    """Per-step scalars (reference data.py:146-228): ``time = cumsum(dt)``."""

    dt: np.ndarray
    time: np.ndarray = dataclasses.field(init=False)
    mu: Union[np.ndarray, None] = None
    theta: Union[np.ndarray, None] = None
    screening_iterations: Union[np.ndarray, None] = None

    def __post_init__(self):
        self.time = np.cumsum(self.dt)

    def time_slice(self, tmin: float = -np.inf, tmax: float = np.inf) -> np.ndarray:
        ts = self.time
        (indices,) = np.where((ts >= tmin) & (ts <= tmax))
        return indices

    def closest_time(self, time: float) -> int:
        return int(np.argmin(np.abs(self.time - time)))

    def voltage(self, i: int = 0, j: int = 1) -> np.ndarray:
        if self.mu is None:
            raise ValueError("No voltage data available.")
        if self.mu.shape[0] == 1:
            raise ValueError("The solution has only one probe point.")
        return self.mu[i] - self.mu[j]

    def phase_difference(self, i: int = 0, j: int = 1) -> np.ndarray:
        if self.theta is None:
            raise ValueError("No phase data available.")
        if self.theta.shape[0] == 1:
            raise ValueError("The solution has only one probe point.")
        return self.theta[i] - self.theta[j]

    def mean_voltage(self, i: int = 0, j: int = 1, tmin: float = -np.inf,
                     tmax: float = np.inf) -> float:
        if self.mu is None:
            raise ValueError("No voltage data available.")
        indices = self.time_slice(tmin, tmax)
        return float(np.average(self.voltage(i, j)[indices], weights=self.dt[indices]))


class SavedSteps:
    """In-memory stand-in for the reference's output file: ``fixed`` holds the arrays
    saved once at the root, ``groups[k]`` the k-th ``data/<k>`` group."""

    def __init__(self):
        self.fixed: Dict[str, np.ndarray] = {}
        self.groups: List[Dict[str, Any]] = []

    def save_fixed_values(self, fixed: Dict[str, np.ndarray]) -> None:
        for k, v in fixed.items():
            self.fixed[k] = np.array(v)

    def save_time_step(self, state: Dict[str, Any], data: Dict[str, np.ndarray],
                       running_state: Optional[Dict[str, np.ndarray]]) -> None:
        grp: Dict[str, Any] = {"attrs": dict(state, timestamp=datetime.now().isoformat())}
        for k, v in data.items():
            grp[k] = np.array(v)
        if running_state is not None:
            grp["running_state"] = {k: np.squeeze(np.array(v)) for k, v in running_state.items()}
        self.groups.append(grp)

    def dynamics(self) -> Optional[DynamicsData]:
        """Concatenate the running-state buffers exactly as ``DynamicsData.from_hdf5``
        does (data.py:396-425): unfilled entries have dt == 0 and are dropped."""
        dts, mus, thetas, sits = [], [], [], []
        for grp in self.groups:
            rs = grp.get("running_state")
            if rs is None:
                continue
            dts.append(np.atleast_1d(rs["dt"]))
            if "mu" in rs:
                mus.append(np.atleast_2d(rs["mu"]))
            if "theta" in rs:
                thetas.append(np.atleast_2d(rs["theta"]))
            if "screening_iterations" in rs:
                sits.append(np.atleast_1d(rs["screening_iterations"]))
        if not dts:
            return None
        dt = np.concatenate(dts)
        mask = dt > 0
        mu = np.concatenate(mus, axis=1)[..., mask] if mus else None
        theta = np.concatenate(thetas, axis=1)[..., mask] if thetas else None
        sit = np.concatenate(sits)[mask] if sits else None
        return DynamicsData(dt=dt[mask], mu=mu, theta=theta, screening_iterations=sit)


class Solution:
    """Results of a simulation (reference solution/solution.py:59-196, data access only)."""

    def __init__(self, *, device, options, saved: SavedSteps, applied_vector_potential=None,
                 terminal_currents=None, disorder_epsilon=None, total_seconds: float = 0.0,
                 path: Optional[str] = None, solver_stats: Optional[dict] = None, mesh=None):
        self.device = device
        self._mesh = mesh if mesh is not None else getattr(device, "mesh", None)
        self.options = options
        self.path = path
        self.applied_vector_potential = applied_vector_potential
        self.terminal_currents = terminal_currents
        self.disorder_epsilon = disorder_epsilon
        self.total_seconds = total_seconds
        self.solver_stats = solver_stats or {}
        self._saved = saved
        self._time_created = datetime.now()
        self.data_range = (0, len(saved.groups) - 1)
        self.dynamics = saved.dynamics()
        self.tdgl_data: Optional[TDGLData] = None
        self._solve_step = -1
        self.load_tdgl_data(-1)

    def load_tdgl_data(self, solve_step: int = -1) -> None:
        step_min, step_max = self.data_range
        if solve_step == 0:
            step = step_min
        elif solve_step < 0:
            step = step_max + 1 + solve_step
        else:
            step = solve_step
        grp = self._saved.groups[step]

        def get(key):
            if key in self._saved.fixed:
                return self._saved.fixed[key]
            return grp.get(key)

        self.tdgl_data = TDGLData(
            step=step, epsilon=get("epsilon"), psi=get("psi"), mu=get("mu"),
            applied_vector_potential=get("applied_vector_potential"),
            induced_vector_potential=get("induced_vector_potential"),
            supercurrent=get("supercurrent"), normal_current=get("normal_current"),
            state={k: v for k, v in grp["attrs"].items()})
        self._solve_step = step

    @property
    def solve_step(self) -> int:
        return self._solve_step

    @solve_step.setter
    def solve_step(self, step: int) -> None:
        self.load_tdgl_data(step)

    @property
    def times(self) -> Optional[np.ndarray]:
        """reference solution.py:134-146"""
        if self.dynamics is None:
            return None
        times = self.dynamics.time
        saved_times = times[:: self.options.save_every]
        if saved_times[-1] == times[-1]:
            return saved_times.copy()
        return np.concatenate([saved_times, times[-1:]])

    def closest_solve_step(self, time: float) -> int:
        return int(np.argmin(np.abs(self.times - time)))

    @property
    def field_units(self) -> str:
        return str(self.options.field_units)

    @property
    def current_units(self) -> str:
        return str(self.options.current_units)

    # -- persistence -----------------------------------------------------------------------
    def _tree(self) -> Dict[str, np.ndarray]:
        """The reference's output tree as flat ``path -> array`` pairs: root-level fixed
        arrays, ``data/<k>/<name>`` per saved step with ``step/time/dt/timestamp`` attributes
        (``data/<k>/attrs/<name>``) and ``data/<k>/running_state/<name>``
        (DataHandler.save_time_step, runner.py:155-183); the mesh arrays under ``mesh/``
        (Mesh.to_hdf5 / EdgeMesh.to_hdf5, finite_volume/mesh.py:345-368) and the solver options
        under ``solution/options`` (Solution._save_to_hdf5_file, solution.py:874-931)."""
        import json

        out: Dict[str, np.ndarray] = {}
        for k, v in self._saved.fixed.items():
            out[k] = v
        for i, grp in enumerate(self._saved.groups):
            for k, v in grp.items():
                if k == "attrs":
                    for ak, av in v.items():
                        out[f"data/{i}/attrs/{ak}"] = np.array(av)
                elif k == "running_state":
                    for rk, rv in v.items():
                        out[f"data/{i}/running_state/{rk}"] = rv
                else:
                    out[f"data/{i}/{k}"] = v
        mesh = getattr(self.device, "mesh", None) if self.device is not None else self._mesh
        if mesh is not None:
            em = mesh.edge_mesh
            out["mesh/sites"] = np.asarray(mesh.sites)
            out["mesh/elements"] = np.asarray(mesh.elements)
            out["mesh/boundary_indices"] = np.asarray(mesh.boundary_indices)
            out["mesh/areas"] = np.asarray(mesh.areas)
            for name in ("centers", "edges", "boundary_edge_indices", "directions",
                         "edge_lengths", "dual_edge_lengths"):
                out[f"mesh/edge_mesh/{name}"] = np.asarray(getattr(em, name))
        if self.options is not None:
            opts = dataclasses.asdict(self.options)
            opts["sparse_solver"] = getattr(opts["sparse_solver"], "value", opts["sparse_solver"])
            if isinstance(opts.get("terminal_psi"), complex):
                opts["terminal_psi"] = [opts["terminal_psi"].real, opts["terminal_psi"].imag]
            out["solution/options"] = np.array(json.dumps(opts))
        out["solution/total_seconds"] = np.array(self.total_seconds)
        out["solution/time_created"] = np.array(self._time_created.isoformat())
        return out

    def save(self, path: str) -> str:
        """What ``SolverOptions.output_file`` means here: the reference's HDF5 layout when h5py
        can be imported and the name does not end in ``.npz``, otherwise the same tree as one
        ``.npz`` (said in the log, because the file then is not what the name promises)."""
        import importlib.util
        import logging

        if not path.endswith(".npz") and importlib.util.find_spec("h5py") is not None:
            return self.to_hdf5(path)
        if not path.endswith(".npz"):
            logging.getLogger("solver").warning(
                "h5py is not installed: writing %s.npz (same tree of keys as the HDF5 file)", path)
        return self.to_npz(path)

    def to_npz(self, path: str) -> str:
        """Write the tree of :meth:`_tree` to one compressed ``.npz`` (h5py / libhdf5 are not
        part of this image; :meth:`to_hdf5` writes the same tree when they are)."""
        np.savez_compressed(path, **self._tree())
        self.path = path if path.endswith(".npz") else path + ".npz"
        return self.path

    def to_hdf5(self, path: str) -> str:
        """The same tree as an HDF5 file with the reference's layout (groups ``data/<k>`` with
        attributes, ``running_state`` sub-groups, ``mesh``, ``solution/options`` attributes).
        Needs h5py, which this image does not ship: raises ImportError without it."""
        import importlib.util
        import json

        if importlib.util.find_spec("h5py") is None:
            raise ImportError("h5py is not installed: use Solution.to_npz (same tree of keys)")
        import h5py

        with h5py.File(path, "x") as f:
            for key, value in self._tree().items():
                parts = key.split("/")
                if len(parts) >= 4 and parts[0] == "data" and parts[2] == "attrs":
                    f.require_group(f"data/{parts[1]}").attrs[parts[3]] = value[()]
                elif key == "solution/options":
                    grp = f.require_group("solution/options")
                    for k, v in json.loads(str(value)).items():
                        if v is not None:
                            grp.attrs[k] = v
                elif parts[0] == "solution":
                    f.require_group("solution").attrs[parts[1]] = value[()]
                else:
                    f[key] = value
        self.path = path
        return path

    @classmethod
    def from_npz(cls, path: str, device=None, options=None) -> "Solution":
        """Load a tree written by :meth:`to_npz` (the counterpart of the reference's
        ``Solution.from_hdf5``, solution/solution.py:933-1005, for the data this path
        produces).  The solver options and the mesh are restored from the file unless given;
        ``Solution.times`` and seeding a new solve (``seed_solution``) work on the result."""
        import json

        from .mesh import EdgeMesh, Mesh
        from .options import SolverOptions

        saved = SavedSteps()
        groups: Dict[int, Dict[str, Any]] = {}
        mesh_arrays: Dict[str, np.ndarray] = {}
        meta: Dict[str, Any] = {}
        with np.load(path, allow_pickle=False) as f:
            for key in f.files:
                parts = key.split("/")
                if parts[0] == "mesh":
                    mesh_arrays["/".join(parts[1:])] = f[key]
                    continue
                if parts[0] == "solution":
                    meta[parts[1]] = f[key]
                    continue
                if parts[0] != "data":
                    saved.fixed[key] = f[key]
                    continue
                grp = groups.setdefault(int(parts[1]), {"attrs": {}})
                if parts[2] == "attrs":
                    v = f[key]
                    grp["attrs"][parts[3]] = v.item() if v.ndim == 0 else v
                elif parts[2] == "running_state":
                    grp.setdefault("running_state", {})[parts[3]] = f[key]
                else:
                    grp[parts[2]] = f[key]
        saved.groups = [groups[k] for k in sorted(groups)]
        if options is None and "options" in meta:
            d = json.loads(str(meta["options"]))
            if isinstance(d.get("terminal_psi"), list):
                d["terminal_psi"] = complex(*d["terminal_psi"])
            known = {fl.name for fl in dataclasses.fields(SolverOptions)}
            options = SolverOptions(**{k: v for k, v in d.items() if k in known})
        mesh = None
        if mesh_arrays:
            g = mesh_arrays
            em = EdgeMesh(g["edge_mesh/centers"], g["edge_mesh/edges"],
                          g["edge_mesh/boundary_edge_indices"], g["edge_mesh/directions"],
                          g["edge_mesh/edge_lengths"], g["edge_mesh/dual_edge_lengths"])
            mesh = Mesh(g["sites"], g["elements"], g["boundary_indices"], areas=g["areas"],
                        edge_mesh=em)
        if options is None:
            raise ValueError(f"{path} holds no solver options: pass options=")
        sol = cls(device=device, options=options, saved=saved, path=path, mesh=mesh,
                  total_seconds=float(meta.get("total_seconds", 0.0)))
        sol.load_tdgl_data(-1)
        return sol
