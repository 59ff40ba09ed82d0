"""Stub-import loader for the UNMODIFIED reference (pyTDGL, /root/reference).

TEST INFRASTRUCTURE ONLY.  This module is used in the build container to (a) validate
the restatement in ``oracle/tdgl_oracle.py`` against the reference's own code and (b)
generate the golden vectors under ``tests/golden/`` (see ``oracle/make_golden.py``).
``/root/reference`` does not exist on the GPU box, so nothing in ``-m gpu`` tests,
``smoke()`` or ``bench.py`` imports this file.

``import tdgl`` fails here (no h5py / pint / shapely / meshpy / matplotlib / IPython), but
the hot-path modules import fine once those six packages are replaced by permissive
stubs and the ``tdgl`` package objects are synthesised so that ``tdgl/__init__.py`` is
never executed.  ``TDGLSolver.__init__`` needs pint, so a solver object
is assembled with ``object.__new__`` and the attributes ``__init__`` would have set
(reference ``tdgl/solver/solver.py:126-320``).
"""

from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Callable, Dict, Optional, Sequence

import numpy as np

REFERENCE_ROOT = os.environ.get("TDGL_REFERENCE_ROOT", "/root/reference")

_STUBBED = [
    "h5py",
    "matplotlib",
    "matplotlib.pyplot",
    "matplotlib.tri",
    "matplotlib.patches",
    "matplotlib.path",
    "matplotlib.colors",
    "matplotlib.cm",
    "matplotlib.animation",
    "shapely",
    "shapely.geometry",
    "shapely.geometry.polygon",
    "shapely.ops",
    "shapely.affinity",
    "shapely.validation",
    "pint",
    "IPython",
    "IPython.display",
    "meshpy",
    "meshpy.triangle",
]


class _Anything:
    """A class whose every attribute is itself; enough for module-level annotations."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tdgl", "solver"))


_loaded: Optional[types.SimpleNamespace] = None


def load() -> types.SimpleNamespace:
    """Import the reference's hot-path modules; returns a namespace of them."""
    global _loaded
    if _loaded is not None:
        return _loaded
    os.environ.setdefault("TQDM_DISABLE", "1")
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in _STUBBED:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                mod = _StubModule(name)
                mod.__path__ = []  # behave as a package
                sys.modules[name] = mod
    root = os.path.join(REFERENCE_ROOT, "tdgl")
    for pkg in ["tdgl", "tdgl.finite_volume", "tdgl.solver", "tdgl.device",
                "tdgl.solution", "tdgl.sources", "tdgl.visualization"]:
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(root, *pkg.split(".")[1:])]
            mod.__package__ = pkg
            sys.modules[pkg] = mod
    ns = types.SimpleNamespace()
    ns.util = importlib.import_module("tdgl.finite_volume.util")
    ns.edge_mesh = importlib.import_module("tdgl.finite_volume.edge_mesh")
    ns.mesh = importlib.import_module("tdgl.finite_volume.mesh")
    ns.options = importlib.import_module("tdgl.solver.options")
    ns.operators = importlib.import_module("tdgl.finite_volume.operators")
    ns.runner = importlib.import_module("tdgl.solver.runner")
    ns.solver = importlib.import_module("tdgl.solver.solver")
    ns.Mesh = ns.mesh.Mesh
    ns.MeshOperators = ns.operators.MeshOperators
    ns.TDGLSolver = ns.solver.TDGLSolver
    ns.SolverOptions = ns.options.SolverOptions
    ns.SparseSolver = ns.options.SparseSolver
    ns.RunningState = ns.runner.RunningState
    ns.TerminalInfo = ns.solver.TerminalInfo
    _loaded = ns
    return ns


def make_reference_mesh(sites: np.ndarray, elements: np.ndarray):
    """The reference's own ``Mesh.from_triangulation`` (mesh.py:104-151)."""
    ref = load()
    import logging

    # silence the per-site tqdm bar of util.py:205
    os.environ.setdefault("TQDM_DISABLE", "1")
    logging.getLogger("tdgl.finite_volume").setLevel(logging.ERROR)
    return ref.Mesh.from_triangulation(np.asarray(sites, float), np.asarray(elements))


def make_reference_solver(
    mesh,
    options,
    *,
    A_applied: np.ndarray,
    epsilon: np.ndarray,
    u: float = 5.79,
    gamma: float = 10.0,
    terminal_info: Sequence = (),
    terminal_currents: Optional[Dict[str, float]] = None,
    current_func: Optional[Callable] = None,
    probe_points: Optional[Sequence[int]] = None,
    A_func: Optional[Callable] = None,
):
    """Assemble a reference ``TDGLSolver`` without running its pint-dependent ``__init__``.
    ``A_func``: t -> [E, 2] makes the vector potential time-dependent (the reference then
    calls ``update_applied_vector_potential`` every step, solver.py:626-642).

    ``A_applied`` is the dimensionless (already ``A_scale``-d) vector potential at the edge
    centres, ``terminal_currents`` are already ``J_scale``-d (solver.py:176-185, 251-256).
    """
    ref = load()
    options.validate()
    s = object.__new__(ref.TDGLSolver)
    s.device = None
    s.options = options
    s.terminal_currents = terminal_currents
    s.seed_solution = None
    s.xp = np
    s.use_cupy = False
    s.probe_points = list(probe_points) if probe_points is not None else None
    s.num_edges = len(mesh.edge_mesh.edges)
    s.u = u
    s.gamma = gamma
    s.dynamic_vector_potential = False
    s.dynamic_epsilon = False
    s.terminal_info = tuple(terminal_info)
    s.terminal_names = [t.name for t in s.terminal_info]
    if current_func is None:
        tc = {name: 0.0 for name in s.terminal_names}
        if terminal_currents:
            tc.update(terminal_currents)

        def current_func(t, _tc=tc):
            return _tc

    s.current_func = current_func
    idx = [np.asarray(t.site_indices, dtype=np.int64) for t in s.terminal_info]
    fixed = np.concatenate(idx) if idx else np.array([], dtype=np.int64)
    s.terminal_current_densities = {name: 0 for name in s.terminal_names}
    terminal_psi = options.terminal_psi
    ops = ref.MeshOperators(
        mesh,
        options.sparse_solver,
        use_cupy=False,
        fixed_sites=fixed,
        fix_psi=(terminal_psi is not None),
    )
    ops.build_operators()
    ops.set_link_exponents(np.asarray(A_applied, float))
    s.operators = ops
    psi_init = np.ones(len(mesh.sites), dtype=np.complex128)
    if terminal_psi is not None:
        psi_init[fixed] = terminal_psi
    s.psi_init = psi_init
    s.mu_init = np.zeros(len(mesh.sites))
    s.epsilon = np.asarray(epsilon, float)
    s.mu_boundary = np.zeros_like(mesh.edge_mesh.boundary_edge_indices, dtype=float)
    s.normalized_directions = mesh.edge_mesh.normalized_directions
    s.current_A_applied = np.asarray(A_applied, float)
    s.new_A_induced = None
    s.areas = None
    s.d_psi_sq_vals = []
    s.tentative_dt = options.dt_init
    s.dt_max = options.dt_max if options.adaptive else options.dt_init
    if A_func is not None:
        s.dynamic_vector_potential = True
        s.update_applied_vector_potential = lambda time: np.asarray(A_func(time), float)
    return s


def run_reference(solver, *, end_time: float, max_steps: Optional[int] = None,
                  psi0=None, mu0=None, record_every: int = 0):
    """Drive ``TDGLSolver.update`` with the bookkeeping of ``Runner._run_stage``
    (runner.py:379-433) but no disk output.  Returns a dict of final fields, the dt
    sequence and (optionally) snapshots every ``record_every`` steps."""
    ref = load()
    opts = solver.options
    names = {"dt": 1}
    if solver.probe_points is not None:
        names["mu"] = len(solver.probe_points)
        names["theta"] = len(solver.probe_points)
    if opts.include_screening:
        names["screening_iterations"] = 1                      # solver.py:777-778
    # one long buffer: never cleared, so it holds the full trace
    nbuf = (max_steps or 0) + 2 if max_steps else 1_000_000
    running = ref.RunningState(names, nbuf)
    E = solver.num_edges
    values = dict(
        psi=solver.psi_init.copy() if psi0 is None else np.array(psi0, complex),
        mu=solver.mu_init.copy() if mu0 is None else np.array(mu0, float),
        supercurrent=np.zeros(E),
        normal_current=np.zeros(E),
        induced_vector_potential=np.zeros((E, 2)),
    )
    if solver.dynamic_vector_potential:
        values["applied_vector_potential"] = solver.current_A_applied
    time = 0.0
    dt = opts.dt_init
    dts = []
    snaps = []
    i = 0
    while True:
        state = {"step": i, "time": time, "dt": dt}
        if record_every and i % record_every == 0:
            snaps.append({"step": i, "time": time, "psi": values["psi"].copy(),
                          "mu": values["mu"].copy()})
        res = solver.update(state, running, dt, **values)
        new_dt = res[0]
        values = dict(zip(values.keys(), res[1:1 + len(values)]))
        dts.append(float(new_dt))
        if time >= end_time or (max_steps is not None and i + 1 >= max_steps):
            break
        dt = new_dt
        running.step += 1
        time += dt
        i += 1
    out = dict(values)
    out["dt"] = np.array(dts)
    out["steps"] = i + 1
    out["time"] = time
    out["running"] = {k: v[:, : i + 1].copy() for k, v in running.values.items()}
    out["snapshots"] = snaps
    return out
