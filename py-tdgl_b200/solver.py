"""``TDGLSolver`` / ``solve``: the reference's public solve API for the hot path, with the
per-step work done by the CUDA engine.

Mirrors ``tdgl.solve`` (tdgl/solver/solve.py:9-52), ``TDGLSolver.__init__`` / ``.solve``
(tdgl/solver/solver.py:117-323, 716-827) and the loop and save cadence of
``Runner.run`` / ``_run_stage`` (tdgl/solver/runner.py:288-454).  What the reference
does per step in NumPy/SciPy (``TDGLSolver.update`` solver.py:580-714) happens inside
``tdgl_advance`` on the device; Python wakes up only at save steps and for
time-dependent terminal currents / epsilon.
"""

from __future__ import annotations

import inspect
import logging
from datetime import datetime
from typing import Callable, Dict, NamedTuple, Optional, Sequence, Union

import numpy as np

from .engine import DeviceEngine, ScreeningFailed, StepFailed
from .options import SolverOptions
from .solution import SavedSteps, Solution
from .synthetic import TerminalInfo

logger = logging.getLogger("solver")


class SolverResult(NamedTuple):
    """Same tuple as the reference's ``SolverResult`` (solver.py:63-85)."""

    dt: float
    psi: np.ndarray
    mu: np.ndarray
    supercurrent: np.ndarray
    normal_current: np.ndarray
    A_induced: np.ndarray
    A_applied: Optional[np.ndarray] = None
    epsilon: Optional[np.ndarray] = None


def validate_terminal_currents(terminal_currents, terminal_info, solver_options,
                               num_evals: int = 100) -> None:
    """Current conservation check, same errors as the reference (solver.py:35-60)."""

    def check_total_current(currents: Dict[str, float]):
        names = set(t.name for t in terminal_info)
        if unknown := set(currents).difference(names):
            raise ValueError(f"Unknown terminal(s) in terminal currents: {list(unknown)}.")
        total_current = sum(currents.values())
        if total_current:
            raise ValueError(
                f"The sum of all terminal currents must be 0 (got {total_current:.2e}).")

    if callable(terminal_currents):
        times = np.random.default_rng().random(num_evals) * solver_options.solve_time
        for t in times:
            check_total_current(terminal_currents(t))
    else:
        check_total_current(terminal_currents)


class _DeferredInterrupt:
    """Ctrl-C while a stage runs sets a flag instead of raising inside a device call: the
    stage loop looks at it between chunks, where the host bookkeeping and the device state
    agree, and then does what ``Runner._run_stage`` does in its ``except KeyboardInterrupt``
    (runner.py:434-451)."""

    def __init__(self):
        self.pending = False
        self._old = None

    def _handler(self, signum, frame):
        self.pending = True

    def __enter__(self):
        import signal
        import threading

        if threading.current_thread() is threading.main_thread():
            self._old = signal.signal(signal.SIGINT, self._handler)
        return self

    def __exit__(self, *exc):
        import signal

        if self._old is not None:
            signal.signal(signal.SIGINT, self._old)
            self._old = None
        return False


class _AsyncSaver:
    """Save pipeline of a stage: the stepping thread only starts a device->pinned-host copy
    (``tdgl_snapshot_begin``, two slots) and goes on stepping; this writer thread waits for
    the copy, moves the arrays into the output container (``SavedSteps`` — the counterpart of
    ``DataHandler.save_time_step``, runner.py:155-183) and frees the slot.  Jobs are
    processed in order, so the groups come out as the reference numbers them."""

    def __init__(self, engine, saved: SavedSteps):
        import queue
        import threading

        self.engine, self.saved = engine, saved
        self.jobs: "queue.Queue" = queue.Queue()
        self.free: "queue.Queue" = queue.Queue()
        for slot in (0, 1):
            self.free.put(slot)
        self.error: Optional[BaseException] = None
        self.thread = threading.Thread(target=self._run, name="tdgl-b200-writer", daemon=True)
        self.thread.start()

    def _check(self) -> None:
        if self.error is not None:
            err, self.error = self.error, None
            raise err

    def save_host(self, state, values, running) -> None:
        """Values that are already on the host (step 0 of a run)."""
        self._check()
        self.jobs.put((None, state, values, running))

    def save_device(self, state, extra, running) -> None:
        """Snapshot the engine's current state; ``extra``: the host-side entries of the group."""
        slot = self.free.get()        # blocks only while both slots are still being written
        self._check()
        self.engine.snapshot_begin(slot)
        self.jobs.put((slot, state, extra, running))

    def _run(self) -> None:
        while True:
            job = self.jobs.get()
            if job is None:
                return
            slot, state, values, running = job
            try:
                if slot is not None:
                    psi, mu, js, jn = self.engine.snapshot_wait(slot)
                    values = dict({"psi": psi, "mu": mu, "supercurrent": js,
                                   "normal_current": jn}, **values)
                self.saved.save_time_step(state, values, running)   # copies out of the views
            except BaseException as exc:  # noqa: BLE001  (re-raised on the stepping thread)
                self.error = exc
            finally:
                if slot is not None:
                    self.free.put(slot)

    def close(self) -> None:
        self.jobs.put(None)
        self.thread.join()
        self._check()


class _RunningState:
    """reference ``RunningState`` (runner.py:186-221)."""

    def __init__(self, names_and_sizes: Dict[str, int], buffer_size: int):
        self.step = 0
        self.buffer_size = buffer_size
        self.names_and_sizes = names_and_sizes
        self.clear()

    def clear(self) -> None:
        self.step = 0
        self.values = {name: np.zeros((size, self.buffer_size))
                       for name, size in self.names_and_sizes.items()}

    def append(self, name: str, value) -> None:
        self.values[name][:, self.step] = value


class TDGLSolver:
    """Drop-in for ``tdgl.TDGLSolver`` on the B200 engine.

    Use ``TDGLSolver(device, options, ...)`` exactly as in the reference, or
    ``TDGLSolver.from_dimensionless(...)`` when the inputs are already dimensionless
    mesh-level arrays (synthetic workloads, tests).
    """

    def __init__(self, device, options: SolverOptions,
                 applied_vector_potential: Union[Callable, float] = 0.0,
                 terminal_currents: Union[Callable, Dict[str, float], None] = None,
                 disorder_epsilon: Union[Callable, float] = 1.0, seed_solution=None):
        from .device import constant_field_vector_potential, unit_scales

        self.device = device
        self.options = options
        options.validate()
        self.terminal_currents = terminal_currents
        self.seed_solution = seed_solution
        mesh = device.mesh
        if mesh is None:
            raise ValueError("The device has no mesh: call device.make_mesh() first.")
        xi = device.layer.coherence_length
        scales = unit_scales(device, options.field_units, options.current_units)
        sites = xi * mesh.sites
        edge_centers = xi * mesh.edge_mesh.centers
        z0 = device.layer.z0 * np.ones(len(edge_centers))
        # applied vector potential at the edge centres (solver.py:164-189); a callable with
        # ``time_dependent = True`` (the reference's ``Parameter``) is evaluated every step
        dynamic_A = bool(getattr(applied_vector_potential, "time_dependent", False))
        if not callable(applied_vector_potential):
            b = float(applied_vector_potential)

            def applied_vector_potential(x, y, z, _b=b):
                return constant_field_vector_potential(x, y, z, Bz=_b)

        self.applied_vector_potential = applied_vector_potential

        def eval_A(t=None, _f=applied_vector_potential, _c=edge_centers, _z=z0,
                   _s=scales.A_scale):
            kw = dict(t=t) if dynamic_A else {}
            return _s * np.asarray(_f(_c[:, 0], _c[:, 1], _z, **kw))[:, :2]

        A = eval_A(0)
        if A.shape != edge_centers.shape:
            raise ValueError(f"Unexpected shape for vector_potential: {A.shape}.")
        # scalar(t) * field(r) (tdgl_b200.sources): the device evaluates the ramp itself
        ramp = None
        sep = getattr(applied_vector_potential, "separable", None) if dynamic_A else None
        if sep is not None:
            spatial, t_knots, f_knots = sep
            A0 = scales.A_scale * np.asarray(
                spatial(edge_centers[:, 0], edge_centers[:, 1], z0))[:, :2]
            ramp = (A0, np.asarray(t_knots, float), np.asarray(f_knots, float))
        # epsilon at the sites (solver.py:191-216)
        dynamic_epsilon = False
        if callable(disorder_epsilon):
            argspec = inspect.getfullargspec(disorder_epsilon)
            dynamic_epsilon = "t" in argspec.kwonlyargs
            vectorized = (argspec.kwonlydefaults is not None
                          and argspec.kwonlydefaults.get("vectorized", False))
            eps_func = disorder_epsilon
        else:
            value = float(disorder_epsilon)
            vectorized = True

            def eps_func(r, _v=value):
                return _v * np.ones(len(r), dtype=float)

        self.disorder_epsilon = disorder_epsilon

        def eval_eps(t=None):
            kw = dict(t=t) if dynamic_epsilon else {}
            if vectorized:
                return np.asarray(eps_func(sites, **kw), dtype=float)
            return np.array([float(eps_func(r, **kw)) for r in sites])

        terminal_info = device.terminal_info()
        names = [t.name for t in terminal_info]
        if terminal_currents is None:
            terminal_currents = {name: 0 for name in names}
        if callable(terminal_currents):
            user_func = terminal_currents
            static_currents = False
        else:
            filled = {name: terminal_currents.get(name, 0) for name in names}
            unknown = set(terminal_currents).difference(names)
            if unknown:
                raise ValueError(f"Unknown terminal(s) in terminal currents: {list(unknown)}.")

            def user_func(t, _c=filled):
                return _c

            static_currents = True
        J_scale = scales.J_scale
        if hasattr(user_func, "knots"):      # a table (sources.PiecewiseLinearCurrents)
            scaled_currents = user_func.scaled(J_scale)
        else:
            scaled_currents = lambda t: {k: J_scale * v for k, v in user_func(t).items()}  # noqa: E731
        screening = None
        if options.include_screening:
            # solver.py:306-309: A_scale = mu_0 / (4 pi) K0 / A0 in 1 / length_units,
            # areas = A_scale * mesh.areas * xi^2, coordinates in length_units
            from .device import MU_0, length_scale

            a_scale = MU_0 / (4 * np.pi) * device.K0 / device.A0 * length_scale(device.length_units)
            screening = (a_scale * xi**2, sites, edge_centers)
        eps_table = None
        if hasattr(disorder_epsilon, "arrays"):          # sources.SeparableEpsilon
            e0, e1 = disorder_epsilon.arrays(sites)
            eps_table = (e0, e1, disorder_epsilon.times, disorder_epsilon.g)
            eval_eps = lambda t=None, _e=disorder_epsilon, _a=(e0, e1): (  # noqa: E731
                _a[0] + np.float64(_e.scale(0.0 if t is None else t)) * _a[1])
            dynamic_epsilon = True
        self._eps_table = eps_table
        self._setup(mesh, options, A, eval_eps, dynamic_epsilon, terminal_info,
                    scaled_currents,
                    static_currents, device.probe_point_indices, device.layer.u,
                    device.layer.gamma, eval_A=eval_A if dynamic_A else None, ramp=ramp,
                    screening=screening)

    @classmethod
    def from_dimensionless(cls, mesh, options: SolverOptions, *, A_applied, epsilon,
                           terminal_info: Sequence[TerminalInfo] = (),
                           terminal_currents: Union[Callable, Dict[str, float], None] = None,
                           probe_point_indices: Optional[Sequence[int]] = None,
                           u: float = 5.79, gamma: float = 10.0, device=None,
                           A_ramp=None, seed_solution=None,
                           screening_scale: float = 0.0) -> "TDGLSolver":
        """Inputs as the reference holds them after ``__init__``: ``A_applied`` [E, 2] in
        units of xi*Bc2 — or a callable ``t -> [E, 2]`` for a time-dependent vector
        potential (host callback every step, like the reference) — currents already
        multiplied by ``J_scale``.  ``A_ramp = (t_knots, f_knots)`` makes the potential
        ``f(t) * A_applied`` with piecewise-linear f, evaluated on the device.  ``epsilon``
        [N], or a callable ``t -> [N]`` for a time-dependent disorder (the reference's
        ``disorder_epsilon(r, *, t)``, solver.py:364-381).  With
        ``options.include_screening``: ``A_induced = screening_scale * sum_j J_j a_j / |c_e -
        r_j|`` in the mesh's own (dimensionless) coordinates and areas."""
        self = object.__new__(cls)
        self.device = device
        self.options = options
        options.validate()
        self.terminal_currents = terminal_currents
        self.seed_solution = seed_solution
        self.applied_vector_potential = None
        self.disorder_epsilon = None
        self._eps_table = None
        dynamic_epsilon = callable(epsilon)
        if hasattr(epsilon, "arrays"):                   # sources.SeparableEpsilon
            e0, e1 = epsilon.arrays(mesh.sites)
            self._eps_table = (e0, e1, epsilon.times, epsilon.g)
            eval_eps = lambda t=None, _e=epsilon, _a=(e0, e1): (  # noqa: E731
                _a[0] + np.float64(_e.scale(0.0 if t is None else t)) * _a[1])
        elif dynamic_epsilon:
            eval_eps = lambda t=None, _f=epsilon: np.asarray(  # noqa: E731
                _f(0.0 if t is None else t), dtype=float)
        else:
            eps = np.asarray(epsilon, dtype=float)
            eval_eps = lambda t=None: eps  # noqa: E731
        names = [t.name for t in terminal_info]
        if terminal_currents is None:
            terminal_currents = {n: 0.0 for n in names}
        if callable(terminal_currents):
            func, static = terminal_currents, False
        else:
            if unknown := set(terminal_currents).difference(names):
                raise ValueError(f"Unknown terminal(s) in terminal currents: {list(unknown)}.")
            filled = {n: terminal_currents.get(n, 0) for n in names}
            func, static = (lambda t, _c=filled: _c), True
        eval_A = None
        ramp = None
        if A_ramp is not None:
            A0 = np.asarray(A_applied, float)
            t_k, f_k = (np.asarray(k, float) for k in A_ramp)
            ramp = (A0, t_k, f_k)
            eval_A = lambda t, _A=A0, _t=t_k, _f=f_k: float(np.interp(t, _t, _f)) * _A  # noqa: E731
            A_applied = eval_A(0.0)
        elif callable(A_applied):
            eval_A = lambda t, _f=A_applied: np.asarray(_f(t), float)  # noqa: E731
            A_applied = eval_A(0.0)
        self._setup(mesh, options, np.asarray(A_applied, float), eval_eps, dynamic_epsilon,
                    tuple(terminal_info), func, static, probe_point_indices, u, gamma,
                    eval_A=eval_A, ramp=ramp,
                    screening=((screening_scale, mesh.sites, mesh.edge_mesh.centers)
                               if options.include_screening else None))
        return self

    # ------------------------------------------------------------------------------------
    def _setup(self, mesh, options, A, eval_eps, dynamic_epsilon, terminal_info, current_func,
               static_currents, probe_points, u, gamma, eval_A=None, ramp=None, screening=None):
        self.mesh = mesh
        self.u, self.gamma = u, gamma
        self.num_edges = len(mesh.edge_mesh.edges)
        self.current_A_applied = A
        self._eval_eps = eval_eps
        self.dynamic_epsilon = dynamic_epsilon
        self._eval_A = eval_A
        self._ramp = ramp     # (A0, t_knots, f_knots): A = f(t) * A0 evaluated on the device
        self.dynamic_vector_potential = eval_A is not None
        d = np.asarray(mesh.edge_mesh.directions, float)
        self.normalized_directions = d / np.linalg.norm(d, axis=1)[:, None]
        epsilon = eval_eps(0.0) if dynamic_epsilon else eval_eps()
        if np.any(epsilon > 1):
            raise ValueError("The disorder parameter epsilon must be <= 1")
        self.epsilon = epsilon
        self.terminal_info = tuple(terminal_info)
        self.terminal_names = [t.name for t in self.terminal_info]
        for term in self.terminal_info:
            if term.length == 0:
                raise ValueError(
                    f"Terminal {term.name!r} does not contain any points"
                    " on the boundary of the mesh.")
        self.current_func = current_func
        self.static_currents = static_currents
        validate_terminal_currents(current_func, self.terminal_info, options)
        idx = [np.asarray(t.site_indices, dtype=np.int64) for t in self.terminal_info]
        fixed = np.concatenate(idx) if idx else np.array([], dtype=np.int64)
        self.terminal_current_densities = {name: 0 for name in self.terminal_names}
        self.probe_points = None if probe_points is None else [int(p) for p in probe_points]
        terminal_psi = options.terminal_psi
        n = len(mesh.sites)
        self.psi_init = np.ones(n, dtype=np.complex128)            # solver.py:285-287
        if terminal_psi is not None:
            self.psi_init[fixed] = terminal_psi
        self.mu_init = np.zeros(n)
        self.mu_boundary = np.zeros(len(mesh.edge_mesh.boundary_edge_indices))
        engine_cls = DeviceEngine
        cuda_device = options.cuda_device
        if options.distributed:
            # None -> DistributedEngine takes torch.cuda.current_device() (LOCAL_RANK under
            # torchrun); an explicit ordinal is honoured but two ranks may not share it
            from .sharded import DistributedEngine as engine_cls
        elif cuda_device is None:
            cuda_device = 0
        self.engine = engine_cls(
            mesh, fixed_sites=fixed, fix_psi=(terminal_psi is not None), gamma=gamma, u=u,
            probe_sites=self.probe_points, device=cuda_device, mu_rtol=options.mu_rtol,
            mu_max_iter=options.mu_max_iterations,
            use_graph=1 if options.use_cuda_graph else 2,
            running_capacity=max(int(options.save_every), 1))
        self.engine.set_link_exponents(A)
        if ramp is not None:
            self.engine.set_vector_potential_ramp(*ramp)
        self.engine.set_epsilon(epsilon)
        # tables the device evaluates itself (no per-step host work)
        self._current_table = False
        knots = getattr(current_func, "knots", None)
        if knots is not None and self.terminal_info:
            t_knots, table = knots
            if unknown := set(table).difference(self.terminal_names):
                raise ValueError(f"Unknown terminal(s) in terminal currents: {list(unknown)}.")
            nb = len(mesh.edge_mesh.boundary_edge_indices)
            term_of = np.full(nb, -1, dtype=np.int32)
            for k, term in enumerate(self.terminal_info):
                term_of[np.asarray(term.boundary_edge_indices, dtype=np.int64)] = k
            cur = np.array([table.get(name, np.zeros(len(t_knots))) for name in self.terminal_names])
            self.engine.set_terminal_current_table(
                term_of, [t.length for t in self.terminal_info], t_knots, cur)
            self._current_table = True
        if getattr(self, "_eps_table", None) is not None:
            self.engine.set_epsilon_table(*self._eps_table)
        self.include_screening = screening is not None
        if screening is not None:
            self.engine.set_screening(
                screening[0], screening[1], screening[2], tolerance=options.screening_tolerance,
                max_iterations=options.max_iterations_per_step,
                step_size=options.screening_step_size, step_drag=options.screening_step_drag)
        self.engine.set_stepper(
            dt_init=options.dt_init, dt_max=options.dt_max, adaptive=options.adaptive,
            adaptive_window=options.adaptive_window,
            max_solve_retries=options.max_solve_retries,
            adaptive_time_step_multiplier=options.adaptive_time_step_multiplier)
        self.stats = dict(steps=0, retries=0, mu_iterations=0, screening_iterations=0)

    def _raise_like_reference(self, exc):
        fi = exc.args[1]
        if isinstance(exc, ScreeningFailed):                      # solver.py:657-663
            o = self.options
            raise RuntimeError(
                f"Screening calculation failed to converge at step {fi.failed_step} after"
                f" {o.max_iterations_per_step} iterations. Relative error in"
                f" induced vector potential: {fi.screening_error:.2e}"
                f" (tolerance: {o.screening_tolerance:.2e}).") from None
        raise RuntimeError(                                       # solver.py:479-483
            f"Solver failed to converge in {self.options.max_solve_retries}"
            f" retries at step {fi.failed_step} with dt = {fi.failed_dt:.2e}."
            f" Try using a smaller dt_init.") from None

    def update_mu_boundary(self, time: float) -> None:
        """reference solver.py:325-345; uploads only when a density changed."""
        currents = self.current_func(time)
        changed = False
        if self._current_table:      # the device evaluates the table itself (k_step_begin)
            for term in self.terminal_info:
                dens = (-1 / term.length) * sum(
                    currents.get(name, 0) for name in self.terminal_names if name != term.name)
                self.terminal_current_densities[term.name] = dens
                self.mu_boundary[np.asarray(term.boundary_edge_indices)] = dens
            return
        for term in self.terminal_info:
            dens = (-1 / term.length) * sum(
                currents.get(name, 0) for name in self.terminal_names if name != term.name)
            if dens != self.terminal_current_densities[term.name]:
                self.terminal_current_densities[term.name] = dens
                self.mu_boundary[np.asarray(term.boundary_edge_indices)] = dens
                changed = True
        if changed:
            self.engine.set_mu_boundary(self.mu_boundary)

    def _update_vector_potential(self, time: float, dt_prev: float, prev_A=None) -> None:
        """reference solver.py:626-642: A(t) at the edge centres, dA/dt as a backward
        difference over the previous step's dt projected on the edge directions, new link
        variables only if A changed."""
        if not self.dynamic_vector_potential:
            return
        if self._ramp is not None:      # the device evaluates the ramp; keep the host copy
            self.current_A_applied = self._eval_A(time)
            return
        A = self._eval_A(time)
        prev = self.current_A_applied if prev_A is None else np.asarray(prev_A, float)
        dA_dt = np.einsum("ij, ij -> i", (A - prev) / dt_prev, self.normalized_directions)
        if not np.allclose(A, self.current_A_applied):
            self.engine.set_link_exponents(A)
        self.engine.set_dA_dt(dA_dt)
        self.current_A_applied = A

    def update(self, state: Dict[str, float], running_state, dt: float, *, psi, mu,
               supercurrent=None, normal_current=None, induced_vector_potential=None,
               applied_vector_potential=None, epsilon=None, out=None) -> SolverResult:
        """One time step with the reference's signature (``TDGLSolver.update``,
        solver.py:580-714), i.e. the function ``Runner._run_stage`` calls once per step
        (runner.py:417-423): host ``psi`` / ``mu`` in, ``SolverResult`` of host arrays out.
        ``supercurrent`` / ``normal_current`` inputs are ignored as in the reference; ``dt``
        (the previous step's dt) only matters for time-dependent vector potentials (the
        backward difference dA/dt, solver.py:632).  ``solve()`` does not use this seam — it
        keeps the state on the device between save steps."""
        step, time = int(state["step"]), float(state["time"])
        self.update_mu_boundary(time)
        self._update_vector_potential(time, float(dt), applied_vector_potential)
        if self.dynamic_epsilon:
            self.epsilon = self._eval_eps(time)
            if self._eps_table is None:
                self.engine.set_epsilon(self.epsilon)
        if self.include_screening and induced_vector_potential is not None:
            self.engine.set_induced_vector_potential(induced_vector_potential)
        try:
            info, (psi1, mu1, js, jn) = self.engine.update(psi, mu, step, time, out=out)
        except (StepFailed, ScreeningFailed) as exc:
            self._raise_like_reference(exc)
        if self.include_screening:
            induced_vector_potential = self.engine.get_induced_vector_potential()
        if running_state is not None:
            running_state.append("dt", info.dt)
            if self.probe_points is not None:
                running_state.append("mu", mu1[self.probe_points])
                running_state.append("theta", np.angle(psi1[self.probe_points]))
            if self.include_screening:
                running_state.append("screening_iterations", info.screening_iterations)
        self.stats["steps"] += 1
        self.stats["retries"] += info.retries
        self.stats["mu_iterations"] += info.mu_iterations
        self.stats["screening_iterations"] += info.screening_iterations
        if induced_vector_potential is None:
            # (one shared read-only zero array: a fresh 16 B/edge allocation per step costs
            # more than the device-to-host copies of a step at 1M sites)
            if getattr(self, "_zero_A_induced", None) is None:
                self._zero_A_induced = np.zeros((self.num_edges, 2))
                self._zero_A_induced.setflags(write=False)
            induced_vector_potential = self._zero_A_induced
        results = [info.dt, psi1, mu1, js, jn, induced_vector_potential]
        if self.dynamic_vector_potential:
            results.append(self.current_A_applied)
        if self.dynamic_epsilon:
            results.append(self.epsilon)
        # positional contract of the seam (reference solver.py:708-714): A_applied only if the
        # vector potential is dynamic, then epsilon only if it is dynamic — Runner unpacks
        # `new_dt, *values` against exactly that many names (runner.py:424-428)
        return SolverResult(*results)

    # ------------------------------------------------------------------------------------
    def _values(self) -> Dict[str, np.ndarray]:
        psi, mu = self.engine.get_state()
        js, jn = self.engine.get_currents()
        out = {"psi": psi, "mu": mu, "supercurrent": js, "normal_current": jn,
               "induced_vector_potential": (self.engine.get_induced_vector_potential()
                                            if self.include_screening
                                            else np.zeros((self.num_edges, 2)))}
        if self.dynamic_vector_potential:
            out["applied_vector_potential"] = self.current_A_applied
        if self.dynamic_epsilon:
            out["epsilon"] = self.epsilon
        return out

    def _host_values(self) -> Dict[str, np.ndarray]:
        """The entries of a saved group that live on the host (the rest comes from the
        engine's snapshot)."""
        out = {"induced_vector_potential": (self.engine.get_induced_vector_potential()
                                            if self.include_screening
                                            else np.zeros((self.num_edges, 2)))}
        if self.dynamic_vector_potential:
            out["applied_vector_potential"] = self.current_A_applied
        if self.dynamic_epsilon:
            out["epsilon"] = self.epsilon
        return out

    def _run_stage(self, name: str, end_time: float, save: bool, saved: SavedSteps,
                   running: _RunningState, first_values: Optional[dict]) -> bool:
        """The loop of ``Runner._run_stage`` (runner.py:379-453) in chunks that end at the
        next save step (or after one step when a host callback is time-dependent)."""
        opts = self.options
        # single-device engines save through the asynchronous pipeline; sharded ones (whose
        # outputs are summed over the shards) synchronously
        saver = (_AsyncSaver(self.engine, saved)
                 if save and opts.async_save and type(self.engine) is DeviceEngine else None)
        try:
            return self._stage_loop(name, end_time, save, saved, running, first_values, saver)
        finally:
            if saver is not None:
                saver.close()

    def _stage_loop(self, name, end_time, save, saved, running, first_values, saver) -> bool:
        opts = self.options
        every = max(int(opts.save_every), 1)
        host_currents = (not self.static_currents) and not self._current_table
        host_epsilon = self.dynamic_epsilon and self._eps_table is None
        per_step_host = (host_currents or host_epsilon
                         or (self.dynamic_vector_potential and self._ramp is None))
        i, time = 0, 0.0
        cancelled = False

        def save_step(step):
            state = {"step": step, "time": time, "dt": self._prev_dt}
            run_vals = None if step == 0 else running.values
            if step == 0 and first_values is not None:
                if saver is not None:
                    saver.save_host(state, first_values, run_vals)
                else:
                    saved.save_time_step(state, first_values, run_vals)
            elif saver is not None:
                saver.save_device(state, self._host_values(), run_vals)
            else:
                saved.save_time_step(state, self._values(), run_vals)

        with _DeferredInterrupt() as intr:
            while True:
                if i % every == 0:
                    if save:
                        save_step(i)
                    running.clear()
                if intr.pending:                       # runner.py:434-451
                    intr.pending = False
                    msg = f"{{}} simulation at step {i} of stage {name!r}."
                    resume = False
                    if opts.pause_on_interrupt:
                        response = input(f"Simulation paused at stage {name!r} (step {i})."
                                         " Continue simulation? [yN]")
                        resume = response.lower().startswith("y")
                    if resume:
                        logger.info(msg.format("Resuming"))
                    else:
                        logger.warning(msg.format("Cancelling"))
                        cancelled = True
                        break
                self.update_mu_boundary(time)
                self._update_vector_potential(time, self._prev_dt)
                if host_epsilon:
                    self.epsilon = self._eval_eps(time)
                    self.engine.set_epsilon(self.epsilon)
                chunk = 1 if per_step_host else every - (i % every)
                try:
                    info = self.engine.advance(chunk, end_time, i, time)
                except (StepFailed, ScreeningFailed) as exc:
                    self._raise_like_reference(exc)
                k = info.steps_done
                t_last = info.time if info.finished else info.time - info.dt
                if self._ramp is not None:
                    # the vector potential of the last step taken (saved with the results)
                    self.current_A_applied = self._eval_A(t_last)
                if self._eps_table is not None:     # epsilon of the last step taken
                    self.epsilon = self._eval_eps(t_last)
                dt, mu_p, th_p = self.engine.get_running(k)
                pos = running.step
                running.values["dt"][0, pos:pos + k] = dt
                if self.probe_points is not None:
                    running.values["mu"][:, pos:pos + k] = mu_p
                    running.values["theta"][:, pos:pos + k] = th_p
                if self.include_screening:
                    running.values["screening_iterations"][0, pos:pos + k] = \
                        self.engine.get_running_screening(k)
                running.step = pos + (k - 1 if info.finished else k)
                self._prev_dt = info.dt
                self.stats["steps"] += k
                self.stats["retries"] += info.retries
                self.stats["mu_iterations"] += info.mu_iterations
                self.stats["screening_iterations"] += info.screening_iterations
                i, time = info.step, info.time
                if opts.progress_interval and (i // opts.progress_interval
                                               != (i - k) // opts.progress_interval):
                    logger.info(f"{name}: Time {time}/{end_time}, dt={info.dt:.2e}")
                if info.finished:
                    break
        if save and (i % every):
            save_step(i)
        return not cancelled

    def solve(self) -> Optional[Solution]:
        """reference ``TDGLSolver.solve`` (solver.py:716-827) + ``Runner.run``
        (runner.py:288-328)."""
        start_time = datetime.now()
        opts = self.options
        opts.validate()
        if self.seed_solution is None:
            psi0, mu0 = self.psi_init, self.mu_init
            first = {"psi": psi0.copy(), "mu": mu0.copy(),
                     "supercurrent": np.zeros(self.num_edges),
                     "normal_current": np.zeros(self.num_edges),
                     "induced_vector_potential": np.zeros((self.num_edges, 2))}
        else:
            if self.device is not None and self.seed_solution.device != self.device:
                raise ValueError(
                    "The seed_solution.device must be equal to the device being simulated.")
            seed = self.seed_solution.tdgl_data
            psi0, mu0 = seed.psi, seed.mu
            first = {"psi": np.array(seed.psi), "mu": np.array(seed.mu),
                     "supercurrent": np.array(seed.supercurrent),
                     "normal_current": np.array(seed.normal_current),
                     "induced_vector_potential": np.array(seed.induced_vector_potential)}
        self.engine.set_state(psi0, mu0)
        if self.include_screening:
            self.engine.set_induced_vector_potential(first["induced_vector_potential"])
        saved = SavedSteps()
        fixed = {}
        if self.dynamic_vector_potential:
            first["applied_vector_potential"] = self.current_A_applied
        else:
            fixed["applied_vector_potential"] = self.current_A_applied
        if self.dynamic_epsilon:
            first["epsilon"] = self.epsilon
        else:
            fixed["epsilon"] = self.epsilon
        saved.save_fixed_values(fixed)
        names = {"dt": 1}
        if self.probe_points is not None:
            names["mu"] = len(self.probe_points)
            names["theta"] = len(self.probe_points)
        if self.include_screening:
            names["screening_iterations"] = 1                       # solver.py:777-778
        running = _RunningState(names, max(int(opts.save_every), 1))
        self._prev_dt = float(opts.dt_init)
        ok = True
        if opts.skip_time:
            ok = self._run_stage("Thermalizing", opts.skip_time, False, saved, running, None)
            running.clear()
            first = None
        if not ok:
            return None
        self._run_stage("Simulating", opts.solve_time, True, saved, running, first)
        end_time = datetime.now()
        logger.info(f"Simulation took {end_time - start_time}")
        solution = Solution(
            device=self.device, options=opts, saved=saved,
            applied_vector_potential=self.applied_vector_potential,
            terminal_currents=self.terminal_currents, disorder_epsilon=self.disorder_epsilon,
            total_seconds=(end_time - start_time).total_seconds(),
            solver_stats=dict(self.stats, **self.engine.info()), mesh=self.mesh)
        if opts.output_file is not None:
            solution.save(opts.output_file)
        return solution


def solve(device, options: SolverOptions,
          applied_vector_potential: Union[Callable, float] = 0,
          terminal_currents: Union[Callable, Dict[str, float], None] = None,
          disorder_epsilon: Union[float, Callable] = 1, seed_solution=None):
    """Same signature as ``tdgl.solve`` (tdgl/solver/solve.py:9-52)."""
    solver = TDGLSolver(device=device, options=options,
                        applied_vector_potential=applied_vector_potential,
                        terminal_currents=terminal_currents, disorder_epsilon=disorder_epsilon,
                        seed_solution=seed_solution)
    return solver.solve()
