"""CPU oracle: a NumPy/SciPy restatement of pyTDGL's per-step hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU baseline — never as a fallback of
the CUDA path (the product fails loudly when ``libtdgl_b200.so`` is missing).

Parity pinning: the reference's own tests hold no golden vectors for this path
(SURVEY.md §4); this restatement is pinned against the UNMODIFIED reference code
executed in the build container through ``oracle/ref_loader.py`` — see
``tests/test_oracle_vs_reference.py`` (runs where ``/root/reference`` exists) and the
committed fixtures ``tests/golden/*.npz`` produced by ``oracle/make_golden.py`` from the
reference itself (checked on every box by ``tests/test_oracle_golden.py``).

Each function cites the reference file:line it follows (paths relative to
``/root/reference/``).  Like the reference, the mu system is solved with SuperLU through
``scipy.sparse.linalg.factorized`` on the *singular* pure-Neumann Laplacian
(``tdgl/finite_volume/operators.py:285,306-308``), so raw ``mu`` carries an arbitrary
additive constant and raw ``psi`` an arbitrary global phase: compare through
``gauge_fix``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, NamedTuple, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class TerminalInfo(NamedTuple):
    """reference ``tdgl/device/device.py:30-46``"""

    name: str
    site_indices: Sequence[int]
    edge_indices: Sequence[int]
    boundary_edge_indices: Sequence[int]
    length: float


# ------------------------------------------------------------------ operators (L0)

def divergence_matrix(edges, areas, dual_len, n_sites) -> sp.csr_array:
    """(e0, e) = s_e / a[e0], (e1, e) = -s_e / a[e1]  — operators.py:59-84"""
    E = len(edges)
    e = np.arange(E)
    i0, i1 = edges[:, 0], edges[:, 1]
    return sp.csr_array(
        (np.concatenate([dual_len / areas[i0], -dual_len / areas[i1]]),
         (np.concatenate([i0, i1]), np.concatenate([e, e]))),
        shape=(n_sites, E))


def link_variables(A, directions) -> np.ndarray:
    """U_e = exp(-i A_e . d_e)  — operators.py:109-111,153-155"""
    return np.exp(-1j * np.einsum("ij, ij -> i", A, directions))


def gradient_matrix(edges, edge_len, n_sites, U=None) -> sp.csr_array:
    """(e, e1) = U_e / l_e, (e, e0) = -1 / l_e  — operators.py:87-117"""
    E = len(edges)
    e = np.arange(E)
    w = 1 / edge_len
    link = np.ones(E) if U is None else U
    return sp.csr_array(
        (np.concatenate([link * w, -w]),
         (np.concatenate([e, e]), np.concatenate([edges[:, 1], edges[:, 0]]))),
        shape=(E, n_sites))


def laplacian_matrix(edges, areas, edge_len, dual_len, n_sites, U=None,
                     fixed_sites=None) -> sp.csc_array:
    """Off-diagonals w_e U_e / a[e0], w_e conj(U_e) / a[e1]; diagonals -w_e / a
    accumulated; rows of ``fixed_sites`` replaced by a unit diagonal
    — operators.py:120-185"""
    i0, i1 = edges[:, 0], edges[:, 1]
    w = dual_len / edge_len
    link = np.ones(len(edges)) if U is None else U
    rows = np.concatenate([i0, i1, i0, i1])
    cols = np.concatenate([i1, i0, i0, i1])
    vals = np.concatenate([w * link / areas[i0], w * link.conjugate() / areas[i1],
                           -w / areas[i0], -w / areas[i1]])
    if fixed_sites is None:
        fixed_sites = np.array([], dtype=np.int64)
    free = np.isin(rows, fixed_sites, invert=True)
    rows = np.concatenate([rows[free], fixed_sites])
    cols = np.concatenate([cols[free], fixed_sites])
    vals = np.concatenate([vals[free], np.ones(len(fixed_sites))])
    return sp.csc_array((vals, (rows, cols)), shape=(n_sites, n_sites))


def neumann_boundary_matrix(edges, areas, edge_len, boundary_edge_indices,
                            n_sites) -> sp.csr_array:
    """(i, b) = (j, b) = l_b / (2 a) for boundary edge b = (i, j)  — operators.py:188-230"""
    nb = len(boundary_edge_indices)
    b = np.arange(nb)
    be = edges[boundary_edge_indices]
    bl = edge_len[boundary_edge_indices]
    return sp.csr_array(
        (np.concatenate([bl / (2 * areas[be[:, 0]]), bl / (2 * areas[be[:, 1]])]),
         (np.concatenate([be[:, 0], be[:, 1]]), np.concatenate([b, b]))),
        shape=(n_sites, nb))


class OracleOperators:
    """reference ``MeshOperators`` (operators.py:233-394) for the SuperLU path."""

    def __init__(self, mesh, fixed_sites=None, fix_psi: bool = True):
        em = mesh.edge_mesh
        self.n = len(mesh.sites)
        self.edges = np.asarray(em.edges)
        self.areas = np.asarray(mesh.areas)
        self.edge_len = np.asarray(em.edge_lengths)
        self.dual_len = np.asarray(em.dual_edge_lengths)
        self.directions = np.asarray(em.directions)
        self.boundary_edge_indices = np.asarray(em.boundary_edge_indices)
        self.fixed_sites = (np.array([], dtype=np.int64) if fixed_sites is None
                            else np.asarray(fixed_sites, dtype=np.int64))
        self.fix_psi = fix_psi
        a = (self.edges, self.areas, self.edge_len, self.dual_len, self.n)
        # build_operators — operators.py:282-308 (mu Laplacian has NO fixed sites)
        self.mu_laplacian = laplacian_matrix(*a)
        self.mu_boundary_laplacian = neumann_boundary_matrix(
            self.edges, self.areas, self.edge_len, self.boundary_edge_indices, self.n)
        self.mu_gradient = gradient_matrix(self.edges, self.edge_len, self.n)
        self.divergence = divergence_matrix(self.edges, self.areas, self.dual_len, self.n)
        spla.use_solver(useUmfpack=False)
        self.mu_laplacian_lu = spla.factorized(self.mu_laplacian)
        self.psi_gradient = None
        self.psi_laplacian = None

    def set_link_exponents(self, A) -> None:
        """operators.py:310-383.  (Rebuilding gives the same matrices as the in-place
        value rewrite of :346-383.)"""
        U = link_variables(np.asarray(A, float), self.directions)
        self.psi_gradient = gradient_matrix(self.edges, self.edge_len, self.n, U)
        self.psi_laplacian = laplacian_matrix(
            self.edges, self.areas, self.edge_len, self.dual_len, self.n, U,
            self.fixed_sites if self.fix_psi else None)

    def get_supercurrent(self, psi):
        """J_s[e] = Im(conj(psi[e0]) * (grad_psi @ psi)[e])  — operators.py:385-394"""
        return (psi.conjugate()[self.edges[:, 0]] * (self.psi_gradient @ psi)).imag


# ------------------------------------------------------------------ screening (row S)

def get_quantity_on_site(edges, normalized_directions, n_sites, quantity_on_edge):
    """Edge quantity -> site vector: the mean over the edges at a site of the quantity times
    the edge's unit vector, divided by 2  — finite_volume/mesh.py:203-243 (vector=True)."""
    flux_x = quantity_on_edge * normalized_directions[:, 0]
    flux_y = quantity_on_edge * normalized_directions[:, 1]
    vertices = np.concatenate([edges[:, 0], edges[:, 1]])
    counts = np.bincount(vertices, minlength=n_sites)
    x = np.bincount(vertices, weights=np.concatenate([flux_x, flux_x]), minlength=n_sites) / counts
    y = np.bincount(vertices, weights=np.concatenate([flux_y, flux_y]), minlength=n_sites) / counts
    return np.array([x, y]).T / 2


def get_A_induced(J_site, site_areas, sites, edge_centers, chunk: int = 2048):
    """A_induced[i, k] = sum_j J_site[j, k] * site_areas[j] / |edge_centers[i] - sites[j]|
    — solver/screening.py:12-42 (the numba kernel; summed here in site order per edge like
    its inner loop, in blocks of edges)."""
    out = np.empty((len(edge_centers), 2))
    w = J_site * site_areas[:, None]
    for i0 in range(0, len(edge_centers), chunk):
        c = edge_centers[i0:i0 + chunk]
        dx = c[:, 0][:, None] - sites[:, 0][None, :]
        dy = c[:, 1][:, None] - sites[:, 1][None, :]
        inv = 1.0 / np.sqrt(dx * dx + dy * dy)
        out[i0:i0 + chunk] = inv @ w
    return out


# ------------------------------------------------------------------ step physics (L1)

def solve_for_psi_squared(psi, abs_sq_psi, mu, epsilon, gamma, u, dt, psi_laplacian):
    """psi^n, mu^n -> psi^{n+1}, |psi^{n+1}|^2 (quadratic root), or None
    — solver.py:383-439 (expression order kept)."""
    U = np.exp(-1j * mu * dt)
    z = U * gamma**2 / 2 * psi
    with np.errstate(all="raise"):
        try:
            w = z * abs_sq_psi + U * (
                psi + (dt / u) * np.sqrt(1 + gamma**2 * abs_sq_psi)
                * ((epsilon - abs_sq_psi) * psi + psi_laplacian @ psi))
            c = w.real * z.real + w.imag * z.imag
            two_c_1 = 2 * c + 1
            w2 = np.absolute(w) ** 2
            disc = two_c_1**2 - 4 * np.absolute(z) ** 2 * w2
        except Exception:
            return None
    if np.any(disc < 0):
        return None
    new_sq = (2 * w2) / (two_c_1 + np.sqrt(disc))
    return w - z * new_sq, new_sq


@dataclass
class OracleOptions:
    """The ``SolverOptions`` fields the hot path reads (options.py:66-89)."""

    solve_time: float = 1.0
    skip_time: float = 0.0
    dt_init: float = 1e-6
    dt_max: float = 1e-1
    adaptive: bool = True
    adaptive_window: int = 10
    max_solve_retries: int = 10
    adaptive_time_step_multiplier: float = 0.25
    terminal_psi: Optional[complex] = 0.0
    save_every: int = 100
    include_screening: bool = False
    max_iterations_per_step: int = 1000
    screening_tolerance: float = 1e-3
    screening_step_size: float = 0.1
    screening_step_drag: float = 0.5


@dataclass
class OracleSolver:
    """State + one-step update of reference ``TDGLSolver`` (solver.py:88-714) with the
    pint/Device layer stripped: all inputs are already dimensionless."""

    mesh: object
    options: OracleOptions
    A_applied: np.ndarray                      # [E, 2], already A_scale-d (solver.py:185)
    epsilon: np.ndarray                        # [N]
    u: float = 5.79
    gamma: float = 10.0
    terminal_info: Sequence[TerminalInfo] = ()
    current_func: Optional[Callable[[float], Dict[str, float]]] = None  # J_scale-d
    probe_points: Optional[Sequence[int]] = None
    d_psi_sq_vals: List[float] = field(default_factory=list)
    # time-dependent vector potential: t -> [E, 2] (already A_scale-d); then ``A_applied``
    # must be its value at t = 0 (solver.py:164-185)
    A_func: Optional[Callable[[float], np.ndarray]] = None
    # time-dependent disorder: t -> [N]; then ``epsilon`` must be its value at t = 0
    # (solver.py:191-216, 364-381, 644-646)
    epsilon_func: Optional[Callable[[float], np.ndarray]] = None
    # screening (solver.py:304-314): A_induced = screening_scale * sum_j J_j a_j / |c_e - r_j|
    # with a_j, c_e, r_j the mesh's own (dimensionless) areas and coordinates, i.e.
    # screening_scale = [mu_0 / (4 pi) K0 / A0 in 1/length_units] * xi
    screening_scale: float = 0.0

    def __post_init__(self):
        o = self.options
        idx = [np.asarray(t.site_indices, dtype=np.int64) for t in self.terminal_info]
        self.fixed_sites = (np.concatenate(idx) if idx else np.array([], dtype=np.int64))
        self.terminal_names = [t.name for t in self.terminal_info]
        if self.current_func is None:
            zero = {n: 0.0 for n in self.terminal_names}
            self.current_func = lambda t: zero
        self.terminal_current_densities = {n: 0 for n in self.terminal_names}
        self.operators = OracleOperators(self.mesh, self.fixed_sites,
                                         fix_psi=(o.terminal_psi is not None))
        self.operators.set_link_exponents(self.A_applied)
        self.current_A_applied = np.asarray(self.A_applied, float)
        d = np.asarray(self.mesh.edge_mesh.directions, float)
        self.normalized_directions = d / np.linalg.norm(d, axis=1)[:, None]
        n = len(self.mesh.sites)
        self.psi_init = np.ones(n, dtype=np.complex128)       # solver.py:285-287
        if o.terminal_psi is not None:
            self.psi_init[self.fixed_sites] = o.terminal_psi
        self.mu_init = np.zeros(n)
        self.mu_boundary = np.zeros(len(self.mesh.edge_mesh.boundary_edge_indices))
        self.tentative_dt = o.dt_init                          # solver.py:318-320
        self.dt_max = o.dt_max if o.adaptive else o.dt_init
        self.epsilon = np.asarray(self.epsilon, float)
        self.A_induced = np.zeros((len(self.mesh.edge_mesh.edges), 2))
        self.screening_iterations = 0
        self.screening_areas = self.screening_scale * np.asarray(self.mesh.areas, float)

    def get_induced_vector_potential(self, current_density, A_induced_vals, velocity):
        """One step of Polyak's method — solver.py:522-578."""
        o = self.options
        em = self.mesh.edge_mesh
        J_site = get_quantity_on_site(np.asarray(em.edges), self.normalized_directions,
                                      len(self.mesh.sites), current_density)
        new_A = get_A_induced(J_site, self.screening_areas, np.asarray(self.mesh.sites, float),
                              np.asarray(em.centers, float))
        A_induced = A_induced_vals[-1]
        dA = new_A - A_induced
        velocity.append((1 - o.screening_step_drag) * velocity[-1] + o.screening_step_size * dA)
        A_induced = A_induced + velocity[-1]
        A_induced_vals.append(A_induced)
        numerator = np.linalg.norm(dA, axis=1)
        denominator = np.maximum(np.linalg.norm(A_induced, axis=1), 1e-20)
        err = float(np.max(numerator / denominator))
        del velocity[:-2]
        del A_induced_vals[:-2]
        return A_induced, err

    def update_mu_boundary(self, time: float) -> None:
        """J_ext,k = -(1/L_k) sum_{j != k} I_j  — solver.py:325-345"""
        currents = self.current_func(time)
        for term in self.terminal_info:
            dens = (-1 / term.length) * sum(
                currents.get(name, 0) for name in self.terminal_names if name != term.name)
            if dens != self.terminal_current_densities[term.name]:
                self.terminal_current_densities[term.name] = dens
                self.mu_boundary[np.asarray(term.boundary_edge_indices)] = dens

    def adaptive_euler_step(self, step, psi, abs_sq_psi, mu, dt):
        """solver.py:441-487"""
        o = self.options
        L = self.operators.psi_laplacian
        res = solve_for_psi_squared(psi, abs_sq_psi, mu, self.epsilon, self.gamma, self.u,
                                    dt, L)
        retries = 0
        while res is None:
            if not o.adaptive or retries > o.max_solve_retries:
                raise RuntimeError(
                    f"Solver failed to converge in {o.max_solve_retries}"
                    f" retries at step {step} with dt = {dt:.2e}."
                    f" Try using a smaller dt_init.")
            dt = dt * o.adaptive_time_step_multiplier
            res = solve_for_psi_squared(psi, abs_sq_psi, mu, self.epsilon, self.gamma,
                                        self.u, dt, L)
            retries += 1
        return res[0], res[1], dt

    def solve_for_observables(self, psi, dA_dt=0.0):
        """solver.py:489-520"""
        ops = self.operators
        js = ops.get_supercurrent(psi)
        rhs = ops.divergence @ (js - dA_dt) - ops.mu_boundary_laplacian @ self.mu_boundary
        mu = ops.mu_laplacian_lu(rhs)
        jn = -(ops.mu_gradient @ mu) - dA_dt
        return mu, js, jn

    def update(self, step: int, time: float, psi, mu, dt_prev: Optional[float] = None,
               A_prev: Optional[np.ndarray] = None):
        """One time step — solver.py:580-714 (static epsilon, no screening).  ``dt_prev`` is
        the ``dt`` argument of the reference's ``update`` (the previous step's dt) and
        ``A_prev`` the ``applied_vector_potential`` value Runner threads through; both only
        matter for a time-dependent vector potential.  Returns (dt, psi', mu', J_s, J_n)."""
        o = self.options
        self.update_mu_boundary(time)
        dA_dt = 0.0
        if self.A_func is not None:                                      # :626-642
            A = np.asarray(self.A_func(time), float)
            prev = self.current_A_applied if A_prev is None else A_prev
            dA_dt = np.einsum("ij, ij -> i", (A - prev) / dt_prev, self.normalized_directions)
            if not np.allclose(A, self.current_A_applied):
                self.operators.set_link_exponents(A)
            self.current_A_applied = A
        if self.epsilon_func is not None:                                # :644-646
            self.epsilon = np.asarray(self.epsilon_func(time), float)
        old_sq = np.absolute(psi) ** 2                                   # :649
        if not o.include_screening:
            dt = self.tentative_dt                                       # :668
            psi, new_sq, dt = self.adaptive_euler_step(step, psi, old_sq, mu, dt)
            mu, js, jn = self.solve_for_observables(psi, dA_dt)
        else:
            # Polyak iteration on the induced vector potential (:650-688).  As in the
            # reference, psi and mu are re-assigned by every pass of the loop (the next pass
            # steps from them) while |psi|^2 of the step's input stays fixed.
            err = np.inf
            A_vals, velocity = [self.A_induced], [0.0]
            A_ind = self.A_induced
            it = 0
            while True:
                if err < o.screening_tolerance:
                    break
                if it > o.max_iterations_per_step:
                    raise RuntimeError(
                        f"Screening calculation failed to converge at step {step} after"
                        f" {o.max_iterations_per_step} iterations. Relative error in"
                        f" induced vector potential: {err:.2e}"
                        f" (tolerance: {o.screening_tolerance:.2e}).")
                if it == 0:
                    dt = self.tentative_dt
                self.operators.set_link_exponents(self.current_A_applied + A_ind)
                psi, new_sq, dt = self.adaptive_euler_step(step, psi, old_sq, mu, dt)
                mu, js, jn = self.solve_for_observables(psi, dA_dt)
                A_ind, err = self.get_induced_vector_potential(js + jn, A_vals, velocity)
                it += 1
            self.A_induced = A_ind
            self.screening_iterations = it                               # :695-696
        if o.adaptive:                                                   # :698-707
            self.d_psi_sq_vals.append(float(np.absolute(new_sq - old_sq).max()))
            if step > o.adaptive_window:
                new_dt = o.dt_init / max(
                    1e-10, np.mean(self.d_psi_sq_vals[-o.adaptive_window:]))
                self.tentative_dt = np.clip(0.5 * (new_dt + dt), 0, self.dt_max)
        return dt, psi, mu, js, jn


def run(solver: OracleSolver, *, end_time: float, max_steps: Optional[int] = None,
        psi0=None, mu0=None, dt0: Optional[float] = None):
    """The loop of ``Runner._run_stage`` (runner.py:379-433) without disk output: one more
    update is performed after ``time >= end_time`` is first reached, ``dt <- new_dt``,
    ``time += dt``.  Returns final fields, the dt sequence and the probe traces.  ``dt0``:
    Runner's ``self.dt`` on entry (``dt_init`` for the first stage, the thermalisation
    stage's last dt for the second, runner.py:262,431)."""
    psi = solver.psi_init.copy() if psi0 is None else np.array(psi0, complex)
    mu = solver.mu_init.copy() if mu0 is None else np.array(mu0, float)
    time = 0.0
    dts, mus, thetas, sits = [], [], [], []
    i = 0
    dt = solver.options.dt_init if dt0 is None else dt0          # Runner's self.dt
    while True:
        dt, psi, mu, js, jn = solver.update(i, time, psi, mu, dt_prev=dt)
        dts.append(float(dt))
        sits.append(solver.screening_iterations)
        if solver.probe_points is not None:                      # solver.py:691-694
            mus.append(mu[list(solver.probe_points)])
            thetas.append(np.angle(psi[list(solver.probe_points)]))
        if time >= end_time or (max_steps is not None and i + 1 >= max_steps):
            break
        time += dt
        i += 1
    out = dict(psi=psi, mu=mu, supercurrent=js, normal_current=jn, dt=np.array(dts),
               steps=i + 1, time=time)
    if solver.options.include_screening:
        out["screening_iterations"] = np.array(sits)
        out["induced_vector_potential"] = solver.A_induced
    if solver.probe_points is not None:
        out["running"] = dict(mu=np.array(mus).T, theta=np.array(thetas).T)
    return out


def run_stages(solver: OracleSolver, psi0=None, mu0=None):
    """``Runner.run`` (runner.py:288-328): an optional thermalisation stage up to
    ``skip_time`` whose results are not saved, then — with step and time reset to 0, the
    solver's ``tentative_dt`` / |psi|^2 history and Runner's ``self.dt`` carried over — the
    stage up to ``solve_time``.  Returns the second stage's ``run`` output."""
    o = solver.options
    dt0 = None
    if o.skip_time:
        th = run(solver, end_time=o.skip_time, psi0=psi0, mu0=mu0)
        # (the loop breaks before ``self.dt = new_dt``, runner.py:429-431: Runner's dt on
        # entry of the second stage is the dt of the LAST BUT ONE thermalisation step)
        psi0, mu0 = th["psi"], th["mu"]
        dt0 = float(th["dt"][-2]) if len(th["dt"]) >= 2 else o.dt_init
    return run(solver, end_time=o.solve_time, psi0=psi0, mu0=mu0, dt0=dt0)


# ------------------------------------------------------------------ gauge-fixed comparison

def gauge_fix(psi, mu, areas, psi_ref=None):
    """Remove what no correct solver can reproduce: the additive constant of mu (the
    reference's is SuperLU roundoff on a singular system) and the global phase of psi
    it integrates to.  mu -> mu - area-weighted mean; psi -> psi * exp(-i phi) with
    phi = arg<psi_ref, psi> (or arg of the area-weighted mean if no reference)."""
    mu0 = mu - np.dot(areas, mu) / areas.sum()
    if psi_ref is None:
        phi = np.angle(np.dot(areas, psi))
    else:
        phi = np.angle(np.vdot(psi_ref, psi))
    return psi * np.exp(-1j * phi), mu0


def compare(a: dict, b: dict, areas) -> dict:
    """Gauge-fixed relative differences (max-norm, relative to the max magnitude)."""
    pa, ma = gauge_fix(a["psi"], a["mu"], areas, b["psi"])
    pb, mb = gauge_fix(b["psi"], b["mu"], areas, b["psi"])

    def rel(x, y):
        return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))

    out = dict(psi=rel(pa, pb), abs_psi=rel(np.abs(a["psi"]), np.abs(b["psi"])),
               mu=rel(ma, mb))
    for k in ("supercurrent", "normal_current"):
        if k in a and k in b and a[k] is not None and b[k] is not None:
            out[k] = rel(a[k], b[k])
    if "dt" in a and "dt" in b and len(a["dt"]) == len(b["dt"]):
        out["dt"] = rel(np.asarray(a["dt"]), np.asarray(b["dt"]))
    return out
