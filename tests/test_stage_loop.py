"""Host logic of ``TDGLSolver.solve`` without a GPU: the stage loop (save cadence, stop rule,
running-state bookkeeping — reference ``Runner._run_stage``, tdgl/solver/runner.py:330-454),
the asynchronous save pipeline and the Ctrl-C handling, driven by a stand-in for the engine
that takes fixed-dt "steps" on the host (the arithmetic of a step is the CUDA engine's and is
tested on the GPU; here only the bookkeeping around ``tdgl_advance`` is under test)."""
import os
import signal
import threading

import numpy as np
import pytest

import tdgl_b200 as tdgl
from tdgl_b200 import solver as solver_mod
from tdgl_b200.engine import AdvanceInfo
from tdgl_b200.mesh import make_film_mesh


class FakeEngine:
    """The slice of ``DeviceEngine`` the stage loop uses.  A step multiplies psi by
    exp(0.1 i) and adds dt to mu; the loop semantics of ``tdgl_advance`` are the reference's:
    the update of step i runs, THEN `time >= t_end` ends the stage without advancing."""

    instances = []

    def __init__(self, mesh, *, probe_sites=None, running_capacity=0, **kw):
        self.n, self.E = len(mesh.sites), len(mesh.edge_mesh.edges)
        self.probes = [] if probe_sites is None else list(probe_sites)
        self.cap = running_capacity
        self.psi = np.ones(self.n, complex)
        self.mu = np.zeros(self.n)
        self.dt = 1e-3
        self.last = []            # (dt, mu_probe, theta_probe) of the last advance
        self.calls = []           # (max_steps, step, time) of every advance
        self.slots = {}
        self.begun = []
        self.fail_wait = False
        self.interrupt_at = None  # raise SIGINT inside the advance that starts at this step
        self.log = []             # setter calls, in order
        FakeEngine.instances.append(self)

    def set_link_exponents(self, A):
        self.log.append(("link", np.array(A)))

    def set_epsilon(self, eps):
        self.log.append(("eps", np.array(eps)))

    def set_mu_boundary(self, mub):
        self.log.append(("mub", np.array(mub)))

    def set_dA_dt(self, v):
        self.log.append(("dadt", np.array(v)))

    def set_vector_potential_ramp(self, A0, t, f):
        self.log.append(("ramp", np.array(A0), np.array(t), np.array(f)))

    def set_terminal_current_table(self, term_of, lengths, t, cur):
        self.log.append(("cur_table", np.array(term_of), np.array(lengths), np.array(t),
                         np.array(cur)))

    def set_epsilon_table(self, e0, e1, t, g):
        self.log.append(("eps_table", np.array(e0), np.array(e1), np.array(t), np.array(g)))

    def update(self, psi, mu, step, time, out=None):
        self.set_state(psi, mu)
        info = self.advance(1, 1e300, step, time)
        return info, (self.psi.copy(), self.mu.copy(), *self.get_currents())

    def close(self): pass

    def set_stepper(self, *, dt_init, **kw):
        self.dt = float(dt_init)

    def set_state(self, psi, mu):
        self.psi, self.mu = np.array(psi, complex), np.array(mu, float)

    def advance(self, max_steps, t_end, step, time):
        assert max_steps <= max(self.cap, 1)      # the ring buffer holds one chunk
        self.calls.append((max_steps, step, time))
        if self.interrupt_at is not None and step >= self.interrupt_at:
            self.interrupt_at = None
            os.kill(os.getpid(), signal.SIGINT)
        self.last, k, finished = [], 0, False
        while k < max_steps:
            self.psi = self.psi * np.exp(0.1j)
            self.mu = self.mu + self.dt
            self.last.append((self.dt, self.mu[self.probes], np.angle(self.psi[self.probes])))
            k += 1
            if time >= t_end:
                finished = True
                break
            time += self.dt
            step += 1
        return AdvanceInfo(k, step, time, self.dt, self.dt, finished, 0, 0, 0.0, 0, 3 * k, 0.0)

    def get_running(self, k):
        assert k == len(self.last)
        dt = np.array([r[0] for r in self.last])
        mu = np.array([r[1] for r in self.last]).T.reshape(len(self.probes), k)
        th = np.array([r[2] for r in self.last]).T.reshape(len(self.probes), k)
        return dt, mu, th

    def get_state(self):
        return self.psi.copy(), self.mu.copy()

    def get_currents(self):
        return np.full(self.E, self.mu[0]), np.full(self.E, -self.mu[0])

    def snapshot_begin(self, slot):
        assert slot not in self.slots, "slot reused before its copy was consumed"
        self.begun.append(slot)
        self.slots[slot] = (self.psi.copy(), self.mu.copy(), *self.get_currents())

    def snapshot_wait(self, slot):
        if self.fail_wait:
            self.slots.pop(slot)
            raise RuntimeError("copy failed")
        return self.slots.pop(slot)

    def info(self):
        return dict(n_sites=self.n)


@pytest.fixture()
def fake(monkeypatch):
    FakeEngine.instances.clear()
    monkeypatch.setattr(solver_mod, "DeviceEngine", FakeEngine)
    mesh = make_film_mesh(6, 4, 0.5)

    def make(**opts):
        o = tdgl.SolverOptions(**dict(dict(solve_time=0.0105, dt_init=1e-3, adaptive=False,
                                           save_every=4), **opts))
        s = tdgl.TDGLSolver.from_dimensionless(
            mesh, o, A_applied=np.zeros((len(mesh.edge_mesh.edges), 2)),
            epsilon=np.ones(len(mesh.sites)), probe_point_indices=[1, 5])
        return s, FakeEngine.instances[-1]

    return make


def _expected_updates(solve_time, dt):
    """Number of updates of a fixed-dt stage (runner.py:379-433): steps 0, 1, ... until the
    update that starts at time >= end."""
    t, i = 0.0, 0
    while t < solve_time:
        t += dt
        i += 1
    return i + 1


@pytest.mark.parametrize("async_save", [True, False])
@pytest.mark.parametrize("solve_time", [0.0105, 0.008, 0.0075])
def test_save_cadence_and_running_state(fake, async_save, solve_time):
    s, eng = fake(solve_time=solve_time, async_save=async_save)
    sol = s.solve()
    updates = _expected_updates(solve_time, 1e-3)
    last = updates - 1                                   # index of the last step taken
    steps = [g["attrs"]["step"] for g in sol._saved.groups]
    want = list(range(0, last + 1, 4)) + ([last] if last % 4 else [])
    assert steps == want
    # chunks end at save steps; the engine is never asked for more than one buffer
    assert all(n <= 4 for n, _, _ in eng.calls)
    assert [c[1] for c in eng.calls] == list(range(0, last + 1, 4))
    # group k holds the state BEFORE step k (group 0: the initial values), like the reference
    g1 = sol._saved.groups[1]
    np.testing.assert_allclose(g1["psi"], np.exp(0.1j * 4) * np.ones(eng.n))
    np.testing.assert_allclose(g1["mu"], 4e-3)
    assert g1["attrs"]["time"] == pytest.approx(4e-3) and g1["attrs"]["dt"] == 1e-3
    assert "running_state" not in sol._saved.groups[0]
    # dynamics: every update whose buffer was saved (the update of a last step that is itself
    # a save step is appended after that save and never written: runner.py:399-402,452-453)
    n_dyn = updates if last % 4 else updates - 1
    assert len(sol.dynamics.dt) == n_dyn and np.all(sol.dynamics.dt == 1e-3)
    assert sol.dynamics.mu.shape == (2, n_dyn)
    np.testing.assert_allclose(sol.dynamics.mu[0], 1e-3 * np.arange(1, n_dyn + 1))
    assert s.stats["steps"] == updates and s.stats["mu_iterations"] == 3 * updates
    if async_save:                                      # two pinned slots, used alternately
        assert eng.begun == [k % 2 for k in range(len(eng.begun))] and not eng.slots


def test_thermalisation_stage_is_not_saved(fake):
    s, eng = fake(skip_time=0.003, solve_time=0.004)
    sol = s.solve()
    t_updates = _expected_updates(0.003, 1e-3)
    steps = [g["attrs"]["step"] for g in sol._saved.groups]
    assert steps[0] == 0 and sol._saved.groups[0]["attrs"]["time"] == 0.0
    # the saved stage starts from the thermalised state, not from the initial values
    np.testing.assert_allclose(sol._saved.groups[0]["psi"], np.exp(0.1j * t_updates) * np.ones(eng.n))
    assert s.stats["steps"] == t_updates + _expected_updates(0.004, 1e-3)


def test_writer_thread_errors_reach_the_stepping_thread(fake):
    s, eng = fake(solve_time=0.02)
    eng.fail_wait = True
    with pytest.raises(RuntimeError, match="copy failed"):
        s.solve()
    assert not [t for t in threading.enumerate() if t.name == "tdgl-b200-writer"]


@pytest.mark.parametrize("pause,answer,done", [(False, None, False), (True, "n", False),
                                               (True, "y", True)])
def test_interrupt_follows_the_reference(fake, monkeypatch, pause, answer, done):
    """Ctrl-C during a stage (runner.py:434-451): cancelled, or — with pause_on_interrupt and
    the answer "y" — resumed; the data saved so far is kept either way."""
    s, eng = fake(solve_time=0.02, pause_on_interrupt=pause)
    eng.interrupt_at = 8
    asked = []
    monkeypatch.setattr("builtins.input", lambda prompt: (asked.append(prompt), answer)[1])
    old = signal.getsignal(signal.SIGINT)
    sol = s.solve()
    assert signal.getsignal(signal.SIGINT) is old            # handler restored
    assert bool(asked) == pause
    steps = [g["attrs"]["step"] for g in sol._saved.groups]
    if done:
        assert steps[-1] == _expected_updates(0.02, 1e-3) - 1
    else:
        assert steps == [0, 4, 8, 12]       # the interrupted chunk (steps 8..11) completed


def test_interrupt_while_thermalising_returns_none(fake, monkeypatch):
    s, eng = fake(skip_time=0.02, solve_time=0.01, pause_on_interrupt=False)
    eng.interrupt_at = 4
    assert s.solve() is None                                 # runner.py:313-314


# ------------------------------------------------------------------------------------------
# time-dependent inputs: what reaches the engine, and when

def _strip(monkeypatch):
    from tdgl_b200.synthetic import film_problem

    FakeEngine.instances.clear()
    monkeypatch.setattr(solver_mod, "DeviceEngine", FakeEngine)
    mesh, A, eps, terms = film_problem(8, 4, 0.5, terminals=True)
    return mesh, A, eps, terms


def test_callable_currents_step_one_at_a_time_and_upload_on_change(monkeypatch):
    """reference solver.py:325-345: J_ext,k = -(1 / L_k) sum_{j != k} I_j(t), written to the
    terminal's boundary edges, only when a density changed; a callable current makes the host
    wake up every step (runner.py:417-423 calls update once per step)."""
    mesh, A, eps, terms = _strip(monkeypatch)
    I = 0.3

    def cur(t):
        f = 1.0 if t < 3.5e-3 else 2.0
        return {"source": I * f, "drain": -I * f}

    o = tdgl.SolverOptions(solve_time=0.0065, dt_init=1e-3, adaptive=False, save_every=4)
    s = tdgl.TDGLSolver.from_dimensionless(mesh, o, A_applied=A, epsilon=eps, terminal_info=terms,
                                           terminal_currents=cur)
    eng = FakeEngine.instances[-1]
    s.solve()
    assert all(n == 1 for n, _, _ in eng.calls)                    # chunk = 1
    ups = [e[1] for e in eng.log if e[0] == "mub"]
    assert len(ups) == 2                                           # t = 0 and the jump at 4e-3
    for k, f in enumerate((1.0, 2.0)):
        want = np.zeros(len(mesh.edge_mesh.boundary_edge_indices))
        for t in terms:
            other = sum(v for n, v in cur(1.0 if f == 2.0 else 0.0).items() if n != t.name)
            want[np.asarray(t.boundary_edge_indices)] = -other / t.length
        np.testing.assert_array_equal(ups[k], want)
    # a dict of currents is uploaded once and the chunks run to the save steps
    s2 = tdgl.TDGLSolver.from_dimensionless(mesh, o, A_applied=A, epsilon=eps, terminal_info=terms,
                                            terminal_currents={"source": I, "drain": -I})
    eng2 = FakeEngine.instances[-1]
    s2.solve()
    assert [n for n, _, _ in eng2.calls] == [4, 4] and len([e for e in eng2.log if e[0] == "mub"]) == 1
    # non-conserving and unknown terminals are the reference's ValueErrors (solver.py:45-53)
    with pytest.raises(ValueError, match="sum of all terminal currents must be 0"):
        tdgl.TDGLSolver.from_dimensionless(mesh, o, A_applied=A, epsilon=eps, terminal_info=terms,
                                           terminal_currents={"source": I, "drain": 0.0})
    with pytest.raises(ValueError, match="Unknown terminal"):
        tdgl.TDGLSolver.from_dimensionless(mesh, o, A_applied=A, epsilon=eps, terminal_info=terms,
                                           terminal_currents={"source": I, "gate": -I})


def test_host_vector_potential_callback_follows_the_reference(monkeypatch):
    """reference solver.py:626-642: dA/dt = (A(t) - A_prev) / dt_prev projected on the
    normalised edge directions every step; the link variables are rebuilt only
    `if not allclose(A, A_prev)`."""
    mesh, A, eps, terms = _strip(monkeypatch)
    A1 = np.ones_like(A) * 0.05

    def A_of_t(t):          # constant up to 2.5e-3, then growing by 20 % per step
        return A1 * (1.0 + 0.2 * max(0.0, round((t - 2e-3) / 1e-3)))

    o = tdgl.SolverOptions(solve_time=0.0045, dt_init=1e-3, adaptive=False, save_every=100)
    s = tdgl.TDGLSolver.from_dimensionless(mesh, o, A_applied=A_of_t, epsilon=eps)
    eng = FakeEngine.instances[-1]
    sol = s.solve()
    links = [e[1] for e in eng.log if e[0] == "link"]
    dadt = [e[1] for e in eng.log if e[0] == "dadt"]
    n_updates = len(eng.calls)
    assert len(dadt) == n_updates and all(n == 1 for n, _, _ in eng.calls)
    d = np.asarray(mesh.edge_mesh.directions)
    nd = d / np.linalg.norm(d, axis=1)[:, None]
    times = [c[2] for c in eng.calls]
    prev = A_of_t(0.0)
    changes = 0
    for k, t in enumerate(times):
        cur = A_of_t(t)
        np.testing.assert_allclose(dadt[k], np.einsum("ij,ij->i", (cur - prev) / 1e-3, nd),
                                   rtol=0, atol=1e-15)
        changes += not np.allclose(cur, prev)
        prev = cur
    assert changes >= 2 and len(links) == 1 + changes      # setup + one rebuild per change
    # the vector potential of every saved step travels with the results (dynamic => per group)
    assert "applied_vector_potential" in sol._saved.groups[-1]
    assert "applied_vector_potential" not in sol._saved.fixed


def test_device_side_tables_are_handed_over_once(monkeypatch):
    """Separable inputs (f(t) A0(r), I_k(t) tables, eps0 + g(t) eps1) go to the engine at setup
    and the stage loop runs whole chunks (no per-step host callback)."""
    from tdgl_b200.sources import PiecewiseLinearCurrents, SeparableEpsilon

    mesh, A, eps, terms = _strip(monkeypatch)
    tk = np.array([0.0, 0.004, 0.01])
    table = PiecewiseLinearCurrents(tk, {"source": [0.1, 0.3, 0.3], "drain": [-0.1, -0.3, -0.3]})
    e = SeparableEpsilon(eps, -0.1 * eps, tk, np.array([0.0, 1.0, 1.0]))
    o = tdgl.SolverOptions(solve_time=0.0075, dt_init=1e-3, adaptive=False, save_every=4)
    s = tdgl.TDGLSolver.from_dimensionless(mesh, o, A_applied=A, epsilon=e, terminal_info=terms,
                                           terminal_currents=table, A_ramp=(tk, [0.0, 1.0, 1.0]))
    eng = FakeEngine.instances[-1]
    sol = s.solve()
    kinds = [x[0] for x in eng.log]
    assert kinds.count("ramp") == 1 and kinds.count("cur_table") == 1 and kinds.count("eps_table") == 1
    assert [c[1] for c in eng.calls] == [0, 4, 8]                  # whole chunks, no per-step callback
    ct = next(x for x in eng.log if x[0] == "cur_table")
    names = [t.name for t in s.terminal_info]
    np.testing.assert_array_equal(ct[4], np.array([table.knots[1][n] for n in names]))
    np.testing.assert_array_equal(ct[2], [t.length for t in s.terminal_info])
    for k, t in enumerate(s.terminal_info):
        assert np.all(ct[1][np.asarray(t.boundary_edge_indices)] == k)
    assert np.sum(ct[1] >= 0) == sum(len(t.boundary_edge_indices) for t in s.terminal_info)
    # saved epsilon / A follow the tables at the time of the last step taken
    g = sol._saved.groups[-1]
    t_last = g["attrs"]["time"]
    np.testing.assert_allclose(g["epsilon"], eps * (1 - 0.1 * min(t_last / 0.004, 1.0)), rtol=1e-12)
    np.testing.assert_allclose(g["applied_vector_potential"], A * min(t_last / 0.004, 1.0), rtol=1e-12)


def test_update_seam_tuple_layout(monkeypatch):
    """`SolverResult` is positional for the reference's Runner (`new_dt, *values`,
    runner.py:424-428): A_applied is present only for a dynamic vector potential, epsilon only
    for a dynamic epsilon, in that order (solver.py:708-714)."""
    mesh, A, eps, terms = _strip(monkeypatch)
    o = tdgl.SolverOptions(solve_time=1.0, dt_init=1e-3, adaptive=False)
    n, E = len(mesh.sites), len(mesh.edge_mesh.edges)
    for dyn_A, dyn_eps, length in ((False, False, 6), (True, False, 7), (False, True, 7), (True, True, 8)):
        s = tdgl.TDGLSolver.from_dimensionless(
            mesh, o, A_applied=(lambda t: A) if dyn_A else A,
            epsilon=(lambda t: eps) if dyn_eps else eps)
        res = s.update({"step": 0, "time": 0.0, "dt": 1e-3}, None, 1e-3,
                       psi=np.ones(n, complex), mu=np.zeros(n))
        got = [x for x in res if x is not None]
        assert len(got) == length
        assert res.dt == 1e-3 and res.psi.shape == (n,) and res.supercurrent.shape == (E,)
        assert res.A_induced.shape == (E, 2) and not res.A_induced.any()
        if dyn_A:
            assert got[6].shape == (E, 2)
        if dyn_eps:
            assert got[-1].shape == (n,)
