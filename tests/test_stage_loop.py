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
        FakeEngine.instances.append(self)

    def set_link_exponents(self, A): pass
    def set_epsilon(self, eps): pass
    def set_mu_boundary(self, mub): pass
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
