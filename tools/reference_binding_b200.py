# tdgl/solver/b200.py  (new file in the reference; kept here as tools/reference_binding_b200.py
# and executed by tests/test_integration_stub.py)
"""ctypes binding of libtdgl_b200.so at pyTDGL's step seam.

``B200Step(solver).update`` has the signature and return value of ``TDGLSolver.update``
(tdgl/solver/solver.py:580-714) and is what ``Runner`` would be handed as ``function``
(solver.py:791-804) when the B200 engine is selected.
"""
import ctypes as C
import os

import numpy as np

_LIB_PATH = os.environ.get("TDGL_B200_LIB", "libtdgl_b200.so")  # built by `python __graft_entry__.py`
_lib = C.CDLL(_LIB_PATH)
_lib.tdgl_last_error.restype = C.c_char_p          # (a c_int default would truncate the pointer)
_lib.tdgl_last_error.argtypes = [C.c_void_p]
for _name in ("tdgl_create", "tdgl_set_link_exponents", "tdgl_set_epsilon", "tdgl_set_mu_boundary",
              "tdgl_set_dA_dt", "tdgl_set_stepper", "tdgl_update"):
    getattr(_lib, _name).restype = C.c_int
_lib.tdgl_destroy.restype = None
_lib.tdgl_destroy.argtypes = [C.c_void_p]


class _Info(C.Structure):                  # tdgl_advance_info, include/tdgl_b200.h
    _fields_ = [("steps_done", C.c_int64), ("step", C.c_int64), ("time", C.c_double),
                ("dt", C.c_double), ("tentative_dt", C.c_double), ("finished", C.c_int32),
                ("status", C.c_int32), ("failed_step", C.c_int64), ("failed_dt", C.c_double),
                ("retries", C.c_int64), ("mu_iterations", C.c_int64),
                ("mu_rel_residual", C.c_double), ("device_ms", C.c_double),
                ("screening_iterations", C.c_int64), ("screening_error", C.c_double)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, t):
    return np.ascontiguousarray(a, t)


class B200Step:
    """Replaces MeshOperators + TDGLSolver.update for one TDGLSolver instance."""

    def __init__(self, solver, mesh=None, result_type=None):
        # solver: a fully constructed tdgl.TDGLSolver; mesh defaults to solver.device.mesh
        mesh = solver.device.mesh if mesh is None else mesh
        em, o = mesh.edge_mesh, solver.options
        fixed = _c(solver.operators.fixed_sites, np.int64)                  # solver.py:258-262
        probes = _c(solver.probe_points if solver.probe_points is not None else [], np.int64)
        self._keep = [_c(em.edges, np.int64), _c(mesh.areas, float), _c(em.edge_lengths, float),
                      _c(em.dual_edge_lengths, float), _c(em.directions, float),
                      _c(em.boundary_edge_indices, np.int64), fixed, _c(mesh.sites, float), probes]
        e, a, l, d, dr, b, fx, xy, pr = self._keep
        self.h = C.c_void_p()
        rc = _lib.tdgl_create(C.byref(self.h), C.c_int64(len(mesh.sites)), C.c_int64(len(e)),
                              C.c_int64(len(b)), _p(e), _p(a), _p(l), _p(d), _p(dr), _p(b),
                              _p(fx), C.c_int64(len(fx)), C.c_int32(o.terminal_psi is not None),
                              _p(xy), C.c_double(solver.gamma), C.c_double(solver.u), _p(pr),
                              C.c_int64(len(pr)), None)
        if rc:
            raise RuntimeError(f"tdgl_create failed ({rc}): " + _lib.tdgl_last_error(None).decode())
        self._check(_lib.tdgl_set_link_exponents(self.h, _p(_c(solver.current_A_applied, float))))  # operators.py:310-383
        self._check(_lib.tdgl_set_epsilon(self.h, _p(_c(solver.epsilon, float))))                    # solver.py:191-216
        self._check(_lib.tdgl_set_stepper(
            self.h, C.c_double(o.dt_init), C.c_double(o.dt_max), C.c_int32(o.adaptive),
            C.c_int32(o.adaptive_window), C.c_int32(o.max_solve_retries),
            C.c_double(o.adaptive_time_step_multiplier)))
        self.solver = solver
        self.n, self.ne = len(mesh.sites), len(e)
        self._mu_boundary_sent = None
        if result_type is None:
            from .solver import SolverResult as result_type   # noqa: N813  (inside the reference)
        self._result = result_type

    def __del__(self):
        if getattr(self, "h", None) is not None and self.h.value:
            _lib.tdgl_destroy(self.h)
            self.h = C.c_void_p()

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"tdgl_b200 error {rc}: " + _lib.tdgl_last_error(self.h).decode())

    def update(self, state, running_state, dt, *, psi, mu, supercurrent, normal_current,
               induced_vector_potential, applied_vector_potential=None, epsilon=None):
        """Same signature and return value as TDGLSolver.update (solver.py:580-714)."""
        s = self.solver
        time = state["time"]
        s.update_mu_boundary(time)                               # solver.py:325-345 (host)
        mub = _c(s.mu_boundary, float)
        if self._mu_boundary_sent is None or not np.array_equal(mub, self._mu_boundary_sent):
            self._check(_lib.tdgl_set_mu_boundary(self.h, _p(mub)))
            self._mu_boundary_sent = mub.copy()
        current_A = applied_vector_potential
        if s.dynamic_vector_potential:                           # solver.py:626-642
            current_A = s.update_applied_vector_potential(time)
            dA_dt = np.einsum("ij, ij -> i", (current_A - applied_vector_potential) / dt,
                              s.normalized_directions)
            if not np.allclose(current_A, s.current_A_applied):
                self._check(_lib.tdgl_set_link_exponents(self.h, _p(_c(current_A, float))))
            self._check(_lib.tdgl_set_dA_dt(self.h, _p(_c(dA_dt, float))))
            s.current_A_applied = current_A
        if s.dynamic_epsilon:                                    # solver.py:644-646
            epsilon = s.update_epsilon(time)
            self._check(_lib.tdgl_set_epsilon(self.h, _p(_c(epsilon, float))))
        psi1, mu1 = np.empty(self.n, complex), np.empty(self.n)
        js, jn = np.empty(self.ne), np.empty(self.ne)
        info = _Info()
        rc = _lib.tdgl_update(self.h, _p(_c(psi, complex)), _p(_c(mu, float)),
                              C.c_int64(state["step"]), C.c_double(time), _p(psi1), _p(mu1),
                              _p(js), _p(jn), C.byref(info))
        if rc == 1:                                              # TDGL_E_STEP_FAILED
            raise RuntimeError(f"Solver failed to converge in {s.options.max_solve_retries}"
                               f" retries at step {info.failed_step} with dt = {info.failed_dt:.2e}."
                               " Try using a smaller dt_init.")   # solver.py:479-483
        self._check(rc)
        running_state.append("dt", info.dt)                      # solver.py:690-694
        if s.probe_points is not None:
            running_state.append("mu", mu1[s.probe_points])
            running_state.append("theta", np.angle(psi1[s.probe_points]))
        results = [info.dt, psi1, mu1, js, jn, induced_vector_potential]
        if s.dynamic_vector_potential:                           # solver.py:708-714
            results.append(current_A)
        if s.dynamic_epsilon:
            results.append(epsilon)
        return self._result(*results)
