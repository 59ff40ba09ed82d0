"""``DeviceEngine``: thin Python owner of one ``tdgl_handle`` (see include/tdgl_b200.h).

It plays the role of the reference's ``MeshOperators`` (tdgl/finite_volume/operators.py:
233-394) plus the arrays ``TDGLSolver.update`` threads through the Runner — but all of
them live in HBM; Python only sees them at save steps.
"""

from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import as_c128, as_f64, as_i64, ptr


class StepFailed(RuntimeError):
    """|psi|^2 solve failed ``max_solve_retries`` times (reference RuntimeError,
    solver.py:478-483)."""


class ScreeningFailed(RuntimeError):
    """The screening iteration did not converge in ``max_iterations_per_step`` passes
    (reference RuntimeError, solver.py:657-663)."""


class AdvanceInfo(NamedTuple):
    steps_done: int
    step: int
    time: float
    dt: float
    tentative_dt: float
    finished: bool
    status: int
    failed_step: int
    failed_dt: float
    retries: int
    mu_iterations: int
    mu_rel_residual: float
    device_ms: float = 0.0
    screening_iterations: int = 0
    screening_error: float = 0.0


class _PinnedBlock:
    """Owner of one page-locked allocation (freed when the last array view dies)."""

    def __init__(self, nbytes: int):
        self._lib = _lib.load()
        self.ptr = self._lib.tdgl_host_alloc(int(nbytes))
        if not self.ptr:
            raise MemoryError(f"tdgl_host_alloc({nbytes}) failed")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                self._lib.tdgl_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _check_out(out, sizes) -> None:
    """Output arrays the library writes through raw pointers: (psi c128, mu f64, J_s f64, J_n
    f64), each None (not wanted) or a C-contiguous 1-D array of at least the given length."""
    if len(out) != 4:
        raise ValueError("out must be (psi, mu, supercurrent, normal_current)")
    names = ("psi", "mu", "supercurrent", "normal_current")
    dtypes = (np.complex128, np.float64, np.float64, np.float64)
    for a, n, name, dt in zip(out, sizes, names, dtypes):
        if a is None:
            continue
        if (not isinstance(a, np.ndarray) or a.dtype != dt or a.ndim != 1 or a.shape[0] < n
                or not a.flags.c_contiguous):
            raise ValueError(f"out[{name}] must be a C-contiguous {np.dtype(dt).name} array of"
                             f" length {n}")


def pinned_empty(shape, dtype) -> np.ndarray:
    """NumPy array in page-locked host memory: the copies of ``DeviceEngine.update`` from /
    to such arrays are true asynchronous DMA transfers."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    block = _PinnedBlock(max(n * dtype.itemsize, 16))
    buf = (C.c_char * block.nbytes).from_address(block.ptr)
    buf._tdgl_block = block  # the array's base keeps the allocation alive
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


class DeviceEngine:
    def __init__(self, mesh, *, fixed_sites: Optional[Sequence[int]] = None,
                 fix_psi: bool = True, gamma: float = 10.0, u: float = 5.79,
                 probe_sites: Optional[Sequence[int]] = None, device: int = 0,
                 mu_rtol: float = 0.0, mu_max_iter: int = 0, amg_theta: float = 0.0,
                 amg_max_coarse: int = 0, use_graph: int = 0, reorder: int = 0,
                 running_capacity: int = 0, world: int = 1, rank: int = 0,
                 replicate_below: int = 0, fuse_coarse: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        em = mesh.edge_mesh
        self.n_sites = len(mesh.sites)
        self.n_edges = len(em.edges)
        self.n_boundary_edges = len(em.boundary_edge_indices)
        self.n_probe = 0 if probe_sites is None else len(probe_sites)
        edges = as_i64(em.edges)
        areas = as_f64(mesh.areas, (self.n_sites,))
        elen = as_f64(em.edge_lengths, (self.n_edges,))
        dlen = as_f64(em.dual_edge_lengths, (self.n_edges,))
        dirs = as_f64(em.directions, (self.n_edges, 2))
        bidx = as_i64(em.boundary_edge_indices)
        fixed = as_i64([] if fixed_sites is None else fixed_sites)
        probes = as_i64([] if probe_sites is None else probe_sites)
        xy = as_f64(mesh.sites, (self.n_sites, 2))
        cfg = _lib.tdgl_config()
        cfg.struct_size = C.sizeof(_lib.tdgl_config)
        cfg.device = device
        cfg.mu_rtol = mu_rtol
        cfg.mu_max_iter = mu_max_iter
        cfg.amg_theta = amg_theta
        cfg.amg_max_coarse = amg_max_coarse
        cfg.use_graph = use_graph
        cfg.reorder = reorder
        cfg.running_capacity = running_capacity
        cfg.world = world
        cfg.rank = rank
        cfg.replicate_below = replicate_below
        cfg.fuse_coarse = fuse_coarse
        self.world, self.rank = int(world), int(rank)
        self.running_capacity = running_capacity or 4096
        rc = self._lib.tdgl_create(
            C.byref(self._h), self.n_sites, self.n_edges, self.n_boundary_edges, ptr(edges),
            ptr(areas), ptr(elen), ptr(dlen), ptr(dirs), ptr(bidx), ptr(fixed), len(fixed),
            1 if fix_psi else 0, ptr(xy), float(gamma), float(u), ptr(probes), len(probes),
            C.byref(cfg))
        if rc != _lib.TDGL_OK:
            msg = self._lib.tdgl_last_error(None).decode()
            self._h = C.c_void_p()
            raise _lib.TDGLLibraryError(f"tdgl_create failed ({rc}): {msg}")

    # -- lifetime -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.tdgl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc == _lib.TDGL_OK:
            return
        msg = self._lib.tdgl_last_error(self._h).decode()
        if rc == _lib.TDGL_E_INVALID:
            raise ValueError(msg)
        raise _lib.TDGLLibraryError(f"tdgl_b200 error {rc}: {msg}")

    @property
    def message(self) -> str:
        return self._lib.tdgl_last_error(self._h).decode()

    # -- sharding: wiring of the peer arenas (include/tdgl_b200.h, "domain decomposition") ---
    def comm_export(self) -> bytes:
        """64-byte CUDA IPC handle of this shard's arena (to be all-gathered)."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.tdgl_comm_export(self._h, buf))
        return buf.raw

    def comm_connect_ipc(self, handles: Sequence[bytes]) -> None:
        """``handles``: the exported handles of all shards in rank order (other processes)."""
        blob = b"".join(handles)
        if len(blob) != 64 * self.world:
            raise ValueError("need one 64-byte handle per shard")
        self._check(self._lib.tdgl_comm_connect_ipc(self._h, blob, self.world))

    def comm_connect_local(self, engines: Sequence["DeviceEngine"]) -> None:
        """``engines``: all shards in rank order, living in this process."""
        arr = (C.c_void_p * len(engines))(*[e._h.value for e in engines])
        self._check(self._lib.tdgl_comm_connect_local(self._h, arr, len(engines)))

    def shard_info(self) -> dict:
        out = (C.c_int64 * 8)()
        self._check(self._lib.tdgl_shard_info(self._h, out, 8))
        keys = ["world", "rank", "n_sites", "n_owned", "n_halo0", "n_halo_all_levels",
                "neighbours0", "n_send0"]
        return dict(zip(keys, [int(v) for v in out]))

    # -- inputs ---------------------------------------------------------------------------
    def set_link_exponents(self, A) -> None:
        self._check(self._lib.tdgl_set_link_exponents(self._h, ptr(as_f64(A, (self.n_edges, 2)))))

    def set_epsilon(self, eps) -> None:
        self._check(self._lib.tdgl_set_epsilon(self._h, ptr(as_f64(eps, (self.n_sites,)))))

    def set_mu_boundary(self, mu_boundary) -> None:
        self._check(self._lib.tdgl_set_mu_boundary(
            self._h, ptr(as_f64(mu_boundary, (self.n_boundary_edges,)))))

    def set_dA_dt(self, dA_dt) -> None:
        """``dA_dt`` [E] (or None for a static vector potential), solver.py:626-634."""
        arr = None if dA_dt is None else as_f64(dA_dt, (self.n_edges,))
        self._check(self._lib.tdgl_set_dA_dt(self._h, ptr(arr)))

    def set_vector_potential_ramp(self, A0, t_knots, f_knots) -> None:
        """A(r, t) = f(t) * A0(r) evaluated on the device (f piecewise linear); ``A0`` None or
        no knots turns it off."""
        if A0 is None or len(t_knots) == 0:
            self._check(self._lib.tdgl_set_vector_potential_ramp(self._h, None, 0, None, None))
            return
        t, f = as_f64(t_knots), as_f64(f_knots)
        if t.shape != f.shape or t.ndim != 1:
            raise ValueError("t_knots and f_knots must be 1-D arrays of the same length")
        self._check(self._lib.tdgl_set_vector_potential_ramp(
            self._h, ptr(as_f64(A0, (self.n_edges, 2))), len(t), ptr(t), ptr(f)))

    def set_terminal_current_table(self, terminal_of_boundary_edge, lengths, t_knots, currents) -> None:
        """Device-side I_k(t) tables (``currents`` [n_terminals, n_knots], J_scale-d); empty
        ``t_knots`` turns them off."""
        if t_knots is None or len(t_knots) == 0:
            self._check(self._lib.tdgl_set_terminal_current_table(self._h, 0, None, None, 0, None, None))
            return
        term = np.ascontiguousarray(terminal_of_boundary_edge, dtype=np.int32)
        if term.shape != (self.n_boundary_edges,):
            raise ValueError("one terminal index per boundary edge")
        L, t = as_f64(lengths), as_f64(t_knots)
        cur = as_f64(currents, (len(L), len(t)))
        self._check(self._lib.tdgl_set_terminal_current_table(
            self._h, len(L), ptr(term), ptr(L), len(t), ptr(t), ptr(cur)))

    def set_epsilon_table(self, eps0, eps1, t_knots, g_knots) -> None:
        """epsilon(r, t) = eps0 + g(t) eps1 evaluated on the device; empty knots: off."""
        if t_knots is None or len(t_knots) == 0:
            self._check(self._lib.tdgl_set_epsilon_table(self._h, None, None, 0, None, None))
            return
        t, g = as_f64(t_knots), as_f64(g_knots)
        self._check(self._lib.tdgl_set_epsilon_table(
            self._h, ptr(as_f64(eps0, (self.n_sites,))), ptr(as_f64(eps1, (self.n_sites,))),
            len(t), ptr(t), ptr(g)))

    def set_screening(self, scale: float, sites_xy, edge_centers, *, tolerance: float,
                      max_iterations: int, step_size: float, step_drag: float) -> None:
        """Turn on the screening iteration (include/tdgl_b200.h, tdgl_set_screening)."""
        self._check(self._lib.tdgl_set_screening(
            self._h, 1, float(scale), ptr(as_f64(sites_xy, (self.n_sites, 2))),
            ptr(as_f64(edge_centers, (self.n_edges, 2))), float(tolerance), int(max_iterations),
            float(step_size), float(step_drag)))

    def set_induced_vector_potential(self, A) -> None:
        self._check(self._lib.tdgl_set_induced_vector_potential(
            self._h, ptr(as_f64(A, (self.n_edges, 2)))))

    def get_induced_vector_potential(self) -> np.ndarray:
        A = np.empty((self.n_edges, 2))
        self._check(self._lib.tdgl_get_induced_vector_potential(self._h, ptr(A)))
        return A

    def get_running_screening(self, steps: int) -> np.ndarray:
        its = np.zeros(max(int(steps), 1), dtype=np.int64)
        self._check(self._lib.tdgl_get_running_screening(self._h, len(its), ptr(its)))
        return its[:steps]

    def set_state(self, psi, mu) -> None:
        self._check(self._lib.tdgl_set_state(
            self._h, ptr(as_c128(psi, (self.n_sites,))), ptr(as_f64(mu, (self.n_sites,)))))

    def set_stepper(self, *, dt_init, dt_max, adaptive=True, adaptive_window=10,
                    max_solve_retries=10, adaptive_time_step_multiplier=0.25) -> None:
        self._check(self._lib.tdgl_set_stepper(
            self._h, float(dt_init), float(dt_max), int(bool(adaptive)), int(adaptive_window),
            int(max_solve_retries), float(adaptive_time_step_multiplier)))

    # -- stepping -------------------------------------------------------------------------
    def advance(self, max_steps: int, t_end: float, step: int, time: float) -> AdvanceInfo:
        info = _lib.tdgl_advance_info()
        rc = self._lib.tdgl_advance(self._h, int(max_steps), float(t_end), int(step),
                                    float(time), C.byref(info))
        out = AdvanceInfo(info.steps_done, info.step, info.time, info.dt, info.tentative_dt,
                          bool(info.finished), info.status, info.failed_step, info.failed_dt,
                          info.retries, info.mu_iterations, info.mu_rel_residual, info.device_ms,
                          info.screening_iterations, info.screening_error)
        if rc == _lib.TDGL_E_STEP_FAILED:
            raise StepFailed(f"step {out.failed_step} dt {out.failed_dt:.2e}", out)
        if out.status == 4:
            raise ScreeningFailed(f"step {out.failed_step}", out)
        self._check(rc)
        return out

    def update(self, psi, mu, step: int, time: float, out=None):
        """One ``TDGLSolver.update`` at the reference's step seam: host arrays in, one step
        on the device, host arrays out.  ``out`` = (psi, mu, supercurrent, normal_current)
        arrays to write into (e.g. pinned); allocated if omitted."""
        if out is None:
            out = (np.empty(self.n_sites, np.complex128), np.empty(self.n_sites),
                   np.empty(self.n_edges), np.empty(self.n_edges))
        else:
            _check_out(out, (self.n_sites, self.n_sites, self.n_edges, self.n_edges))
        info = _lib.tdgl_advance_info()
        psi = as_c128(psi, (self.n_sites,))
        mu = as_f64(mu, (self.n_sites,))
        rc = self._lib.tdgl_update(self._h, ptr(psi), ptr(mu), int(step), float(time),
                                   ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]),
                                   C.byref(info))
        res = AdvanceInfo(info.steps_done, info.step, info.time, info.dt, info.tentative_dt,
                          bool(info.finished), info.status, info.failed_step, info.failed_dt,
                          info.retries, info.mu_iterations, info.mu_rel_residual, info.device_ms,
                          info.screening_iterations, info.screening_error)
        if rc == _lib.TDGL_E_STEP_FAILED:
            raise StepFailed(f"step {res.failed_step} dt {res.failed_dt:.2e}", res)
        if res.status == 4:
            raise ScreeningFailed(f"step {res.failed_step}", res)
        self._check(rc)
        return res, out

    def local_maps(self):
        """(owned_sites, owned_edges): caller ids of the sites / edges this shard owns, in the
        order of the local arrays of ``update_local`` (all of them for a single shard)."""
        sizes = np.zeros(2, dtype=np.int64)
        self._check(self._lib.tdgl_local_maps(self._h, ptr(sizes), None, None))
        sites = np.zeros(max(int(sizes[0]), 1), dtype=np.int64)
        edges = np.zeros(max(int(sizes[1]), 1), dtype=np.int64)
        self._check(self._lib.tdgl_local_maps(self._h, ptr(sizes), ptr(sites), ptr(edges)))
        return sites[:sizes[0]], edges[:sizes[1]]

    def update_local(self, psi_local, mu_local, step: int, time: float, out):
        """The step seam of one shard: psi / mu of the OWNED sites in, psi', mu' of the owned
        sites and J_s, J_n of the owned edges out (``out`` = four arrays of those sizes, e.g.
        pinned).  Every shard must make the call for every step."""
        if getattr(self, "_local_sizes", None) is None:
            sizes = np.zeros(2, dtype=np.int64)
            self._check(self._lib.tdgl_local_maps(self._h, ptr(sizes), None, None))
            self._local_sizes = (int(sizes[0]), int(sizes[1]))
        ns, ne = self._local_sizes
        psi_local = as_c128(psi_local, (ns,))
        mu_local = as_f64(mu_local, (ns,))
        _check_out(out, (ns, ns, ne, ne))
        info = _lib.tdgl_advance_info()
        rc = self._lib.tdgl_update_local(self._h, ptr(psi_local), ptr(mu_local), int(step),
                                         float(time), ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                         ptr(out[3]), C.byref(info))
        res = AdvanceInfo(info.steps_done, info.step, info.time, info.dt, info.tentative_dt,
                          bool(info.finished), info.status, info.failed_step, info.failed_dt,
                          info.retries, info.mu_iterations, info.mu_rel_residual, info.device_ms,
                          info.screening_iterations, info.screening_error)
        if rc == _lib.TDGL_E_STEP_FAILED:
            raise StepFailed(f"step {res.failed_step} dt {res.failed_dt:.2e}", res)
        self._check(rc)
        return res, out

    # -- outputs --------------------------------------------------------------------------
    def get_state(self):
        psi = np.empty(self.n_sites, dtype=np.complex128)
        mu = np.empty(self.n_sites, dtype=np.float64)
        self._check(self._lib.tdgl_get_state(self._h, ptr(psi), ptr(mu)))
        return psi, mu

    def get_currents(self):
        js = np.empty(self.n_edges)
        jn = np.empty(self.n_edges)
        self._check(self._lib.tdgl_get_currents(self._h, ptr(js), ptr(jn)))
        return js, jn

    def snapshot_begin(self, slot: int) -> None:
        """Start an asynchronous copy of psi, mu, J_s, J_n into pinned host slot 0 / 1."""
        self._check(self._lib.tdgl_snapshot_begin(self._h, int(slot)))

    def snapshot_wait(self, slot: int):
        """Block until the slot's copy has landed; returns NumPy views of the pinned buffers
        (psi, mu, J_s, J_n), valid until the slot's next ``snapshot_begin``.  Safe to call
        from a writer thread while the stepping thread keeps calling ``advance``."""
        ptrs = [C.c_void_p() for _ in range(4)]
        rc = self._lib.tdgl_snapshot_wait(self._h, int(slot), *[C.byref(p) for p in ptrs])
        if rc != _lib.TDGL_OK:
            raise _lib.TDGLLibraryError(f"tdgl_snapshot_wait failed ({rc})")

        def view(p, n, dtype):
            buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(p.value)
            return np.frombuffer(buf, dtype=dtype, count=n)

        return (view(ptrs[0], self.n_sites, np.complex128), view(ptrs[1], self.n_sites, np.float64),
                view(ptrs[2], self.n_edges, np.float64), view(ptrs[3], self.n_edges, np.float64))

    def get_running(self, steps: int):
        cap = max(int(steps), 1)
        dt = np.zeros(cap)
        mu = np.zeros((max(self.n_probe, 1), cap))
        th = np.zeros((max(self.n_probe, 1), cap))
        self._check(self._lib.tdgl_get_running(self._h, cap, ptr(dt), ptr(mu), ptr(th)))
        return dt[:steps], mu[: self.n_probe, :steps], th[: self.n_probe, :steps]

    # -- single operators -------------------------------------------------------------------
    def psi_laplacian(self, x):
        y = np.empty(self.n_sites, dtype=np.complex128)
        self._check(self._lib.tdgl_op_psi_laplacian(self._h, ptr(as_c128(x, (self.n_sites,))), ptr(y)))
        return y

    def psi_step(self, psi, mu, dt):
        out = np.empty(self.n_sites, dtype=np.complex128)
        sq = np.empty(self.n_sites)
        failed = C.c_int32(0)
        self._check(self._lib.tdgl_op_psi_step(
            self._h, ptr(as_c128(psi, (self.n_sites,))), ptr(as_f64(mu, (self.n_sites,))),
            float(dt), ptr(out), ptr(sq), C.byref(failed)))
        return out, sq, bool(failed.value)

    def mu_rhs(self, psi):
        rhs = np.empty(self.n_sites)
        self._check(self._lib.tdgl_op_mu_rhs(self._h, ptr(as_c128(psi, (self.n_sites,))), ptr(rhs)))
        return rhs

    def mu_laplacian(self, x):
        y = np.empty(self.n_sites)
        self._check(self._lib.tdgl_op_mu_laplacian(self._h, ptr(as_f64(x, (self.n_sites,))), ptr(y)))
        return y

    def mu_solve(self, rhs):
        mu = np.empty(self.n_sites)
        it = C.c_int32(0)
        rr = C.c_double(0)
        self._check(self._lib.tdgl_op_mu_solve(self._h, ptr(as_f64(rhs, (self.n_sites,))), ptr(mu),
                                               C.byref(it), C.byref(rr)))
        return mu, it.value, rr.value

    def time_kernel(self, which: int, reps: int = 20, flush_l2: bool = True) -> float:
        ms = C.c_double(0)
        self._check(self._lib.tdgl_time_kernel(self._h, int(which), int(reps),
                                               1 if flush_l2 else 0, C.byref(ms)))
        return ms.value

    def time_cusparse(self, which: int, reps: int = 20, flush_l2: bool = True) -> float:
        """cusparseSpMV on the engine's CSR arrays (comparator; 0: real f64, 1: complex128)."""
        ms = C.c_double(0)
        self._check(self._lib.tdgl_time_cusparse(self._h, int(which), int(reps),
                                                 1 if flush_l2 else 0, C.byref(ms)))
        return ms.value

    def get_trace(self, capacity: int = 512):
        """In-loop timeline of the CG iteration's kernels (``tdgl_get_trace``; needs the engine
        to have been created under ``TDGL_B200_TRACE``): list of dicts with name, rows, in / go /
        out in microseconds from the earliest launch of the pass, and the pass count."""
        n = C.c_int32(0)
        names = C.create_string_buffer(64 * capacity)
        t_in = np.zeros(capacity)
        t_go = np.zeros(capacity)
        t_out = np.zeros(capacity)
        cnt = np.zeros(capacity, dtype=np.int64)
        self._check(self._lib.tdgl_get_trace(self._h, capacity, C.byref(n), names, ptr(t_in),
                                             ptr(t_go), ptr(t_out), ptr(cnt)))
        out = []
        for k in range(n.value):
            label = names.raw[64 * k:64 * (k + 1)].split(b"\0", 1)[0].decode()
            name, _, rows = label.partition("rows=")
            out.append(dict(name=name.strip(), rows=int(rows) if rows.strip() else 0,
                            t_in=float(t_in[k]), t_go=float(t_go[k]), t_out=float(t_out[k]),
                            count=int(cnt[k])))
        return out

    def info(self) -> dict:
        out = (C.c_int64 * 8)()
        self._check(self._lib.tdgl_get_info(self._h, out, 8))
        keys = ["n_sites", "n_edges", "nnz", "amg_levels", "amg_nnz", "amg_coarsest",
                "launches", "graph_mode"]
        return dict(zip(keys, [int(v) for v in out]))


def host_shard_probe(mesh, world: int, rhs=None, theta=0.0, max_coarse=0, max_iter=200,
                     rtol=1e-10, replicate_below=0):
    """Host-only: the domain decomposition the sharded engine builds for ``world`` shards,
    and (optionally) the sharded AMG-PCG emulated in this process."""
    lib = _lib.load()
    em = mesh.edge_mesh
    n, E = len(mesh.sites), len(em.edges)
    nl = C.c_int32(0)
    off = np.zeros((32, 9), dtype=np.int64)
    halo = np.zeros((32, 8), dtype=np.int64)
    perm = np.zeros(n, dtype=np.int64)
    it = C.c_int32(0)
    x = None
    if rhs is not None:
        rhs = as_f64(rhs, (n,))
        x = np.zeros(n)
    rc = lib.tdgl_host_shard_probe(n, E, ptr(as_i64(em.edges)), ptr(as_f64(em.edge_lengths)),
                                   ptr(as_f64(em.dual_edge_lengths)),
                                   ptr(as_f64(mesh.sites, (n, 2))), world, theta, max_coarse,
                                   replicate_below, C.byref(nl), ptr(off), ptr(halo), ptr(perm), ptr(rhs), ptr(x),
                                   max_iter, rtol, C.byref(it))
    if rc != 0:
        raise _lib.TDGLLibraryError(lib.tdgl_last_error(None).decode())
    L = nl.value
    return dict(levels=L, rep=int(off[31, 0]), offsets=off[:L, :world + 1].copy(),
                halo_sizes=halo[:L, :world].copy(),
                perm=perm, x=x, iterations=it.value)


def host_shard_lists(mesh, world: int, rank: int):
    """Host-only: level-0 exchange lists of one shard (caller site ids): owned, halo, and
    ``send[q]`` = the owned sites sent to shard q, ordered as q's halo."""
    lib = _lib.load()
    em = mesh.edge_mesh
    n, E = len(mesh.sites), len(em.edges)
    args = (n, E, ptr(as_i64(em.edges)), ptr(as_f64(em.edge_lengths)),
            ptr(as_f64(em.dual_edge_lengths)), ptr(as_f64(mesh.sites, (n, 2))), world, rank)
    counts = np.zeros(3, dtype=np.int64)
    rc = lib.tdgl_host_shard_lists(*args, ptr(counts), None, None, None, None)
    if rc != 0:
        raise _lib.TDGLLibraryError(lib.tdgl_last_error(None).decode())
    owned = np.zeros(counts[0], dtype=np.int64)
    halo = np.zeros(max(counts[1], 1), dtype=np.int64)
    send_ptr = np.zeros(world + 1, dtype=np.int64)
    send = np.zeros(max(counts[2], 1), dtype=np.int64)
    rc = lib.tdgl_host_shard_lists(*args, ptr(counts), ptr(owned), ptr(halo), ptr(send_ptr),
                                   ptr(send))
    if rc != 0:
        raise _lib.TDGLLibraryError(lib.tdgl_last_error(None).decode())
    return dict(owned=owned, halo=halo[:counts[1]],
                send=[send[send_ptr[q]:send_ptr[q + 1]] for q in range(world)])


def host_amg_probe(mesh, theta=0.08, max_coarse=200, rhs=None, max_iter=200, rtol=1e-10):
    """Host-only: build the AMG hierarchy in C++ and (optionally) run host PCG with it."""
    lib = _lib.load()
    em = mesh.edge_mesh
    n, E = len(mesh.sites), len(em.edges)
    nl = C.c_int32(0)
    rows = (C.c_int64 * 32)()
    nnz = (C.c_int64 * 32)()
    it = C.c_int32(0)
    x = None
    if rhs is not None:
        rhs = as_f64(rhs, (n,))
        x = np.zeros(n)
    rc = lib.tdgl_host_amg_probe(n, E, ptr(as_i64(em.edges)), ptr(as_f64(em.edge_lengths)),
                                 ptr(as_f64(em.dual_edge_lengths)), theta, max_coarse,
                                 C.byref(nl), rows, nnz, ptr(rhs), ptr(x), max_iter, rtol,
                                 C.byref(it))
    if rc != 0:
        raise _lib.TDGLLibraryError(lib.tdgl_last_error(None).decode())
    L = nl.value
    return dict(levels=L, rows=[int(rows[i]) for i in range(L)],
                nnz=[int(nnz[i]) for i in range(L)], x=x, iterations=it.value)
