"""Domain-decomposed engines (no counterpart in the reference).

The mesh is cut into ``world`` shards (contiguous ranges of the Z-order site numbering, on
every AMG level); each shard is one ``tdgl_handle`` on one GPU.  Halo exchanges and scalar
all-reduces of the step are device code on the peers' memory (``csrc/comm.cuh``); the host
side only wires the shards once and sums the whole-mesh outputs at save steps.

``DistributedEngine``  one shard per process (``torchrun``): ``torch.distributed`` carries
                       the CUDA IPC handles at setup and the sums at save steps.
``LocalShardGroup``    all shards in this process, one host thread each (several GPUs with
                       peer access, or several shards on one GPU for tests).

Both expose the methods of ``DeviceEngine`` that ``TDGLSolver`` uses, with whole-mesh
inputs and outputs, so the solver code is the same for 1 and N GPUs.
"""

from __future__ import annotations

import threading
from typing import List, Optional, Sequence

import numpy as np

from ._lib import ptr
from .engine import AdvanceInfo, DeviceEngine


class DistributedEngine(DeviceEngine):
    """This process's shard; every rank must make the same calls in the same order."""

    def __init__(self, mesh, *, group=None, device: Optional[int] = None,
                 allow_shared_device: bool = False, **kw):
        import socket

        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self._dist, self._torch, self._group = dist, torch, group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if device is None:
            device = torch.cuda.current_device()
        if world > 1 and not allow_shared_device:
            # one shard per GPU: two ranks on one device would serialise (and the NCCL
            # communicator of the output sums is bound to the rank's own device)
            where: List[Optional[tuple]] = [None] * world
            dist.all_gather_object(where, (socket.gethostname(), int(device)), group=group)
            if len(set(where)) != world:
                raise RuntimeError(
                    f"ranks share a CUDA device {where}: call torch.cuda.set_device(LOCAL_RANK)"
                    " before building the solver (or set SolverOptions.cuda_device per rank);"
                    " pass allow_shared_device=True to shard one GPU on purpose")
        super().__init__(mesh, device=device, world=world, rank=rank, **kw)
        self._reduce_device = (torch.device("cuda", device)
                               if dist.get_backend(group) == "nccl" else torch.device("cpu"))
        if world > 1:
            handles: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(handles, self.comm_export(), group=group)
            self.comm_connect_ipc(handles)
            dist.barrier(group=group)

    def _sum(self, *arrays: np.ndarray):
        """In-place sum over the shards (each holds its own entries, zeros elsewhere)."""
        if self.world == 1:
            return arrays
        torch = self._torch
        for a in arrays:
            flat = a.view(np.float64).reshape(-1)
            t = torch.from_numpy(flat).to(self._reduce_device)
            self._dist.all_reduce(t, group=self._group)
            flat[:] = t.cpu().numpy()
        return arrays

    def _outputs_on_device(self, what: int, psi=None, mu=None, js=None, jn=None):
        """Sum the staged whole-mesh outputs over the shards on the devices (NCCL over
        NVLink), then fetch them once."""
        import ctypes as C

        torch = self._torch
        ptrs = (C.c_void_p * 4)()
        counts = (C.c_int64 * 4)()
        self._check(self._lib.tdgl_stage_outputs(self._h, what, ptrs, counts))

        class _Dev:  # zero-copy view of an engine buffer
            def __init__(self, p, n):
                self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8",
                                                 "data": (int(p), False), "version": 3}

        want = [k for k, a in enumerate((psi, mu, js, jn)) if a is not None]
        for k in want:
            t = torch.as_tensor(_Dev(ptrs[k], counts[k]), device=self._reduce_device)
            self._dist.all_reduce(t, group=self._group)
        torch.cuda.current_stream(self._reduce_device).synchronize()
        self._check(self._lib.tdgl_fetch_outputs(self._h, ptr(psi), ptr(mu), ptr(js), ptr(jn)))

    @property
    def _device_reduce(self) -> bool:
        return self.world > 1 and self._reduce_device.type == "cuda"

    def set_state(self, psi, mu) -> None:
        # every shard fills its own halo mailboxes from the whole-mesh arrays; no shard may
        # still be stepping (and storing into a peer's mailbox) while that happens
        if self.world > 1:
            self._dist.barrier(group=self._group)
        super().set_state(psi, mu)
        if self.world > 1:
            self._dist.barrier(group=self._group)

    def get_state(self):
        if not self._device_reduce:
            return self._sum(*super().get_state())
        psi = np.empty(self.n_sites, dtype=np.complex128)
        mu = np.empty(self.n_sites)
        self._outputs_on_device(1, psi=psi, mu=mu)
        return psi, mu

    def get_currents(self):
        if not self._device_reduce:
            return self._sum(*super().get_currents())
        js, jn = np.empty(self.n_edges), np.empty(self.n_edges)
        self._outputs_on_device(2, js=js, jn=jn)
        return js, jn

    def get_running(self, steps: int):
        dt, mu, th = super().get_running(steps)
        if self.n_probe:
            mu, th = np.ascontiguousarray(mu), np.ascontiguousarray(th)
            self._sum(mu, th)
        return dt, mu, th

    def update(self, psi, mu, step: int, time: float, out=None):
        if not self._device_reduce:
            info, out = super().update(psi, mu, step, time, out=out)
            self._sum(*out)
            return info, out
        if out is None:
            out = (np.empty(self.n_sites, np.complex128), np.empty(self.n_sites),
                   np.empty(self.n_edges), np.empty(self.n_edges))
        info, _ = super().update(psi, mu, step, time, out=(None, None, None, None))
        self._outputs_on_device(3, *out)
        return info, out


class LocalShardGroup:
    """All shards of one mesh inside this process.  ``devices[r]`` is the CUDA ordinal of
    shard r (default: all on device 0 — the exchange kernels then run through the same
    code path, on one GPU's memory)."""

    def __init__(self, mesh, world: int, devices: Optional[Sequence[int]] = None, **kw):
        devices = list(devices) if devices is not None else [0] * world
        self.world = world
        self.engines = [DeviceEngine(mesh, device=devices[r], world=world, rank=r, **kw)
                        for r in range(world)]
        for e in self.engines:
            e.comm_connect_local(self.engines)
        e0 = self.engines[0]
        self.n_sites, self.n_edges, self.n_probe = e0.n_sites, e0.n_edges, e0.n_probe
        self.n_boundary_edges = e0.n_boundary_edges

    def close(self):
        for e in self.engines:
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _all(self, fn):
        """Run fn(engine) for every shard concurrently (the shards wait for each other on
        the device, so their host calls must overlap)."""
        out: list = [None] * self.world
        err: list = [None] * self.world

        def work(r):
            try:
                out[r] = fn(self.engines[r])
            except BaseException as exc:  # noqa: BLE001  (re-raised below)
                err[r] = exc

        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        failed = [(r, e) for r, e in enumerate(err) if e is not None]
        if len(failed) > 1:    # usually one root cause and the others timing out on it
            raise type(failed[0][1])(
                "; ".join(f"shard {r}: {e}" for r, e in failed)) from failed[0][1]
        if failed:
            raise failed[0][1]
        return out

    # inputs are whole-mesh arrays, every shard takes its part
    def set_link_exponents(self, A):
        self._all(lambda e: e.set_link_exponents(A))

    def set_epsilon(self, eps):
        self._all(lambda e: e.set_epsilon(eps))

    def set_mu_boundary(self, mub):
        self._all(lambda e: e.set_mu_boundary(mub))

    def set_dA_dt(self, dA_dt):
        self._all(lambda e: e.set_dA_dt(dA_dt))

    def set_vector_potential_ramp(self, A0, t_knots, f_knots):
        self._all(lambda e: e.set_vector_potential_ramp(A0, t_knots, f_knots))

    def set_terminal_current_table(self, *a):
        self._all(lambda e: e.set_terminal_current_table(*a))

    def set_epsilon_table(self, *a):
        self._all(lambda e: e.set_epsilon_table(*a))

    def set_state(self, psi, mu):
        self._all(lambda e: e.set_state(psi, mu))

    def set_stepper(self, **kw):
        self._all(lambda e: e.set_stepper(**kw))

    def advance(self, max_steps: int, t_end: float, step: int, time: float) -> AdvanceInfo:
        infos = self._all(lambda e: e.advance(max_steps, t_end, step, time))
        a = infos[0]
        for b in infos[1:]:
            if (a.steps_done, a.step, a.time, a.dt, a.retries, a.mu_iterations) != (
                    b.steps_done, b.step, b.time, b.dt, b.retries, b.mu_iterations):
                raise RuntimeError(f"shards disagree on the step bookkeeping: {a} vs {b}")
        return a._replace(device_ms=max(i.device_ms for i in infos))

    def local_maps(self):
        return [e.local_maps() for e in self.engines]

    def update_local(self, psi_parts, mu_parts, step: int, time: float, outs):
        """Shard-local step seam on every shard at once: ``psi_parts[r]`` / ``mu_parts[r]`` are
        shard r's owned entries, ``outs[r]`` its four output arrays."""
        res = self._all(lambda e: e.update_local(psi_parts[e.rank], mu_parts[e.rank], step, time,
                                                 outs[e.rank]))
        return res[0][0], outs

    def get_state(self):
        parts = self._all(lambda e: e.get_state())
        return sum(p[0] for p in parts), sum(p[1] for p in parts)

    def get_currents(self):
        parts = self._all(lambda e: e.get_currents())
        return sum(p[0] for p in parts), sum(p[1] for p in parts)

    def get_running(self, steps: int):
        parts = [e.get_running(steps) for e in self.engines]
        return parts[0][0], sum(p[1] for p in parts), sum(p[2] for p in parts)

    def info(self) -> dict:
        d = self.engines[0].info()
        d["launches"] = sum(e.info()["launches"] for e in self.engines)
        return d

    def shard_info(self):
        return [e.shard_info() for e in self.engines]
