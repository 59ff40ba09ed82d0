"""Field sources for ``applied_vector_potential`` — the subset of the reference's
``tdgl.sources`` / ``tdgl.Parameter`` algebra the hot path meets (tdgl/sources/constant.py,
tdgl/sources/scaling.py, tdgl/parameter.py:355-373): a constant field, a linear ramp, and
their product.

The reference evaluates a time-dependent ``Parameter`` through a Python callback every step
(solver.py:626-642).  A product ``scalar(t) * field(r)`` of these sources is *separable*, and
``TDGLSolver`` hands it to the device (``tdgl_set_vector_potential_ramp``): the ramp is then
evaluated inside the device-side step loop and the run needs no per-step host work.  Any
other callable with ``time_dependent = True`` and signature ``f(x, y, z, *, t)`` takes the
generic path (host callback every step, like the reference).
"""

from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from .device import constant_field_vector_potential


class Field:
    """A spatial field ``f(x, y, z) -> [n, 3]``, optionally times a piecewise-linear scalar
    ``s(t)`` given by knots."""

    def __init__(self, spatial: Callable, knots: Optional[Tuple[Sequence[float],
                                                                 Sequence[float]]] = None):
        self.spatial = spatial
        self.knots = None
        if knots is not None:
            t, v = (np.asarray(k, float) for k in knots)
            if t.ndim != 1 or t.shape != v.shape or len(t) < 2 or np.any(np.diff(t) <= 0):
                raise ValueError("knots must be increasing times and as many values")
            self.knots = (t, v)

    @property
    def time_dependent(self) -> bool:
        return self.knots is not None

    #: (spatial callable, t_knots, f_knots) when the field is scalar(t) * field(r)
    @property
    def separable(self):
        return None if self.knots is None else (self.spatial, self.knots[0], self.knots[1])

    def scale(self, t: float) -> float:
        if self.knots is None:
            return 1.0
        return float(np.interp(t, self.knots[0], self.knots[1]))

    def __call__(self, x, y, z, *, t: Optional[float] = None):
        A = np.asarray(self.spatial(x, y, z))
        return A if self.knots is None else self.scale(0.0 if t is None else t) * A

    def __mul__(self, other):
        if isinstance(other, Ramp):
            return other * self
        if np.isscalar(other):
            s = float(other)
            k = None if self.knots is None else (self.knots[0], s * self.knots[1])
            return Field((lambda x, y, z, _f=self.spatial: s * np.asarray(_f(x, y, z)))
                         if k is None else self.spatial, k)
        return NotImplemented

    __rmul__ = __mul__


class Ramp:
    """Scalar ``s(t)``: ``initial`` before ``tmin``, linear to ``final`` at ``tmax``, ``final``
    after (reference ``linear_ramp``, sources/scaling.py:4-14)."""

    time_dependent = True

    def __init__(self, tmin: float, tmax: float, initial: float = 0.0, final: float = 1.0):
        if not tmax > tmin:
            raise ValueError("tmax must be greater than tmin")
        self.tmin, self.tmax, self.initial, self.final = tmin, tmax, initial, final

    def __call__(self, x=None, y=None, z=None, *, t: float):
        if t < self.tmin:
            return self.initial
        if t < self.tmax:
            return self.initial + (self.final - self.initial) * (t - self.tmin) / (
                self.tmax - self.tmin)
        return self.final

    def __mul__(self, other):
        if isinstance(other, Field) and other.knots is None:
            return Field(other.spatial, ([self.tmin, self.tmax], [self.initial, self.final]))
        return NotImplemented

    __rmul__ = __mul__


def ConstantField(value: float = 0, field_units: str = "mT", length_units: str = "um") -> Field:
    """Uniform out-of-plane field ``value``, symmetric gauge (reference ``ConstantField``,
    sources/constant.py:25-39, same signature).  The returned vector potential is in
    ``field_units * length_units`` like the reference's; its magnitude ``B x r / 2`` does not
    depend on the unit names (positions go to metres and back, constant.py:18-22), so they
    are kept only as attributes."""
    f = Field(lambda x, y, z, _b=float(value): constant_field_vector_potential(x, y, z, Bz=_b))
    f.field_units, f.length_units = str(field_units), str(length_units)
    return f


def LinearRamp(*, tmin: float, tmax: float, initial: float = 0.0, final: float = 1.0) -> Ramp:
    """reference ``LinearRamp`` (sources/scaling.py:17-40)."""
    return Ramp(tmin, tmax, initial, final)


def _table_lookup(t_knots: np.ndarray, values: np.ndarray, t: float) -> float:
    """Piecewise-linear interpolation, constant outside the knots — the arithmetic of the
    device's ``table_lookup`` (csrc/kernels.cuh), operation for operation."""
    n = len(t_knots)
    if t >= t_knots[n - 1]:
        return float(values[n - 1])
    if t <= t_knots[0]:
        return float(values[0])
    k = 0
    while k + 2 < n and t >= t_knots[k + 1]:
        k += 1
    w = (np.float64(t) - t_knots[k]) / (t_knots[k + 1] - t_knots[k])
    return float(values[k] + w * (values[k + 1] - values[k]))


class PiecewiseLinearCurrents:
    """Time-dependent terminal currents ``t -> {name: I}`` given by knots.  Usable wherever the
    reference takes a callable ``terminal_currents`` (solver.py:234-249); because it is a table,
    ``TDGLSolver`` hands it to the device (``tdgl_set_terminal_current_table``) and the run needs
    no per-step host callback — a generic callable is evaluated from Python every step, like in
    the reference.

    ``times``: increasing knots; ``currents``: ``{terminal name: values at the knots}``."""

    def __init__(self, times: Sequence[float], currents):
        t = np.asarray(times, dtype=np.float64)
        if t.ndim != 1 or len(t) < 2 or np.any(np.diff(t) <= 0):
            raise ValueError("times must be at least two increasing knots")
        self.times = t
        self.currents = {str(k): np.asarray(v, dtype=np.float64) for k, v in currents.items()}
        for k, v in self.currents.items():
            if v.shape != t.shape:
                raise ValueError(f"currents[{k!r}] must have one value per knot")

    @property
    def knots(self):
        return self.times, self.currents

    def scaled(self, factor: float) -> "PiecewiseLinearCurrents":
        return PiecewiseLinearCurrents(self.times, {k: factor * v for k, v in self.currents.items()})

    def __call__(self, t: float):
        return {k: _table_lookup(self.times, v, t) for k, v in self.currents.items()}


class SeparableEpsilon:
    """``epsilon(r, t) = e0(r) + g(t) * e1(r)`` with piecewise-linear ``g``: a time-dependent
    disorder parameter (the reference's ``disorder_epsilon(r, *, t)``, solver.py:364-381) in a
    form the device evaluates inside the psi step (``tdgl_set_epsilon_table``).  ``e0`` / ``e1``:
    callables ``r -> value`` evaluated at the sites (vectorised: ``r`` is [N, 2]) or arrays
    [N]; ``times`` / ``g``: the knots of g."""

    def __init__(self, e0, e1, times: Sequence[float], g: Sequence[float]):
        self.e0, self.e1 = e0, e1
        self.times = np.asarray(times, dtype=np.float64)
        self.g = np.asarray(g, dtype=np.float64)
        if self.times.ndim != 1 or len(self.times) < 2 or np.any(np.diff(self.times) <= 0):
            raise ValueError("times must be at least two increasing knots")
        if self.g.shape != self.times.shape:
            raise ValueError("g must have one value per knot")

    def arrays(self, sites: np.ndarray):
        def ev(f):
            return (np.asarray(f(sites), dtype=np.float64) if callable(f)
                    else np.asarray(f, dtype=np.float64))
        return ev(self.e0), ev(self.e1)

    def scale(self, t: float) -> float:
        return _table_lookup(self.times, self.g, t)

    def __call__(self, r, *, t: float, vectorized: bool = True):
        e0, e1 = self.arrays(np.atleast_2d(r))
        return e0 + np.float64(self.scale(t)) * e1
