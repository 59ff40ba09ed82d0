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
