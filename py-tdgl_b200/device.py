"""``Device`` / ``Layer`` / ``Polygon``: the slice of the reference's device layer that
``TDGLSolver`` touches: ``.mesh``, ``.layer.{u,gamma,z0,coherence_length}``,
``.length_units``, ``.K0``, ``.Bc2``, ``.probe_points``, ``.probe_point_indices``,
``.terminal_info()``.  pint / shapely / meshpy / matplotlib are not available here, so
units reduce to a table of SI prefixes and meshing to a Delaunay triangulation of a
jittered lattice (reference: tdgl/device/device.py, layer.py, polygon.py, meshing.py).
Geometry and meshing are setup-time code outside the hot path.
"""

from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np

from .mesh import Mesh, morton_order
from .synthetic import TerminalInfo

PHI_0 = 2.067833848e-15      # Wb
MU_0 = 1.25663706212e-6      # N / A^2

_PREFIX = {"": 1.0, "k": 1e3, "c": 1e-2, "m": 1e-3, "u": 1e-6, "µ": 1e-6, "n": 1e-9,
           "p": 1e-12}


def _unit(unit: str, base: Sequence[str]) -> float:
    """Scale factor to SI of a prefixed unit such as 'um', 'mT', 'uA'."""
    unit = unit.strip()
    aliases = {"meter": "m", "tesla": "T", "ampere": "A", "amp": "A", "micron": "um",
               "gauss": "G", "angstrom": "angstrom"}
    unit = aliases.get(unit, unit)
    if unit == "angstrom":
        return 1e-10
    for b in base:
        if unit.endswith(b):
            pre = unit[: len(unit) - len(b)]
            if pre in _PREFIX:
                scale = _PREFIX[pre]
                if b == "G":
                    scale *= 1e-4
                return scale
    raise ValueError(f"Unknown unit {unit!r}")


def length_scale(unit: str) -> float:
    return _unit(unit, ["m"])


def field_scale(unit: str) -> float:
    return _unit(unit, ["T", "G"])


def current_scale(unit: str) -> float:
    return _unit(unit, ["A"])


class Layer:
    """reference tdgl/device/layer.py:6-41"""

    def __init__(self, *, london_lambda: float, coherence_length: float, thickness: float,
                 conductivity: Union[float, None] = None, u: float = 5.79,
                 gamma: float = 10.0, z0: float = 0):
        self.london_lambda = london_lambda
        self.coherence_length = coherence_length
        self.thickness = thickness
        self.conductivity = conductivity
        self.u = u
        self.gamma = gamma
        self.z0 = z0

    @property
    def Lambda(self) -> float:
        return self.london_lambda**2 / self.thickness

    def copy(self) -> "Layer":
        return Layer(london_lambda=self.london_lambda, coherence_length=self.coherence_length,
                     thickness=self.thickness, conductivity=self.conductivity, u=self.u,
                     gamma=self.gamma, z0=self.z0)

    def __eq__(self, other):
        return isinstance(other, Layer) and self.__dict__ == other.__dict__


def box(width: float, height: Optional[float] = None, points: int = 101,
        center: Tuple[float, float] = (0, 0)) -> np.ndarray:
    """Rectangle outline (reference tdgl/geometry.py:85-134, without rotation)."""
    width = abs(width)
    height = width if height is None else abs(height)
    x0, y0 = center
    perimeter = 2 * (width + height)
    nx = max(int(round(points * width / perimeter)), 2)
    ny = max(int(round(points * height / perimeter)), 2)
    xs = np.linspace(-width / 2, width / 2, nx + 1)
    ys = np.linspace(-height / 2, height / 2, ny + 1)
    pts = np.concatenate([
        np.stack([xs[:-1], np.full(nx, -height / 2)], 1),
        np.stack([np.full(ny, width / 2), ys[:-1]], 1),
        np.stack([xs[:0:-1], np.full(nx, height / 2)], 1),
        np.stack([np.full(ny, -width / 2), ys[:0:-1]], 1)])
    return pts + np.array([x0, y0])


def circle(radius: float, points: int = 100, center: Tuple[float, float] = (0, 0)):
    """reference tdgl/geometry.py:62-82"""
    th = 2 * np.pi * np.arange(points) / points
    return np.stack([center[0] + radius * np.cos(th), center[1] + radius * np.sin(th)], 1)


class Polygon:
    """A named closed polygon (reference tdgl/device/polygon.py, containment only)."""

    def __init__(self, name: Optional[str] = None, *, points):
        self.name = name
        pts = np.asarray(points, dtype=float)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError(f"Expected shape (n, 2), but got {pts.shape}.")
        if np.allclose(pts[0], pts[-1]):
            pts = pts[:-1]
        self.points = pts

    def contains_points(self, points, index: bool = False, radius: float = 0):
        """Even-odd rule; points within ``1e-9 * extent`` (+ ``radius``) of the outline count
        as inside, which is what the reference's terminal/boundary tests rely on."""
        p = np.atleast_2d(np.asarray(points, dtype=float))
        v = self.points
        x, y = p[:, 0], p[:, 1]
        inside = np.zeros(len(p), dtype=bool)
        near = np.zeros(len(p), dtype=bool)
        tol = 1e-9 * max(np.ptp(v[:, 0]), np.ptp(v[:, 1]), 1e-300) + radius
        n = len(v)
        for k in range(n):
            x0, y0 = v[k]
            x1, y1 = v[(k + 1) % n]
            cond = (y0 > y) != (y1 > y)
            with np.errstate(divide="ignore", invalid="ignore"):
                xint = (x1 - x0) * (y - y0) / (y1 - y0) + x0
            inside ^= cond & (x < xint)
            # distance to the segment
            dx, dy = x1 - x0, y1 - y0
            L2 = dx * dx + dy * dy
            t = np.clip(((x - x0) * dx + (y - y0) * dy) / max(L2, 1e-300), 0, 1)
            near |= np.hypot(x - (x0 + t * dx), y - (y0 + t * dy)) <= tol
        out = inside | near
        return np.where(out)[0] if index else out

    @property
    def extents(self):
        return np.ptp(self.points[:, 0]), np.ptp(self.points[:, 1])

    def copy(self) -> "Polygon":
        return Polygon(self.name, points=self.points.copy())

    def __eq__(self, other):
        return (isinstance(other, Polygon) and self.name == other.name
                and self.points.shape == other.points.shape
                and np.allclose(self.points, other.points))


class UnitScales(NamedTuple):
    A_scale: float
    J_scale: float


def unit_scales(device: "Device", field_units: str, current_units: str) -> UnitScales:
    """``A_scale`` (solver.py:176-180) and ``J_scale`` (solver.py:251-253) without pint."""
    xi = device.layer.coherence_length
    L = length_scale(device.length_units)
    Bc2 = device.Bc2
    A_scale = field_scale(field_units) / (Bc2 * xi)
    J_scale = 4 * (current_scale(current_units) / L) / device.K0
    return UnitScales(A_scale, J_scale)


def constant_field_vector_potential(x, y, z, *, Bz: float):
    """Uniform field ``Bz`` (in field units) in the symmetric gauge about the bounding box
    of the evaluation points; returns A in field_units * length_units
    (reference sources/constant.py:7-22, em.py:437-472)."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    xs = x - (x.min() + np.ptp(x) / 2)
    ys = y - (y.min() + np.ptp(y) / 2)
    return np.stack([-Bz * ys / 2, Bz * xs / 2, np.zeros_like(xs)], axis=1)


class Device:
    """reference tdgl/device/device.py:49-256 (solver-facing subset)."""

    def __init__(self, name: str, *, layer: Layer, film: Polygon,
                 holes: Optional[List[Polygon]] = None,
                 terminals: Optional[List[Polygon]] = None,
                 probe_points: Optional[Sequence[Tuple[float, float]]] = None,
                 length_units: str = "um"):
        self.name = name
        self.layer = layer
        self.film = film
        self.holes = list(holes) if holes is not None else []
        self.terminals = tuple(terminals) if terminals is not None else tuple()
        names = set()
        for term in self.terminals:
            term.mesh = False
            if term.name is None:
                raise ValueError("All current terminals must have a unique name.")
            if term.name in names:
                raise ValueError("All current terminals must have a unique name.")
            names.add(term.name)
        if probe_points is not None:
            probe_points = np.asarray(probe_points).squeeze()
            if probe_points.ndim == 1:
                probe_points = probe_points[None, :]
            if probe_points.ndim != 2 or probe_points.shape[1] != 2:
                raise ValueError(
                    f"Probe points must have shape (n, 2), got {probe_points.shape}.")
        self.probe_points = probe_points
        self._length_units = length_units
        self.mesh: Optional[Mesh] = None

    # -- units (device.py:120-168) ----------------------------------------------------------
    @property
    def length_units(self) -> str:
        return self._length_units

    @property
    def coherence_length(self) -> float:
        return self.layer.coherence_length

    @property
    def Bc2(self) -> float:
        """Upper critical field in tesla, Phi_0 / (2 pi xi^2)."""
        xi_m = self.layer.coherence_length * length_scale(self.length_units)
        return PHI_0 / (2 * np.pi * xi_m**2)

    @property
    def A0(self) -> float:
        return self.Bc2 * self.layer.coherence_length * length_scale(self.length_units)

    @property
    def K0(self) -> float:
        """Sheet current density scale in A/m, 4 xi Bc2 / (mu_0 Lambda)."""
        L = length_scale(self.length_units)
        return 4 * self.layer.coherence_length * L * self.Bc2 / (MU_0 * self.layer.Lambda * L)

    # -- mesh-derived (device.py:221-306) ----------------------------------------------------
    @property
    def points(self):
        return None if self.mesh is None else self.mesh.sites * self.layer.coherence_length

    @property
    def edge_lengths(self):
        if self.mesh is None:
            return None
        return self.mesh.edge_mesh.edge_lengths * self.layer.coherence_length

    @property
    def probe_point_indices(self):
        if self.mesh is None or self.probe_points is None:
            return None
        xi = self.layer.coherence_length
        return [self.mesh.closest_site(xy) for xy in self.probe_points / xi]

    def terminal_info(self) -> Tuple[TerminalInfo, ...]:
        xi = self.layer.coherence_length
        mesh = self.mesh
        sites = self.points
        edge_positions = xi * mesh.edge_mesh.centers
        ix_boundary = mesh.edge_mesh.boundary_edge_indices
        edge_lengths = self.edge_lengths[ix_boundary]
        boundary_edge_positions = edge_positions[ix_boundary]
        info = []
        for terminal in self.terminals:
            sites_index = np.intersect1d(terminal.contains_points(sites, index=True),
                                         mesh.boundary_indices)
            edges_index = np.intersect1d(
                terminal.contains_points(edge_positions, index=True), ix_boundary)
            boundary_edges_index = terminal.contains_points(boundary_edge_positions, index=True)
            length = float(edge_lengths[boundary_edges_index].sum())
            info.append(TerminalInfo(terminal.name, sites_index, edges_index,
                                     boundary_edges_index, length))
        return tuple(sorted(info, key=lambda t: t.length))

    def contains_points(self, points):
        inside = self.film.contains_points(points)
        for hole in self.holes:
            inside &= ~hole.contains_points(points, radius=-1e-9)
        return inside

    def make_mesh(self, max_edge_length: Optional[float] = None,
                  min_points: Optional[int] = None, smooth: int = 0, jitter: float = 0.15,
                  seed: int = 0, reorder: bool = True, **_ignored) -> None:
        """Generates the triangular mesh with the reference's contract (``Device.make_mesh``,
        tdgl/device/device.py:520-566; ``generate_mesh``, tdgl/device/meshing.py:15-123):

        * ``max_edge_length``: no edge of the result is longer (default 1.0 x coherence length;
          <= 0: the density follows the point density of the film / hole polygons only);
        * ``min_points``: the result has at least this many vertices;
        * ``smooth``: that many Laplacian smoothing sweeps of the interior vertices
          (``Mesh.smooth``) afterwards;
        * like the reference the generator refines until both bounds hold.

        The generator itself is not Triangle (meshpy is a setup-time dependency this package
        does not take): outline points of film and holes resampled at the pitch + a jittered
        hexagonal interior lattice, Delaunay-triangulated (Qhull), elements outside the film
        or inside holes removed.  ``meshpy_kwargs`` such as ``min_angle`` are accepted and
        ignored (a jittered hexagonal lattice has angles of 40-80 degrees)."""
        xi = self.layer.coherence_length
        w, hgt = self.film.extents
        if max_edge_length is None:
            max_edge_length = 1.0 * xi
        max_edge_length = float(max_edge_length)
        min_points = int(min_points) if min_points else 0
        if max_edge_length <= 0:
            # density of the polygons' own points (reference: "determined solely by the density
            # of points in the Device's film and holes"), still subject to min_points
            seg = np.concatenate([np.linalg.norm(np.diff(np.vstack([p.points, p.points[:1]]),
                                                         axis=0), axis=1)
                                  for p in [self.film] + list(self.holes)])
            h = float(np.median(seg[seg > 0]))
            max_edge_length = np.inf
        else:
            h = 0.8 * max_edge_length
        if min_points:
            # a hexagonal lattice of pitch h has 2 / (sqrt(3) h^2) points per unit area
            h = min(h, np.sqrt(self._area() * 2 / np.sqrt(3) / min_points))
        for attempt in range(40):
            pts, tri = self._triangulate(h, jitter, seed)
            longest = _max_edge_length(pts, tri)
            if len(pts) >= min_points and longest <= max_edge_length:
                break
            # (the reference shrinks Triangle's max_volume by min(0.98, sqrt(target / longest))
            # per pass, meshing.py:117-120; the pitch is a length, so the same factor applies
            # to it directly and converges in a pass or two)
            shrink = 0.98
            if np.isfinite(max_edge_length) and longest > max_edge_length:
                shrink = min(shrink, 0.98 * max_edge_length / longest)
            if len(pts) < min_points:
                shrink = min(shrink, 0.98 * np.sqrt(len(pts) / min_points))
            h *= shrink
        else:
            raise RuntimeError("make_mesh: could not satisfy min_points / max_edge_length")
        if smooth:
            m = Mesh.from_triangulation(pts, tri, create_submesh=False).smooth(
                int(smooth), create_submesh=False)
            pts, tri = m.sites, m.elements
        if reorder:
            perm = morton_order(pts)
            inv = np.empty_like(perm)
            inv[perm] = np.arange(len(perm))
            pts, tri = pts[perm], inv[tri]
        self.mesh = Mesh.from_triangulation(pts / xi, tri)

    def _area(self) -> float:
        def poly_area(v):
            x, y = v[:, 0], v[:, 1]
            return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        return poly_area(self.film.points) - sum(poly_area(hh.points) for hh in self.holes)

    def _triangulate(self, h: float, jitter: float, seed: int):
        """Points (length units) and counter-clockwise triangles at lattice pitch ``h``."""
        from scipy.spatial import Delaunay

        from .mesh import _hex_lattice

        def resample(poly):
            v = poly.points
            out = []
            for k in range(len(v)):
                a, b = v[k], v[(k + 1) % len(v)]
                m = max(int(np.ceil(np.linalg.norm(b - a) / h)), 1)
                t = np.arange(m)[:, None] / m
                out.append(a + t * (b - a))
            return np.concatenate(out)

        outlines = [resample(self.film)] + [resample(p) for p in self.holes]
        boundary = np.concatenate(outlines)
        xmin, ymin = self.film.points.min(axis=0)
        xmax, ymax = self.film.points.max(axis=0)
        rng = np.random.default_rng(seed)
        pts = _hex_lattice(xmin, xmax, ymin, ymax, h)
        pts = pts + rng.uniform(-jitter * h, jitter * h, size=pts.shape)
        keep = self.contains_points(pts)
        # keep clear of every outline
        for outline in outlines:
            d = _min_distance(pts, outline)
            keep &= d > 0.6 * h
        allpts = np.concatenate([boundary, pts[keep]])
        tri = Delaunay(allpts).simplices.astype(np.int64)
        p = allpts[tri]
        area2 = ((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                 - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0]))
        flip = area2 < 0
        tri[flip] = tri[flip][:, [0, 2, 1]]
        cen = p.mean(axis=1)
        good = (np.abs(area2) > 1e-10 * np.abs(area2).max()) & self.film.contains_points(cen)
        for hole in self.holes:
            good &= ~hole.contains_points(cen, radius=-1e-9 * h)
        tri = tri[good]
        used = np.zeros(len(allpts), dtype=bool)
        used[tri.ravel()] = True
        remap = np.cumsum(used) - 1
        return allpts[used], remap[tri]

    def copy(self) -> "Device":
        d = Device(self.name, layer=self.layer.copy(), film=self.film.copy(),
                   holes=[h.copy() for h in self.holes],
                   terminals=[t.copy() for t in self.terminals],
                   probe_points=None if self.probe_points is None else self.probe_points.copy(),
                   length_units=self.length_units)
        d.mesh = self.mesh
        return d

    def __eq__(self, other):
        if other is self:
            return True
        if not isinstance(other, Device):
            return False
        return (self.name == other.name and self.layer == other.layer
                and self.film == other.film and self.holes == other.holes
                and list(self.terminals) == list(other.terminals)
                and self.length_units == other.length_units)


def _max_edge_length(points: np.ndarray, elements: np.ndarray) -> float:
    """reference ``get_max_edge_length`` (finite_volume/util.py:45-56)."""
    e = np.concatenate([elements[:, c] for c in ((0, 1), (1, 2), (2, 0))])
    return float(np.linalg.norm(points[e[:, 1]] - points[e[:, 0]], axis=1).max())


def _min_distance(pts: np.ndarray, outline: np.ndarray) -> np.ndarray:
    """Distance from each point to the nearest outline vertex (outlines are resampled at the
    mesh pitch, so vertex distance is a good proxy for distance to the curve)."""
    from scipy.spatial import cKDTree

    d, _ = cKDTree(outline).query(pts)
    return d
