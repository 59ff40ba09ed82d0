"""Finite-volume mesh arrays: the input contract of the hot path.

Mirrors the *data* of the reference's ``tdgl.finite_volume.Mesh`` / ``EdgeMesh``
(reference ``tdgl/finite_volume/mesh.py:24-69``, ``edge_mesh.py:9-42``): same attribute
names, shapes, dtypes, units (lengths in units of xi) and ordering (edge rows sorted and
lexicographically unique, ``util.py:15-28``).  The construction here is vectorised
NumPy (the reference loops over edges and sites in Python, ``util.py:59-97,169-255``;
108 s at 1M sites); it is setup-time code, not part of the timed path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np


class EdgeMesh:
    """Edges of the triangulation (reference ``edge_mesh.py:9-42``)."""

    def __init__(self, centers, edges, boundary_edge_indices, directions,
                 edge_lengths, dual_edge_lengths):
        self.centers = np.asarray(centers, dtype=np.float64)
        self.edges = np.asarray(edges, dtype=np.int64)
        self.boundary_edge_indices = np.asarray(boundary_edge_indices, dtype=np.int64)
        self.directions = np.asarray(directions, dtype=np.float64)
        self.normalized_directions = (
            self.directions / np.linalg.norm(self.directions, axis=1)[:, np.newaxis]
        )
        self.edge_lengths = np.asarray(edge_lengths, dtype=np.float64)
        self.dual_edge_lengths = np.asarray(dual_edge_lengths, dtype=np.float64)

    @property
    def x(self):
        return self.centers[:, 0]

    @property
    def y(self):
        return self.centers[:, 1]


def get_edges(elements: np.ndarray, num_sites: Optional[int] = None):
    """Unique sorted edges and the is-boundary mask (same result as reference
    ``util.py:15-28``; done on a packed 64-bit key instead of a row-wise unique)."""
    elements = np.asarray(elements, dtype=np.int64)
    if num_sites is None:
        num_sites = int(elements.max()) + 1
    pairs = np.concatenate([elements[:, [0, 1]], elements[:, [1, 2]], elements[:, [2, 0]]])
    lo = pairs.min(axis=1)
    hi = pairs.max(axis=1)
    key = lo * np.int64(num_sites) + hi
    ukey, counts = np.unique(key, return_counts=True)
    edges = np.stack([ukey // num_sites, ukey % num_sites], axis=1)
    return edges, counts == 1, key, ukey


def circumcenters(sites: np.ndarray, elements: np.ndarray) -> np.ndarray:
    """Voronoi vertices = triangle circumcentres (reference ``util.py:100-124``)."""
    A = sites[elements[:, 0]]
    B = sites[elements[:, 1]] - A
    C = sites[elements[:, 2]] - A
    D = 2 * B[:, 0] * C[:, 1] - 2 * B[:, 1] * C[:, 0]
    b2 = (B**2).sum(axis=1)
    c2 = (C**2).sum(axis=1)
    Ux = (C[:, 1] * b2 - B[:, 1] * c2) / D
    Uy = (B[:, 0] * c2 - C[:, 0] * b2) / D
    return np.array([Ux, Uy]).T + A


class Mesh:
    """Triangular mesh + Voronoi dual (reference ``mesh.py:24-69``)."""

    def __init__(self, sites, elements, boundary_indices, areas=None, dual_sites=None,
                 edge_mesh: Optional[EdgeMesh] = None, voronoi_polygons=None):
        self.sites = np.asarray(sites, dtype=np.float64)
        self.elements = np.asarray(elements, dtype=np.int64)
        self.boundary_indices = np.asarray(boundary_indices, dtype=np.int64)
        self.areas = None if areas is None else np.asarray(areas, dtype=np.float64)
        self.dual_sites = None if dual_sites is None else np.asarray(dual_sites)
        self.edge_mesh = edge_mesh
        self.voronoi_polygons = voronoi_polygons

    @property
    def x(self):
        return self.sites[:, 0]

    @property
    def y(self):
        return self.sites[:, 1]

    def closest_site(self, xy) -> int:
        """reference ``mesh.py:92-101``"""
        return int(np.argmin(np.linalg.norm(self.sites - np.atleast_2d(xy), axis=1)))

    def smooth(self, iterations: int, create_submesh: bool = True) -> "Mesh":
        """Laplacian smoothing: every interior vertex moves to the arithmetic mean of its
        neighbours, boundary vertices stay (reference ``Mesh.smooth``, ``mesh.py:245-283``;
        here the neighbour sums are two ``bincount`` passes per coordinate over the unique
        edges and the dual mesh is built once, after the last sweep)."""
        sites = np.array(self.sites, dtype=np.float64)
        if iterations <= 0:
            return Mesh.from_triangulation(sites, self.elements, create_submesh=create_submesh)
        n = len(sites)
        edges = get_edges(self.elements, n)[0]
        e0, e1 = edges[:, 0], edges[:, 1]
        num_neighbors = np.bincount(edges.ravel(), minlength=n).astype(np.float64)
        boundary = self.boundary_indices
        for _ in range(int(iterations)):
            new = np.empty_like(sites)
            for d in range(2):
                # (same order of accumulation as the reference: edge[:, 0] targets first)
                acc = np.bincount(e0, sites[e1, d], minlength=n)
                acc += np.bincount(e1, sites[e0, d], minlength=n)
                new[:, d] = acc / num_neighbors
            new[boundary] = sites[boundary]
            sites = new
        return Mesh.from_triangulation(sites, self.elements, create_submesh=create_submesh)

    @staticmethod
    def from_triangulation(sites, elements, create_submesh: bool = True) -> "Mesh":
        """Equivalent of reference ``Mesh.from_triangulation`` (``mesh.py:104-151``) +
        ``EdgeMesh.from_mesh`` (``edge_mesh.py:54-92``): the edge / dual-mesh arrays come from
        the library's multi-threaded host builder (``tdgl_host_mesh_dual``,
        ``csrc/host_mesh.h``), the Voronoi areas of the boundary sites from
        :func:`_voronoi_areas`.  :meth:`_from_triangulation_numpy` is the same construction
        in vectorised NumPy and gives the same arrays bit for bit (tests)."""
        sites, elements = _check_triangulation(sites, elements)
        if not create_submesh or len(elements) == 0:
            return Mesh._from_triangulation_numpy(sites, elements, create_submesh)
        from . import _lib

        lib = _lib.load()
        n, T = len(sites), len(elements)
        sites = np.ascontiguousarray(sites)
        elements = np.ascontiguousarray(elements)
        cap = 3 * T
        edges = np.empty((cap, 2), dtype=np.int64)
        is_boundary = np.empty(cap, dtype=np.uint8)
        dual = np.empty((T, 2))
        centers = np.empty((cap, 2))
        directions = np.empty((cap, 2))
        lengths = np.empty(cap)
        dual_len = np.empty(cap)
        areas = np.empty(n)
        n_edges = C.c_int64(0)
        rc = lib.tdgl_host_mesh_dual(n, T, _lib.ptr(sites), _lib.ptr(elements), C.byref(n_edges),
                                     _lib.ptr(edges), _lib.ptr(is_boundary), _lib.ptr(dual),
                                     _lib.ptr(centers), _lib.ptr(directions), _lib.ptr(lengths),
                                     _lib.ptr(dual_len), _lib.ptr(areas))
        if rc != 0:
            raise ValueError(lib.tdgl_last_error(None).decode())
        E = n_edges.value
        edges, centers, directions = edges[:E].copy(), centers[:E].copy(), directions[:E].copy()
        lengths, dual_len = lengths[:E].copy(), dual_len[:E].copy()
        is_boundary = is_boundary[:E].astype(bool)
        boundary_edge_indices = np.where(is_boundary)[0]
        boundary_indices = np.unique(edges[is_boundary].ravel())
        edge_mesh = EdgeMesh(centers, edges, boundary_edge_indices, directions, lengths, dual_len)
        areas = _voronoi_areas(sites, dual, elements, edges, lengths, dual_len,
                               boundary_indices, boundary_edge_indices, interior=areas)
        return Mesh(sites, elements, boundary_indices, areas=areas, dual_sites=dual,
                    edge_mesh=edge_mesh)

    @staticmethod
    def _from_triangulation_numpy(sites, elements, create_submesh: bool = True) -> "Mesh":
        """The construction of :meth:`from_triangulation` in vectorised NumPy (packed-key
        unique edges, searchsorted adjacency, bincount areas)."""
        sites, elements = _check_triangulation(sites, elements)
        n = len(sites)
        edges, is_boundary, tri_keys, ukey = get_edges(elements, n)
        boundary_edge_indices = np.where(is_boundary)[0]
        boundary_indices = np.unique(edges[is_boundary].ravel())
        if not create_submesh:
            return Mesh(sites, elements, boundary_indices)
        dual = circumcenters(sites, elements)
        coords = sites[edges]
        centers = coords.mean(axis=1)
        directions = coords[:, 1] - coords[:, 0]
        lengths = np.linalg.norm(directions, axis=1)
        # triangles adjacent to each edge (one for boundary edges, two otherwise)
        T = len(elements)
        tri_of_key = np.tile(np.arange(T, dtype=np.int64), 3)
        eidx = np.searchsorted(ukey, tri_keys)
        order = np.argsort(eidx, kind="stable")
        eidx_s = eidx[order]
        tri_s = tri_of_key[order]
        first = np.searchsorted(eidx_s, np.arange(len(edges)))
        t0 = tri_s[first]
        t1 = tri_s[np.minimum(first + 1, len(tri_s) - 1)]
        dual_len = np.where(
            is_boundary,
            np.linalg.norm(dual[t0] - centers, axis=1),
            np.linalg.norm(dual[t0] - dual[t1], axis=1),
        )
        edge_mesh = EdgeMesh(centers, edges, boundary_edge_indices, directions, lengths,
                             dual_len)
        areas = _voronoi_areas(sites, dual, elements, edges, lengths, dual_len,
                               boundary_indices, boundary_edge_indices)
        return Mesh(sites, elements, boundary_indices, areas=areas, dual_sites=dual,
                    edge_mesh=edge_mesh)


def _check_triangulation(sites, elements):
    sites = np.asarray(sites, dtype=np.float64).squeeze()
    elements = np.asarray(elements, dtype=np.int64).squeeze()
    if sites.ndim != 2 or sites.shape[1] != 2:
        raise ValueError(
            f"The site coordinates must have shape (n, 2), got {sites.shape!r}")
    if elements.ndim != 2 or elements.shape[1] != 3:
        raise ValueError(f"The elements must have shape (m, 3), got {elements.shape!r}.")
    return sites, elements


def _voronoi_areas(sites, dual, elements, edges, lengths, dual_len, boundary_indices,
                   boundary_edge_indices, interior=None) -> np.ndarray:
    """Voronoi cell areas.

    Interior sites: the cell is the convex polygon of the surrounding circumcentres,
    whose area is the sum over incident edges of the triangle (site, dual edge) =
    ``edge_length * dual_edge_length / 4`` (every edge at an interior site is interior).
    Boundary sites follow the reference's convention (``util.py:205-254``): convex hull
    of circumcentres + the two adjacent boundary-edge midpoints + the site, minus the
    (midpoint, midpoint, site) triangle when the site is not a hull vertex.  There are
    only O(sqrt(N)) of them, so that loop stays in Python.
    """
    from scipy.spatial import ConvexHull, QhullError

    n = len(sites)
    if interior is not None:       # the same sums, already formed by the native builder
        areas = interior
    else:
        quarter = 0.25 * lengths * dual_len
        areas = np.bincount(edges[:, 0], quarter, n) + np.bincount(edges[:, 1], quarter, n)
    if len(boundary_indices) == 0:
        return areas
    # incident triangles of boundary sites
    is_b = np.zeros(n, dtype=bool)
    is_b[boundary_indices] = True
    flat = elements.ravel()
    sel = np.nonzero(is_b[flat])[0]
    site_of = flat[sel]
    tri_of = sel // 3
    order = np.argsort(site_of, kind="stable")
    site_of = site_of[order]
    tri_of = tri_of[order]
    starts = np.searchsorted(site_of, boundary_indices)
    ends = np.searchsorted(site_of, boundary_indices, side="right")
    bedges = edges[boundary_edge_indices]
    bmid = sites[bedges].mean(axis=1)
    # the two boundary edges at each boundary site
    bflat = bedges.ravel()
    border = np.argsort(bflat, kind="stable")
    bsite = bflat[border]
    bedge_of = border // 2
    bstarts = np.searchsorted(bsite, boundary_indices)
    bends = np.searchsorted(bsite, boundary_indices, side="right")

    def hull_area(pts):
        try:
            hull = ConvexHull(pts)
        except QhullError:
            return 0.0, True
        return hull.volume, len(hull.vertices) == len(pts)

    for k, s in enumerate(boundary_indices):
        mids = bmid[bedge_of[bstarts[k]:bends[k]]]
        pts = np.concatenate([dual[tri_of[starts[k]:ends[k]]], mids, sites[s][None, :]])
        a, convex = hull_area(pts)
        if not convex:
            tri, _ = hull_area(np.concatenate([mids, sites[s][None, :]]))
            a -= tri
        areas[s] = a
    return areas


# ----------------------------------------------------------------------------------------
# Synthetic meshes: jittered hexagonal lattice + exact boundary points,
# Delaunay-triangulated.  meshpy/Triangle (reference device/meshing.py) is not available.
# ----------------------------------------------------------------------------------------

def _hex_lattice(xmin, xmax, ymin, ymax, h):
    dy = h * np.sqrt(3) / 2
    ny = int(np.floor((ymax - ymin) / dy)) + 1
    nx = int(np.floor((xmax - xmin) / h)) + 2
    y0 = ymin + 0.5 * ((ymax - ymin) - (ny - 1) * dy)
    ys = y0 + dy * np.arange(ny)
    xs = xmin + h * np.arange(-1, nx)
    X, Y = np.meshgrid(xs, ys)
    X = X + (np.arange(ny) % 2)[:, None] * (h / 2)
    return np.stack([X.ravel(), Y.ravel()], axis=1)


def make_film_points(width: float, height: float, h: float,
                     holes: Sequence[Tuple[float, float, float]] = (),
                     jitter: float = 0.15, seed: int = 0):
    """Points of a ``width x height`` rectangular film centred at the origin with
    circular ``holes`` [(cx, cy, r), ...]: exact, evenly spaced boundary points and a
    jittered hexagonal interior lattice of pitch ``h``."""
    rng = np.random.default_rng(seed)
    x0, x1 = -width / 2, width / 2
    y0, y1 = -height / 2, height / 2
    nx = max(int(round(width / h)), 2)
    ny = max(int(round(height / h)), 2)
    bx = np.linspace(x0, x1, nx + 1)
    by = np.linspace(y0, y1, ny + 1)
    boundary = [
        np.stack([bx[:-1], np.full(nx, y0)], 1),
        np.stack([np.full(ny, x1), by[:-1]], 1),
        np.stack([bx[:0:-1], np.full(nx, y1)], 1),
        np.stack([np.full(ny, x0), by[:0:-1]], 1),
    ]
    for (cx, cy, r) in holes:
        m = max(int(round(2 * np.pi * r / h)), 8)
        th = 2 * np.pi * np.arange(m) / m
        boundary.append(np.stack([cx + r * np.cos(th), cy + r * np.sin(th)], 1))
    pts = _hex_lattice(x0, x1, y0, y1, h)
    pts = pts + rng.uniform(-jitter * h, jitter * h, size=pts.shape)
    margin = 0.6 * h
    keep = ((pts[:, 0] > x0 + margin) & (pts[:, 0] < x1 - margin)
            & (pts[:, 1] > y0 + margin) & (pts[:, 1] < y1 - margin))
    for (cx, cy, r) in holes:
        keep &= np.hypot(pts[:, 0] - cx, pts[:, 1] - cy) > r + margin
    return np.concatenate(boundary + [pts[keep]])


_STRIP_WORKER = r"""
import sys
import numpy as np
from scipy.spatial import Delaunay
path, out, lo, hi, pad = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5])
pts = np.load(path, mmap_mode="r")
x = np.asarray(pts[:, 0])
idx = np.where((x >= lo - pad) & (x <= hi + pad))[0]
tri = idx[Delaunay(np.asarray(pts[idx])).simplices]
cx = x[tri].mean(axis=1)
np.save(out, tri[(cx >= lo) & (cx < hi)].astype(np.int64))
"""


def _delaunay(points: np.ndarray, min_parallel: int = 400_000) -> np.ndarray:
    """Delaunay triangulation (Qhull).  Large quasi-uniform point sets are cut into vertical
    strips that overlap by a few lattice spacings and are triangulated by parallel worker
    processes: a triangle away from the padded strip's edge is the same in every strip that
    contains it (the Delaunay triangulation is local), and each one is kept by the strip its
    centroid falls into.  The single-threaded Qhull call is 70 % of the synthetic-mesh
    set-up (setup only: not part of any measured region)."""
    import os
    import subprocess
    import sys
    import tempfile

    from scipy.spatial import Delaunay

    n = len(points)
    nproc = min(os.cpu_count() or 1, 16, n // 200_000)
    if n < min_parallel or nproc < 2 or os.environ.get("TDGL_B200_SERIAL_MESH"):
        return Delaunay(points).simplices.astype(np.int64)
    span = np.ptp(points, axis=0)
    pad = 8.0 * np.sqrt(span[0] * span[1] / n)
    cuts = np.quantile(points[:, 0], np.linspace(0, 1, nproc + 1))
    cuts[0], cuts[-1] = -1e300, 1e300
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm) as tmp:
        path = os.path.join(tmp, "points.npy")
        np.save(path, np.ascontiguousarray(points))
        env = dict(os.environ, OMP_NUM_THREADS="1")
        procs = [subprocess.Popen([sys.executable, "-c", _STRIP_WORKER, path,
                                   os.path.join(tmp, f"tri{i}.npy"), repr(float(cuts[i])),
                                   repr(float(cuts[i + 1])), repr(float(pad))], env=env)
                 for i in range(nproc)]
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("strip triangulation worker failed")
        parts = [np.load(os.path.join(tmp, f"tri{i}.npy")) for i in range(nproc)]
    return np.concatenate(parts)


def triangulate(points: np.ndarray, holes: Sequence[Tuple[float, float, float]] = ()):
    """Delaunay triangulation (Qhull); triangles inside holes and degenerate slivers on
    straight boundaries are dropped; triangles are oriented counter-clockwise."""
    tri = _delaunay(points)
    p = points[tri]
    area2 = ((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
             - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0]))
    flip = area2 < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    keep = np.abs(area2) > 1e-10 * np.abs(area2).max()
    cen = p.mean(axis=1)
    for (cx, cy, r) in holes:
        keep &= np.hypot(cen[:, 0] - cx, cen[:, 1] - cy) > r * (1 - 1e-9) - 1e-12
    tri = tri[keep]
    # drop unused points (none expected) and renumber
    used = np.zeros(len(points), dtype=bool)
    used[tri.ravel()] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        points = points[used]
        tri = remap[tri]
    return points, tri


def make_film_mesh(width: float, height: float, h: float,
                   holes: Sequence[Tuple[float, float, float]] = (),
                   jitter: float = 0.15, seed: int = 0, reorder: bool = True) -> Mesh:
    """Synthetic film mesh; ``reorder`` sorts the sites along a space-filling (Morton)
    curve so that CSR rows that are close in memory are close in space."""
    pts = make_film_points(width, height, h, holes, jitter, seed)
    pts, tri = triangulate(pts, holes)
    if reorder:
        perm = morton_order(pts)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        pts = pts[perm]
        tri = inv[tri]
    return Mesh.from_triangulation(pts, tri)


def morton_order(points: np.ndarray, bits: int = 20) -> np.ndarray:
    """Permutation sorting 2D points along a Z-order curve."""
    p = points - points.min(axis=0)
    span = p.max()
    q = np.minimum((p / span * ((1 << bits) - 1)).astype(np.uint64), (1 << bits) - 1)

    def spread(v):
        v = v & np.uint64(0xFFFFFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return np.argsort(code, kind="stable")
