"""Synthetic workloads of BASELINE.json's configs, in dimensionless form:
lengths in xi, fields in B_c2, sheet currents in K_0/4 (what the reference works in after
``TDGLSolver.__init__`` applied ``A_scale`` / ``J_scale``, solver.py:176-185,251-256)."""

from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np

from .mesh import Mesh, make_film_mesh


class TerminalInfo(NamedTuple):
    """Same fields as the reference's ``TerminalInfo`` (device/device.py:30-46)."""

    name: str
    site_indices: Sequence[int]
    edge_indices: Sequence[int]
    boundary_edge_indices: Sequence[int]
    length: float


def uniform_field_vector_potential(edge_centers: np.ndarray, b: float) -> np.ndarray:
    """Symmetric-gauge A of a uniform field ``b`` (units of B_c2) centred on the bounding
    box of the evaluation points, as the reference's ``ConstantField`` does
    (em.py:463-471): A = b/2 * (-y, x)."""
    xs = edge_centers[:, 0]
    ys = edge_centers[:, 1]
    xs = xs - (xs.min() + np.ptp(xs) / 2)
    ys = ys - (ys.min() + np.ptp(ys) / 2)
    return np.stack([-b * ys / 2, b * xs / 2], axis=1)


def box_terminal(mesh: Mesh, name: str, xmin, xmax, ymin, ymax) -> TerminalInfo:
    """Terminal = boundary sites / boundary edges inside an axis-aligned box, with the
    index conventions of the reference's ``Device.terminal_info`` (device.py:221-256):
    ``boundary_edge_indices`` index INTO the list of boundary edges."""
    em = mesh.edge_mesh

    def inside(p):
        return (p[:, 0] >= xmin) & (p[:, 0] <= xmax) & (p[:, 1] >= ymin) & (p[:, 1] <= ymax)

    sites = np.intersect1d(np.where(inside(mesh.sites))[0], mesh.boundary_indices)
    ib = em.boundary_edge_indices
    b_in = np.where(inside(em.centers[ib]))[0]
    length = float(em.edge_lengths[ib][b_in].sum())
    return TerminalInfo(name, sites, ib[b_in], b_in, length)


def gaussian_disorder(sites: np.ndarray, depth: float = 0.8, width2: float = 8.0):
    """epsilon(r) = 1 - depth * exp(-|r|^2 / width2): the deterministic perturbation
    used for parity runs (SURVEY.md §8c last bullet)."""
    r2 = (sites**2).sum(axis=1)
    return 1.0 - depth * np.exp(-r2 / width2)


def film_problem(width, height, h, b=0.0, holes=(), seed=0, disorder=False,
                 terminals=False, reorder=True):
    """Mesh + dimensionless inputs of a rectangular-film workload.  ``terminals`` puts a
    ``source`` on the full left edge and a ``drain`` on the full right edge."""
    mesh = make_film_mesh(width, height, h, holes=holes, seed=seed, reorder=reorder)
    A = uniform_field_vector_potential(mesh.edge_mesh.centers, b)
    eps = gaussian_disorder(mesh.sites) if disorder else np.ones(len(mesh.sites))
    terms = ()
    if terminals:
        tol = 1e-9 * max(width, height)
        x0, x1 = -width / 2, width / 2
        big = 10 * max(width, height)
        terms = (box_terminal(mesh, "source", x0 - tol, x0 + tol, -big, big),
                 box_terminal(mesh, "drain", x1 - tol, x1 + tol, -big, big))
        terms = tuple(sorted(terms, key=lambda t: t.length))
    return mesh, A, eps, terms


def vortex_state(mesh: Mesh, holes=(), fixed_sites=None, q: float = 0.0, b: float = 0.0,
                 seed: int = 0):
    """A deterministic analytic state with the features of the developed dynamics — vortex /
    antivortex pairs leaving the holes, a few vortices that entered from the film edges, and
    the phase gradient ``q`` of a transport current — for benchmarks that have to start both
    the CUDA engine and the CPU reference from the SAME developed-looking state without
    running thousands of warm-up steps on the CPU.  Returns (psi [N] complex, mu [N] = 0).

    psi = prod_k (z - z_k)^{+-1 direction} / sqrt(|z - z_k|^2 + 2)  *  exp(i q x)
    (each factor is the usual single-vortex ansatz with a core of ~xi)."""
    z = mesh.sites[:, 0] + 1j * mesh.sites[:, 1]
    x0, x1 = mesh.sites[:, 0].min(), mesh.sites[:, 0].max()
    y0, y1 = mesh.sites[:, 1].min(), mesh.sites[:, 1].max()
    rng = np.random.default_rng(seed)
    centres, signs = [], []
    for (hx, hy, hr) in holes:
        centres += [complex(hx, hy + 1.5 * hr), complex(hx, hy - 1.5 * hr)]
        signs += [+1, -1]
    n_edge = 8 if (holes or b) else 0
    for k in range(n_edge):
        centres.append(complex(x0 + (x1 - x0) * rng.uniform(0.1, 0.9),
                               (y1 - 4.0 * (1 + k % 3)) if k % 2 == 0 else (y0 + 4.0 * (1 + k % 3))))
        signs.append(+1 if (k % 2 == 0 or b) else -1)
    psi = np.ones(len(z), dtype=np.complex128)
    for c, s in zip(centres, signs):
        d = z - c
        f = d / np.sqrt(np.abs(d) ** 2 + 2.0)
        psi *= f if s > 0 else np.conj(f)
    if q:
        psi *= np.exp(1j * q * mesh.sites[:, 0])
    if fixed_sites is not None and len(fixed_sites):
        psi[np.asarray(fixed_sites, dtype=np.int64)] = 0.0
    return psi, np.zeros(len(z))
