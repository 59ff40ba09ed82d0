"""Prototype (CPU, SciPy) of the mu solver design: smoothed-aggregation AMG as the
preconditioner of CG on the area-symmetrised Neumann Laplacian.  Used to choose the
parameters of the CUDA implementation and as a host-side cross-check of its hierarchy;
TEST INFRASTRUCTURE, never on the product path."""

from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp

try:
    import numba
except ImportError:  # pragma: no cover
    numba = None


def sym_mu_matrix(edges, edge_len, dual_len, n):
    """A = -diag(areas) @ mu_laplacian: A_ij = -w_e, A_ii = sum w_e (SPSD, null = 1)."""
    w = dual_len / edge_len
    i0, i1 = edges[:, 0], edges[:, 1]
    A = sp.coo_array((np.concatenate([-w, -w, w, w]),
                      (np.concatenate([i0, i1, i0, i1]), np.concatenate([i1, i0, i0, i1]))),
                     shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def _aggregate_py(indptr, indices, data, theta):
    n = len(indptr) - 1
    diag = np.zeros(n)
    for i in range(n):
        for k in range(indptr[i], indptr[i + 1]):
            if indices[k] == i:
                diag[i] = data[k]
    agg = -np.ones(n, dtype=np.int64)
    nagg = 0
    # pass 1: root + all strong neighbours free
    for i in range(n):
        if agg[i] >= 0:
            continue
        ok = True
        cnt = 0
        for k in range(indptr[i], indptr[i + 1]):
            j = indices[k]
            if j == i:
                continue
            if data[k] * data[k] < theta * theta * diag[i] * diag[j]:
                continue
            cnt += 1
            if agg[j] >= 0:
                ok = False
                break
        if ok and cnt > 0:
            agg[i] = nagg
            for k in range(indptr[i], indptr[i + 1]):
                j = indices[k]
                if j != i and data[k] * data[k] >= theta * theta * diag[i] * diag[j]:
                    agg[j] = nagg
            nagg += 1
    # pass 2: join the most strongly connected neighbouring aggregate (of pass 1)
    agg2 = agg.copy()
    for i in range(n):
        if agg[i] >= 0:
            continue
        best = -1
        bestv = 0.0
        for k in range(indptr[i], indptr[i + 1]):
            j = indices[k]
            if j != i and agg[j] >= 0 and -data[k] > bestv:
                bestv = -data[k]
                best = agg[j]
        if best >= 0:
            agg2[i] = best
    # pass 3: leftovers
    for i in range(n):
        if agg2[i] < 0:
            agg2[i] = nagg
            for k in range(indptr[i], indptr[i + 1]):
                j = indices[k]
                if agg2[j] < 0:
                    agg2[j] = nagg
            nagg += 1
    return agg2, nagg


_aggregate = numba.njit(cache=True)(_aggregate_py) if numba is not None else _aggregate_py


def rho_DinvA(A, iters=30, seed=0):
    d = A.diagonal()
    rng = np.random.default_rng(seed)
    x = rng.normal(size=A.shape[0])
    lam = 1.0
    for _ in range(iters):
        y = (A @ x) / d
        lam = np.linalg.norm(y) / np.linalg.norm(x)
        x = y / np.linalg.norm(y)
    return lam


class Level:
    pass


def build_hierarchy(A, theta=0.0, omega_scale=4.0 / 3.0, max_coarse=200, max_levels=12,
                    smooth_P=True, verbose=False):
    levels = []
    B = np.ones(A.shape[0])          # near-null-space vector carried down the hierarchy
    while True:
        lv = Level()
        lv.A = A.tocsr()
        lv.d = A.diagonal()
        lv.rho = rho_DinvA(lv.A)
        levels.append(lv)
        n = A.shape[0]
        if n <= max_coarse or len(levels) >= max_levels:
            break
        agg, nagg = _aggregate(lv.A.indptr.astype(np.int64), lv.A.indices.astype(np.int64),
                               lv.A.data, theta)
        nrm = np.sqrt(np.bincount(agg, weights=B * B, minlength=nagg))
        T = sp.csr_array((B / nrm[agg], (np.arange(n), agg)), shape=(n, nagg))
        lv.B = B
        B = nrm
        if smooth_P:
            omega = omega_scale / lv.rho
            P = T - sp.diags_array(omega / lv.d) @ (lv.A @ T)
        else:
            P = T
        P = P.tocsr()
        lv.P = P
        lv.R = P.T.tocsr()
        A = (lv.R @ lv.A @ P).tocsr()
        A.sum_duplicates()
        if verbose:
            print(f"  level {len(levels)-1}: n={n} nnz={lv.A.nnz} ({lv.A.nnz/n:.1f}/row)"
                  f" -> nagg={nagg} P nnz/row={P.nnz/n:.2f} rho={lv.rho:.3f}")
    last = levels[-1]
    last.B = B
    Ad = last.A.toarray()
    g = np.trace(Ad) / len(B) / (B @ B)
    last.pinv = np.linalg.inv(Ad + g * np.outer(B, B))
    if verbose:
        print(f"  coarsest: n={last.A.shape[0]} nnz={last.A.nnz}")
    return levels


def op_complexity(levels):
    return sum(l.A.nnz for l in levels) / levels[0].A.nnz


def cheb_coeffs(lo, hi, degree):
    """Coefficients for the Chebyshev smoother iteration on [lo, hi]."""
    theta = 0.5 * (hi + lo)
    delta = 0.5 * (hi - lo)
    return theta, delta


def smooth(lv, x, b, kind, sweeps, zero_guess=False):
    A, d = lv.A, lv.d
    if kind == "jacobi":
        w = (4.0 / 3.0) / lv.rho
        for s in range(sweeps):
            if zero_guess and s == 0:
                x = w * b / d
            else:
                x = x + w * (b - A @ x) / d
        return x
    if kind == "l1jacobi":
        l1 = np.asarray(abs(A).sum(axis=1)).ravel()
        for s in range(sweeps):
            if zero_guess and s == 0:
                x = b / l1
            else:
                x = x + (b - A @ x) / l1
        return x
    if kind == "cheb":
        # Chebyshev polynomial smoother of degree `sweeps` on D^-1 A, interval
        # [rho/30... use rho*a, 1.1 rho]
        hi = 1.1 * lv.rho
        lo = hi / lv.cheb_ratio
        theta = 0.5 * (hi + lo)
        delta = 0.5 * (hi - lo)
        sigma = theta / delta
        rho_k = 1.0 / sigma
        r = b if zero_guess else b - A @ x
        dvec = (r / d) / theta
        x = dvec.copy() if zero_guess else x + dvec
        for _ in range(sweeps - 1):
            rho_n = 1.0 / (2 * sigma - rho_k)
            r = b - A @ x
            dvec = rho_n * rho_k * dvec + (2 * rho_n / delta) * (r / d)
            x = x + dvec
            rho_k = rho_n
        return x
    raise ValueError(kind)


def vcycle(levels, b, k=0, kind="jacobi", pre=1, post=1, counter=None):
    lv = levels[k]
    if k == len(levels) - 1:
        return lv.pinv @ b
    x = smooth(lv, None, b, kind, pre, zero_guess=True)
    r = b - lv.A @ x
    rc = lv.R @ r
    xc = vcycle(levels, rc, k + 1, kind, pre, post, counter)
    x = x + lv.P @ xc
    x = smooth(lv, x, b, kind, post)
    return x


def pcg(A, b, M, x0=None, rtol=1e-10, maxiter=200, bnorm=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    bn = np.linalg.norm(b) if bnorm is None else bnorm
    z = M(r)
    p = z.copy()
    rz = r @ z
    hist = [np.linalg.norm(r) / bn]
    for it in range(maxiter):
        if hist[-1] < rtol:
            break
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        hist.append(np.linalg.norm(r) / bn)
        if hist[-1] < rtol:
            break
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, hist


def work_per_cycle(levels, kind, pre, post):
    """fine-SpMV equivalents (by nnz) of one V-cycle + the CG SpMV."""
    nnz0 = levels[0].A.nnz
    w = 1.0  # CG's own SpMV
    for lv in levels[:-1]:
        spmv = lv.A.nnz / nnz0
        pr = (lv.P.nnz * 2) / nnz0
        # pre (first sweep free with zero guess), residual, post
        w += spmv * ((pre - 1) + 1 + post) + pr
    return w


if __name__ == "__main__":
    import sys

    sys.path.insert(0, "/root/repo")
    from tdgl_b200.mesh import make_film_mesh

    size = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    t0 = time.time()
    mesh = make_film_mesh(size, size, 0.43)
    n = len(mesh.sites)
    em = mesh.edge_mesh
    print(f"mesh {n} sites in {time.time()-t0:.1f}s")
    A = sym_mu_matrix(em.edges, em.edge_lengths, em.dual_edge_lengths, n)
    rng = np.random.default_rng(1)
    b = rng.normal(size=n)
    b -= b.mean()
    # a smooth physical-like rhs too
    bs = mesh.areas * np.sin(mesh.sites[:, 0] / 7.0) * np.cos(mesh.sites[:, 1] / 5.0)
    bs -= bs.mean()
    for theta in (0.0,):
        t0 = time.time()
        levels = build_hierarchy(A, theta=theta, verbose=True)
        print(f"theta={theta}: setup {time.time()-t0:.1f}s levels={len(levels)}"
              f" op complexity={op_complexity(levels):.3f}")
        for kind, pre, post, ratio in [("jacobi", 1, 1, 0), ("jacobi", 2, 2, 0),
                                       ("cheb", 2, 2, 10.0), ("cheb", 3, 3, 10.0),
                                       ("cheb", 2, 2, 4.0), ("cheb", 3, 3, 30.0),
                                       ("l1jacobi", 1, 1, 0)]:
            for lv in levels:
                lv.cheb_ratio = ratio
            M = lambda r: vcycle(levels, r, 0, kind, pre, post)  # noqa: E731
            for name, rhs in (("random", b), ("smooth", bs)):
                x, hist = pcg(A, rhs, M)
                w = work_per_cycle(levels, kind, pre, post)
                print(f"   {kind}({pre},{post}) r={ratio}: {name}: its={len(hist)-1}"
                      f" work/it={w:.2f} total={w*(len(hist)-1):.1f} spmv-eq"
                      f" final={hist[-1]:.1e}")
