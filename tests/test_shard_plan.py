"""Domain decomposition, host side (no GPU): the plan the sharded engine builds, validated by
running the sharded AMG-PCG with all shards emulated on the host (same local operators, same
exchange lists as the device path), and a world_size-2 gloo run that performs the level-0
halo exchange for real between two processes."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from tdgl_b200.engine import host_amg_probe, host_shard_lists, host_shard_probe
from tdgl_b200.synthetic import film_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sym_mu_matrix(mesh):
    em = mesh.edge_mesh
    n = len(mesh.sites)
    w = em.dual_edge_lengths / em.edge_lengths
    i0, i1 = em.edges[:, 0], em.edges[:, 1]
    return sp.csr_array((np.concatenate([-w, -w, w, w]),
                         (np.concatenate([i0, i1, i0, i1]), np.concatenate([i1, i0, i0, i1]))),
                        shape=(n, n))


@pytest.fixture(scope="module")
def problem():
    mesh, A, eps, _ = film_problem(40, 30, 0.43, b=0.1, holes=((5.0, 3.0, 4.0),))
    rng = np.random.default_rng(3)
    x0 = rng.normal(size=len(mesh.sites))
    return mesh, _sym_mu_matrix(mesh) @ x0


@pytest.mark.parametrize("world,replicate_below", [(2, 0), (3, 0), (4, 0), (8, 0), (2, 100),
                                                   (4, 500), (8, 1)])
def test_sharded_pcg_matches_single(problem, world, replicate_below):
    """replicate_below: AMG levels with at most that many rows are computed redundantly by
    every shard (default 32768: only the fine level of this small mesh is partitioned;
    100 / 500 / 1 partition two, one-or-two and all levels)."""
    mesh, rhs = problem
    g = host_amg_probe(mesh, rhs=rhs)
    p = host_shard_probe(mesh, world, rhs=rhs, replicate_below=replicate_below)
    assert 1 <= p["rep"] <= p["levels"] - 1
    n = len(mesh.sites)
    off = p["offsets"]
    assert off[0, 0] == 0 and off[0, -1] == n
    assert np.all(np.diff(off, axis=1) >= 0)
    assert sorted(p["perm"]) == list(range(n))
    # same preconditioner up to a renumbering of the aggregates: same convergence
    assert abs(p["iterations"] - g["iterations"]) <= 2, (p["iterations"], g["iterations"])
    a, b = p["x"] - p["x"].mean(), g["x"] - g["x"].mean()
    assert np.abs(a - b).max() < 1e-8 * np.abs(b).max()
    A = _sym_mu_matrix(mesh)
    assert np.linalg.norm(A @ p["x"] - rhs) < 1e-9 * np.linalg.norm(rhs)


def test_exchange_lists_are_consistent(problem):
    mesh, _ = problem
    world = 4
    lists = [host_shard_lists(mesh, world, r) for r in range(world)]
    n = len(mesh.sites)
    owned = np.concatenate([s["owned"] for s in lists])
    assert sorted(owned) == list(range(n))                      # a partition of the sites
    A = _sym_mu_matrix(mesh)
    for r, s in enumerate(lists):
        mine = set(s["owned"].tolist())
        assert not mine.intersection(s["halo"].tolist())
        # every neighbour of an owned site is owned or in the halo
        cols = set(A[s["owned"]].indices.tolist())
        assert cols <= mine.union(s["halo"].tolist())
        assert len(s["send"][r]) == 0
        # what q sends to r is exactly r's halo entries owned by q, in halo order
        got = np.concatenate([lists[q]["send"][r] for q in range(world)])
        owner = np.empty(n, dtype=int)
        for q in range(world):
            owner[lists[q]["owned"]] = q
        order = np.argsort(owner[s["halo"]], kind="stable")
        np.testing.assert_array_equal(got, s["halo"][order])


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
from tdgl_b200.engine import host_shard_lists
from tdgl_b200.synthetic import film_problem
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mesh, A, eps, _ = film_problem(40, 30, 0.43, b=0.1, holes=((5.0, 3.0, 4.0),))
n = len(mesh.sites)
em = mesh.edge_mesh
w = em.dual_edge_lengths / em.edge_lengths
i0, i1 = em.edges[:, 0], em.edges[:, 1]
L = sp.csr_array((np.concatenate([-w, -w, w, w]), (np.concatenate([i0, i1, i0, i1]),
                 np.concatenate([i1, i0, i0, i1]))), shape=(n, n))
s = host_shard_lists(mesh, world, rank)
x = np.random.default_rng(7).normal(size=n)          # the same whole-mesh vector everywhere
local = np.concatenate([x[s["owned"]], np.full(len(s["halo"]), np.nan)])   # halo unknown
# halo exchange over gloo: send my boundary entries, receive the peers' into my halo slots
owner = np.empty(n, dtype=int)
owned_all = [None] * world
dist.all_gather_object(owned_all, s["owned"])
for q in range(world):
    owner[owned_all[q]] = q
reqs, recv_bufs = [], {}
for q in range(world):
    if q == rank:
        continue
    out = torch.from_numpy(np.ascontiguousarray(x[s["send"][q]]))
    if len(out):
        reqs.append(dist.isend(out, q))
    cnt = int(np.sum(owner[s["halo"]] == q))
    if cnt:
        recv_bufs[q] = torch.empty(cnt, dtype=torch.float64)
        reqs.append(dist.irecv(recv_bufs[q], q))
for r in reqs:
    r.wait()
pos = len(s["owned"])
for q in sorted(recv_bufs):                           # the halo is grouped by owner, in rank order
    k = len(recv_bufs[q])
    local[pos:pos + k] = recv_bufs[q].numpy()
    pos += k
assert pos == len(local) and not np.isnan(local).any()
ids = np.concatenate([s["owned"], s["halo"]])
np.testing.assert_array_equal(local, x[ids])           # every halo slot got the right entry
# local rows of the operator applied to [owned | halo] == the same rows of the global product
g2l = -np.ones(n, dtype=int); g2l[ids] = np.arange(len(ids))
rows = L[s["owned"]]
y_local = np.array([np.dot(rows.data[rows.indptr[i]:rows.indptr[i + 1]],
                           local[g2l[rows.indices[rows.indptr[i]:rows.indptr[i + 1]]]])
                    for i in range(len(s["owned"]))])
np.testing.assert_allclose(y_local, (L @ x)[s["owned"]], rtol=1e-13, atol=1e-13)
# a dot product reduced in rank order gives every rank the same bits
part = torch.tensor([float(np.dot(x[s["owned"]], y_local))], dtype=torch.float64)
parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(parts, part)
total = 0.0
for p in parts:
    total += float(p)
assert abs(total - float(x @ (L @ x))) < 1e-9 * abs(total)
dist.barrier()
if rank == 0:
    print("GLOO_OK", world, len(s["owned"]), len(s["halo"]))
dist.destroy_process_group()
'''


def test_halo_exchange_world2_gloo(tmp_path):
    """Two processes, gloo backend: the level-0 halo exchange with the lists the sharded
    engine uses, then the local SpMV and a rank-ordered reduction."""
    import subprocess

    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER)
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "GLOO_OK 2" in res.stdout


def test_shard_plan_edge_cases(problem):
    """What cannot be sharded is refused loudly: more than 8 shards, a mesh so small that its
    AMG hierarchy has a single level; one shard is the identity plan."""
    from tdgl_b200._lib import TDGLLibraryError
    from tdgl_b200.mesh import make_film_mesh

    mesh, rhs = problem
    with pytest.raises(TDGLLibraryError, match="at most 8"):
        host_shard_probe(mesh, 9)
    tiny = make_film_mesh(4, 3, 0.5)          # ~60 sites: one level (<= 200 rows)
    with pytest.raises(TDGLLibraryError, match="too small to shard"):
        host_shard_probe(tiny, 2)
    one = host_shard_probe(mesh, 1, rhs=rhs)
    g = host_amg_probe(mesh, rhs=rhs)
    assert one["iterations"] == g["iterations"] and one["halo_sizes"].sum() == 0
    np.testing.assert_allclose(one["x"], g["x"], rtol=0, atol=1e-9 * np.abs(g["x"]).max())


def test_setup_is_bitwise_independent_of_the_thread_count(monkeypatch):
    """The setup-time loops (site graph, power iteration, sparse products, Z-order sort, shard
    extraction: csrc/host_csr.h `parallel_chunks`) split rows over host threads; every rank
    of a sharded run must still build the SAME hierarchy, whatever cores it was given.  The
    host probes run the whole setup + AMG-PCG, so equal solutions bit for bit and equal
    permutations / halo plans mean equal hierarchies."""
    mesh = film_problem(110, 100, 0.4, b=0.1, holes=((5.0, 3.0, 9.0),))[0]   # ~79k sites
    assert len(mesh.sites) > 4 * 16384   # enough rows for several chunks
    rng = np.random.default_rng(5)
    rhs = _sym_mu_matrix(mesh) @ rng.normal(size=len(mesh.sites))
    got = {}
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("TDGL_B200_HOST_THREADS", threads)
        got[threads] = (host_amg_probe(mesh, rhs=rhs), host_shard_probe(mesh, 4, rhs=rhs))
    g1, p1 = got["1"]
    for threads in ("3", "8"):
        g, p = got[threads]
        assert g["rows"] == g1["rows"] and g["nnz"] == g1["nnz"]
        assert g["iterations"] == g1["iterations"] and p["iterations"] == p1["iterations"]
        assert np.array_equal(g["x"], g1["x"])
        assert np.array_equal(p["x"], p1["x"])
        assert np.array_equal(p["perm"], p1["perm"])
        assert np.array_equal(p["offsets"], p1["offsets"])
        assert np.array_equal(p["halo_sizes"], p1["halo_sizes"])
