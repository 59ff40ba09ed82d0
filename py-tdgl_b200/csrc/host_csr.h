// Host-side CSR containers and the setup-time algebra of the engine (no CUDA here).
//
// Everything in this header runs once per tdgl_create(): building the finite-volume
// operators from the mesh arrays (reference tdgl/finite_volume/operators.py:59-230) and
// the smoothed-aggregation hierarchy that preconditions the mu solve.  The per-step work
// is in kernels.cuh.
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <numeric>
#include <cstdlib>
#include <stdexcept>
#include <system_error>
#include <thread>
#include <vector>

#include <sched.h>

namespace tdgl {

// Threads of the setup-time loops: the cores this process may run on (at most 16), or
// TDGL_B200_HOST_THREADS.  Every parallel loop below splits ROWS into contiguous chunks and
// leaves each row's arithmetic (and every reduction over rows) in the serial order, so the
// operators and the hierarchy are bit-for-bit the same for any thread count — the ranks of a
// sharded run must build identical hierarchies.
inline int host_threads() {
  if (const char* e = std::getenv("TDGL_B200_HOST_THREADS")) return std::max(1, std::atoi(e));
  cpu_set_t set;
  int n = 0;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (n <= 0) n = static_cast<int>(std::thread::hardware_concurrency());
  return std::max(1, std::min(n, 16));
}

// fn(chunk, lo, hi) over `chunks` contiguous ranges of [0, n); the caller's thread takes the
// last one.  An exception in a worker is rethrown here.
template <typename F>
inline void parallel_chunks(int64_t n, int chunks, F&& fn) {
  if (chunks <= 1) { fn(0, int64_t{0}, n); return; }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> err(chunks);
  auto run = [&](int c) {
    try { fn(c, n * c / chunks, n * (c + 1) / chunks); } catch (...) { err[c] = std::current_exception(); }
  };
  pool.reserve(chunks);
  for (int c = 0; c + 1 < chunks; ++c) {
    // (no thread to be had — a pids limit, say: the chunk runs here)
    try { pool.emplace_back(run, c); } catch (const std::system_error&) { run(c); }
  }
  run(chunks - 1);
  for (auto& t : pool) t.join();
  for (auto& e : err) if (e) std::rethrow_exception(e);
}

// Number of chunks for a loop over n rows (small loops stay on the caller's thread).
inline int chunks_for(int64_t n, int64_t min_rows_per_chunk = 16384) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(host_threads(), n / min_rows_per_chunk)));
}

template <typename T>
struct HostCsr {
  int64_t rows = 0, cols = 0;
  std::vector<int32_t> ptr;  // rows + 1
  std::vector<int32_t> idx;  // nnz, sorted within a row
  std::vector<T> val;        // nnz
  int64_t nnz() const { return static_cast<int64_t>(idx.size()); }
};

// Site adjacency of the triangulation: for every site the incident edges, sorted by
// neighbour index, plus a leading diagonal slot.  All site operators (covariant
// Laplacian, mu Laplacian, divergence) share this one structure.
struct SiteGraph {
  int64_t n = 0;
  std::vector<int32_t> ptr;   // n + 1
  std::vector<int32_t> nbr;   // neighbour site (the row itself for the diagonal slot)
  std::vector<int32_t> edge;  // edge index of the slot (-1 for the diagonal slot)
  std::vector<int8_t> head;   // 1 if the row is edges[e,0] (link variable U_e), 0 if
                              // edges[e,1] (conj U_e); 0 on the diagonal slot
};

inline SiteGraph build_site_graph(int64_t n, int64_t n_edges, const int32_t* e0,
                                  const int32_t* e1) {
  SiteGraph g;
  g.n = n;
  g.ptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) g.ptr[i + 1] = 1;  // diagonal slot
  for (int64_t e = 0; e < n_edges; ++e) {
    if (e0[e] < 0 || e0[e] >= n || e1[e] < 0 || e1[e] >= n || e0[e] == e1[e])
      throw std::invalid_argument("edge index out of range");
    g.ptr[e0[e] + 1]++;
    g.ptr[e1[e] + 1]++;
  }
  for (int64_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
  const int64_t nnz = g.ptr[n];
  g.nbr.resize(nnz);
  g.edge.resize(nnz);
  g.head.resize(nnz);
  std::vector<int32_t> fill(g.ptr.begin(), g.ptr.end() - 1);
  for (int64_t i = 0; i < n; ++i) {
    int32_t k = fill[i]++;
    g.nbr[k] = static_cast<int32_t>(i);
    g.edge[k] = -1;
    g.head[k] = 0;
  }
  for (int64_t e = 0; e < n_edges; ++e) {
    int32_t k = fill[e0[e]]++;
    g.nbr[k] = e1[e];
    g.edge[k] = static_cast<int32_t>(e);
    g.head[k] = 1;
    k = fill[e1[e]]++;
    g.nbr[k] = e0[e];
    g.edge[k] = static_cast<int32_t>(e);
    g.head[k] = 0;
  }
  // sort every row by neighbour index (diagonal lands in its natural position)
  parallel_chunks(n, chunks_for(n), [&](int, int64_t lo, int64_t hi) {
    std::vector<int32_t> order;
    std::vector<int32_t> tn, te;
    std::vector<int8_t> th;
    for (int64_t i = lo; i < hi; ++i) {
      const int32_t b = g.ptr[i], len = g.ptr[i + 1] - b;
      order.resize(len);
      std::iota(order.begin(), order.end(), 0);
      std::sort(order.begin(), order.end(),
                [&](int32_t a, int32_t c) { return g.nbr[b + a] < g.nbr[b + c]; });
      tn.resize(len); te.resize(len); th.resize(len);
      for (int32_t k = 0; k < len; ++k) {
        tn[k] = g.nbr[b + order[k]]; te[k] = g.edge[b + order[k]]; th[k] = g.head[b + order[k]];
      }
      for (int32_t k = 0; k < len; ++k) {
        if (k > 0 && tn[k] == tn[k - 1]) throw std::invalid_argument("duplicate edge");
        g.nbr[b + k] = tn[k]; g.edge[b + k] = te[k]; g.head[b + k] = th[k];
      }
    }
  });
  return g;
}

// y = A x
template <typename T>
inline void spmv(const HostCsr<T>& A, const std::vector<T>& x, std::vector<T>& y) {
  y.resize(A.rows);
  parallel_chunks(A.rows, chunks_for(A.rows), [&](int, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      T s = T(0);
      for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) s += A.val[k] * x[A.idx[k]];
      y[i] = s;
    }
  });
}

template <typename T>
inline HostCsr<T> transpose(const HostCsr<T>& A) {
  HostCsr<T> B;
  B.rows = A.cols; B.cols = A.rows;
  B.ptr.assign(B.rows + 1, 0);
  for (int64_t k = 0; k < A.nnz(); ++k) B.ptr[A.idx[k] + 1]++;
  for (int64_t i = 0; i < B.rows; ++i) B.ptr[i + 1] += B.ptr[i];
  B.idx.resize(A.nnz()); B.val.resize(A.nnz());
  std::vector<int32_t> fill(B.ptr.begin(), B.ptr.end() - 1);
  for (int64_t i = 0; i < A.rows; ++i)
    for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      int32_t d = fill[A.idx[k]]++;
      B.idx[d] = static_cast<int32_t>(i);
      B.val[d] = A.val[k];
    }
  return B;  // rows come out sorted because i ascends
}

// C = A * B (Gustavson, dense accumulator over B.cols); rows of C sorted.  Row chunks are
// multiplied by separate threads (each with its own accumulator) and stitched together.
template <typename T>
inline HostCsr<T> spgemm(const HostCsr<T>& A, const HostCsr<T>& B) {
  if (A.cols != B.rows) throw std::invalid_argument("spgemm shape");
  HostCsr<T> C;
  C.rows = A.rows; C.cols = B.cols;
  C.ptr.assign(C.rows + 1, 0);
  const int chunks = chunks_for(A.rows);
  std::vector<std::vector<int32_t>> cidx(chunks);
  std::vector<std::vector<T>> cval(chunks);
  std::vector<int64_t> first(chunks + 1, 0);
  parallel_chunks(A.rows, chunks, [&](int c_, int64_t lo, int64_t hi) {
    std::vector<T> acc(B.cols, T(0));
    std::vector<int32_t> mark(B.cols, -1), touched;
    std::vector<int32_t> oi;   // (locals: the headers of cidx[c_] of neighbouring chunks share
    std::vector<T> ov;         //  a cache line, and push_back writes the header)
    first[c_] = lo;
    {
      int64_t bound = 0;   // products formed in this chunk >= entries produced
      for (int64_t i = lo; i < hi; ++i)
        for (int32_t ka = A.ptr[i]; ka < A.ptr[i + 1]; ++ka) bound += B.ptr[A.idx[ka] + 1] - B.ptr[A.idx[ka]];
      oi.reserve(bound);
      ov.reserve(bound);
    }
    for (int64_t i = lo; i < hi; ++i) {
      touched.clear();
      for (int32_t ka = A.ptr[i]; ka < A.ptr[i + 1]; ++ka) {
        const int32_t j = A.idx[ka];
        const T a = A.val[ka];
        for (int32_t kb = B.ptr[j]; kb < B.ptr[j + 1]; ++kb) {
          const int32_t c = B.idx[kb];
          if (mark[c] != i) { mark[c] = static_cast<int32_t>(i); acc[c] = T(0); touched.push_back(c); }
          acc[c] += a * B.val[kb];
        }
      }
      std::sort(touched.begin(), touched.end());
      for (int32_t c : touched) { oi.push_back(c); ov.push_back(acc[c]); }
      C.ptr[i + 1] = static_cast<int32_t>(oi.size());   // chunk-relative until stitched
    }
    cidx[c_] = std::move(oi);
    cval[c_] = std::move(ov);
  });
  first[chunks] = A.rows;
  int64_t total = 0;
  std::vector<int64_t> base(chunks, 0);
  for (int c = 0; c < chunks; ++c) { base[c] = total; total += static_cast<int64_t>(cidx[c].size()); }
  if (total > 0x7FFFFFF0ll) throw std::runtime_error("spgemm result exceeds 32-bit indexing");
  C.idx.resize(total);
  C.val.resize(total);
  parallel_chunks(chunks, chunks, [&](int, int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1; ++c) {
      std::copy(cidx[c].begin(), cidx[c].end(), C.idx.begin() + base[c]);
      std::copy(cval[c].begin(), cval[c].end(), C.val.begin() + base[c]);
      if (base[c] != 0)
        for (int64_t i = first[c]; i < first[c + 1]; ++i) C.ptr[i + 1] += static_cast<int32_t>(base[c]);
    }
  });
  return C;
}

template <typename T>
inline std::vector<T> diagonal(const HostCsr<T>& A) {
  std::vector<T> d(A.rows, T(0));
  for (int64_t i = 0; i < A.rows; ++i)
    for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (A.idx[k] == i) d[i] = A.val[k];
  return d;
}

// Largest eigenvalue of D^-1 A by power iteration (deterministic start vector).
inline double rho_dinv_a(const HostCsr<double>& A, const std::vector<double>& d, int iters = 30) {
  const int64_t n = A.rows;
  std::vector<double> x(n), y(n);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (int64_t i = 0; i < n; ++i) {  // splitmix64 -> uniform(-1, 1)
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    x[i] = (static_cast<double>(z >> 11) / 9007199254740992.0) * 2.0 - 1.0;
  }
  double lam = 2.0;
  const int chunks = chunks_for(n);
  for (int it = 0; it < iters; ++it) {
    spmv(A, x, y);
    parallel_chunks(n, chunks, [&](int, int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i) y[i] /= d[i];
    });
    double ny = 0, nx = 0;   // (serial sums: the same bits for any thread count)
    for (int64_t i = 0; i < n; ++i) { ny += y[i] * y[i]; nx += x[i] * x[i]; }
    if (ny == 0 || nx == 0) return 2.0;
    lam = std::sqrt(ny / nx);
    const double inv = 1.0 / std::sqrt(ny);
    parallel_chunks(n, chunks, [&](int, int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i) x[i] = y[i] * inv;
    });
  }
  return lam;
}

}  // namespace tdgl
