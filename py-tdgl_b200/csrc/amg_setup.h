// Host-side setup of the smoothed-aggregation hierarchy that preconditions the mu solve.
//
// The reference factors the (singular, pure-Neumann) mu Laplacian once with SuperLU and
// does a forward/back solve per step (tdgl/finite_volume/operators.py:285,306-308;
// tdgl/solver/solver.py:513-516).  On the GPU the same system is solved by CG on the
// area-symmetrised matrix  A = -diag(areas) * mu_laplacian  (A_ij = -w_e, A_ii = sum w_e,
// w_e = dual_edge_length / edge_length; SPSD with null space span{1}), preconditioned by
// one V-cycle of this hierarchy.  The matrix never changes during a solve, so the setup
// is amortised over 10^3..10^6 steps.
#pragma once

#include <thread>

#include "host_csr.h"

namespace tdgl {

struct AmgLevel {
  HostCsr<double> A;
  std::vector<double> dinv;  // 1 / diag(A)
  double rho = 2.0;          // spectral radius of D^-1 A
  HostCsr<double> P, R;      // prolongation to this level from the next, R = P^T
  std::vector<double> B;     // near-null-space vector on this level
};

struct AmgHierarchy {
  std::vector<AmgLevel> levels;
  std::vector<double> coarse_inv;  // dense (A_c + g B B^T)^-1, row-major nc x nc
  int64_t nc = 0;
  // Domain decomposition: off[l][r] .. off[l][r+1] is the contiguous row range of level l
  // owned by rank r (one entry {0, n} per level for a single rank).
  std::vector<std::vector<int64_t>> off;
};

// Rank owning row i of a level partitioned into the contiguous ranges off[r]..off[r+1].
inline int owner_of(const std::vector<int64_t>& off, int64_t i) {
  return static_cast<int>(std::upper_bound(off.begin(), off.end(), i) - off.begin()) - 1;
}

// Greedy aggregation on the strength graph  a_ij^2 >= theta^2 a_ii a_jj:
//  pass 1: a free node all of whose strong neighbours are free seeds an aggregate,
//  pass 2: remaining nodes join the pass-1 aggregate they are most strongly tied to,
//  pass 3: leftovers form aggregates with their still-free neighbours.
inline int64_t aggregate(const HostCsr<double>& A, const std::vector<double>& d, double theta,
                         std::vector<int32_t>& agg) {
  const int64_t n = A.rows;
  agg.assign(n, -1);
  const double t2 = theta * theta;
  auto strong = [&](int64_t i, int32_t k) {
    const int32_t j = A.idx[k];
    return j != i && A.val[k] * A.val[k] >= t2 * d[i] * d[j];
  };
  int32_t nagg = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    bool ok = true;
    int cnt = 0;
    for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      if (!strong(i, k)) continue;
      ++cnt;
      if (agg[A.idx[k]] >= 0) { ok = false; break; }
    }
    if (!ok || cnt == 0) continue;
    agg[i] = nagg;
    for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k)) agg[A.idx[k]] = nagg;
    ++nagg;
  }
  std::vector<int32_t> agg2(agg);
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    int32_t best = -1;
    double bestv = 0.0;
    for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int32_t j = A.idx[k];
      if (j != i && agg[j] >= 0 && -A.val[k] > bestv) { bestv = -A.val[k]; best = agg[j]; }
    }
    if (best >= 0) agg2[i] = best;
  }
  for (int64_t i = 0; i < n; ++i) {
    if (agg2[i] >= 0) continue;
    agg2[i] = nagg;
    for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (agg2[A.idx[k]] < 0) agg2[A.idx[k]] = nagg;
    ++nagg;
  }
  agg.swap(agg2);
  return nagg;
}

// In-place Cholesky inverse of a dense SPD matrix (row-major n x n).  The coarsest level may
// hold a couple of thousand rows (fewer, larger levels = fewer kernel launches per V-cycle), so
// the O(n^3) work is arranged in contiguous dot products and the n right-hand sides of the
// inverse are split over the host cores.
inline void dense_spd_inverse(std::vector<double>& M, int64_t n) {
  std::vector<double> L(M);   // lower triangle, row-major: row i = L[i*n .. i*n+i]
  for (int64_t j = 0; j < n; ++j) {
    const double* lj = &L[j * n];
    double s = lj[j];
    for (int64_t k = 0; k < j; ++k) s -= lj[k] * lj[k];
    if (!(s > 0)) throw std::runtime_error("coarse matrix not positive definite");
    const double ljj = std::sqrt(s);
    L[j * n + j] = ljj;
    const double inv = 1.0 / ljj;
    for (int64_t i = j + 1; i < n; ++i) {
      const double* li = &L[i * n];
      double t = li[j];
      for (int64_t k = 0; k < j; ++k) t -= li[k] * lj[k];
      L[i * n + j] = t * inv;
    }
  }
  std::vector<double> Lt(n * n, 0.0);   // L^T, so that the back substitution reads rows too
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = 0; k <= i; ++k) Lt[k * n + i] = L[i * n + k];
  // column c of the inverse: L y = e_c (y_i = 0 for i < c), L^T x = y
  auto columns = [&](int64_t c0, int64_t c1) {
    std::vector<double> y(n), x(n);
    for (int64_t c = c0; c < c1; ++c) {
      for (int64_t i = 0; i < c; ++i) y[i] = 0.0;
      for (int64_t i = c; i < n; ++i) {
        const double* li = &L[i * n];
        double t = (i == c) ? 1.0 : 0.0;
        for (int64_t k = c; k < i; ++k) t -= li[k] * y[k];
        y[i] = t / li[i];
      }
      for (int64_t i = n - 1; i >= 0; --i) {
        const double* ri = &Lt[i * n];
        double t = y[i];
        for (int64_t k = i + 1; k < n; ++k) t -= ri[k] * x[k];
        x[i] = t / ri[i];
      }
      for (int64_t i = 0; i < n; ++i) M[i * n + c] = x[i];
    }
  };
  parallel_chunks(n, n < 256 ? 1 : host_threads(), [&](int, int64_t c0, int64_t c1) { columns(c0, c1); });
  // symmetrise against roundoff
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = i + 1; j < n; ++j) {
      const double a = 0.5 * (M[i * n + j] + M[j * n + i]);
      M[i * n + j] = M[j * n + i] = a;
    }
}

// `off0` (optional): world+1 offsets partitioning the rows of A into contiguous ranges, one
// per rank.  Every coarse level is then partitioned too: an aggregate belongs to the rank
// that owns its first (lowest-numbered) fine node, and aggregates are renumbered so that
// each rank's aggregates are contiguous (H.off).  With one rank nothing is renumbered.
inline AmgHierarchy build_amg(HostCsr<double> A, double theta, int64_t max_coarse,
                              int max_levels, const std::vector<int64_t>* off0 = nullptr) {
  AmgHierarchy H;
  std::vector<double> B(A.rows, 1.0);
  std::vector<int64_t> cur_off = off0 != nullptr ? *off0 : std::vector<int64_t>{0, A.rows};
  if (cur_off.size() < 2 || cur_off.front() != 0 || cur_off.back() != A.rows)
    throw std::invalid_argument("bad partition offsets");
  const int world = static_cast<int>(cur_off.size()) - 1;
  while (true) {
    H.off.push_back(cur_off);
    AmgLevel lv;
    lv.A = std::move(A);
    const int64_t n = lv.A.rows;
    std::vector<double> d = diagonal(lv.A);
    lv.dinv.resize(n);
    for (int64_t i = 0; i < n; ++i) {
      if (!(d[i] > 0)) throw std::runtime_error("non-positive diagonal in the mu operator");
      lv.dinv[i] = 1.0 / d[i];
    }
    lv.rho = rho_dinv_a(lv.A, d);
    lv.B = B;
    const bool last = n <= max_coarse || static_cast<int>(H.levels.size()) + 1 >= max_levels;
    if (last) { H.levels.push_back(std::move(lv)); break; }
    std::vector<int32_t> agg;
    const int64_t nagg = aggregate(lv.A, d, theta, agg);
    if (nagg >= n) { H.levels.push_back(std::move(lv)); break; }  // cannot coarsen
    if (world > 1) {
      // group the aggregates by owning rank (stable: keeps the creation order inside a rank)
      std::vector<int> own(nagg, -1);
      for (int64_t i = 0; i < n; ++i)
        if (own[agg[i]] < 0) own[agg[i]] = owner_of(cur_off, i);  // i ascends: first node wins
      std::vector<int64_t> next_off(world + 1, 0);
      for (int64_t c = 0; c < nagg; ++c) next_off[own[c] + 1]++;
      for (int r = 0; r < world; ++r) next_off[r + 1] += next_off[r];
      std::vector<int64_t> fill(next_off.begin(), next_off.end() - 1);
      std::vector<int32_t> relabel(nagg);
      for (int64_t c = 0; c < nagg; ++c) relabel[c] = static_cast<int32_t>(fill[own[c]]++);
      for (int64_t i = 0; i < n; ++i) agg[i] = relabel[agg[i]];
      cur_off = next_off;
    } else {
      cur_off = {0, nagg};
    }
    // tentative prolongator T (one entry per row), normalised so that T B_c = B
    std::vector<double> nrm(nagg, 0.0);
    for (int64_t i = 0; i < n; ++i) nrm[agg[i]] += B[i] * B[i];
    for (auto& v : nrm) v = std::sqrt(v);
    HostCsr<double> T;
    T.rows = n; T.cols = nagg;
    T.ptr.resize(n + 1); T.idx.resize(n); T.val.resize(n);
    for (int64_t i = 0; i < n; ++i) { T.ptr[i] = static_cast<int32_t>(i); T.idx[i] = agg[i]; T.val[i] = B[i] / nrm[agg[i]]; }
    T.ptr[n] = static_cast<int32_t>(n);
    // P = (I - omega D^-1 A) T
    const double omega = (4.0 / 3.0) / lv.rho;
    HostCsr<double> P = spgemm(lv.A, T);
    for (int64_t i = 0; i < n; ++i)
      for (int32_t k = P.ptr[i]; k < P.ptr[i + 1]; ++k) {
        P.val[k] *= -omega * lv.dinv[i];
        if (P.idx[k] == agg[i]) P.val[k] += T.val[i];
      }
    HostCsr<double> R = transpose(P);
    HostCsr<double> Ac = spgemm(R, spgemm(lv.A, P));
    lv.P = std::move(P);
    lv.R = std::move(R);
    H.levels.push_back(std::move(lv));
    A = std::move(Ac);
    B = nrm;
  }
  // coarsest level: dense inverse of A_c + g B B^T (the B-component of a compatible
  // right-hand side is zero, so this acts as the pseudo-inverse)
  const AmgLevel& c = H.levels.back();
  const int64_t nc = c.A.rows;
  H.nc = nc;
  H.coarse_inv.assign(nc * nc, 0.0);
  double tr = 0, bb = 0;
  for (int64_t i = 0; i < nc; ++i) {
    for (int32_t k = c.A.ptr[i]; k < c.A.ptr[i + 1]; ++k) {
      H.coarse_inv[i * nc + c.A.idx[k]] = c.A.val[k];
      if (c.A.idx[k] == i) tr += c.A.val[k];
    }
    bb += c.B[i] * c.B[i];
  }
  const double g = tr / static_cast<double>(nc) / bb;
  for (int64_t i = 0; i < nc; ++i)
    for (int64_t j = 0; j < nc; ++j) H.coarse_inv[i * nc + j] += g * c.B[i] * c.B[j];
  dense_spd_inverse(H.coarse_inv, nc);
  return H;
}

}  // namespace tdgl
