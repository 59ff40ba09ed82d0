// Host-side construction of the finite-volume (dual) mesh arrays from a triangulation — the
// input contract of the hot path (no CUDA here).
//
// The reference builds these arrays with Python loops over edges and sites
// (tdgl/finite_volume/util.py:15-28 get_edges, :100-124 circumcentres, :59-97 dual edge
// lengths; edge_mesh.py:54-92; 108 s at 1M sites).  Here the triangle edges are sorted on a
// packed 64-bit key by several threads and everything per edge / per triangle is one parallel
// pass.  Every floating-point expression is evaluated in the order the vectorised NumPy builder
// (py-tdgl_b200/mesh.py::_from_triangulation_numpy) evaluates it, so both give the same bits;
// the Voronoi areas of the O(sqrt N) boundary sites follow the reference's convex-hull
// convention and stay in Python (mesh.py::_voronoi_areas).
#pragma once

#include <utility>

#include "host_csr.h"

namespace tdgl {

// Sorts v with the host threads: chunk sorts + pairwise merges.  With a strict total order on
// distinct elements the result is the one std::sort gives.
template <typename T>
inline void parallel_sort(std::vector<T>& v, int64_t min_per_chunk = 1 << 17) {
  const int64_t n = static_cast<int64_t>(v.size());
  const int chunks = chunks_for(n, min_per_chunk);
  if (chunks <= 1) { std::sort(v.begin(), v.end()); return; }
  std::vector<int64_t> cut(chunks + 1);
  for (int c = 0; c <= chunks; ++c) cut[c] = n * c / chunks;
  parallel_chunks(n, chunks, [&](int, int64_t lo, int64_t hi) { std::sort(v.begin() + lo, v.begin() + hi); });
  for (int width = 1; width < chunks; width *= 2) {
    const int pairs = (chunks + 2 * width - 1) / (2 * width);
    parallel_chunks(pairs, pairs, [&](int, int64_t p0, int64_t p1) {
      for (int64_t p = p0; p < p1; ++p) {
        const int a = static_cast<int>(p) * 2 * width, m = std::min(a + width, chunks), b = std::min(a + 2 * width, chunks);
        if (m < b) std::inplace_merge(v.begin() + cut[a], v.begin() + cut[m], v.begin() + cut[b]);
      }
    });
  }
}

// Outputs have room for 3 * n_tri edges; returns the number of unique edges E.
//   edges[E,2]        each row (lo, hi), rows lexicographically ascending (util.py:25-28)
//   is_boundary[E]    1 if the edge belongs to exactly one triangle
//   dual[T,2]         circumcentres
//   centers, directions [E,2], lengths[E], dual_lengths[E]
//   areas[n]          sum over incident edges of length * dual_length / 4 (the Voronoi cell
//                     area of an interior site; boundary sites are redone by the caller)
inline int64_t build_dual_mesh(int64_t n, int64_t T, const double* sites, const int64_t* tri,
                               int64_t* edges, uint8_t* is_boundary, double* dual,
                               double* centers, double* directions, double* lengths,
                               double* dual_lengths, double* areas) {
  if (n < 3 || T < 1) throw std::invalid_argument("a mesh needs at least one triangle");
  if (n > 0x7FFFFFFFll || 3 * T > 0xFFFFFFF0ll) throw std::invalid_argument("mesh too large");
  // (key, position in [pairs (0,1) of all triangles | pairs (1,2) | pairs (2,0)]): the order
  // a stable sort of the concatenated pair list gives
  std::vector<std::pair<uint64_t, uint32_t>> key(3 * T);
  const int tchunks = chunks_for(T);
  parallel_chunks(T, tchunks, [&](int, int64_t lo, int64_t hi) {
    for (int64_t t = lo; t < hi; ++t) {
      const int64_t v[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
      for (int s = 0; s < 3; ++s) {
        const int64_t a = v[s], b = v[(s + 1) % 3];
        if (a < 0 || a >= n || b < 0 || b >= n) throw std::invalid_argument("element index out of range");
        const uint64_t l = static_cast<uint64_t>(std::min(a, b)), h = static_cast<uint64_t>(std::max(a, b));
        key[s * T + t] = {l * static_cast<uint64_t>(n) + h, static_cast<uint32_t>(s * T + t)};
      }
    }
  });
  parallel_sort(key);
  // unique edges; first[e] = position of the edge's first entry in `key`
  std::vector<int64_t> first;
  first.reserve(3 * T / 2 + 16);
  for (int64_t k = 0; k < 3 * T; ++k)
    if (k == 0 || key[k].first != key[k - 1].first) first.push_back(k);
  const int64_t E = static_cast<int64_t>(first.size());
  first.push_back(3 * T);
  // circumcentres (mesh.py::circumcenters, same expression order)
  parallel_chunks(T, tchunks, [&](int, int64_t lo, int64_t hi) {
    for (int64_t t = lo; t < hi; ++t) {
      const double ax = sites[2 * tri[3 * t]], ay = sites[2 * tri[3 * t] + 1];
      const double bx = sites[2 * tri[3 * t + 1]] - ax, by = sites[2 * tri[3 * t + 1] + 1] - ay;
      const double cx = sites[2 * tri[3 * t + 2]] - ax, cy = sites[2 * tri[3 * t + 2] + 1] - ay;
      const double d = 2 * bx * cy - 2 * by * cx;
      const double b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
      const double ux = (cy * b2 - by * c2) / d, uy = (bx * c2 - cx * b2) / d;
      dual[2 * t] = ux + ax;
      dual[2 * t + 1] = uy + ay;
    }
  });
  parallel_chunks(E, chunks_for(E), [&](int, int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; ++e) {
      const int64_t k = first[e], cnt = first[e + 1] - k;
      const int64_t i0 = static_cast<int64_t>(key[k].first / static_cast<uint64_t>(n));
      const int64_t i1 = static_cast<int64_t>(key[k].first % static_cast<uint64_t>(n));
      edges[2 * e] = i0;
      edges[2 * e + 1] = i1;
      is_boundary[e] = cnt == 1;
      const double x0 = sites[2 * i0], y0 = sites[2 * i0 + 1], x1 = sites[2 * i1], y1 = sites[2 * i1 + 1];
      const double mx = (x0 + x1) / 2.0, my = (y0 + y1) / 2.0;
      const double dx = x1 - x0, dy = y1 - y0;
      centers[2 * e] = mx; centers[2 * e + 1] = my;
      directions[2 * e] = dx; directions[2 * e + 1] = dy;
      lengths[e] = std::sqrt(dx * dx + dy * dy);
      const int64_t t0 = key[k].second % T;
      double ex, ey;
      if (cnt == 1) {
        ex = dual[2 * t0] - mx; ey = dual[2 * t0 + 1] - my;
      } else {
        const int64_t t1 = key[k + 1].second % T;
        ex = dual[2 * t0] - dual[2 * t1]; ey = dual[2 * t0 + 1] - dual[2 * t1 + 1];
      }
      dual_lengths[e] = std::sqrt(ex * ex + ey * ey);
    }
  });
  // areas = bincount(e0, q) + bincount(e1, q) with q = 0.25 * length * dual_length: two
  // accumulators per site, each filled in ascending edge order
  std::vector<double> second(n, 0.0);
  for (int64_t i = 0; i < n; ++i) areas[i] = 0.0;
  for (int64_t e = 0; e < E; ++e) {
    const double q = 0.25 * lengths[e] * dual_lengths[e];
    areas[edges[2 * e]] += q;
    second[edges[2 * e + 1]] += q;
  }
  for (int64_t i = 0; i < n; ++i) areas[i] += second[i];
  return E;
}

}  // namespace tdgl
