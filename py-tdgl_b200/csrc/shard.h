// Domain decomposition of the mesh operators and of the AMG hierarchy (host side, no CUDA).
//
// The reference has no multi-GPU path; this is the B200 build's own.  Sites
// are numbered along a Z-order curve, rank r owns the contiguous range off[0][r]..off[0][r+1]
// of that numbering (a coordinate partitioner: compact subdomains, no METIS needed), and
// every coarse level is partitioned by build_amg() along with it.  A rank stores only its
// rows; a local vector is laid out [owned | halo], the halo being the sorted list of
// non-owned columns its rows reference.  Because ownership ranges are contiguous, a sorted
// halo is automatically grouped by owning rank, and what rank p sends to rank q is simply
// the part of q's halo that falls into p's range — both sides derive the same lists from
// the same global plan, nothing has to be negotiated at run time.
#pragma once

#include "amg_setup.h"

namespace tdgl {

constexpr int kMaxWorld = 8;

struct ShardPlan {
  int world = 1;
  int levels = 0;
  // First replicated level.  Levels l < rep are partitioned (a rank computes its owned rows,
  // vectors are [owned | halo]); on level rep every rank holds the whole vector, laid out
  // [owned | all others], computes ALL rows redundantly and only the restriction into it is
  // partitioned (followed by an all-gather); levels l > rep are plain replicas in the global
  // numbering.  The coarse levels are tiny (<= kReplicateBelow rows): recomputing them on
  // every GPU costs nothing and removes four exchanges per level from every V-cycle.
  int rep = 0;
  std::vector<std::vector<int64_t>> off;                 // [level][world + 1]
  std::vector<std::vector<std::vector<int32_t>>> halo;   // [level][rank] sorted global ids

  int64_t rows(int l) const { return off[l][world]; }
  int64_t owned(int l, int r) const { return l > rep ? rows(l) : off[l][r + 1] - off[l][r]; }
  int64_t local_size(int l, int r) const {
    return l > rep ? rows(l) : off[l][r + 1] - off[l][r] + static_cast<int64_t>(halo[l][r].size());
  }
  // rows rank r computes on level l
  int64_t compute_rows(int l, int r) const { return l < rep ? owned(l, r) : local_size(l, r); }
  // position of global id g in rank r's local numbering of level l (-1: not present)
  int64_t local_index(int l, int r, int64_t g) const {
    if (l > rep) return g;
    if (g >= off[l][r] && g < off[l][r + 1]) return g - off[l][r];
    const auto& h = halo[l][r];
    auto it = std::lower_bound(h.begin(), h.end(), static_cast<int32_t>(g));
    if (it == h.end() || *it != g) return -1;
    return (off[l][r + 1] - off[l][r]) + (it - h.begin());
  }
  int64_t global_index(int l, int r, int64_t k) const {
    if (l > rep) return k;
    const int64_t n = off[l][r + 1] - off[l][r];
    return k < n ? off[l][r] + k : halo[l][r][k - n];
  }
};

constexpr int64_t kReplicateBelow = 32768;

// Equal split of n rows over `world` ranks.
inline std::vector<int64_t> equal_offsets(int64_t n, int world) {
  std::vector<int64_t> off(world + 1);
  for (int r = 0; r <= world; ++r) off[r] = n * r / world;
  return off;
}

// ---- site numbering ---------------------------------------------------------------------------
// Z-order (Morton) numbering of the sites: a window of consecutive CSR rows references mostly
// itself, and a contiguous range of the numbering is a compact subdomain.
inline std::vector<int> morton_permutation(const double* xy, int64_t n) {
  double xmin = xy[0], xmax = xy[0], ymin = xy[1], ymax = xy[1];
  for (int64_t i = 0; i < n; ++i) {
    xmin = std::min(xmin, xy[2 * i]); xmax = std::max(xmax, xy[2 * i]);
    ymin = std::min(ymin, xy[2 * i + 1]); ymax = std::max(ymax, xy[2 * i + 1]);
  }
  const double span = std::max(std::max(xmax - xmin, ymax - ymin), 1e-300);
  auto spread = [](uint64_t v) {
    v &= 0xFFFFFFFFull;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
  };
  std::vector<std::pair<uint64_t, int>> key(n);
  const double scale = static_cast<double>((1u << 20) - 1) / span;
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t qx = static_cast<uint64_t>((xy[2 * i] - xmin) * scale);
    const uint64_t qy = static_cast<uint64_t>((xy[2 * i + 1] - ymin) * scale);
    key[i] = {spread(qx) | (spread(qy) << 1), static_cast<int>(i)};
  }
  // (keys are unique pairs: chunk sorts + merges give the same order as one sort)
  const int chunks = chunks_for(n, 1 << 17);
  std::vector<int64_t> cut(chunks + 1);
  for (int c = 0; c <= chunks; ++c) cut[c] = n * c / chunks;
  parallel_chunks(n, chunks, [&](int, int64_t lo, int64_t hi) { std::sort(key.begin() + lo, key.begin() + hi); });
  for (int width = 1; width < chunks; width *= 2) {
    const int pairs = (chunks + 2 * width - 1) / (2 * width);
    parallel_chunks(pairs, pairs, [&](int, int64_t p0, int64_t p1) {
      for (int64_t p = p0; p < p1; ++p) {
        const int a = static_cast<int>(p) * 2 * width, m = std::min(a + width, chunks), b = std::min(a + 2 * width, chunks);
        if (m < b) std::inplace_merge(key.begin() + cut[a], key.begin() + cut[m], key.begin() + cut[b]);
      }
    });
  }
  std::vector<int> perm(n);
  for (int64_t i = 0; i < n; ++i) perm[i] = key[i].second;
  return perm;
}

// Sharded numbering: Z-order, cut into `world` equal ranges, and inside every range the rows
// near the cut (within `depth` edges of a row of another range) moved to the front.  A rank's
// kernels then compute — and store into the neighbours' mailboxes — the rows the neighbours
// wait for FIRST, and the CTAs that read halo columns (the same rows) run while the
// neighbour's matching stores, issued at the start of its previous kernel, have long landed:
// the exchange has a whole kernel of slack instead of sitting on the critical path.
inline std::vector<int> shard_permutation(const double* xy, int64_t n, int64_t n_edges,
                                          const int64_t* edges, int world, int depth = 2) {
  std::vector<int> perm = morton_permutation(xy, n);
  if (world <= 1) return perm;
  std::vector<int> inv(n);
  for (int64_t i = 0; i < n; ++i) inv[perm[i]] = static_cast<int>(i);
  std::vector<int64_t> off(world + 1);
  for (int r = 0; r <= world; ++r) off[r] = n * r / world;
  auto owner = [&](int pos) {
    return static_cast<int>(std::upper_bound(off.begin(), off.end(), static_cast<int64_t>(pos)) - off.begin()) - 1;
  };
  // level[pos]: 0 = not near a cut, k = reached in the k-th sweep
  std::vector<unsigned char> near(n, 0);
  for (int64_t e = 0; e < 2 * n_edges; ++e)
    if (edges[e] < 0 || edges[e] >= n) throw std::invalid_argument("edge index out of range");
  for (int64_t e = 0; e < n_edges; ++e) {
    const int a = inv[edges[2 * e]], b = inv[edges[2 * e + 1]];
    if (owner(a) != owner(b)) near[a] = near[b] = 1;
  }
  for (int d = 2; d <= depth; ++d) {
    std::vector<unsigned char> next(near);
    for (int64_t e = 0; e < n_edges; ++e) {
      const int a = inv[edges[2 * e]], b = inv[edges[2 * e + 1]];
      if (owner(a) != owner(b)) continue;
      if (near[a] && !near[b]) next[b] = static_cast<unsigned char>(d);
      if (near[b] && !near[a]) next[a] = static_cast<unsigned char>(d);
    }
    near.swap(next);
  }
  std::vector<int> out(n);
  for (int r = 0; r < world; ++r) {
    int64_t w = off[r];
    for (int64_t p = off[r]; p < off[r + 1]; ++p)
      if (near[p]) out[w++] = perm[p];
    for (int64_t p = off[r]; p < off[r + 1]; ++p)
      if (!near[p]) out[w++] = perm[p];
  }
  return out;
}

// Adds to halo[r] the non-owned columns of the rows row_off[r]..row_off[r+1] of G, whose
// columns live on a level partitioned by col_off.
inline void collect_halo(const HostCsr<double>& G, const std::vector<int64_t>& row_off,
                         const std::vector<int64_t>& col_off,
                         std::vector<std::vector<int32_t>>& halo) {
  const int world = static_cast<int>(row_off.size()) - 1;
  for (int r = 0; r < world; ++r) {
    const int64_t c0 = col_off[r], c1 = col_off[r + 1];
    for (int64_t i = row_off[r]; i < row_off[r + 1]; ++i)
      for (int32_t k = G.ptr[i]; k < G.ptr[i + 1]; ++k) {
        const int32_t j = G.idx[k];
        if (j < c0 || j >= c1) halo[r].push_back(j);
      }
  }
}

inline ShardPlan make_plan(const AmgHierarchy& H, int64_t replicate_below = kReplicateBelow) {
  ShardPlan p;
  p.levels = static_cast<int>(H.levels.size());
  p.off = H.off;
  p.world = static_cast<int>(H.off[0].size()) - 1;
  p.halo.assign(p.levels, std::vector<std::vector<int32_t>>(p.world));
  p.rep = 0;
  if (p.world == 1) return p;
  if (p.world > kMaxWorld) throw std::invalid_argument("at most 8 ranks");
  if (p.levels < 2) throw std::invalid_argument("mesh too small to shard (single-level hierarchy)");
  p.rep = p.levels - 1;
  for (int l = 1; l < p.levels; ++l)
    if (H.levels[l].A.rows <= replicate_below) { p.rep = l; break; }
  for (int l = 0; l <= p.rep; ++l) {
    const AmgLevel& lv = H.levels[l];
    if (l == p.rep) {
      // gathered level: every other rank's rows are "halo"
      const int64_t n = lv.A.rows;
      for (int r = 0; r < p.world; ++r)
        for (int64_t g = 0; g < n; ++g)
          if (g < p.off[l][r] || g >= p.off[l][r + 1]) p.halo[l][r].push_back(static_cast<int32_t>(g));
      continue;
    }
    collect_halo(lv.A, p.off[l], p.off[l], p.halo[l]);       // smoothers, SpMV
    collect_halo(lv.R, p.off[l + 1], p.off[l], p.halo[l]);   // restriction reads level-l residuals
    if (l > 0) collect_halo(H.levels[l - 1].P, p.off[l - 1], p.off[l], p.halo[l]);  // prolongation
    for (int r = 0; r < p.world; ++r) {
      auto& h = p.halo[l][r];
      std::sort(h.begin(), h.end());
      h.erase(std::unique(h.begin(), h.end()), h.end());
    }
  }
  return p;
}

// The global rows rank `rank` computes on level l, in its local order.
inline std::vector<int64_t> compute_row_list(const ShardPlan& plan, int l, int rank) {
  std::vector<int64_t> rows(plan.compute_rows(l, rank));
  for (size_t k = 0; k < rows.size(); ++k) rows[k] = plan.global_index(l, rank, static_cast<int64_t>(k));
  return rows;
}

// The given global rows of G (in that order), columns renumbered into rank `rank`'s local
// layout of the column level `cl`.
inline HostCsr<double> extract_rows(const HostCsr<double>& G, const std::vector<int64_t>& rows,
                                    const ShardPlan& plan, int cl, int rank) {
  HostCsr<double> L;
  L.rows = static_cast<int64_t>(rows.size());
  L.cols = plan.local_size(cl, rank);
  L.ptr.assign(L.rows + 1, 0);
  for (int64_t i = 0; i < L.rows; ++i) L.ptr[i + 1] = L.ptr[i] + (G.ptr[rows[i] + 1] - G.ptr[rows[i]]);
  L.idx.resize(L.ptr[L.rows]);
  L.val.resize(L.ptr[L.rows]);
  parallel_chunks(L.rows, chunks_for(L.rows), [&](int, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      int32_t d = L.ptr[i];
      for (int32_t k = G.ptr[rows[i]]; k < G.ptr[rows[i] + 1]; ++k, ++d) {
        const int64_t li = plan.local_index(cl, rank, G.idx[k]);
        if (li < 0) throw std::runtime_error("halo plan misses a column");
        L.idx[d] = static_cast<int32_t>(li);
        L.val[d] = G.val[k];
      }
    }
  });
  return L;
}

// Rows r0..r1 of G with columns renumbered into rank `rank`'s local layout of the column
// level `cl`.  Column order inside a row is kept (the kernels do not need sorted rows).
inline HostCsr<double> extract_local(const HostCsr<double>& G, int64_t r0, int64_t r1,
                                     const ShardPlan& plan, int cl, int rank) {
  std::vector<int64_t> rows(r1 - r0);
  for (int64_t i = r0; i < r1; ++i) rows[i - r0] = i;
  return extract_rows(G, rows, plan, cl, rank);
}

// What `rank` sends to `peer` on level l: local (owned) indices, in the order in which they
// appear in the peer's halo, and the position of that block inside the peer's halo.
struct SendBlock {
  int peer = -1;
  std::vector<int32_t> idx;   // local owned indices on `rank`
  int64_t dst_pos = 0;        // first halo slot on the peer (relative to the start of its halo)
};

inline std::vector<SendBlock> send_blocks(const ShardPlan& plan, int l, int rank) {
  std::vector<SendBlock> out;
  if (l > plan.rep) return out;
  const int64_t g0 = plan.off[l][rank], g1 = plan.off[l][rank + 1];
  for (int q = 0; q < plan.world; ++q) {
    if (q == rank) continue;
    const auto& h = plan.halo[l][q];
    auto b = std::lower_bound(h.begin(), h.end(), static_cast<int32_t>(g0));
    auto e = std::lower_bound(h.begin(), h.end(), static_cast<int32_t>(g1));
    if (b == e) continue;
    SendBlock s;
    s.peer = q;
    s.dst_pos = b - h.begin();
    s.idx.reserve(e - b);
    for (auto it = b; it != e; ++it) s.idx.push_back(static_cast<int32_t>(*it - g0));
    out.push_back(std::move(s));
  }
  return out;
}

// Ranks that send to `rank` on level l (the flags it waits for).
inline std::vector<int> recv_peers(const ShardPlan& plan, int l, int rank) {
  std::vector<int> out;
  if (l > plan.rep) return out;
  const auto& h = plan.halo[l][rank];
  for (int q = 0; q < plan.world; ++q) {
    if (q == rank) continue;
    auto b = std::lower_bound(h.begin(), h.end(), static_cast<int32_t>(plan.off[l][q]));
    auto e = std::lower_bound(h.begin(), h.end(), static_cast<int32_t>(plan.off[l][q + 1]));
    if (b != e) out.push_back(q);
  }
  return out;
}

// ---- arena layout -------------------------------------------------------------------------
// Everything a peer writes into lives in ONE device allocation per rank (the arena), so that
// one CUDA IPC handle per rank makes it addressable.  The layout is a pure function of the
// plan, hence every rank can compute every peer's offsets.  Units: 8-byte words.
//   [0, kArenaHeader)   all-reduce mailbox: [2 parities][8 ranks][4 values][2 words]
//   vectors             psi0, psi1 (complex: 2 words per entry), mu, cg_r, cg_p, then per
//                       level x, r, b, y; each start aligned to 32 words
//   halo mailboxes      one per CHANNEL (= exchanged vector): [2 parities][halo entries of
//                       the vector's level][2 words per double]
// Mailbox words use the "low latency" format of comm.cuh: 32 bits of payload + a 32-bit tag
// in one 8-byte store, so that a value and its arrival flag are one atomic write.  A
// consumer kernel reads halo entries straight out of the mailbox (polling the tag), so there
// is no separate exchange step: the producer kernel stores its boundary rows into the
// peers' mailboxes as it computes them.
constexpr int64_t kArenaRedBox = 0, kArenaHeader = 256;
enum : int { kVecPsi0 = 0, kVecPsi1, kVecMu, kVecCgR, kVecCgP, kVecLevel0 };
inline int vec_id(int level, int which /*0 x, 1 r, 2 b, 3 y*/) { return kVecLevel0 + 4 * level + which; }
// channels are numbered like the vectors they carry
inline int ch_words(int ch) { return ch <= kVecPsi1 ? 4 : 2; }
inline int ch_level(int ch) { return ch < kVecLevel0 ? 0 : (ch - kVecLevel0) / 4; }

struct ArenaLayout {
  std::vector<int64_t> off;      // per vector id, in words from the arena base
  std::vector<int64_t> box_off;  // per channel: start of the mailbox (parity 0); 0: none
  std::vector<int64_t> box_cap;  // per channel: words per parity
  int64_t total = 0;             // words
};

inline ArenaLayout arena_layout(const ShardPlan& plan, int rank) {
  ArenaLayout a;
  int64_t pos = kArenaHeader;
  auto add = [&](int64_t words) {
    const int64_t at = pos;
    pos += (words + 31) / 32 * 32;
    return at;
  };
  const int64_t nx0 = plan.local_size(0, rank);
  a.off.push_back(add(2 * nx0)); a.off.push_back(add(2 * nx0));
  a.off.push_back(add(nx0)); a.off.push_back(add(nx0)); a.off.push_back(add(nx0));
  for (int l = 0; l < plan.levels; ++l) {
    const int64_t nx = plan.local_size(l, rank);
    for (int w = 0; w < 4; ++w) a.off.push_back(add(nx));
  }
  const int nch = kVecLevel0 + 4 * plan.levels;
  a.box_off.assign(nch, 0);
  a.box_cap.assign(nch, 0);
  if (plan.world > 1)
    for (int ch = 0; ch < nch; ++ch) {
      const int l = ch_level(ch);
      if (l > plan.rep) continue;
      a.box_cap[ch] = std::max<int64_t>(static_cast<int64_t>(plan.halo[l][rank].size()) * ch_words(ch), 2);
      a.box_off[ch] = add(2 * a.box_cap[ch]);
    }
  a.total = pos;
  return a;
}

}  // namespace tdgl
