// Device-side exchange steps of the sharded engine (sm_100a, NVLink peer memory).
//
// Every rank maps every peer's arena (CUDA IPC across processes, or plainly the same
// address space when several shards share one process) and the exchange steps are plain
// device code on those mappings — no NCCL call, no host, no staging copy: a time step needs
// a few hundred exchanges of <= 100 KB each, which is a latency problem, not a bandwidth one.
//
// Mailbox format ("low latency", the idea of NCCL's LL protocol): every 8-byte word a peer
// writes carries 32 bits of payload and a 32-bit sequence tag, stored with ONE 64-bit store
// (single-copy atomic).  The receiver polls the word until the tag equals the sequence
// number it expects: data and arrival flag are the same write, so there is no
// fence + flag round trip — the cost of an exchange is one NVLink traversal.
//
//   halo exchange  there is no exchange kernel.  A kernel that produces a vector stores
//                  its boundary rows into the neighbours' mailboxes as it computes them
//                  (push_row); a kernel that gathers from a vector reads halo columns out of
//                  its own mailbox (halo_get), polling only for the entries it needs — CTAs
//                  that touch no halo column never wait, so the NVLink latency overlaps
//                  the interior of the subdomain.
//   all-reduce     comm_allreduce, called by ONE warp (in the last block of a reducing kernel
//                  or in a controller kernel): lane q stores the partial into rank q's
//                  mailbox and polls this rank's mailbox for rank q's partial; the partials
//                  are combined in rank order, so every rank gets the same bits and takes
//                  the same loop decisions (dt retry, CG continue, stop).
//
// Every poll is bounded by a wall-clock timeout that raises Ctl::status = 3 (reported as
// TDGL_E_CUDA) instead of hanging the GPU.
#pragma once

// (included by kernels.cuh right after the definition of Ctl)
#include "shard.h"

namespace tdgl {

constexpr int kMaxLevels = 24;

struct Comm {
  int rank = 0, world = 1;
  unsigned int rseq = 0;               // all-reduce sequence number (device mutated)
  unsigned long long* peer[kMaxWorld] = {};  // arena base of every rank as mapped here
};

constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;  // 4 s

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_word(unsigned long long* p, unsigned int payload,
                                        unsigned int tag) {
  const unsigned long long w = (static_cast<unsigned long long>(tag) << 32) | payload;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_f64(unsigned long long* p, double v, unsigned int tag) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  st_word(p, static_cast<unsigned int>(b), tag);
  st_word(p + 1, static_cast<unsigned int>(b >> 32), tag);
}

// Poll one mailbox word until its tag equals `tag`; returns the payload.  On timeout (or when
// another failure was raised meanwhile) sets *ok = false.
__device__ __forceinline__ unsigned int poll_word(Ctl* ctl, const unsigned long long* p,
                                                  unsigned int tag, bool* ok) {
  unsigned long long w = ld_word(p);
  if (static_cast<unsigned int>(w >> 32) == tag) return static_cast<unsigned int>(w);
  const unsigned long long t0 = global_ns();
  unsigned int n = 0;
  while (true) {
    w = ld_word(p);
    if (static_cast<unsigned int>(w >> 32) == tag) return static_cast<unsigned int>(w);
    if ((++n & 255u) == 0) {
      if (global_ns() - t0 > kSpinTimeoutNs ||
          *reinterpret_cast<volatile int*>(&ctl->status) != 0) {
        if (atomicCAS(&ctl->status, 0, 3) == 0) {
          ctl->fail_tag = tag;
          ctl->fail_addr = reinterpret_cast<unsigned long long>(p);
        }
        *ok = false;
        return 0u;
      }
    }
  }
}
// A double travels as two tagged words (low half, high half) in one aligned 16-byte pair:
// fetch both with one vector load (each word is validated by its own tag, so the pair needs
// no atomicity of its own) and re-poll until both tags match.
__device__ __forceinline__ void ld_pair(const unsigned long long* p, unsigned long long* lo,
                                        unsigned long long* hi) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(*lo), "=l"(*hi) : "l"(p) : "memory");
}
__device__ __forceinline__ bool pair_ready(unsigned long long lo, unsigned long long hi,
                                           unsigned int tag) {
  return static_cast<unsigned int>(lo >> 32) == tag && static_cast<unsigned int>(hi >> 32) == tag;
}
__device__ __forceinline__ double pair_value(unsigned long long lo, unsigned long long hi) {
  return __longlong_as_double(static_cast<long long>((hi << 32) | (lo & 0xffffffffull)));
}
__device__ __forceinline__ double poll_f64(Ctl* ctl, const unsigned long long* p,
                                           unsigned int tag, bool* ok) {
  unsigned long long lo, hi;
  ld_pair(p, &lo, &hi);
  if (pair_ready(lo, hi, tag)) return pair_value(lo, hi);
  const unsigned long long t0 = global_ns();
  unsigned int n = 0;
  while (true) {
    ld_pair(p, &lo, &hi);
    if (pair_ready(lo, hi, tag)) return pair_value(lo, hi);
    if ((++n & 255u) == 0) {
      if (global_ns() - t0 > kSpinTimeoutNs ||
          *reinterpret_cast<volatile int*>(&ctl->status) != 0) {
        if (atomicCAS(&ctl->status, 0, 3) == 0) {
          ctl->fail_tag = tag;
          ctl->fail_addr = reinterpret_cast<unsigned long long>(p);
        }
        *ok = false;
        return 0.0;
      }
    }
  }
}

// v[0..n) <- reduction over all ranks (sum in rank order, or max), n <= 4.  Must be called by
// all 32 lanes of ONE warp with the same arguments; the result is valid in every lane.
__device__ __forceinline__ void comm_allreduce(Ctl* ctl, Comm* c, double* v, int n, bool is_max) {
  const int lane = threadIdx.x & 31;
  const int me = c->rank, world = c->world;
  unsigned int s = c->rseq + 1u;
  if (s == 0u) s = 1u;  // tag 0 is the mailbox's initial state
  __syncwarp();
  const int par = static_cast<int>(s & 1u);
  bool ok = true;
  double got[4] = {0.0, 0.0, 0.0, 0.0};
  if (lane < world) {
    unsigned long long* box = c->peer[lane] + kArenaRedBox + ((par * kMaxWorld + me) * 4) * 2;
    for (int k = 0; k < n; ++k) st_f64(box + 2 * k, v[k], s);
    const unsigned long long* mine = c->peer[me] + kArenaRedBox + ((par * kMaxWorld + lane) * 4) * 2;
    for (int k = 0; k < n; ++k) got[k] = poll_f64(ctl, mine + 2 * k, s, &ok);
  }
  ok = __all_sync(0xffffffffu, ok);
  for (int k = 0; k < n; ++k) {
    double acc = __shfl_sync(0xffffffffu, got[k], 0);
    for (int q = 1; q < world; ++q) {
      const double x = __shfl_sync(0xffffffffu, got[k], q);
      acc = is_max ? fmax(acc, x) : acc + x;
    }
    if (ok) v[k] = acc;
  }
  __syncwarp();
  if (lane == 0) c->rseq = s;
  __syncwarp();   // (a second all-reduce by the same warp must see the new sequence number)
}

// ---- tags ------------------------------------------------------------------------------------
// A mailbox word is valid for a consumer when its tag equals the tag the consumer expects.
// Tags are derived from counters of the control block that are identical on all ranks and
// constant while the kernels that use them run, so nothing has to be signalled:
//   kTagIter      (solve_epoch, cg_it)      vectors of the V-cycle / CG of iteration cg_it
//   kTagIterNext  (solve_epoch, cg_it + 1)  the residual k_cg_fused leaves for the next one
//   kTagIter0     (solve_epoch, 0)          the residual the rhs kernel leaves for iteration 0
//   kTagPsiNew    psi_epoch + 1             psi written by the current attempt of the psi step
//   kTagPsiCur    psi_tag[cur]              the accepted psi
//   kTagMu        solve_epoch               mu of this step's solve (after the gauge shift)
//   kTagMuPrev    solve_epoch - 1           mu of the previous step (read by the rhs kernel)
//   kTagMuPrev2   solve_epoch - 2           mu of the step before (still in the other parity
//                                           buffer of the mailbox; the rhs kernel's mu_prev)
// Each channel's mailbox is double-buffered by tag parity; between two stores into the same
// buffer lies at least one all-reduce that every rank has passed, so its readers are done.
enum : int { kTagIter = 0, kTagIterNext, kTagIter0, kTagPsiNew, kTagPsiCur, kTagMu, kTagMuPrev,
             kTagMuPrev2 };

__device__ __forceinline__ unsigned int comm_tag(const Ctl* ctl, int mode) {
  const unsigned int s = static_cast<unsigned int>(ctl->solve_epoch) << 10;
  switch (mode) {
    case kTagIter: return s | (static_cast<unsigned int>(ctl->cg_it) & 1023u);
    case kTagIterNext: return s | ((static_cast<unsigned int>(ctl->cg_it) + 1u) & 1023u);
    case kTagIter0: return s;
    case kTagPsiNew: return static_cast<unsigned int>(ctl->psi_epoch) + 1u;
    case kTagPsiCur: return static_cast<unsigned int>(ctl->psi_tag[ctl->cur]);
    case kTagMu: return static_cast<unsigned int>(ctl->solve_epoch);
    case kTagMuPrev: return static_cast<unsigned int>(ctl->solve_epoch) - 1u;
    default: return static_cast<unsigned int>(ctl->solve_epoch) - 2u;
  }
}

// ---- producer side: boundary rows go straight into the peers' mailboxes ------------------------
struct PushArgs {
  const unsigned char* bnd = nullptr;  // per 32 rows: any of them is sent somewhere (null: off)
  const int* rptr = nullptr;           // per row: range of `ent`
  const int2* ent = nullptr;           // {peer rank, halo entry on that peer}
  long long box[2][kMaxWorld] = {};    // the channel's mailbox on every peer, per parity (words)
  int tag_mode = 0;
};

template <typename T>
__device__ __forceinline__ void push_row(const Comm* c, const PushArgs& p, unsigned int tag,
                                         int row, T v) {
  if (p.bnd == nullptr || !p.bnd[row >> 5]) return;   // uniform per warp
  constexpr int W = sizeof(T) / 4;
  const double* d = reinterpret_cast<const double*>(&v);
  const int par = static_cast<int>(tag & 1u);
  for (int k = __ldg(p.rptr + row), ke = __ldg(p.rptr + row + 1); k < ke; ++k) {
    const int2 e = __ldg(p.ent + k);
    unsigned long long* dst = c->peer[e.x] + p.box[par][e.x] + static_cast<long long>(e.y) * W;
#pragma unroll
    for (int w = 0; w < W / 2; ++w) st_f64(dst + 2 * w, d[w], tag);
  }
}

// ---- consumer side: halo columns are read out of this rank's mailbox ---------------------------
struct HaloArgs {
  int n_owned = 0x7fffffff;   // columns >= n_owned are halo entries (default: there are none)
  long long box[2] = {};      // this rank's mailbox of the channel, per parity (words)
  int tag_mode = 0;
};

struct HaloView {             // HaloArgs resolved for one kernel invocation
  int n_owned;
  const unsigned long long* box;
  unsigned int tag;
};
__device__ __forceinline__ HaloView halo_view(const Ctl* ctl, const Comm* c, const HaloArgs& h) {
  HaloView v;
  v.n_owned = h.n_owned;
  v.box = nullptr;
  v.tag = 0;
  if (c != nullptr && h.n_owned != 0x7fffffff) {
    v.tag = comm_tag(ctl, h.tag_mode);
    v.box = c->peer[c->rank] + h.box[v.tag & 1u];
  } else {
    v.n_owned = 0x7fffffff;
  }
  return v;
}
__device__ __forceinline__ double halo_get(Ctl* ctl, const HaloView& h, const double* __restrict__ x,
                                           int j) {
  if (j < h.n_owned) return __ldg(x + j);
  bool ok = true;
  return poll_f64(ctl, h.box + static_cast<long long>(j - h.n_owned) * 2, h.tag, &ok);
}
// (float arrays of the V-cycle: the mailbox carries the owner's float value widened to double)
__device__ __forceinline__ float halo_get(Ctl* ctl, const HaloView& h, const float* __restrict__ x,
                                          int j) {
  if (j < h.n_owned) return __ldg(x + j);
  bool ok = true;
  return static_cast<float>(poll_f64(ctl, h.box + static_cast<long long>(j - h.n_owned) * 2, h.tag, &ok));
}
__device__ __forceinline__ double2 halo_get(Ctl* ctl, const HaloView& h,
                                            const double2* __restrict__ x, int j) {
  if (j < h.n_owned) return __ldg(x + j);
  bool ok = true;
  const unsigned long long* p = h.box + static_cast<long long>(j - h.n_owned) * 4;
  double2 v;
  v.x = poll_f64(ctl, p, h.tag, &ok);
  v.y = poll_f64(ctl, p + 2, h.tag, &ok);
  return v;
}

// Copies a channel's mailbox into the halo slots of its vector (waiting for the entries):
// for the consumers that read plain arrays — the replicated AMG levels, whose right-hand side
// is all-gathered this way once per V-cycle, and the edge-current kernel at save steps.
template <typename T>
__global__ void __launch_bounds__(1024)
k_unpack(Ctl* ctl, Comm* c, HaloArgs h, int n_halo, T* __restrict__ vec) {
  griddep_enter();
  if (ctl->status != 0) return;
  const HaloView v = halo_view(ctl, c, h);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_halo; e += gridDim.x * blockDim.x)
    vec[h.n_owned + e] = halo_get(ctl, v, vec, h.n_owned + e);
}

// Fills a channel of this rank's OWN mailbox from a whole-mesh array (tdgl_set_state: every
// rank is handed the whole psi / mu, nothing has to travel).
template <typename T>
__global__ void __launch_bounds__(256)
k_fill_box(Comm* c, long long box0, long long box1, unsigned int tag, int n_owned, int n_halo,
           const T* __restrict__ vec) {
  constexpr int W = sizeof(T) / 4;
  unsigned long long* box = c->peer[c->rank] + ((tag & 1u) ? box1 : box0);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_halo; e += gridDim.x * blockDim.x) {
    const T v = vec[n_owned + e];
    const double* d = reinterpret_cast<const double*>(&v);
#pragma unroll
    for (int w = 0; w < W / 2; ++w) st_f64(box + static_cast<long long>(e) * W + 2 * w, d[w], tag);
  }
}

}  // namespace tdgl
