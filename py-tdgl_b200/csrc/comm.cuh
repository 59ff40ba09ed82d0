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
//   halo exchange  k_halo_exchange (one CTA): scatters the rank's boundary entries into the
//                  neighbours' mailboxes, then polls its own mailbox and unpacks it into the
//                  halo slots of the vector.  Mailboxes are per level and double-buffered by
//                  the parity of the per-level sequence number: a neighbour can be at most
//                  one exchange of that level ahead (it needs this rank's data to go on).
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
  unsigned int lseq[kMaxLevels] = {};  // halo-exchange sequence number per level
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
        atomicCAS(&ctl->status, 0, 3);
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
        atomicCAS(&ctl->status, 0, 3);
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
}

// One exchange of one vector on one level, as seen by one rank.
struct ExchArgs {
  int level = 0;
  int nnbr = 0;
  int n_owned = 0;                      // halo slots of the vector start here
  int n_halo = 0;                       // entries to receive
  int nbr[kMaxWorld - 1] = {};          // ranks exchanged with: everyone this rank sends to or
                                        // receives from (a symmetric relation); each pair also
                                        // swaps one "present" word per exchange, so that neither
                                        // side can run more than one exchange of the level ahead
  int send_begin[kMaxWorld] = {};       // ranges of send_idx per neighbour
  long long dst_word[2][kMaxWorld - 1] = {};  // the neighbour's mailbox of the level, per parity
  int dst_entry[kMaxWorld - 1] = {};    // first halo entry there that this rank fills
  long long ack_word[2][kMaxWorld - 1] = {};  // this rank's "present" word on the neighbour
  long long box_word[2] = {};           // this rank's mailbox of the level, per parity
  long long box_ack[2] = {};            // the neighbours' "present" words in it (indexed by rank)
  const int* send_idx = nullptr;        // local owned indices
};

// T = double (W = 2 mailbox words per entry) or double2 (W = 4); the level-0 mailbox is laid
// out with 4 words per entry, coarser ones with 2 (shard.h: ll_words).
template <typename T>
__device__ __forceinline__ void halo_exchange_body(Ctl* ctl, Comm* c, const ExchArgs& a,
                                                   T* __restrict__ vec) {
  constexpr int W = sizeof(T) / 4;
  unsigned int s = c->lseq[a.level] + 1u;
  if (s == 0u) s = 1u;
  const int par = static_cast<int>(s & 1u);
  for (int j = 0; j < a.nnbr; ++j) {
    unsigned long long* dst = c->peer[a.nbr[j]] + a.dst_word[par][j] +
                              static_cast<long long>(a.dst_entry[j]) * W;
    const int b = a.send_begin[j], e = a.send_begin[j + 1];
    for (int k = b + threadIdx.x; k < e; k += blockDim.x) {
      const T v = vec[a.send_idx[k]];
      const double* d = reinterpret_cast<const double*>(&v);
#pragma unroll
      for (int w = 0; w < W / 2; ++w) st_f64(dst + static_cast<long long>(k - b) * W + 2 * w, d[w], s);
    }
  }
  if (threadIdx.x < a.nnbr)
    st_word(c->peer[a.nbr[threadIdx.x]] + a.ack_word[par][threadIdx.x], 0u, s);
  const unsigned long long* box = c->peer[c->rank] + a.box_word[par];
  bool ok = true;
  if (threadIdx.x < a.nnbr)
    poll_word(ctl, c->peer[c->rank] + a.box_ack[par] + a.nbr[threadIdx.x], s, &ok);
  // receive: four entries per thread in flight (first-try loads issued back to back), then
  // only the entries whose words have not landed yet are polled again
  constexpr int P = W / 2;  // 16-byte pairs per entry
  for (int h0 = threadIdx.x; h0 < a.n_halo && ok; h0 += 4 * blockDim.x) {
    unsigned long long lo[4][P], hi[4][P];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int h = h0 + u * blockDim.x;
      if (h < a.n_halo) {
#pragma unroll
        for (int w = 0; w < P; ++w) ld_pair(box + static_cast<long long>(h) * W + 2 * w, &lo[u][w], &hi[u][w]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int h = h0 + u * blockDim.x;
      if (h < a.n_halo) {
        T v;
        double* d = reinterpret_cast<double*>(&v);
#pragma unroll
        for (int w = 0; w < P; ++w)
          d[w] = pair_ready(lo[u][w], hi[u][w], s)
                     ? pair_value(lo[u][w], hi[u][w])
                     : poll_f64(ctl, box + static_cast<long long>(h) * W + 2 * w, s, &ok);
        if (ok) vec[a.n_owned + h] = v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) c->lseq[a.level] = s;
}

template <typename T>
__global__ void __launch_bounds__(1024)
k_halo_exchange(Ctl* ctl, Comm* c, ExchArgs a, T* __restrict__ vec) {
  griddep_enter();
  if (ctl->status != 0) return;
  halo_exchange_body<T>(ctl, c, a, vec);
}

// psi is double-buffered: exchange the buffer that holds the current psi.
__global__ void __launch_bounds__(1024)
k_halo_exchange_psi(Ctl* ctl, Comm* c, ExchArgs a, double2* psi0, double2* psi1) {
  griddep_enter();
  if (ctl->status != 0) return;
  halo_exchange_body<double2>(ctl, c, a, ctl->cur ? psi1 : psi0);
}

}  // namespace tdgl
