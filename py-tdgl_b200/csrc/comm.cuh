// Device-side exchange steps of the sharded engine (sm_100a, NVLink peer memory).
//
// Every rank maps every peer's arena (CUDA IPC across processes, or plainly the same
// address space when several shards share one process) and the exchange steps are plain
// device code on those mappings:
//
//   halo exchange  k_halo_exchange: st.global of the rank's boundary entries straight into
//                  the neighbours' halo slots, fence.sys, st.release.sys of the exchange
//                  sequence number into each neighbour's flag word, ld.acquire.sys spin on
//                  the rank's own flag words.  No staging buffers, no host, no NCCL call:
//                  the step needs ~300 exchanges of <= 100 KB, which is a latency problem.
//   all-reduce     comm_allreduce: called by ONE thread (the last block of a reducing
//                  kernel / a controller kernel): stores the partial into every rank's slot,
//                  then reads all slots in rank order, so every rank gets the same bits and
//                  takes the same loop decisions.
//
// Safety: flags carry monotonically increasing sequence numbers (never reset), reduction
// slots are double-buffered by sequence parity, every spin is bounded by a wall-clock
// timeout that raises Ctl::status = 3 (reported as TDGL_E_CUDA) instead of hanging.
#pragma once

// (included by kernels.cuh right after the definition of Ctl)
#include "shard.h"

namespace tdgl {

struct Comm {
  int rank = 0, world = 1;
  unsigned long long hseq = 0;   // halo-exchange sequence number (device mutated)
  unsigned long long rseq = 0;   // all-reduce sequence number (device mutated)
  double* peer[kMaxWorld] = {};  // arena base of every rank as mapped on this device
};

constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;  // 4 s

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Spin until *flag >= want.  Returns false on timeout / when another failure was raised.
__device__ __forceinline__ bool spin_until(Ctl* ctl, const unsigned long long* flag,
                                           unsigned long long want) {
  if (ld_acquire_sys(flag) >= want) return true;
  const unsigned long long t0 = global_ns();
  unsigned int n = 0;
  while (ld_acquire_sys(flag) < want) {
    if ((++n & 1023u) == 0) {
      if (global_ns() - t0 > kSpinTimeoutNs ||
          *reinterpret_cast<volatile int*>(&ctl->status) != 0) {
        atomicCAS(&ctl->status, 0, 3);
        return false;
      }
    }
  }
  return true;
}

// v[0..n) <- reduction over all ranks (sum in rank order, or max).  One calling thread.
__device__ __forceinline__ void comm_allreduce(Ctl* ctl, Comm* c, double* v, int n, bool is_max) {
  const unsigned long long s = c->rseq + 1;
  const int par = static_cast<int>(s & 1ull);
  const int me = c->rank, world = c->world;
  for (int q = 0; q < world; ++q) {
    double* slot = c->peer[q] + kArenaRedSlot + (par * kMaxWorld + me) * 4;
    for (int k = 0; k < n; ++k) reinterpret_cast<volatile double*>(slot)[k] = v[k];
  }
  __threadfence_system();
  for (int q = 0; q < world; ++q)
    st_release_sys(reinterpret_cast<unsigned long long*>(c->peer[q] + kArenaRedFlag) + me, s);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const unsigned long long* flags =
      reinterpret_cast<const unsigned long long*>(c->peer[me] + kArenaRedFlag);
  bool ok = true;
  for (int q = 0; q < world && ok; ++q) {
    ok = spin_until(ctl, flags + q, s);
    const double* slot = c->peer[me] + kArenaRedSlot + (par * kMaxWorld + q) * 4;
    for (int k = 0; k < n; ++k) {
      const double x = ld_relaxed_sys_f64(slot + k);
      acc[k] = (q == 0) ? x : (is_max ? fmax(acc[k], x) : acc[k] + x);
    }
  }
  if (ok)
    for (int k = 0; k < n; ++k) v[k] = acc[k];
  c->rseq = s;
}

// One exchange of one vector on one level, as seen by one rank.
struct ExchArgs {
  int nnbr = 0;
  int nbr[kMaxWorld - 1] = {};         // ranks exchanged with (symmetric relation)
  int send_begin[kMaxWorld] = {};      // ranges of send_idx per neighbour
  long long dst_off[kMaxWorld - 1] = {};  // first destination element on the neighbour, in
                                          // units of T from its arena base
  const int* send_idx = nullptr;       // local owned indices
};

template <typename T>
__device__ __forceinline__ void halo_exchange_body(Ctl* ctl, Comm* c, const ExchArgs& a,
                                                   const T* __restrict__ src) {
  const unsigned long long s = c->hseq + 1;
  for (int j = 0; j < a.nnbr; ++j) {
    T* dst = reinterpret_cast<T*>(c->peer[a.nbr[j]]) + a.dst_off[j];
    const int b = a.send_begin[j], e = a.send_begin[j + 1];
    for (int k = b + threadIdx.x; k < e; k += blockDim.x) dst[k - b] = src[a.send_idx[k]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.nnbr) {
    const int q = a.nbr[threadIdx.x];
    st_release_sys(reinterpret_cast<unsigned long long*>(c->peer[q] + kArenaHaloFlag) + c->rank, s);
    spin_until(ctl, reinterpret_cast<const unsigned long long*>(c->peer[c->rank] + kArenaHaloFlag) + q, s);
  }
  __syncthreads();
  if (threadIdx.x == 0) c->hseq = s;
}

template <typename T>
__global__ void __launch_bounds__(1024)
k_halo_exchange(Ctl* ctl, Comm* c, ExchArgs a, const T* __restrict__ src) {
  if (ctl->status != 0) return;
  halo_exchange_body<T>(ctl, c, a, src);
}

// psi is double-buffered: exchange the buffer that holds the current psi.
__global__ void __launch_bounds__(1024)
k_halo_exchange_psi(Ctl* ctl, Comm* c, ExchArgs a0, ExchArgs a1, const double2* psi0,
                    const double2* psi1) {
  if (ctl->status != 0) return;
  if (ctl->cur) halo_exchange_body<double2>(ctl, c, a1, psi1);
  else halo_exchange_body<double2>(ctl, c, a0, psi0);
}

}  // namespace tdgl
