// Device kernels of the TDGL stepper (sm_100a).  Everything here is HBM-bound fp64 /
// complex128 streaming over CSR row windows: no dense contraction, hence no tensor cores.
//
// Conventions
//  * all per-step scalars (dt, CG alpha/beta, loop conditions) live in the device-resident
//    control block `Ctl`, so the same kernels run from a host-driven launch sequence or
//    from inside one CUDA graph with device-side WHILE loops (no host round trips);
//  * reductions are deterministic: per-block partials, last block to finish adds them in
//    a fixed order (the reference is bitwise deterministic run to run);
//  * every kernel returns immediately once `ctl->status != 0`.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tdgl {

constexpr int kMaxWindow = 1024;   // adaptive_window cap
constexpr int kMaxProbes = 64;
constexpr int kBlock = 256;
constexpr int kMaxPartials = 8192;  // grid-size cap of reducing kernels
constexpr int kMaxKnots = 32;       // knots of a device-side piecewise-linear table
constexpr int kMaxTerminals = 8;    // terminals with device-side current tables

struct Ctl {
  // --- configuration (host writes) -----------------------------------------------------
  double dt_init, dt_max, multiplier, gamma, u, mu_rtol;
  int adaptive, window, max_retries, cg_max_iter;
  int n_probe, running_capacity;
  // --- Runner state ----------------------------------------------------------------------
  double time, t_end, dt, tentative_dt;
  long long step;
  long long steps_left;
  long long steps_done;
  int finished, status;
  long long failed_step;
  double failed_dt;
  // --- psi step ---------------------------------------------------------------------------
  int cur;            // index of the buffer holding the current psi
  int retries;
  int psi_go;         // loop condition of the dt-retry loop
  int disc_flag;
  unsigned long long max_dpsi_bits;
  long long total_retries;
  // --- adaptive controller (solver.py:698-707) -------------------------------------------
  double dpsi_hist[kMaxWindow];
  long long n_hist;
  // --- CG ---------------------------------------------------------------------------------
  double rz_new, rz_prev, pAp, rr, bb;   // gamma = r.z (this / previous iteration), delta = z.Az
  double alpha_prev;                     // step length of the previous iteration
  // initial guess mu + c1 (mu - mu_prev) + c2 (mu - mu_pp) from the last three solutions:
  // r.d1, d1.d1, r.d2, d1.d2, d2.d2 (d_k = A mu_history_k - A mu) and the optimal c1, c2
  double rd, dd, rd2, d1d2, d2d2, guess_c, guess_c2;
  int cg_it, cg_go;
  long long total_cg_it;
  int step_go;
  // --- scratch for means ------------------------------------------------------------------
  double mu_mean;
  // --- epochs the sharded engine derives its mailbox tags from (comm.cuh) -------------------
  int solve_epoch;    // mu solves started so far (+1): bumped at the start of every step
  int psi_epoch;      // attempts of the psi step so far (+1)
  int psi_tag[2];     // psi_epoch of the attempt that produced each psi buffer
  // --- time-dependent terminal currents I_k(t) and epsilon(r, t) = e0(r) + g(t) e1(r), both as
  //     piecewise-linear tables evaluated by k_step_begin (solver.py:325-345, 364-381, 644-646) --
  int cur_on, cur_nterm, cur_knots, cur_changed;
  double cur_t[kMaxKnots];
  double cur_v[kMaxTerminals][kMaxKnots];   // J_scale-d currents, terminal order of the caller
  double cur_len[kMaxTerminals];            // terminal lengths
  double cur_dens[kMaxTerminals];           // current densities in use (mu_boundary values)
  int eps_on, eps_knots;
  double eps_g;                             // g at the current step
  double eps_t[kMaxKnots], eps_v[kMaxKnots];
  // --- diagnostics of a shard-exchange timeout (status 3): what was waited for ----------------
  unsigned int fail_tag;
  unsigned long long fail_addr;
  // --- screening: Polyak iteration on the induced vector potential (solver.py:650-688) -------
  int scr_on, scr_it, scr_go, scr_max_it;
  double scr_tol, scr_alpha, scr_beta, scr_err;
  unsigned long long scr_err_bits;    // max over the edges of |dA| / |A_induced| (double bits)
  long long total_scr_it;
  // --- separable time-dependent vector potential A(r, t) = f(t) A0(r) (device-side ramp) ----
  // f is piecewise linear through (ramp_t[k], ramp_v[k]), constant outside
  int ramp_on, ramp_changed, ramp_knots, ramp_pad;
  double ramp_f;       // f at the current step
  double ramp_dfdt;    // (f(t) - f(t_prev)) / dt_prev, the reference's backward difference
  double ramp_t[kMaxKnots], ramp_v[kMaxKnots];
  double ramp_amax;    // max |component of A0|: decides the reference's allclose() test, see k_step_begin
  double ramp_f_links; // f of the last rebuild of the link variables (what the operators hold)
  // --- debug timeline (TDGL_B200_TRACE=1): 4 words per traced launch {first CTA in, first CTA
  //     past griddepcontrol.wait, last CTA out, launches}, %globaltimer ns; null = off ----------
  unsigned long long* trace;
};

// Programmatic dependent launch: `griddep_wait` blocks until the kernels this launch depends
// on have completed and flushed their results (a no-op for an ordinary launch);
// `griddep_launch_dependents` lets the next kernel of the stream become resident early.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// Debug timeline of the stepping kernels (Ctl::trace; Engine::trace_report).  `id` 0 = untraced.
__device__ __forceinline__ unsigned long long trace_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_in(const Ctl* ctl, int id, int word) {
  if (id > 0 && ctl->trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->trace[4 * id + word] = trace_ns();
    if (word == 0) ctl->trace[4 * id + 3] += 1ull;
  }
}
__device__ __forceinline__ void trace_out(const Ctl* ctl, int id) {
  if (id > 0 && ctl->trace != nullptr && threadIdx.x == 0) atomicMax(ctl->trace + 4 * id + 2, trace_ns());
}
// kernels without a static prologue: let the successor in, then wait for the predecessor
__device__ __forceinline__ void griddep_enter() {
  griddep_launch_dependents();
  griddep_wait();
}

}  // namespace tdgl
#include "comm.cuh"
namespace tdgl {

// ------------------------------------------------------------------------------------------
// helpers

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int W>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, W);
  return v;
}

// Block-wide sum of `v` (every thread contributes); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = 0.0;
  if (warp == 0) {
    t = (lane < nw) ? smem[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;
}

// Deterministic grid reduction.  Each block calls this with its partial (thread 0's value
// is used); returns true in ALL threads of the last block to arrive, with *total holding
// the sum of all partials added in block order.  `counter` must be zero on entry and is
// re-armed on exit.
__device__ __forceinline__ bool grid_sum_last(double partial, double* partials,
                                              unsigned int* counter, double* smem,
                                              double* total) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = partial;
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  // (the tail of every reducing kernel: all loads of a trip are issued before the first add —
  // one L2 round trip per eight partials instead of one per partial; same order of additions)
  double acc = 0.0;
  const unsigned int n = gridDim.x, bd = blockDim.x;
  unsigned int i = threadIdx.x;
  for (; i + 7u * bd < n; i += 8u * bd) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(partials + i + k * bd);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  for (; i < n; i += bd) acc += __ldcg(partials + i);
  acc = block_sum(acc, smem);
  if (threadIdx.x == 0) {
    *total = acc;
    *counter = 0u;
  }
  return true;
}

// Deterministic grid reduction of TWO sums with two block barriers in all (a CTA's warps reach
// them at different times — every barrier is a wait for the slowest warp): warp sums -> shared
// memory -> thread 0 adds them in warp order and publishes the block's pair -> last block adds
// the pairs in block order.  v0 / v1: every thread's contributions.  Same contract as
// grid_sum_last otherwise (smem: >= 64 doubles).
__device__ __forceinline__ bool grid_sum2_fused(double v0, double v1, double* partials,
                                                unsigned int* counter, double* smem, double* t0,
                                                double* t1) {
  __shared__ int s_last2f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v0 = warp_sum(v0);
  v1 = warp_sum(v1);
  if (lane == 0) { smem[warp] = v0; smem[32 + warp] = v1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < nw; ++k) { a += smem[k]; b += smem[32 + k]; }
    reinterpret_cast<double2*>(partials)[blockIdx.x] = make_double2(a, b);
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last2f = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last2f) return false;
  __threadfence();
  double a0 = 0.0, a1 = 0.0;
  const double2* p2 = reinterpret_cast<const double2*>(partials);
  const unsigned int n = gridDim.x, bd = blockDim.x;
  unsigned int i = threadIdx.x;
  for (; i + 7u * bd < n; i += 8u * bd) {
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(p2 + i + k * bd);
#pragma unroll
    for (int k = 0; k < 8; ++k) { a0 += v[k].x; a1 += v[k].y; }
  }
  for (; i < n; i += bd) {
    const double2 v = __ldcg(p2 + i);
    a0 += v.x;
    a1 += v.y;
  }
  a0 = block_sum(a0, smem);
  a1 = block_sum(a1, smem);
  if (threadIdx.x == 0) {
    *t0 = a0;
    *t1 = a1;
    *counter = 0u;
  }
  return true;
}

#define TDGL_COND_ARG cudaGraphConditionalHandle

__device__ __forceinline__ void set_cond(cudaGraphConditionalHandle h, int v) {
  if (h != 0) cudaGraphSetConditional(h, v ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------
// (the CSR kernels live in csr_window.cuh)

// Coarsest level: x = Minv b with a dense row-major rows x ld float matrix (ld = nc rounded up
// to a multiple of 4, zero padded).  The level may hold a couple of thousand rows (fewer AMG
// levels = fewer dependent launches per V-cycle), i.e. ~10 MB per application: two rows per
// CTA, four warps per row, each warp streaming 512-byte pieces of its quarter of the row with
// 16-byte loads — rows x 4 warps keep the whole machine loading.  Partial sums are combined in
// a fixed order (bitwise reproducible).
template <typename TB, typename TX>
__global__ void __launch_bounds__(kBlock)
k_dense_matvec(const Ctl* __restrict__ ctl, int rows, int nc, int ld, const float* __restrict__ M,
               const TB* __restrict__ b, TX* __restrict__ x, int trace_id) {
  trace_in(ctl, trace_id, 0);
  griddep_enter();
  trace_in(ctl, trace_id, 1);
  if (ctl->status != 0) return;
  __shared__ double part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 2 + (warp >> 2);
  double s = 0.0;
  if (row < rows) {
    const float4* __restrict__ m4 = reinterpret_cast<const float4*>(M + static_cast<size_t>(row) * ld);
#pragma unroll 4
    for (int j = (warp & 3) * 128 + lane * 4; j < nc; j += 512) {
      const float4 m = __ldg(m4 + (j >> 2));
      const double b0 = static_cast<double>(b[j]);
      const double b1 = j + 1 < nc ? static_cast<double>(b[j + 1]) : 0.0;
      const double b2 = j + 2 < nc ? static_cast<double>(b[j + 2]) : 0.0;
      const double b3 = j + 3 < nc ? static_cast<double>(b[j + 3]) : 0.0;
      s = fma(static_cast<double>(m.x), b0, s);
      s = fma(static_cast<double>(m.y), b1, s);
      s = fma(static_cast<double>(m.z), b2, s);
      s = fma(static_cast<double>(m.w), b3, s);
    }
  }
  s = warp_sum(s);
  if (lane == 0) part[warp] = s;
  __syncthreads();
  if (threadIdx.x < 2) {
    const int r = blockIdx.x * 2 + threadIdx.x;
    const double* q = part + 4 * threadIdx.x;
    if (r < rows) x[r] = static_cast<TX>(((q[0] + q[1]) + q[2]) + q[3]);
  }
  trace_out(ctl, trace_id);
}

// ------------------------------------------------------------------------------------------
// CG vector kernels

// One iteration of the single-reduction form of preconditioned CG (Chronopoulos & Gear): with
// z = M r, w = A z, gamma = r.z and delta = z.w from the SpMV kernel (ONE reduction),
//   beta = gamma / gamma_prev ;  alpha = gamma / (delta - beta gamma / alpha_prev)
//   p = z + beta p ;  s = w + beta s  (= A p) ;  x += alpha p ;  r -= alpha s ;  rr = ||r||^2
// in one pass over the vectors.  The same iterates as textbook PCG (p = z + beta p, alpha =
// r.z / p.Ap) in exact arithmetic, with one launch and one grid reduction less per iteration.
// Last block: bookkeeping + loop condition of the CG loop.
template <bool SH>
__global__ void __launch_bounds__(kBlock, 5)
k_cg_fused(Ctl* ctl, Comm* comm, PushArgs push, int n, const float* __restrict__ z,
           const double* __restrict__ w, double* __restrict__ p, double* __restrict__ s,
           double* __restrict__ x, double* __restrict__ r, const float* __restrict__ dinv,
           float* __restrict__ x0, double omega, double* partials, unsigned int* counter,
           cudaGraphConditionalHandle cond, int trace_id) {
  trace_in(ctl, trace_id, 0);
  griddep_enter();
  trace_in(ctl, trace_id, 1);
  __shared__ double red[32];
  if (ctl->status != 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ctl->cg_go = 0;   // (the host-driven loop reads this, the graph the conditional)
      set_cond(cond, 0);
    }
    return;
  }
  const double gamma = ctl->rz_new, delta = ctl->pAp;
  const bool first = ctl->cg_it == 0;
  const double beta = first ? 0.0 : gamma / ctl->rz_prev;
  const double denom = first ? delta : delta - beta * gamma / ctl->alpha_prev;
  const double alpha = gamma / denom;
  const unsigned int tag = (SH && comm != nullptr) ? comm_tag(ctl, push.tag_mode) : 0u;
  double d = 0.0;
  // two independent elements per trip: twelve loads in flight per thread
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
    const int j = i + stride;
    const bool two = j < n;
    const double zi = z[i], wi = w[i], xi = x[i], ri0 = r[i];
    const double pi0 = first ? 0.0 : p[i], si0 = first ? 0.0 : s[i];
    double zj = 0.0, wj = 0.0, xj = 0.0, rj0 = 0.0, pj0 = 0.0, sj0 = 0.0;
    if (two) {
      zj = z[j]; wj = w[j]; xj = x[j]; rj0 = r[j];
      if (!first) { pj0 = p[j]; sj0 = s[j]; }
    }
    const double pi = first ? zi : zi + beta * pi0;
    const double si = first ? wi : wi + beta * si0;
    p[i] = pi;
    s[i] = si;
    x[i] = xi + alpha * pi;
    const double ri = ri0 - alpha * si;
    r[i] = ri;
    // the next V-cycle's pre-smoothed iterate from a zero guess, x0 = omega D^-1 r, is formed
    // here (one float per row), so that its residual kernel gathers ONE float per matrix entry
    // instead of r and 1/diag; sharded: x0's boundary rows travel, not r's
    if (x0 != nullptr) {
      const float ti = static_cast<float>(omega * static_cast<double>(dinv[i]) * ri);
      x0[i] = ti;
      if (SH && comm != nullptr) push_row(comm, push, tag, i, static_cast<double>(ti));
    }
    d += ri * ri;
    if (two) {
      const double pj = first ? zj : zj + beta * pj0;
      const double sj = first ? wj : wj + beta * sj0;
      p[j] = pj;
      s[j] = sj;
      x[j] = xj + alpha * pj;
      const double rj = rj0 - alpha * sj;
      r[j] = rj;
      if (x0 != nullptr) {
        const float tj = static_cast<float>(omega * static_cast<double>(dinv[j]) * rj);
        x0[j] = tj;
        if (SH && comm != nullptr) push_row(comm, push, tag, j, static_cast<double>(tj));
      }
      d += rj * rj;
    }
  }
  const double bs = block_sum(d, red);
  double total;
  if (grid_sum_last(bs, partials, counter, red, &total) && threadIdx.x < 32) {
    total = __shfl_sync(0xffffffffu, total, 0);
    if (SH && comm != nullptr) comm_allreduce(ctl, comm, &total, 1, false);
    if (threadIdx.x == 0) {
      ctl->rr = total;
      ctl->rz_prev = gamma;
      ctl->alpha_prev = alpha;
      const int it = ctl->cg_it + 1;
      ctl->cg_it = it;
      ctl->total_cg_it += 1;
      const double tol2 = ctl->mu_rtol * ctl->mu_rtol * ctl->bb;
      int go = (total > tol2) ? 1 : 0;
      if (ctl->status != 0) {  // exchange failure raised meanwhile
        go = 0;
      } else if (!(total == total) || !(denom > 0.0)) {  // NaN / loss of positivity: breakdown
        ctl->status = 2; ctl->failed_step = ctl->step; ctl->failed_dt = ctl->dt; go = 0;
      } else if (go && it >= ctl->cg_max_iter) {
        ctl->status = 2; ctl->failed_step = ctl->step; ctl->failed_dt = ctl->dt; go = 0;
      }
      ctl->cg_go = go;
      set_cond(cond, go);
    }
  }
  trace_out(ctl, trace_id);
}

// Decide whether the CG loop has to run at all (warm start may already satisfy the
// tolerance) and arm its counters.
__global__ void k_cg_begin(Ctl* ctl, cudaGraphConditionalHandle cond, int with_guess) {
  griddep_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int go = 0;
  // optimal extrapolation of the initial guess (see kw_mu_rhs): minimise
  // || r + c1 d1 + c2 d2 ||^2 — a 2 x 2 least-squares problem; falls back to the one-term
  // guess c1 = -r.d1 / d1.d1 when the history is degenerate (first steps, d2 ~ d1), and to
  // the plain warm start when there is no history at all
  double c = 0.0, c2 = 0.0;
  if (with_guess && ctl->dd > 0.0 && ctl->status == 0) {
    const double g11 = ctl->dd, g12 = ctl->d1d2, g22 = ctl->d2d2;
    const double det = g11 * g22 - g12 * g12;
    bool two = with_guess > 1 && g22 > 0.0 && det > 1e-8 * g11 * g22;
    if (two) {
      c = (-ctl->rd * g22 + ctl->rd2 * g12) / det;
      c2 = (-ctl->rd2 * g11 + ctl->rd * g12) / det;
      if (!(c == c) || !(c2 == c2) || fabs(c) > 16.0 || fabs(c2) > 16.0) two = false;
    }
    if (!two) {
      c2 = 0.0;
      c = -ctl->rd / ctl->dd;
      if (!(c == c) || fabs(c) > 8.0) c = 0.0;   // (degenerate history: plain warm start)
    }
    const double rr = ctl->rr + 2.0 * (c * ctl->rd + c2 * ctl->rd2) + c * c * g11 +
                      2.0 * c * c2 * g12 + c2 * c2 * g22;
    ctl->rr = rr > 0.0 ? rr : 0.0;
  }
  ctl->guess_c2 = c2;
  ctl->guess_c = c;
  if (ctl->status == 0) {
    const double tol2 = ctl->mu_rtol * ctl->mu_rtol * ctl->bb;
    go = (ctl->rr > tol2) ? 1 : 0;
    if (!(ctl->rr == ctl->rr) || !(ctl->bb == ctl->bb)) {
      ctl->status = 2; ctl->failed_step = ctl->step; ctl->failed_dt = ctl->dt; go = 0;
    }
  }
  ctl->cg_it = 0;
  ctl->cg_go = go;
  set_cond(cond, go);
}

// x0 = omega D^-1 r for a residual that no stepping kernel produced (single-operator calls)
__global__ void k_x0_from_r(int n, const double* __restrict__ r, const float* __restrict__ dinv,
                            double omega, float* __restrict__ x0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x0[i] = static_cast<float>(omega * static_cast<double>(dinv[i]) * r[i]);
}

// Applies the extrapolated initial guess chosen by k_cg_begin:
//   mu <- mu + c1 (mu - mu_prev) + c2 (mu - mu_pp),  r <- r + c1 d1 + c2 d2,
// and shifts the history (mu_pp <- mu_prev, mu_prev <- old mu).  Sharded: the boundary rows of
// x0 = omega D^-1 r go to the neighbours (iteration 0's V-cycle input), and the halo slots of mu_pp are filled
// from mu_prev's mailbox copy (threads n .. nx-1) before this step's solution overwrites it.
__global__ void __launch_bounds__(kBlock)
k_mu_guess(Ctl* ctl, const Comm* comm, PushArgs push, HaloArgs prev_halo, int n, int nx,
           double* __restrict__ mu, double* __restrict__ mu_prev, double* __restrict__ mu_pp,
           double* __restrict__ r, const double* __restrict__ d1, const double* __restrict__ d2,
           const float* __restrict__ dinv, float* __restrict__ x0, double omega) {
  griddep_enter();
  if (ctl->status != 0) return;
  const double c1 = ctl->guess_c, c2 = ctl->guess_c2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double old = mu[i], prev = mu_prev[i];
    mu[i] = old + c1 * (old - prev) + c2 * (old - mu_pp[i]);
    mu_pp[i] = prev;
    mu_prev[i] = old;
    const double ri = r[i] + c1 * d1[i] + c2 * d2[i];
    r[i] = ri;
    if (x0 != nullptr) {   // iteration 0's pre-smoothed iterate (see k_cg_fused)
      const float ti = static_cast<float>(omega * static_cast<double>(dinv[i]) * ri);
      x0[i] = ti;
      if (comm != nullptr) push_row(comm, push, comm_tag(ctl, push.tag_mode), i, static_cast<double>(ti));
    }
  } else if (comm != nullptr && i < nx) {
    const HaloView hv = halo_view(ctl, comm, prev_halo);
    mu_pp[i] = halo_get(ctl, hv, mu_prev, i);
  }
}

// ------------------------------------------------------------------------------------------
// pointwise part of the psi step: closed-form |psi|^2 update
// (reference TDGLSolver.solve_for_psi_squared, tdgl/solver/solver.py:418-438)

struct PsiOut {
  double2 psi;
  double sq;
  int failed;
};

// abs2 = |psi|^2 of the step's INPUT state (solver.py:649): equal to |psi|^2 of the psi passed
// here except in the 2nd, 3rd, ... pass of a screening iteration, where the reference steps
// from the previous pass's psi but keeps the step's first |psi|^2 (solver.py:676-678).
__device__ __forceinline__ PsiOut psi_update(double2 psi, double2 lap, double mu, double eps,
                                             double gamma, double u, double dt, double abs2) {
  // U = exp(-i mu dt)
  double s, c;
  sincos(-mu * dt, &s, &c);
  const double2 U = make_double2(c, s);
  // z = U * gamma^2 / 2 * psi
  const double g2h = gamma * gamma / 2.0;
  const double2 Ug = make_double2(U.x * g2h, U.y * g2h);
  const double2 z = make_double2(Ug.x * psi.x - Ug.y * psi.y, Ug.x * psi.y + Ug.y * psi.x);
  // w = z |psi|^2 + U (psi + dt/u sqrt(1 + gamma^2 |psi|^2) ((eps - |psi|^2) psi + L psi))
  const double f = (dt / u) * sqrt(1.0 + gamma * gamma * abs2);
  const double e = eps - abs2;
  const double2 inner = make_double2(psi.x + f * (e * psi.x + lap.x),
                                     psi.y + f * (e * psi.y + lap.y));
  const double2 w = make_double2(z.x * abs2 + (U.x * inner.x - U.y * inner.y),
                                 z.y * abs2 + (U.x * inner.y + U.y * inner.x));
  const double cc = w.x * z.x + w.y * z.y;
  const double two_c_1 = 2.0 * cc + 1.0;
  const double w2 = w.x * w.x + w.y * w.y;
  const double z2 = z.x * z.x + z.y * z.y;
  const double disc = two_c_1 * two_c_1 - 4.0 * z2 * w2;
  PsiOut o;
  // the reference gives up on disc < 0 or any floating-point error (solver.py:420-436)
  o.failed = !(disc >= 0.0) || !isfinite(w2) || !isfinite(disc);
  const double sq = (2.0 * w2) / (two_c_1 + sqrt(disc));
  o.sq = sq;
  o.psi = make_double2(w.x - z.x * sq, w.y - z.y * sq);
  if (!o.failed && !(isfinite(sq) && isfinite(o.psi.x) && isfinite(o.psi.y))) o.failed = 1;
  return o;
}

// Piecewise-linear table lookup, constant outside the knots (the same arithmetic as the host
// classes in tdgl_b200/sources.py, which the parity tests feed to the oracle).
__device__ __forceinline__ double table_lookup(const double* __restrict__ t, const double* __restrict__ v,
                                               int n, double x) {
  if (x >= t[n - 1]) return v[n - 1];
  if (x <= t[0]) return v[0];
  int k = 0;
  while (k + 2 < n && x >= t[k + 1]) ++k;
  const double w = __ddiv_rn(__dsub_rn(x, t[k]), __dsub_rn(t[k + 1], t[k]));
  return __dadd_rn(v[k], __dmul_rn(w, __dsub_rn(v[k + 1], v[k])));
}

// Start of TDGLSolver.update (solver.py:649-668): dt <- tentative_dt, retries <- 0.
__global__ void k_step_begin(Ctl* ctl, cudaGraphConditionalHandle cond_psi,
                             cudaGraphConditionalHandle cond_scr) {
  griddep_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (ctl->ramp_on) {
    // TDGLSolver.update for a time-dependent vector potential (solver.py:626-642) with
    // A(r, t) = f(t) A0(r): dA/dt = (f(t) - f_prev) / dt_prev * A0, dt_prev = Runner's dt
    // argument (the previous step's dt); link variables are rebuilt only if f changed
    const double t = ctl->time;
    const int n = ctl->ramp_knots;
    double f = ctl->ramp_v[0];
    if (t >= ctl->ramp_t[n - 1]) {
      f = ctl->ramp_v[n - 1];
    } else if (t > ctl->ramp_t[0]) {
      int k = 0;
      while (k + 2 < n && t >= ctl->ramp_t[k + 1]) ++k;
      const double w = (t - ctl->ramp_t[k]) / (ctl->ramp_t[k + 1] - ctl->ramp_t[k]);
      f = ctl->ramp_v[k] + w * (ctl->ramp_v[k + 1] - ctl->ramp_v[k]);
    }
    ctl->ramp_dfdt = (f - ctl->ramp_f) / ctl->dt;
    // The reference rebuilds the link variables only `if not allclose(A_new, A_prev)`
    // (solver.py:635-638; numpy's rtol = 1e-5, atol = 1e-8; A_prev = the previous step's A
    // whether or not it was used), i.e. when some component violates
    // |df| |A0| <= atol + rtol |f_prev| |A0|.  With A = f A0 the component with the largest |A0|
    // decides: rebuild iff (|df| - rtol |f_prev|) max|A0| > atol.  (A slow ramp therefore keeps
    // the link variables of an older f — the reference's behaviour, reproduced on purpose.)
    const double df = fabs(f - ctl->ramp_f);
    const bool moved = (df - 1e-5 * fabs(ctl->ramp_f)) * ctl->ramp_amax > 1e-8;
    ctl->ramp_changed = moved || ctl->ramp_changed == 2;  // 2: forced (set-up)
    ctl->ramp_f = f;
    if (ctl->ramp_changed) ctl->ramp_f_links = f;
  }
  if (ctl->cur_on) {
    // update_mu_boundary (solver.py:325-345): J_ext,k = -(1 / L_k) sum_{j != k} I_j(t)
    int changed = ctl->cur_changed == 2;   // 2: forced (set-up)
    for (int k = 0; k < ctl->cur_nterm; ++k) {
      double sum = 0.0;
      for (int j = 0; j < ctl->cur_nterm; ++j)
        if (j != k) sum = __dadd_rn(sum, table_lookup(ctl->cur_t, ctl->cur_v[j], ctl->cur_knots, ctl->time));
      const double dens = __dmul_rn(__ddiv_rn(-1.0, ctl->cur_len[k]), sum);
      if (dens != ctl->cur_dens[k]) { ctl->cur_dens[k] = dens; changed = 1; }
    }
    ctl->cur_changed = changed;
  }
  if (ctl->eps_on) ctl->eps_g = table_lookup(ctl->eps_t, ctl->eps_v, ctl->eps_knots, ctl->time);
  ctl->dt = ctl->tentative_dt;
  ctl->scr_it = 0;
  ctl->scr_err_bits = 0ull;
  ctl->scr_err = 1e300;
  ctl->solve_epoch += 1;
  ctl->retries = 0;
  ctl->disc_flag = 0;
  ctl->max_dpsi_bits = 0ull;
  const int go = (ctl->status == 0) ? 1 : 0;
  ctl->psi_go = go;
  ctl->scr_go = go;
  set_cond(cond_psi, go);
  set_cond(cond_scr, go);
}

// adaptive_euler_step's retry logic (solver.py:475-485)
__global__ void k_psi_control(Ctl* ctl, Comm* comm, cudaGraphConditionalHandle cond_psi) {
  griddep_enter();
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  int go = 0;
  if (comm != nullptr && ctl->status == 0) {  // one full warp (launched with 32 threads)
    // any(disc < 0) and max |d psi^2| over all shards
    double v[2] = {ctl->disc_flag ? 1.0 : 0.0,
                   __longlong_as_double(static_cast<long long>(ctl->max_dpsi_bits))};
    comm_allreduce(ctl, comm, v, 2, true);
    if (threadIdx.x == 0) {
      ctl->disc_flag = v[0] > 0.0 ? 1 : 0;
      ctl->max_dpsi_bits = static_cast<unsigned long long>(__double_as_longlong(v[1]));
    }
  }
  if (threadIdx.x != 0) return;
  ctl->psi_epoch += 1;  // = the tag the attempt just made stored its boundary rows with
  if (ctl->status == 0) {
    if (ctl->disc_flag) {
      if (!ctl->adaptive || ctl->retries > ctl->max_retries) {
        ctl->status = 1;
        ctl->failed_step = ctl->step;
        ctl->failed_dt = ctl->dt;
      } else {
        ctl->dt = ctl->dt * ctl->multiplier;
        ctl->retries += 1;
        ctl->total_retries += 1;
        ctl->disc_flag = 0;
        ctl->max_dpsi_bits = 0ull;
        go = 1;
      }
    } else {
      ctl->cur ^= 1;  // accept: the freshly written buffer becomes the current psi
      ctl->psi_tag[ctl->cur] = ctl->psi_epoch;
    }
  }
  ctl->psi_go = go;
  set_cond(cond_psi, go);
}

// area-weighted mean of mu (deterministic), then mu -= mean
__global__ void __launch_bounds__(kBlock)
k_weighted_sum(Ctl* ctl, Comm* comm, int n, const double* __restrict__ w,
               const double* __restrict__ x, double* partials, unsigned int* counter,
               double inv_total_weight) {
  griddep_enter();
  __shared__ double red[32];
  if (ctl->status != 0) return;
  double d = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    d += w[i] * x[i];
  const double bs = block_sum(d, red);
  double total;
  if (grid_sum_last(bs, partials, counter, red, &total) && threadIdx.x < 32) {
    total = __shfl_sync(0xffffffffu, total, 0);
    if (comm != nullptr) comm_allreduce(ctl, comm, &total, 1, false);
    if (threadIdx.x == 0) ctl->mu_mean = total * inv_total_weight;
  }
}

__global__ void __launch_bounds__(kBlock)
k_shift(const Ctl* __restrict__ ctl, const Comm* comm, PushArgs push, int n,
        double* __restrict__ x) {
  griddep_enter();
  if (ctl->status != 0) return;
  const double m = ctl->mu_mean;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = x[i] - m;
    x[i] = v;
    if (comm != nullptr) push_row(comm, push, comm_tag(ctl, push.tag_mode), i, v);
  }
}

// End of update() + the Runner bookkeeping (solver.py:690-707, runner.py:429-433).
// Thread p < n_probe records the probe values; thread 0 runs the controller.
__global__ void k_step_end(Ctl* ctl, const double2* __restrict__ psi_buf0,
                           const double2* __restrict__ psi_buf1, const double* __restrict__ mu,
                           const int* __restrict__ probes, double* run_dt, double* run_mu,
                           double* run_theta, long long* run_scr /* null: no screening */,
                           cudaGraphConditionalHandle cond_step) {
  griddep_enter();
  if (blockIdx.x != 0) return;
  if (ctl->status != 0) {
    if (threadIdx.x == 0) { ctl->step_go = 0; set_cond(cond_step, 0); }
    return;
  }
  const long long pos = ctl->steps_done;
  const int cap = ctl->running_capacity;
  if (pos < cap && threadIdx.x < ctl->n_probe) {
    const double2* psi = ctl->cur ? psi_buf1 : psi_buf0;
    const int s = probes[threadIdx.x];  // < 0: the probe site belongs to another shard
    run_mu[(size_t)threadIdx.x * cap + pos] = s >= 0 ? mu[s] : 0.0;
    const double2 p = s >= 0 ? psi[s] : make_double2(1.0, 0.0);
    run_theta[(size_t)threadIdx.x * cap + pos] = s >= 0 ? atan2(p.y, p.x) : 0.0;
  }
  if (threadIdx.x != 0) return;
  const double dt = ctl->dt;
  if (pos < cap) run_dt[pos] = dt;
  if (pos < cap && run_scr != nullptr) run_scr[pos] = ctl->scr_it;   // solver.py:695-696
  if (ctl->adaptive) {
    const double d = __longlong_as_double((long long)ctl->max_dpsi_bits);
    ctl->dpsi_hist[ctl->n_hist % kMaxWindow] = d;
    ctl->n_hist += 1;
    const int win = ctl->window;
    if (ctl->step > win) {
      // mean of the last `win` entries (fewer only if the history is shorter)
      long long cnt = ctl->n_hist < win ? ctl->n_hist : win;
      double sum = 0.0;
      for (long long k = ctl->n_hist - cnt; k < ctl->n_hist; ++k)
        sum += ctl->dpsi_hist[k % kMaxWindow];
      const double mean = sum / (double)cnt;
      const double new_dt = ctl->dt_init / fmax(1e-10, mean);
      double t = 0.5 * (new_dt + dt);
      t = fmin(fmax(t, 0.0), ctl->dt_max);
      ctl->tentative_dt = t;
    }
  }
  ctl->steps_done = pos + 1;
  ctl->steps_left -= 1;
  int go = 0;
  if (ctl->time >= ctl->t_end) {
    ctl->finished = 1;
  } else {
    ctl->time += dt;
    ctl->step += 1;
    go = ctl->steps_left > 0 ? 1 : 0;
  }
  ctl->step_go = go;
  set_cond(cond_step, go);
}

// ------------------------------------------------------------------------------------------
// setup / IO kernels

// bterm_i = sum over boundary edges b=(i,j) of  l_b / (2 a_i) * mu_boundary[b]
// (mu_boundary_laplacian @ mu_boundary, operators.py:188-230).  Each boundary site has
// two boundary edges and a + b == b + a, so the atomics are order-independent.
__global__ void k_boundary_term(int nb, const int* __restrict__ be0, const int* __restrict__ be1,
                                const double* __restrict__ blen, const double* __restrict__ areas,
                                const double* __restrict__ mub, double* bterm) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const double v = blen[b] * mub[b];
  const int i = be0[b], j = be1[b];  // < 0: the site belongs to another shard
  if (i >= 0) atomicAdd(bterm + i, v / (2.0 * areas[i]));
  if (j >= 0) atomicAdd(bterm + j, v / (2.0 * areas[j]));
}

// Time-dependent vector potential (reference solver.py:626-642, 507-510, 519): the rhs gains
// -(divergence @ dA_dt).  divergence has (e0, e) = s_e / a[e0], (e1, e) = -s_e / a[e1]
// (operators.py:59-84); gathered per site over its incident edges (deterministic), and folded
// with the boundary-current term into the one site vector the rhs kernel subtracts:
//   bterm_eff_i = bterm_i + sum_e (+-) s_e / a_i * dA_dt[e]
__global__ void k_site_terms(int n, const int* __restrict__ ptr, const int* __restrict__ eidx,
                             const signed char* __restrict__ head,
                             const double* __restrict__ weight, const double* __restrict__ elen,
                             const double* __restrict__ areas, const double* __restrict__ dadt,
                             const double* __restrict__ bterm, double* __restrict__ bterm_eff) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  double s = 0.0;
  if (dadt != nullptr)
    for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
      const int e = eidx[k];
      if (e < 0) continue;
      const double dual = weight[e] * elen[e];
      s += (head[k] ? dual : -dual) / areas[row] * dadt[e];
    }
  bterm_eff[row] = bterm[row] + s;
}

// Device-side ramp: the link variables for A = ramp_f * A0 (theta holds A0 . d), rebuilt by
// the step that sees f change (one pass over the matrix values, ~3 % of a step at 1M sites).
__global__ void k_link_values_ramp(const Ctl* __restrict__ ctl, int n, const int* __restrict__ ptr,
                                   const int* __restrict__ eidx,
                                   const signed char* __restrict__ head,
                                   const double* __restrict__ weight,
                                   const double* __restrict__ theta0,
                                   const double* __restrict__ areas, double2* __restrict__ lval) {
  griddep_enter();
  if (ctl->status != 0 || !ctl->ramp_changed) return;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double f = ctl->ramp_f;
  double diag = 0.0;
  int kd = -1;
  for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
    const int e = eidx[k];
    if (e < 0) { kd = k; continue; }
    const double w = weight[e];
    double s, c;
    sincos(-(f * theta0[e]), &s, &c);
    if (!head[k]) s = -s;
    lval[k] = make_double2(w * c / areas[row], w * s / areas[row]);
    diag += -w / areas[row];
  }
  if (kd >= 0) lval[kd] = make_double2(diag, 0.0);
}

// Values of the covariant Laplacian (all rows kept, see k_mu_rhs) from the link variables
//   U_e = exp(-i A_e . d_e) ; off-diagonals w_e U_e / a_i (row = edges[e,0]) or
//   w_e conj(U_e) / a_i (row = edges[e,1]) ; diagonal -sum w_e / a_i
// (operators.py:149-169, 346-383).
__global__ void k_link_values(int n, const int* __restrict__ ptr, const int* __restrict__ eidx,
                              const signed char* __restrict__ head,
                              const double* __restrict__ weight /* w_e */,
                              const double* __restrict__ theta /* A_e . d_e */,
                              const double* __restrict__ areas, double2* __restrict__ lval) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double inv_a = 1.0 / areas[row];
  double diag = 0.0;
  int kd = -1;
  for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
    const int e = eidx[k];
    if (e < 0) { kd = k; continue; }
    const double w = weight[e];
    double s, c;
    sincos(-theta[e], &s, &c);
    if (!head[k]) s = -s;
    lval[k] = make_double2(w * c / areas[row], w * s / areas[row]);
    diag += -w / areas[row];
  }
  (void)inv_a;
  if (kd >= 0) lval[kd] = make_double2(diag, 0.0);
}

// J_s[e] = Im(conj(psi[e0]) (U_e psi[e1] - psi[e0]) / l_e)      (operators.py:385-394)
// J_n[e] = -(mu[e1] - mu[e0]) / l_e                              (solver.py:519, static A)
// Edge arrays are in the caller's edge order, site indices are internal.
__global__ void k_currents(int ne, const int* __restrict__ elist /* null: all edges */,
                           const int* __restrict__ e0, const int* __restrict__ e1,
                           const double* __restrict__ elen, const double* __restrict__ theta,
                           const double2* __restrict__ psi /* null: the current buffer of ... */,
                           const double2* __restrict__ psi_buf0, const double2* __restrict__ psi_buf1,
                           int mode /* 1: J_s, 2: J_n, 3: both */,
                           const double* __restrict__ mu,
                           const double* __restrict__ dadt /* may be null */,
                           const Ctl* __restrict__ ctl,
                           const double* __restrict__ ramp_proj /* null: no device-side ramp */,
                           const double2* __restrict__ aind /* null: no screening */,
                           const double2* __restrict__ edir,
                           double* __restrict__ js, double* __restrict__ jn) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // output slot
  if (t >= ne) return;
  const int e = elist != nullptr ? elist[t] : t;         // (shard-local output: owned edges only)
  const int i = e0[e], j = e1[e];
  if (i < 0) {  // the edge belongs to another shard (owner = shard of edges[e,0])
    if (mode & 1) js[t] = 0.0;
    if (mode & 2) jn[t] = 0.0;
    return;
  }
  if (psi == nullptr) psi = ctl->cur ? psi_buf1 : psi_buf0;
  const double inv_l = 1.0 / elen[e];
  double s, c;
  // (device-side ramp: theta holds A0 . d; the gradient operator carries the link exponents of
  // the last rebuild, ramp_f_links * A0, like the reference's — with screening they are rebuilt
  // every pass from the current A, ramp_f * A0)
  double th = theta[e];
  if (ramp_proj != nullptr) th *= (aind != nullptr) ? ctl->ramp_f : ctl->ramp_f_links;
  if (aind != nullptr) th += aind[e].x * edir[e].x + aind[e].y * edir[e].y;
  sincos(-th, &s, &c);
  const double2 pi = psi[i], pj = psi[j];
  // g = (U psi_j) * (1/l) + psi_i * (-1/l)   as the CSR gradient row computes it
  const double gx = (c * pj.x - s * pj.y) * inv_l - pi.x * inv_l;
  const double gy = (c * pj.y + s * pj.x) * inv_l - pi.y * inv_l;
  if (mode & 1) js[t] = pi.x * gy - pi.y * gx;
  double da = dadt != nullptr ? dadt[e] : 0.0;
  if (ramp_proj != nullptr) da = ctl->ramp_dfdt * ramp_proj[e];
  if (mode & 2) jn[t] = -(mu[j] * inv_l - mu[i] * inv_l) - da;
}

// Device-side terminal currents: the boundary term of the rhs on the sites that touch a terminal
// edge, rewritten by the step that sees a current density change —
//   bterm_i = sum over the boundary edges b at site i of (l_b * mu_boundary[b]) / (2 a_i)
//             (mu_boundary_laplacian @ mu_boundary, operators.py:188-230, as k_boundary_term)
//           + the dA/dt part of the site (as k_site_terms).
// One thread per listed site; site_ptr / site_bedge: its boundary edges.
__global__ void __launch_bounds__(kBlock)
k_terminal_sites(const Ctl* __restrict__ ctl, int n_list, const int* __restrict__ site,
                 const int* __restrict__ site_ptr, const int* __restrict__ site_bedge,
                 const int* __restrict__ bedge_term, const double* __restrict__ blen,
                 const double* __restrict__ areas, const int* __restrict__ ptr,
                 const int* __restrict__ eidx, const signed char* __restrict__ head,
                 const double* __restrict__ weight, const double* __restrict__ elen,
                 const double* __restrict__ dadt /* may be null */, double* __restrict__ bterm_base,
                 double* __restrict__ bterm) {
  griddep_enter();
  if (ctl->status != 0 || !ctl->cur_changed) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_list) return;
  const int row = site[t];
  double base = 0.0;
  for (int k = site_ptr[t]; k < site_ptr[t + 1]; ++k) {
    const int b = site_bedge[k];
    const int term = bedge_term[b];
    if (term < 0) continue;
    base += (blen[b] * ctl->cur_dens[term]) / (2.0 * areas[row]);
  }
  double s = 0.0;
  if (dadt != nullptr)
    for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
      const int e = eidx[k];
      if (e < 0) continue;
      const double dual = weight[e] * elen[e];
      s += (head[k] ? dual : -dual) / areas[row] * dadt[e];
    }
  bterm_base[row] = base;
  bterm[row] = base + s;
}

// ------------------------------------------------------------------------------------------
// screening: the induced vector potential
//   A_ind[e] = sum_j J_site[j] a~_j / |c_e - r_j|        (solver/screening.py:12-42)
// with J_site the site average of the edge current J_s + J_n (finite_volume/mesh.py:203-243),
// iterated with Polyak's method inside every time step (solver/solver.py:522-578, 650-688).

// |psi|^2 of the step's input state, kept for all passes of the screening loop
__global__ void __launch_bounds__(kBlock)
k_scr_old_sq(const Ctl* __restrict__ ctl, int n, const double2* __restrict__ psi0,
             const double2* __restrict__ psi1, double* __restrict__ old_sq) {
  griddep_enter();
  if (ctl->status != 0) return;
  const double2* psi = ctl->cur ? psi1 : psi0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) old_sq[i] = psi[i].x * psi[i].x + psi[i].y * psi[i].y;
}

// Link variables for A = A_applied + A_induced (solver.py:670-673): theta holds A_applied . d
// (times ramp_f under a device-side ramp), aind the current induced potential.
__global__ void __launch_bounds__(kBlock)
k_link_values_scr(const Ctl* __restrict__ ctl, int n, const int* __restrict__ ptr,
                  const int* __restrict__ eidx, const signed char* __restrict__ head,
                  const double* __restrict__ weight, const double* __restrict__ theta,
                  const double2* __restrict__ aind, const double2* __restrict__ edir,
                  const double* __restrict__ areas, double2* __restrict__ lval) {
  griddep_enter();
  if (ctl->status != 0) return;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double f = ctl->ramp_on ? ctl->ramp_f : 1.0;
  double diag = 0.0;
  int kd = -1;
  for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
    const int e = eidx[k];
    if (e < 0) { kd = k; continue; }
    const double w = weight[e];
    const double th = f * theta[e] + aind[e].x * edir[e].x + aind[e].y * edir[e].y;
    double s, c;
    sincos(-th, &s, &c);
    if (!head[k]) s = -s;
    lval[k] = make_double2(w * c / areas[row], w * s / areas[row]);
    diag += -w / areas[row];
  }
  if (kd >= 0) lval[kd] = make_double2(diag, 0.0);
}

// w_i = a~_i * J_site[i], J_site[i] = (1/2) mean over the edges at site i of (J_s + J_n)[e] e_hat
// (Mesh.get_quantity_on_site, mesh.py:203-243).  One thread per site walks its incident edges
// and evaluates each edge's current on the fly (operators.py:385-394, solver.py:519): no
// edge-sized temporary, fixed summation order.
__global__ void __launch_bounds__(kBlock)
k_scr_site_current(const Ctl* __restrict__ ctl, int n, const int* __restrict__ ptr,
                   const int* __restrict__ nbr, const int* __restrict__ eidx,
                   const signed char* __restrict__ head, const double* __restrict__ elen,
                   const double* __restrict__ theta, const double2* __restrict__ aind,
                   const double2* __restrict__ edir, const double2* __restrict__ psi0,
                   const double2* __restrict__ psi1, const double* __restrict__ mu,
                   const double* __restrict__ dadt /* may be null */,
                   const double* __restrict__ ramp_proj /* may be null */,
                   const double* __restrict__ scr_area, double2* __restrict__ wsite) {
  griddep_enter();
  if (ctl->status != 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2* psi = ctl->cur ? psi1 : psi0;
  const double f = ctl->ramp_on ? ctl->ramp_f : 1.0;
  double sx = 0.0, sy = 0.0;
  int cnt = 0;
  for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
    const int e = eidx[k];
    if (e < 0) continue;
    const int j = nbr[k];
    const int a = head[k] ? i : j, b = head[k] ? j : i;   // the edge runs a -> b
    const double inv_l = 1.0 / elen[e];
    const double th = f * theta[e] + aind[e].x * edir[e].x + aind[e].y * edir[e].y;
    double s, c;
    sincos(-th, &s, &c);
    const double2 pa = psi[a], pb = psi[b];
    const double gx = (c * pb.x - s * pb.y) * inv_l - pa.x * inv_l;
    const double gy = (c * pb.y + s * pb.x) * inv_l - pa.y * inv_l;
    const double js = pa.x * gy - pa.y * gx;
    double da = dadt != nullptr ? dadt[e] : 0.0;
    if (ramp_proj != nullptr) da = ctl->ramp_dfdt * ramp_proj[e];
    const double jn = -(mu[b] * inv_l - mu[a] * inv_l) - da;
    const double J = js + jn;
    const double dl = sqrt(edir[e].x * edir[e].x + edir[e].y * edir[e].y);
    sx += J * (edir[e].x / dl);
    sy += J * (edir[e].y / dl);
    ++cnt;
  }
  const double sc = cnt > 0 ? scr_area[i] / (2.0 * cnt) : 0.0;
  wsite[i] = make_double2(sx * sc, sy * sc);
}

// A_new[e] = sum_j w_j / |c_e - r_j|: all-pairs sum, tiled through shared memory (one tile of
// kScrTile sites = positions + weights, 32 bytes per site, is read once per CTA and used by
// every edge of the CTA).  Each thread owns one edge and adds the sites in index order, so the
// result does not depend on the launch geometry.  fp64 compute-bound (sqrt + divide per
// pair): the reference's algorithm is O(E N) and so is this.
constexpr int kScrTile = 128;
__global__ void __launch_bounds__(kScrTile)
k_scr_a_induced(const Ctl* __restrict__ ctl, int n_edges, int n_sites,
                const double2* __restrict__ ecent, const double2* __restrict__ sxy,
                const double2* __restrict__ wsite, double2* __restrict__ a_new) {
  griddep_enter();
  __shared__ double2 s_r[kScrTile];
  __shared__ double2 s_w[kScrTile];
  if (ctl->status != 0) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const double2 c = e < n_edges ? ecent[e] : make_double2(0.0, 0.0);
  double ax = 0.0, ay = 0.0;
  for (int j0 = 0; j0 < n_sites; j0 += kScrTile) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    s_r[threadIdx.x] = j < n_sites ? sxy[j] : make_double2(1e300, 1e300);
    s_w[threadIdx.x] = j < n_sites ? wsite[j] : make_double2(0.0, 0.0);
    __syncthreads();
    const int m = min(kScrTile, n_sites - j0);
#pragma unroll 4
    for (int t = 0; t < m; ++t) {
      const double dx = c.x - s_r[t].x, dy = c.y - s_r[t].y;
      const double inv = 1.0 / sqrt(dx * dx + dy * dy);
      ax = fma(s_w[t].x, inv, ax);
      ay = fma(s_w[t].y, inv, ay);
    }
  }
  if (e < n_edges) a_new[e] = make_double2(ax, ay);
}

// Polyak update (solver.py:564-577): dA = A_new - A ; v = (1 - beta) v + alpha dA ; A += v ;
// error = max_e |dA_e| / max(|A_e|, 1e-20).  The velocity starts from 0 in every time step.
// a_new is left holding the OLD A: the potential the link variables of this pass were built
// from, i.e. the one the supercurrent the reference returns belongs to (solver.py:670-681).
__global__ void __launch_bounds__(kBlock)
k_scr_polyak(Ctl* ctl, int n_edges, double2* __restrict__ a_new, double2* __restrict__ aind,
             double2* __restrict__ vel) {
  griddep_enter();
  __shared__ double s_max[8];
  if (ctl->status != 0) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool first = ctl->scr_it == 0;
  double err = 0.0;
  if (e < n_edges) {
    const double2 a = aind[e], an = a_new[e];
    const double2 v0 = first ? make_double2(0.0, 0.0) : vel[e];
    const double dx = an.x - a.x, dy = an.y - a.y;
    const double2 v = make_double2((1.0 - ctl->scr_beta) * v0.x + ctl->scr_alpha * dx,
                                   (1.0 - ctl->scr_beta) * v0.y + ctl->scr_alpha * dy);
    const double2 a1 = make_double2(a.x + v.x, a.y + v.y);
    vel[e] = v;
    aind[e] = a1;
    a_new[e] = a;
    err = sqrt(dx * dx + dy * dy) / fmax(sqrt(a1.x * a1.x + a1.y * a1.y), 1e-20);
    if (!(err == err)) err = 1e300;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    double mx = s_max[0];
    for (int k = 1; k < static_cast<int>(blockDim.x >> 5); ++k) mx = fmax(mx, s_max[k]);
    if (mx > 0.0) atomicMax(&ctl->scr_err_bits, (unsigned long long)__double_as_longlong(mx));
  }
}

// Loop control of the screening iteration (solver.py:654-663): stop when the error is below
// the tolerance; give up (status 4) after max_iterations_per_step passes.  Re-arms the
// per-pass state of the psi step (adaptive_euler_step starts with retries = 0 every pass).
__global__ void k_scr_control(Ctl* ctl, cudaGraphConditionalHandle cond_scr,
                              cudaGraphConditionalHandle cond_psi) {
  griddep_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int go = 0;
  if (ctl->status == 0) {
    const double err = __longlong_as_double((long long)ctl->scr_err_bits);
    ctl->scr_err = err;
    ctl->scr_it += 1;
    ctl->total_scr_it += 1;
    if (!(err < ctl->scr_tol)) {
      if (ctl->scr_it > ctl->scr_max_it) {
        ctl->status = 4;
        ctl->failed_step = ctl->step;
        ctl->failed_dt = ctl->dt;
      } else {
        go = 1;
        ctl->scr_err_bits = 0ull;
        ctl->retries = 0;
        ctl->disc_flag = 0;
        ctl->max_dpsi_bits = 0ull;
        ctl->solve_epoch += 1;
      }
    }
  }
  ctl->scr_go = go;
  ctl->psi_go = go;
  set_cond(cond_scr, go);
  set_cond(cond_psi, go);
}

template <typename T>
__global__ void k_gather(int n, const int* __restrict__ perm, const T* __restrict__ src,
                         T* __restrict__ dst) {  // dst[i] = src[perm[i]]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

template <typename T>
__global__ void k_scatter(int n, const int* __restrict__ perm, const T* __restrict__ src,
                          T* __restrict__ dst) {  // dst[perm[i]] = src[i]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i];
}

// dst[perm[i]] = (current psi buffer)[i]: for launches enqueued before the host knows which
// buffer the device-side psi loop will accept into
__global__ void k_scatter_psi(const Ctl* __restrict__ ctl, int n, const int* __restrict__ perm,
                              const double2* __restrict__ psi_buf0,
                              const double2* __restrict__ psi_buf1, double2* __restrict__ dst) {
  const double2* src = ctl->cur ? psi_buf1 : psi_buf0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i];
}

__global__ void k_scale_neg_area(int n, const double* __restrict__ areas,
                                 const double* __restrict__ rhs, double* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = -areas[i] * rhs[i];
}

// *out = dot(a, b)   (deterministic)
__global__ void __launch_bounds__(kBlock)
k_dot(Ctl* ctl, Comm* comm, int n, const double* __restrict__ a, const double* __restrict__ b,
      double* partials, unsigned int* counter, double* out) {
  griddep_enter();
  __shared__ double red[32];
  if (ctl->status != 0) return;
  double d = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    d += a[i] * b[i];
  const double bs = block_sum(d, red);
  double total;
  if (grid_sum_last(bs, partials, counter, red, &total) && threadIdx.x < 32) {
    total = __shfl_sync(0xffffffffu, total, 0);
    if (comm != nullptr) comm_allreduce(ctl, comm, &total, 1, false);
    if (threadIdx.x == 0) *out = total;
  }
}

// ---- shard-local step seam (sharded engine) ---------------------------------------------------
// All ranks have finished reading the mailboxes of the previous step (one all-reduce).
__global__ void k_comm_barrier(Ctl* ctl, Comm* comm) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  double v = 0.0;
  comm_allreduce(ctl, comm, &v, 1, false);
}

// Does the state the caller hands in equal the state the engine already holds (bit for bit)?
// The reference's Runner feeds every step's output back as the next step's input
// (runner.py:417-423): then nothing has to be replaced, and the history of the mu solve's
// initial guess and the mailboxes stay valid.  *differ is set when any word differs.
__global__ void __launch_bounds__(kBlock)
k_state_differs(int n, const double2* __restrict__ psi_in, const double2* psi_buf0,
                const double2* psi_buf1, const Ctl* __restrict__ ctl, const double* __restrict__ mu_in,
                const double* __restrict__ mu, int* differ) {
  const double2* __restrict__ psi = ctl->cur ? psi_buf1 : psi_buf0;
  bool d = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 a = psi_in[i], b = psi[i];
    d |= __double_as_longlong(a.x) != __double_as_longlong(b.x) ||
         __double_as_longlong(a.y) != __double_as_longlong(b.y) ||
         __double_as_longlong(mu_in[i]) != __double_as_longlong(mu[i]);
  }
  if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0) atomicOr(differ, 1);
}
// ... on ANY rank (all ranks must take the same path): max over the ranks, left in *differ.
__global__ void k_state_vote(Ctl* ctl, Comm* comm, int* differ) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  double v = static_cast<double>(*differ);
  if (comm != nullptr) comm_allreduce(ctl, comm, &v, 1, true);
  if (threadIdx.x == 0) *differ = v != 0.0 ? 1 : 0;
}

// The boundary rows of a state vector this rank has just been handed go to the neighbours'
// mailboxes (the counterpart of k_fill_box when every rank only holds its own rows).
template <typename T>
__global__ void __launch_bounds__(kBlock)
k_push_state(const Comm* comm, PushArgs push, unsigned int tag, int n, const T* __restrict__ vec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) push_row(comm, push, tag, i, vec[i]);
}

// Reads a buffer larger than L2 so that the next kernel starts from a cold cache
// (microbenchmarks only).
__global__ void k_flush_l2(const double* __restrict__ buf, size_t n, double* sink) {
  double s = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    s += buf[i];
  if (s == 123.456) *sink = s;
}

// y = -y / areas   (A -> mu_laplacian)
__global__ void k_neg_div(int n, const double* __restrict__ areas, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = -y[i] / areas[i];
}

}  // namespace tdgl
