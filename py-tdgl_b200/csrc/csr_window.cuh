// CSR row-window kernels (sm_100a): every sparse operator of the step streams its CSR
// arrays through shared memory with the TMA engine and does the per-row arithmetic there.
//
//   one CTA  = one window of blockDim.x consecutive rows, one thread per row
//   thread 0 = reads the two row pointers that bound the window and issues one
//              cp.async.bulk (global -> shared, mbarrier complete_tx) per CSR array for
//              the window's whole nnz extent: the matrix moves in a few multi-KB
//              transactions that need no registers, several CTAs per SM keep > 100 KB in
//              flight per SM, which is what saturating HBM3e takes;
//   others   = meanwhile fetch their own row pointers and row-local vector entries;
//   then     = every thread walks its row out of shared memory, gathers x[col] through
//              L1/L2 (sites are numbered along a Z-order curve, so a window's columns are
//              mostly the window itself) and applies the operator's epilogue.
//
// The CSR arrays are static for a whole solve; only the vectors change.  All streaming
// traffic of a kernel is therefore issued BEFORE griddep_wait(): under programmatic
// dependent launch the next kernel's matrix windows are already in flight while the
// previous kernel drains.
//
// Alignment: cp.async.bulk needs 16-byte aligned addresses and sizes, so a window stages
// the nnz range [ptr[r0] & ~3, (ptr[r1] + 3) & ~3); the arrays carry 4 padding elements.
#pragma once

#include "kernels.cuh"

namespace tdgl {

constexpr int kWinRows = 256;  // default rows per window (= threads per CTA)

struct WinCsr {
  int rows = 0;
  int cap = 0;               // max over windows of the aligned nnz extent (elements)
  const int* ptr = nullptr;  // rows + 1
  const int* idx = nullptr;  // nnz (+4 padding)
  // per window {first staged nnz (multiple of 4), staged nnz count}: 8 bytes per CTA that stay
  // L2-resident across launches, so that the bulk copies of a window are issued after one
  // L2 hit instead of after a DRAM round trip for the two row pointers bounding it
  const int2* win = nullptr;
};

// ---- PTX wrappers ----------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- staging ---------------------------------------------------------------------------------

struct WinRow {
  int row;     // global row of this thread (may be >= rows)
  int kb, ke;  // extent of the row inside the staged arrays
};

// Issues the window's bulk copies and returns this thread's row extent.  SA / SB are the
// element sizes of up to two value arrays that share the matrix structure (SB = 0: one).
// Shared layout: [valA: cap*SA][valB: cap*SB][idx: cap*4].  Ends with __syncthreads().
// LPR = lanes per row: 1 (a thread per row) or, for matrices with long rows, 4 (blockDim / 4
// rows per window; the lanes of a row split its entries).
template <int SA, int SB, int LPR = 1>
__device__ __forceinline__ WinRow window_stage(const WinCsr& m, const void* valA,
                                               const void* valB, unsigned char* smem,
                                               uint64_t* bar) {
  griddep_launch_dependents();
  const int r0 = blockIdx.x * (blockDim.x / LPR);
  const int2 wd = __ldg(m.win + blockIdx.x);
  const int k0a = wd.x;
  if (threadIdx.x == 0) {
    const uint32_t n = static_cast<uint32_t>(wd.y);
    mbar_init(bar, 1);
    mbar_arrive_expect_tx(bar, n * (SA + SB + 4));
    if (n > 0) {
      bulk_g2s(smem, static_cast<const unsigned char*>(valA) + static_cast<size_t>(k0a) * SA,
               n * SA, bar);
      if (SB > 0)
        bulk_g2s(smem + static_cast<size_t>(m.cap) * SA,
                 static_cast<const unsigned char*>(valB) + static_cast<size_t>(k0a) * SB, n * SB,
                 bar);
      bulk_g2s(smem + static_cast<size_t>(m.cap) * (SA + SB), m.idx + k0a, n * 4, bar);
    }
  }
  WinRow w;
  w.row = r0 + threadIdx.x / LPR;
  w.kb = w.ke = 0;
  if (w.row < m.rows) {
    w.kb = __ldg(m.ptr + w.row) - k0a;
    w.ke = __ldg(m.ptr + w.row + 1) - k0a;
  }
  __syncthreads();  // the initialised barrier is visible to every waiter
  return w;
}

// sum_k val[k] * x[idx[k]] over the staged row, four gathers in flight per thread, added in
// column order.
// Halo columns — sharded engine only — come out of the mailbox (comm.cuh).  The gathers are
// issued unconditionally from the array (index clamped), so that they stay independent and
// in flight together; the rare halo entries are patched afterwards.
template <bool SH>
__device__ __forceinline__ double row_dot(const double* __restrict__ sv,
                                          const int* __restrict__ si, int kb, int ke,
                                          const double* __restrict__ x, Ctl* ctl,
                                          const HaloView& h) {
  double s = 0.0;
  const int last = ke - 1;
  const int top = SH ? h.n_owned - 1 : 0x7fffffff;
#pragma unroll 2
  for (int k = kb; k < ke; k += 4) {
    const int k1 = min(k + 1, last), k2 = min(k + 2, last), k3 = min(k + 3, last);
    const int j0 = si[k], j1 = si[k1], j2 = si[k2], j3 = si[k3];
    double x0 = __ldg(x + min(j0, top)), x1 = __ldg(x + min(j1, top)), x2 = __ldg(x + min(j2, top)),
           x3 = __ldg(x + min(j3, top));
    if (SH && max(max(j0, j1), max(j2, j3)) > top) {
      if (j0 > top) x0 = halo_get(ctl, h, x, j0);
      if (j1 > top) x1 = halo_get(ctl, h, x, j1);
      if (j2 > top) x2 = halo_get(ctl, h, x, j2);
      if (j3 > top) x3 = halo_get(ctl, h, x, j3);
    }
    const double v0 = sv[k], v1 = (k + 1 < ke) ? sv[k1] : 0.0, v2 = (k + 2 < ke) ? sv[k2] : 0.0,
                 v3 = (k + 3 < ke) ? sv[k3] : 0.0;
    s = fma(v0, x0, s);
    s = fma(v1, x1, s);
    s = fma(v2, x2, s);
    s = fma(v3, x3, s);
  }
  return s;
}

// complex: sum_k val[k] * x[idx[k]]
template <bool SH>
__device__ __forceinline__ double2 row_dot_c(const double2* __restrict__ sv,
                                             const int* __restrict__ si, int kb, int ke,
                                             const double2* __restrict__ x, Ctl* ctl,
                                             const HaloView& h) {
  double sx = 0.0, sy = 0.0;
  const int last = ke - 1;
  const int top = SH ? h.n_owned - 1 : 0x7fffffff;
#pragma unroll 2
  for (int k = kb; k < ke; k += 2) {
    const int k1 = min(k + 1, last);
    const int j0 = si[k], j1 = si[k1];
    double2 x0 = __ldg(x + min(j0, top)), x1 = __ldg(x + min(j1, top));
    if (SH && max(j0, j1) > top) {
      if (j0 > top) x0 = halo_get(ctl, h, x, j0);
      if (j1 > top) x1 = halo_get(ctl, h, x, j1);
    }
    const double2 v0 = sv[k];
    double2 v1 = sv[k1];
    if (k + 1 >= ke) v1 = make_double2(0.0, 0.0);
    sx += v0.x * x0.x - v0.y * x0.y;
    sy += v0.x * x0.y + v0.y * x0.x;
    sx += v1.x * x1.x - v1.y * x1.y;
    sy += v1.x * x1.y + v1.y * x1.x;
  }
  return make_double2(sx, sy);
}

// ---- real operators (mu system, AMG levels) ----------------------------------------------------

enum : int { kOpSpmvDot = 0, kOpResidual, kOpPresmooth, kOpJacobi, kOpPlain, kOpPlainAdd,
             kOpSpmvCg };

struct RealArgs {
  const double* val = nullptr;
  const double* x = nullptr;     // gathered vector
  double* y = nullptr;           // row output
  const double* b = nullptr;     // right-hand side (residual / smoothers)
  const double* dinv = nullptr;  // 1 / diag (smoothers)
  const double* w = nullptr;     // kOpJacobi: dot(w, y) -> *red_out
  double* r = nullptr;           // kOpPresmooth: residual output
  double omega = 0.0;
  double* red_out = nullptr;     // reduction result (deterministic), may be null
  double* red2_out = nullptr;    // kOpSpmvCg: the second sum
  // sharded engine: where the halo columns of the gathered vector (x; b for kOpPresmooth)
  // are found, and where the boundary rows of the output (r for kOpPresmooth, y otherwise)
  // go.  Defaults: no halo, nothing to send.
  HaloArgs halo;
  PushArgs push;
};

// psi is double-buffered, and so are its mailboxes: [b] belongs to buffer b
struct PsiComm {
  HaloArgs halo[2];
  PushArgs push[2];
};

//  kOpSpmvDot   y = A x ;                         red = dot(x, y)
//  kOpSpmvCg    y = A x ;                         red = dot(b, x), red2 = dot(x, y)
//               (the CG iteration's SpMV: x = z, b = r -> gamma = r.z and delta = z.Az)
//  kOpResidual  y = b - A x ;                     red = ||y||^2
//  kOpPresmooth y = omega D^-1 b ; r = b - A y    (smoothing from a zero guess + residual)
//  kOpJacobi    y = x + omega D^-1 (b - A x) ;    red = dot(w, y)
//  kOpPlain     y = A x ;  kOpPlainAdd  y += A x  (restriction / prolongation)
// SH: instantiated for the sharded engine (halo columns out of the mailbox, boundary rows
// pushed to the peers); the single-GPU instantiation carries none of that code.
// LPR: lanes per row.  1 = a thread per row (the mesh operators, ~7 entries per row); 4 = four
// lanes share a row, entries k, k + 4, ... each (the AMG operators: rows of 15-30 entries and
// few of them — a quarter of the dependent-gather chain per thread, four times the gathers in
// flight; measured 25.8 -> ~12 us for the fine-level restriction).
template <int OP, bool SH, int LPR>
__global__ void __launch_bounds__(kWinRows, 8)   // 8 CTAs/SM = 2048 threads: <= 32 registers
kw_real(Ctl* ctl, Comm* comm, WinCsr m, RealArgs a, double* partials, unsigned int* counter) {
  extern __shared__ __align__(128) unsigned char win_smem[];
  __shared__ uint64_t bar;
  __shared__ double red[32];
  const WinRow w = window_stage<8, 0, LPR>(m, a.val, nullptr, win_smem, &bar);
  const double* sv = reinterpret_cast<const double*>(win_smem);
  const int* si = reinterpret_cast<const int*>(win_smem + static_cast<size_t>(m.cap) * 8);
  griddep_wait();
  const bool live = (ctl->status == 0);
  const bool in = live && w.row < m.rows;
  const int sub = threadIdx.x & (LPR - 1);   // lane of the row (LPR lanes share a long row)
  const bool lead = in && sub == 0;          // the lane that owns the row's epilogue
  HaloView hv;
  hv.n_owned = 0x7fffffff; hv.box = nullptr; hv.tag = 0u;
  unsigned int tag_out = 0u;
  if (SH) {
    hv = halo_view(ctl, comm, a.halo);
    if (comm != nullptr && a.push.bnd != nullptr) tag_out = comm_tag(ctl, a.push.tag_mode);
  }
  // row-local operands travel while the window lands
  double bi = 0.0, di = 0.0, xi = 0.0, wi = 0.0;
  if (lead) {
    if (OP == kOpResidual || OP == kOpPresmooth || OP == kOpJacobi || OP == kOpSpmvCg)
      bi = a.b[w.row];
    if (OP == kOpPresmooth || OP == kOpJacobi) di = a.dinv[w.row];
    if (OP == kOpSpmvDot || OP == kOpJacobi || OP == kOpSpmvCg) xi = a.x[w.row];
    if (OP == kOpPlainAdd) xi = a.y[w.row];
    if (OP == kOpJacobi && a.w != nullptr) wi = a.w[w.row];
  }
  mbar_wait(&bar, 0);
  if (!live) return;
  double d = 0.0, d2 = 0.0;
  double s = 0.0;
  if (in) {
    const int top = SH ? hv.n_owned - 1 : 0x7fffffff;
    if (OP == kOpPresmooth) {
      // x_j = omega dinv_j b_j on the fly
      const int last = w.ke - 1;
      for (int k = w.kb + sub; k < w.ke; k += 2 * LPR) {
        const int k1 = min(k + LPR, last);
        const bool two = k + LPR < w.ke;
        const int j0 = si[k], j1 = si[k1];
        double b0 = __ldg(a.b + min(j0, top)), b1 = __ldg(a.b + min(j1, top));
        if (SH && max(j0, j1) > top) {
          if (j0 > top) b0 = halo_get(ctl, hv, a.b, j0);
          if (j1 > top) b1 = halo_get(ctl, hv, a.b, j1);
        }
        const double t0 = __ldg(a.dinv + j0) * b0;
        const double t1 = __ldg(a.dinv + j1) * b1;
        s = fma(sv[k], a.omega * t0, s);
        s = fma(two ? sv[k1] : 0.0, a.omega * t1, s);
      }
    } else if (LPR == 1) {
      s = row_dot<SH>(sv, si, w.kb, w.ke, a.x, ctl, hv);
    } else {
#pragma unroll 2
      for (int k = w.kb + sub; k < w.ke; k += 2 * LPR) {
        const int k1 = k + LPR;
        const bool two = k1 < w.ke;
        const int j0 = si[k], j1 = two ? si[k1] : j0;
        double x0 = __ldg(a.x + min(j0, top)), x1 = __ldg(a.x + min(j1, top));
        if (SH && max(j0, j1) > top) {
          if (j0 > top) x0 = halo_get(ctl, hv, a.x, j0);
          if (j1 > top) x1 = halo_get(ctl, hv, a.x, j1);
        }
        s = fma(sv[k], x0, s);
        s = fma(two ? sv[k1] : 0.0, x1, s);
      }
    }
  }
  if (LPR > 1) s = group_sum<LPR>(s);
  if (lead) {
    double out;  // the value other shards may need
    if (OP == kOpSpmvDot) {
      a.y[w.row] = out = s;
      d = s * xi;
    } else if (OP == kOpSpmvCg) {
      a.y[w.row] = out = s;
      d = bi * xi;
      d2 = s * xi;
    } else if (OP == kOpResidual) {
      const double ri = bi - s;
      a.y[w.row] = out = ri;
      d = ri * ri;
    } else if (OP == kOpPresmooth) {
      a.y[w.row] = a.omega * di * bi;
      a.r[w.row] = out = bi - s;
    } else if (OP == kOpJacobi) {
      const double yi = xi + a.omega * di * (bi - s);
      a.y[w.row] = out = yi;
      d = wi * yi;
    } else if (OP == kOpPlain) {
      a.y[w.row] = out = s;
    } else {
      a.y[w.row] = out = xi + s;
    }
    if (SH && tag_out != 0u) push_row(comm, a.push, tag_out, w.row, out);
    (void)out;
  }
  if (OP == kOpSpmvCg) {
    const double b0 = block_sum(d, red);
    const double b1 = block_sum(d2, red);
    double t0, t1;
    if (grid_sum2_last(b0, b1, partials, counter, red, &t0, &t1) && threadIdx.x < 32) {
      double v[2];
      v[0] = __shfl_sync(0xffffffffu, t0, 0);
      v[1] = __shfl_sync(0xffffffffu, t1, 0);
      if (SH && comm != nullptr) comm_allreduce(ctl, comm, v, 2, false);
      if (threadIdx.x == 0) {
        *a.red_out = v[0];
        *a.red2_out = v[1];
      }
    }
  }
  if ((OP == kOpSpmvDot || OP == kOpResidual || OP == kOpJacobi) && a.red_out != nullptr) {
    const double bs = block_sum(d, red);
    double total;
    if (grid_sum_last(bs, partials, counter, red, &total) && threadIdx.x < 32) {
      total = __shfl_sync(0xffffffffu, total, 0);
      if (SH && comm != nullptr) comm_allreduce(ctl, comm, &total, 1, false);
      if (threadIdx.x == 0) *a.red_out = total;
    }
  }
}

// ---- psi step ----------------------------------------------------------------------------------
// Fused covariant-Laplacian SpMV + closed-form |psi|^2 update (reference
// TDGLSolver.solve_for_psi_squared, tdgl/solver/solver.py:418-438); fixed[i] != 0 marks rows
// the reference replaces by the identity (operators.py:170-184): there (L psi)_i = psi_i.
template <bool SH>
__global__ void __launch_bounds__(kWinRows)
kw_psi_step(Ctl* ctl, const Comm* comm, PsiComm pc, WinCsr m, const double2* __restrict__ lval,
            const unsigned char* __restrict__ fixed, const double2* psi_buf0,
            const double2* psi_buf1, double2* out_buf0, double2* out_buf1,
            const double* __restrict__ mu, const double* __restrict__ eps,
            double* __restrict__ sq_out /* may be null */, double dt_override /* < 0: ctl->dt */,
            const double* __restrict__ old_sq /* screening: |psi|^2 of the step's input; else null */) {
  extern __shared__ __align__(128) unsigned char win_smem[];
  __shared__ uint64_t bar;
  __shared__ double s_max[8];
  __shared__ int s_flag;
  const WinRow w = window_stage<16, 0>(m, lval, nullptr, win_smem, &bar);
  const double2* sv = reinterpret_cast<const double2*>(win_smem);
  const int* si = reinterpret_cast<const int*>(win_smem + static_cast<size_t>(m.cap) * 16);
  if (threadIdx.x == 0) s_flag = 0;
  griddep_wait();
  const bool live = (ctl->status == 0);
  const int cur = ctl->cur;
  const double2* __restrict__ psi = cur ? psi_buf1 : psi_buf0;
  double2* __restrict__ out = cur ? out_buf0 : out_buf1;
  const double dt = dt_override >= 0.0 ? dt_override : ctl->dt;
  const bool in = live && w.row < m.rows;
  // sharded: halo columns of the current psi come out of its buffer's mailbox, the boundary
  // rows of the new psi go into the other buffer's mailbox on the neighbours
  HaloView hv;
  hv.n_owned = 0x7fffffff; hv.box = nullptr; hv.tag = 0u;
  unsigned int tag_out = 0u;
  if (SH && comm != nullptr) {
    hv = halo_view(ctl, comm, pc.halo[cur]);
    tag_out = comm_tag(ctl, kTagPsiNew);
  }
  double2 p = make_double2(0.0, 0.0);
  double mui = 0.0, epsi = 0.0;
  bool fx = false;
  if (in) {
    p = psi[w.row];
    mui = mu[w.row];
    epsi = eps[w.row];
    fx = fixed[w.row] != 0;
  }
  double abs2 = p.x * p.x + p.y * p.y;
  if (in && old_sq != nullptr) abs2 = old_sq[w.row];
  mbar_wait(&bar, 0);
  if (!live) return;
  double dmax = 0.0;
  int failed = 0;
  if (in) {
    double2 lap = row_dot_c<SH>(sv, si, w.kb, w.ke, psi, ctl, hv);
    if (fx) lap = p;
    const PsiOut o = psi_update(p, lap, mui, epsi, ctl->gamma, ctl->u, dt, abs2);
    out[w.row] = o.psi;
    if (SH && comm != nullptr) push_row(comm, pc.push[cur ^ 1], tag_out, w.row, o.psi);
    if (sq_out != nullptr) sq_out[w.row] = o.sq;
    failed = o.failed;
    const double d = fabs(o.sq - abs2);
    dmax = (d == d) ? d : 0.0;
  }
  // max / any are order-free: warp shuffle, then one atomic per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    failed |= __shfl_xor_sync(0xffffffffu, failed, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_max[warp] = dmax;
    if (failed) atomicOr(&s_flag, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mx = s_max[0];
    for (int k = 1; k < static_cast<int>(blockDim.x >> 5); ++k) mx = fmax(mx, s_max[k]);
    if (mx > 0.0) atomicMax(&ctl->max_dpsi_bits, (unsigned long long)__double_as_longlong(mx));
    if (s_flag) atomicOr(&ctl->disc_flag, 1);
  }
}

// ---- right-hand side of the mu system ------------------------------------------------------------
//   rhs_i = (divergence @ J_s)_i - (mu_boundary_laplacian @ mu_boundary)_i
//         = Im(conj(psi_i) (L~ psi)_i) - bterm_i          (L~: Laplacian without fixed rows)
//   b_i   = -areas_i * rhs_i ;   r_i = b_i - (A mu)_i ;  bb = ||b||^2, rr = ||r||^2
// (reference solve_for_observables, solver.py:507-510; identity: SURVEY.md appendix A).
// The complex and the real matrix share one CSR structure and are staged together.
template <bool SH>
__global__ void __launch_bounds__(kWinRows)
kw_mu_rhs(Ctl* ctl, Comm* comm, PsiComm pc, HaloArgs mu_halo, HaloArgs mu_prev_halo, WinCsr m,
          const double2* __restrict__ lval, const double* __restrict__ aval,
          const double2* psi_buf0, const double2* psi_buf1, const double* __restrict__ mu,
          const double* __restrict__ mu_prev, const double* __restrict__ mu_pp /* plain halo */,
          double* __restrict__ d_out, double* __restrict__ d2_out,
          const double* __restrict__ areas, const double* __restrict__ bterm,
          const double* __restrict__ ramp_div /* null: no device-side ramp */,
          double* __restrict__ b, double* __restrict__ r,
          double* __restrict__ rhs_raw /* may be null: un-symmetrised rhs */, double* partials,
          unsigned int* counter) {
  extern __shared__ __align__(128) unsigned char win_smem[];
  __shared__ uint64_t bar;
  __shared__ double red[32];
  __shared__ int s_last;
  const WinRow w = window_stage<16, 8>(m, lval, aval, win_smem, &bar);
  const double2* sl = reinterpret_cast<const double2*>(win_smem);
  const double* sa = reinterpret_cast<const double*>(win_smem + static_cast<size_t>(m.cap) * 16);
  const int* si = reinterpret_cast<const int*>(win_smem + static_cast<size_t>(m.cap) * 24);
  griddep_wait();
  const bool live = (ctl->status == 0);
  const double2* __restrict__ psi = ctl->cur ? psi_buf1 : psi_buf0;
  const bool in = live && w.row < m.rows;
  HaloView hpsi, hmu, hmp;
  hpsi.n_owned = hmu.n_owned = hmp.n_owned = 0x7fffffff;
  hpsi.box = hmu.box = hmp.box = nullptr;
  hpsi.tag = hmu.tag = hmp.tag = 0u;
  if (SH && comm != nullptr) {
    hpsi = halo_view(ctl, comm, pc.halo[ctl->cur]);
    hmu = halo_view(ctl, comm, mu_halo);
    hmp = halo_view(ctl, comm, mu_prev_halo);   // (the other parity buffer of mu's mailbox)
  }
  double2 p = make_double2(0.0, 0.0);
  double ai = 0.0, bt = 0.0;
  if (in) {
    p = psi[w.row];
    ai = areas[w.row];
    bt = bterm[w.row];
    // device-side ramp: + divergence @ dA_dt with dA_dt = ramp_dfdt * (A0 . e_hat)
    if (ramp_div != nullptr) bt += ctl->ramp_dfdt * ramp_div[w.row];
  }
  mbar_wait(&bar, 0);
  if (!live) return;
  // Initial guess of the solve: mu_g = mu + c1 (mu - mu_prev) + c2 (mu - mu_pp) (mu, mu_prev,
  // mu_pp: the last three solutions).  Its residual is r + c1 d1 + c2 d2 with r = b - A mu,
  // d1 = A mu_prev - A mu, d2 = A mu_pp - A mu, so the c1, c2 that minimise it follow from
  // seven dot products (k_cg_begin) — never worse than the plain warm start (c = 0), and
  // several CG iterations cheaper while the dynamics are smooth.
  double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (in) {
    const double2 lap = row_dot_c<SH>(sl, si, w.kb, w.ke, psi, ctl, hpsi);
    const double am = row_dot<SH>(sa, si, w.kb, w.ke, mu, ctl, hmu);
    const double amp = row_dot<SH>(sa, si, w.kb, w.ke, mu_prev, ctl, hmp);
    HaloView plain;
    plain.n_owned = 0x7fffffff; plain.box = nullptr; plain.tag = 0u;
    const double ampp = row_dot<false>(sa, si, w.kb, w.ke, mu_pp, ctl, plain);
    const double rhs = (p.x * lap.y - p.y * lap.x) - bt;
    if (rhs_raw != nullptr) rhs_raw[w.row] = rhs;
    const double bi = -ai * rhs;
    const double ri = bi - am;
    const double d1 = amp - am, d2 = ampp - am;
    b[w.row] = bi;
    r[w.row] = ri;
    d_out[w.row] = d1;
    d2_out[w.row] = d2;
    acc[0] = bi * bi;
    acc[1] = ri * ri;
    acc[2] = ri * d1;
    acc[3] = d1 * d1;
    acc[4] = ri * d2;
    acc[5] = d1 * d2;
    acc[6] = d2 * d2;
  }
  // seven sums through one deterministic reduction
  double v[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = block_sum(acc[k], red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) partials[8 * blockIdx.x + k] = v[k];
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x)
#pragma unroll
      for (int k = 0; k < 7; ++k) a[k] += reinterpret_cast<volatile double*>(partials)[8 * i + k];
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] = block_sum(a[k], red);
    if (threadIdx.x < 32) {  // block_sum leaves the total in every lane of warp 0
      if (SH && comm != nullptr) {
        comm_allreduce(ctl, comm, a, 4, false);
        comm_allreduce(ctl, comm, a + 4, 4, false);
      }
      if (threadIdx.x == 0) {
        ctl->bb = a[0];
        ctl->rr = a[1];
        ctl->rd = a[2];
        ctl->dd = a[3];
        ctl->rd2 = a[4];
        ctl->d1d2 = a[5];
        ctl->d2d2 = a[6];
        *counter = 0u;
      }
    }
  }
}

// y = psi_laplacian @ x with the reference's fixed rows (identity)  — parity / microbench op
__global__ void __launch_bounds__(kWinRows)
kw_psi_laplacian(WinCsr m, const double2* __restrict__ lval,
                 const unsigned char* __restrict__ fixed, const double2* __restrict__ x,
                 double2* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char win_smem[];
  __shared__ uint64_t bar;
  const WinRow w = window_stage<16, 0>(m, lval, nullptr, win_smem, &bar);
  const double2* sv = reinterpret_cast<const double2*>(win_smem);
  const int* si = reinterpret_cast<const int*>(win_smem + static_cast<size_t>(m.cap) * 16);
  griddep_wait();
  mbar_wait(&bar, 0);
  if (w.row < m.rows) {
    HaloView none;
    none.n_owned = 0x7fffffff; none.box = nullptr; none.tag = 0;
    const double2 lap = row_dot_c<false>(sv, si, w.kb, w.ke, x, nullptr, none);
    y[w.row] = fixed[w.row] ? x[w.row] : lap;
  }
}

}  // namespace tdgl

namespace tdgl {

// ---- fused coarse levels -------------------------------------------------------------------------
// The coarse part of the V-cycle — every level with at most kFuseBelow rows, down to the dense
// coarsest solve and back up — as ONE kernel: a single thread-block cluster of 8 CTAs walks
// through the phases with hardware cluster barriers between them instead of one kernel launch
// per operator per level.  These levels hold ~0.2 % of the unknowns but, launched one by one,
// cost a third of the V-cycle's launches; their matrices and vectors stay L2-resident,
// so no shared-memory staging is needed — a phase is a few dependent L2 round trips.
//   down:  x = w D^-1 b ; r = b - A x ; b' = R r          (per level)
//   coarsest: y = Minv b (dense)
//   up:    x += P y' ; y = x + w D^-1 (b - A x)            (per level)
// Input: b of level `first`; output: y of level `first`.  In the sharded engine these levels
// are replicated on every shard (shard.h), so the kernel contains no exchange.

constexpr int kFuseBelow = 4096;    // rows: what one 8-CTA cluster turns around in a few microseconds
constexpr int kFuseCtas = 8;        // portable cluster size
constexpr int kFuseThreads = 1024;

struct FusedCsr {
  int rows = 0;
  const int* ptr = nullptr;
  const int* idx = nullptr;
  const double* val = nullptr;
};
struct FusedLevel {
  int n = 0;
  FusedCsr A, P, R;   // P: n x n_coarse, R: n_coarse x n
  const double* dinv = nullptr;
  double omega = 0.0;
  double *b = nullptr, *x = nullptr, *r = nullptr, *y = nullptr;
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One row = one group of 8 lanes (the rows of these levels are long, 15-30 entries, and there
// are few of them: spreading a row over lanes keeps the dependent-load chains short).
// `term(k)` returns the k-th product of the row.  All 32 lanes of a warp must call this.
template <typename F>
__device__ __forceinline__ double fused_row_sum(const FusedCsr& m, int row, bool live, int sub,
                                                F term) {
  double s = 0.0;
  if (live) {
    const int kb = __ldg(m.ptr + row), ke = __ldg(m.ptr + row + 1);
    for (int k = kb + sub; k < ke; k += 8) s += term(k);
  }
  return group_sum<8>(s);
}

__global__ void __cluster_dims__(kFuseCtas, 1, 1) __launch_bounds__(kFuseThreads)
k_coarse_cycle(const Ctl* __restrict__ ctl, const FusedLevel* __restrict__ lv, int first,
               int n_levels, const double* __restrict__ coarse_inv, int nc) {
  griddep_enter();
  if (ctl->status != 0) return;  // uniform over the cluster
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  const int grp = tid >> 3, sub = tid & 7, ngrp = nth >> 3;
  for (int l = first; l + 1 < n_levels; ++l) {
    const FusedLevel L = lv[l];
    // x = w D^-1 b ; r = b - A x  (x_j formed on the fly from b_j, as kw_real<presmooth>)
    for (int base = 0; base < L.n; base += ngrp) {
      const int i = base + grp;
      const bool live = i < L.n;
      const double s = fused_row_sum(L.A, i, live, sub, [&](int k) {
        const int j = __ldg(L.A.idx + k);
        return __ldg(L.A.val + k) * (L.omega * (__ldg(L.dinv + j) * __ldcg(L.b + j)));
      });
      if (live && sub == 0) {
        const double bi = __ldcg(L.b + i);
        __stcg(L.x + i, L.omega * __ldg(L.dinv + i) * bi);
        __stcg(L.r + i, bi - s);
      }
    }
    cluster_sync_all();
    const FusedLevel C = lv[l + 1];
    for (int base = 0; base < L.R.rows; base += ngrp) {
      const int i = base + grp;
      const bool live = i < L.R.rows;
      const double s = fused_row_sum(L.R, i, live, sub, [&](int k) {
        return __ldg(L.R.val + k) * __ldcg(L.r + __ldg(L.R.idx + k));
      });
      if (live && sub == 0) __stcg(C.b + i, s);
    }
    cluster_sync_all();
  }
  {
    const FusedLevel C = lv[n_levels - 1];
    const int warp = tid >> 5, lane = threadIdx.x & 31, nwarp = nth >> 5;
    for (int i = warp; i < nc; i += nwarp) {
      const double* row = coarse_inv + static_cast<size_t>(i) * nc;
      double s = 0.0;
      for (int j = lane; j < nc; j += 32) s += __ldg(row + j) * __ldcg(C.b + j);
      s = warp_sum(s);
      if (lane == 0) __stcg(C.y + i, s);
    }
    cluster_sync_all();
  }
  for (int l = n_levels - 2; l >= first; --l) {
    const FusedLevel L = lv[l];
    const FusedLevel C = lv[l + 1];
    for (int base = 0; base < L.n; base += ngrp) {
      const int i = base + grp;
      const bool live = i < L.n;
      const double s = fused_row_sum(L.P, i, live, sub, [&](int k) {
        return __ldg(L.P.val + k) * __ldcg(C.y + __ldg(L.P.idx + k));
      });
      if (live && sub == 0) __stcg(L.x + i, __ldcg(L.x + i) + s);
    }
    cluster_sync_all();
    for (int base = 0; base < L.n; base += ngrp) {
      const int i = base + grp;
      const bool live = i < L.n;
      const double s = fused_row_sum(L.A, i, live, sub, [&](int k) {
        return __ldg(L.A.val + k) * __ldcg(L.x + __ldg(L.A.idx + k));
      });
      if (live && sub == 0)
        __stcg(L.y + i, __ldcg(L.x + i) + L.omega * __ldg(L.dinv + i) * (__ldcg(L.b + i) - s));
    }
    if (l > first) cluster_sync_all();
  }
}

}  // namespace tdgl
