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
// flags in the top bits of a window descriptor's count (engine.h window_descriptors)
constexpr int kWinFlagShift = 28;
constexpr int kWinHalo = 1;    // the window references a halo column
constexpr int kWinPush = 2;    // the window holds a row that is sent to a peer

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
  int flags;   // kWinHalo | kWinPush of the window (uniform over the CTA)
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
    const uint32_t n = static_cast<uint32_t>(wd.y) & ((1u << kWinFlagShift) - 1u);
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
  w.flags = wd.y >> kWinFlagShift;
  w.kb = w.ke = 0;
  if (w.row < m.rows) {
    w.kb = __ldg(m.ptr + w.row) - k0a;
    w.ke = __ldg(m.ptr + w.row + 1) - k0a;
  }
  __syncthreads();  // the initialised barrier is visible to every waiter
  return w;
}

// Halo-aware load of a gathered vector entry as double (the mailbox always carries doubles).
template <typename T>
__device__ __forceinline__ double halo_val(Ctl* ctl, const HaloView& h, const T* __restrict__ x, int j) {
  if (j < h.n_owned) return static_cast<double>(__ldg(x + j));
  bool ok = true;
  return poll_f64(ctl, h.box + static_cast<long long>(j - h.n_owned) * 2, h.tag, &ok);
}

// sum_k val[k] * x[idx[k]] over the staged row, four gathers in flight per thread, added in
// column order, accumulated in double whatever the storage types (TV: matrix values, TX: the
// gathered vector — float for the operators of the V-cycle, see kw_real).
// Halo columns — sharded engine only — come out of the mailbox (comm.cuh).  The gathers are
// issued unconditionally from the array (index clamped), so that they stay independent and
// in flight together; the rare halo entries are patched afterwards.
template <bool SH, typename TV = double, typename TX = double>
__device__ __forceinline__ double row_dot(const TV* __restrict__ sv,
                                          const int* __restrict__ si, int kb, int ke,
                                          const TX* __restrict__ x, Ctl* ctl,
                                          const HaloView& h) {
  double s = 0.0;
  const int last = ke - 1;
  const int top = SH ? h.n_owned - 1 : 0x7fffffff;
  if constexpr (!SH && sizeof(TX) == 4) {
    // float vector, no halo: EIGHT gathers in flight per trip — a mesh row (7-8 entries) is one
    // round trip to L1/L2 instead of two.  Same products, same order of additions (entries past
    // the end of the row contribute fma(0, x, s) = s).
    for (int k = kb; k < ke; k += 8) {
      int j[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) j[q] = si[min(k + q, last)];
      TX xv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) xv[q] = __ldg(x + j[q]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double v = (k + q < ke) ? static_cast<double>(sv[min(k + q, last)]) : 0.0;
        s = fma(v, static_cast<double>(xv[q]), s);
      }
    }
    return s;
  } else {
#pragma unroll 2
  for (int k = kb; k < ke; k += 4) {
    const int k1 = min(k + 1, last), k2 = min(k + 2, last), k3 = min(k + 3, last);
    const int j0 = si[k], j1 = si[k1], j2 = si[k2], j3 = si[k3];
    double x0 = __ldg(x + min(j0, top)), x1 = __ldg(x + min(j1, top)), x2 = __ldg(x + min(j2, top)),
           x3 = __ldg(x + min(j3, top));
    if (SH && max(max(j0, j1), max(j2, j3)) > top) {
      if (j0 > top) x0 = halo_val(ctl, h, x, j0);
      if (j1 > top) x1 = halo_val(ctl, h, x, j1);
      if (j2 > top) x2 = halo_val(ctl, h, x, j2);
      if (j3 > top) x3 = halo_val(ctl, h, x, j3);
    }
    const double v0 = sv[k], v1 = (k + 1 < ke) ? static_cast<double>(sv[k1]) : 0.0,
                 v2 = (k + 2 < ke) ? static_cast<double>(sv[k2]) : 0.0,
                 v3 = (k + 3 < ke) ? static_cast<double>(sv[k3]) : 0.0;
    s = fma(v0, x0, s);
    s = fma(v1, x1, s);
    s = fma(v2, x2, s);
    s = fma(v3, x3, s);
  }
  return s;
  }
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
  // element types are fixed by the kernel instantiation (RealTypes below)
  const void* val = nullptr;
  const void* x = nullptr;       // gathered vector
  void* y = nullptr;             // row output
  const void* b = nullptr;       // right-hand side (residual / smoothers; gathered by kOpPresmooth)
  const void* dinv = nullptr;    // 1 / diag (smoothers)
  const void* w = nullptr;       // kOpJacobi: dot(w, y) -> *red_out (type of b)
  void* r = nullptr;             // kOpPresmooth: residual output
  double omega = 0.0;
  double* red_out = nullptr;     // reduction result (deterministic), may be null
  double* red2_out = nullptr;    // kOpSpmvCg: the second sum
  // sharded engine: where the halo columns of the gathered vector (x; b for kOpPresmooth)
  // are found, and where the boundary rows of the output (r for kOpPresmooth, y otherwise)
  // go.  Defaults: no halo, nothing to send.
  HaloArgs halo;
  PushArgs push;
  int trace_id = 0;              // debug timeline slot (kernels.cuh trace_in / trace_out)
};

// Storage types of one kw_real instantiation.  The CG iteration works in double (kTypesD).
// The V-cycle is only a preconditioner: its operators (A, P, R, 1/diag of every level) and its
// vectors are stored in float — two thirds of the matrix bytes (4 + 4 instead of 8 + 4 per
// entry) and half of the vector bytes of the kernels that dominate the step — while every row
// is still accumulated in double.  The two places where the cycle touches CG's double vectors
// read CG's double residual r have their own type set (kTypesP0: the fine-level smoothers; the
// post-smoother's output z is float too), and CG's SpMV gathers that float z (kTypesZ).
template <typename TV_, typename TX_, typename TB_, typename TY_, typename TR_>
struct RealTypes {
  using V = TV_;   // matrix values, 1 / diag
  using X = TX_;   // gathered vector x
  using B = TB_;   // right-hand side b (and w)
  using Y = TY_;   // output y
  using R = TR_;   // residual output r (kOpPresmooth)
};
using kTypesD = RealTypes<double, double, double, double, double>;
using kTypesF = RealTypes<float, float, float, float, float>;
using kTypesP0 = RealTypes<float, float, double, float, float>;
using kTypesZ = RealTypes<double, float, double, double, double>;   // CG's SpMV: w = A z, z float

// psi is double-buffered, and so are its mailboxes: [b] belongs to buffer b
struct PsiComm {
  HaloArgs halo[2];
  PushArgs push[2];
};

// The accumulated part of one row of kw_real (this lane's share of it when LPR lanes split the
// row).  HALO: the window references halo columns (sharded engine) — they come out of the
// mailbox; without it the loop carries no exchange code at all.
template <int OP, bool HALO, int LPR, typename T>
__device__ __forceinline__ double real_row_sum(const typename T::V* __restrict__ sv,
                                               const int* __restrict__ si, const WinRow& w, int sub,
                                               const RealArgs& a, Ctl* ctl, const HaloView& hv) {
  using TV = typename T::V;
  using TX = typename T::X;
  using TB = typename T::B;
  const TX* ax = static_cast<const TX*>(a.x);
  const TB* ab = static_cast<const TB*>(a.b);
  const TV* adinv = static_cast<const TV*>(a.dinv);
  const int top = HALO ? hv.n_owned - 1 : 0x7fffffff;
  double s = 0.0;
  if (OP == kOpPresmooth) {
    // x_j = omega dinv_j b_j on the fly
    const int last = w.ke - 1;
    for (int k = w.kb + sub; k < w.ke; k += 2 * LPR) {
      const int k1 = min(k + LPR, last);
      const bool two = k + LPR < w.ke;
      const int j0 = si[k], j1 = si[k1];
      double b0, b1;
      if (HALO) {
        b0 = __ldg(ab + min(j0, top)); b1 = __ldg(ab + min(j1, top));
        if (max(j0, j1) > top) {
          if (j0 > top) b0 = halo_val(ctl, hv, ab, j0);
          if (j1 > top) b1 = halo_val(ctl, hv, ab, j1);
        }
      } else {
        b0 = __ldg(ab + j0); b1 = __ldg(ab + j1);
      }
      const double t0 = static_cast<double>(__ldg(adinv + j0)) * b0;
      const double t1 = static_cast<double>(__ldg(adinv + j1)) * b1;
      s = fma(static_cast<double>(sv[k]), a.omega * t0, s);
      s = fma(two ? static_cast<double>(sv[k1]) : 0.0, a.omega * t1, s);
    }
  } else if (LPR == 1) {
    s = row_dot<HALO, TV, TX>(sv, si, w.kb, w.ke, ax, ctl, hv);
  } else {
#pragma unroll 2
    for (int k = w.kb + sub; k < w.ke; k += 2 * LPR) {
      const int k1 = k + LPR;
      const bool two = k1 < w.ke;
      const int j0 = si[k], j1 = two ? si[k1] : j0;
      double x0, x1;
      if (HALO) {
        x0 = __ldg(ax + min(j0, top)); x1 = __ldg(ax + min(j1, top));
        if (max(j0, j1) > top) {
          if (j0 > top) x0 = halo_val(ctl, hv, ax, j0);
          if (j1 > top) x1 = halo_val(ctl, hv, ax, j1);
        }
      } else {
        x0 = __ldg(ax + j0); x1 = __ldg(ax + j1);
      }
      s = fma(static_cast<double>(sv[k]), x0, s);
      s = fma(two ? static_cast<double>(sv[k1]) : 0.0, x1, s);
    }
  }
  return s;
}

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
template <int OP, bool SH, int LPR, typename T = kTypesD>
__global__ void __launch_bounds__(kWinRows, 8)   // 8 CTAs/SM = 2048 threads: <= 32 registers
kw_real(Ctl* ctl, Comm* comm, WinCsr m, RealArgs a, double* partials, unsigned int* counter) {
  using TV = typename T::V;
  using TX = typename T::X;
  using TB = typename T::B;
  using TY = typename T::Y;
  using TR = typename T::R;
  extern __shared__ __align__(128) unsigned char win_smem[];
  __shared__ uint64_t bar;
  __shared__ double red[64];
  trace_in(ctl, a.trace_id, 0);
  const WinRow w = window_stage<sizeof(TV), 0, LPR>(m, a.val, nullptr, win_smem, &bar);
  const TV* sv = reinterpret_cast<const TV*>(win_smem);
  const int* si = reinterpret_cast<const int*>(win_smem + static_cast<size_t>(m.cap) * sizeof(TV));
  const TX* ax = static_cast<const TX*>(a.x);
  const TB* ab = static_cast<const TB*>(a.b);
  const TB* aw = static_cast<const TB*>(a.w);
  const TV* adinv = static_cast<const TV*>(a.dinv);
  TY* ay = static_cast<TY*>(a.y);
  TR* ar = static_cast<TR*>(a.r);
  griddep_wait();
  trace_in(ctl, a.trace_id, 1);
  const bool live = (ctl->status == 0);
  const bool in = live && w.row < m.rows;
  const int sub = threadIdx.x & (LPR - 1);   // lane of the row (LPR lanes share a long row)
  const bool lead = in && sub == 0;          // the lane that owns the row's epilogue
  // sharded: only the CTAs whose window references a halo column / holds a row that is sent
  // (flags of the window descriptor; the rows near the cuts are numbered first) run the
  // exchange-aware code — every other CTA runs the single-GPU instruction stream
  const bool hcta = SH && (w.flags & kWinHalo) != 0;
  HaloView hv;
  hv.n_owned = 0x7fffffff; hv.box = nullptr; hv.tag = 0u;
  unsigned int tag_out = 0u;
  if (SH) {
    if (hcta) hv = halo_view(ctl, comm, a.halo);
    if ((w.flags & kWinPush) != 0 && comm != nullptr && a.push.bnd != nullptr)
      tag_out = comm_tag(ctl, a.push.tag_mode);
  }
  // row-local operands travel while the window lands
  double bi = 0.0, di = 0.0, xi = 0.0, wi = 0.0;
  if (lead) {
    if (OP == kOpResidual || OP == kOpPresmooth || OP == kOpJacobi || OP == kOpSpmvCg)
      bi = ab[w.row];
    if (OP == kOpPresmooth || OP == kOpJacobi) di = adinv[w.row];
    if (OP == kOpSpmvDot || OP == kOpJacobi || OP == kOpSpmvCg) xi = ax[w.row];
    if (OP == kOpPlainAdd) xi = ay[w.row];
    if (OP == kOpJacobi && aw != nullptr) wi = aw[w.row];
  }
  mbar_wait(&bar, 0);
  if (!live) return;
  double d = 0.0, d2 = 0.0;
  double s = 0.0;
  if (in) {
    if (hcta)
      s = real_row_sum<OP, true, LPR, T>(sv, si, w, sub, a, ctl, hv);
    else
      s = real_row_sum<OP, false, LPR, T>(sv, si, w, sub, a, ctl, hv);
  }
  if (LPR > 1) s = group_sum<LPR>(s);
  if (lead) {
    double out;  // the value other shards may need
    if (OP == kOpSpmvDot) {
      ay[w.row] = static_cast<TY>(out = s);
      d = s * xi;
    } else if (OP == kOpSpmvCg) {
      ay[w.row] = static_cast<TY>(out = s);
      d = bi * xi;
      d2 = s * xi;
    } else if (OP == kOpResidual) {
      const double ri = bi - s;
      ay[w.row] = static_cast<TY>(out = ri);
      d = ri * ri;
    } else if (OP == kOpPresmooth) {
      ay[w.row] = static_cast<TY>(a.omega * di * bi);
      ar[w.row] = static_cast<TR>(out = bi - s);
    } else if (OP == kOpJacobi) {
      const double yi = xi + a.omega * di * (bi - s);
      ay[w.row] = static_cast<TY>(out = yi);
      d = wi * yi;
    } else if (OP == kOpPlain) {
      ay[w.row] = static_cast<TY>(out = s);
    } else {
      ay[w.row] = static_cast<TY>(out = xi + s);
    }
    // (what travels is what the owner stored: a float-typed output is rounded first, so that
    // a halo copy equals the owner's value bit for bit)
    if (SH && tag_out != 0u) {
      const double sent = (OP == kOpPresmooth) ? static_cast<double>(static_cast<TR>(out))
                                               : static_cast<double>(static_cast<TY>(out));
      push_row(comm, a.push, tag_out, w.row, sent);
    }
    (void)out;
  }
  if (OP == kOpSpmvCg) {
    double t0, t1;
    if (grid_sum2_fused(d, d2, partials, counter, red, &t0, &t1) && threadIdx.x < 32) {
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
  trace_out(ctl, a.trace_id);
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
            const double* __restrict__ old_sq /* screening: |psi|^2 of the step's input; else null */,
            const double* __restrict__ eps1 /* epsilon = eps + eps_g(t) * eps1; null: static */) {
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
  const bool hcta = SH && comm != nullptr && (w.flags & kWinHalo) != 0;
  const bool pcta = SH && comm != nullptr && (w.flags & kWinPush) != 0;
  HaloView hv;
  hv.n_owned = 0x7fffffff; hv.box = nullptr; hv.tag = 0u;
  unsigned int tag_out = 0u;
  if (hcta) hv = halo_view(ctl, comm, pc.halo[cur]);
  if (pcta) tag_out = comm_tag(ctl, kTagPsiNew);
  double2 p = make_double2(0.0, 0.0);
  double mui = 0.0, epsi = 0.0;
  bool fx = false;
  if (in) {
    p = psi[w.row];
    mui = mu[w.row];
    epsi = eps[w.row];
    // (no FMA contraction: the host evaluates e0 + g * e1 with a rounded product)
    if (eps1 != nullptr) epsi = __dadd_rn(epsi, __dmul_rn(ctl->eps_g, eps1[w.row]));
    fx = fixed[w.row] != 0;
  }
  double abs2 = p.x * p.x + p.y * p.y;
  if (in && old_sq != nullptr) abs2 = old_sq[w.row];
  mbar_wait(&bar, 0);
  if (!live) return;
  double dmax = 0.0;
  int failed = 0;
  if (in) {
    double2 lap = hcta ? row_dot_c<true>(sv, si, w.kb, w.ke, psi, ctl, hv)
                       : row_dot_c<false>(sv, si, w.kb, w.ke, psi, ctl, hv);
    if (fx) lap = p;
    const PsiOut o = psi_update(p, lap, mui, epsi, ctl->gamma, ctl->u, dt, abs2);
    out[w.row] = o.psi;
    if (pcta) push_row(comm, pc.push[cur ^ 1], tag_out, w.row, o.psi);
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
  const bool hcta = SH && comm != nullptr && (w.flags & kWinHalo) != 0;
  if (hcta) {
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
    HaloView plain;
    plain.n_owned = 0x7fffffff; plain.box = nullptr; plain.tag = 0u;
    double2 lap;
    double am, amp, ampp;
    if (hcta) {
      lap = row_dot_c<true>(sl, si, w.kb, w.ke, psi, ctl, hpsi);
      am = row_dot<true>(sa, si, w.kb, w.ke, mu, ctl, hmu);
      amp = row_dot<true>(sa, si, w.kb, w.ke, mu_prev, ctl, hmp);
      ampp = row_dot<false>(sa, si, w.kb, w.ke, mu_pp, ctl, plain);
    } else {
      // one walk over the row for all four products: per entry one complex and three real
      // gathers at the same column, two entries in flight (added in column order like row_dot)
      double lx = 0.0, ly = 0.0;
      am = amp = ampp = 0.0;
      const int last = w.ke - 1;
#pragma unroll 1
      for (int k = w.kb; k < w.ke; k += 2) {
        const int k1 = min(k + 1, last);
        const bool two = k + 1 < w.ke;
        const int j0 = si[k], j1 = si[k1];
        const double2 p0 = __ldg(psi + j0), p1 = __ldg(psi + j1);
        const double m0 = __ldg(mu + j0), m1 = __ldg(mu + j1);
        const double q0 = __ldg(mu_prev + j0), q1 = __ldg(mu_prev + j1);
        const double r0 = __ldg(mu_pp + j0), r1 = __ldg(mu_pp + j1);
        const double2 l0 = sl[k];
        double2 l1 = sl[k1];
        const double a0 = sa[k], a1 = two ? sa[k1] : 0.0;
        if (!two) l1 = make_double2(0.0, 0.0);
        lx += l0.x * p0.x - l0.y * p0.y;
        ly += l0.x * p0.y + l0.y * p0.x;
        lx += l1.x * p1.x - l1.y * p1.y;
        ly += l1.x * p1.y + l1.y * p1.x;
        am = fma(a0, m0, am); am = fma(a1, m1, am);
        amp = fma(a0, q0, amp); amp = fma(a1, q1, amp);
        ampp = fma(a0, r0, ampp); ampp = fma(a1, r1, ampp);
      }
      lap = make_double2(lx, ly);
    }
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
    {
      // four blocks' partials (16 x 16-byte loads) in flight per trip; same order of additions
      const double2* p2 = reinterpret_cast<const double2*>(partials);
      const unsigned int n = gridDim.x, bd = blockDim.x;
      unsigned int i = threadIdx.x;
      for (; i + 3u * bd < n; i += 4u * bd) {
        double2 v[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) v[q][k] = __ldcg(p2 + 4 * (i + q * bd) + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[0] += v[q][0].x; a[1] += v[q][0].y; a[2] += v[q][1].x; a[3] += v[q][1].y;
          a[4] += v[q][2].x; a[5] += v[q][2].y; a[6] += v[q][3].x;
        }
      }
      for (; i < n; i += bd) {
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] += __ldcg(partials + 8 * i + k);
      }
    }
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
