// Host side of the engine: device-memory ownership, operator upload, launch sequences and
// the CUDA-graph form of Runner._run_stage (reference tdgl/solver/runner.py:379-433).
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "amg_setup.h"
#include "kernels.cuh"
#include "csr_window.cuh"

namespace tdgl {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define TDGL_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      char buf__[512];                                                                    \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call,                       \
               cudaGetErrorString(err__), __FILE__, __LINE__);                            \
      throw ::tdgl::CudaError(buf__);                                                     \
    }                                                                                     \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool owned = true;  // false: a view into the arena (see shard.h)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p && owned) cudaFree(p); p = nullptr; n = 0; owned = true; }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) TDGL_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void view(void* base, size_t count) {  // non-owning window on memory someone else owns
    release();
    p = static_cast<T*>(base);
    n = count;
    owned = false;
  }
  void zero(cudaStream_t s) { if (n) TDGL_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count > n) alloc(count);
    if (count) TDGL_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) TDGL_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

// A real CSR matrix as the window kernels see it: structure + values + rows per CTA.
struct CsrView {
  WinCsr m;
  const void* val = nullptr;   // float (operators of the V-cycle) or double (level-0 CG matrix)
  int win = kWinRows;  // rows per window
  int lpr = 1;         // lanes per row: threads per CTA = win * lpr
  int vbytes = 8;      // bytes per value
};

// The operators of the V-cycle (A of the levels >= 1, P and R of every level) are stored in
// float: the cycle is a preconditioner, CG itself runs in double (csr_window.cuh, RealTypes).
struct DevCsr {
  int rows = 0, cols = 0, win = kWinRows, cap = 0;
  int lpr = 1;            // lanes per row of the kernels that apply it (4: long rows)
  int64_t nnz = 0;
  DevBuf<int> ptr, idx;   // idx / val carry 4 padding elements (see csr_window.cuh)
  DevBuf<int2> wdesc;     // per window {first staged nnz, staged nnz count}
  DevBuf<float> val;
  CsrView view() const { return CsrView{WinCsr{rows, cap, ptr.p, idx.p, wdesc.p}, val.p, win, lpr, 4}; }
};

struct DevLevel {
  int n = 0;            // rows owned by this shard
  int nx = 0;           // owned + halo entries of a vector on this level
  DevCsr A, P, R;       // P: n x n_coarse, R: n_coarse x n
  DevBuf<float> dinv;   // nx entries (the halo part is static, filled at setup)
  double omega = 0.0;   // Jacobi weight (4/3) / rho(D^-1 A)
  DevBuf<float> b, x, y, r;   // work vectors in the arena (level 0: b = CG's r, y = CG's z, double)
  // sharded engine, partitioned / gathered levels: per owned row the (peer, halo entry)
  // pairs its value is stored to, and a per-32-rows "anything to send" byte
  DevBuf<int> push_rptr;
  DevBuf<int2> push_ent;
  DevBuf<unsigned char> push_bnd;
};

// Largest aligned nnz extent of any window of `win` rows (shared-memory elements a CTA
// needs per CSR array).
inline int window_cap(const std::vector<int32_t>& ptr, int64_t rows, int win) {
  int cap = 4;
  for (int64_t r0 = 0; r0 < rows; r0 += win) {
    const int64_t r1 = std::min<int64_t>(r0 + win, rows);
    cap = std::max(cap, ((ptr[r1] + 3) & ~3) - (ptr[r0] & ~3));
  }
  return cap;
}

// {first staged nnz, staged nnz count | flags} of every window of `win` rows.  Flags (sharded
// engine; csr_window.cuh kWinHalo / kWinPush): the window references a halo column
// (idx >= n_owned_cols), the window holds a row whose value is sent to a peer (push_rptr: the
// per-row ranges of the row space's send list).  CTAs whose window has neither run the plain
// single-GPU instruction stream.
inline std::vector<int2> window_descriptors(const std::vector<int32_t>& ptr, int64_t rows, int win,
                                            const int32_t* idx = nullptr, int64_t n_owned_cols = 0,
                                            const std::vector<int>* push_rptr = nullptr) {
  std::vector<int2> d;
  for (int64_t r0 = 0; r0 < rows; r0 += win) {
    const int64_t r1 = std::min<int64_t>(r0 + win, rows);
    const int k0a = ptr[r0] & ~3, k1a = (ptr[r1] + 3) & ~3;
    int flags = 0;
    if (idx != nullptr)
      for (int32_t k = ptr[r0]; k < ptr[r1]; ++k)
        if (idx[k] >= n_owned_cols) { flags |= kWinHalo; break; }
    if (push_rptr != nullptr && !push_rptr->empty()) {
      const int64_t n = static_cast<int64_t>(push_rptr->size()) - 1;
      const int64_t a = std::min(r0, n), b = std::min(r1, n);
      if ((*push_rptr)[b] > (*push_rptr)[a]) flags |= kWinPush;
    }
    if (k1a - k0a >= (1 << kWinFlagShift)) throw std::runtime_error("CSR window too large");
    d.push_back(make_int2(k0a, (k1a - k0a) | (flags << kWinFlagShift)));
  }
  if (d.empty()) d.push_back(make_int2(0, 0));
  return d;
}

// Rows per window: as many as fit `budget` bytes of shared memory at `bytes_per_nnz`.
inline int pick_window(const std::vector<int32_t>& ptr, int64_t rows, int bytes_per_nnz,
                       int budget, int* cap_out, int max_rows = kWinRows) {
  for (int win = max_rows; win >= 32; win /= 2) {
    const int cap = window_cap(ptr, rows, win);
    if (static_cast<int64_t>(cap) * bytes_per_nnz <= budget || win == 32) {
      if (static_cast<int64_t>(cap) * bytes_per_nnz > 200 * 1024)
        throw std::runtime_error("a CSR row window does not fit in shared memory");
      *cap_out = cap;
      return win;
    }
  }
  return 32;
}

struct Config {
  int device = 0;
  double mu_rtol = 1e-10;
  int mu_max_iter = 500;
  double amg_theta = 0.08;
  int amg_max_coarse = 2000;
  int use_graph = 1;
  int reorder = 1;
  int running_capacity = 4096;
  int world = 1;   // number of shards (one GPU / process each, or several per process)
  int rank = 0;    // this engine's shard
  int replicate_below = 0;  // AMG levels with at most this many rows are replicated (0: 32768)
  int fuse_coarse = 0;      // (accepted, ignored: the fused coarse-level kernel of round 1 was
                            // measured not to be faster and removed)
};

class Engine {
 public:
  Engine(int64_t n_sites, int64_t n_edges, int64_t n_bedges, const int64_t* edges,
         const double* areas, const double* edge_len, const double* dual_len,
         const double* directions, const int64_t* bedge_idx, const int64_t* fixed_sites,
         int64_t n_fixed, int fix_psi, const double* sites_xy, double gamma, double u,
         const int64_t* probe_sites, int64_t n_probe, const Config& cfg);
  ~Engine();

  void set_link_exponents(const double* A);
  void set_epsilon(const double* eps);
  void set_mu_boundary(const double* mub);
  void set_dA_dt(const double* dadt);
  void set_ramp(const double* A0, int n_knots, const double* t_knots, const double* f_knots);
  void set_state(const double* psi, const double* mu, bool reset_history = true);
  void set_terminal_currents(int n_term, const int32_t* term_of_bedge, const double* lengths,
                             int n_knots, const double* t_knots, const double* values);
  void set_epsilon_table(const double* eps0, const double* eps1, int n_knots, const double* t_knots,
                         const double* g_knots);
  void set_screening(int enable, double scale, const double* sites_xy, const double* edge_centers,
                     double tolerance, int max_iterations, double step_size, double drag);
  void set_induced(const double* A);
  void get_induced(double* A);
  void get_running_screening(int64_t capacity, int64_t* iterations);
  // asynchronous save pipeline: the state is staged in device buffers on the stepping stream
  // (two scatter + one edge kernel), then drained to pinned host memory on the copy stream
  // while the next chunk of steps runs; snapshot_wait blocks on that slot's copy only
  void snapshot_begin(int slot);
  void snapshot_wait(int slot, double** psi, double** mu, double** js, double** jn);
  void set_stepper(double dt_init, double dt_max, int adaptive, int window, int max_retries,
                   double multiplier);
  struct AdvanceInfo {
    int64_t steps_done, step; double time, dt, tentative_dt; int finished, status;
    int64_t failed_step; double failed_dt; int64_t retries, mu_iterations; double mu_rel_residual;
    double device_ms;
    int64_t screening_iterations; double screening_error;
  };
  AdvanceInfo advance(int64_t max_steps, double t_end, int64_t step, double time);
  AdvanceInfo update(const double* psi, const double* mu, int64_t step, double time,
                     double* psi_out, double* mu_out, double* js, double* jn);
  void local_maps(int64_t* sizes, int64_t* sites, int64_t* edges);
  AdvanceInfo update_local(const double* psi_loc, const double* mu_loc, int64_t step, double time,
                           double* psi_out, double* mu_out, double* js, double* jn);
  void stage_outputs(int what, void** ptrs, int64_t* counts);
  void fetch_outputs(double* psi, double* mu, double* js, double* jn);
  void get_state(double* psi, double* mu);
  void get_currents(double* js, double* jn);
  void get_running(int64_t capacity, double* dt, double* mu_probe, double* theta_probe);

  void op_psi_laplacian(const double* x, double* y);
  void op_psi_step(const double* psi, const double* mu, double dt, double* psi_out,
                   double* sq_out, int* failed);
  void op_mu_rhs(const double* psi, double* rhs);
  void op_mu_laplacian(const double* x, double* y);
  void op_mu_solve(const double* rhs, double* mu, int* iterations, double* rel_res);
  double time_kernel(int which, int reps, int flush_l2);
  double time_cusparse(int which, int reps, int flush_l2);
  void get_info(int64_t* out, int n);

  // ---- sharded engine: wiring of the peer arenas -------------------------------------------
  void comm_export(void* handle_out /* 64 bytes: cudaIpcMemHandle_t */);
  void comm_connect_ipc(const void* handles /* world x 64 bytes, rank order */);
  void comm_connect_local(Engine* const* peers /* world engines of this process */);
  void shard_info(int64_t* out, int n);
  int world() const { return world_; }
  int n_boundary_edges() const { return Eb_; }
  void make_current() { TDGL_CUDA(cudaSetDevice(cfg_.device)); }

  std::string last_error;
  std::string failure_detail();   // what a shard-exchange timeout (status 3) was waiting for

 private:
  // ---- sizes / host copies --------------------------------------------------------------
  Config cfg_;
  int Ng_ = 0;  // sites of the whole mesh
  int N_ = 0;   // sites (rows) owned by this shard (= Ng_ for a single shard)
  int Nx_ = 0;  // owned + halo sites: length of every level-0 vector that is gathered from
  int E_ = 0, Eb_ = 0, nprobe_ = 0;
  int world_ = 1, rank_ = 0;
  bool comm_on_ = false;        // exchanges enabled in the launches being enqueued
  bool connected_ = true;       // peer arenas mapped (always true for a single shard)
  ShardPlan plan_;
  std::vector<ArenaLayout> layouts_;  // arena layout of every rank
  DevBuf<double> arena_;
  DevBuf<Comm> comm_;
  std::vector<void*> ipc_opened_;

  int64_t nnz_ = 0;
  double gamma_, u_, total_area_ = 0.0;
  std::vector<int> perm_;      // internal index -> caller index
  std::vector<int> inv_perm_;  // caller index -> internal index
  std::vector<double> h_dirs_; // [E,2] caller edge order
  int64_t launches_ = 0;
  int graph_mode_ = 2;
  int64_t last_steps_done_ = 0;

  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;

  // ---- site operators (shared structure) --------------------------------------------------
  DevBuf<int> ptr_, idx_, eidx_;
  DevBuf<int2> wdesc0_;
  DevBuf<signed char> head_;
  DevBuf<double2> lval_;           // covariant Laplacian values (all rows kept)
  DevBuf<unsigned char> fixed_;    // rows the reference replaces by identity
  DevBuf<double> areas_, eps_, bterm_;   // bterm_: what the rhs kernel subtracts per site
  DevBuf<double> bterm_base_, dadt_;     // boundary-current term; dA/dt on the edges
  bool has_dadt_ = false;
  void refresh_site_terms();
  // device-side ramp A(r, t) = f(t) A0(r): A0 . e_hat per edge, divergence of it per site
  bool ramp_on_ = false;
  DevBuf<double> ramp_proj_, ramp_div_, ramp_zero_;
  void enqueue_ramp_links();
  int win0_ = kWinRows, cap0_ = 0;  // window geometry of the site operators
  // ---- device-side tables: terminal currents I_k(t), epsilon(r, t) = e0(r) + g(t) e1(r) ----------
  bool cur_on_ = false, eps_dyn_ = false;
  DevBuf<int> ts_site_, ts_ptr_, ts_bedge_, bedge_term_;
  DevBuf<double> eps1_;
  int n_ts_ = 0;
  std::vector<int> h_b0_, h_b1_;   // local site indices of the boundary edges' ends (-1: not owned)
  void enqueue_step_inputs();      // what follows k_step_begin: ramp links, terminal sites
  // ---- screening (row S): induced vector potential, Polyak iteration inside the step ---------
  bool scr_on_ = false;
  DevBuf<double2> aind_, aind_new_, vel_, edir_, ecent_, sxy_, wsite_;
  DevBuf<double> old_sq_, scr_area_;
  DevBuf<long long> run_scr_;
  std::vector<double> h_areas_int_;   // areas in internal site order (host copy)
  void enqueue_screening_pass_begin();
  void enqueue_screening_pass_end(cudaGraphConditionalHandle cond_scr, cudaGraphConditionalHandle cond_psi);
  void rebuild_graph();
  // ---- edges (caller edge order, internal site indices) -----------------------------------
  DevBuf<int> e0_, e1_, own_edges_;
  std::vector<int> h_own_edges_;   // edges owned by this shard (edges[e,0] is an owned site)
  DevBuf<double> elen_, weight_, theta_;
  DevBuf<int> be0_, be1_;
  DevBuf<double> blen_, mub_;
  // ---- state ------------------------------------------------------------------------------
  DevBuf<double2> psi_[2];
  DevBuf<double> mu_;
  DevBuf<int> dperm_, probes_;
  DevBuf<double> run_dt_, run_mu_, run_theta_;
  // ---- mu solver ----------------------------------------------------------------------------
  std::vector<DevLevel> levels_;
  DevBuf<float> coarse_inv_;
  int nc_ld_ = 0;               // leading dimension of coarse_inv_ (nc_ rounded up to 4)
  int nc_ = 0;
  int64_t amg_nnz_ = 0;
  DevBuf<double> cg_b_, cg_r_, cg_p_, cg_Ap_, cg_z_, cg_s_;   // cg_Ap_: w = A z; cg_s_: A p
  // z = M r, the V-cycle's output, is stored in float like everything else the cycle computes
  // (its entries carry float accuracy anyway); CG forms gamma, delta, p from exactly these
  // values in double.  (cg_z_ survives as the rhs kernel's scratch for the guess's d2.)
  DevBuf<float> cg_zf_;
  DevBuf<double> mu_prev_, mu_pp_;   // the two solutions before the last one (extrapolated
                                     // initial guess of the next solve)
  int guess_terms_ = 2;              // 0: warm start, 1: one-term, 2: two-term extrapolation
  DevBuf<double> partials_;
  DevBuf<unsigned int> counter_;
  // ---- scratch for IO ------------------------------------------------------------------------
  DevBuf<double2> tmp_c_;
  DevBuf<double> tmp_d_, tmp_d2_, tmp_e_, tmp_e2_;
  DevBuf<double> flush_;  // 256 MB scratch read between timed launches (time_kernel)
  // ---- control ------------------------------------------------------------------------------
  DevBuf<Ctl> ctl_;
  Ctl* h_ctl_ = nullptr;  // pinned mirror
  cudaGraph_t graph_ = nullptr;
  cudaGraphExec_t graph_exec_ = nullptr;
  cudaGraphConditionalHandle h_step_ = 0, h_psi_ = 0, h_cg_ = 0, h_scr_ = 0;
  // The step seam (update(): host arrays in and out every step) runs ONE step as two graphs
  // — step begin + psi loop | rhs + mu solve + step end — so that psi' and J_s, final after
  // the first, travel to the host on a copy stream while the second runs.
  cudaGraph_t graph_a_ = nullptr, graph_b_ = nullptr;
  cudaGraphExec_t graph_a_exec_ = nullptr, graph_b_exec_ = nullptr;
  cudaGraphConditionalHandle h_psi_a_ = 0, h_cg_b_ = 0;
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t ev_psi_ = nullptr, ev_copy_ = nullptr;
  struct SnapSlot {
    DevBuf<double2> d_psi;
    DevBuf<double> d_mu, d_js, d_jn;
    double *h_psi = nullptr, *h_mu = nullptr, *h_js = nullptr, *h_jn = nullptr;   // pinned
    cudaEvent_t staged = nullptr, done = nullptr;
    bool pending = false;
  };
  SnapSlot snap_[2];
  void build_split_graphs();
  void destroy_graphs();
  void prepare_advance(int64_t max_steps, double t_end, int64_t step, double time);
  AdvanceInfo collect_advance(float dev_ms);

  // ---- launch sequences ------------------------------------------------------------------
  // Every kernel of the stepping sequence is launched with programmatic stream
  // serialisation (PDL): it may become resident while its predecessor drains, issues the
  // bulk copies of its (static) CSR window, and blocks in griddepcontrol.wait until the
  // predecessor's results are visible.  TDGL_B200_PDL=0 turns the attribute off.
  bool pdl_ = true;
  // fine-level pre-smoothed iterate x0 = omega D^-1 r, written by the kernels that produce r
  // (null for a single-level hierarchy: the dense solve needs no smoother)
  const float* x0_dinv() const { return levels_.size() > 1 ? levels_[0].dinv.p : nullptr; }
  float* x0_out() const { return levels_.size() > 1 ? levels_[0].x.p : nullptr; }
  double x0_omega() const { return levels_.size() > 1 ? levels_[0].omega : 0.0; }
  // debug timeline (TDGL_B200_TRACE=1): one slot per enqueued (captured) launch of the CG
  // iteration's kernels; trace_report() prints the last pass through every slot
  bool trace_on_ = false;
  bool trace_print_ = false;
  DevBuf<unsigned long long> trace_;
  std::vector<std::string> trace_names_;
  int trace_slot(const char* name, int rows);
 public:
  void trace_report();
  int trace_collect(int capacity, char* names, double* in_us, double* go_us, double* out_us,
                    int64_t* counts);
 private:
  template <typename... KArgs, typename... Args>
  void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream_;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_ ? 1 : 0;
    TDGL_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  }
  static int grid_win(int rows, int win) { return (rows + win - 1) / win; }
  WinCsr site_csr() const { return WinCsr{N_, cap0_, ptr_.p, idx_.p, wdesc0_.p}; }
  int sm_count_ = 148;
  template <int OP, typename T = kTypesD>
  void launch_real(const CsrView& A, const RealArgs& a);
  int grid_flat(int n) const {
    int g = (n + kBlock - 1) / kBlock;
    return g < 1 ? 1 : (g > 1184 ? 1184 : g);  // 148 SMs x 8 resident blocks
  }
  // k_cg_fused: 48 registers -> 5 resident blocks per SM, one wave, two elements per trip
  int grid_fused(int n) const {
    const int g = (n + 2 * kBlock - 1) / (2 * kBlock);
    return g < 1 ? 1 : (g > 5 * sm_count_ ? 5 * sm_count_ : g);
  }
  void upload_csr(const HostCsr<double>& h, DevCsr& d, int lanes_per_row = 0,
                  int64_t n_owned_cols = -1, const std::vector<int>* push_rptr = nullptr);
  void launch_spmv(const CsrView& A, const double* x, double* y, double* dot_out);
  void enqueue_vcycle(double* r_in, float* z_out);
  void enqueue_psi_step(double* sq_out, double dt_override);
  void enqueue_mu_rhs(double* rhs_raw);
  void enqueue_cg_iteration(cudaGraphConditionalHandle cond);
  void enqueue_mu_finish();
  void host_solve_loop(bool with_guess = false);   // host-driven CG loop on the current b/r
  void enqueue_solve_begin(cudaGraphConditionalHandle cond);
  Comm* comm() const { return comm_on_ ? comm_.p : nullptr; }
  PushArgs make_push(int level, int channel, int tag_mode) const;
  HaloArgs make_halo(int level, int channel, int tag_mode) const;
  PsiComm make_psi_comm() const;
  void enqueue_unpack(int level, int channel, int tag_mode, float* vec);
  void fill_state_boxes();
  void unpack_state_halos(int cur);
  void upload_comm(double* const* peers);
  void configure_kernels();
  void build_graph();
  void sync_ctl_to_host();
  void push_ctl();
  DevBuf<double> aval_;   // level-0 mu matrix values (CG, rhs); structure shared with ptr_/idx_
  DevBuf<float> aval32_;  // the same values in float: the fine-level smoothers of the V-cycle
  CsrView A0() const { return CsrView{site_csr(), aval_.p, win0_, 1, 8}; }
  CsrView A0f() const { return CsrView{site_csr(), aval32_.p, win0_, 1, 4}; }
  CsrView levelA(size_t l) const { return l == 0 ? A0f() : levels_[l].A.view(); }
};

}  // namespace tdgl
