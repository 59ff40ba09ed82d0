/*
 * tdgl_b200 — C ABI of the B200-native TDGL time stepper.
 *
 * This is the drop-in boundary for ONE path of pyTDGL (loganbvh/py-tdgl v0.8.3): the
 * per-step hot path of the adaptive-Euler solver.  The reference has no FFI layer; its
 * only backend seam is the Python call `TDGLSolver.update(state, running_state, dt,
 * **values)` made once per step by `Runner._run_stage` (tdgl/solver/runner.py:417-423)
 * plus the operator container `MeshOperators` (tdgl/finite_volume/operators.py:233-394).
 * The entry points below are what a ctypes binding inside the reference would call
 * instead (see INTEGRATION.md for that stub).
 *
 * Conventions
 *   - plain C types only; every pointer argument is a caller-owned HOST buffer that is
 *     copied during the call; the handle owns all device memory.
 *   - all quantities are dimensionless exactly as inside the reference after
 *     TDGLSolver.__init__ (lengths in xi, A in xi*Bc2, current density in K0/4).
 *   - complex arrays are interleaved (re, im) doubles, i.e. numpy complex128.
 *   - index arrays are int64 like the reference's (finite_volume/mesh.py:59-60,
 *     edge_mesh.py:36).
 *   - every call returns 0 on success or a TDGL_E_* code; tdgl_last_error() gives text.
 *   - one host thread per handle; handles are independent (no global state).
 */
#ifndef TDGL_B200_H
#define TDGL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tdgl_handle tdgl_handle;

enum {
  TDGL_OK = 0,
  TDGL_E_STEP_FAILED = 1,  /* |psi|^2 solve failed max_solve_retries times
                              (reference RuntimeError, solver/solver.py:478-483) */
  TDGL_E_MU_SOLVER = 2,    /* mu solver did not reach tolerance in max iterations; also: the
                              screening iteration did not converge (info.status == 4, reference
                              RuntimeError solver/solver.py:657-663) */
  TDGL_E_CUDA = 3,         /* CUDA / NCCL runtime error */
  TDGL_E_INVALID = 4       /* bad argument */
};

/* Engine knobs that have no counterpart in the reference.  Zero-initialise and set
 * struct_size = sizeof(tdgl_config); fields left 0 take the default in brackets. */
typedef struct tdgl_config {
  int32_t struct_size;
  int32_t device;           /* CUDA device ordinal [0] */
  double mu_rtol;           /* ||r|| <= mu_rtol * ||b|| for the mu solve [1e-10] */
  int32_t mu_max_iter;      /* CG iteration cap [500] */
  double amg_theta;         /* strength threshold of the aggregation [0.08] */
  int32_t amg_max_coarse;   /* stop coarsening at this size [200] */
  int32_t use_graph;        /* 1: one CUDA graph with device-side loops per advance();
                               2: host-driven launches (debug) [1] */
  int32_t reorder;          /* 1: renumber sites along a Z-order curve when coordinates
                               are given; 2: keep the caller's numbering [1] */
  int32_t running_capacity; /* steps of running state kept per advance() [4096] */
  int32_t world;            /* number of shards the mesh is decomposed into, 1..8 [1] */
  int32_t rank;             /* the shard this handle computes, 0..world-1 [0] */
  int32_t replicate_below;  /* sharded: AMG levels with at most this many rows are computed
                               redundantly by every shard instead of exchanged [32768] */
  int32_t fuse_coarse;      /* 1: AMG levels with <= 4096 rows run as ONE thread-block-cluster
                               kernel; 2: one kernel per operator per level.  Measured at 1M
                               sites the fused form is not faster, hence [2] */
} tdgl_config;

/* Mesh + material -> device-resident operators.  Replaces MeshOperators.__init__ +
 * build_operators (operators.py:245-308): mu Laplacian (no fixed sites), Neumann
 * boundary matrix, gradient, divergence; the SuperLU factorisation (operators.py:306-308)
 * is replaced by the setup of an algebraic-multigrid preconditioner.
 *   edges[E][2], areas[N], edge_lengths[E], dual_edge_lengths[E], directions[E][2],
 *   boundary_edge_indices[Eb]: Mesh / EdgeMesh arrays (mesh.py:24-69, edge_mesh.py:9-42)
 *   fixed_sites[n_fixed], fix_psi: MeshOperators(fixed_sites=, fix_psi=)  (solver.py:270-276)
 *   sites_xy[N][2]: optional (may be NULL) coordinates, used only for memory locality
 *   gamma, u: Layer parameters (solver.py:152-153)
 *   probe_sites[n_probe]: device.probe_point_indices (solver.py:142)                 */
int tdgl_create(tdgl_handle** out, int64_t n_sites, int64_t n_edges, int64_t n_boundary_edges,
                const int64_t* edges, const double* areas, const double* edge_lengths,
                const double* dual_edge_lengths, const double* directions,
                const int64_t* boundary_edge_indices, const int64_t* fixed_sites,
                int64_t n_fixed, int32_t fix_psi, const double* sites_xy, double gamma,
                double u, const int64_t* probe_sites, int64_t n_probe,
                const tdgl_config* config);

void tdgl_destroy(tdgl_handle* h);

const char* tdgl_last_error(const tdgl_handle* h);

/* MeshOperators.set_link_exponents (operators.py:310-383): (re)builds the values of the
 * covariant Laplacian and gradient from A[E][2] (dimensionless, at edge centres). */
int tdgl_set_link_exponents(tdgl_handle* h, const double* A);

/* epsilon[N] (solver.py:191-216, 644-646). */
int tdgl_set_epsilon(tdgl_handle* h, const double* epsilon);

/* mu_boundary[Eb]: the terminal current densities on boundary edges written by
 * TDGLSolver.update_mu_boundary (solver.py:325-345). */
int tdgl_set_mu_boundary(tdgl_handle* h, const double* mu_boundary);

/* dA_dt[E]: time derivative of the applied vector potential projected on the normalised
 * edge directions, as TDGLSolver.update forms it for a time-dependent Parameter
 * (solver.py:626-634); enters the rhs (solver.py:508) and the normal current (solver.py:519).
 * NULL: static vector potential (dA_dt = 0).  Call tdgl_set_link_exponents with the new A
 * as well (solver.py:635-639). */
int tdgl_set_dA_dt(tdgl_handle* h, const double* dA_dt);

/* Separable time-dependent vector potential A(r, t) = f(t) * A0(r) (what the reference's
 * `LinearRamp(...) * ConstantField(...)` is, tdgl/sources, parameter.py:355-373), evaluated on
 * the DEVICE every step instead of through a host callback (solver.py:626-642): A0[E][2]
 * dimensionless at the edge centres, f piecewise linear through (t_knots[k], f_knots[k]),
 * 2 <= n_knots <= 32, constant outside.  Link variables are rebuilt and dA/dt =
 * (f(t) - f(t_prev)) / dt_prev * A0 enters the rhs and J_n exactly as in the reference.
 * n_knots = 0 turns the ramp off. */
int tdgl_set_vector_potential_ramp(tdgl_handle* h, const double* A0, int32_t n_knots,
                                   const double* t_knots, const double* f_knots);

/* psi[N] (complex128) and mu[N]: the `psi`, `mu` values Runner threads through update(). */
int tdgl_set_state(tdgl_handle* h, const double* psi, const double* mu);

/* Time-dependent terminal currents without the per-step host callback (the reference calls
 * current_func(t) from Python every step, solver.py:325-345): I_k(t) piecewise linear through
 * (t_knots[j], currents[k][j]) (constant outside), already J_scale-d, terminals in the caller's
 * order; terminal_of_boundary_edge[Eb] = terminal index of every boundary edge or -1;
 * terminal_lengths[n_terminals].  The device evaluates J_ext,k = -(1 / L_k) sum_{j != k} I_j(t)
 * at the start of every step and rewrites the boundary term of the sites at the terminals when
 * a density changed.  n_knots = 0 turns the table off (tdgl_set_mu_boundary applies again). */
int tdgl_set_terminal_current_table(tdgl_handle* h, int32_t n_terminals, const int32_t* terminal_of_boundary_edge,
                                    const double* terminal_lengths, int32_t n_knots,
                                    const double* t_knots, const double* currents);
/* epsilon(r, t) = epsilon0(r) + g(t) * epsilon1(r), g piecewise linear through the knots
 * (update_epsilon, solver.py:364-381, 644-646, evaluated inside the psi step).  n_knots = 0: off. */
int tdgl_set_epsilon_table(tdgl_handle* h, const double* epsilon0, const double* epsilon1,
                           int32_t n_knots, const double* t_knots, const double* g_knots);

/* Screening (SolverOptions.include_screening; reference solver/solver.py:304-314, 522-578,
 * 650-688, solver/screening.py:12-42, finite_volume/mesh.py:203-243): every time step iterates
 * Polyak's method on the induced vector potential
 *     A_induced[e] = scale * sum_j J_site[j] * areas[j] / |edge_centers[e] - sites_xy[j]|,
 * J_site = site average of the edge current J_s + J_n, with the link variables rebuilt from
 * A_applied + A_induced in every pass, until the relative change is below `tolerance`
 * (status 4 / TDGL_E_MU_SOLVER after `max_iterations` passes, like the reference's RuntimeError).
 *   sites_xy[N][2], edge_centers[E][2]: the coordinates the reference uses (xi * mesh coordinates);
 *   scale: the reference's mu_0 / (4 pi) * K0 / A0 (in 1 / length_units) * xi^2 per unit mesh area.
 * enable = 0 turns screening off.  Not available on a sharded engine. */
int tdgl_set_screening(tdgl_handle* h, int32_t enable, double scale, const double* sites_xy,
                       const double* edge_centers, double tolerance, int32_t max_iterations,
                       double step_size, double step_drag);
/* The `induced_vector_potential` value Runner threads through update(): [E][2]. */
int tdgl_set_induced_vector_potential(tdgl_handle* h, const double* A_induced);
int tdgl_get_induced_vector_potential(tdgl_handle* h, double* A_induced);
/* RunningState "screening_iterations" of the steps of the last advance() (solver.py:695-696). */
int tdgl_get_running_screening(tdgl_handle* h, int64_t capacity, int64_t* iterations);

/* The SolverOptions fields the step reads (options.py:66-89) and the controller state
 * TDGLSolver keeps between steps (solver.py:316-320): resets tentative_dt = dt_init and
 * clears the |psi|^2-change history. */
int tdgl_set_stepper(tdgl_handle* h, double dt_init, double dt_max, int32_t adaptive,
                     int32_t adaptive_window, int32_t max_solve_retries,
                     double adaptive_time_step_multiplier);

typedef struct tdgl_advance_info {
  int64_t steps_done;      /* update() calls performed by this advance() */
  int64_t step;            /* Runner's step index after the call */
  double time;             /* Runner's time after the call */
  double dt;               /* dt used by the last step (SolverResult.dt) */
  double tentative_dt;     /* TDGLSolver.tentative_dt for the next step */
  int32_t finished;        /* 1 if time >= t_end was reached (Runner breaks) */
  int32_t status;          /* TDGL_OK or the failure that stopped the stepping */
  int64_t failed_step;     /* step index at which status was raised */
  double failed_dt;        /* dt at that point (for the reference's error text) */
  int64_t retries;         /* total dt-shrinking retries in this call */
  int64_t mu_iterations;   /* total CG iterations in this call */
  double mu_rel_residual;  /* ||r||/||b|| of the last mu solve */
  double device_ms;        /* CUDA-event time of the stepping on the engine's stream */
  int64_t screening_iterations; /* total passes of the screening (Polyak) loop in this call */
  double screening_error;  /* relative change of A_induced in the last pass (solver.py:571-575) */
} tdgl_advance_info;

/* Runs the loop of Runner._run_stage (runner.py:379-433) on the device: up to
 * `max_steps` calls of update(); after each one stops if time >= t_end (the reference
 * performs that last update and then breaks), else dt <- new_dt, time += dt, step += 1.
 * `step`/`time` are the Runner's current values (they restart at 0 for each stage,
 * runner.py:294-318, while the solver's adaptive history persists). */
int tdgl_advance(tdgl_handle* h, int64_t max_steps, double t_end, int64_t step, double time,
                 tdgl_advance_info* info);

/* One call of TDGLSolver.update() exactly at the reference's step seam
 * (runner.py:417-423, solver.py:580-714): host psi[N] (complex128) / mu[N] in, ONE time
 * step on the device, host psi', mu', supercurrent[E], normal_current[E] out (any output
 * pointer may be NULL).  `step` / `time` are Runner's state["step"] / state["time"]; the
 * dt used is returned in info->dt.  All copies are asynchronous on the engine's stream
 * when the host buffers are pinned (tdgl_host_alloc). */
int tdgl_update(tdgl_handle* h, const double* psi, const double* mu, int64_t step, double time,
                double* psi_out, double* mu_out, double* supercurrent, double* normal_current,
                tdgl_advance_info* info);

/* Page-locked host memory for the arrays that cross the boundary every step. */
/* Shard-local step seam (domain decomposition): the rank hands in and gets back ONLY the
 * entries it owns — 1/world of the whole-mesh traffic of tdgl_update per rank, no sums over
 * zero-padded arrays; halo values travel device to device.
 *   tdgl_local_maps: sizes[2] = {owned sites, owned edges}; sites[owned sites] = caller site
 *     ids in the order of the local arrays; edges[owned edges] = caller edge ids (an edge is
 *     owned by the shard that owns edges[e][0]).  Null arrays are skipped (size query).
 *   tdgl_update_local: like tdgl_update with psi_local / mu_local (in) and psi_out / mu_out /
 *     supercurrent / normal_current (out) in that local order.  Every rank must call it for
 *     every step (the ranks meet in an on-device barrier).  With world = 1 it is tdgl_update
 *     without the renumbering.  When the state handed in equals, bit for bit and on every rank,
 *     the state the engine holds — a Runner-style loop feeding each step's output back
 *     (reference runner.py:417-423) — nothing is replaced and the step is identical to a step
 *     of tdgl_advance (the mu solve keeps its initial-guess history); any other state resets
 *     that history like tdgl_set_state.  psi_out / supercurrent are final before the mu solve
 *     and are copied out under it. */
int tdgl_local_maps(tdgl_handle* h, int64_t* sizes, int64_t* sites, int64_t* edges);
int tdgl_update_local(tdgl_handle* h, const double* psi_local, const double* mu_local, int64_t step,
                      double time, double* psi_out, double* mu_out, double* supercurrent,
                      double* normal_current, tdgl_advance_info* info);

void* tdgl_host_alloc(int64_t bytes);
void tdgl_host_free(void* p);

/* Current psi[N] (complex128) and mu[N]. */
int tdgl_get_state(tdgl_handle* h, double* psi, double* mu);

/* Supercurrent and normal current on the edges for the current state
 * (operators.py:385-394, solver.py:519); computed on demand (save steps). */
int tdgl_get_currents(tdgl_handle* h, double* supercurrent, double* normal_current);

/* Running state of the last advance(): dt[k], mu[n_probe][k], theta[n_probe][k] for
 * k < steps_done (solver.py:690-694); `capacity` is the row stride of the outputs. */
/* Asynchronous save pipeline (the reference saves synchronously from the stepping thread,
 * DataHandler.save_time_step, solver/runner.py:155-183).  tdgl_snapshot_begin stages psi, mu,
 * J_s, J_n of the current state in device buffers on the stepping stream and starts their copy
 * into the slot's page-locked host buffers on a second stream; it returns at once and the
 * next tdgl_advance overlaps the copy.  tdgl_snapshot_wait blocks until that slot's copy has
 * landed and returns the host pointers (whole-mesh arrays in the caller's numbering: psi
 * complex128[N], mu f64[N], currents f64[E]), valid until the slot's next tdgl_snapshot_begin.
 * Two slots (0, 1).  tdgl_snapshot_wait may be called from a second host thread (a writer). */
int tdgl_snapshot_begin(tdgl_handle* h, int32_t slot);
int tdgl_snapshot_wait(tdgl_handle* h, int32_t slot, double** psi, double** mu,
                       double** supercurrent, double** normal_current);

int tdgl_get_running(tdgl_handle* h, int64_t capacity, double* dt, double* mu_probe,
                     double* theta_probe);

/* ---- single operators, for parity tests and microbenchmarks ------------------------- */

/* y = psi_laplacian @ x  (complex CSR SpMV, fixed rows are identity; solver.py:426) */
int tdgl_op_psi_laplacian(tdgl_handle* h, const double* x, double* y);
/* TDGLSolver.solve_for_psi_squared (solver.py:383-439) for one dt; returns new psi and
 * the |psi|^2 root; *failed = 1 where the reference would return None. */
int tdgl_op_psi_step(tdgl_handle* h, const double* psi, const double* mu, double dt,
                     double* psi_out, double* sq_out, int32_t* failed);
/* rhs = divergence @ supercurrent(psi) - mu_boundary_laplacian @ mu_boundary
 * (solver.py:507-510). */
int tdgl_op_mu_rhs(tdgl_handle* h, const double* psi, double* rhs);
/* y = mu_laplacian @ x (operators.py:285). */
int tdgl_op_mu_laplacian(tdgl_handle* h, const double* x, double* y);
/* Solve mu_laplacian @ mu = rhs (solver.py:513-516) to mu_rtol; the result has
 * area-weighted mean zero.  *iterations / *rel_residual may be NULL. */
int tdgl_op_mu_solve(tdgl_handle* h, const double* rhs, double* mu, int32_t* iterations,
                     double* rel_residual);

/* Time one kernel (or kernel sequence) of the step with CUDA events on the engine's
 * stream; returns the mean milliseconds per launch over `reps` launches after one warm-up.
 *   which: 0 psi step (fused SpMV + update)        1 mu rhs (+ warm-start residual)
 *          2 fine-level CG SpMV (w = A z, r.z, z.w) 3 one whole V-cycle
 *          4 one full mu solve from a zero guess   5 fine-level pre-smoothing + residual
 *          6 fine-level Jacobi post-smoothing      7 fine-level restriction
 *          8 fine-level prolongation
 *   flush_l2 != 0: every launch is timed on its own after a kernel that reads a 256 MB
 *   scratch buffer, so that nothing of the operator is left in the 126 MB L2. */
int tdgl_time_kernel(tdgl_handle* h, int32_t which, int32_t reps, int32_t flush_l2,
                     double* mean_ms);

/* Comparator, measurement only: NVIDIA cuSPARSE's generic CSR SpMV (cusparseSpMV) on the
 * engine's own device-resident CSR arrays — the library kernel SURVEY.md section 2.3 names as
 * the bar for the reference's sparse products (operators.py:291-293, 341-342).
 *   which: 0 real f64 mu operator (what kw_real<spmv_cg> applies)
 *          1 complex128 covariant Laplacian (what kw_psi_step applies)
 * libcusparse is opened with dlopen() inside this call only; it is not a load-time dependency
 * and never on the stepping path.  Returns TDGL_E_INVALID when the library is not present. */
int tdgl_time_cusparse(tdgl_handle* h, int32_t which, int32_t reps, int32_t flush_l2,
                       double* mean_ms);

/* which = 2: fine-level CG SpMV w = A z with r.z and z.w (kw_real<spmv_cg>); 9: the fused CG
 * vector update (k_cg_fused). */

/* Sizes and setup facts: [0] N, [1] E, [2] nnz of a site operator, [3] AMG levels,
 * [4] sum of level nnz, [5] coarsest size, [6] kernels launched so far (host count),
 * [7] graph mode actually in use (1/2). */
int tdgl_get_info(tdgl_handle* h, int64_t* out, int32_t n);

/* Measurement only: the in-loop timeline of the kernels of the CG iteration (V-cycle, SpMV,
 * vector update).  A handle created while the environment variable TDGL_B200_TRACE is set (1:
 * also printed to stderr after every tdgl_advance, 2: recorded only) lets these kernels write
 * %globaltimer when their first CTA becomes resident (in), when it passes griddepcontrol.wait
 * (go: the predecessor has drained) and when their last CTA is done (out) — inside the captured
 * CUDA graph, i.e. back to back with warm caches, which no host-side timer can see.  Returns
 * the launches passed since the last call (the LAST pass through each slot of a loop body),
 * ordered by `in`, times in microseconds from the earliest `in`; names: capacity x 64 chars.
 * n_out = 0 when the handle was created without the variable. */
int tdgl_get_trace(tdgl_handle* h, int32_t capacity, int32_t* n_out, char* names, double* in_us,
                   double* go_us, double* out_us, int64_t* counts);

/* ---- domain decomposition (no counterpart in the reference) ------
 * A handle created with tdgl_config.world = P > 1 computes shard `rank` of the mesh: every
 * rank passes the SAME whole-mesh arrays to tdgl_create (and to the tdgl_set_* calls); the
 * engine keeps only its rows (a contiguous range of the Z-order numbering, on every AMG
 * level) plus a halo.  Halo exchanges and scalar all-reduces are device code on the peers'
 * memory (NVLink), so the shards must be wired once before the first tdgl_advance():
 *   one process per GPU : tdgl_comm_export() -> all-gather the 64-byte handles with the
 *                         host library of your choice (torch.distributed) ->
 *                         tdgl_comm_connect_ipc()
 *   one process, P handles (one or several GPUs): tdgl_comm_connect_local()
 * All shards must then call tdgl_advance()/tdgl_update() with the same arguments, each from
 * its own host thread or process.  Outputs (tdgl_get_state, tdgl_get_currents, tdgl_update,
 * tdgl_get_running probes) are whole-mesh arrays holding this shard's sites / edges and
 * zeros elsewhere: sum them over the shards. */
int tdgl_comm_export(tdgl_handle* h, void* handle_out /* 64 bytes */);
int tdgl_comm_connect_ipc(tdgl_handle* h, const void* handles /* world x 64 bytes */,
                          int32_t world);
int tdgl_comm_connect_local(tdgl_handle* h, tdgl_handle* const* peers, int32_t world);
/* Sharded outputs without the detour through host memory: tdgl_stage_outputs leaves the
 * whole-mesh psi (complex128), mu, supercurrent, normal current (what & 1: state, what & 2:
 * currents) in device buffers — this shard's entries, zeros elsewhere — and returns their
 * device pointers and lengths in doubles; the caller sums them over the shards on the
 * devices (e.g. ncclAllReduce in place) and reads the result with tdgl_fetch_outputs (any
 * pointer may be NULL). */
int tdgl_stage_outputs(tdgl_handle* h, int32_t what, void** device_ptrs /* 4 */,
                       int64_t* counts /* 4 */);
int tdgl_fetch_outputs(tdgl_handle* h, double* psi, double* mu, double* supercurrent,
                       double* normal_current);
/* [0] world, [1] rank, [2] sites of the mesh, [3] sites owned, [4] level-0 halo sites,
 * [5] halo entries over all levels, [6] level-0 neighbour shards, [7] level-0 entries sent
 * per exchange. */
int tdgl_shard_info(tdgl_handle* h, int64_t* out, int32_t n);

/* ---- host-only helpers (no GPU needed): used by CPU tests ---------------------------- */

/* Finite-volume (dual) mesh arrays of a triangulation — what the reference computes in
 * Mesh.from_triangulation / EdgeMesh.from_mesh (tdgl/finite_volume/mesh.py:104-151,
 * edge_mesh.py:54-92, util.py:15-28,59-124) with Python loops: unique sorted edges, the
 * is-boundary flag (edge of exactly one triangle), circumcentres, edge centres / directions /
 * lengths, dual (Voronoi) edge lengths, and per site the sum of length * dual_length / 4 over
 * its edges (the Voronoi area of an interior site; the caller redoes boundary sites with the
 * reference's convex-hull convention).  Multi-threaded; the edge arrays need room for
 * 3 * n_triangles rows, *n_edges returns the number filled. */
int tdgl_host_mesh_dual(int64_t n_sites, int64_t n_triangles, const double* sites_xy,
                        const int64_t* elements /* [n_triangles,3] */, int64_t* n_edges,
                        int64_t* edges /* [.,2] */, uint8_t* is_boundary,
                        double* dual_sites /* [n_triangles,2] */, double* centers /* [.,2] */,
                        double* directions /* [.,2] */, double* edge_lengths,
                        double* dual_edge_lengths, double* areas /* [n_sites] */);

/* Builds the AMG hierarchy on the host and reports per-level sizes; optionally applies
 * `n_cycles` of preconditioned CG on the host to rhs to validate the hierarchy.
 * level_rows/level_nnz have room for 32 entries. */
int tdgl_host_amg_probe(int64_t n_sites, int64_t n_edges, const int64_t* edges,
                        const double* edge_lengths, const double* dual_edge_lengths,
                        double theta, int32_t max_coarse, int32_t* n_levels,
                        int64_t* level_rows, int64_t* level_nnz, const double* rhs,
                        double* x, int32_t max_iter, double rtol, int32_t* iterations);

/* Builds the domain decomposition of a mesh for `world` shards exactly as the sharded engine
 * does and validates it on the host: per-level ownership offsets (level_off[32][9]), halo
 * sizes (halo_sizes[32][8]; level_off[31][0] holds the first replicated level), the Z-order permutation (site_owner_perm[n_sites]: position
 * in the ordering -> caller site), and, if rhs/x are given, the sharded AMG-PCG with all
 * shards emulated in this process (same local operators, same exchange lists). */
int tdgl_host_shard_probe(int64_t n_sites, int64_t n_edges, const int64_t* edges,
                          const double* edge_lengths, const double* dual_edge_lengths,
                          const double* sites_xy, int32_t world, double theta,
                          int32_t max_coarse, int64_t replicate_below, int32_t* n_levels,
                          int64_t* level_off, int64_t* halo_sizes, int64_t* site_owner_perm,
                          const double* rhs, double* x, int32_t max_iter, double rtol,
                          int32_t* iterations);

/* Level-0 exchange lists of one shard in the caller's site numbering: owned sites, halo
 * sites, and per peer the owned sites sent to it (send_ptr[world+1] ranges into send_sites,
 * ordered as the peer's halo).  counts[3] = {n_owned, n_halo, n_send}; call with null
 * arrays first to size them. */
int tdgl_host_shard_lists(int64_t n_sites, int64_t n_edges, const int64_t* edges,
                          const double* edge_lengths, const double* dual_edge_lengths,
                          const double* sites_xy, int32_t world, int32_t rank, int64_t* counts,
                          int64_t* owned, int64_t* halo, int64_t* send_ptr, int64_t* send_sites);

const char* tdgl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TDGL_B200_H */
