/* mono_abi.h - C ABI of the B200-native operator-split monodomain step.
 *
 * This is the drop-in boundary for ONE path of finsberg/fenicsx-beat (v0.5.0):
 *     beat.MonodomainSplittingSolver.step((t0, t1))
 *       = pointwise ionic ODE update + theta-rule diffusion solve + stimulus
 * The reference has no FFI of its own (it is pure Python over dolfinx/PETSc/NumPy); every entry point
 * below names the reference call it replaces (path:line under the reference tree) so that a
 * maintainer can bind it with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C, no torch / CUDA types; all arrays are HOST pointers unless the name says `_dev`.
 *   - all arithmetic is fp64; indices are int32 (CSR) / int64 (sizes).
 *   - every function returns 0 on success and a negative MONO_E_* code on failure; the message is
 *     available from mono_last_error(ctx) (ctx == NULL: error of the last failed context-free call,
 *     i.e. mono_ctx_create, mono_fem_assemble_p1 or mono_csr_row_patterns, of this thread).
 *     CUDA / NCCL failures never abort the process.
 *   - one context per process-rank and GPU; a context is NOT thread-safe; work is queued on the
 *     context's own CUDA stream and is asynchronous unless documented otherwise (`get`/`info` calls
 *     synchronise).
 *   - there is no CPU fallback: without a CUDA device mono_ctx_create fails.
 */
#ifndef MONO_ABI_H
#define MONO_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MONO_ABI_VERSION 1

/* error codes */
#define MONO_OK 0
#define MONO_E_INVALID (-1)   /* bad argument / call order */
#define MONO_E_CUDA (-2)      /* CUDA runtime error (message has the CUDA string) */
#define MONO_E_NCCL (-3)      /* NCCL error */
#define MONO_E_NOMEM (-4)
#define MONO_E_UNSUPPORTED (-5)

/* cell models compiled into the library (generated from odes/ by fenicsx-beat_b200/codegen) */
#define MONO_MODEL_FHN 0    /* FitzHugh-Nagumo, README.md:58-129 */
#define MONO_MODEL_TP06 1   /* odes/tentusscher_panfilov_2006/tentusscher_panfilov_2006_epi_cell.ode */
#define MONO_MODEL_TORORD 2 /* odes/torord/ToRORd_dynCl_endo.ode */
#define MONO_MODEL_SIMPLE 3 /* v' = -omega s, s' = omega v: the linear test model of tests/test_monodomain_solver.py:25-30 */

#define MONO_SCHEME_FORWARD_EULER 0 /* y + dt f */
#define MONO_SCHEME_GRL1 1          /* generalized Rush-Larsen, first order (gotranx scheme) */

/* preconditioners / norms / initial guess of the device CG (maps petsc_options, base_model.py:136-157) */
#define MONO_PC_NONE 0
#define MONO_PC_JACOBI 1
#define MONO_PC_CHEBYSHEV 2 /* polynomial: k Chebyshev steps for D^-1 A on [b/kappa, b], b = Gershgorin bound (pipecg only) */
#define MONO_NORM_PRECONDITIONED 0 /* PETSc default for KSPCG: ||M^-1 r||_2 */
#define MONO_NORM_UNPRECONDITIONED 1
#define MONO_NORM_NATURAL 2        /* sqrt(r . M^-1 r) */
#define MONO_KSP_CG 0     /* PETSc KSPCG: two global reductions per iteration */
#define MONO_KSP_PIPECG 1 /* PETSc KSPPIPECG (Ghysels-Vanroose pipelined CG): same iterates in exact arithmetic,
                             one global synchronisation per iteration */
#define MONO_X0_ZERO 0             /* PETSc default (KSPSetInitialGuessNonzero false) */
#define MONO_X0_PREVIOUS 1         /* start from v_ ; fewer iterations, same fixed point */

/* KSP-style converged reasons reported by mono_ksp_info (PETSc sign convention, telemetry.py:67-76) */
#define MONO_KSP_CONVERGED_RTOL 2
#define MONO_KSP_CONVERGED_ATOL 3
#define MONO_KSP_DIVERGED_ITS (-3)
#define MONO_KSP_DIVERGED_NAN (-9)

typedef struct mono_ctx mono_ctx;

/* ---- life cycle ------------------------------------------------------------------------------ */
int mono_abi_version(void);
const char *mono_last_error(const mono_ctx *ctx);
/* device: CUDA ordinal.  Fails (MONO_E_CUDA) when no CUDA device is usable. */
int mono_ctx_create(int device, mono_ctx **out);
int mono_ctx_destroy(mono_ctx *ctx);
/* block until all queued work of this context has finished */
int mono_sync(mono_ctx *ctx);
/* write *n_sm, *cc_major, *cc_minor, *mem_bytes of the context's device */
int mono_device_info(mono_ctx *ctx, int *n_sm, int *cc_major, int *cc_minor, int64_t *mem_bytes);

/* Page-locked host memory for the host mirrors of device vectors (pde.state.x.array & co): transfers
 * from/to it run at full PCIe rate.  Needs a CUDA device (MONO_E_CUDA otherwise).
 * The vector setters (mono_set_v, mono_set_v_prev, mono_set_v_ode, mono_ode_set_state_row) return after the
 * copy when the source is pageable; from page-locked memory the copy is ASYNCHRONOUS on the context's stream:
 * do not modify the buffer before the next synchronising call (mono_sync or any get/info call). */
int mono_host_alloc(int64_t nbytes, void **out);
int mono_host_free(void *ptr);

/* ---- multi-GPU (one rank per GPU; replaces the MPI communicator inside PETSc/dolfinx,
 *      base_model.py:203-206,236,242).  id is the 128-byte NCCL unique id made by rank 0 and
 *      distributed by the caller (bench.py uses torch.distributed).  NCCL is loaded with dlopen at this
 *      point: a libnccl.so.2 already in the process (torch's), else $MONO_NCCL_LIB, else the default path. */
int mono_comm_unique_id(void *id_out128);
int mono_comm_init(mono_ctx *ctx, int nranks, int rank, const void *id128);
/* Halo pattern of this rank's dofs (dolfinx IndexMap: owned dofs first, then ghosts grouped by
 * owner).  For neighbour k: send_idx[send_ptr[k]:send_ptr[k+1]] are OWNED local indices whose values
 * the neighbour needs; ghosts [recv_ptr[k], recv_ptr[k+1]) (offsets into the ghost block) come from it. */
int mono_set_halo(mono_ctx *ctx, int n_nbr, const int32_t *nbr_ranks, const int32_t *send_ptr,
                  const int32_t *send_idx, const int32_t *recv_ptr);
/* With more than one rank mono_set_halo is COLLECTIVE (every rank calls it, after mono_pde_set_matrices and
 * mono_comm_init): the ranks exchange CUDA-IPC handles of their exchange buffers so that the persistent PDE
 * kernel can store boundary values and partial dot products straight into the neighbours' memory over NVLink.
 * From then on every stepping call (mono_pde_step, mono_split_step, mono_split_solve) must be made by all
 * ranks in the same order, like the MPI collectives inside the reference's KSP solve.
 * mono_halo_refresh_nccl: ghost refresh of the PDE solution through ncclSend/ncclRecv - the plain-library
 * exchange the tests compare the in-kernel one with (state.x.scatter_forward(), base_model.py:242). */
int mono_halo_refresh_nccl(mono_ctx *ctx);

/* ---- ODE stage: DolfinODESolver / ODESystemSolver (odesolver.py:46-79,135-225) ----------------- */
/* num_points = owned + ghost dofs (odesolver.py:189-190); v_index = row of the membrane potential. */
int mono_ode_create(mono_ctx *ctx, int model_id, int scheme_id, int64_t num_points, int v_index);
/* states: (num_states, num_points) row-major with row stride ld (SoA, odesolver.py:149-153) */
int mono_ode_set_states(mono_ctx *ctx, const double *states, int64_t ld);
int mono_ode_get_states(mono_ctx *ctx, double *states, int64_t ld);
int mono_ode_set_state_row(mono_ctx *ctx, int row, const double *values);
int mono_ode_get_state_row(mono_ctx *ctx, int row, double *values);
/* parameters: per_node == 0: (num_params,) shared by all nodes, `derived` = the model's
 * parameter-only intermediates (n_derived values from the generated host module);
 * per_node == 1: (num_params, num_points) row-major with row stride ld, derived ignored.
 * (both shapes occur in the reference: demos/pace_train.py:133-167)                              */
int mono_ode_set_params(mono_ctx *ctx, const double *params, int num_params, int per_node, int64_t ld,
                        const double *derived, int n_derived);
/* Per-region parameter sets (DolfinMultiODESolver, odesolver.py:228-354, for ONE cell model): params (n_regions,
 * num_params) and derived (n_regions, n_derived) row-major, region_of_node[num_points] in [0, n_regions).  Pass
 * region_of_node == NULL to update the values only (a region's parameters were mutated in place).               */
int mono_ode_set_region_params(mono_ctx *ctx, int n_regions, const double *params, int num_params, const double *derived,
                               int n_derived, const int32_t *region_of_node);
/* states[:] = fun(states, t0, parameters, dt)          (odesolver.py:67-79) */
int mono_ode_step(mono_ctx *ctx, double t0, double dt);
/* v_ode <- states[v_index]   (DolfinODESolver.to_dolfin, odesolver.py:164-166) */
int mono_ode_to_dolfin(mono_ctx *ctx);
/* states[v_index] <- v_ode   (DolfinODESolver.from_dolfin, odesolver.py:168-170) */
int mono_ode_from_dolfin(mono_ctx *ctx);
/* v_pde <- v_ode / v_ode <- v_pde for identical P1 spaces (utils.py:52-54 via odesolver.py:101-115) */
int mono_ode_to_pde(mono_ctx *ctx);
int mono_pde_to_ode(mono_ctx *ctx);
int mono_get_v_ode(mono_ctx *ctx, double *v);
int mono_set_v_ode(mono_ctx *ctx, const double *v);

/* ---- PDE stage: MonodomainModel / BaseModel (monodomain_model.py:27-98, base_model.py:73-297) -- */
/* Constant P1 matrices of this rank in CSR over LOCAL indices: n_owned rows, n_owned+n_ghost columns,
 * identical sparsity for both.  mass_ij = int phi_i phi_j, stiff_ij = int (M grad phi_j).grad phi_i
 * (the two pieces of the form in monodomain_model.py:83-96).                                       */
int mono_pde_set_matrices(mono_ctx *ctx, int64_t n_owned, int64_t n_ghost, const int64_t *indptr,
                          const int32_t *indices, const double *mass, const double *stiff);
/* C_m (monodomain_model.py:38), theta of the theta-rule (base_model.py:159), CG controls
 * (petsc_options ksp_rtol / ksp_atol / ksp_max_it / pc_type / ksp_norm_type).                      */
int mono_pde_config(mono_ctx *ctx, double C_m, double theta, double rtol, double atol, int max_it,
                    int pc_type, int norm_type, int x0_mode);
/* MONO_PC_CHEBYSHEV: number of Chebyshev steps (1..4, default 3; 1 = Jacobi) and kappa (default 4).  Any kappa keeps
 * the preconditioner positive definite; it only tunes where the polynomial is accurate.  With several ranks call it
 * (and mono_pde_config) BEFORE mono_set_halo: the number of exchange buffers depends on it. */
int mono_pde_set_chebyshev(mono_ctx *ctx, int steps, double kappa);
/* Krylov driver (petsc_options ksp_type "cg" | "pipecg"); default MONO_KSP_CG. */
int mono_pde_set_ksp_type(mono_ctx *ctx, int ksp_type);
/* (Re)build A = C_m*Mass + dt*theta*K and B = C_m*Mass - dt*(1-theta)*K; called by mono_pde_step
 * itself when |dt - current| >= 1e-12 (base_model.py:225-230, _update_matrices :188-194).          */
int mono_pde_set_dt(mono_ctx *ctx, double dt);
/* Stimulus k: load vector s_k (sparse: idx/val over local OWNED dofs, = int phi_i dz(marker)),
 * amplitude, and the closed time window [t_start, t_end] tested at time t0 + theta*dt
 * (stimulation.py:264-272, base_model.py:216-223).  Returns the id (>= 0).                         */
int mono_stim_add(mono_ctx *ctx, int64_t nnz, const int32_t *idx, const double *val, double t_start,
                  double t_end, double amplitude);
int mono_stim_set_amplitude(mono_ctx *ctx, int stim_id, double amplitude); /* Stimulus.assign, stimulation.py:23-24 */
int mono_stim_set_window(mono_ctx *ctx, int stim_id, double t_start, double t_end);
/* one theta-rule step: b = B v_ + dt*sum_k a_k(t) s_k ; solve A v = b ; ghost refresh  (base_model.py:208-245) */
int mono_pde_step(mono_ctx *ctx, double t0, double t1);
/* v_ <- v   (MonodomainModel.assign_previous, monodomain_model.py:59-60) */
int mono_pde_assign_previous(mono_ctx *ctx);
int mono_get_v(mono_ctx *ctx, double *v);            /* pde.state.x.array (owned+ghost) */
int mono_set_v(mono_ctx *ctx, const double *v);
int mono_get_v_prev(mono_ctx *ctx, double *v);       /* pde.v_.x.array */
int mono_set_v_prev(mono_ctx *ctx, const double *v);
/* iterations, residual norm (in the configured norm) and reason of the last solve
 * (ksp.getIterationNumber / getResidualNorm / getConvergedReason, telemetry.py:67-76); synchronises. */
int mono_ksp_info(mono_ctx *ctx, int *iterations, double *residual_norm, int *reason);
/* sum of iterations over all solves since the context was created (no per-step sync needed) */
int mono_ksp_total_iterations(mono_ctx *ctx, int64_t *total, int64_t *solves);

/* EXPERIMENTAL - stencil dictionary of the matrices (set MONO_PDE_DICT=1 in the environment before
 * mono_pde_set_matrices): what it found (patterns kept, fraction of owned rows they cover) and whether the solver mode
 * chosen at the last step uses it (streaming KSPCG only).  Results are bit-identical with and without it. */
int mono_pde_dictionary_info(mono_ctx *ctx, int *n_patterns, double *rows_covered, int *active);

/* ---- the fused path: MonodomainSplittingSolver.step (monodomain_solver.py:53-116) -------------- */
/* One operator-split step on the device: ODE(theta_split*dt) -> PDE(dt) [-> ODE((1-theta_split)*dt)],
 * with the v_ode/v_pde/v_ hand-offs of :72-97 fused away.  Post-condition (as after the reference's
 * step): states[v_index] == v_ode == v == v_ .                                                    */
int mono_split_step(mono_ctx *ctx, double t0, double t1, double theta_split);
/* `nsteps` consecutive steps of size dt starting at t0 without returning to the host
 * (MonodomainSplittingSolver.solve, monodomain_solver.py:39-51).                                   */
int mono_split_solve(mono_ctx *ctx, double t0, double dt, int64_t nsteps, double theta_split);

/* ---- observers (demos/niederer_benchmark.py:281-289) --------------- */
/* Point probe = sum_j w_j v[node_j] (P1 interpolation inside one cell; nodes are local indices). */
int mono_probe_add(mono_ctx *ctx, int n_nodes, const int32_t *nodes, const double *weights);
int mono_probe_values(mono_ctx *ctx, double *values); /* current value of every probe; synchronises */
/* activation time of probe p = start time t0 of the first split step after which probe > threshold
 * (niederer_benchmark.py:284-287); -1 while not activated.                                         */
int mono_probe_activation(mono_ctx *ctx, double threshold);
int mono_probe_activation_times(mono_ctx *ctx, double *times);
/* Whole-field observers evaluated on the device after every split step, in the same launch as the probes, so that a
 * run never has to copy V to the host to follow it (what the demos do with state.x.array every step:
 * demos/niederer_benchmark.py:271-287 logs v.max()/v.min() and compares point values with 0, README.md:197-201):
 *   activation map: for every OWNED node the start time t0 of the first split step after which v > threshold, -1 before;
 *   min/max:        extrema of v over the owned nodes after the last split step.                                  */
int mono_observe_config(mono_ctx *ctx, int activation_map, double threshold, int minmax);
int mono_activation_map(mono_ctx *ctx, double *times_owned);            /* n_owned values; synchronises */
int mono_v_minmax(mono_ctx *ctx, double *vmin, double *vmax);           /* of the last split step; synchronises */
/* Strided snapshot: out[k] = v[offset + k*stride], k < count (gathered on the device, one small copy: frames for
 * plotting / coarse output without moving the whole field).                                                   */
int mono_get_v_strided(mono_ctx *ctx, int64_t offset, int64_t stride, int64_t count, double *out);

/* ---- host-side set-up helper (no GPU involved) -------------------------------------------------- */
/* P1 simplex assembly of the OWNED rows (0..n_owned-1) of the mass matrix  v*w*dx  and the stiffness matrix
 * inner(M grad v, grad w)*dx  (the two forms of monodomain_model.py:96-118 without their constant factors) over the
 * local columns, both on one sparsity with sorted columns: what dolfinx.fem.assemble_matrix provides in the reference
 * and mono_pde_set_matrices takes here.  For hosts that bring a mesh as plain arrays instead of a dolfinx mesh.
 *   tdim 1..3; cells: n_cells x (tdim+1) local vertex ids, row-major; x: n_local rows of x_ld >= tdim doubles;
 *   m_kind 0: scalar M[0], 1: constant tensor M[tdim*tdim], 2: one tensor per cell M[n_cells*tdim*tdim] (row-major).
 * Two calls: with indices == NULL only indptr[n_owned+1] is written (indptr[n_owned] = nnz); the second call, with
 * that indptr, fills indices/mass/stiff.  Multi-threaded (MONO_HOST_THREADS, default min(cores, 16)), deterministic.
 * Errors are reported through mono_last_error(NULL). */
int mono_fem_assemble_p1(int tdim, int64_t n_local, int64_t n_owned, int64_t n_cells, const int64_t *cells,
                         const double *x, int x_ld, int m_kind, const double *M, int64_t *indptr, int32_t *indices,
                         double *mass, double *stiff);

/* Stencil dictionary of a (mass, stiffness) CSR pair (host threads, no GPU): owned rows are classified by
 * (column offsets relative to the row, mass values, stiffness values), compared bit for bit.  The max_patterns (1..255)
 * most frequent stencils are numbered 0.. by frequency (ties: first row); pattern_of_row[r] is that number or 255 for a
 * row outside the dictionary; representative_row / rows_per_pattern (max_patterns entries each) describe the kept ones.
 * A structured box (the Niederer slab) has 27 stencils; a mapped or unstructured mesh has as many as rows.  Used to decide
 * whether the matrix stream of the CG iteration is compressible. */
int mono_csr_row_patterns(int64_t n_owned, const int64_t *indptr, const int32_t *indices, const double *mass,
                          const double *stiff, int max_patterns, uint8_t *pattern_of_row, int32_t *n_patterns,
                          int64_t *representative_row, int64_t *rows_per_pattern);

/* ---- measurement helpers (bench.py) -------------------------------------------------------------- */
/* CUDA-event stopwatch on the context's stream: start/stop record events, elapsed synchronises. */
int mono_timer_start(mono_ctx *ctx, int slot);
int mono_timer_stop(mono_ctx *ctx, int slot);
int mono_timer_elapsed_ms(mono_ctx *ctx, int slot, float *ms);
/* numbered CUDA-event marks on the context's stream (the pool grows on demand; bench.py brackets every
 * timed step with a pair so the L2 flush between steps stays outside the measurement).
 * mono_event_elapsed_ms synchronises on mark idx1. */
int mono_event_record(mono_ctx *ctx, int idx);
int mono_event_elapsed_ms(mono_ctx *ctx, int idx0, int idx1, float *ms);
/* overwrite a scratch buffer larger than L2 (evicts the working set between timed steps) */
int mono_l2_flush(mono_ctx *ctx);
/* per-stage device time accumulated since the last reset, in ms: [0]=ode [1]=pde(rhs+cg+halo) */
int mono_stage_times_ms(mono_ctx *ctx, double *ms2, int64_t *steps, int reset);
int mono_stage_timing(mono_ctx *ctx, int enable);
/* fp64 FMA micro-benchmark (the denominator of the ODE-stage roofline): achieved DFMA TFLOP/s */
int mono_bench_dfma(mono_ctx *ctx, double *tflops);
/* cost of one in-kernel grid-wide synchronisation + reduction of the persistent PDE kernel, in us
 * (n back-to-back synchronisations in one cooperative launch; needs the PDE matrices to be set) */
int mono_bench_grid_sync(mono_ctx *ctx, int n, float *us_per_sync);
/* measurement: with enable != 0 the persistent PDE kernel records %globaltimer (ns) of CTA 0 at its phase
 * boundaries; stamps64[0] = count, stamps64[1..] = the stamps of the last launch (synchronises). */
int mono_debug_timeline(mono_ctx *ctx, int enable, uint64_t *stamps64);
/* number of kernel launches issued by this context so far */
int mono_launch_count(mono_ctx *ctx, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* MONO_ABI_H */
