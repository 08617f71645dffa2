// Internal state behind the opaque mono_ctx of include/mono_abi.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mono_abi.h"

struct ncclComm;

// one stimulus as the device sees it (array of these lives in device memory)
struct StimDev {
  int64_t nnz;
  const int32_t* idx;
  const double* val;
  double t_start, t_end, amp;
};

// result block written by the PDE kernel (device memory, read back on demand)
struct KspResult {
  int iterations;
  int reason;
  double rnorm;
  long long total_iterations;
  long long solves;
  int error;  // non-zero: kernel-side failure (e.g. barrier timeout)
};

// one record of the in-kernel grid synchronisation / reduction (pde_kernels.cu grid_allreduce)
struct alignas(16) SyncRec {
  double val;
  unsigned long long gen;
};

constexpr int kMaxRanks = 16;  // ranks of one multi-GPU run (one node)
constexpr int kMaxCheb = 4;    // Chebyshev steps of the polynomial preconditioner (1 = Jacobi)
constexpr int kMaxTb = 3 * kMaxCheb;  // tagged exchange buffers (pde_kernels.cu)

struct ProbeDev {
  int n;           // nodes in this probe (<= 4)
  int32_t node[4];
  double w[4];
};

struct mono_ctx {
  int device = -1;
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;

  // ---- comm -------------------------------------------------------------------------------
  int nranks = 1, rank = 0;
  ncclComm* comm = nullptr;
  int n_nbr = 0;
  std::vector<int32_t> nbr_ranks, send_ptr, recv_ptr;
  int32_t* send_idx_dev = nullptr;  // concatenated owned indices to pack
  double* send_buf = nullptr;       // packed values, one segment per neighbour
  int64_t n_send = 0;
  double* red_buf = nullptr;        // scalars for all-reduce (device)

  // ---- ODE ----------------------------------------------------------------------------------
  bool has_ode = false;
  int model_id = -1, scheme_id = -1, ns = 0, np = 0, nd = 0, v_index = 0;
  int64_t npts = 0, ld = 0;
  double* states = nullptr;  // ns x ld (row = state, SoA)
  double* v_ode = nullptr;   // npts
  bool per_node = false;
  double* params_dev = nullptr;  // np x ld when per_node
  std::vector<double> params_host;  // np shared parameters followed by nd derived constants
  bool have_params = false;
  double* region_table = nullptr;      // [n_regions][np + max(nd,1)]: shared parameters + derived constants per region
  int32_t* region_of_node = nullptr;   // npts (padded to ld); non-null selects the per-region kernel
  int n_regions = 0;

  // ---- PDE ----------------------------------------------------------------------------------
  bool has_pde = false;
  int64_t n_owned = 0, n_ghost = 0, n_local = 0;
  int64_t n_slices = 0, sell_nnz = 0;   // SELL-32 storage
  int64_t* slice_ptr = nullptr;          // n_slices+1 element offsets
  int32_t* cols = nullptr;               // sell_nnz
  double *mass = nullptr, *stiff = nullptr, *A = nullptr, *B = nullptr;  // sell_nnz each
  double* dinv = nullptr;                // 1/diag(A) (or 1 for PC none), n_owned
  double *x = nullptr, *v_prev = nullptr;  // solution and previous solution, n_local each
  double* work[10] = {};                   // thread-private CG vectors in global memory (streaming mode only, lazy)
  // everything a peer rank writes into lives in ONE allocation (exported with CUDA IPC): see pde_setup_launch_config
  SyncRec* exch = nullptr;
  int64_t exch_recs = 0, exch_nl = 0, exch_off_xg = 0, exch_off_xrecs = 0;  // sizes / offsets in records
  int exch_nb = 2;                         // tagged exchange buffers of exch_nl records each, at the start of `exch`
  int cheb_k = 3;                          // Chebyshev steps when pc_type == MONO_PC_CHEBYSHEV
  double cheb_kappa = 4.0;                 // interval [b/kappa, b], b = Gershgorin bound of D^-1 A
  double cheb_inv_theta = 1.0, cheb_c1[kMaxCheb] = {}, cheb_c2[kMaxCheb] = {}, gershgorin = 0.0;
  SyncRec* xg = nullptr;                   // landing zone of the neighbours' final x values (n_ghost)
  SyncRec* xrecs = nullptr;                // cross-rank reduction records [2][4][kMaxRanks]
  unsigned long long* gen_state = nullptr; // device-resident generation counter of the persistent kernels
  double spin_timeout_ms = 4000.0;         // a poll that waits longer flags the launch as failed
  // peers (filled by mono_set_halo when nranks > 1)
  bool peers_ready = false;
  void* peer_base[kMaxRanks] = {};         // cudaIpcOpenMemHandle mappings of the other ranks' `exch`
  SyncRec* peer_xrecs[kMaxRanks] = {};
  int32_t* send_of_row_dev = nullptr;      // n_owned: first send entry of a row, -1 for interior rows
  void* send_ents_dev = nullptr;           // SendEnt[n_send] (pde_kernels.cu)
  uint8_t* slice_send_dev = nullptr;       // n_slices: 1 when a row of the slice has a send list
  int ksp_type = MONO_KSP_CG;
  double C_m = 1.0, theta = 0.5, rtol = 1e-5, atol = 1e-50;
  int max_it = 10000, pc_type = MONO_PC_JACOBI, norm_type = MONO_NORM_PRECONDITIONED, x0_mode = MONO_X0_ZERO;
  double cur_dt = -1.0;
  bool have_dt = false;
  std::vector<StimDev> stims_host;
  std::vector<void*> stim_allocs;
  double* stim_vec = nullptr;                    // dense sum_k a_k(t) s_k of the currently active stimuli
  std::vector<std::pair<int, double>> stim_sig;  // (stimulus id, amplitude) pairs stim_vec was built from

  // grid synchronisation records of the persistent PDE kernel: 2 parities x 4 scalars x blocks
  SyncRec* recs = nullptr;
  int pde_blocks = 0, pde_workers = 0, pde_threads = 0, rows_per_thread = 1, max_width = 0;
  unsigned long long* timeline_dev = nullptr;  // measurement: phase time stamps of the last PDE kernel (64 slots)
  bool mode_dirty = true;           // resident / matsmem / staged must be re-derived (mesh, ksp_type or pc_type changed)
  bool matsmem = false;             // ... and so do the A entries (one row per thread)
  bool resident = false;            // the CG vectors of a CTA's rows fit in shared memory
  bool staged = false;              // streaming mode: SELL slices reach the SpMV through TMA-staged shared memory
  bool stream_plain = false;        // streaming KSPCG runs pde_cg_stream_kernel (plain vectors + fenced barrier)
  size_t resident_smem = 0;
  // stencil dictionary of the matrices (default, MONO_PDE_DICT=0 disables; pde_build_sell): rows whose (column offsets, values)
  // repeat take their entries from a small table in shared memory instead of the SELL stream
  int n_pat = 0;                    // patterns in the dictionary (0: none)
  double dict_cover = 0.0;          // fraction of the owned rows the dictionary covers
  bool dict_active = false;         // the current mode runs the dictionary kernel (streaming KSPCG only)
  uint8_t* pat_dev = nullptr;       // n_owned: pattern of a row, 255 = not in the dictionary
  int32_t* dict_off_dev = nullptr;  // [n_pat][16] column offsets relative to the row
  int32_t* dict_w_dev = nullptr;    // [n_pat] entries per pattern
  int64_t* dict_src_dev = nullptr;  // [n_pat][16] SELL position of the representative row's entries (-1: padding)
  double *dict_A_dev = nullptr, *dict_B_dev = nullptr;  // [n_pat][16] gathered from A / B after every rebuild
  // shared-memory ring of the dictionary SpMV (pde_cg_stream_kernel, MODE 2): the column offsets of all stencils fall
  // into a few clusters (planes of a structured mesh); per cluster a ring of 2^ring_cap_log2 elements follows the rows
  bool ring_ok = false;             // the offsets cluster tightly enough for the rings to fit in shared memory
  bool ring_active = false;         // the current mode runs the ring kernel
  int ring_ncl = 0, ring_cap_log2 = 0, ring_depth = 2;
  int64_t ring_lo[4] = {}, ring_hi[4] = {};  // offset bounds of each cluster (inclusive)
  int32_t* dict_cl_dev = nullptr;   // [n_pat][16] cluster of each entry
  int64_t* dict_rep_dev = nullptr;  // [n_pat] a row that has the stencil
  double* dict_dinv_dev = nullptr;  // [n_pat] its Jacobi diagonal (gathered after every rebuild of A)
  uint8_t* slice_stim_dev = nullptr;  // n_slices (+ padding): 1 where the current stim_vec is non-zero on a row of the slice
  // rows outside the dictionary (next to a ghost layer, irregular spots) in a compact row-major side table
  int64_t n_nd = 0;
  std::vector<int32_t> nd_rows_host;
  int32_t* nd_rows_dev = nullptr;   // [n_nd] row
  int32_t* nd_w_dev = nullptr;      // [n_nd] entries
  int32_t* nd_cols_dev = nullptr;   // [n_nd][16]
  int64_t* nd_src_dev = nullptr;    // [n_nd][16] SELL position of each entry (-1: padding)
  double *nd_A_dev = nullptr, *nd_B_dev = nullptr;  // [n_nd][16]
  int64_t* nd_cta_ptr_dev = nullptr;  // [pde_workers + 1] side rows of each CTA's run of slices
  KspResult* ksp_dev = nullptr;
  KspResult* ksp_host = nullptr;  // pinned

  // fused-step bookkeeping: after mono_split_step the solution x is authoritative and
  // states[v_index], v_ode, v_prev are refreshed lazily (canonicalize()).
  bool fused_pending = false;

  // ---- observers ------------------------------------------------------------------------------
  std::vector<ProbeDev> probes_host;
  ProbeDev* probes_dev = nullptr;
  double* probe_vals_dev = nullptr;
  double* probe_act_dev = nullptr;
  bool probes_dirty = true;
  bool act_enabled = false;
  double act_threshold = 0.0;
  bool actmap_enabled = false, minmax_enabled = false;  // whole-field observers (mono_observe_config)
  double actmap_threshold = 0.0;
  double* actmap_dev = nullptr;          // n_owned: activation time of every owned node, -1 before
  unsigned long long* minmax_dev = nullptr;  // [2]: order-preserving integer images of min and max v of the last step
  double* snap_dev = nullptr;            // staging of mono_get_v_strided
  int64_t snap_cap = 0;

  // ---- measurement ---------------------------------------------------------------------------
  cudaEvent_t timers[8][2] = {};
  std::vector<cudaEvent_t> marks;    // mono_event_record pool
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  bool stage_timing = false;
  std::vector<cudaEvent_t> ev_pool;  // [ode_start, ode_stop, pde_start, pde_stop] per recorded step
  size_t ev_used = 0;
  std::vector<int> ev_tags;          // stage (0 = ode, 1 = pde) of each recorded event pair
  double stage_ms[2] = {0, 0};
  int64_t stage_steps = 0;
};

// ---- helpers shared by the translation units ---------------------------------------------------
int mono_fail(mono_ctx* c, int code, const std::string& msg);
// Both macros refuse a NULL context first, so every entry point that starts with one of them is safe to call with NULL.
#define MONO_CUDA(c, call)                                                                          \
  do {                                                                                              \
    if (!(c)) return mono_fail(nullptr, MONO_E_INVALID, "ctx is NULL");                             \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return mono_fail((c), MONO_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
  } while (0)
#define MONO_NEED_CTX(c)                                                \
  do {                                                                  \
    if (!(c)) return mono_fail(nullptr, MONO_E_INVALID, "ctx is NULL"); \
  } while (0)
#define MONO_CHECK(c, cond, msg)                                        \
  do {                                                                  \
    if (!(c)) return mono_fail(nullptr, MONO_E_INVALID, "ctx is NULL"); \
    if (!(cond)) return mono_fail((c), MONO_E_INVALID, (msg));          \
  } while (0)

// ode_kernels.cu
int ode_model_dims(int model_id, int* ns, int* np);
int ode_model_num_derived(int model_id, int scheme_id);
// launches the cell-model kernel over [0, n): V is read from v_in when non-null (else from the
// states row), new V additionally written to v_out1 / v_out2 when non-null.
int ode_launch(mono_ctx* c, double t, double dt, const double* v_in, double* v_out1, double* v_out2);

// pde_kernels.cu
int pde_build_sell(mono_ctx* c, const int64_t* indptr, const int32_t* indices, const double* mass, const double* stiff);
int pde_update_matrices(mono_ctx* c, double dt);
int pde_launch_step(mono_ctx* c, double t_eval, double dt);
int pde_setup_launch_config(mono_ctx* c);
int probes_launch(mono_ctx* c, double t0);  // probes + whole-field observers, one launch
int strided_pack_launch(mono_ctx* c, int64_t offset, int64_t stride, int64_t count, double* out_dev);
int pde_bench_sync(mono_ctx* c, int n, float* us_per_sync);

int pde_build_send_table(mono_ctx* c, const std::vector<int32_t>& row, const std::vector<std::vector<void*>>& dst_t,
                         const std::vector<void*>& dst_xg);

// halo.cu
int halo_refresh(mono_ctx* c, double* vec);  // owner -> ghost copy of an n_local vector through NCCL (no-op for 1 rank)
int halo_destroy(mono_ctx* c);
int halo_allreduce_max(mono_ctx* c, double* v);  // max over ranks through the bootstrap communicator (set-up only)
